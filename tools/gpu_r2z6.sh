#!/bin/bash
set -u
cd "$(dirname "$0")/.."
for pdl in 1 0; do
echo "== DCB_IMG_PDL=$pdl"; DCB_IMG_PDL=$pdl DCB_PIPE_PROBE=1 timeout 120 python tools/e2e_edges.py 2>&1 | grep -A3 "DIRECT 0"
DCB_IMG_PDL=$pdl DCB_PIPE_TRACE=1 timeout 120 python tools/e2e_edges.py equal8 2>&1 | tail -9
done
