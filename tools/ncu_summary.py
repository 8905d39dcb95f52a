"""Summarise an .ncu-rep (run where ncu is installed, no GPU needed):
key raw metrics of the first captured kernel + executed-instruction mix and
stall reasons from the SASS source page.  Usage: ncu_summary.py rep [npx]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
npx = float(sys.argv[2]) if len(sys.argv) > 2 else 4096.0 * 4096.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum.pct", "sm__inst_executed_pipe_xu.sum.pct",
        "sm__inst_executed_pipe_alu.sum.pct", "sm__inst_executed_pipe_fma.sum.pct",
        "sm__inst_executed_pipe_lsu.sum.pct", "smsp__issue_active.avg.pct", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max", "lts__t_sectors_srcunit_tex_op_read.sum",
        "smsp__cycles_active.avg"]
r = rows[2]
print("kernel:", r[hdr.index("Kernel Name")])
for i, k in enumerate(hdr):
    if any(k.startswith(x) for x in KEYS) and "per_second" not in k:
        print("  %-75s %-12s %s" % (k, units[i], r[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ops, samp, stalls = collections.Counter(), collections.Counter(), collections.Counter()
total = 0
for row in rows[2:]:
    if row and row[0] == "Kernel Name":
        break
    if len(row) < len(hdr):
        continue
    m = re.match(r"(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", row[iS].strip())
    op = m.group(1) if m else "?"
    e, s = int(row[iE]), int(row[iN])
    ops[op] += e
    samp[op] += s
    total += e
    for i in stall_cols:
        stalls[hdr[i]] += int(row[i])
print("thread-instructions per pixel: %.1f" % (total * 32 / npx))
for op, c in ops.most_common(28):
    print("  %-10s %6.2f /px   samples %d" % (op, c * 32 / npx, samp[op]))
print("stalls:", stalls.most_common(10))
