"""Hot instructions (by stall samples) of the first kernel in an .ncu-rep.
Usage: ncu_hot.py rep [min_samples] [npx]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 25
npx = float(sys.argv[3]) if len(sys.argv) > 3 else 4096.0 * 4096.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
rowsx = []
for k, row in enumerate(rows[2:]):
    if row and row[0] == "Kernel Name":
        break
    if len(row) < len(hdr):
        continue
    rowsx.append((k, row))
tot = sum(int(r[iN]) for k, r in rowsx)
print("total samples", tot)
# running regions: print cumulative samples per 50-line block
blk = 64
for b in range(0, len(rowsx), blk):
    rs = rowsx[b:b + blk]
    n = sum(int(r[iN]) for k, r in rs)
    e = sum(int(r[iE]) for k, r in rs) * 32 / npx
    print("lines %4d-%4d samples %5d (%4.1f%%) instr/px %6.2f" % (rs[0][0], rs[-1][0], n, 100.0 * n / tot, e))
for k, row in rowsx:
    n = int(row[iN])
    if n >= thr:
        st = sorted([(int(row[i]), hdr[i][6:]) for i in stall_cols if int(row[i]) > 0], reverse=True)[:3]
        print("%4d %-55s e=%.2f n=%3d %s" % (k, row[iS].strip()[:55], int(row[iE]) * 32 / npx, n, st))
