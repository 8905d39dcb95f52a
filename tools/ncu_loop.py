"""Every SASS line of the first kernel in an .ncu-rep that executed at least `frac` of the hottest
line's count, with its stall samples: the sampling loop, instruction by instruction.
Usage: ncu_loop.py rep [frac] [npx]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
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
    if len(row) >= len(hdr):
        rowsx.append((k, row))
emax = max(int(r[iE]) for k, r in rowsx)
tot = sum(int(r[iN]) for k, r in rowsx)
hot = [(k, r) for k, r in rowsx if int(r[iE]) >= frac * emax]
print("total samples %d; hot lines %d, their samples %d (%.1f %%), instr/px %.2f" % (
    tot, len(hot), sum(int(r[iN]) for k, r in hot), 100.0 * sum(int(r[iN]) for k, r in hot) / tot,
    sum(int(r[iE]) for k, r in hot) * 32 / npx))
for k, row in hot:
    n = int(row[iN])
    st = sorted([(int(row[i]), hdr[i][6:]) for i in stall_cols if int(row[i]) > 0], reverse=True)[:3]
    print("%4d %-58s n=%3d %s" % (k, row[iS].strip()[:58], n, " ".join("%s:%d" % (b, a) for a, b in st)))
