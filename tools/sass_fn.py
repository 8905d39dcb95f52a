"""Print the SASS of the functions of a .so whose (mangled) name matches all
the given substrings.  Usage: sass_fn.py lib.so substr [substr ...]"""
import subprocess
import sys

lib, pats = sys.argv[1], sys.argv[2:]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, keep = None, False
for line in out.splitlines():
    if "Function :" in line:
        name = line.split("Function :")[1].strip()
        keep = all(p in name for p in pats)
        if keep:
            print("==", name)
        continue
    if keep and "/*" in line and line.strip().startswith("/*") and ";" in line:
        print(line.split("*/", 1)[1].split(";")[0].strip())
