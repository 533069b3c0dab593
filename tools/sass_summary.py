"""SASS mnemonic counts of chosen kernels of an object file (compile-time evidence: TMA loads, mbarrier waits,
128-bit accesses):   python tools/sass_summary.py build/kernels_stencil.o 'stencil7_dot_tma.*TileCfgILi64ELi16ELi4E' ..."""
import re
import subprocess
import sys

OPS = ["UTMALDG", "SYNCS", "MEMBAR.SC.SYS", "MEMBAR.SC.GPU", "DFMA", "DMUL", "DADD", "LDG.E.128", "LDG.E.64", "STG.E.128",
       "STG.E.64", "LDS.128", "LDS.64", "STS", "BAR.SYNC", "ATOMG", "RED"]
sass = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", sass)[1:]
print("kernel | | " + " | ".join(OPS))
for pat in sys.argv[2:]:
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        if not re.search(pat, name):
            continue
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", dem).split("(")[0]
        body = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", f)
        counts = [sum(1 for b in body if b == op or b.startswith(op + ".") or (op.endswith("STS") and b.startswith("STS"))) for op in OPS]
        print(dem + " | | " + " | ".join(str(c) for c in counts))
