#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): tcgen05.mma ->
UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UTMAREDG/UBLKCP, tcgen05.commit -> UTCBAR, setmaxnreg ->
USETMAXREG, and the legacy HMMA path (must be 0).   python tools/sass_opcodes.py > profiles/sass_opcodes.txt"""
import _path  # noqa: F401  (repository root on sys.path)
import re
import subprocess
import sys

sys.path.insert(0, ".")
from aigv_assessor_b200 import _lib

OPS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UBLKPF", "USETMAXREG",
       "SYNCS", "MUFU.EX2", "FFMA2", "HMMA"]
out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
funcs = re.split(r"\n\s*Function : ", out)[1:]
print(f"# {_lib.LIB_PATH.split('/')[-1]}: SASS mnemonic counts per kernel (cuobjdump -sass), sm_100a")
print("%-86s %6s " % ("kernel", "instrs") + " ".join("%8s" % o for o in OPS))
tot = {o: 0 for o in OPS}
for f in funcs:
    name = f.split("\n")[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    dem = re.sub(r"\(anonymous namespace\)::", "", dem)
    dem = re.sub(r"\(CUtensorMap_st.*", "(...)", dem)[:84]
    n = len(re.findall(r"/\*[0-9a-f]{4,5}\*/\s+\S", f))
    counts = [len(re.findall(r"\b" + re.escape(o), f)) for o in OPS]
    for o, c in zip(OPS, counts):
        tot[o] += c
    print("%-86s %6d " % (dem, n) + " ".join("%8d" % c for c in counts))
print("%-86s %6s " % ("TOTAL", "") + " ".join("%8d" % tot[o] for o in OPS))
