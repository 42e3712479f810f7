#!/bin/bash
# Per-kernel counts of the Blackwell-native SASS mnemonics in the shipped library (tcgen05.mma = UTCHMMA, tcgen05.ld/st =
# LDTM/STTM, TMA loads = UTMALDG, TMA stores = UTMASTG, tcgen05.commit = UTCBAR, legacy warp MMA = HMMA) -> profiles/sass_tcgen05.txt
LIB=${1:-make-an-audio-3_b200/csrc/libma3b200.so}
cuobjdump -sass "$LIB" | python3 -c '
import re, subprocess, sys
names, cur = {}, None
keys = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "MUFU", "instrs"]
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = names.setdefault(m.group(1), dict.fromkeys(keys, 0)); continue
    if cur is None or not re.match(r"\s+/\*[0-9a-f]+\*/", line):
        continue
    cur["instrs"] += 1
    for k in keys[:-1]:
        if re.search(r"\b" + k, line): cur[k] += 1
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
print("# %s: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)" % sys.argv[1])
print("%8s %5s %5s %8s %8s %7s %5s %5s %7s  kernel" % tuple(keys))
rows = sorted(zip(dem, names.values()), key=lambda r: (-r[1]["UTCHMMA"], -r[1]["HMMA"], r[0]))
for d, c in rows:
    if c["UTCHMMA"] + c["LDTM"] + c["UTMALDG"] + c["HMMA"] == 0: continue
    d = re.sub(r"\((ma3::\w+|float|int|__half|long long|__nv_bfloat16|unsigned).*$", "(...)", d)
    print("%8d %5d %5d %8d %8d %7d %5d %5d %7d  %s" % (*[c[k] for k in keys], d))
print("# kernels without tensor-core / TMA instructions (elementwise, norms, layout): %d" % sum(1 for c in names.values() if c["UTCHMMA"] + c["LDTM"] + c["UTMALDG"] + c["HMMA"] == 0))
' "$(basename $LIB)"
