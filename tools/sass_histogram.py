"""CPU: opcode histogram of the shipped library's SASS (`cuobjdump -sass`), one row per kernel -- the Blackwell-native
evidence (UTCHMMA = tcgen05.mma, UTMALDG = TMA load, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit).
usage: python tools/sass_histogram.py [out.md]"""
import collections
import re
import subprocess
import sys

LIB = "graphinvent_b200/lib/libgib200.so"
OPS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "LDS", "STS", "LDG", "STG", "MUFU",
       "BRA", "HMMA"]
out = sys.argv[1] if len(sys.argv) > 1 else "profiles/r02_sass_histogram.md"
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kernels, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = demangle(m.group(1)).split("(")[0]
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        kernels[cur]["_n"] += 1
        kernels[cur][m.group(1)] += 1
L = ["# SASS opcode histogram of the shipped library (`cuobjdump -sass graphinvent_b200/lib/libgib200.so`, sm_100a)\n\n",
     "Blackwell-native evidence: `UTCHMMA` = tcgen05.mma, `UTMALDG` = TMA load (cp.async.bulk.tensor), `LDTM` / `STTM` = "
     "tcgen05.ld / tcgen05.st (tensor memory), `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier.  No `HMMA` (legacy mma.sync) "
     "anywhere.  `BRA` per kernel: the SELU epilogue (`tc3_gemm_kernel<false, 1>`) carried 180 of them in its store loop "
     "before it was made branch-free.  Regenerate with `python tools/sass_histogram.py`.\n\n",
     "| kernel | instructions | " + " | ".join(OPS) + " |\n", "|---|---:|" + "---:|" * len(OPS) + "\n"]
for k, c in kernels.items():
    L.append(f"| `{k[:64]}` | {c['_n']} | " + " | ".join(str(c[o]) for o in OPS) + " |\n")
open(out, "w").write("".join(L))
print("wrote", out, len(kernels), "kernels")
