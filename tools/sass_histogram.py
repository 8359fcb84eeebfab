"""sass_histogram.py — per-kernel SASS opcode evidence for the Blackwell-native instructions.

    python tools/sass_histogram.py > profiles/r02_sass_histogram.md

Dumps `cuobjdump -sass` of the in-tree library and counts, per kernel, the mnemonics that prove tcgen05 / TMEM / TMA
(B200_PROFILING.md: tcgen05.mma → UTC*MMA, tcgen05.ld/st → LDTM/STTM, cp.async.bulk.tensor → UTMALDG/UTMASTG) next to
the legacy tensor path (HMMA = mma.sync) and MUFU."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "bridgelang_b200" / "libbridgelang_b200.so"
WANT = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "LDSM", "LDGSTS", "MUFU.EX2",
        "FMNMX3", "USETMAXREG"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WANT:
                if op == w or op.startswith(w + ".") or (w == "UTCHMMA" and op.startswith("UTCHMMA")):
                    kernels[cur][w] += 1
    names = demangle(list(kernels))
    print("# SASS opcode histogram of `bridgelang_b200/libbridgelang_b200.so` (sm_100a)\n")
    print("`cuobjdump -sass`, counted per kernel by `tools/sass_histogram.py`.  UTCHMMA = tcgen05.mma, LDTM/STTM = "
          "tcgen05.ld/st (TMEM), UTMALDG = cp.async.bulk.tensor (TMA load), UTCBAR = tcgen05.commit, SYNCS = mbarrier, "
          "HMMA/LDSM = mma.sync/ldmatrix (only the fallback attention kernels for non-tower shapes).\n")
    print("| kernel | instr | " + " | ".join(WANT) + " |")
    print("|---|---|" + "---|" * len(WANT))
    tot = collections.Counter()
    for k, c in kernels.items():
        name = names.get(k, k)
        name = re.sub(r"blb::\(anonymous namespace\)::|blb::", "", name)
        name = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", name)[:110]
        print(f"| `{name}` | {c['_total']} | " + " | ".join(str(c[w]) if c[w] else "" for w in WANT) + " |")
        tot.update(c)
    print(f"| **all {len(kernels)} kernels** | {tot['_total']} | " + " | ".join(str(tot[w]) for w in WANT) + " |")


if __name__ == "__main__":
    sys.exit(main())
