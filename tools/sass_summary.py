"""Mnemonic counts per kernel from `cuobjdump -sass` of the built library (no GPU needed): the evidence that the
tensor-core kernels are tcgen05 / TMEM / TMA native and that no legacy mma.sync path exists.

    python tools/sass_summary.py > profiles/r02_sass_summary.md
"""

import os
import re
import subprocess
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "audiopure_b200", "libaudiopure_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACMDFLUSH",
         "SYNCS", "UCGABAR", "ACQBULK", "HMMA", "IMMA", "MUFU", "LDG", "STG", "LDS", "STS", "LDGSTS", "ATOMG", "RED", "SHFL",
         "BAR", "FFMA", "ELECT"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            kernels[name] = Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and name:
            kernels[name][m.group(1)] += 1
            kernels[name]["_total"] += 1
            if m.group(1) in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM", "MUFU"):
                kernels[name][m.group(1) + m.group(2)] += 1
    print("# SASS mnemonic counts per kernel (`cuobjdump -sass audiopure_b200/libaudiopure_b200.so`, sm_100a)\n")
    print("tcgen05.mma -> `UTCHMMA` (`.2CTA` = cta_group::2), tcgen05.ld -> `LDTM`, tcgen05.commit -> `UTCBAR`, "
          "TMA loads / stores -> `UTMALDG` / `UTMASTG`, mbarrier -> `SYNCS`, cluster barrier -> `UCGABAR`; "
          "`HMMA` / `IMMA` (legacy mma.sync) must not appear.\n")
    short = lambda n: re.sub(r"\(.*", "", n).replace("void ", "")
    cols = [w for w in WATCH if any(k[w] for k in kernels.values())]
    print("| kernel | instructions | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for n, c in kernels.items():
        print("| `%s` | %d | " % (short(n), c["_total"]) + " | ".join(str(c[w]) if c[w] else "" for w in cols) + " |")
    print("\n## Variants of the Blackwell-specific instructions\n")
    for n, c in kernels.items():
        det = sorted((k, v) for k, v in c.items() if "." in k)
        if det:
            print("* `%s`: " % short(n) + ", ".join("`%s` x%d" % kv for kv in det))
    legacy = sum(c["HMMA"] + c["IMMA"] for c in kernels.values())
    print("\nLegacy `HMMA` / `IMMA` instructions in the library: **%d**." % legacy)


if __name__ == "__main__":
    main()
