"""Opcode evidence from the built library: per kernel, how many tensor-core / TMEM / bulk-copy instructions the SASS
holds (B200_PROFILING.md: UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP =
cp.async.bulk, UTMALDG/UTMASTG = tensor-map TMA, HMMA/IMMA = legacy mma.sync).

    python tools/sass_histogram.py > profiles/sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gdrf_b200", "libgdrf_b200.so")
WATCH = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "UTMALDG", "UTMASTG",
         "SYNCS", "HMMA", "IMMA", "DFMA", "FFMA", "MUFU", "F2FP", "F2F", "ATOMS", "RED", "ATOM", "STG", "LDG")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
            if m.group(1) == "UTCHMMA" and ".2CTA" in m.group(2):
                kernels[cur]["UTCHMMA.2CTA"] += 1
            kernels[cur]["_total"] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: instruction counts per kernel (static SASS)")
    tot = collections.Counter()
    for name, c in kernels.items():
        hits = {k: c[k] for k in WATCH + ("UTCHMMA.2CTA",) if c.get(k)}
        for k, v in hits.items():
            tot[k] += v
        print(f"{name[:110]:110s} total {c['_total']:6d}  " + "  ".join(f"{k} {v}" for k, v in hits.items()))
    print("\n# library totals: " + "  ".join(f"{k} {v}" for k, v in sorted(tot.items())))
    legacy = tot.get("HMMA", 0) + tot.get("IMMA", 0)
    print(f"# tcgen05 MMAs (UTCHMMA) {tot.get('UTCHMMA', 0)}, of which cta_group::2 {tot.get('UTCHMMA.2CTA', 0)}; "
          f"legacy mma.sync (HMMA/IMMA) {legacy}; tensor-map TMA (UTMALDG/UTMASTG) {tot.get('UTMALDG', 0) + tot.get('UTMASTG', 0)}; "
          f"bulk copies (UBLKCP) {tot.get('UBLKCP', 0)}")


if __name__ == "__main__":
    main()
