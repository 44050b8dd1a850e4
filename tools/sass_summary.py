#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-only SASS instructions in libcic.so -> profiles/sass_summary.txt.

  python tools/sass_summary.py            (build.py runs it after every link)

UTCHMMA / UTCHMMA.2CTA = tcgen05.mma (one- / two-CTA), LDTM = tcgen05.ld (TMEM -> registers), UTMALDG = TMA tensor load
(cp.async.bulk.tensor), UTCBAR = tcgen05.commit -> mbarrier, UTMASTG = TMA store, SYNCS = mbarrier arrive / try_wait.
The mnemonics are the ones /opt/skills/guides/B200_PROFILING.md lists as proof of tcgen05 / TMA use.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "contextual-image-compression_b200", "libcic.so")
OUT = os.path.join(ROOT, "profiles", "sass_summary.txt")
MNEMONICS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCCP", "SYNCS", "HMMA", "FFMA", "DFMA"]


def main():
    cuobjdump = os.environ.get("CUOBJDUMP", "/usr/local/cuda/bin/cuobjdump")
    p = subprocess.run([cuobjdump, "-sass", LIB], capture_output=True, text=True)
    if p.returncode != 0:
        sys.exit(f"cuobjdump failed: {p.stderr[:400]}")
    counts = collections.OrderedDict()
    cur = None
    fn = re.compile(r"^\s*Function : (\S+)")
    ins = re.compile(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
    for line in p.stdout.splitlines():
        m = fn.match(line)
        if m:
            cur = counts.setdefault(m.group(1), collections.Counter())
            continue
        m = ins.match(line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                cur["UTCHMMA.2CTA"] += 1
            for k in MNEMONICS:
                if k != "UTCHMMA.2CTA" and (op == k or op.startswith(k + ".")):
                    cur[k] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    rows = []
    for (name, c), pretty in zip(counts.items(), demangle):
        short = re.sub(r"\(.*", "", pretty).replace("void ", "").replace("cic::", "")
        rows.append((short, c))
    rows.sort(key=lambda r: (-r[1]["UTCHMMA"], r[0]))
    cols = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTCBAR", "SYNCS", "FFMA", "DFMA", "_total"]
    with open(OUT, "w") as f:
        f.write("# cuobjdump -sass contextual-image-compression_b200/libcic.so: instruction counts per kernel (tools/sass_summary.py)\n")
        f.write("# UTCHMMA = tcgen05.mma (.2CTA: cta_group::2), LDTM = tcgen05.ld, UTMALDG = TMA load, UTCBAR = tcgen05.commit, SYNCS = mbarrier\n")
        f.write(f"{'kernel':96s} " + " ".join(f"{c:>12s}" for c in cols) + "\n")
        tot = collections.Counter()
        for short, c in rows:
            f.write(f"{short[:96]:96s} " + " ".join(f"{c[k]:12d}" for k in cols) + "\n")
            tot.update(c)
        f.write(f"{'TOTAL (' + str(len(rows)) + ' kernels)':96s} " + " ".join(f"{tot[k]:12d}" for k in cols) + "\n")
    print(f"wrote {OUT}: {len(rows)} kernels, UTCHMMA {tot['UTCHMMA']} ({tot['UTCHMMA.2CTA']} .2CTA), LDTM {tot['LDTM']}, UTMALDG {tot['UTMALDG']}")


if __name__ == "__main__":
    main()
