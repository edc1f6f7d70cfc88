#!/usr/bin/env python
"""profiles/sass_summary.md: per-object counts of the SASS mnemonics that prove a Blackwell-native kernel
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UBLKCP, commit -> UTCBAR),
from `cuobjdump -sass` of the objects __graft_entry__.build() links into libciaosr_b200.so, plus the per-kernel
breakdown of the tcgen05 kernels.  Run after build():  python tools/sass_summary.py"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "LDGSTS"]


def main():
    objs = sorted(glob.glob(os.path.join(ROOT, "ciaosr_b200", "csrc", "build", "*.o")))
    out = ["# SASS evidence per object (`cuobjdump -sass`, sm_100a)", "",
           "Counts of instruction mnemonics; `UTCHMMA` = `tcgen05.mma kind::f16`, `UTCBAR` = `tcgen05.commit`, `LDTM`/`STTM` = "
           "`tcgen05.ld`/`st`, `UTMALDG` = `cp.async.bulk.tensor` (TMA tile load), `UBLKCP` = `cp.async.bulk`, `SYNCS` = mbarrier "
           "ops.  `HMMA` (legacy mma.sync) must be 0.", "",
           "| object | " + " | ".join(PAT) + " |", "|---|" + "---:|" * len(PAT)]
    per_kernel = []
    for o in objs:
        sass = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
        tot = collections.Counter()
        kern, kc = None, collections.Counter()
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                if kern and (kc["UTCHMMA"] or kc["UTMALDG"] or kc["LDTM"]):
                    per_kernel.append((os.path.basename(o), kern, dict(kc)))
                kern, kc = m.group(1), collections.Counter()
                continue
            for p in PAT:
                if re.search(r"\b" + p + r"[\.\s]", line):
                    tot[p] += 1
                    kc[p] += 1
        if kern and (kc["UTCHMMA"] or kc["UTMALDG"] or kc["LDTM"]):
            per_kernel.append((os.path.basename(o), kern, dict(kc)))
        out.append(f"| {os.path.basename(o)} | " + " | ".join(str(tot[p]) for p in PAT) + " |")
    out += ["", "## tcgen05 / TMA kernels", "", "| object | kernel (demangled) | " + " | ".join(PAT[:9]) + " |",
            "|---|---|" + "---:|" * 9]
    for o, k, c in per_kernel:
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)[:110]
        out.append(f"| {o} | `{name}` | " + " | ".join(str(c.get(p, 0)) for p in PAT[:9]) + " |")
    path = os.path.join(ROOT, "profiles", "sass_summary.md")
    open(path, "w").write("\n".join(out) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
