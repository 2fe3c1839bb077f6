#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell path (B200_PROFILING.md): UTC*MMA = tcgen05.mma,
LDTM/STTM = tcgen05.ld/st, UTMALDG = cp.async.bulk.tensor (TMA), SYNCS = mbarrier, UTCBAR = tcgen05.commit, plus FFMA /
HMMA for contrast.  Reads the in-tree libdcs.so with cuobjdump.

    python tools/sass_counts.py > profiles/r2_sass_counts.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "deepconvsep_b200", "libdcs.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCATOM", "SYNCS", "LDGSTS", "HMMA", "FFMA", "FADD", "FMUL", "MUFU"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, counts, total = None, collections.OrderedDict(), collections.Counter()
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", name).replace("void ", "")
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for p in PAT:
                if op.startswith(p):
                    counts[cur][p] += 1
    print("SASS mnemonic counts per kernel of deepconvsep_b200/libdcs.so (cuobjdump -sass, sm_100a)")
    print("%-78s %7s " % ("kernel", "instr") + " ".join("%7s" % p for p in PAT))
    for k, c in counts.items():
        if c["_total"] == 0:
            continue
        print("%-78s %7d " % (k[:78], c["_total"]) + " ".join("%7d" % c[p] for p in PAT))
        total.update(c)
    print("%-78s %7d " % ("TOTAL", total["_total"]) + " ".join("%7d" % total[p] for p in PAT))


if __name__ == "__main__":
    main()
