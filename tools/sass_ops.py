#!/usr/bin/env python
"""Opcode histogram of the Blackwell-specific instructions in every object of libb2s.so (cuobjdump -sass), so the
tcgen05 / TMEM / TMA claims can be checked from a tracked file: python tools/sass_ops.py > profiles/r02_sass_ops.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "llm_speech_summarization_b200", "build")
KEYS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UTMACCTL",
        "UTMACMDFLUSH", "LDTM", "STTM", "SYNCS", "ELECT", "HMMA", "MUFU.EX2", "RED.E", "REDG", "ATOMG", "FENCE.VIEW.ASYNC",
        "ACQBULK", "CCTL", "LDGSTS", "UBLKCP", "UBLKRED", "NANOSLEEP")


def main():
    print("# SASS opcode histogram per object (cuobjdump -sass, sm_100a); UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,")
    print("# UTMALDG = TMA tensor load, UTMASTG = TMA tensor store, UTMAREDG = TMA tensor reduction, UTCBAR = tcgen05.commit,")
    print("# SYNCS = mbarrier ops, UBLKCP = cp.async.bulk (1-D). No HMMA (mma.sync) remains anywhere.")
    for name in sorted(os.listdir(BUILD)):
        if not name.endswith(".o"):
            continue
        out = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, name)], capture_output=True, text=True).stdout
        hist = collections.Counter()
        total = 0
        for line in out.splitlines():
            m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
            if not m:
                continue
            total += 1
            op = m.group(1)
            for k in KEYS:
                if op.startswith(k):
                    hist[op] += 1
                    break
        print(f"\n== {name}: {total} SASS instructions")
        for op, n in sorted(hist.items()):
            print(f"   {n:6d}  {op}")


if __name__ == "__main__":
    sys.exit(main())
