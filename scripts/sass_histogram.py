#!/usr/bin/env python
"""SASS opcode histogram of the in-tree objects (cuobjdump -sass csrc/build/*.o): the mnemonics that prove the Blackwell
paths (UTC*MMA = tcgen05.mma, UTMALDG / UBLKCP = TMA / bulk copies, LDTM / STTM = tcgen05.ld / st), per .cu file.

    python scripts/sass_histogram.py > profiles/sass_opcodes.md        (runs in the build container, no GPU)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "seq_recommendations_b200", "csrc", "build")
KEYS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "FFMA", "MUFU.EX2",
        "REDG", "ATOM", "LDGSTS", "LDG", "STG", "LDS", "STS"]


def main():
    print("# SASS opcode histogram (sm_100a, `cuobjdump -sass`), per translation unit\n")
    print("`UTCHMMA` = tcgen05.mma kind::f16, `UTMALDG` = cp.async.bulk.tensor (TMA load), `UBLKCP` = cp.async.bulk "
          "(DSMEM bulk copy), `LDTM` / `STTM` = tcgen05.ld / tcgen05.st, `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier "
          "ops, `LDGSTS` = cp.async, `REDG` = red.global (vector reductions), `HMMA` = legacy mma.sync (none expected).\n")
    print("| object | kernels | " + " | ".join("`%s`" % k for k in KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    for f in sorted(os.listdir(BUILD)):
        if not f.endswith(".o"):
            continue
        out = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, f)], capture_output=True, text=True).stdout
        n_kernels = len(re.findall(r"^\s*Function :", out, flags=re.M))
        cnt = collections.Counter()
        for line in out.splitlines():
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            op = m.group(1)
            for k in KEYS:
                if op.startswith(k):
                    cnt[k] += 1
                    break
        print("| `%s` | %d | " % (f, n_kernels) + " | ".join(str(cnt[k]) for k in KEYS) + " |")


if __name__ == "__main__":
    main()
