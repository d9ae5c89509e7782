#!/usr/bin/env python
"""profiles/sass_summary.txt: per-kernel SASS mnemonic counts of the built library (cuobjdump -sass), the evidence
that the tensor-core / tensor-memory / bulk-copy / packed-fp32 instructions are really in the shipped code.

    python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "flid_b200", "libflid_b200.so")
COLS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "FFMA2", "FMUL2", "FADD2", "SYNCS"]
EXTRA = ["MUFU.EX2", "LDG.E.128", "STG.E.128", "SHFL", "ATOMG", "REDG", "LDGSTS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = {}
    try:
        import shutil
        if shutil.which("c++filt"):
            pass
    except Exception:
        pass
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            per[cur]["instrs"] += 1
            for c in COLS + EXTRA:
                if op == c or op.startswith(c + "."):
                    per[cur][c] += 1
    dem = subprocess.run(["c++filt"], input="\n".join(per.keys()), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    print("SASS summary of flid_b200/libflid_b200.so (sm_100a), `cuobjdump -sass` mnemonic counts per kernel (tools/sass_summary.py)")
    print("UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (tensor memory), UBLKCP = cp.async.bulk (TMA bulk copy),")
    print("SYNCS = mbarrier ops, FFMA2/FMUL2/FADD2 = packed fp32 pairs (sm_100).\n")
    print("whole library: " + ", ".join(f"{c} {total[c]}" for c in COLS + EXTRA) + "\n")
    print(f"{'kernel':112s} {'instrs':>6s} " + " ".join(f"{c:>7s}" for c in COLS))
    rows = sorted(zip(dem, per.values()), key=lambda kv: -kv[1]["instrs"])
    for name, c in rows:
        if c["instrs"] < 64 and not any(c[k] for k in COLS):
            continue
        print(f"{name[:112]:112s} {c['instrs']:6d} " + " ".join(f"{c[k]:7d}" for k in COLS))


if __name__ == "__main__":
    sys.exit(main())
