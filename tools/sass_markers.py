#!/usr/bin/env python
"""SASS marker census of the shipped library: `cuobjdump -sass` + per-kernel counts of the mnemonics that prove tcgen05 / TMEM / TMA /
mbarrier use.    python tools/sass_markers.py > profiles/r02_sass_markers.txt"""
import collections
import os
import re
import subprocess

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "f16-model_b200", "lib", "libf16b200.so")
MARKS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UBLKPF", "SYNCS", "ELECT", "FFMA2", "MUFU.TANH", "MUFU.RCP", "MUFU.EX2", "MUFU.LG2", "MUFU.SIN",
         "MUFU.COS", "DFMA", "F2FP", "SHFL", "HFMA2", "HMUL2", "HADD2", "ATOMS", "ATOMG", "VOTE", "ACQBULK", "MEMBAR", "ERRBAR", "CCTL", "REDUX", "BAR.SYNC", "BAR.ARV"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()  # noqa: E731
    per, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = [0, collections.Counter()]
            order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if cur and m:
            per[cur][0] += 1
            text = m.group(1)
            for k in MARKS:
                if re.search(r"(^|\s)" + re.escape(k) + r"(\.|\s|$)", text):
                    per[cur][1][k] += 1
    total = collections.Counter()
    for _, c in per.values():
        total.update(c)
    print("# SASS marker census of f16-model_b200/lib/libf16b200.so (sm_100a, final round-2 build), `cuobjdump -sass` + grep, per kernel (tools/sass_markers.py).")
    print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA), UBLKPF = bulk L2 prefetch,")
    print("# SYNCS = mbarrier ops, ELECT = elect.sync, FFMA2 = packed fp32 FMA (sm_100), MUFU.TANH = tanh.approx, HFMA2 / HMUL2 / HADD2 = packed-half arithmetic")
    print("total: " + ", ".join(f"{k} {v}" for k, v in total.most_common()))
    print()
    for name in order:
        n, c = per[name]
        short = re.sub(r"\(.*", "", demangle(name).replace("(anonymous namespace)::", ""))
        print(f"{short}  [{n} instr]: " + ", ".join(f"{k} {v}" for k, v in c.most_common()))


if __name__ == "__main__":
    main()
