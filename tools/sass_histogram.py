#!/usr/bin/env python3
"""Instruction histogram of the hot loops from the built library's SASS (cuobjdump; no GPU needed).

  python tools/sass_histogram.py > profiles/r02_sass_histograms.txt

For loglik_kernel<2,2,1,128,3,0> the seven Gauss-Radau substeps are unrolled: the region between the first and the last pair of
MUFU.RSQ64H of the predictor-corrector loop is one iteration (7 substeps).  For var2_kernel<2,2,96,4,168> the substep loop is
rolled: producer part = from the mbarrier arrive (SYNCS.ARRIVE) to the try-wait (SYNCS.PHASECHK), second-order part = from
the try-wait to the next group barrier; both contain the 7-way corrector switch (static count; one case runs per substep).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rvel_mcmc_b200", "librvgpu.so")
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")


def sass(fun):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, LIB], capture_output=True, text=True).stdout
    ins = []
    for l in out.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", l)
        if m:
            ins.append((m.group(1), l))
    return ins


def hist(ins, title):
    c = collections.Counter(op for op, _ in ins)
    n = sum(c.values())
    f = sum(v for k, v in c.items() if k in FP64)
    print("%s: %d instructions, FP64 %d (%.1f %%)" % (title, n, f, 100.0 * f / max(n, 1)))
    print("   " + "  ".join("%s %d" % kv for kv in c.most_common(18)))


def main():
    ll = sass("_ZN2rv13loglik_kernelILi2ELi2ELi1ELi128ELi3ELi0EEEvNS_10LoglikArgsE")
    idx = [i for i, (op, l) in enumerate(ll) if op == "MUFU" and "RSQ64H" in l]
    # the unrolled predictor-corrector loop holds 14 consecutive rsqrt seeds (7 substeps x (star pair, planet pair)): take the
    # tightest window of 15 seeds -- from the first seed of substep 1 to the first seed of the next code region
    best = None
    for k in range(len(idx) - 14):
        span = idx[k + 14] - idx[k]
        if best is None or span < best[0]:
            best = (span, k)
    if best is not None:
        a, b = idx[best[1]], idx[best[1] + 14]
        hist(ll[a:b], "loglik_kernel<2,2,1,128,3,0>  seven Gauss-Radau substeps (one predictor-corrector iteration)")
        print("   per substep: %.1f instructions" % ((b - a) / 7.0))
    hist(ll, "loglik_kernel<2,2,1,128,3,0>  whole kernel (static)")
    v2 = sass("_ZN2rv11var2_kernelILi2ELi2ELi96ELi4ELi168EEEvNS_7VarArgsENS_10Var2LayoutE")
    arr = [i for i, (op, l) in enumerate(v2) if op == "SYNCS" and "ARRIVE" in l]
    chk = [i for i, (op, l) in enumerate(v2) if op == "SYNCS" and "PHASECHK" in l]
    if arr and chk:
        nb = next(i for i, (op, l) in enumerate(v2) if i > chk[0] and op == "BAR")
        pb = max(i for i, (op, l) in enumerate(v2) if i < arr[0] and op == "BAR")
        hist(v2[pb:arr[0]], "var2_kernel<2,2,96,4,168>  producer: predict + publish + star sums (per substep)")
        hist(v2[arr[0]:chk[0]], "var2_kernel  producer force + corrector switch, then the second-order lanes' predictor (static)")
        hist(v2[chk[0]:nb], "var2_kernel  second-order lanes: wait, force on the whole set, corrector switch (static)")
        loc = sum(1 for op, l in v2[chk[0]:nb] if op in ("LDL", "STL"))
        print("   local-memory (spill) instructions in the second-order substep: %d" % loc)
    hist(v2, "var2_kernel<2,2,96,4,168>  whole kernel (static)")


if __name__ == "__main__":
    main()
