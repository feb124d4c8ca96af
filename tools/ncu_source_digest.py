#!/usr/bin/env python3
"""Digest of the per-instruction (SASS) page of an ncu report captured with --set full --import-source on (read HERE, no GPU).

  python tools/ncu_source_digest.py gpurun_out/prof.ncu-rep [--min-share 0.004] [--top 12] > profiles/r02x_source_digest.txt

Prints the dynamic (executed) instruction mix, then splits the kernel into regions of equal execution count (straight-line
code that runs the same number of times: the predictor-corrector loop body, the per-attempt code before / after it, ...) with
each region's share of executed instructions, share of the stall samples (= share of time), FP64 fraction, static instruction
mix and stall-reason breakdown.  issue-rate estimate of a region = selected share x resident warps per scheduler.
"""
import argparse, collections, csv, io, re, subprocess

FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--min-share", type=float, default=0.004)
    ap.add_argument("--top", type=int, default=12)
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print("# " + (rows[h - 1][1] if h > 0 and len(rows[h - 1]) > 1 else ""))
    hdr, data = rows[h], [r for r in rows[h + 1:] if len(r) >= len(rows[h])]
    isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    st = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
    ins = []
    for r in data:
        m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", r[isrc])
        ins.append((m.group(1) if m else "?", int(r[iex]), int(r[ismp]), r))
    tot, tots = sum(i[1] for i in ins), sum(i[2] for i in ins)
    c = collections.Counter()
    for op, ex, smp, r in ins:
        c[op] += ex
    print("# %d SASS instructions, %.4g executed (warp level), %d stall samples" % (len(ins), tot, tots))
    print("dynamic mix: " + "  ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in c.most_common(18)))
    print("FP64 share of executed instructions: %.3f" % (sum(v for k, v in c.items() if k in FP64) / tot))
    segs, cur = [], None
    for k, (op, ex, smp, r) in enumerate(ins):
        if cur is None or not (0.8 * cur["ex"] <= ex <= 1.25 * cur["ex"]):
            cur = {"start": k, "ex": max(ex, 1), "n": 0, "sum": 0, "smp": 0, "ops": collections.Counter(), "st": collections.Counter()}
            segs.append(cur)
        cur["n"] += 1; cur["sum"] += ex; cur["smp"] += smp; cur["ops"][op] += 1
        for i in st:
            if r[i] and r[i] != "0":
                cur["st"][hdr[i][6:]] += int(r[i])
        cur["ex"] = cur["sum"] / cur["n"] if cur["sum"] else 1
    print("regions of equal execution count (share of executed instructions >= %g):" % a.min_share)
    for s in segs:
        if s["sum"] / tot < a.min_share:
            continue
        f = sum(v for k, v in s["ops"].items() if k in FP64)
        print("  @%5d  %4d instr x %.3e  instr share %.3f  time share %.3f  FP64 %.2f | %s" % (
            s["start"], s["n"], s["ex"], s["sum"] / tot, s["smp"] / tots, f / s["n"], " ".join("%s:%d" % kv for kv in s["ops"].most_common(8))))
        if s["smp"]:
            print("          stalls: " + "  ".join("%s %.2f" % (k, v / s["smp"]) for k, v in s["st"].most_common(6)))
    print("top instructions by stall samples:")
    for k in sorted(sorted(range(len(ins)), key=lambda k: -ins[k][2])[:a.top]):
        op, ex, smp, r = ins[k]
        rs = sorted(((int(r[i]), hdr[i][6:]) for i in st if r[i] and r[i] != "0"), reverse=True)[:3]
        print("  @%5d %7d  %-64s %s" % (k, smp, r[isrc].strip()[:64], " ".join("%s:%d" % (b, v) for v, b in rs)))


if __name__ == "__main__":
    main()
