#!/usr/bin/env python3
"""Summarise an ncu report per CUDA source line (needs -lineinfo + --import-source on).

    python tools/ncu_lines.py gpurun_out/recon_v1.ncu-rep [top_n] [--ranges a-b:label,...]

Prints the lines that execute the most warp instructions with their stall samples, and an
optional roll-up over line ranges (phases of the kernel).  Runs here, without a GPU.
"""
import csv
import subprocess
import sys


def load(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    out = []  # (file, line, source, inst, samples, thread_inst)
    cur_file, hdr = None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            i_inst = hdr.index("Instructions Executed")
            i_smp = hdr.index("# Samples")
            i_tin = hdr.index("Thread Instructions Executed")
            continue
        if hdr and len(r) == len(hdr) and r[0].isdigit():
            try:
                out.append((cur_file, int(r[0]), r[1], int(r[i_inst]), int(r[i_smp]), int(r[i_tin])))
            except ValueError:
                pass
    return out


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 40
    ranges = []
    for a in sys.argv[2:]:
        if a.startswith("--ranges="):
            for part in a[9:].split(","):
                rng, label = part.split(":")
                lo, hi = rng.split("-")
                ranges.append((int(lo), int(hi), label))
    L = load(rep)
    tot = sum(x[3] for x in L)
    smp = sum(x[4] for x in L)
    print("total warp instructions (all captured launches): %d, samples %d" % (tot, smp))
    byfile = {}
    for f, ln, src, n, s, t in L:
        a = byfile.setdefault(f, [0, 0])
        a[0] += n
        a[1] += s
    for f, (n, s) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print("  %-28s inst %5.1f%%  samples %5.1f%%" % (f, 100.0 * n / tot, 100.0 * s / max(smp, 1)))
    print("top lines:")
    for f, ln, src, n, s, t in sorted(L, key=lambda x: -x[3])[:top]:
        print("  %5.2f%% smp %5.2f%% thr/inst %4.1f %s:%d  %s" % (100.0 * n / tot, 100.0 * s / max(smp, 1), t / max(n, 1), f, ln,
                                                              src.strip()[:90]))
    if ranges:
        print("ranges (kernels.cu lines; inlined device_math.cuh lines are not attributed):")
        for lo, hi, label in ranges:
            n = sum(x[3] for x in L if x[0] == "kernels.cu" and lo <= x[1] <= hi)
            s = sum(x[4] for x in L if x[0] == "kernels.cu" and lo <= x[1] <= hi)
            print("  %-24s %4d-%-4d inst %5.1f%% samples %5.1f%%" % (label, lo, hi, 100.0 * n / tot, 100.0 * s / max(smp, 1)))


if __name__ == "__main__":
    main()
