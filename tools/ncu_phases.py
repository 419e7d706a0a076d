#!/usr/bin/env python3
"""Per-phase instruction / stall-sample shares of recon_tile_kernel from an ncu report.
   python tools/ncu_phases.py gpurun_out/recon_v3.ncu-rep [launches_in_report=2]
Phases are located by marker comments in recon_tile.cu; helper functions by their signature."""
import collections, sys
sys.path.insert(0, 'tools')
import ncu_lines
rep = sys.argv[1]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 2
L = ncu_lines.load(rep)
agg = collections.OrderedDict()
for f, ln, src, n, s, t in L:
    a = agg.setdefault((f, ln), [src, 0, 0, 0]); a[1] += n; a[2] += s; a[3] += t
tot = sum(a[1] for a in agg.values()); smp = sum(a[2] for a in agg.values())
print('warp inst per launch %.1fM' % (tot / nl / 1e6))
src = open('h263_rs_b200/csrc/recon_tile.cu').read().split('\n')
marks = []
for i, l in enumerate(src):
    ls = l.strip()
    if ls.startswith('// =====') or ls.startswith('// ---- ') or ls.startswith('__device__ __forceinline__') or ls.startswith('__global__'):
        marks.append((i + 1, ls[:70]))
marks.append((len(src) + 1, 'end'))
for i in range(len(marks) - 1):
    lo, hi = marks[i][0], marks[i + 1][0] - 1
    n = sum(a[1] for (ff, ln), a in agg.items() if ff == 'recon_tile.cu' and lo <= ln <= hi)
    s = sum(a[2] for (ff, ln), a in agg.items() if ff == 'recon_tile.cu' and lo <= ln <= hi)
    t = sum(a[3] for (ff, ln), a in agg.items() if ff == 'recon_tile.cu' and lo <= ln <= hi)
    if n:
        print("%4d-%-4d inst %5.1f%% (%6.1fM) smp %5.1f%% thr/inst %4.1f  %s" % (lo, hi, 100 * n / tot, n / nl / 1e6, 100 * s / smp, t / max(n, 1), marks[i][1]))
for f in sorted(set(ff for (ff, ln) in agg)):
    if f != 'recon_tile.cu':
        n = sum(a[1] for (ff, ln), a in agg.items() if ff == f); s = sum(a[2] for (ff, ln), a in agg.items() if ff == f)
        print("%-24s inst %5.1f%% (%6.1fM) smp %5.1f%%" % (f, 100 * n / tot, n / nl / 1e6, 100 * s / smp))
        if f == 'device_math.cuh':
            for (ff, ln), a in agg.items():
                if ff == f and a[1] / tot > 0.003:
                    print("      %4d %5.1f%%  %s" % (ln, 100 * a[1] / tot, a[0].strip()[:80]))
