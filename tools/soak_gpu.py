#!/usr/bin/env python3
"""Randomised GPU-vs-oracle soak: N seeded stream configurations (picture size, quantiser range, vector mode, 4MV, escapes,
zig-zag overflows, intra share, truncation, event density, flavour, deblocking), every picture compared bit for bit with the
oracle (planes + RGBA).  Run on a B200:  python tools/soak_gpu.py [n_cases] [seed]  -> one summary line per 10 cases.
The oracle is the checker here, as in tests/ (this is test tooling, not product code)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import oracle_decode_stream  # noqa: E402
from h263_rs_b200 import _lib, api, synth  # noqa: E402

SIZES = [(176, 144), (352, 288), (128, 96), (16, 16), (32, 48), (160, 120), (200, 100), (161, 99), (47, 31), (64, 64), (240, 176), (704, 576),
         (24, 40), (17, 33), (96, 80), (320, 240), (15, 15), (480, 360)]


def one_case(rng, k):
    w, h = SIZES[int(rng.integers(len(SIZES)))]
    big = w * h > 200000
    qlo = int(rng.integers(1, 32))
    kw = dict(
        mv_mode=int(rng.integers(3)), pct_fourmv=int(rng.integers(0, 40)), pct_dquant=int(rng.integers(0, 50)),
        pct_escape=int(rng.integers(0, 40)), permille_overflow=int(rng.integers(0, 40)), pct_intra=int(rng.integers(0, 60)),
        pct_uncoded=int(rng.integers(0, 60)), pct_cbp_inter=int(rng.integers(5, 100)), pct_cbp_intra=int(rng.integers(20, 101)),
        mean_events_x10=int(rng.integers(5, 260)), qp_min=qlo, qp_max=int(rng.integers(qlo, 32)),
        truncate_permille=int(rng.integers(0, 2)) * int(rng.integers(0, 400)), intra_period=int(rng.integers(0, 5)),
        deblock_flag=int(rng.integers(2)),
    )
    flavour = int(rng.integers(0, 5) == 0)
    if flavour:
        kw["flavour"] = 1
        if (w, h) not in ((176, 144), (352, 288), (128, 96), (704, 576)):
            w, h = 176, 144
    else:
        kw["version"] = int(rng.integers(2))
    n = 3 if big else int(rng.integers(3, 8))
    packets = synth.make_stream(w, h, n, int(rng.integers(1 << 30)), **kw)
    opt = 0 if flavour else 1
    deblock = bool(kw["deblock_flag"]) and rng.integers(2) == 1
    ref = oracle_decode_stream(packets, opt, deblock=deblock)
    st = api.H263State(opt, deblock=deblock)
    pics = 0
    for i, pk in enumerate(packets):
        if isinstance(ref[i], int):
            try:
                st.decode_next_picture(pk)
                return "case %d %dx%d picture %d: oracle fails with %d, product decodes" % (k, w, h, i, ref[i]), pics
            except _lib.H263Error:
                continue
        st.decode_next_picture(pk)
        y, cb, cr = st.get_last_picture().as_yuv()
        ok = np.array_equal(y, ref[i]["y"]) and np.array_equal(cb, ref[i]["cb"]) and np.array_equal(cr, ref[i]["cr"]) and \
            np.array_equal(st.get_last_rgba(), ref[i]["rgba"])
        if not ok:
            return "case %d %dx%d picture %d MISMATCH %r" % (k, w, h, i, kw), pics
        pics += 1
    return None, pics


def one_batch(rng, k):
    """n streams of one picture size decoded in lock step through h263cu_decode_step (tiles of four macroblocks straddle
    pictures whenever the macroblock count is not a multiple of four), every stream with its own content."""
    w, h = SIZES[int(rng.integers(len(SIZES)))]
    if w * h > 200000:
        w, h = 176, 144
    n, steps = int(rng.integers(2, 14)), int(rng.integers(2, 6))
    streams, refs = [], []
    for s in range(n):
        qlo = int(rng.integers(1, 32))
        kw = dict(mv_mode=int(rng.integers(3)), pct_fourmv=int(rng.integers(0, 40)), pct_escape=int(rng.integers(0, 30)),
                  pct_intra=int(rng.integers(0, 50)), pct_uncoded=int(rng.integers(0, 60)), pct_cbp_inter=int(rng.integers(5, 100)),
                  mean_events_x10=int(rng.integers(5, 200)), qp_min=qlo, qp_max=int(rng.integers(qlo, 32)),
                  permille_overflow=int(rng.integers(0, 30)), intra_period=int(rng.integers(0, 4)))
        streams.append(synth.make_stream(w, h, steps, int(rng.integers(1 << 30)), **kw))
        refs.append(oracle_decode_stream(streams[-1], 1))
    dec = api.BatchDecoder(n, w, h, threads=4)
    pics = 0
    for t in range(steps):
        errs = dec.decode_step([streams[s][t] for s in range(n)])
        dec.ctx.sync()
        for s in range(n):
            r = refs[s][t]
            if isinstance(r, int):
                if not errs[s]:
                    return "batch %d %dx%d stream %d step %d: oracle fails, product decodes" % (k, w, h, s, t), pics
                continue
            y, cb, cr = dec.ctx.read_yuv(s)
            if errs[s] or not (np.array_equal(y, r["y"]) and np.array_equal(cb, r["cb"]) and np.array_equal(cr, r["cr"]) and
                               np.array_equal(dec.ctx.read_rgba(s), r["rgba"])):
                return "batch %d %dx%d stream %d step %d MISMATCH (err %d)" % (k, w, h, s, t, int(errs[s])), pics
            pics += 1
    dec.ctx.close()
    return None, pics


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 2026
    rng = np.random.default_rng(seed)
    t0 = time.time()
    bad, pics = [], 0
    for k in range(n):
        err, p = one_batch(rng, k) if k % 4 == 3 else one_case(rng, k)
        pics += p
        if err:
            bad.append(err)
            print(err, flush=True)
        if (k + 1) % 10 == 0:
            print("%d cases, %d pictures compared, %d mismatching cases, %.0f s" % (k + 1, pics, len(bad), time.time() - t0), flush=True)
    print("SOAK %s: %d cases (seed %d), %d pictures bit-exact vs the oracle, %d failures" % ("FAILED" if bad else "ok", n, seed, pics, len(bad)))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
