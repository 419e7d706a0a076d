"""Parity at BASELINE.json's full sizes (configs[2] and configs[3]) through size-independent properties.

The oracle cannot decode 1024 CIF streams in test time, so a full-size step is built from a few unique streams
replicated over all stream slots: every replica sits at a different place of the step (other tiles, other
plane slots, other CTAs) and must produce the checksums the oracle computes for its unique stream; and the two
independent device implementations (tiled kernel, generic warp-per-macroblock kernel) must agree on every one
of the step's macroblocks."""
import numpy as np
import pytest

from helpers import oracle_decode_stream, weighted_sum
from h263_rs_b200 import _lib, api, synth

pytestmark = pytest.mark.gpu


def oracle_sums(ref):
    return [[weighted_sum(r["y"]), weighted_sum(r["cb"]), weighted_sum(r["cr"]), weighted_sum(r["rgba"])] for r in ref]


def run_full(n_streams, w, h, unique, t_steps, out_flags, deblock, **gen):
    streams = [synth.make_stream(w, h, t_steps, 9000 + u, deblock_flag=1 if deblock else 0, **gen) for u in range(unique)]
    want = [oracle_sums(oracle_decode_stream(p, deblock=deblock)) for p in streams]
    dec = api.BatchDecoder(n_streams, w, h, threads=0)
    all_sums = []
    for t in range(t_steps):
        errs = dec.decode_step([streams[s % unique][t] for s in range(n_streams)], out_flags)
        assert not errs.any()
        dec.ctx.sync()
        sums = dec.ctx.checksums(np.arange(n_streams))
        for s in range(n_streams):
            assert [int(v) for v in sums[s]] == want[s % unique][t], (s, t)
        all_sums.append(np.array(sums, dtype=np.uint64).copy())
    dec.ctx.close()
    return all_sums


def test_config3_1024_cif_streams_replicas_match_oracle():
    """configs[2]: 1024 concurrent CIF streams on one GPU, fused MC + IDCT + YUV->RGBA."""
    run_full(1024, 352, 288, unique=24, t_steps=4, out_flags=_lib.OUT_RGBA, deblock=False)


def test_config4_256_4cif_streams_deblock_border_vectors():
    """configs[3]: 256 concurrent 4CIF streams, deblocking on, vectors biased across all four borders."""
    run_full(256, 704, 576, unique=6, t_steps=3, out_flags=_lib.OUT_RGBA | _lib.OUT_DEBLOCK, deblock=True, mv_mode=2)


def test_full_size_tiled_and_generic_kernels_agree(monkeypatch):
    """Two independent device implementations over a full-size step: identical checksums for all 1024 streams
    (the unique streams differ per slot here: 1024 different seeds, no oracle involved)."""
    n, t_steps = 1024, 3
    streams = [synth.make_stream(352, 288, t_steps, 20000 + s, mv_mode=s % 3, pct_fourmv=(s * 7) % 40) for s in range(n)]
    got = {}
    for force in ("tile", "mb"):
        monkeypatch.setenv("H263CU_KERNEL", force)
        dec = api.BatchDecoder(n, 352, 288, threads=0)
        per_step = []
        for t in range(t_steps):
            assert not dec.decode_step([streams[s][t] for s in range(n)]).any()
            dec.ctx.sync()
            per_step.append(np.array(dec.ctx.checksums(np.arange(n)), dtype=np.uint64).copy())
        got[force] = per_step
        dec.ctx.close()
    for t in range(t_steps):
        assert np.array_equal(got["tile"][t], got["mb"][t]), t
