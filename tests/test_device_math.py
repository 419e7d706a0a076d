"""CPU self-test of the kernels' arithmetic primitives (h263_rs_b200/csrc/device_math.cuh
compiled for the host) against the oracle and the reference's constant tables.  This is
how arithmetic bugs are caught here, where no GPU exists; the -m gpu tests then cover
the real kernels end to end."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "device_math_host.cpp")
OUT = os.path.join(HERE, "native", "libdevice_math_host.so")


@pytest.fixture(scope="module")
def dm():
    dep = os.path.join(HERE, "..", "h263_rs_b200", "csrc", "device_math.cuh")
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(dep)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", OUT, SRC])
    L = C.CDLL(OUT)
    L.dm_basis.restype = C.c_float
    for f in ("dm_round_residual", "dm_round_residual_scaled", "dm_round_residual_dc"):
        getattr(L, f).argtypes = [C.c_float]
    L.dm_avg2.restype = C.c_uint32
    L.dm_avg4.restype = C.c_uint32
    L.dm_avg2.argtypes = [C.c_uint32] * 2
    L.dm_avg4.argtypes = [C.c_uint32] * 4
    L.dm_yuv_pixel.restype = C.c_uint32
    L.dm_block_transform.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p]
    return L


def test_tables_match_reference(dm):
    k = json.load(open(os.path.join(HERE, "golden", "kat_constants.json")))
    for p, (x, y) in enumerate(k["dezigzag_xy"]):
        assert dm.dm_dezigzag(p) == y * 8 + x
    basis = [np.float32(t) for t in k["basis_table_f32_literals"]]
    for f in range(8):
        for i in range(8):
            assert np.float32(dm.dm_basis(f, i)) == basis[f * 8 + i]
    assert [dm.dm_q2s(q) for q in range(32)] == k["quant_to_strength"]


def test_dequant_matches_oracle_including_i16_wrap(dm):
    for quant in range(1, 32):
        for level in list(range(-1024, 1024, 7)) + [-1024, -1023, -529, -528, -1, 1, 528, 529, 1023]:
            if level == 0:
                continue
            cls, blk = O.inverse_rle(None, [0], [level], quant)
            assert dm.dm_dequant(level, quant) == int(blk[0, 0]), (quant, level)
    for code in range(1, 256):
        if code != 128:
            cls, blk = O.inverse_rle(code, [], [], 1)
            assert dm.dm_intradc_level(code) == int(blk[0, 0])


def test_block_transform_matches_oracle_idct(dm):
    rng = np.random.default_rng(5)
    pred = np.full((8, 8), 128, np.uint8)  # residual range [-128,127] observable; use two preds
    for trial in range(3000):
        kind = trial % 4
        c = np.zeros((8, 8), np.float32)
        if kind == 0:  # sparse low-frequency (typical)
            n = rng.integers(1, 6)
            c[rng.integers(0, 3, n), rng.integers(0, 3, n)] = rng.integers(-300, 300, n)
        elif kind == 1:  # dense
            c[:] = rng.integers(-2048, 2048, (8, 8))
        elif kind == 2:  # column only -> Vert
            c[: rng.integers(2, 9), 0] = rng.integers(-2048, 2048, 1)[0] or 5
            c[1, 0] = c[1, 0] or 7
        elif trial % 8 == 3:  # (0,0) + (0,4): the family where Vert and Full round differently
            c[0, 0], c[4, 0] = rng.integers(-60, 60), rng.integers(1, 60)
        else:  # row only -> Horiz (computed as Full by the kernel)
            c[0, : rng.integers(2, 9)] = rng.integers(-2048, 2048, 1)[0] or 5
            c[0, 1] = c[0, 1] or 7
        nz_rows = [y for y in range(8) if c[y].any()]
        nz_cols = [x for x in range(8) if c[:, x].any()]
        if not nz_rows or (nz_rows == [0] and nz_cols == [0]):
            continue
        if nz_cols == [0]:
            cls_o, cls_k = 3, 3
        elif nz_rows == [0]:
            cls_o, cls_k = 2, 4
        else:
            cls_o, cls_k = 4, 4
        rows = sum(1 << y for y in nz_rows)
        res = np.zeros(64, np.int32)
        cc = np.ascontiguousarray(c.reshape(64))
        dm.dm_block_transform(cls_k, cc.ctypes.data, rows, res.ctypes.data)
        res = res.reshape(8, 8)
        for base in (0, 255, 128):
            p = np.full((8, 8), base, np.uint8)
            exp = O.idct_block(cls_o, c, p)
            got = np.clip(p.astype(np.int32) + res, 0, 255).astype(np.uint8)
            assert np.array_equal(exp, got), (trial, cls_o)


def test_rounding_forms(dm):
    rng = np.random.default_rng(6)
    for dc in range(-2048, 2048):
        c = np.zeros((8, 8), np.float32)
        c[0, 0] = dc
        exp = int(O.idct_block(1, c, np.full((8, 8), 128, np.uint8))[0, 0]) - 128
        got = dm.dm_round_residual_dc(C.c_float(dc))
        assert max(-128, min(127, got)) == exp


def test_packed_averages(dm):
    rng = np.random.default_rng(7)
    v = rng.integers(0, 2**32, (20000, 4), dtype=np.uint64).astype(np.uint32)
    corner = np.array([[0, 0, 0, 0], [0xFFFFFFFF] * 4, [0xFF00FF00, 0x00FF00FF, 0xFFFFFFFF, 0], [0x01010101, 0, 0, 0],
                       [0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFE, 0xFFFFFFFF]], np.uint32)
    for a, b, c, d in np.concatenate([v, corner]):
        ab = [(int(a) >> s) & 255 for s in (0, 8, 16, 24)]
        bb = [(int(b) >> s) & 255 for s in (0, 8, 16, 24)]
        cb = [(int(c) >> s) & 255 for s in (0, 8, 16, 24)]
        db = [(int(d) >> s) & 255 for s in (0, 8, 16, 24)]
        e2 = sum((((x + y + 1) >> 1) << s) for x, y, s in zip(ab, bb, (0, 8, 16, 24)))
        e4 = sum((((x + y + z + w + 2) >> 2) << s) for x, y, z, w, s in zip(ab, bb, cb, db, (0, 8, 16, 24)))
        assert dm.dm_avg2(int(a), int(b)) == e2
        assert dm.dm_avg4(int(a), int(b), int(c), int(d)) == e4


def test_average_sum_of_mvs(dm):
    for s in range(-128, 128):
        assert dm.dm_average_sum_of_mvs(s) == O.lib().orc_average_sum_of_mvs(s)


def test_yuv_pixel_matches_oracle_exhaustively_on_a_grid(dm):
    ys = np.arange(0, 256, 3)
    for cb in range(0, 256, 5):
        for cr in range(0, 256, 5):
            y = ys.astype(np.uint8)
            exp = O.yuv420_to_rgba(y, np.full((len(y) + 1) // 2, cb, np.uint8), np.full((len(y) + 1) // 2, cr, np.uint8),
                                   len(y)).view(np.uint32)
            got = np.array([dm.dm_yuv_pixel(int(v), cb, cr) for v in ys], np.uint32)
            assert np.array_equal(exp, got)


def test_deblock_process_both_semantics(dm):
    rng = np.random.default_rng(8)
    quads = rng.integers(0, 256, (20000, 4))
    for q in quads:
        s = int(rng.integers(0, 13))
        for trunc in (0, 1):
            arr = (C.c_int * 4)(*[int(v) for v in q])
            dm.dm_deblock_process(arr, s, trunc)
            exp = O.deblock_process([int(v) for v in q], s, simd=not trunc) if s > 0 else None
            if exp is not None:
                assert list(arr) == exp, (q, s, trunc)
            else:
                assert list(arr) == [int(v) for v in q]  # strength 0: no-op


def test_identities_of_the_v15_recon_kernel_hold_for_every_input():
    """The recon kernel's epilogue replaces three pieces of the reference's arithmetic by identities (recon_tile.cu);
    each is checked here over its whole input domain, in integers, with the semantics of the SASS instructions used:
      * BT.601 (bt601.rs:12-59) on complemented terms: clamp(0xFFFFFF - x, 0, 0xFFFFFF) carries 255 - clamp(x >> 16, 0, 255)
        in byte 2 and 0 in byte 3 -- all 2^24 (y, cb, cr);
      * average_sum_of_mvs (types.rs:759-768) = ((s + 13) >> 4) + ((s + 2) >> 4) -- every sum of four i8 vectors;
      * the even lanes of a + b as (a + b) - 256 (O_a + O_b) modulo 2^32 -- random words."""
    y = np.arange(256, dtype=np.int64)[:, None, None]
    cb = np.arange(256, dtype=np.int64)[None, :, None]
    cr = np.arange(256, dtype=np.int64)[None, None, :]
    gray = (y - 16) * 76309
    ref = {
        "r": np.clip((gray + (cr - 128) * 104597 + 32768) >> 16, 0, 255) + 0 * cb,
        "g": np.clip((gray + (cr - 128) * -53279 + (cb - 128) * -25675 + 32768) >> 16, 0, 255),
        "b": np.clip((gray + (cb - 128) * 132201 + 32768) >> 16, 0, 255) + 0 * cr,
    }
    # the kernel's terms: t'_c = 0xFFFFFF - t_c with t_c = chroma part + 32768 - 16 * 76309 (folded constants)
    tr = cr * -104597 + (0xFFFFFF - (32768 - 128 * 104597 - 16 * 76309))
    tg = cr * 53279 + cb * 25675 + (0xFFFFFF - (32768 + 128 * 53279 + 128 * 25675 - 16 * 76309))
    tb = cb * -132201 + (0xFFFFFF - (32768 - 128 * 132201 - 16 * 76309))
    xr = y * -76309 + tr  # one IMAD
    dg, db = tg - tr, tb - tr  # green / blue as differences from red (added inside VIADDMNMX.RELU)
    for name, w in (("r", xr + 0 * cb), ("g", xr + dg), ("b", xr + db)):
        assert np.abs(w).max() < 2 ** 31  # no 32-bit overflow anywhere
        wc = np.clip(w, 0, 0xFFFFFF)  # max(min(a + b, c), 0)
        assert np.array_equal(255 - ((wc >> 16) & 0xFF), ref[name]), name
        assert not (wc >> 24).any()

    s = np.arange(-4 * 128, 4 * 127 + 1)
    whole, frac = (s >> 4) << 1, s & 15
    ref_avg = np.where(frac <= 2, whole, np.where(frac >= 14, whole + 2, whole + 1))
    assert np.array_equal(((s + 13) >> 4) + ((s + 2) >> 4), ref_avg)

    rng = np.random.default_rng(5)
    a = rng.integers(0, 2 ** 32, 200000, dtype=np.uint64)
    b = rng.integers(0, 2 ** 32, 200000, dtype=np.uint64)
    M = np.uint64(0x00FF00FF)
    odd = ((a >> np.uint64(8)) & M) + ((b >> np.uint64(8)) & M)
    even = ((a + b) - np.uint64(256) * odd) & np.uint64(0xFFFFFFFF)
    assert np.array_equal(even, (a & M) + (b & M))
