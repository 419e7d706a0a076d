"""Two independent restatements (C++ oracle, numpy) of the functions the reference never
tests -- IDCT, inverse RLE classes, motion compensation -- have to agree bit for bit."""
import numpy as np

import np_restatement as N
import oracle_lib as O


def rand_block(rng, cls):
    c = np.zeros((8, 8), np.float32)
    vals = lambda n: rng.integers(-2048, 2048, n).astype(np.float32)
    if cls == 1:
        c[0, 0] = vals(1)[0] or 3
    elif cls == 2:
        k = rng.integers(1, 8)
        c[0, : k + 1] = vals(k + 1)
    elif cls == 3:
        k = rng.integers(1, 8)
        c[: k + 1, 0] = vals(k + 1)
    elif cls == 4:
        n = rng.integers(2, 40)
        ys, xs = rng.integers(0, 8, n), rng.integers(0, 8, n)
        c[ys, xs] = vals(n)
    return c


def test_idct_classes_agree():
    rng = np.random.default_rng(1)
    for cls in (0, 1, 2, 3, 4):
        for _ in range(400):
            c = rand_block(rng, cls)
            pred = rng.integers(0, 256, (8, 8)).astype(np.uint8)
            assert np.array_equal(O.idct_block(cls, c, pred), N.idct_block(cls, c, pred)), cls


def test_dc_tie_values_and_small_magnitudes():
    # dc = 4 (mod 8) are exact ties of dc/8; Dc differs from Full there
    pred = np.full((8, 8), 128, np.uint8)
    diff = 0
    for dc in list(range(-2048, 2048)):
        c = np.zeros((8, 8), np.float32)
        c[0, 0] = dc
        a = O.idct_block(1, c, pred)
        assert np.array_equal(a, N.idct_block(1, c, pred))
        exp = np.sign(dc) * ((abs(dc) + 4) >> 3)
        assert int(a[0, 0]) == int(np.clip(128 + np.clip(exp, -256, 255), 0, 255))
        diff += not np.array_equal(a, O.idct_block(4, c, pred))
    assert diff > 100  # the classes really are observable


def test_horiz_equals_full_but_vert_does_not_always():
    rng = np.random.default_rng(2)
    pred = np.zeros((8, 8), np.uint8) + 128
    for _ in range(3000):
        c = rand_block(rng, 2)
        assert np.array_equal(O.idct_block(2, c, pred), O.idct_block(4, c, pred))
    # Vert is observable: with coefficients at (0,0) and (0,4) only, about 8 % of the blocks
    # round differently from the Full path (ties of v/4), so the class must be reproduced.
    vdiff = n = 0
    for dc in range(-60, 60, 3):
        for c4 in range(-60, 60):
            if c4 == 0:
                continue
            c = np.zeros((8, 8), np.float32)
            c[0, 0], c[4, 0] = dc, c4
            a, b = O.idct_block(3, c, pred), O.idct_block(4, c, pred)
            assert np.array_equal(a, N.idct_block(3, c, pred)) and np.array_equal(b, N.idct_block(4, c, pred))
            n += 1
            vdiff += not np.array_equal(a, b)
    assert vdiff > n // 50


def test_gather_block_agrees_all_modes_and_borders():
    rng = np.random.default_rng(3)
    src = rng.integers(0, 256, (40, 56)).astype(np.uint8)
    for _ in range(3000):
        pos = (int(rng.integers(0, 7)) * 8, int(rng.integers(0, 5)) * 8)
        mv = (int(rng.integers(-70, 70)), int(rng.integers(-70, 70)))
        dst = np.zeros_like(src)
        got = O.gather_block(src, pos, mv, dst)[pos[1] : pos[1] + 8, pos[0] : pos[0] + 8]
        assert np.array_equal(got, N.gather_block(src, pos, mv)), (pos, mv)


def test_inverse_rle_semantics():
    # overflow drops the whole block, DC included (rle.rs:125-127)
    cls, blk = O.inverse_rle(100, [10, 60], [1, 1], 5)
    assert cls == 0 and not blk.any()
    # an inter block whose only event sits at index 0 is Dc, never Zero
    cls, blk = O.inverse_rle(None, [0], [-1], 4)
    assert cls == 1 and blk[0, 0] == -(4 * 3 - 1)
    # parity rule: odd QP -> no -1
    cls, blk = O.inverse_rle(None, [0], [2], 5)
    assert blk[0, 0] == 25
    # clamp to [-2048, 2047]
    cls, blk = O.inverse_rle(None, [0], [127], 31)
    assert blk[0, 0] == 2047
    cls, blk = O.inverse_rle(None, [0], [-127], 31)
    assert blk[0, 0] == -2048
    # release-mode i16 wrap: 31 * (2*1023+1) = 63457 -> -2079 as i16 -> clamp -2048
    cls, blk = O.inverse_rle(None, [0], [1023], 31)
    assert blk[0, 0] == -2048
    # intra DC 0xFF => 1024, and class Vert/Horiz/Full
    cls, blk = O.inverse_rle(255, [0], [1], 2)  # idx 1 = (x=1,y=0) -> Horiz
    assert cls == 2 and blk[0, 0] == 1024
    cls, blk = O.inverse_rle(10, [1], [1], 2)  # idx 2 = (x=0,y=1) -> Vert
    assert cls == 3 and blk[0, 0] == 80 and blk[1, 0] == 5
    cls, blk = O.inverse_rle(10, [3], [1], 2)  # idx 4 = (1,1) -> Full
    assert cls == 4


def test_mv_helpers():
    L = O.lib()
    for s in range(-128, 128):
        whole, frac = (s >> 4) << 1, s & 15
        exp = whole if frac <= 2 else (whole + 2 if frac >= 14 else whole + 1)
        assert L.orc_average_sum_of_mvs(s) == exp
    for a in range(-5, 6):
        for b in range(-5, 6):
            for c in range(-5, 6):
                assert L.orc_median_of(a, b, c) == sorted((a, b, c))[1]
    for pred in range(-32, 32):
        for mvd in range(-32, 32):
            out = L.orc_halfpel_decode(pred, mvd)
            assert -32 <= out < 32 and (out - pred - mvd) % 64 == 0
