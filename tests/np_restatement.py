"""Independent numpy-f32 restatement of the reference's IDCT and motion compensation.

Written from the reference's formulas (idct.rs:52-65,82-201; gather.rs:16-126), NOT from
the C++ oracle: it exists so that two independently written restatements have to agree,
because the reference itself has no tests for these functions ("parity unpinned").
Vectorised with numpy; every float operation is a separately rounded float32 op.
"""
import numpy as np

F = np.float32


def basis_table():
    # the reference's f32 literals (idct.rs:39-48), parsed from decimal text by numpy
    txt = """0.70710677 0.70710677 0.70710677 0.70710677 0.70710677 0.70710677 0.70710677 0.70710677
    0.98078525 0.8314696 0.5555702 0.19509023 -0.19509032 -0.55557036 -0.83146966 -0.9807853
    0.9238795 0.38268343 -0.38268352 -0.9238796 -0.9238795 -0.38268313 0.3826836 0.92387956
    0.8314696 -0.19509032 -0.9807853 -0.55557 0.55557007 0.98078525 0.19509007 -0.8314698
    0.70710677 -0.70710677 -0.70710665 0.707107 0.70710677 -0.70710725 -0.70710653 0.7071068
    0.5555702 -0.9807853 0.19509041 0.83146936 -0.8314698 -0.19508928 0.9807853 -0.55557007
    0.38268343 -0.9238795 0.92387974 -0.3826839 -0.38268384 0.9238793 -0.92387974 0.3826839
    0.19509023 -0.55557 0.83146936 -0.9807852 0.98078525 -0.83147013 0.55557114 -0.19508967"""
    return np.array([F(v) for v in txt.split()], F).reshape(8, 8)


B = basis_table()


def idct_1d(v):
    """v: (..., 8) float32 -> (..., 8); out[i] = sum_freq v[freq]*B[freq][i], left to right from 0."""
    v = np.asarray(v, F)
    acc = np.zeros(v.shape[:-1] + (8,), F)
    for freq in range(8):
        acc = (acc + (v[..., freq : freq + 1] * B[freq]).astype(F)).astype(F)
    return acc


def _round(v, scale=None):
    v = np.asarray(v, F)
    x = v if scale is None else (v * F(scale)).astype(F)
    sgn = np.where(np.signbit(v), F(-0.5), F(0.5)).astype(F)
    r = ((x / F(4.0)).astype(F) + sgn).astype(F)
    r = np.trunc(r).astype(np.int32)
    return np.clip(r, -256, 255)


def idct_block(cls, coefs, pred):
    """cls: 0 Zero, 1 Dc, 2 Horiz, 3 Vert, 4 Full; coefs (8,8) [y][x]; pred (8,8) u8 -> (8,8) u8."""
    coefs = np.asarray(coefs, F).reshape(8, 8)
    pred = np.asarray(pred, np.int32).reshape(8, 8)
    if cls == 0:
        res = np.zeros((8, 8), np.int32)
    elif cls == 1:
        dc = coefs[0, 0]
        v = ((dc * F(0.5)).astype(F) / F(4.0)).astype(F)
        sgn = F(-0.5) if np.signbit(dc) else F(0.5)
        r = int(np.clip(np.trunc(F(v + sgn)), -256, 255))
        res = np.full((8, 8), r, np.int32)
    elif cls == 2:
        o = idct_1d(coefs[0])
        res = np.tile(_round(o, B[0, 0])[None, :], (8, 1))
    elif cls == 3:
        o = idct_1d(coefs[:, 0])
        res = np.tile(_round(o, B[0, 0])[:, None], (1, 8))
    else:
        t = idct_1d(coefs)  # t[y][i]
        out = idct_1d(t.T.copy())  # out[i][j]; pixel (x=i, y=j)
        res = _round(out).T
    return np.clip(pred + res, 0, 255).astype(np.uint8)


def gather_block(src, pos, mv):
    """8x8 prediction at pos=(x,y) from plane src (h,w) with half-pel mv=(mvx,mvy)."""
    h, w = src.shape
    mvx, mvy = mv
    dx, ix = mvx >> 1, mvx & 1  # floor division / oddness == HalfPel::into_lerp_parameters
    dy, iy = mvy >> 1, mvy & 1
    xs = np.arange(8) + pos[0] + dx
    ys = np.arange(8) + pos[1] + dy
    s = src.astype(np.uint16)

    def samp(yy, xx):
        return s[np.clip(yy, 0, h - 1)[:, None], np.clip(xx, 0, w - 1)[None, :]]

    a = samp(ys, xs)
    if ix and iy:
        return ((a + samp(ys, xs + 1) + samp(ys + 1, xs) + samp(ys + 1, xs + 1) + 2) // 4).astype(np.uint8)
    if ix:
        return ((a + samp(ys, xs + 1) + 1) // 2).astype(np.uint8)
    if iy:
        return ((a + samp(ys + 1, xs) + 1) // 2).astype(np.uint8)
    return a.astype(np.uint8)
