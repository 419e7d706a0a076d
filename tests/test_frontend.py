"""Host front end (product parser + stream generator) against the oracle's parser: every
symbol the generator emits must be parsed back identically by both, and both must fail
with the same error on damaged streams."""
import numpy as np
import pytest

import oracle_lib as O
from helpers import compare_parse_with_oracle
from h263_rs_b200 import _lib, frontend, synth

CASES = [
    # BASELINE.json config 1: QCIF, 1 I + 29 P
    dict(w=176, h=144, n=30, seed=1),
    dict(w=352, h=288, n=6, seed=2, mv_mode=1),
    dict(w=352, h=288, n=5, seed=3, mv_mode=2, pct_escape=20, permille_overflow=30, truncate_permille=300),
    dict(w=352, h=288, n=4, seed=4, version=0, pct_escape=30),
    dict(w=176, h=144, n=6, seed=5, flavour=1, pct_escape=10),
    dict(w=160, h=120, n=6, seed=6, pct_fourmv=30, pct_dquant=40),
    dict(w=200, h=100, n=6, seed=7, intra_period=3),
    dict(w=704, h=576, n=2, seed=8, deblock_flag=1),
    dict(w=128, h=96, n=4, seed=9, mean_events_x10=200, pct_cbp_inter=90),  # > 32 events per MB
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_s%d" % (c["w"], c["h"], c["seed"]))
def test_generator_roundtrip_and_parser_agreement(case):
    kw = {k: v for k, v in case.items() if k not in ("w", "h", "n", "seed")}
    packets = synth.make_stream(case["w"], case["h"], case["n"], case["seed"], **kw)
    opt = 0 if kw.get("flavour", 0) == 1 else 1
    nmb, nev, nerr = compare_parse_with_oracle(packets, opt)
    mbw, mbh = (case["w"] + 15) // 16, (case["h"] + 15) // 16
    assert nerr == 0 and nmb == case["n"] * mbw * mbh and nev > 0


def test_generator_is_deterministic():
    a = synth.make_stream(176, 144, 3, 42)
    b = synth.make_stream(176, 144, 3, 42)
    c = synth.make_stream(176, 144, 3, 43)
    assert a == b and a != c


def test_damaged_streams_fail_identically():
    """Bit flips, truncation and garbage: product and oracle agree on error code or on the
    parsed content (same transactional behaviour: a failed picture leaves the state alone)."""
    rng = np.random.default_rng(5)  # seed 5 reaches packets whose trailing COD=1 bits run past the picture's capacity
    bases = [synth.make_stream(176, 144, 4, 21, pct_escape=10), synth.make_stream(64, 48, 5, 22, pct_fourmv=30, mv_mode=1),
             synth.make_stream(176, 144, 3, 23, flavour=1), synth.make_stream(48, 32, 6, 24, version=0, pct_escape=25)]
    for trial in range(900):
        base = bases[trial % len(bases)]
        pk = [bytearray(p) for p in base]
        victim = int(rng.integers(0, len(pk)))
        mode = trial % 3
        if mode == 0:
            for _ in range(int(rng.integers(1, 4))):
                pos = int(rng.integers(0, len(pk[victim])))  # header bytes included
                pk[victim][pos] ^= 1 << int(rng.integers(0, 8))
        elif mode == 1:
            pk[victim] = pk[victim][: int(rng.integers(1, len(pk[victim])))]
        else:
            pk[victim] = pk[victim] + bytes(rng.integers(0, 256, int(rng.integers(1, 6)), dtype=np.uint8))
        compare_parse_with_oracle([bytes(p) for p in pk], 0 if base is bases[2] else 1)


def test_error_values():
    ps = frontend.Parser(1)
    with pytest.raises(_lib.H263Error) as e:
        ps.parse_picture(b"\x00\x00")  # fewer than 17 bits: UnexpectedEof
    assert e.value.code == -16 and e.value.is_eof_error()
    with pytest.raises(_lib.H263Error) as e:
        ps.parse_picture(b"\xff\xff\xff\xff\xff")  # no start code
    assert e.value.code == -2
    # a P picture without a reference: UncodedIFrameBlocks (gather.rs:149)
    pk = synth.make_stream(176, 144, 2, 3)
    with pytest.raises(_lib.H263Error) as e:
        frontend.Parser(1).parse_picture(pk[1])
    assert e.value.code == -15
    st = O.OracleState(1)
    with pytest.raises(O.OracleError) as oe:
        st.decode_next_picture(pk[1])
    assert oe.value.code == 15
    # reserved Sorenson size code 7 -> PictureFormatInvalid
    bad = bytearray(pk[0])
    bad[3] = (bad[3] & 0xFC) | 0x03  # size code bits straddle bytes 3/4: xxxxxx11 1xxxxxxx
    bad[4] |= 0x80
    with pytest.raises(_lib.H263Error) as e:
        frontend.Parser(1).parse_picture(bytes(bad))
    with pytest.raises(O.OracleError) as oe:
        O.OracleState(1).decode_next_picture(bytes(bad))
    assert e.value.code == -oe.value.code == -14


def test_peek_picture():
    pk = synth.make_stream(352, 288, 2, 5, deblock_flag=1)
    h = frontend.peek_picture(pk[1])
    assert (h["width"], h["height"], h["mb_w"], h["mb_h"], h["n_mbs"]) == (352, 288, 22, 18, 396)
    assert h["pic_type"] == _lib.PIC_P and h["flags"] & 1 and h["temporal_reference"] == 1


def test_parse_step_threaded_matches_serial():
    n = 12
    streams = [synth.make_stream(176, 144, 3, 100 + s, mv_mode=s % 3) for s in range(n)]
    par = [frontend.Parser(1) for _ in range(n)]
    ser = [frontend.Parser(1) for _ in range(n)]
    for t in range(3):
        packets = [streams[s][t] for s in range(n)]
        if t == 2:
            packets[5] = b"\x00\x00\x80"  # one stream fails; the others are packed densely
        pics, mbs, events, errs, pic_of = frontend.parse_step(par, packets, threads=4)
        k = 0
        for s in range(n):
            try:
                pic, m, ev = ser[s].parse_picture(packets[s], stream=s)
            except _lib.H263Error as e:
                assert errs[s] == e.code and pic_of[s] == -1
                continue
            assert errs[s] == 0 and pic_of[s] == k
            p = pics[k]
            assert p["stream"] == s and p["n_mbs"] == len(m) and p["n_event_units"] == len(ev)
            assert p["flags"] & 4  # H263CU_PICFLAG_MV_IN_RANGE: parsed vectors never leave [-32, 31]
            got_m = mbs[p["first_mb"] : p["first_mb"] + p["n_mbs"]].copy()
            assert (got_m["pic"] == k).all()
            got_m["pic"] = 0
            assert np.array_equal(got_m, m)
            assert np.array_equal(events[p["first_event"] : p["first_event"] + p["n_event_units"]], ev)
            k += 1
        assert k == len(pics)


OPT_DISPOSABLE = 0x100  # H263CU_OPT_DECODE_DISPOSABLE / ORC_OPT_DECODE_DISPOSABLE: an extension beyond the reference


def test_disposable_pictures_extension_and_reference_behaviour():
    """Sorenson disposable P pictures (type code 2).  Without the extension bit product and oracle fail them exactly
    like the reference (UnimplementedDecoding at the first coded macroblock, macroblock.rs:461-465) and the stream
    state stays put; with it both parse them like P pictures, flag them, and keep predicting from the last
    non-disposable picture."""
    pk = synth.make_stream(176, 144, 10, 77, pct_disposable=50, mv_mode=1, pct_fourmv=10)
    types = [int(frontend.peek_picture(p)["pic_type"]) for p in pk]
    assert types[0] == _lib.PIC_I and _lib.PIC_DISPOSABLE_P in types and _lib.PIC_P in types
    # reference behaviour: the first disposable picture is refused with error 17 by both
    nmb, nev, nerr = compare_parse_with_oracle(pk, 1)
    assert nerr >= 1
    first = types.index(_lib.PIC_DISPOSABLE_P)
    with pytest.raises(_lib.H263Error) as e:
        ps = frontend.Parser(1)
        for p in pk[: first + 1]:
            ps.parse_picture(p)
    assert e.value.code == -17
    # extension: everything parses, identically in product and oracle
    nmb, nev, nerr = compare_parse_with_oracle(pk, 1 | OPT_DISPOSABLE)
    assert nerr == 0 and nmb == 10 * 99
    ps = frontend.Parser(1 | OPT_DISPOSABLE)
    for p, t in zip(pk, types):
        pic, _, _ = ps.parse_picture(p)
        assert bool(pic["flags"][0] & 8) == (t == _lib.PIC_DISPOSABLE_P)  # H263CU_PICFLAG_DISPOSABLE
