"""The reference's own parser known-answer tests (reader.rs:448-559, macroblock.rs:551-1010, block.rs:757-2124),
replayed on the PRODUCT front end -- its 64-bit-window bit reader, its table-driven VLC decode and its block
decoder -- through the h263cu_test_* hooks of libh263cu.so.  (tests/test_oracle_kats.py replays the same vectors on
the oracle.)  No GPU needed."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from h263_rs_b200 import _lib

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TABLE_ID = {"MCBPC_I_TABLE": 0, "MCBPC_P_TABLE": 1, "CBPY_TABLE_INTRA": 2, "MVD_TABLE": 3, "TCOEF_TABLE": 4}
WIDTH = {"u8": 8, "i8": 8, "u16": 16, "i16": 16, "u32": 32, "i32": 32, "u64": 64, None: 32}


def load(name):
    with open(os.path.join(G, name)) as f:
        return json.load(f)


def read_bits(data, pos, n, signed=False, peek=False):
    p = C.c_size_t(pos)
    v = C.c_int64()
    e = _lib.lib().h263cu_test_read_bits(data, len(data), C.byref(p), n, int(signed), int(peek), C.byref(v))
    return e, v.value, p.value


# reader.rs:448-559
@pytest.mark.parametrize("case", load("kat_reader.json"), ids=lambda c: c["name"])
def test_reader_kats_on_the_product_reader(case):
    data = bytes(case["data"])
    pos = 0
    for op in case["ops"]:
        name = op["op"]
        if op.get("bits", 0) and op["bits"] > 32:
            continue  # the product reader never reads more than 32 bits at once (the longest field is 17 bits)
        if name == "skip_bits":
            e, _, pos = read_bits(data, pos, op["bits"])
            assert e == 0
        elif name == "recognize_start_code":
            if op["in_error"]:
                continue  # the error-resynchronisation search is the GOB stub's (gob.rs:50-71); not a product path
            sk = C.c_int()
            e = _lib.lib().h263cu_test_start_code(data, len(data), pos, C.byref(sk))
            if op["expect"] is None:
                assert e != 0 and sk.value == -1
            else:
                assert e == 0 and sk.value == op["expect"]
        else:
            signed = "signed" in name
            peek = name.startswith("peek")
            e, v, newpos = read_bits(data, pos, op["bits"], signed, peek)
            if op.get("expect_err"):
                assert e == _lib.ERR_UNHANDLED_IO_ERROR
                continue
            assert e == 0
            exp = op["expect"]
            if signed and op.get("cast") is None and op.get("type", "i")[0] == "u":
                exp &= (1 << WIDTH[op["type"]]) - 1
                v &= (1 << WIDTH[op["type"]]) - 1
            assert v == exp, (op, v)
            pos = newpos


def run_vlc_seq(t):
    data = bytes(t["data"])
    pos = C.c_size_t(0)
    out = (C.c_int * 4)()
    n = 0
    for step in t["seq"]:
        if step["table"] == "MODB_TABLE":
            return n  # MODB (PB frames) is not on any decodable path (macroblock.rs:461-465)
        e = _lib.lib().h263cu_test_read_vlc(TABLE_ID[step["table"]], data, len(data), C.byref(pos), out)
        assert e == 0
        got = list(out)
        exp = step["expect"]
        if exp[0] != 0:
            assert got[0] == exp[0], (n, got, exp)
        else:
            assert got == exp, (n, got, exp)
        n += 1
    return n


# macroblock.rs:561-1009
@pytest.mark.parametrize("case", load("kat_mb_tables.json"), ids=lambda c: c["name"])
def test_macroblock_table_kats_on_the_product_tables(case):
    n = run_vlc_seq(case)
    if case["name"] != "macroblock_modb_table":
        assert n == len(case["seq"]) and n >= 10


# block.rs:768-1705
def test_tcoef_table_kat_on_the_product_table():
    assert run_vlc_seq(load("kat_block.json")["tcoef_table"]) == 102


# block.rs:1707-2123
@pytest.mark.parametrize("case", load("kat_block.json")["decode_block"], ids=lambda c: c["name"])
def test_decode_block_kats_on_the_product_block_decoder(case):
    data = bytes(case["data"])
    pos = C.c_size_t(0)
    dc, n, ovf = C.c_int(), C.c_int(), C.c_int()
    run = np.zeros(64, np.uint8)
    level = np.zeros(64, np.int16)
    e = _lib.lib().h263cu_test_decode_block(data, len(data), C.byref(pos), 1 if case["sorenson"] else 0, case["version"],
                                            int(case["intra"]), int(case["tcoef_present"]), C.byref(dc), C.byref(n),
                                            run.ctypes.data, level.ctypes.data, C.byref(ovf))
    assert e == 0 and not ovf.value
    if case["expect_intradc_level"] is None:
        assert dc.value == -1
    else:
        lvl = 1024 if dc.value == 255 else dc.value << 3  # IntraDc::into_level
        assert lvl == case["expect_intradc_level"]
    assert n.value == len(case["expect_tcoef"])
    for i, ev in enumerate(case["expect_tcoef"]):
        # is_short (block.rs:700-741) is parser-internal: both forms reach the device as (run, level)
        assert (int(run[i]), int(level[i])) == (ev["run"], ev["level"])


def test_every_code_of_every_table_decodes_to_itself():
    """Walks every code of the generated tables through the product's LUT decode: each code, followed by arbitrary
    bits, must come back as its own symbol with its own length (prefix-freeness + LUT fill)."""
    import re

    src = open(os.path.join(os.path.dirname(_lib.HERE), "h263_rs_b200", "csrc", "vlc_codes.inc")).read()
    tables = re.findall(r"static const VlcCode (\w+)_CODES\[\] = \{(.*?)\};", src, re.S)
    ids = {"MCBPC_I": 0, "MCBPC_P": 1, "CBPY": 2, "MVD": 3, "TCOEF": 4}
    seen = 0
    for name, body in tables:
        for bits, ln, kind, a, b, c in re.findall(r'\{"([01]+)",\s*(\d+),\s*(-?\d+),\s*(-?\d+),\s*(-?\d+),\s*(-?\d+)\}', body):
            for tail in ("0" * 24, "1" * 24, "10" * 12):
                s = bits + tail
                data = bytes(int(s[i : i + 8].ljust(8, "0"), 2) for i in range(0, len(s), 8))
                pos = C.c_size_t(0)
                out = (C.c_int * 4)()
                assert _lib.lib().h263cu_test_read_vlc(ids[name], data, len(data), C.byref(pos), out) == 0
                assert pos.value == int(ln) and list(out) == [int(kind), int(a), int(b), int(c)], (name, bits)
            seen += 1
    assert seen > 150
