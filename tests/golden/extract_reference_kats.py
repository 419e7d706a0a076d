#!/usr/bin/env python3
"""Extract the reference's own known-answer tests into JSON fixtures.

Runs ONLY in the build container (reads /root/reference, which does not exist on the
GPU box).  It scrapes the `#[cfg(test)]` modules of the reference crates and writes
the vectors -- inputs and expected outputs, nothing else -- as JSON next to this
script.  tests/test_oracle_kats.py replays them against the oracle, and the GPU tests
replay the colour-conversion and deblock ones against the CUDA kernels.

Sources (relative to /root/reference):
  h263/src/parser/reader.rs:444-560      11 bit-reader tests        -> kat_reader.json
  h263/src/parser/macroblock.rs:551-1010  4 VLC-table tests (+MODB, unused) -> kat_mb_tables.json
  h263/src/parser/block.rs:757-2124      TCOEF table + 8 decode_block tests -> kat_block.json
  yuv/src/bt601.rs:198-483               yuv_to_rgb / yuv420_to_rgba -> kat_yuv.json
  deblock/src/deblock.rs:319-559         process table + 11x17 picture -> kat_deblock.json
  h263/src/decoder/cpu/rle.rs:6-71, idct.rs:39-48, deblock.rs:5-8  constant tables -> kat_constants.json
"""
import json
import re
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def strip_comments(s):
    return re.sub(r"//[^\n]*", "", s)


def test_fns(text):
    """Yield (name, body) for every `#[test] fn name() { ... }`."""
    for m in re.finditer(r"#\[test\](?:\s*#\[[^\]]*\])*\s*fn\s+(\w+)\s*\(\)\s*\{", text):
        i = m.end()
        depth = 1
        while depth:
            c = text[i]
            depth += c == "{"
            depth -= c == "}"
            i += 1
        yield m.group(1), text[m.end() : i - 1]


def rust_int(tok):
    tok = tok.strip().replace("_", "")
    tok = re.sub(r"(u8|u16|u32|i8|i16|i32|usize)$", "", tok)
    neg = tok.startswith("-")
    if neg:
        tok = tok[1:]
    v = int(tok, 0)
    return -v if neg else v


def int_list(s):
    return [rust_int(t) for t in re.findall(r"-?(?:0b[01_]+|0x[0-9A-Fa-f_]+|\d[\d_]*)(?:u8|i16|u16)?", s)]


def first_array(body, var):
    m = re.search(r"let\s+%s\s*(?::[^=]*)?=\s*&?\[(.*?)\];" % var, body, re.S)
    assert m, var
    return int_list(m.group(1))


# ----------------------------------------------------------------------------- reader
def extract_reader():
    text = strip_comments((REF / "h263/src/parser/reader.rs").read_text())
    text = text[text.index("#[cfg(test)]") :]
    tests = []
    for name, body in test_fns(text):
        data = first_array(body, "data")
        ops = []
        for line in body.split(";"):
            line = " ".join(line.split())
            m = re.search(r"assert_eq!\(\s*(None|Some\((\d+)\))\s*,\s*reader\.recognize_start_code\((true|false)\)\.unwrap\(\)", line)
            if m:
                ops.append({"op": "recognize_start_code", "in_error": m.group(3) == "true",
                            "expect": None if m.group(1) == "None" else int(m.group(2))})
                continue
            m = re.search(r"assert_eq!\(\s*(-?\w+)\s*,\s*reader\.(\w+)(?:::<(\w+)>)?\((\w*)\)\.unwrap\(\)(?: as (\w+))?\s*\)", line)
            if m:
                exp, fn, ty, arg, cast = m.groups()
                if fn == "recognize_start_code":
                    continue
                ops.append({"op": fn, "bits": rust_int(arg) if arg else 8, "type": ty, "cast": cast,
                            "expect": rust_int(exp)})
                continue
            m = re.search(r"reader\.(\w+)(?:::<(\w+)>)?\((\w*)\)\.unwrap_err\(\)", line)
            if m:
                fn, ty, arg = m.groups()
                ops.append({"op": fn, "bits": rust_int(arg) if arg else 8, "type": ty, "expect_err": True})
                continue
            m = re.search(r"reader\.skip_bits\((\d+)\)\.unwrap\(\)", line)
            if m:
                ops.append({"op": "skip_bits", "bits": int(m.group(1))})
        tests.append({"name": name, "data": data, "ops": ops})
    assert len(tests) == 11, len(tests)
    return tests


# ----------------------------------------------------------------------------- tables
MBTYPE = {"Inter": 0, "InterQ": 1, "Inter4V": 2, "Intra": 3, "IntraQ": 4, "Inter4Vq": 5}


def norm_expected(table, exp):
    """Map the Rust expected value to our (kind, a, b, c) row (tools/gen_vlc_tables.py)."""
    exp = " ".join(exp.split())
    if table in ("MCBPC_I_TABLE", "MCBPC_P_TABLE"):
        if "Stuffing" in exp:
            return [1, 0, 0, 0]
        if "Invalid" in exp:
            return [2, 0, 0, 0]
        m = re.search(r"Valid\(MacroblockType::(\w+), (true|false), (true|false)\)", exp)
        return [0, MBTYPE[m.group(1)], int(m.group(2) == "true"), int(m.group(3) == "true")]
    if table == "CBPY_TABLE_INTRA":
        if exp.startswith("None"):
            return [2, 0, 0, 0]
        v = 0
        for t in re.findall(r"true|false", exp):
            v = (v << 1) | (t == "true")
        return [0, v, 0, 0]
    if table == "MVD_TABLE":
        if exp.startswith("None"):
            return [2, 0, 0, 0]
        import math

        m = re.search(r"Some\((-?[0-9.]+)\)", exp)
        return [0, int(math.floor(float(m.group(1)) * 2)), 0, 0]
    if table == "TCOEF_TABLE":
        if exp.startswith("None"):
            return [2, 0, 0, 0]
        if "EscapeToLong" in exp:
            return [3, 0, 0, 0]
        m = re.search(r"last: (true|false), run: (\d+), level: (\d+)", exp)
        return [0, int(m.group(1) == "true"), int(m.group(2)), int(m.group(3))]
    if table == "MODB_TABLE":
        t = re.findall(r"true|false", exp)
        return [0, int(t[0] == "true"), int(t[1] == "true"), 0]
    raise ValueError(table)


def extract_vlc_test(body):
    data = first_array(body, "bit_pattern")
    seq = []
    for m in re.finditer(r"assert_eq!\(\s*reader\.read_vlc\(&(\w+)\)\.unwrap\(\),\s*(.*?)\s*\);", body, re.S):
        seq.append({"table": m.group(1), "expect": norm_expected(m.group(1), m.group(2))})
    return {"data": data, "seq": seq}


def extract_mb_tables():
    text = strip_comments((REF / "h263/src/parser/macroblock.rs").read_text())
    text = text[text.index("#[cfg(test)]") :]
    tests = []
    for name, body in test_fns(text):
        t = extract_vlc_test(body)
        t["name"] = name
        tests.append(t)
    assert len(tests) == 5, len(tests)
    return tests


# ----------------------------------------------------------------------------- block
def extract_block():
    text = strip_comments((REF / "h263/src/parser/block.rs").read_text())
    text = text[text.index("#[cfg(test)]") :]
    out = {"tcoef_table": None, "decode_block": []}
    for name, body in test_fns(text):
        if name == "tcoef_table":
            out["tcoef_table"] = extract_vlc_test(body)
            continue
        data = first_array(body, "bitstream")
        version = re.search(r"version:\s*(None|Some\((\d+)\))", body)
        ver = -1 if version.group(1) == "None" else int(version.group(2))
        call = re.search(r"decode_block\(\s*&mut reader,\s*DecoderOption::(\w+)(?:\(\))?,\s*&picture,\s*PictureOption::empty\(\),\s*MacroblockType::(\w+),\s*(true|false)", body, re.S)
        sorenson = call.group(1) == "SORENSON_SPARK_BITSTREAM"
        exp = re.search(r"Block\s*\{\s*intradc:\s*(None|IntraDc::from_level\((0x[0-9A-Fa-f]+|\d+)\)),\s*tcoef:\s*vec!\[(.*?)\]\s*\}", body, re.S)
        dc_level = None if exp.group(1) == "None" else rust_int(exp.group(2))
        events = [
            {"is_short": m.group(1) == "true", "run": int(m.group(2)), "level": int(m.group(3))}
            for m in re.finditer(r"is_short:\s*(true|false),\s*run:\s*(\d+),\s*level:\s*(-?\d+)", exp.group(3))
        ]
        out["decode_block"].append({
            "name": name, "data": data, "version": ver, "sorenson": sorenson,
            "intra": MBTYPE[call.group(2)] in (3, 4), "tcoef_present": call.group(3) == "true",
            "expect_intradc_level": dc_level, "expect_tcoef": events,
        })
    assert out["tcoef_table"] and len(out["decode_block"]) == 8, len(out["decode_block"])
    return out


# ----------------------------------------------------------------------------- yuv
def extract_yuv():
    text = strip_comments((REF / "yuv/src/bt601.rs").read_text())
    single = []
    pictures = []
    for name, body in test_fns(text):
        if name == "test_yuv_to_rgb" or name == "test_yuv420_to_rgba_tiny":
            for m in re.finditer(r"assert_eq!\(\s*yuv_to_rgb\(\((\d+),\s*(\d+),\s*(\d+)\)\),\s*\((\d+),\s*(\d+),\s*(\d+)\)\s*\)", body):
                v = list(map(int, m.groups()))
                single.append({"yuv": v[:3], "rgb": v[3:]})
        if name.startswith("test_yuv420_to_rgba"):
            for m in re.finditer(r"assert_eq!\(\s*yuv420_to_rgba\(\s*&\[(.*?)\],\s*&\[(.*?)\],\s*&\[(.*?)\],\s*(\d+)\s*,?\s*\),\s*vec!\[(.*?)\]\s*\)", body, re.S):
                y, cb, cr, w, exp = m.groups()
                exp = exp.strip()
                if ";" in exp:  # vec![0u8; 0]
                    val, cnt = exp.split(";")
                    rgba = [rust_int(val)] * int(cnt)
                else:
                    rgba = int_list(exp)
                pictures.append({"test": name, "y": int_list(y), "cb": int_list(cb), "cr": int_list(cr),
                                 "width": int(w), "rgba": rgba})
    assert len(single) == 11 and len(pictures) == 10, (len(single), len(pictures))
    return {"yuv_to_rgb": single, "yuv420_to_rgba": pictures}


# ----------------------------------------------------------------------------- deblock
def extract_deblock():
    text = strip_comments((REF / "deblock/src/deblock.rs").read_text())
    text = text[text.index("#[cfg(test)]") :]
    res = {}
    for name, body in test_fns(text):
        if name == "test_process":
            rows = []
            for m in re.finditer(r"\(\((\d+),\s*(\d+),\s*(\d+),\s*(\d+)\),\s*(\d+),\s*\((\d+),\s*(\d+),\s*(\d+),\s*(\d+)\)\)", body):
                v = list(map(int, m.groups()))
                rows.append({"in": v[0:4], "strength": v[4], "out": v[5:9]})
            res["process"] = rows
        if name == "test_deblock":
            data = first_array(body, "data")
            cases = []
            for s in (4, 8, 12):
                exp = first_array(body, "expected_%d" % s)
                cases.append({"strength": s, "expected": exp})
            m = re.search(r"deblock\(data,\s*(\d+),", body)
            res["picture"] = {"width": int(m.group(1)), "data": data, "cases": cases}
    assert len(res["process"]) == 37, len(res["process"])
    assert len(res["picture"]["data"]) == 11 * 17
    return res


# ----------------------------------------------------------------------------- constants
def extract_constants():
    rle = strip_comments((REF / "h263/src/decoder/cpu/rle.rs").read_text())
    m = re.search(r"DEZIGZAG_MAPPING[^=]*=\s*\[(.*?)\];", rle, re.S)
    dz = [[int(a), int(b)] for a, b in re.findall(r"\((\d+),\s*(\d+)\)", m.group(1))]
    assert len(dz) == 64
    idct = (REF / "h263/src/decoder/cpu/idct.rs").read_text()
    idct = idct[idct.index("const BASIS_TABLE") :]
    m = re.search(r"=\s*\[(.*?)\];", idct, re.S)
    basis = [t for t in re.findall(r"-?\d+\.\d+", m.group(1))]
    assert len(basis) == 64
    deb = strip_comments((REF / "deblock/src/deblock.rs").read_text())
    m = re.search(r"QUANT_TO_STRENGTH[^=]*=\s*\[(.*?)\];", deb, re.S)
    q2s = int_list(m.group(1))
    assert len(q2s) == 32
    return {"dezigzag_xy": dz, "basis_table_f32_literals": basis, "quant_to_strength": q2s}


def main():
    files = {
        "kat_reader.json": extract_reader(),
        "kat_mb_tables.json": extract_mb_tables(),
        "kat_block.json": extract_block(),
        "kat_yuv.json": extract_yuv(),
        "kat_deblock.json": extract_deblock(),
        "kat_constants.json": extract_constants(),
    }
    for name, obj in files.items():
        (OUT / name).write_text(json.dumps(obj, indent=None, separators=(",", ":")) + "\n")
        print(name, "written")


if __name__ == "__main__":
    main()
