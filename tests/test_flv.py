"""FLV container feed: scan / mux round trips and malformed input, no GPU needed.
The reference has no demuxer (its caller hands it one packet per FLV video tag); these tests pin the
container layout against hand-built bytes and check that the packets reach the parser unchanged."""
import struct

import numpy as np
import pytest

from h263_rs_b200 import _lib, flv, frontend, synth


def hand_tag(tag_type, ts, payload):
    h = bytes([tag_type]) + struct.pack(">I", len(payload))[1:] + struct.pack(">I", ts & 0xFFFFFF)[1:] + bytes([ts >> 24]) + b"\0\0\0"
    return h + payload + struct.pack(">I", 11 + len(payload))


def hand_flv(tags):
    return b"FLV\x01\x05" + struct.pack(">I", 9) + struct.pack(">I", 0) + b"".join(tags)


def test_scan_hand_built_file():
    pic_a, pic_b = b"\x00\x00\x84\x01\x02\x03", b"\x00\x00\x84\xff"
    data = hand_flv([
        hand_tag(18, 0, b"\x02\x00\x0aonMetaData"),           # script
        hand_tag(9, 0, b"\x12" + pic_a),                       # key frame, codec 2
        hand_tag(8, 5, b"\x2f\x00\x01"),                       # audio
        hand_tag(9, 40, b"\x14" + b"vp6 data"),                # video, codec 4 (VP6): skipped
        hand_tag(9, 0x01000028, b"\x22" + pic_b),              # inter frame, extended timestamp
        hand_tag(9, 90, b"\x52\x00"),                          # video info/command frame: skipped
        hand_tag(0x29, 100, b"\x22" + pic_b),                  # filtered (encrypted) video tag: skipped
    ])
    pk, other = flv.scan(data)
    assert len(pk) == 2 and other == 5
    assert bytes(data[pk[0]["offset"] : pk[0]["offset"] + pk[0]["size"]]) == pic_a
    assert bytes(data[pk[1]["offset"] : pk[1]["offset"] + pk[1]["size"]]) == pic_b
    assert list(pk["frame_type"]) == [1, 2] and list(pk["codec_id"]) == [2, 2]
    assert list(pk["timestamp_ms"]) == [0, 0x01000028]
    assert flv.packets(data) == [pic_a, pic_b]


def test_mux_scan_round_trip_and_parse():
    pics = synth.make_stream(176, 144, 6, 11, mv_mode=1)
    for filler in (0, 1, 4):
        data = flv.mux(pics, ms_per_picture=33, filler_every=filler)
        pk, other = flv.scan(data)
        assert len(pk) == len(pics)
        assert other == (0 if filler == 0 else 2 * len(range(0, len(pics), filler)))
        assert flv.packets(data) == pics
        assert list(pk["timestamp_ms"]) == [33 * i for i in range(len(pics))]
        assert list(pk["frame_type"]) == [1] + [2] * (len(pics) - 1)
    # the demuxed packets drive the parser exactly like the raw ones
    a, b = frontend.Parser(1), frontend.Parser(1)
    for raw, dem in zip(pics, flv.packets(flv.mux(pics, filler_every=2))):
        pa, ma, ea = a.parse_picture(raw)
        pb, mb, eb = b.parse_picture(dem)
        assert pa.tobytes() == pb.tobytes() and ma.tobytes() == mb.tobytes() and ea.tobytes() == eb.tobytes()


def test_truncated_and_malformed_input():
    pics = synth.make_stream(176, 144, 3, 5)
    data = flv.mux(pics)
    # a stream cut anywhere yields the complete tags before the cut and never an error
    full, _ = flv.scan(data)
    ends = [int(p["offset"]) + int(p["size"]) for p in full]
    for cut in (len(data) - 1, ends[1] + 3, ends[1] - 1, ends[0], 13, 9):
        pk, _ = flv.scan(data[:cut])
        assert len(pk) == sum(1 for e in ends if e <= cut), cut
    for bad in (b"", b"FL", b"FLX\x01\x05\x00\x00\x00\x09", b"FLV\x01\x05\x00\x00\x00\x05", b"FLV\x01\x05\x00\x00\x01\x00"):
        with pytest.raises(_lib.H263Error) as e:
            flv.scan(bad)
        assert e.value.code == _lib.ERR_INVALID_BITSTREAM
    assert len(flv.scan(b"FLV\x01\x05\x00\x00\x00\x09")[0]) == 0  # header only
    # empty video tag (no codec byte) and zero-length picture are not packets / empty packets
    data = hand_flv([hand_tag(9, 0, b""), hand_tag(9, 0, b"\x12")])
    pk, other = flv.scan(data)
    assert len(pk) == 1 and other == 1 and pk[0]["size"] == 0


def test_frame_types_pass_through():
    pics = [b"\x01\x02", b"\x03", b"\x04\x05\x06"]
    data = flv.mux(pics, frame_types=[1, 3, 2])
    pk, _ = flv.scan(data)
    assert list(pk["frame_type"]) == [1, 3, 2]
    assert flv.packets(data) == pics
