#!/usr/bin/env python3
"""Writes a synthetic stream in the packet-file format of the Rust cross-check and the oracle's
hash line per picture (same format as src/main.rs prints).  Runs here (CPU only).

    python oracle/rust_harness/dump_stream.py out.h263pk out_hashes.txt [width height pictures seed]
"""
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def fnv(data):
    h = 0xCBF29CE484222325
    for b in bytes(data):
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def main():
    import oracle_lib as O  # noqa: F401  (the oracle: test infrastructure)
    from helpers import oracle_decode_stream
    from h263_rs_b200 import synth

    out, hashes = sys.argv[1], sys.argv[2]
    w, h, n, seed = (int(a) for a in (sys.argv[3:7] + ["176", "144", "30", "1"][len(sys.argv) - 3:]))
    packets = synth.make_stream(w, h, n, seed, mv_mode=2, pct_escape=10, pct_fourmv=10)
    with open(out, "wb") as f:
        for p in packets:
            f.write(struct.pack("<I", len(p)))
            f.write(p)
    plain = oracle_decode_stream(packets)
    deb = oracle_decode_stream(packets, deblock=True)
    with open(hashes, "w") as f:
        for i, (a, b) in enumerate(zip(plain, deb)):
            if isinstance(a, int):
                f.write("%d error %d\n" % (i, a))
            else:
                f.write("%d %016x %016x %016x %016x %016x\n" % (i, fnv(a["y"]), fnv(a["cb"]), fnv(a["cr"]), fnv(a["rgba"]), fnv(b["rgba"])))


if __name__ == "__main__":
    main()
