#!/usr/bin/env python3
"""BASELINE.json configs[3] alone (256 x 4CIF, deblock flag, border vectors): recon + deblock/RGBA kernel times.
    python tools/config4_bench.py [n_steps]      (H263CU_LIB selects an experiment build)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from h263_rs_b200 import api, frontend  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
r = bench.config4_deblock(api, frontend, 0, bench.host_threads(), n_steps=n)
print(json.dumps({k: r[k] for k in ("ms_per_step", "recon_ms_per_step", "deblock_rgba_ms_per_step", "value", "bit_exact_vs_oracle_stream0")}))
