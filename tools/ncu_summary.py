#!/usr/bin/env python3
"""Condense an `ncu --set full` report into the JSON summary committed under profiles/.

    python tools/ncu_summary.py gpurun_out/recon_v1.ncu-rep profiles/r01_v1_recon_ncu.json [--latest]

With --latest the summary is also written to profiles/recon_ncu_summary.json, the file
bench.py reads `roofline.traffic` (dram bytes per launch) from.  Runs here, without a GPU.
"""
import csv
import json
import os
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_mb",
    "dram__bytes_write.sum": "dram_write_mb",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_data_pipe_wavefronts_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "l1_wavefronts_shared",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum": "l1_wavefronts_global_ld",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum": "l1_wavefronts_global_st",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_instruction",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__cycles_elapsed.max": "sm_cycles",
}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    launches = []
    for r in data:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k, name in KEYS.items():
            if k in hdr:
                v = r[hdr.index(k)].replace(",", "")
                try:
                    d[name] = float(v)
                except ValueError:
                    d[name] = v
                if name in ("dram_read_mb", "dram_write_mb"):
                    u = units[hdr.index(k)]
                    scale = {"Mbyte": 1.0, "Gbyte": 1e3, "Kbyte": 1e-3, "byte": 1e-6}.get(u, 1.0)
                    d[name] = d[name] * scale
                if name == "duration_us":
                    u = units[hdr.index(k)]
                    d[name] = d[name] * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
        launches.append(d)
    n = max(len(launches), 1)
    dram = sum((l.get("dram_read_mb", 0) + l.get("dram_write_mb", 0)) for l in launches) / n * 1e6
    summary = {
        "report": os.path.basename(rep),
        "command": "ncu --set full --clock-control none --import-source on -k regex:recon_ -s 5 -c 2 python bench.py --steps 4 --warmup 3 --skip-extras",
        "note": "per-launch values; ncu serialises and replays, so durations are not bench values",
        "launches": launches,
        "dram_bytes_per_launch": dram,
    }
    with open(out, "w") as f:
        json.dump(summary, f, indent=1)
    if "--latest" in sys.argv:
        with open(os.path.join(os.path.dirname(out), "recon_ncu_summary.json"), "w") as f:
            json.dump(summary, f, indent=1)
    print(json.dumps(summary, indent=1)[:1500])


if __name__ == "__main__":
    main()
