#!/usr/bin/env python3
"""profiles/r2_traffic.json (what bench.py reports as roofline.traffic / ncu_same_launch) from a condensed capture.
Usage: python profiles/traffic.py profiles/<capture>.csv <rays per launch> <geodesic frequency> "<how it was taken>" """
import csv
import json
import os
import sys


def main(path, rays, frequency, source):
    m = {}
    with open(path) as f:
        for row in csv.reader(f):
            if len(row) == 3 and row[0] != "metric":
                m[row[0]] = (row[1], row[2])

    def val(name, scale=1.0):
        unit, v = m[name]
        v = float(v)
        if unit == "Gbyte":
            v *= 1e9
        elif unit == "Mbyte":
            v *= 1e6
        return v * scale

    stalls = {k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: float(v[1])
              for k, v in m.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
    top = sorted(((v, k) for k, v in stalls.items() if k != "selected"), reverse=True)[:5]
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    out = {
        "kernel": "trace_persistent<closest>",
        "workload": {"frequency": frequency, "rays_per_launch": rays},
        "dram_bytes_read": round(rd), "dram_bytes_write": round(wr), "dram_bytes_per_launch": round(rd + wr),
        "source": "profiles/%s: %s" % (os.path.basename(path), source),
        "ncu": {
            "gpu_time_ms": round(val("gpu__time_duration.sum"), 2),
            "lts_throughput_pct": round(val("lts__throughput.avg.pct_of_peak_sustained_elapsed"), 1),
            "gpu_dram_throughput_pct": round(val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), 2),
            "lts_sector_hit_rate_pct": round(val("lts__t_sector_hit_rate.pct"), 1),
            "l1tex_sector_hit_rate_pct": round(val("l1tex__t_sector_hit_rate.pct"), 2),
            "issue_active_pct": round(val("smsp__issue_active.avg.pct_of_peak_sustained_active"), 2),
            "pipe_alu_pct": round(val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), 2),
            "pipe_fma_pct": round(val("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), 2),
            "threads_per_warp_instruction": round(val("smsp__thread_inst_executed_per_inst_executed.ratio"), 2),
            "warp_instructions_per_ray": round(val("smsp__inst_executed.sum") / rays, 1),
            "local_loads_per_ray": round(val("smsp__sass_inst_executed_op_local_ld.sum") / rays, 3),
            "top_stalls_per_issue": {k: round(v, 2) for v, k in top},
        },
    }
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "r2_traffic.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=2)
        f.write("\n")
    print(open(dst).read())


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
