"""Turns an `ncu --set full` report into a compact per-launch JSON (the metrics DESIGN.md / bench.py quote).

    ncu -i gpurun_out/x.ncu-rep --page raw --csv | python profiles/summarize_full.py > profiles/r2_x.json
"""
import csv
import json
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "lts__t_sectors.sum": "l2_sectors",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio": "stall_sleeping",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "sm__inst_executed.sum": "warp_instructions",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_pct_of_peak",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
}


def main():
    rows = list(csv.reader(l for l in sys.stdin if not l.startswith("==")))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, r):
            if h == "Kernel Name":
                d["kernel"] = v.split("(")[0].replace("void ", "").replace("efgh::<unnamed>::", "").replace("<unnamed>::", "")
            elif h in KEYS:
                try:
                    x = float(v.replace(",", ""))
                except ValueError:
                    continue
                k = KEYS[h]
                if k == "duration_us":
                    x *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u.replace("second", "s").replace("usecond", "us"), 1.0) if u else 1.0
                    x = x * ({"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(u, 1.0) if u in ("nsecond", "usecond", "msecond", "second") else 1.0)
                if k.endswith("_MB"):
                    x *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                d[k] = round(x, 3)
        out.append(d)
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
