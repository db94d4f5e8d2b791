"""Summarise `ncu --page raw --csv` exports into the handful of metrics the roofline needs.

  python tools/ncu_summary.py gpurun_out/r1_*.raw.csv [--json profiles/traffic.json]
Prints one line per profiled launch and (optionally) writes {kernel: dram bytes per launch} for bench.py.
"""
import csv
import json
import sys

KEYS = [
    ("gpu__time_duration.sum", "dur_us", 1e-3),
    ("dram__bytes_read.sum", "dram_rd_MB", 1e-6),
    ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct", 1),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__grid_size", "grid", 1),
    ("launch__block_size", "block", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct", 1),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct", 1),
    ("smsp__cycles_active.avg", "cycles", 1),
]
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "usecond": 1e3, "nsecond": 1, "msecond": 1e6}


def rows(path):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    r = list(csv.reader(lines))
    hdr, units, data = r[0], r[1], r[2:]
    return hdr, units, data


def main():
    args = sys.argv[1:]
    out_json = None
    if "--json" in args:
        i = args.index("--json")
        out_json = args[i + 1]
        del args[i:i + 2]
    traffic = {}
    for path in args:
        hdr, units, data = rows(path)
        idx = {h: i for i, h in enumerate(hdr)}
        for d in data:
            name = d[idx["Kernel Name"]]
            short = name.split("(")[0].split("<")[0].split("::")[-1]
            parts = [f"{path.split('/')[-1].replace('.raw.csv', ''):14s} {short[:34]:34s}"]
            vals = {}
            for key, label, scale in KEYS:
                if key not in idx:
                    continue
                try:
                    v = float(d[idx[key]].replace(",", ""))
                except ValueError:
                    continue
                u = units[idx[key]]
                v *= UNIT_SCALE.get(u, 1)
                vals[label] = v * scale
                parts.append(f"{label}={v * scale:.4g}")
            print(" ".join(parts))
            if "dram_rd_MB" in vals:
                traffic.setdefault(short, int((vals["dram_rd_MB"] + vals.get("dram_wr_MB", 0)) * 1e6))
    if out_json:
        try:
            old = json.load(open(out_json))
        except Exception:
            old = {}
        old.update(traffic)
        json.dump(old, open(out_json, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
