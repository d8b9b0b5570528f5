"""Summarise an .ncu-rep (read here with `ncu -i ... --page raw --csv`) into one JSON line per captured launch.

Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_ncu_full.jsonl
HBM GB/s = (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration.sum.
"""
import csv, io, json, subprocess, sys

KEEP = {
    "gpu__time_duration.sum": "us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_peak",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_peak",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pct_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "xbar2sm_MB",
    "smsp__cycles_elapsed.avg": "cycles",
}
UNIT = {"Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "byte": 1e-6, "ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}

def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        rec = {"id": int(d["ID"]), "kernel": d["Kernel Name"].split("(")[0]}
        for k, name in KEEP.items():
            if k in d and d[k] != "":
                v = float(d[k].replace(",", ""))
                u = units[hdr.index(k)]
                if u in UNIT:
                    v *= UNIT[u]
                rec[name] = round(v, 3)
        if "us" in rec and "dram_read_MB" in rec:
            rec["hbm_GBps"] = round((rec["dram_read_MB"] + rec["dram_write_MB"]) / rec["us"] * 1e3, 1)
        print(json.dumps(rec))

if __name__ == "__main__":
    main(sys.argv[1])
