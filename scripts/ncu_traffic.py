"""Summarise one `ncu --set full` capture of the headline bench into profiles/r02_p4_dram_traffic.json (read by bench.py's
roofline.traffic).  Usage:
  ncu -i gpurun_out/r02_p4.ncu-rep --page raw --csv > /tmp/raw.csv
  python scripts/ncu_traffic.py /tmp/raw.csv --samples 100 --chunk 50 --source "<the ncu command>" [--out profiles/...json]
Keeps every launch's duration / DRAM bytes / tensor-pipe activity next to the mean the bench reports."""
import argparse
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--samples", type=int, default=100)
    ap.add_argument("--chunk", type=int, default=50)
    ap.add_argument("--kernel", default="umma_conv")
    ap.add_argument("--source", default="ncu --set full --clock-control none")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_p4_dram_traffic.json"))
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv)))
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}

    def val(r, name, want_unit):
        v = float(r[col[name]].replace(",", ""))
        u = units[col[name]]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "%": 1.0}[u]
        return v * scale

    launches = []
    for r in rows[2:]:
        if a.kernel not in r[col["Kernel Name"]]:
            continue
        launches.append({
            "kernel": r[col["Kernel Name"]].split("(")[0].replace("void <unnamed>::", ""),
            "grid": r[col["launch__grid_size"]],
            "us": round(val(r, "gpu__time_duration.sum", "us"), 2),
            "dram_read_bytes": val(r, "dram__bytes_read.sum", "byte"),
            "dram_write_bytes": val(r, "dram__bytes_write.sum", "byte"),
            "tensor_pipe_pct": float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]),
            "warps_active_pct": float(r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]]),
        })
    tot = sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in launches)
    out = {"samples_per_rank": a.samples, "chunk": a.chunk, "conv_launches": len(launches),
           "dram_bytes_per_launch": tot / max(1, len(launches)), "dram_bytes_per_step": tot,
           "source": "dram__bytes_read.sum + dram__bytes_write.sum, mean over the %d conv launches of one S=%d step (chunks of %d), %s"
                     % (len(launches), a.samples, a.chunk, a.source),
           "launches": launches}
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print(a.out, len(launches), "launches", "%.1f MB per launch" % (out["dram_bytes_per_launch"] / 1e6))


if __name__ == "__main__":
    main()
