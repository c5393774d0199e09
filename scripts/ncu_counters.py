"""ncu report (raw page CSV) of the bench workload -> profiles/r02_counters.json: per kernel, per launch of 64 hypotheses,
the warp instructions executed and the DRAM bytes moved, tagged with the sha of the kernel sources they were measured on.
Usage (GPU box): ncu -i gpurun_out/X.ncu-rep --page raw --csv > gpurun_out/X_raw.csv ; python scripts/ncu_counters.py gpurun_out/X_raw.csv out.json"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    src, dst = sys.argv[1], sys.argv[2]
    import bench

    rows = list(csv.reader(open(src)))
    head = rows[0]
    col = {n: i for i, n in enumerate(head)}
    units = rows[1]
    kernels = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        key = "pixel_kernel" if "pixel_kernel" in name else "raster_kernel" if "raster_kernel" in name else "iter_kernel" if "iter_kernel" in name else "bin_kernel" if "bin_kernel" in name else None
        if key is None:
            continue

        def val(metric, scale_units=True):
            i = col[metric]
            v = float(r[i].replace(",", ""))
            u = units[i].lower()
            if scale_units:
                for pre, m in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("msecond", 1e3), ("usecond", 1.0), ("nsecond", 1e-3), ("second", 1e6)):
                    if u.startswith(pre):
                        return v * m
            return v

        rec = {"kernel_name": name, "inst_executed": val("smsp__inst_executed.sum", False), "dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
               "duration_us_under_ncu": val("gpu__time_duration.sum"), "registers": val("launch__registers_per_thread", False),
               "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active", False), "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active", False)}
        kernels.setdefault(key, []).append(rec)
    out = {"source_sha": bench.source_sha(), "workload": "bench workload (config 2), 64 hypotheses per launch, DDOPE_PARTS=1", "kernels": {}}
    for k, recs in kernels.items():
        n = len(recs)
        out["kernels"][k] = {m: sum(r[m] for r in recs) / n for m in recs[0] if m != "kernel_name"}
        out["kernels"][k]["kernel_name"] = recs[0]["kernel_name"]
        out["kernels"][k]["launches_averaged"] = n
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
