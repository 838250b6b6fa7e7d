"""Summarise an `ncu --set full` report (ncu -i REP --page raw --csv > raw.csv) into the JSON kept under profiles/.
usage: python tools/summarize_ncu_full.py raw.csv "source command line" > profiles/rNN_ncu_full_<config>.json"""
import csv
import json
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio"]
STALLS = "smsp__average_warp_latency_issue_stalled_"  # older ncu: smsp__average_warps_issue_stalled_*_per_issue_active.ratio


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return x


def main(path, source=""):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = {"source": source, "units": {k: units[col[k]] for k in KEEP if k in col}, "launches": []}
    stall_cols = [h for h in hdr if "issue_stalled" in h and h.endswith("_per_warp_active.pct") is False and "ratio" in h]
    for r in data:
        e = {"kernel": r[col["Kernel Name"]].split("(")[0]}
        for k in KEEP:
            if k in col:
                e[k] = num(r[col[k]])
        st = {}
        for h in stall_cols:
            v = num(r[col[h]])
            if isinstance(v, float):
                st[h.split("issue_stalled_")[-1].split("_per_")[0]] = round(v, 2)
        e["top_stalls"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:5])
        out["launches"].append(e)
    ones = [e for e in out["launches"] if "k_onesweep<0, 0>" in e["kernel"] or e["kernel"].endswith("k_onesweep<0, 0>")]
    if ones:
        def gb(e, k):
            v, u = e[k], out["units"][k].lower()
            return v * (1e9 if u.startswith("g") else 1e6 if u.startswith("m") else 1e3 if u.startswith("k") else 1)
        big = max(ones, key=lambda e: e["gpu__time_duration.sum"])
        out["k_onesweep_traffic_bytes_per_launch"] = gb(big, "dram__bytes_read.sum") + gb(big, "dram__bytes_write.sum")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:])
