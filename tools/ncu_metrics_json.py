"""`ncu --page raw --csv` dump of tools/batch_step.py  ->  per-stage metrics of ONE multi-view batch as JSON.

    ncu -i capture.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_metrics_json.py raw.csv profiles/r02_stage_metrics_C3_batch8.json [--batch -1] [--note "..."]

bench.py reads the result for the `roofline` / `kernels` entries of its JSON line: warp instructions executed
(-> issue-rate fraction) and DRAM bytes (-> `traffic`) per stage.  A batch starts at a `preprocess_kernel` launch;
radix passes are attributed to the depth sort before the emission kernels and to the tile sort after them.
"""
import argparse
import csv
import json

STAGES = ["preprocess", "depth_sort", "emit", "tile_sort", "ranges", "blend_fwd", "blend_bwd", "preprocess_bwd"]


def stage_of(name, seen_emit):
    if "preprocess_bwd" in name:
        return "preprocess_bwd"
    if "preprocess_kernel" in name:
        return "preprocess"
    if "radix" in name or "sort_" in name:
        return "tile_sort" if seen_emit else "depth_sort"
    if "emit" in name:
        return "emit"
    if "ranges" in name:
        return "ranges"
    if "tile_order" in name or "blend_fwd" in name:
        return "blend_fwd"
    if "unit_build" in name or "blend_bwd" in name:
        return "blend_bwd"
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw_csv")
    ap.add_argument("out_json")
    ap.add_argument("--batch", type=int, default=-1, help="which batch of the capture (default: the last complete one)")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.raw_csv)))
    hdr, units = rows[0], rows[1]
    col = {n: hdr.index(n) for n in ("Kernel Name", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                     "gpu__time_duration.sum", "launch__registers_per_thread",
                                     "smsp__issue_active.avg.pct_of_peak_sustained_active",
                                     "smsp__thread_inst_executed_per_inst_executed.ratio") if n in hdr}

    def num(r, n, scale_by_unit=False):
        if n not in col or r[col[n]] in ("", "n/a"):
            return None
        v = float(r[col[n]].replace(",", ""))
        if scale_by_unit:
            u = units[col[n]].lower()
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        return v

    batches, cur = [], None
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if "preprocess_kernel" in name:
            cur = {"seen_emit": False, "stages": {}}
            batches.append(cur)
        if cur is None:
            continue
        st = stage_of(name, cur["seen_emit"])
        if "emit" in name:
            cur["seen_emit"] = True
        if st is None:
            continue
        e = cur["stages"].setdefault(st, {"kernels": [], "warp_instructions": 0.0, "dram_bytes": 0.0, "ncu_ms": 0.0})
        u_t = units[col["gpu__time_duration.sum"]].lower()
        t = num(r, "gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u_t, 1.0)
        e["kernels"].append({"name": name.split("(")[0].replace("void ", "").replace("tgr::", ""), "ncu_ms": t,
                             "warp_instructions": num(r, "smsp__inst_executed.sum"),
                             "dram_bytes": (num(r, "dram__bytes_read.sum", True) or 0) + (num(r, "dram__bytes_write.sum", True) or 0),
                             "registers": num(r, "launch__registers_per_thread"),
                             "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                             "lanes_per_instruction": num(r, "smsp__thread_inst_executed_per_inst_executed.ratio")})
        e["warp_instructions"] += num(r, "smsp__inst_executed.sum") or 0
        e["dram_bytes"] += e["kernels"][-1]["dram_bytes"]
        e["ncu_ms"] += t
    full = [b for b in batches if "preprocess_bwd" in b["stages"]] or batches
    b = full[a.batch]
    out = {"source": a.raw_csv, "note": a.note, "stages": b["stages"]}
    json.dump(out, open(a.out_json, "w"), indent=1)
    for s in STAGES:
        if s in b["stages"]:
            e = b["stages"][s]
            print("%-15s %8.3f ms  %12.0f warp-inst  %8.1f MB dram  (%d launches)" % (s, e["ncu_ms"], e["warp_instructions"],
                                                                                 e["dram_bytes"] / 1e6, len(e["kernels"])))


if __name__ == "__main__":
    main()
