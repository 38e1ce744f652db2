#!/usr/bin/env python
"""DRAM traffic per launch of the kernels of an ncu --set full capture -> profiles/r2_traffic.json (read by bench.py for
`roofline.traffic`), plus a per-kernel key-counter table on stdout.
    python profiles/ncu_traffic.py gpurun_out/r2_path.ncu-rep gpurun_out/r2_windowing.ncu-rep"""
import csv, json, subprocess, sys
from collections import defaultdict
from pathlib import Path

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def to_bytes(val: str, unit: str) -> float:
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    per = defaultdict(list)
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[col["Kernel Name"]].split("(")[0].split("<")[0].split("::")[-1]
            rec = {k: (r[i], units[i]) for k, i in col.items() if any(k.startswith(p) for p in KEYS)}
            rec["_dram"] = to_bytes(*rec["dram__bytes_read.sum"]) + to_bytes(*rec["dram__bytes_write.sum"])
            rec["_read"], rec["_write"] = to_bytes(*rec["dram__bytes_read.sum"]), to_bytes(*rec["dram__bytes_write.sum"])
            per[name].append(rec)
    traffic = {"batch": {}, "windowing": {}}
    for name, recs in per.items():
        last = recs[-1]                                   # the last captured launch (warm)
        print(f"== {name}  ({len(recs)} launches captured)")
        for k, vu in sorted(last.items()):
            if not k.startswith("_"):
                print(f"   {k:84s} {vu[0]:>16s} {vu[1]}")
        print(f"   DRAM read {last['_read'] / 1e6:.1f} MB  write {last['_write'] / 1e6:.1f} MB")
        if name.startswith("tokenizer2_kernel"):
            traffic["batch"]["A/tc/tokenizer"] = int(last["_dram"])
        elif name.startswith("transformer_bf16_kernel"):
            traffic["batch"]["A/tc/transformer"] = int(last["_dram"])
        elif name.startswith("k_gather"):
            traffic["windowing"]["k_gather"] = int(last["_dram"])
        elif name.startswith("score_kernel"):
            traffic["windowing"]["sf_normality_score"] = int(last["_dram"])
        elif name.startswith(("k_flag", "k_scan", "k_compact")):
            traffic["windowing"][name] = int(last["_dram"])
    w = traffic["windowing"]
    if "k_gather" in w:
        w["sf_window_normalize"] = sum(w.get(k, 0) for k in ("k_flag", "k_scan", "k_compact", "k_gather"))
    path = Path(__file__).resolve().parent / "r2_traffic.json"
    json.dump(traffic, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
