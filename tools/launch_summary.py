"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`)
into markdown + a per-kernel DRAM-traffic table (JSON) that bench.py reads for `roofline.traffic`.
Usage: python tools/launch_summary.py gpurun_out/launches.csv profiles/rNN_ncu_launches.md profiles/ncu_traffic.json"""
import collections
import csv
import json
import re
import sys

FAMILY = [  # kernel-name regex -> C-ABI family used in bench.py's kernel table
    (r"gemm_tn(_ts)?_kernel", "mmfm_gemm_tn"), (r"gemm_wgrad_kernel", "mmfm_gemm_wgrad"),
    (r"attn_fwd", "mmfm_attention_fwd"), (r"attn_bwd", "mmfm_attention_bwd"),
    (r"layernorm_fwd", "mmfm_layernorm_fwd"), (r"layernorm_bwd", "mmfm_layernorm_bwd"),
    (r"loss_kernel", "mmfm_loss_fwd_bwd"), (r"adamw", "mmfm_adamw_step"),
]


def main(src, md, js):
    lines = [ln for ln in open(src, errors="ignore") if ln.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.defaultdict(lambda: {"n": set(), "ns": 0.0, "rd": 0.0, "wr": 0.0})
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        name = r[ix["Kernel Name"]]
        metric, unit, val = r[ix["Metric Name"]], r[ix["Metric Unit"]], float(r[ix["Metric Value"]].replace(",", ""))
        d = per[name]
        d["n"].add(r[ix["ID"]])
        scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        if metric.startswith("gpu__time"):
            d["ns"] += val * scale
        elif "bytes_read" in metric:
            d["rd"] += val * scale
        elif "bytes_write" in metric:
            d["wr"] += val * scale
    tot = sum(d["ns"] for d in per.values())
    n_l = sum(len(d["n"]) for d in per.values())
    with open(md, "w") as f:
        f.write(f"# ncu launch list (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                f"--clock-control none ... python bench.py --steps 2 --warmup 3 --no-cpu-baseline`)\n\n")
        f.write(f"{n_l} launches captured (about one forward+backward step at B=256, N=668), {tot / 1e6:.2f} ms of kernel "
                f"time (cold-cache, serialised: compare shares).\n\n")
        f.write("| kernel | launches | total ms | share | avg us | DRAM read MB/launch | DRAM write MB/launch |\n|---|---|---|---|---|---|---|\n")
        for name, d in sorted(per.items(), key=lambda kv: -kv[1]["ns"]):
            n = len(d["n"])
            short = re.sub(r"\(.*", "", name)[:80]
            f.write(f"| `{short}` | {n} | {d['ns'] / 1e6:.3f} | {100 * d['ns'] / tot:.1f}% | {d['ns'] / n / 1e3:.1f} | "
                    f"{d['rd'] / n / 1e6:.2f} | {d['wr'] / n / 1e6:.2f} |\n")
    fam = collections.defaultdict(lambda: {"launches": 0, "bytes": 0.0, "ns": 0.0})
    for name, d in per.items():
        for rx, key in FAMILY:
            if re.search(rx, name):
                fam[key]["launches"] += len(d["n"])
                fam[key]["bytes"] += d["rd"] + d["wr"]
                fam[key]["ns"] += d["ns"]
                break
    out = {k: {"dram_bytes_per_launch": v["bytes"] / v["launches"], "launches": v["launches"],
               "share_of_captured_time": v["ns"] / tot} for k, v in fam.items()}
    json.dump({"source": src, "note": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the "
               "launches of one captured step", "kernels": out}, open(js, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:4])
