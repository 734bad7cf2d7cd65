#!/usr/bin/env python
"""Turn what one `scripts/gpu_suite.sh tag=<tag> ...` call left under gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py <tag> <round-prefix, e.g. r02>

* every bench JSON line            -> profiles/<prefix>_bench_<name>.json (name from the workload / mode / GPUs of the line)
* the ncu launch list              -> profiles/<prefix>_launches_10m_bench.csv (as captured) + <prefix>_kernel_shares.txt
* `ncu --set full` raw pages       -> profiles/<prefix>_ncu_full_<kernel>_metrics.csv (the metrics quoted in DESIGN.md) and
                                      profiles/traffic_<prefix>.json (DRAM bytes per launch: what bench.py quotes as `traffic`)
* `ncu --page source` of bm25      -> profiles/<prefix>_bm25_source_hotspots.txt (stall samples per source line and region)
"""
import csv
import glob
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
           "launch__registers_per_thread", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
           "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct",
           "l1tex__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def to_bytes(value: str, unit: str) -> float:
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value.replace(",", "")) * scale.get(unit, 1.0)


def bench_lines(tag: str, prefix: str):
    for path in sorted(glob.glob(str(OUT / f"{tag}_bench*.json"))):
        lines = [ln for ln in open(path).read().splitlines() if ln.startswith("{")]
        if not lines:
            continue
        d = json.loads(lines[-1])
        wl = d.get("config", {}).get("workload", "")
        if d.get("impl") == "reference":
            name = "reference_arm"
        elif "full-fusion" in wl:
            name = "full_fusion_exhaustive" if "get_scores matrix" in wl else "full_fusion"
        elif "MC-Dropout" in wl:
            name = "c4"
        elif "batch 1 " in wl or "GEMV" in wl:
            name = "c2"
        else:
            passages = wl.split(";")[-1].strip().split(" ")[0] if ";" in wl else ""
            name = {"10000000": "10m", "100000000": "c5_100m"}.get(passages, passages or "bench")
        name += f"_{d.get('n_gpus', 1)}gpu"
        dst = PROF / f"{prefix}_bench_{name}.json"
        dst.write_text(json.dumps(d, indent=1) + "\n")
        print("wrote", dst.name, d.get("value"))


def launches(tag: str, prefix: str):
    paths = sorted(glob.glob(str(OUT / f"{tag}_launches*.csv")))
    if not paths:
        return
    text = open(paths[-1]).read()
    (PROF / f"{prefix}_launches_10m_bench.csv").write_text(text)
    rows = list(csv.reader(text.splitlines()))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv, gs = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = {}
    for r in rows[hi + 1:]:
        if len(r) > mv:
            try:
                v = float(r[mv].replace(",", ""))
            except ValueError:
                continue
            a = agg.setdefault(r[kn].split("(")[0] + " grid " + r[gs], [0, 0.0])
            a[0] += 1
            a[1] += v
    total = sum(a[1] for a in agg.values())
    out = ["kernel (grid) | launches | share of the GPU time inside bench_timed | per launch",
           "(ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include bench_timed/; serialised, cold cache)"]
    for name, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"{name:95s} {a[0]:4d} {100 * a[1] / total:6.2f} % {a[1] / a[0] / 1e3:10.1f} us")
    (PROF / f"{prefix}_kernel_shares.txt").write_text("\n".join(out) + "\n")
    print("\n".join(out[:8]))


def ncu_full(tag: str, prefix: str):
    traffic = {"source": f"profiles/{prefix}_ncu_full_*_metrics.csv (ncu --set full --clock-control none, bench.py --steps 2 --warmup 1, "
                         f"launches inside the NVTX range bench_timed)",
               "workload": {"passages": 10_000_000, "batch": 1024, "k": 10, "pool": 50, "n_gpus": 1, "mode": "pool"}, "kernels": {}}
    for path in sorted(glob.glob(str(OUT / f"{tag}_ncu_*_raw.csv"))):
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        kern = Path(path).name[len(tag) + 5:-8]
        keep = [hdr.index(m) for m in ["Kernel Name"] + METRICS if m in hdr]
        with open(PROF / f"{prefix}_ncu_full_{kern}_metrics.csv", "w", newline="") as fh:
            w = csv.writer(fh)
            w.writerow([hdr[i] for i in keep])
            w.writerow([units[i] for i in keep])
            for r in rows[2:]:
                w.writerow([r[i] for i in keep])
        # the launch with the longest duration is the one quoted (the dense kernel runs twice per step)
        t = hdr.index("gpu__time_duration.sum")
        best = max(rows[2:], key=lambda r: float(r[t].replace(",", "")))
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        traffic["kernels"][kern] = {"dram_bytes_read": to_bytes(best[rd], units[rd]), "dram_bytes_write": to_bytes(best[wr], units[wr]),
                                    "gpu_time_ms_under_ncu": float(best[t].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[t], 1.0),
                                    "kernel": best[hdr.index("Kernel Name")]}
        print(kern, traffic["kernels"][kern])
    if traffic["kernels"]:
        (PROF / f"traffic_{prefix}.json").write_text(json.dumps(traffic, indent=1) + "\n")


def bm25_source(tag: str, prefix: str):
    rep = OUT / f"{tag}_ncu_bm25_kernel.ncu-rep"
    if not rep.exists():
        return
    res = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(res.stdout.splitlines()))
    fileof, data = None, []
    for r in rows:
        if r and r[0] == "File Path":
            fileof = r[1].split("/")[-1]
            continue
        if len(r) > 8 and r[2] == "-" and r[0].isdigit():
            try:
                data.append((fileof, int(r[0]), r[1], int(r[4]), int(r[7])))
            except ValueError:
                pass
    total = sum(d[3] for d in data) or 1
    out = [f"bm25_kernel<false>, 10M passages x 1024 queries, pool 50: warp-stall samples per CUDA source line (ncu --set full --import-source on, {total} samples)",
           "file line share-of-samples source"]
    for f, line, src, smp, _ in sorted(data, key=lambda d: -d[3])[:40]:
        out.append(f"{f:24s} {line:5d} {100 * smp / total:5.1f} %  {src[:110]}")
    (PROF / f"{prefix}_bm25_source_hotspots.txt").write_text("\n".join(out) + "\n")
    print("\n".join(out[:12]))


if __name__ == "__main__":
    tag, prefix = sys.argv[1], sys.argv[2]
    PROF.mkdir(exist_ok=True)
    bench_lines(tag, prefix)
    launches(tag, prefix)
    ncu_full(tag, prefix)
    bm25_source(tag, prefix)
