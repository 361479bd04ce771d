"""profiles/r2_ncu_summary.json from the raw ncu CSVs of scripts/gpu_ncu.sh (metrics pass over every launch of one warm frame, per tile-shard
denominator N).  Per workload and N: DRAM bytes of the frame's k_trace launches (roofline.traffic), their share of the frame's device time,
lanes per instruction, ALU-pipe / issue / L1TEX utilisation (time-weighted over the k_trace launches).
    python scripts/ncu_summary.py gpurun_out/r2_ncu_metrics_c3_n{1,2,4,8}.csv > profiles/r2_ncu_summary.json"""
import csv, json, re, sys

CMD = ("ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
       "smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,"
       "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,"
       "l1tex__throughput.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct "
       "--csv --log-file gpurun_out/r2_ncu_metrics_<wl>_n<N>.csv python scripts/ncu_frame.py <wl> <N>   (scripts/gpu_ncu.sh)")


def to_float(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "nsecond": 1e-6, "ns": 1e-6, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0, "second": 1e3}
    return v * scale.get(u, 1.0)


def load(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append(r)
    launches = {}
    for r in rows:
        k = int(r["ID"])
        d = launches.setdefault(k, {"kernel": re.sub(r"^void ", "", r["Kernel Name"])})
        d[r["Metric Name"]] = to_float(r["Metric Value"], r["Metric Unit"])
    return [launches[k] for k in sorted(launches)]


out = {"command": CMD}
for path in sys.argv[1:]:
    m = re.search(r"metrics_(\w+?)_n(\d+)\.csv", path)
    wl, n = m.group(1), m.group(2)
    L = load(path)
    total_ms = sum(l["gpu__time_duration.sum"] for l in L)
    tr = [l for l in L if l["kernel"].startswith("k_trace") and l["gpu__time_duration.sum"] > 0.005]
    t_ms = sum(l["gpu__time_duration.sum"] for l in tr)
    w = lambda key: sum(l[key] * l["gpu__time_duration.sum"] for l in tr) / max(t_ms, 1e-9)
    per_kernel = {}
    for l in L:
        name = re.sub(r"[<(].*", "", l["kernel"])
        e = per_kernel.setdefault(name, {"launches": 0, "ms": 0.0, "dram_gb": 0.0})
        e["launches"] += 1; e["ms"] += l["gpu__time_duration.sum"]; e["dram_gb"] += (l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"]) / 1e9
    for e in per_kernel.values():
        e["ms"] = round(e["ms"], 3); e["dram_gb"] = round(e["dram_gb"], 3); e["share"] = round(e["ms"] / total_ms, 4)
    out.setdefault(wl, {})[n] = {
        "csv": path.replace("gpurun_out/", "profiles/"), "launches_in_frame": len(L), "frame_ms_under_ncu": round(total_ms, 3),
        "k_trace_launches": len(tr), "k_trace_ms_under_ncu": round(t_ms, 3), "k_trace_share_of_frame": round(t_ms / total_ms, 4),
        "dram_bytes_per_step": sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in tr),
        "dram_read_gb": round(sum(l["dram__bytes_read.sum"] for l in tr) / 1e9, 3), "dram_write_gb": round(sum(l["dram__bytes_write.sum"] for l in tr) / 1e9, 3),
        "lanes_per_inst": round(w("smsp__thread_inst_executed_per_inst_executed.ratio"), 2),
        "alu_pipe_pct": round(w("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), 1),
        "fma_pipe_pct": round(w("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), 1),
        "issue_active_pct": round(w("smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
        "l1tex_throughput_pct": round(w("l1tex__throughput.avg.pct_of_peak_sustained_active"), 1),
        "l1_hit_pct": round(w("l1tex__t_sector_hit_rate.pct"), 1), "l2_hit_pct": round(w("lts__t_sector_hit_rate.pct"), 1),
        "per_kernel": per_kernel,
    }
print(json.dumps(out, indent=1))
