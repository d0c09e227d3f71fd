"""Read the ncu artefacts in gpurun_out/ (launch list csv + .ncu-rep) and write profiles/<round>_* summaries.
usage: python scripts/summarize_profiles.py r02"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size"]


def launches():
    rows = list(csv.reader(open(os.path.join(G, "launches_%s.csv" % R))))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        t = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        name = r[ki].split("(")[0].replace("nnpops::<unnamed>::", "").replace("void ", "")
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
    tot = sum(v[1] for v in agg.values())
    return [{"kernel": k, "launches": v[0], "us": round(v[1], 1), "share": round(v[1] / tot, 4)} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])], tot


def raw(rep):
    out = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(P, rep.replace(".ncu-rep", "_raw.csv").replace("prof_", R + "_").replace("_" + R + "_raw", "_raw")), "w").write(out)
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                d[w] = r[i] if w == "Kernel Name" else (float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else None)
                if w.startswith("dram__bytes") and d[w] is not None:
                    d[w] *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(units[i], 1.0)
                if w == "gpu__time_duration.sum" and d[w] is not None:
                    d[w] *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[i], 1.0)
        d["Kernel Name"] = d["Kernel Name"].split("(")[0].replace("nnpops::<unnamed>::", "").replace("void ", "")
        res.append(d)
    return res


if __name__ == "__main__":
    os.makedirs(P, exist_ok=True)
    ll, tot = launches()
    summary = {"round": R, "launch_list": ll, "launch_list_total_us": round(tot, 1)}
    import shutil
    shutil.copy(os.path.join(G, "launches_%s.csv" % R), os.path.join(P, "%s_launches.csv" % R))
    for key, rep in (("gemm", "prof_gemm_%s.ncu-rep" % R), ("aev", "prof_aev_%s.ncu-rep" % R)):
        if os.path.exists(os.path.join(G, rep)):
            summary[key] = raw(rep)
    if "gemm" in summary:
        g = summary["gemm"]
        summary["gemm_step_totals"] = {
            "launches": len(g), "time_us": round(sum(x["gpu__time_duration.sum"] for x in g), 1),
            "dram_bytes": sum((x["dram__bytes_read.sum"] or 0) + (x["dram__bytes_write.sum"] or 0) for x in g),
            "tensor_pipe_active_pct_time_weighted": round(sum(x["gpu__time_duration.sum"] * x["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] for x in g) /
                                                          sum(x["gpu__time_duration.sum"] for x in g), 2)}
    if "gemm" in summary and any("mlp_chain" in x["Kernel Name"] for x in summary["gemm"]):
        # round 2: ONE fused chain launch per evaluation -- bench.py reads these as the static-capture traffic of its roofline
        c = [x for x in summary["gemm"] if "mlp_chain" in x["Kernel Name"]][0]
        summary["chain_step_totals"] = {
            "launches": 1, "time_us": round(c["gpu__time_duration.sum"], 1),
            "dram_bytes": (c["dram__bytes_read.sum"] or 0) + (c["dram__bytes_write.sum"] or 0),
            "tensor_pipe_active_pct": c.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")}
    json.dump(summary, open(os.path.join(P, "%s_summary.json" % R), "w"), indent=1)
    print(json.dumps({k: v for k, v in summary.items() if k in ("gemm_step_totals", "launch_list_total_us")}, indent=1))
    for x in ll[:12]:
        print(x)
    for key in ("gemm", "aev"):
        for x in summary.get(key, []):
            print(key, x["Kernel Name"][:40], round(x["gpu__time_duration.sum"], 1), "us tensor%", x.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                  "dramMB", round(((x["dram__bytes_read.sum"] or 0) + (x["dram__bytes_write.sum"] or 0)) / 1e6, 1), "issue%", x.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                  "fma%", x.get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"), "xu%", x.get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"))
