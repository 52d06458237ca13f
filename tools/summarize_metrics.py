"""Summarise an `ncu --metrics ... --csv` per-launch log (long format) into a per-kernel table.

usage: python tools/summarize_metrics.py gpurun_out/step_metrics_v4.csv [top_n]
       python tools/summarize_metrics.py <csv> --traffic-json <precision> <out.json>    # DRAM bytes per GEMM launch for bench.py
"""
import csv, json, os, sys, collections, re

GEMM_KERNELS = ("gemm_bf16_kernel", "gemm_tc_kernel")       # the tensor-core GEMM launches bench.py's roofline counts

def load(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    per = collections.OrderedDict()
    for r in rd:
        k = r["ID"]
        d = per.setdefault(k, {"name": r["Kernel Name"], "grid": r.get("Grid Size", ""), "block": r.get("Block Size", "")})
        v = r["Metric Value"].replace(",", "")
        try: v = float(v)
        except ValueError: pass
        u = r["Metric Unit"]
        if isinstance(v, float):
            if u in ("Kbyte",): v *= 1e3
            elif u in ("Mbyte",): v *= 1e6
            elif u in ("Gbyte",): v *= 1e9
            elif u in ("us", "usecond"): v *= 1e3
            elif u in ("ms", "msecond"): v *= 1e6
            elif u in ("s", "second"): v *= 1e9
        d[r["Metric Name"]] = v
    return list(per.values())

def traffic_json(rows, precision, out, src):
    """profiles/r2_gemm_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of the tensor-core GEMM launches of ONE step."""
    sel = [r for r in rows if any(k in r["name"] for k in GEMM_KERNELS)]
    b = sum(r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0) for r in sel)
    doc = {}
    if os.path.exists(out):
        with open(out) as f:
            doc = json.load(f)
    doc[precision] = {"dram_bytes_per_launch": b / max(1, len(sel)), "launches": len(sel), "dram_bytes_per_step": b,
                      "source": "%s: dram__bytes_read.sum + dram__bytes_write.sum over the %d %s launches of one C2 step (ncu, cold caches, serialised)"
                                % (src, len(sel), " / ".join(GEMM_KERNELS))}
    with open(out, "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc[precision]))


def main():
    if len(sys.argv) > 4 and sys.argv[2] == "--traffic-json":
        return traffic_json(load(sys.argv[1]), sys.argv[3], sys.argv[4], os.path.relpath(sys.argv[1]))
    rows = load(sys.argv[1]); top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    agg = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r["name"]); name = re.sub(r"<.*", "", name).split("::")[-1]
        a = agg.setdefault(name, dict(n=0, t=0.0, rd=0.0, wr=0.0, tens=0.0, smt=0.0, warps=0.0, regs=0))
        t = r.get("gpu__time_duration.sum", 0.0)
        a["n"] += 1; a["t"] += t; a["rd"] += r.get("dram__bytes_read.sum", 0.0); a["wr"] += r.get("dram__bytes_write.sum", 0.0)
        a["tens"] += t * r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
        a["smt"] += t * r.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0.0)
        a["warps"] += t * r.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0)
        a["regs"] = max(a["regs"], int(r.get("launch__registers_per_thread", 0)))
    tot = sum(a["t"] for a in agg.values())
    print(f"{len(rows)} launches, {tot/1e6:.3f} ms summed kernel time (cold-cache, serialised)")
    print(f"{'kernel':38s} {'n':>4s} {'ms':>8s} {'share':>6s} {'dramMB':>9s} {'GB/s':>7s} {'tensor%':>7s} {'sm%':>6s} {'warps%':>6s} {'regs':>4s}")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"])[:top]:
        b = a["rd"] + a["wr"]; t = a["t"]
        print(f"{name[:38]:38s} {a['n']:4d} {t/1e6:8.3f} {100*t/tot:5.1f}% {b/1e6:9.1f} {b/max(t,1):7.1f} {a['tens']/max(t,1):7.1f} {a['smt']/max(t,1):6.1f} {a['warps']/max(t,1):6.1f} {a['regs']:4d}")
    print(f"total dram traffic {sum(a['rd']+a['wr'] for a in agg.values())/1e6:.1f} MB / step")

if __name__ == "__main__":
    main()
