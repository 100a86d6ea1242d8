"""Aggregate an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,...` log of tools/ncu_forward.py per
kernel name and write profiles/r02_traffic.json (per-launch DRAM traffic bench.py reports as roofline.traffic).
Usage: python tools/summarise_forward_metrics.py gpurun_out/r02_forward_metrics_b50.csv [batch]"""
import collections, csv, json, re, sys
path, batch = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "?")
rows = list(csv.reader(open(path, errors="ignore")))
hdr, K = None, collections.OrderedDict()
for r in rows:
    if r and r[0] == "ID": hdr = r; continue
    if hdr is None or len(r) != len(hdr) or not r[0].isdigit(): continue
    d = dict(zip(hdr, r))
    k = K.setdefault(int(d["ID"]), {"name": d["Kernel Name"]})
    try: v = float(d["Metric Value"].replace(",", ""))
    except ValueError: v = float("nan")
    u = d["Metric Unit"]
    if d["Metric Name"].startswith("dram__bytes"): v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    if d["Metric Name"] == "gpu__time_duration.sum": v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(u, 1)
    k[d["Metric Name"]] = v
short = lambda n: re.sub(r"\(.*", "", re.sub(r"<unnamed>::", "", re.sub(r"^void ", "", n)))[:110]
agg = collections.OrderedDict()
for k in K.values():
    a = agg.setdefault(short(k["name"]), dict(n=0, us=0.0, rd=0.0, wr=0.0, tp=0.0))
    t = k.get("gpu__time_duration.sum", 0.0)
    a["n"] += 1; a["us"] += t; a["rd"] += k.get("dram__bytes_read.sum", 0.0); a["wr"] += k.get("dram__bytes_write.sum", 0.0)
    a["tp"] += k.get("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * t
tot = sum(a["us"] for a in agg.values())
print("| kernel | launches | total us | share % | DRAM read MB | DRAM write MB | DRAM GB/s | tensor pipe (hmma) % active |\n|---|---|---|---|---|---|---|---|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    if a["us"] / tot < 0.004: continue
    print(f"| `{name}` | {a['n']} | {a['us']:.1f} | {100 * a['us'] / tot:.1f} | {a['rd'] / 1e6:.1f} | {a['wr'] / 1e6:.1f} | {(a['rd'] + a['wr']) / a['us'] / 1e3:.0f} | {a['tp'] / a['us']:.1f} |")
out = {}
for key, name in (("conv1x1_bias_act", "conv1x1_tc_kernel"), ("bias_act", "bias_act_kernel")):
    if name in agg:
        c = agg[name]
        out[key] = {"dram_bytes_per_launch": (c["rd"] + c["wr"]) / c["n"], "dram_read_bytes_per_launch": c["rd"] / c["n"],
                    "dram_write_bytes_per_launch": c["wr"] / c["n"], "launches": c["n"], "avg_us_under_ncu": c["us"] / c["n"],
                    "tensor_pipe_hmma_pct_active": c["tp"] / c["us"],
                    "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the {c['n']} {name} launches of one eager YOLOX-m forward "
                              f"(batch {batch}, the bench default), tools/ncu_forward.py + tools/summarise_forward_metrics.py, profiles/r02_forward_metrics.md"}
json.dump(out, open("profiles/r02_traffic.json", "w"), indent=1)
print(json.dumps(out)[:400], file=sys.stderr)
