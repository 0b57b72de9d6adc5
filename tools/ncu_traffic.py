"""Turns an ncu CSV (metrics gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum) of
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv \
        --log-file launches.csv python bench.py --clips 1 --streams 1 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline
into profiles/ncu_traffic_r1.json + a readable per-kernel summary.  Only the LAST decoder call of the capture is kept
(a call starts with the three layout launches of the multi-scale features).
  python tools/ncu_traffic.py launches.csv profiles/ncu_traffic_r1.json profiles/launches_summary.txt"""
import csv, collections, json, sys
src, out_json, out_txt = sys.argv[1:4]
title = sys.argv[4] if len(sys.argv) > 4 else "one decoder call, clips=1, cfg 2 (36x736x1280, Q=100)"
lines = [l for l in open(src) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
per = collections.OrderedDict()
for r in rows:
    k = (r["ID"], r["Kernel Name"].split("(")[0])
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    m = r["Metric Name"]
    if m.startswith("gpu__time"):
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)          # -> us
    else:
        v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    per.setdefault(k, {})[m] = v
# keep the last decoder call: it starts at the first of the last three consecutive token-layout launches
keys = list(per.keys())
starts = [i for i, k in enumerate(keys) if "tokens_prep" in k[1] or "nchw_to_tokens" in k[1]]
if len(starts) >= 3:
    first = starts[-3]
    per = collections.OrderedDict((k, per[k]) for k in keys[first:])
def fam(n, m=None):
    if "xattn" in n: return "xattn"
    if "maskfeat" in n or "nchw" in n or "tokens_prep" in n: return "prep"
    if "gemm_tn_bs_kernel<256>" in n: return "kv_proj"
    if "gemm_tn_bs_kernel<128>" in n:       # one launch writes the full-resolution logits, the others write bits
        return "mask_logits" if (m or {}).get("dram__bytes_write.sum", 0) > 1e8 else "mask_bits"
    if "match_" in n or "reorder_queries" in n: return "query_matching"
    if "san_attn" in n or "san_pool" in n: return "clip_side_path_attention"
    if "postprocess" in n or "topk" in n: return "post_processing"
    if "gemm" in n or "ln_reduce" in n or "self_attn" in n or "unfold" in n: return "query_side"
    return "other"
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
byname = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for (i, n), m in per.items():
    for d in (agg[fam(n, m)], byname[n]):
        d[0] += 1; d[1] += m.get("gpu__time_duration.sum", 0); d[2] += m.get("dram__bytes_read.sum", 0); d[3] += m.get("dram__bytes_write.sum", 0)
js = {f: {"launches": v[0], "us": v[1], "dram_bytes_per_clip": v[2] + v[3], "read": v[2], "write": v[3]} for f, v in agg.items()}
json.dump(js, open(out_json, "w"), indent=1)
with open(out_txt, "w") as f:
    f.write(f"# {src}: {title}; ncu times are cold-cache & serialised\n")
    f.write(f"{'kernel':60s} {'n':>4s} {'us':>10s} {'dram_read_MB':>13s} {'dram_write_MB':>14s}\n")
    for n, v in sorted(byname.items(), key=lambda x: -x[1][1]):
        f.write(f"{n[:60]:60s} {v[0]:4d} {v[1]:10.1f} {v[2]/1e6:13.1f} {v[3]/1e6:14.1f}\n")
print(open(out_txt).read())
