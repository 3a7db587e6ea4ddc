#!/usr/bin/env python
"""Turn gpurun_out/ ncu outputs into the committed summaries under profiles/ (run here, no GPU needed)."""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
lines = [l for l in open(os.path.join(G, "launches.csv")) if not l.startswith("==")]
rows = list(csv.reader(lines)); hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi: continue
    n = r[ki]
    n = ("apa::" + n[n.index("apa_"):].split("(")[0].split("<")[0]) if "apa_" in n else "torch: " + n.split("<")[0][-60:]
    try: v = float(r[vi])
    except ValueError: continue
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, tag + "_launches_summary.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 120 --csv :\n")
    f.write("#   python bench.py --steps 30 --warmup 3 --no-graph --no-pdl --streams 1 --stream-ctas 0 (library default: 148 STREAM CTAs)\n")
    f.write("# (serialised, cold-cache launch times under the profiler: compare SHARES, not absolutes)\n")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write("%6d launches  %10.1f us total  %5.1f%%  avg %8.2f us  %s\n" % (c, t / 1e3, 100 * t / tot, t / 1e3 / c, n))
print(open(os.path.join(P, tag + "_launches_summary.txt")).read())
out = subprocess.run("ncu -i %s --page raw --csv" % os.path.join(G, "prof_stream.ncu-rep"), shell=True, capture_output=True, text=True).stdout
open(os.path.join(P, tag + "_stream_ncu_raw.csv"), "w").write(out)
rows = list(csv.reader(out.splitlines())); hdr = rows[0]; units = rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_dim_x", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
summ = []
for r in rows[2:]:
    summ.append({w: (r[hdr.index(w)] + " " + units[hdr.index(w)]).strip() for w in want if w in hdr})
json.dump(summ, open(os.path.join(P, tag + "_stream_ncu_summary.json"), "w"), indent=1)
for k, v in summ[0].items(): print("  ", k, v[:90])
def num(s):
    v, u = s.split(" ", 1); return float(v) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(u.strip(), 1)
tr = [num(d["dram__bytes_read.sum"]) + num(d["dram__bytes_write.sum"]) for d in summ]
json.dump({"dram_bytes_per_launch": sum(tr) / len(tr),
           "source": "profiles/%s_stream_ncu_raw.csv (apa_flow_stream_kernel, ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum, mean of %d launches)" % (tag, len(tr))},
          open(os.path.join(P, "traffic.json"), "w"), indent=1)
for src, dst in (("bench_default.log", "_bench_1gpu.json"), ("bench_streams1.log", "_bench_1gpu_streams1.json"), ("bench_reference.log", "_bench_reference_arm.json")):
    p = os.path.join(G, src)
    if os.path.exists(p):
        l = [x for x in open(p) if x.startswith("{")]
        if l: open(os.path.join(P, tag + dst), "w").write(l[-1])
for src, dst in (("timeline_local_solo.log", "_local_timeline_solo.txt"), ("timeline_local_steady.log", "_local_timeline_steady.txt"), ("bench_dense.log", "_dense_configs.txt")):
    p = os.path.join(G, src)
    if os.path.exists(p): shutil.copy(p, os.path.join(P, tag + dst))
print(open(os.path.join(P, "traffic.json")).read())
