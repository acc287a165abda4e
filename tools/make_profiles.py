"""Turns the round-end GPU pass (tools/final_gpu_run.sh -> gpurun_out/<R>_*) into the committed profiles/ files.
usage: python tools/make_profiles.py r02"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

for f in ("bench.json", "bench_reference.json", "bench_g1.json", "bench_three_humanoids.json", "bench_convex_mesh.json", "bench_2gpu.json", "gpu_tests.log", "launches.csv"):
  src = os.path.join(G, f"{R}_{f}")
  if os.path.exists(src):
    shutil.copyfile(src, os.path.join(P, f"{R}_{f}"))

# ---- launch list of the bench command
rows = [r for r in csv.reader(open(os.path.join(G, f"{R}_launches.csv"))) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
  a = agg.setdefault(r[4], [0, 0.0])
  a[0] += 1
  a[1] += float(r[14]) / 1e3
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, f"{R}_launches.md"), "w") as f:
  f.write(f"# ncu launch list of `bench.py --no-graph --steps 8 --warmup 3 --no-cpu` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)\n\n")
  f.write("The step is pipelined over two halves of the 8192 worlds, so every step kernel appears twice per step.\n\n| kernel | launches | mean us | share |\n|---|---|---|---|\n")
  for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    f.write(f"| {k[:90]} | {n} | {t / n:.1f} | {t / tot:.3f} |\n")

# ---- per-kernel ncu --set full summary + source hot spots
raw = list(csv.reader(open(os.path.join(G, f"{R}_raw.csv"))))
hdr, idx = raw[0], {h: i for i, h in enumerate(raw[0])}
cols = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "l1tex__t_sector_hit_rate.pct"]
st = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")]
traffic = {}
with open(os.path.join(P, f"{R}_kernels.md"), "w") as f:
  f.write(f"# ncu --set full --clock-control none, {R} kernels (humanoid, 8192 worlds, 1xB200, MJB_SPLIT=1 so each kernel covers all worlds; step 25 from the squat keyframe)\n\n")
  f.write("Units as ncu reports them (time us, DRAM Mbyte, shared memory Kbyte/block).  `thread_inst / inst` = average active lanes per warp instruction.\n\n")
  f.write("| kernel | " + " | ".join(c.split("__")[-1].replace(".sum", "").replace(".avg.pct_of_peak_sustained_", " % ").replace(".ratio", "")[:34] for c in cols) + " | top stalls (warps per issue-active cycle) |\n")
  f.write("|---|" + "---|" * (len(cols) + 1) + "\n")
  for r in raw[2:]:
    name = r[idx["Kernel Name"]].replace("void <unnamed>::", "").replace("<unnamed>::", "").split("(")[0]
    vals = sorted([(float(r[idx[h]]), h.split("stalled_")[1].split("_per")[0]) for h in st], reverse=True)[:4]
    f.write(f"| {name} | " + " | ".join(f"{float(r[idx[c]]):.4g}" if c in idx and r[idx[c]] else "" for c in cols) + " | " + ", ".join(f"{n} {v:.2f}" for v, n in vals) + " |\n")
    key = "k_" + name.split("<")[0].replace("k_", "").replace("euler_flat", "euler")
    traffic[key] = (float(r[idx["dram__bytes_read.sum"]]) + float(r[idx["dram__bytes_write.sum"]])) * 1e6
  f.write("\n")
  out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, f"{R}_raw.csv"), os.path.join(G, f"{R}_src.csv"), "14"], capture_output=True, text=True).stdout
  sect = [l for l in out.splitlines() if not l.startswith("   ") and not l.startswith("== ")]
  f.write("## top source lines by stall samples (per kernel and source file: % of the file's instructions, % of its stall samples, stall classes)\n\n```\n" + "\n".join(l[:200] for l in sect) + "\n```\n")
json.dump({"source": f"ncu --set full, profiles/{R}_kernels.md (dram__bytes_read.sum + dram__bytes_write.sum per launch, bytes; 8192 worlds)", "traffic_bytes": traffic},
          open(os.path.join(P, f"{R}_traffic.json"), "w"), indent=1)

# ---- SASS evidence of the bulk-async (TMA) staging
so = os.path.join(ROOT, "mujoco_warp_b200", "libmjb200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn, counts = None, collections.OrderedDict()
for line in sass.splitlines():
  if "Function :" in line:
    fn = line.split("Function :")[1].strip()
  for mn in ("UBLKCP", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK", "FENCE.VIEW.ASYNC", "UTMALDG", "LDGSTS"):
    if fn and mn in line:
      counts.setdefault(fn, collections.Counter())[mn] += 1
with open(os.path.join(P, f"{R}_sass_tma.md"), "w") as f:
  f.write("# Bulk-async (1-D TMA) staging in the shipped SASS (`cuobjdump -sass mujoco_warp_b200/libmjb200.so`)\n\n")
  f.write("`cp.async.bulk` appears as `UBLKCP` (global -> shared with `.S.G`, shared -> global with `.G.S`), the mbarrier transaction count as `SYNCS.ARRIVE.TRANS64`, the wait as\n`SYNCS.PHASECHK.TRANS64.TRYWAIT`, the generic -> async proxy fence as `FENCE.VIEW.ASYNC` (B200_PROFILING.md, B300_MICROARCH.md).  Instruction counts per kernel:\n\n| kernel | UBLKCP | SYNCS.ARRIVE.TRANS64 | SYNCS.PHASECHK | FENCE.VIEW.ASYNC |\n|---|---|---|---|---|\n")
  for k, c in counts.items():
    if c["UBLKCP"]:
      f.write(f"| `{k[:110]}` | {c['UBLKCP']} | {c['SYNCS.ARRIVE.TRANS64']} | {c['SYNCS.PHASECHK']} | {c['FENCE.VIEW.ASYNC']} |\n")
  ex = [l for l in sass.splitlines() if "UBLKCP" in l][:6]
  f.write("\nexcerpt:\n```\n" + "\n".join(l.rstrip()[:150] for l in ex) + "\n```\n")
print("profiles written:", sorted(x for x in os.listdir(P) if x.startswith(R)))
