"""Summarise an `ncu --set full` capture exported with --page raw / --page source (see tools/final_gpu_run.sh).
usage: python tools/ncu_summary.py <raw.csv> <src.csv> [topN]"""
import collections
import csv
import sys

raw, src = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
rows = list(csv.reader(open(raw)))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "smsp__inst_executed_op_global_ld.sum", "sm__inst_executed_pipe_lsu.sum", "sm__cycles_active.avg"]
st = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")]
for r in rows[2:]:
  print("==", r[idx["Kernel Name"]][:70])
  for w in want:
    if w in idx:
      print(f"   {w}: {r[idx[w]]}")
  vals = sorted([(round(float(r[idx[h]]), 2), h.split("stalled_")[1].split("_per")[0]) for h in st], reverse=True)[:7]
  print("   stalls:", vals)
rows = list(csv.reader(open(src)))
sections, cur = [], None
for r in rows:
  if len(r) >= 2 and r[0] == "Function Name":
    cur = {"name": r[1], "rows": [], "hdr": None}
    sections.append(cur)
  elif cur is not None and len(r) > 8 and r[0] == "Line No":
    cur["hdr"] = r
  elif cur is not None and cur["hdr"] is not None and r and r[0].isdigit():
    cur["rows"].append(r)
tot_i = sum(int(r[s["hdr"].index("Instructions Executed")] or 0) for s in sections if s["hdr"] for r in s["rows"] if r[s["hdr"].index("Instructions Executed")].isdigit()) or 1
for s in sections:
  h = s["hdr"]
  if h is None or len(s["rows"]) < 12:
    continue
  ii, wi = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
  cols = [h.index(n) for n in ("stall_long_sb", "stall_no_inst", "stall_short_sb", "stall_wait", "stall_mio")]
  agg = collections.OrderedDict()
  for r in s["rows"]:
    try:
      key = (int(r[0]), r[1].strip()[:105])
    except ValueError:
      continue
    a = agg.setdefault(key, [0] * 7)
    for k, c in enumerate([ii, wi] + cols):
      try:
        a[k] += int(r[c] or 0)
      except ValueError:
        pass
  ti = sum(a[0] for a in agg.values()) or 1
  ts = sum(a[1] for a in agg.values()) or 1
  print("=====", s["name"][:50], "inst", ti, f"({100*ti/tot_i:.0f}% of all)", "samples", ts)
  for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{a[0]/ti*100:5.1f}%i {a[1]/ts*100:5.1f}%s lsb {a[2]:4d} noi {a[3]:4d} ssb {a[4]:4d} wait {a[5]:4d} mio {a[6]:4d} | {k[0]:4d} {k[1]}")
