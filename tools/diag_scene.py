"""Diagnostic: one golden scene, teacher-forced from the reference's state after step s-1; GPU forward vs the fp32 oracle, field by field.
usage: python tools/diag_scene.py <scene> <s>"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mujoco_warp_b200 as mjw
from tests import util
from tests.test_oracle_golden_pipeline import GOLD_DIR, load_scene

name, s = sys.argv[1], int(sys.argv[2])
g = np.load(os.path.join(GOLD_DIR, f"pipeline_{name}.npz"))
mjm = load_scene(name)
nworld = g["in/qpos"].shape[0]
ncm, njm = int(g["in/nconmax"]), int(g["in/njmax"])
o = util.make_oracle(mjm, nworld, ncm, njm, dtype=np.float32)
st = dict(qpos=g[f"step{s-1}/qpos"], qvel=g[f"step{s-1}/qvel"], ctrl=g["in/ctrl"], qacc_warmstart=g[f"step{s-1}/qacc_warmstart"]) if s > 0 else dict(qpos=g["in/qpos"], qvel=g["in/qvel"], ctrl=g["in/ctrl"], qacc_warmstart=g["in/qacc_warmstart"])
o.set_state(**st); o.forward()
m = mjw.put_model(mjm)
d = mjw.make_data(mjm, nworld=nworld, nconmax=ncm, njmax=njm, m=m)
f32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32)).cuda()
for k, v in st.items():
  getattr(d, k).copy_(f32(v))
mjw.forward(m, d); torch.cuda.synchronize()
print("nefc gpu", d.nefc.cpu().numpy().tolist(), "oracle", o.d["nefc"].tolist(), "niter gpu", d.solver_niter.cpu().numpy().tolist(), "oracle", o.d["solver_niter"].tolist(), "ovf", d.overflow.cpu().numpy().tolist())
for f in ("qacc_smooth", "qfrc_smooth", "qacc", "qfrc_constraint", "M", "qLD"):
  a, b = getattr(d, f).cpu().numpy().reshape(nworld, -1), np.asarray(o.d[f]).reshape(nworld, -1)
  n = min(a.shape[1], b.shape[1]); e = np.abs(a[:, :n] - b[:, :n])
  i = np.unravel_index(e.argmax(), e.shape)
  print(f"{f:16s} max err {e.max():.4g} at {i} gpu {a[i]:.6g} oracle {b[i]:.6g} scale {np.abs(b).max():.4g}")
J = util.dense_J(d)
for w in range(nworld):
  ne = int(o.d["nefc"][w])
  for f in ("D", "aref", "force", "pos", "vel"):
    a, b = getattr(d.efc, f)[w, :ne].cpu().numpy(), o.d["efc_" + f][w][:ne]
    e = np.abs(a - b); i = int(e.argmax()) if ne else 0
    print(f"w{w} efc_{f:6s} max err {e.max() if ne else 0:.4g} at row {i} gpu {a[i] if ne else 0:.6g} oracle {b[i] if ne else 0:.6g}")
  e = np.abs(J[w, :ne, :mjm.nv] - o.d["efc_J"][w][:ne, :mjm.nv]); print(f"w{w} efc_J max err {e.max():.4g}")
  st_g, st_o = d.efc.state[w, :ne].cpu().numpy(), o.d["efc_state"][w][:ne]
  print(f"w{w} efc_state differs in {int((st_g != st_o).sum())} rows")
# ---- contacts per geom pair, GPU vs oracle
import collections
for w in range(nworld):
  ids = util.world_contacts(d, w)
  gg = d.contact.geom[ids].cpu().numpy(); gd = d.contact.dist[ids].cpu().numpy()
  nc = int(o.d["ncon"][w]); og = o.d["con_geom"][w][:nc]; od_ = o.d["con_dist"][w][:nc]
  cg, co = collections.Counter(map(tuple, gg.tolist())), collections.Counter(map(tuple, og.tolist()))
  for pair in sorted(set(cg) | set(co)):
    if cg[pair] != co[pair]:
      a, b = pair
      print(f"w{w} pair {pair}: gpu {cg[pair]} contacts dist {gd[[i for i in range(len(gg)) if tuple(gg[i]) == pair]]}, oracle {co[pair]} dist {od_[[i for i in range(nc) if tuple(og[i]) == pair]]}")
      for gi in pair:
        print(f"   geom {gi} type {int(mjm.geom_type[gi])} size {np.asarray(mjm.geom_size[gi]).tolist()}")
        print("   gpu xpos", repr(d.geom_xpos[w, gi].cpu().numpy().astype(np.float64).tolist()), "xmat", repr(d.geom_xmat[w, gi].cpu().numpy().reshape(-1).astype(np.float64).tolist()))
        print("   orc xpos", repr(np.asarray(o.d["geom_xpos"][w, gi], dtype=np.float64).tolist()), "xmat", repr(np.asarray(o.d["geom_xmat"][w, gi], dtype=np.float64).reshape(-1).tolist()))
