"""Per-stage event trace with the reference's nested key structure (reference _src/warp_util.py:51-145 EventTracer / event_scope,
flattened by testspeed.py:75-89 into "step.forward.fwd_position.kinematics"-style keys that benchmarks/run.py compares per stage).

The fused step runs a handful of kernels, each covering several of the reference's stage functions, so a trace with the reference's
keys is taken from ONE stage-wise step: every public stage function is launched on its own (the same kernels with a stage mask; the
stage-wise path is bit-identical to the fused one, tests/test_gpu_parity.py) between CUDA event pairs, nested exactly like the
reference's call tree (forward.py:1369 step -> :1342 forward -> :636 fwd_position -> :616 fwd_kinematics -> smooth.kinematics ...).
A parent's time is measured around its children, like the reference's decorator."""

from __future__ import annotations

import torch

from . import constants as C
from . import forward as F


def _tree(m):
  """(name, fn | None, children) call tree of step(m, d) in the reference's stage order."""
  sens = bool(getattr(m, "nsensor", 0))
  kin = ("fwd_kinematics", None, [("kinematics", F.kinematics, []), ("com_pos", F.com_pos, []), ("camlight", F.camlight, [])])
  pos = ("fwd_position", None, [kin, ("crb", F.crb, []), ("collision", F.collision, []), ("make_constraint", F.make_constraint, []), ("transmission", F.transmission, [])])
  # one launch: actuator velocities + com_vel + passive + rne (running the three children alone would skip the actuator velocities)
  vel = ("fwd_velocity", F.fwd_velocity, [])
  fwd = [pos] + ([("sensor_pos", F.sensor_pos, [])] if sens else []) + [vel] + ([("sensor_vel", F.sensor_vel, [])] if sens else [])
  fwd += [("fwd_actuation", F.fwd_actuation, []), ("fwd_acceleration", F.fwd_acceleration, []), ("solve", F.solve, [])]
  fwd += [("sensor_acc", F.sensor_acc, [])] if sens else []
  if m.opt.integrator == C.INT_RK4:
    integ = ("rungekutta4", F.rungekutta4, [])
  elif m.opt.integrator in (C.INT_IMPLICITFAST, C.INT_IMPLICIT):
    integ = ("implicit", F.implicit, [])
  else:
    integ = ("euler", F.euler, [])
  return ("step", None, [("forward", None, fwd), integ])


def event_trace_step(m, d) -> dict:
  """Runs one step stage by stage; returns the reference's trace structure {name: ((elapsed_ms,), sub_trace)}."""
  stream = torch.cuda.current_stream()
  recs = []

  def run(node):
    name, fn, children = node
    beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    beg.record(stream)
    if fn is not None:
      fn(m, d)
    sub = [run(c) for c in children]
    end.record(stream)
    rec = (name, beg, end, sub)
    recs.append(rec)
    return rec

  root = run(_tree(m))
  stream.synchronize()

  def build(rec):
    name, beg, end, sub = rec
    return name, ((beg.elapsed_time(end),), dict(build(s) for s in sub))

  k, v = build(root)
  return {k: v}


def flatten_trace(trace: dict, scale: float = 1.0) -> dict:
  """testspeed.py:75-89 _flatten_trace: {"step": t, "step.forward": t, ...} (sum over recorded events times `scale`)."""
  out = {}

  def rec(prefix, tr):
    for k, (times, sub) in tr.items():
      out[prefix + k] = scale * sum(times)
      rec(prefix + k + ".", sub)

  rec("", trace)
  return out
