"""Public pipeline functions: step, forward and the individually callable stages.

Same names and `(m, d)` signatures as /root/reference/mujoco_warp/__init__.py:26-123 (forward.py:1368 step, :1341 forward,
:635 fwd_position, smooth.py kinematics/com_pos/camlight/crb/factor_m/transmission, collision_driver.py:884 collision,
constraint.py:4897 make_constraint, forward.py:732/1152/1290 fwd_velocity/fwd_actuation/fwd_acceleration, solver.py:3671 solve,
forward.py:387 euler).  Each is one call through the C-ABI on the current torch CUDA stream; `d` is mutated in place;
nothing synchronises the device, so sequences are capturable with torch.cuda.graphs.
"""

from __future__ import annotations

import torch

from . import _lib
from .types import Data, Model


def _call(name: str, m: Model, d: Data):
  if d._model is not m and d._model._handle != m._handle:
    raise ValueError("Data was created for a different Model")
  stream = torch.cuda.current_stream().cuda_stream
  _lib.check(getattr(_lib.lib(), name)(m._handle, d._handle, stream))


def step(m: Model, d: Data):
  """Advance simulation (forward dynamics + Euler integration)."""
  _call("mjb_step", m, d)


def forward(m: Model, d: Data):
  """Forward dynamics."""
  _call("mjb_forward", m, d)


def fwd_position(m: Model, d: Data):
  """Position-dependent computations (kinematics .. make_constraint, transmission)."""
  _call("mjb_fwd_position", m, d)


def kinematics(m: Model, d: Data):
  _call("mjb_kinematics", m, d)


def com_pos(m: Model, d: Data):
  _call("mjb_com_pos", m, d)


def camlight(m: Model, d: Data):
  _call("mjb_camlight", m, d)


def crb(m: Model, d: Data):
  _call("mjb_crb", m, d)


def factor_m(m: Model, d: Data):
  _call("mjb_factor_m", m, d)


def transmission(m: Model, d: Data):
  _call("mjb_transmission", m, d)


def collision(m: Model, d: Data):
  _call("mjb_collision", m, d)


def make_constraint(m: Model, d: Data):
  _call("mjb_make_constraint", m, d)


def fwd_velocity(m: Model, d: Data):
  _call("mjb_fwd_velocity", m, d)


def fwd_actuation(m: Model, d: Data):
  _call("mjb_fwd_actuation", m, d)


def fwd_acceleration(m: Model, d: Data):
  _call("mjb_fwd_acceleration", m, d)


def solve(m: Model, d: Data):
  _call("mjb_solve", m, d)


def euler(m: Model, d: Data):
  _call("mjb_euler", m, d)


def sensor_pos(m: Model, d: Data):
  """Position-stage sensors (reference sensor.py:810); forward() / step() already evaluate every stage after the solver."""
  _call("mjb_sensor_pos", m, d)


def sensor_vel(m: Model, d: Data):
  """Velocity-stage sensors, with subtree_vel when a sensor needs it (reference sensor.py:1432)."""
  _call("mjb_sensor_vel", m, d)


def sensor_acc(m: Model, d: Data):
  """Acceleration-stage sensors, with the cacc part of rne_postconstraint when an accelerometer needs it (reference sensor.py:2512)."""
  _call("mjb_sensor_acc", m, d)


def rungekutta4(m: Model, d: Data):
  """Runge-Kutta 4 integrator, to be called after forward() (reference forward.py:523); the model must use the RK4 integrator."""
  from . import constants as C

  if m.opt.integrator != C.INT_RK4:
    raise NotImplementedError("rungekutta4(): the model was put with another integrator (the RK scratch is allocated per model)")
  _call("mjb_rungekutta4", m, d)


def implicit(m: Model, d: Data):
  """Integrates implicitly in velocity (reference forward.py:578): the full velocity derivative with an LU solve when the model's
  integrator is IMPLICIT, the symmetric implicitfast variant when it is IMPLICITFAST.  (The reference also runs the implicitfast branch
  for Euler / RK4 models; here the factor-and-solve scratch is sized by the model's integrator at put_model, so those raise.)"""
  from . import constants as C

  if m.opt.integrator not in (C.INT_IMPLICITFAST, C.INT_IMPLICIT):
    raise NotImplementedError("implicit(): the model was put with the Euler / RK4 integrator (scratch is sized per integrator)")
  _call("mjb_implicit", m, d)


def fwd_kinematics(m: Model, d: Data):
  """kinematics, com_pos, camlight (reference forward.py:613-632; no flex / tendon in this version)."""
  kinematics(m, d)
  com_pos(m, d)
  camlight(m, d)


def com_vel(m: Model, d: Data):
  """Body velocities cvel and cdof_dot (reference smooth.py:2261)."""
  _call("mjb_com_vel", m, d)


def passive(m: Model, d: Data):
  """Passive joint forces: springs and dampers (reference passive.py:1257)."""
  _call("mjb_passive", m, d)


def rne(m: Model, d: Data, flg_acc: bool = False):
  """Bias forces by recursive Newton-Euler (reference smooth.py:1499)."""
  if flg_acc:
    raise NotImplementedError("rne(flg_acc=True) is not implemented")
  _call("mjb_rne", m, d)


def _vec_call(name: str, m: Model, d: Data, out: torch.Tensor, inp: torch.Tensor):
  for t in (out, inp):
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous() or tuple(t.shape) != (d.nworld, m.nv):
      raise ValueError(f"expected a contiguous CUDA float32 tensor of shape ({d.nworld}, {m.nv})")
  stream = torch.cuda.current_stream().cuda_stream
  _lib.check(getattr(_lib.lib(), name)(m._handle, d._handle, out.data_ptr(), inp.data_ptr(), stream))


def solve_m(m: Model, d: Data, x: torch.Tensor, y: torch.Tensor):
  """x = M^-1 y using the factor in d.qLD (reference smooth.py:3214)."""
  _vec_call("mjb_solve_m", m, d, x, y)


def mul_m(m: Model, d: Data, res: torch.Tensor, vec: torch.Tensor):
  """res = M vec (reference support.py:153)."""
  _vec_call("mjb_mul_m", m, d, res, vec)


def contact_force(m: Model, d: Data, contact_ids: torch.Tensor, to_world_frame: bool, force: torch.Tensor):
  """6D force / torque of the listed contacts into force (n, 6), in the contact frame unless to_world_frame
  (reference support.py:445; contact ids index the global contact pool)."""
  n = int(contact_ids.numel())
  if contact_ids.dtype != torch.int32 or not contact_ids.is_cuda or not contact_ids.is_contiguous():
    raise ValueError("contact_ids: expected a contiguous CUDA int32 tensor")
  if force.dtype != torch.float32 or not force.is_cuda or not force.is_contiguous() or tuple(force.shape) != (n, 6):
    raise ValueError(f"force: expected a contiguous CUDA float32 tensor of shape ({n}, 6)")
  stream = torch.cuda.current_stream().cuda_stream
  _lib.check(_lib.lib().mjb_contact_force(m._handle, d._handle, contact_ids.data_ptr(), n, int(bool(to_world_frame)), force.data_ptr(), stream))


_STATE_ORDER = ("TIME", "QPOS", "QVEL", "ACT", "HISTORY", "WARMSTART", "CTRL", "QFRC_APPLIED", "XFRC_APPLIED", "EQ_ACTIVE", "MOCAP_POS", "MOCAP_QUAT", "USERDATA")


def _state_fields(m: Model, d: Data, sig: int):
  """(tensor, width) of every state component selected by sig, in bit order (reference support.py:722-790)."""
  from .types import State

  if sig >= (1 << int(State.NSTATE)):
    raise ValueError(f"invalid state signature {sig} >= 2^mjNSTATE")
  nw = d.nworld
  out = []
  for name in _STATE_ORDER:
    if not (int(getattr(State, name)) & sig):
      continue
    t = {
      "TIME": lambda: d.time.reshape(nw, 1), "QPOS": lambda: d.qpos, "QVEL": lambda: d.qvel, "WARMSTART": lambda: d.qacc_warmstart,
      "CTRL": lambda: d.ctrl, "QFRC_APPLIED": lambda: d.qfrc_applied, "XFRC_APPLIED": lambda: d.xfrc_applied.reshape(nw, -1),
      "ACT": lambda: d.act, "EQ_ACTIVE": lambda: d.eq_active.reshape(nw, -1), "MOCAP_POS": lambda: d.mocap_pos.reshape(nw, -1), "MOCAP_QUAT": lambda: d.mocap_quat.reshape(nw, -1),
    }.get(name)
    if t is None:  # HISTORY / USERDATA: no history buffers, no user data in this build (width 0)
      continue
    out.append(t())
  return out


def get_state(m: Model, d: Data, state: torch.Tensor, sig: int, active: torch.Tensor = None):
  """Copy the state components selected by the State bit flags in sig from Data into state (nworld, size), concatenated in
  bit order; worlds whose `active` entry is False are left untouched (reference support.py:674)."""
  fields = _state_fields(m, d, int(sig))
  if not fields:
    return
  cat = torch.cat([f.to(torch.float32) for f in fields], dim=1)
  dst = state[:, : cat.shape[1]]
  dst.copy_(cat if active is None else torch.where(active.reshape(-1, 1).bool(), cat, dst))


def set_state(m: Model, d: Data, state: torch.Tensor, sig: int, active: torch.Tensor = None):
  """Inverse of get_state (reference support.py:829)."""
  adr = 0
  for f in _state_fields(m, d, int(sig)):
    w = f.shape[1]
    src = state[:, adr : adr + w].to(f.dtype)
    f.copy_(src if active is None else torch.where(active.reshape(-1, 1).bool(), src, f))
    adr += w


def step1(m: Model, d: Data):
  """First half of a split step, before the user sets controls (reference forward.py:1384 step1: position and velocity stages
  with their sensors; energy is not computed in this build)."""
  fwd_position(m, d)
  if getattr(m, "nsensor", 0):
    d.sensordata.zero_()
    sensor_pos(m, d)
  fwd_velocity(m, d)
  if getattr(m, "nsensor", 0):
    sensor_vel(m, d)


def step2(m: Model, d: Data):
  """Second half of a split step (reference forward.py step2)."""
  from . import constants as C

  fwd_actuation(m, d)
  fwd_acceleration(m, d)
  solve(m, d)
  if getattr(m, "nsensor", 0):
    sensor_acc(m, d)
  if m.opt.integrator in (C.INT_IMPLICITFAST, C.INT_IMPLICIT):
    implicit(m, d)
  else:
    euler(m, d)


def ctrl_noise(m: Model, d: Data, step_index: int, ctrl_center: torch.Tensor | None = None, noise_std: float = 0.01, noise_rate: float = 0.1):
  """Harness OU control noise (reference cli.py:103-145), deterministic Halton sequence per (step, world, actuator)."""
  stream = torch.cuda.current_stream().cuda_stream
  ptr = ctrl_center.data_ptr() if ctrl_center is not None else None
  _lib.check(_lib.lib().mjb_ctrl_noise(m._handle, d._handle, ptr, int(step_index), float(noise_std), float(noise_rate), stream))


def last_launch_count() -> int:
  return int(_lib.lib().mjb_last_launch_count())


KERNEL_NAMES = ("position", "collision", "constraint", "velocity", "solver", "integrate")


def step_profile(m: Model, d: Data):
  """One step with per-kernel CUDA-event timing; returns {kernel: ms}. Synchronises (profiling aid)."""
  import ctypes

  out = (ctypes.c_float * 6)()
  stream = torch.cuda.current_stream().cuda_stream
  _lib.check(_lib.lib().mjb_step_profile(m._handle, d._handle, stream, out))
  return dict(zip(KERNEL_NAMES, [float(x) for x in out]))
