"""Shared helpers for parity tests: scene setup, seeded states, oracle/CUDA field comparison."""

import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES = os.path.join(ROOT, "mujoco_warp_b200", "test_data")
HUMANOID = os.path.join(SCENES, "humanoid.npz")
G1 = os.path.join(SCENES, "unitree_g1_flat.npz")
G1_TRAJ = os.path.join(SCENES, "unitree_g1_shuffle_dance.npz")


def seeded_state(mjm, nworld, key=0, seed=42, qpos_noise=0.05, qvel_noise=0.5, ctrl_noise=0.5):
  """Per-world states around a keyframe, like the reference fixture's seeded uniform noise (test_data/__init__.py:82-98)."""
  from mujoco_warp_b200._src import constants as C

  rng = np.random.default_rng(seed)
  qpos = np.tile(mjm.key_qpos[key] if key is not None and mjm.nkey > key else mjm.qpos0, (nworld, 1)).astype(np.float64)
  qpos += qpos_noise * rng.uniform(-1, 1, qpos.shape)
  # world 0 stays exactly at the keyframe
  qpos[0] = mjm.key_qpos[key] if key is not None and mjm.nkey > key else mjm.qpos0
  for j in range(mjm.njnt):
    qa = mjm.jnt_qposadr[j]
    if mjm.jnt_type[j] == C.JNT_FREE:
      qpos[:, qa + 3 : qa + 7] /= np.linalg.norm(qpos[:, qa + 3 : qa + 7], axis=1, keepdims=True)
    elif mjm.jnt_type[j] == C.JNT_BALL:
      qpos[:, qa : qa + 4] /= np.linalg.norm(qpos[:, qa : qa + 4], axis=1, keepdims=True)
  qvel = qvel_noise * rng.uniform(-1, 1, (nworld, mjm.nv))
  qvel[0] = 0
  ctrl = ctrl_noise * rng.uniform(-1, 1, (nworld, mjm.nu))
  warm = rng.uniform(-1, 1, (nworld, mjm.nv))
  return qpos, qvel, ctrl, warm


def make_oracle(mjm, nworld, nconmax, njmax, dtype=np.float64):
  from mujoco_warp_b200._src import mjcf
  from oracle import orc

  kin = mjcf.kinematics_np(mjm, mjm.qpos0)
  return orc.Oracle(mjm, nworld=nworld, nconmax=nconmax, njmax=njmax, dtype=dtype, static_kin=kin)


def world_contacts(d, w):
  """Indices of world w's contacts in the global pool, in pool order."""
  nacon = int(d.nacon.cpu()[0])
  wid = d.contact.worldid[:nacon].cpu().numpy()
  return np.nonzero(wid == w)[0]


SMOOTH_FIELDS = [
  "xpos", "xquat", "xmat", "xipos", "ximat", "xanchor", "xaxis", "geom_xpos", "geom_xmat", "cam_xpos", "cam_xmat", "light_xpos", "light_xdir",
  "subtree_com", "cdof", "cinert", "crb", "M", "actuator_length", "actuator_moment", "actuator_velocity", "cvel", "cdof_dot", "qfrc_bias",
  "qfrc_spring", "qfrc_damper", "qfrc_passive", "actuator_force", "qfrc_actuator", "qfrc_smooth", "qacc_smooth", "cacc", "cfrc_int", "qLD",
]


def assert_close(name, a, b, atol, rtol):
  a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
  assert a.shape == b.shape, f"{name}: shape {a.shape} vs {b.shape}"
  err = np.abs(a - b)
  tol = atol + rtol * np.abs(b)
  if not (err <= tol).all():
    i = np.unravel_index(np.argmax(err - tol), err.shape)
    raise AssertionError(f"{name}: max violation at {i}: got {a[i]:.8g}, want {b[i]:.8g} (|err|={err[i]:.3g}, tol={tol[i]:.3g})")
