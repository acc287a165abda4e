"""The reference's GJK / EPA known-answer tests (collision_gjk_test.py:368-1002, transcribed in test_oracle_gjk_vectors.CASES) through
the public API on the GPU: each geom sits in a free body posed through qpos, `forward()` runs the collision pipeline and the
contact pool is checked against the reference's expected depth / contact count / witness points.  Only penetrating convex-path
cases apply (the pipeline reports contacts, not distances; sphere-sphere is a primitive pair)."""
import numpy as np
import pytest
import torch

from tests.test_oracle_gjk_vectors import CASES, posed_geoms

pytestmark = pytest.mark.gpu

GPU_CASES = [n for n in sorted(CASES) if not n.startswith("sphere")]
TYPE_NAME = {2: "sphere", 3: "capsule", 4: "ellipsoid", 5: "cylinder", 6: "box"}


def mat2quat(m):
  m = np.asarray(m, dtype=np.float64).reshape(3, 3)
  K = np.array([
    [m[0, 0] - m[1, 1] - m[2, 2], 0, 0, 0], [m[0, 1] + m[1, 0], m[1, 1] - m[0, 0] - m[2, 2], 0, 0],
    [m[0, 2] + m[2, 0], m[1, 2] + m[2, 1], m[2, 2] - m[0, 0] - m[1, 1], 0], [m[2, 1] - m[1, 2], m[0, 2] - m[2, 0], m[1, 0] - m[0, 1], m[0, 0] + m[1, 1] + m[2, 2]]]) / 3.0
  w, v = np.linalg.eigh(K)
  q = v[[3, 0, 1, 2], np.argmax(w)]
  return q * (1.0 if q[0] >= 0 else -1.0)


@pytest.mark.parametrize("name", GPU_CASES)
def test_gpu_reproduces_reference_gjk_vectors(built, name):
  import mujoco_warp_b200 as mjw

  case = CASES[name]
  geoms = posed_geoms(case)
  bodies = "".join(
    f'<body><freejoint/><geom type="{TYPE_NAME[t]}" size="{" ".join(repr(float(x)) for x in s)}"/></body>' for t, s, _, _ in geoms)
  xml = f'<mujoco><option ccd_iterations="{case.get("iterations", 35)}" gravity="0 0 0"/><worldbody>{bodies}</worldbody></mujoco>'
  mjm = mjw.mjcf.load_string(xml)
  m = mjw.put_model(mjm)
  d = mjw.make_data(mjm, nworld=1, nconmax=8, njmax=32, m=m)
  qpos = np.concatenate([np.concatenate([p, mat2quat(R)]) for _, _, p, R in geoms]).astype(np.float32)
  d.qpos.copy_(torch.from_numpy(qpos[None]))
  mjw.forward(m, d)
  torch.cuda.synchronize()
  assert int(d.overflow.cpu()[0]) == 0
  n = int(d.nacon.cpu()[0])
  if n == 0:
    # box_box_float: the reference only asserts dist < 1e-4 (true depth ~2e-6); GJK / EPA report 0.0 there (the oracle does too,
    # in both widths), and with margin 0 a distance of 0 is not a contact (collision_convex.py:860)
    assert "dist" not in case and not case.get("multiccd", False)
    return
  dist = d.contact.dist[:n].cpu().numpy().astype(np.float64)
  assert np.ptp(dist) == 0.0  # every contact of a pair carries the EPA depth
  if "dist" in case:
    # the expectations come from fp32 runs; allow a few fp32 roundings of the pose on top of the reference's own places
    tol = 0.5 * 10.0 ** -case.get("places", 7) + 4.0 * np.finfo(np.float32).eps * max(1.0, float(np.abs(qpos).max()))
    assert abs(dist[0] - case["dist"]) < tol, (dist[0], case["dist"])
  if "dist_less" in case:
    assert dist[0] < case["dist_less"]
  if "ncon" in case and case.get("multiccd", False):
    assert n == case["ncon"], (n, case["ncon"])
  frame = d.contact.frame[0].cpu().numpy().reshape(3, 3)
  pos = d.contact.pos[:n].cpu().numpy().astype(np.float64)
  if "normal" in case:  # frame[0] points from geom1 to geom2 = witness1 - witness2 while penetrating
    np.testing.assert_allclose(frame[0], case["normal"], atol=1e-5)
  if "x1" in case:  # contact position is the midpoint of the witness points
    np.testing.assert_allclose(pos[0], 0.5 * (np.asarray(case["x1"]) + np.asarray(case["x2"])), atol=case["xtol"])
