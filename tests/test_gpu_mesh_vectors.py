"""The reference's mesh known-answer tests (collision_gjk_test.py:341 box_mesh_distance is a separation, the other four penetrate:
:405 mesh_mesh_contact, :441 mesh_mesh_contact2, :528 mesh_mesh_ccd, :648 sphere_mesh_margin) through the public API on the GPU:
the geoms are world-fixed, `forward()` runs the collision pipeline of the mesh build of the kernel, and the contact pool is checked
against the reference's expected depth / contact count."""
import numpy as np
import pytest
import torch

from tests.test_mesh_gjk_vectors import CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_reproduces_reference_mesh_vectors(built, name):
  import mujoco_warp_b200 as mjw

  case = CASES[name]
  # the two geoms go into separate free bodies at the origin (geom pos / quat carry the pose) so that they form a dynamic pair
  geoms = case["body"].replace("/><geom", "/>|<geom").split("|")
  body = "".join(f"<body><freejoint/>{g}</body>" for g in geoms)
  margin = f' margin="{case["margin"]}"' if "margin" in case else ""
  body = body.replace("<geom ", f"<geom{margin} ") if margin else body
  xml = f'<mujoco><option gravity="0 0 0"/><asset>{case["asset"]}</asset><worldbody>{body}</worldbody></mujoco>'
  mjm = mjw.mjcf.load_string(xml)
  m = mjw.put_model(mjm)
  d = mjw.make_data(mjm, nworld=2, nconmax=8, njmax=32, m=m)
  mjw.forward(m, d)
  torch.cuda.synchronize()
  assert (d.overflow.cpu().numpy() == 0).all()
  n = int(d.nacon.cpu()[0])
  if case.get("dist", -1.0) > 0 and "margin" not in case:
    assert n == 0  # separated pair: no contact
    return
  assert n >= 2 and n % 2 == 0  # two identical worlds
  wid = d.contact.worldid[:n].cpu().numpy()
  dist = d.contact.dist[:n].cpu().numpy().astype(np.float64)[wid == 0]
  assert np.ptp(dist) == 0.0
  if "dist" in case:
    # sphere_mesh_margin: the pipeline reports the distance between the un-inflated surfaces (collision_convex.py:862-868)
    want = case["dist"] + case.get("margin", 0.0)
    assert abs(dist[0] - want) < 5e-7, (dist[0], want)
  if "ncon" in case and case.get("multiccd", False):
    assert len(dist) == case["ncon"], (len(dist), case["ncon"])
