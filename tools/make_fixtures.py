"""Generates the committed model fixtures from the reference's MJCF scenes (run in the build container only).

  python tools/make_fixtures.py

/root/reference does not exist on the GPU box, so tests / bench / smoke load these .npz files (MjModel-named arrays
produced by mujoco_warp_b200.mjcf from the reference's benchmark scene) instead of the XML.
"""

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mujoco_warp_b200._src import mjcf  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "mujoco_warp_b200", "test_data")

SCENES = {
  "humanoid": "benchmarks/humanoid/humanoid.xml",  # BASELINE configs[1]: iterations=100, ls_iterations=50
  "three_humanoids": "benchmarks/humanoid/three_humanoids.xml",  # benchmarks/humanoid/__init__.py second entry: nv 81, <replicate> + <attach>
  "unitree_g1_flat": "benchmarks/unitree_g1/scene_flat.xml",  # BASELINE configs[2] (visual mesh geoms skipped: STL not in tree)
}


def convex_mesh_xml():
  """Stand-in for BASELINE configs[3] (aloha_pot: a convex-mesh GJK / EPA narrowphase stress whose STL assets are not in the tree):
  an octagonal "pot" and its lid (16-vertex hulls -> hill climbing on the hull graph), mesh wedges and a tetrahedron (exhaustive
  support search), a box tower (discrete box-box CCD + multi-contact), box-mesh and mesh-mesh stacks on a plane, and two actuated
  pushers that shove the pot around.  Inline `vertex=` meshes in the style of collision_gjk_test.py:341-465."""
  import numpy as np

  def prism(r, h, n=8):
    a = np.arange(n) * 2 * np.pi / n
    ring = np.stack([r * np.cos(a), r * np.sin(a)], axis=1)
    v = np.concatenate([np.c_[ring, np.full(n, -h)], np.c_[ring, np.full(n, h)]])
    return " ".join(f"{x:.6f}" for x in v.reshape(-1))

  return f"""
<mujoco model="convex_mesh_stress">
  <option timestep="0.002" iterations="50" ls_iterations="50" ccd_iterations="35" integrator="implicitfast"/>
  <asset>
    <mesh name="pot" vertex="{prism(0.08, 0.03)}"/>
    <mesh name="lid" vertex="{prism(0.085, 0.006)}"/>
    <mesh name="wedge" vertex="0 0 0  1 0 0  0 1 0  0 0 1  1 1 0  0.3 0.3 0.1  1 1 1" scale="0.12 0.12 0.12"/>
    <mesh name="tetra" vertex="0 0 0  1 0 0  0 1 0  0 0 1" scale="0.1 0.1 0.1"/>
    <mesh name="brick" vertex="-1 -1 -1  1 -1 -1  -1 1 -1  1 1 -1  -1 -1 1  1 -1 1  -1 1 1  1 1 1" scale="0.06 0.04 0.025"/>
  </asset>
  <default><geom friction="0.8 0.01 0.002" density="700" condim="3"/></default>
  <worldbody>
    <geom name="floor" type="plane" size="0 0 .05"/>
    <body name="pot" pos="0 0 0.0305"><freejoint/><geom type="mesh" mesh="pot"/></body>
    <body name="lid" pos="0.002 0.001 0.0675"><freejoint/><geom type="mesh" mesh="lid"/></body>
    <body name="box_a" pos="0.3 0 0.0305"><freejoint/><geom type="box" size="0.05 0.05 0.03"/></body>
    <body name="box_b" pos="0.305 0.004 0.0915" euler="0 0 20"><freejoint/><geom type="box" size="0.045 0.045 0.03"/></body>
    <body name="brick_on_box" pos="0.3 0 0.1475" euler="0 0 35"><freejoint/><geom type="mesh" mesh="brick"/></body>
    <body name="brick_a" pos="-0.3 0.1 0.0255"><freejoint/><geom type="mesh" mesh="brick"/></body>
    <body name="brick_b" pos="-0.295 0.105 0.0765" euler="0 0 50"><freejoint/><geom type="mesh" mesh="brick"/></body>
    <body name="wedge" pos="-0.1 -0.3 0.125" euler="180 0 0"><freejoint/><geom type="mesh" mesh="wedge"/></body>
    <body name="tetra" pos="0.15 -0.3 0.002"><freejoint/><geom type="mesh" mesh="tetra"/></body>
    <body name="pusher_l" pos="-0.2 0 0.04"><joint name="push_l" type="slide" axis="1 0 0" range="-0.05 0.12" limited="true" damping="5"/><geom type="box" size="0.02 0.06 0.03" density="2000"/></body>
    <body name="pusher_r" pos="0.2 0.02 0.04"><joint name="push_r" type="slide" axis="-1 0 0" range="-0.05 0.12" limited="true" damping="5"/><geom type="capsule" size="0.02 0.05" euler="90 0 0" density="2000"/></body>
  </worldbody>
  <actuator>
    <position name="act_l" joint="push_l" kp="120" ctrlrange="-0.05 0.12" ctrllimited="true"/>
    <position name="act_r" joint="push_r" kp="120" ctrlrange="-0.05 0.12" ctrllimited="true"/>
  </actuator>
  <keyframe>
    <key name="start" ctrl="0.06 0.06"/>
  </keyframe>
</mujoco>
"""


# replay trajectories are benchmark INPUT data (ctrl sequences), copied verbatim
TRAJECTORIES = {"unitree_g1_shuffle_dance.npz": "benchmarks/unitree_g1/shuffle_dance.npz"}

if __name__ == "__main__":
  os.makedirs(OUT, exist_ok=True)
  for name, rel in SCENES.items():
    m = mjcf.load(os.path.join(REF, rel))
    mjcf.save_npz(m, os.path.join(OUT, name + ".npz"))
    print(name, "nq", m.nq, "nv", m.nv, "nbody", m.nbody, "ngeom", m.ngeom, "->", os.path.join(OUT, name + ".npz"))
  m = mjcf.load_string(convex_mesh_xml())
  mjcf.save_npz(m, os.path.join(OUT, "convex_mesh.npz"))
  print("convex_mesh", "nq", m.nq, "nv", m.nv, "nbody", m.nbody, "ngeom", m.ngeom, "nmesh", m.nmesh)
  import shutil

  for dst, rel in TRAJECTORIES.items():
    shutil.copyfile(os.path.join(REF, rel), os.path.join(OUT, dst))
    print("copied", rel, "->", dst)
