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
# replay trajectories are benchmark INPUT data (ctrl sequences), copied verbatim
TRAJECTORIES = {"unitree_g1_shuffle_dance.npz": "benchmarks/unitree_g1/shuffle_dance.npz"}

if __name__ == "__main__":
  os.makedirs(OUT, exist_ok=True)
  for name, rel in SCENES.items():
    m = mjcf.load(os.path.join(REF, rel))
    mjcf.save_npz(m, os.path.join(OUT, name + ".npz"))
    print(name, "nq", m.nq, "nv", m.nv, "nbody", m.nbody, "ngeom", m.ngeom, "->", os.path.join(OUT, name + ".npz"))
  import shutil

  for dst, rel in TRAJECTORIES.items():
    shutil.copyfile(os.path.join(REF, rel), os.path.join(OUT, dst))
    print("copied", rel, "->", dst)
