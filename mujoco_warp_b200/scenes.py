"""Benchmark scenes shipped with the package (compiled from the reference's benchmarks/ MJCF by tools/make_fixtures.py)
and the workload table of bench.py / testspeed: sizes follow the reference's benchmark definitions
(/root/reference/benchmarks/humanoid/__init__.py, unitree_g1/__init__.py, aloha/__init__.py:17-26)."""

import os

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "test_data")
HUMANOID = os.path.join(DATA, "humanoid.npz")
THREE_HUMANOIDS = os.path.join(DATA, "three_humanoids.npz")
G1 = os.path.join(DATA, "unitree_g1_flat.npz")
G1_TRAJ = os.path.join(DATA, "unitree_g1_shuffle_dance.npz")
CONVEX_MESH = os.path.join(DATA, "convex_mesh.npz")

# name -> model, worlds per GPU, nconmax, njmax, trajectory to replay (None: keyframe 0 + Ornstein-Uhlenbeck ctrl noise)
WORKLOADS = {
  "humanoid": dict(model=HUMANOID, nworld=8192, nconmax=24, njmax=64, replay=None,
                   label="humanoid.xml nworld=8192/GPU nconmax=24 njmax=64 keyframe=squat Newton/pyramidal/Euler, OU ctrl noise (Halton)"),
  "three_humanoids": dict(model=THREE_HUMANOIDS, nworld=8192, nconmax=100, njmax=192, replay=None,
                          label="three_humanoids (benchmarks/humanoid, nv=81) nworld=8192/GPU nconmax=100 njmax=192, OU ctrl noise"),
  "g1": dict(model=G1, nworld=4096, nconmax=48, njmax=192, replay=G1_TRAJ,
             label="unitree_g1_flat nworld=4096/GPU nconmax=48 njmax=192 implicitfast, replay of shuffle_dance (zero-order hold)"),
  "convex_mesh": dict(model=CONVEX_MESH, nworld=2048, nconmax=64, njmax=256, replay=None,
                      label="convex-mesh stress stand-in for aloha_pot (its STL assets are not in the tree): box / inline-vertex mesh stacks, "
                            "GJK / EPA + multi-contact, nworld=2048/GPU nconmax=64 njmax=256"),
}
