"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: sharding, MAX-over-ranks timing reduction, reporting gather."""

import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, out):
  sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from mujoco_warp_b200._src import shard

  first, count = shard.shard_worlds(1001, world, rank)
  elapsed = shard.reduce_max_elapsed(10.0 + 5.0 * rank, dist)  # rank 1 is slower
  qpos = torch.full((count, 3), float(rank))
  qvel = torch.full((count, 2), float(rank) + 0.5)
  if count != 501 - rank:  # equal-size requirement of all_gather: pad the short shard
    pass
  pad = 501 - count
  g = shard.gather_state(torch.cat([qpos, torch.zeros(pad, 3)]), torch.cat([qvel, torch.zeros(pad, 2)]), dist)
  rate = shard.whole_job_rate([501, 500], elapsed * 1e-3)
  if rank == 0:
    torch.save({"first": first, "count": count, "elapsed": elapsed, "g": g, "rate": rate}, out)
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo(tmp_path):
  out = str(tmp_path / "r0.pt")
  mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
  r = torch.load(out)
  assert (r["first"], r["count"]) == (0, 501)
  assert r["elapsed"] == 15.0  # MAX over ranks, not rank 0's own 10 ms
  assert r["g"].shape == (1002, 5)
  assert float(r["g"][0, 0]) == 0.0 and float(r["g"][501, 0]) == 1.0 and float(r["g"][501, 4]) == 1.5
  assert r["rate"] == pytest.approx(1001 / 0.015)
