"""World sharding across the GPUs of one box: worlds never interact, so rank r of R owns a contiguous
block of worlds with its own Data and a replicated Model; there is no collective inside `step`.  The only collective on
the measurement path is the MAX-reduction of the per-rank elapsed time (and an optional post-step gather for reporting)."""

from __future__ import annotations


def shard_worlds(nworld_total: int, world_size: int, rank: int) -> tuple[int, int]:
  """(first_world, count) of rank's contiguous block; blocks differ by at most one world."""
  if not (0 <= rank < world_size):
    raise ValueError("rank out of range")
  base, rem = divmod(nworld_total, world_size)
  count = base + (1 if rank < rem else 0)
  first = rank * base + min(rank, rem)
  return first, count


def whole_job_rate(units_per_rank: list[int] | int, elapsed_s_max: float, world_size: int | None = None) -> float:
  """Whole-job throughput = units all ranks processed / max-over-ranks elapsed time.  An int means the same count on every
  rank and needs world_size."""
  if isinstance(units_per_rank, int):
    if world_size is None:
      raise ValueError("whole_job_rate: world_size is required when units_per_rank is a single count")
    total = units_per_rank * world_size
  else:
    total = sum(units_per_rank)
  return total / elapsed_s_max


def reduce_max_elapsed(elapsed_ms: float, dist=None, device=None) -> float:
  """MAX over ranks of a device-timed duration (torch.distributed all_reduce); identity for a single process."""
  if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
    return float(elapsed_ms)
  import torch

  if device is None:  # NCCL reduces device tensors only; gloo takes CPU tensors
    device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
  t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
  return float(t[0])


def gather_state(qpos, qvel, dist=None):
  """Optional post-step reporting gather of (nworld_local, nq+nv) state over NCCL/NVLink (all_gather); off the timed path."""
  import torch

  x = torch.cat([qpos, qvel], dim=1).contiguous()
  if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
    return x
  out = [torch.empty_like(x) for _ in range(dist.get_world_size())]
  dist.all_gather(out, x)
  return torch.cat(out, dim=0)
