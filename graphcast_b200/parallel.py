"""Multi-GPU plumbing: one process per GPU over torch.distributed.

The GraphCast step shards trivially over independent forecasts (ensemble members
/ batch elements) -- what the reference does with `pmap` in
`chunked_prediction_generator_multiple_runs` (utils/rollout.py:158-306,
`replicate_dataset` :91-155) -- with NO data-path collective.  These helpers
assign members to ranks and reduce timings (max over ranks, the rule for every
multi-GPU number).  They run on any backend (nccl on GPUs, gloo in CPU tests)."""

from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def members_for_rank(num_members: int, rank: int, world_size: int) -> List[int]:
  """Contiguous block partition of ensemble members over ranks (sizes differ by <= 1)."""
  if num_members < 0 or world_size <= 0 or not 0 <= rank < world_size:
    raise ValueError("bad rank / world_size / num_members")
  base, extra = divmod(num_members, world_size)
  lo = rank * base + min(rank, extra)
  hi = lo + base + (1 if rank < extra else 0)
  return list(range(lo, hi))


def max_over_ranks(value: float, device=None) -> float:
  """Max of a host scalar over all ranks (identity when not initialised)."""
  if not (dist.is_available() and dist.is_initialized()):
    return float(value)
  t = torch.tensor([value], dtype=torch.float64, device=device)
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
  return float(t.item())


def gather_member_outputs(local: torch.Tensor, num_members: int) -> List[torch.Tensor]:
  """All-gather per-member result tensors (e.g. a scalar metric per member) so
  rank 0 can assemble the ensemble; ranks may hold different member counts."""
  if not (dist.is_available() and dist.is_initialized()):
    return [local[i] for i in range(local.shape[0])]
  world = dist.get_world_size()
  counts = [len(members_for_rank(num_members, r, world)) for r in range(world)]
  width = max(counts)
  pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
  pad[:local.shape[0]] = local
  bufs = [torch.empty_like(pad) for _ in range(world)]
  dist.all_gather(bufs, pad)
  out = []
  for r, c in enumerate(counts):
    out.extend(bufs[r][i] for i in range(c))
  return out
