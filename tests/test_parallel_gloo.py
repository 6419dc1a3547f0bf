"""N>1 host logic on CPU: world_size-2 gloo processes exercise member sharding,
max-over-ranks timing and the gather of per-member results."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graphcast_b200 import parallel


def test_members_for_rank_partitions():
  for members in (0, 1, 4, 5, 8):
    for world in (1, 2, 3, 8):
      got = [m for r in range(world) for m in parallel.members_for_rank(members, r, world)]
      assert got == list(range(members))
      sizes = [len(parallel.members_for_rank(members, r, world)) for r in range(world)]
      assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    mine = parallel.members_for_rank(5, rank, world)
    t = parallel.max_over_ranks(10.0 + rank)
    local = torch.tensor([[float(m), float(m) ** 2] for m in mine])
    gathered = parallel.gather_member_outputs(local, 5)
    q.put((rank, mine, t, [g.tolist() for g in gathered]))
  finally:
    dist.destroy_process_group()


def test_two_rank_gloo():
  s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  results = sorted(q.get(timeout=120) for _ in procs)
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert results[0][1] == [0, 1, 2] and results[1][1] == [3, 4]
  for _, _, t, gathered in results:
    assert t == 11.0                                        # max over ranks
    assert gathered == [[float(m), float(m) ** 2] for m in range(5)]
