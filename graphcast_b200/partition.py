"""Mesh partition plan for a node-partitioned processor (host side, once per model).

BASELINE config 4 / SURVEY.md section 8(e): the 16 multi-mesh message-passing
steps of one forecast can be split over P GPUs by giving every rank a block of
mesh nodes and ALL edges whose receiver it owns (so the aggregation of
`_node_update`, utils/typed_graph_net.py:532-538, stays rank-local and needs no
reduce-scatter).  The only exchange is, once per step, the latents of the
"halo" nodes: remote *senders* of owned edges.  The reference's analogue is the
all_gather-before-gather / psum_scatter-after-segment_sum pair of
`sparse_transformer`-style sharding; here the receiver-owned edge rule removes
the second collective.

This module builds that plan with numpy only:

* `recursive_coordinate_bisection` - partition of the unit-sphere vertices into
  P = 2^k parts of (almost) equal size by splitting the longest coordinate axis
  at the median; for the level-6 multi-mesh and P = 8 it gives 5120/5121 nodes,
  ~41 k edges, 2.7 % cross edges and ~410-450 halo nodes per part (SURVEY.md
  section 8e measured the same with the reference's icosahedral_mesh).
* `build_partition_plan` - per rank: owned nodes, halo nodes grouped by owner,
  the owned edges in their original relative order with sender / receiver
  indices rewritten into the local table `[owned | halo]`, and the send lists
  (which owned rows every peer needs).
* `exchange_halo` - the per-step exchange over `torch.distributed` (any backend:
  NCCL on GPUs, gloo in the CPU tests): one `all_to_all_single`-equivalent built
  from batched isend / irecv so that it also runs on gloo.

`graphcast_b200/partitioned.py` (round 2) builds on the same bisection and ownership
rule for the WHOLE step on GPUs (all three graphs, grid nodes, NCCL exchange);
`tests/test_partition.py` checks this module's plan invariants, the halo sizes quoted
above, and - with the CPU oracle on 2 gloo ranks - that the partitioned processor
reproduces the single-rank processor.
"""

from __future__ import annotations

import dataclasses
from typing import Dict, List, Sequence

import numpy as np


def recursive_coordinate_bisection(points: np.ndarray, num_parts: int) -> np.ndarray:
  """Part id in [0, num_parts) for every point ([N, 3] coordinates).

  num_parts must be a power of two.  Every split is at the median of the axis
  with the largest extent, ties broken by index (deterministic), so sibling
  parts differ by at most one point."""
  if num_parts < 1 or num_parts & (num_parts - 1):
    raise ValueError("num_parts must be a power of two")
  points = np.asarray(points, dtype=np.float64)
  part = np.zeros(points.shape[0], dtype=np.int32)

  def split(ids: np.ndarray, first: int, count: int) -> None:
    if count == 1:
      part[ids] = first
      return
    p = points[ids]
    axis = int(np.argmax(p.max(axis=0) - p.min(axis=0)))
    order = np.lexsort((ids, p[:, axis]))          # by coordinate, then by index
    half = (len(ids) + 1) // 2
    split(ids[order[:half]], first, count // 2)
    split(ids[order[half:]], first + count // 2, count // 2)

  split(np.arange(points.shape[0]), 0, num_parts)
  return part


@dataclasses.dataclass
class RankPlan:
  """What one rank needs to run its share of a processor step."""
  rank: int
  owned_nodes: np.ndarray        # [n_own] global node ids, ascending
  halo_nodes: np.ndarray         # [n_halo] global ids of remote senders, grouped by owner rank
  halo_owner: np.ndarray         # [n_halo] owner rank of every halo node (non-decreasing)
  edge_ids: np.ndarray           # [e_own] global edge ids (receiver owned), ascending
  local_senders: np.ndarray      # [e_own] index into the local table [owned | halo]
  local_receivers: np.ndarray    # [e_own] index into owned_nodes
  send_rows: Dict[int, np.ndarray]   # peer -> indices into owned_nodes that the peer needs
  recv_counts: Dict[int, int]        # peer -> number of halo rows received from it

  @property
  def num_local_nodes(self) -> int:
    return int(self.owned_nodes.size + self.halo_nodes.size)


def build_partition_plan(node_part: np.ndarray, senders: np.ndarray,
                         receivers: np.ndarray, num_parts: int) -> List[RankPlan]:
  """Plans for all ranks from a node -> part map and the (global) edge list.

  Edges are owned by the part of their receiver.  Within a rank edges keep their
  global relative order, so a receiver-sorted global list stays receiver-sorted."""
  node_part = np.asarray(node_part)
  senders = np.asarray(senders, dtype=np.int64)
  receivers = np.asarray(receivers, dtype=np.int64)
  n_nodes = node_part.shape[0]
  edge_part = node_part[receivers]
  plans: List[RankPlan] = []
  needs: List[Dict[int, np.ndarray]] = []        # needs[r][owner] = global ids r wants from owner
  for r in range(num_parts):
    owned = np.flatnonzero(node_part == r)
    edge_ids = np.flatnonzero(edge_part == r)
    snd, rcv = senders[edge_ids], receivers[edge_ids]
    remote = np.unique(snd[node_part[snd] != r])
    owner = node_part[remote]
    order = np.lexsort((remote, owner))            # grouped by owner, ascending id inside
    halo, halo_owner = remote[order], owner[order]
    # global id -> local row
    local_of = np.full(n_nodes, -1, dtype=np.int64)
    local_of[owned] = np.arange(owned.size)
    local_of[halo] = owned.size + np.arange(halo.size)
    plans.append(RankPlan(
        rank=r, owned_nodes=owned.astype(np.int64), halo_nodes=halo.astype(np.int64),
        halo_owner=halo_owner.astype(np.int32), edge_ids=edge_ids.astype(np.int64),
        local_senders=local_of[snd], local_receivers=local_of[rcv],
        send_rows={}, recv_counts={}))
    needs.append({int(o): halo[halo_owner == o] for o in np.unique(halo_owner)})
  for r, plan in enumerate(plans):
    plan.recv_counts = {o: int(ids.size) for o, ids in needs[r].items()}
    row_of = np.full(n_nodes, -1, dtype=np.int64)
    row_of[plan.owned_nodes] = np.arange(plan.owned_nodes.size)
    for peer in range(num_parts):
      if peer != r and r in needs[peer]:
        plan.send_rows[peer] = row_of[needs[peer][r]]
  return plans


def plan_statistics(plans: Sequence[RankPlan], senders: np.ndarray, receivers: np.ndarray,
                    node_part: np.ndarray) -> Dict[str, object]:
  """Sizes worth printing: nodes / edges / halo per part and the cross-edge fraction."""
  cross = int(np.count_nonzero(node_part[np.asarray(senders)] != node_part[np.asarray(receivers)]))
  return {
      "nodes_per_part": [int(p.owned_nodes.size) for p in plans],
      "edges_per_part": [int(p.edge_ids.size) for p in plans],
      "halo_per_part": [int(p.halo_nodes.size) for p in plans],
      "peers_per_part": [len(p.recv_counts) for p in plans],
      "cross_edges": cross,
      "cross_edge_fraction": cross / max(1, len(senders)),
  }


def exchange_halo(plan: RankPlan, owned_rows, group=None):
  """One halo exchange: returns the local node table `[owned | halo]`.

  owned_rows: torch tensor [n_own, ...] of this rank's node latents.  Uses
  batched isend / irecv (works on NCCL and gloo); message sizes are fixed by the
  plan, so no size negotiation is needed."""
  import torch
  import torch.distributed as dist
  tail = tuple(owned_rows.shape[1:])
  halo = torch.empty((plan.halo_nodes.size,) + tail, dtype=owned_rows.dtype,
                     device=owned_rows.device)
  ops, keep = [], []
  offset = 0
  for peer in sorted(plan.recv_counts):
    n = plan.recv_counts[peer]
    ops.append(dist.P2POp(dist.irecv, halo[offset:offset + n], peer, group))
    offset += n
  for peer in sorted(plan.send_rows):
    rows = torch.as_tensor(plan.send_rows[peer], device=owned_rows.device)
    buf = owned_rows.index_select(0, rows).contiguous()
    keep.append(buf)
    ops.append(dist.P2POp(dist.isend, buf, peer, group))
  if ops:
    for req in dist.batch_isend_irecv(ops):
      req.wait()
  return torch.cat([owned_rows, halo], dim=0)
