"""Mesh partition plan: invariants of the plan, the halo sizes
measured for the level-6 multi-mesh, and - on 2 gloo ranks with the CPU oracle's
MLPs - that a node-partitioned processor with one halo exchange per step equals
the single-rank processor."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graphcast_b200 import icosahedral_mesh as im
from graphcast_b200 import partition
from oracle import gnn as oracle_gnn


def _multimesh(splits):
  meshes = im.get_hierarchy_of_triangular_meshes_for_sphere(splits=splits)
  merged = im.merge_meshes(meshes)
  snd, rcv = im.faces_to_edges(merged.faces)
  return meshes[-1].vertices, np.asarray(snd), np.asarray(rcv)


def test_rcb_balanced_and_deterministic():
  rng = np.random.default_rng(0)
  pts = rng.standard_normal((1001, 3))
  pts /= np.linalg.norm(pts, axis=1, keepdims=True)
  for parts in (1, 2, 4, 8):
    a = partition.recursive_coordinate_bisection(pts, parts)
    b = partition.recursive_coordinate_bisection(pts, parts)
    np.testing.assert_array_equal(a, b)
    sizes = np.bincount(a, minlength=parts)
    assert sizes.sum() == 1001 and sizes.max() - sizes.min() <= parts.bit_length()
  try:
    partition.recursive_coordinate_bisection(pts, 3)
    assert False, "3 parts must be rejected"
  except ValueError:
    pass


def test_plan_invariants_small_mesh():
  verts, snd, rcv = _multimesh(3)
  for parts in (2, 4):
    node_part = partition.recursive_coordinate_bisection(verts, parts)
    plans = partition.build_partition_plan(node_part, snd, rcv, parts)
    seen = np.zeros(len(snd), dtype=np.int32)
    for p in plans:
      seen[p.edge_ids] += 1
      table = np.concatenate([p.owned_nodes, p.halo_nodes])
      # local indices map back to the global edge list; receivers are always owned
      np.testing.assert_array_equal(table[p.local_senders], snd[p.edge_ids])
      np.testing.assert_array_equal(p.owned_nodes[p.local_receivers], rcv[p.edge_ids])
      assert np.all(node_part[p.owned_nodes] == p.rank)
      assert np.all(node_part[p.halo_nodes] == p.halo_owner)
      assert np.all(p.halo_owner != p.rank)
      assert np.all(np.diff(p.edge_ids) > 0)             # original relative order kept
      assert sum(p.recv_counts.values()) == p.halo_nodes.size
    assert np.all(seen == 1)                              # every edge owned exactly once
    for p in plans:                                       # send lists mirror the peers' halos
      for peer, rows in p.send_rows.items():
        want = plans[peer].halo_nodes[plans[peer].halo_owner == p.rank]
        np.testing.assert_array_equal(p.owned_nodes[rows], want)


def test_level6_halo_sizes_match_survey():
  """SURVEY.md section 8(e), P = 8: 5120 nodes and ~41 k edges per part, 2.7 % cross edges,
  about 410-450 halo nodes per part."""
  verts, snd, rcv = _multimesh(6)
  assert verts.shape[0] == 40962 and len(snd) == 327660
  node_part = partition.recursive_coordinate_bisection(verts, 8)
  plans = partition.build_partition_plan(node_part, snd, rcv, 8)
  st = partition.plan_statistics(plans, snd, rcv, node_part)
  assert set(st["nodes_per_part"]) <= {5120, 5121}
  assert 39000 < min(st["edges_per_part"]) and max(st["edges_per_part"]) < 43000
  assert 0.02 < st["cross_edge_fraction"] < 0.035
  assert 350 <= min(st["halo_per_part"]) and max(st["halo_per_part"]) <= 520
  assert max(st["peers_per_part"]) <= 7


def _processor_single(orc, v, e, snd, rcv, steps):
  s, r = torch.as_tensor(snd), torch.as_tensor(rcv)
  for k in range(steps):
    m = orc.mlp(oracle_gnn.mlp_name("mesh_gnn", f"processor_edges_{k}_", "mesh"), [e, v[s], v[r]])
    agg = orc.segment_sum(m, r, v.shape[0])
    v = v + orc.mlp(oracle_gnn.mlp_name("mesh_gnn", f"processor_nodes_{k}_", "mesh_nodes"), [v, agg])
    e = e + m
  return v


def _case(steps):
  verts, snd, rcv = _multimesh(2)                      # 162 nodes, 1260 edges
  params = oracle_gnn.init_params(c_in=5, n_out=3, msg_steps=steps, seed=3)
  rng = np.random.default_rng(11)
  v0 = torch.as_tensor(rng.standard_normal((verts.shape[0], 1, 512)).astype(np.float32))
  e0 = torch.as_tensor(rng.standard_normal((len(snd), 1, 512)).astype(np.float32))
  return verts, snd, rcv, params, v0, e0


def _worker(rank, world, port, q):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    torch.set_num_threads(2)
    steps = 2
    verts, snd, rcv, params, v0, e0 = _case(steps)
    node_part = partition.recursive_coordinate_bisection(verts, world)
    plan = partition.build_partition_plan(node_part, snd, rcv, world)[rank]
    orc = oracle_gnn.Oracle(params, torch.float32)
    v = v0[torch.as_tensor(plan.owned_nodes)]
    e = e0[torch.as_tensor(plan.edge_ids)]
    ls, lr = torch.as_tensor(plan.local_senders), torch.as_tensor(plan.local_receivers)
    for k in range(steps):
      table = partition.exchange_halo(plan, v)          # [owned | halo]
      m = orc.mlp(oracle_gnn.mlp_name("mesh_gnn", f"processor_edges_{k}_", "mesh"),
                  [e, table[ls], table[lr]])
      agg = orc.segment_sum(m, lr, v.shape[0])
      v = v + orc.mlp(oracle_gnn.mlp_name("mesh_gnn", f"processor_nodes_{k}_", "mesh_nodes"), [v, agg])
      e = e + m
    q.put((rank, plan.owned_nodes.tolist(), v.numpy()))
  finally:
    dist.destroy_process_group()


def test_partitioned_processor_matches_single_rank_on_gloo():
  steps = 2
  verts, snd, rcv, params, v0, e0 = _case(steps)
  want = _processor_single(oracle_gnn.Oracle(params, torch.float32), v0, e0, snd, rcv, steps).numpy()
  s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  results = [q.get(timeout=240) for _ in procs]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  got = np.zeros_like(want)
  covered = np.zeros(want.shape[0], dtype=bool)
  for _, owned, v in results:
    got[owned] = v
    covered[owned] = True
  assert covered.all()
  scale = np.abs(want).max()
  assert np.abs(got - want).max() / scale < 2e-6
