"""Host logic of the node-partitioned step (graphcast_b200/partitioned.py), on the CPU: every
rank's local graph is run through the ORACLE's stages, with the halo exchange emulated by copies,
and the owned rows are assembled -- the result must be the full oracle step.  This pins the
ownership rules, the local numbering ([owned | padding | halo]), the receiver-owned edge sets, the
grid halo (recomputed encoder latents) and the send / receive lists without a GPU."""
import numpy as np
import pytest
import torch

import _cases
from graphcast_b200 import partitioned
from oracle import gnn as oracle_gnn


def _run_partitioned(g, params, x, P):
  orc = oracle_gnn.Oracle(params, torch.float64)
  lgs = [partitioned.build_local_graph(g, P, r) for r in range(P)]
  # every node / edge is owned exactly once
  assert sorted(np.concatenate([l.mesh_owned for l in lgs]).tolist()) == list(range(g.num_mesh_nodes))
  assert sorted(np.concatenate([l.grid_owned for l in lgs]).tolist()) == list(range(g.num_grid_nodes))
  assert sum(l.graph.mesh_senders.size for l in lgs) == g.mesh_senders.size
  assert sum(l.graph.g2m_senders.size for l in lgs) == g.g2m_senders.size
  assert sum(l.graph.m2g_senders.size for l in lgs) == g.m2g_senders.size
  for l in lgs:
    assert l.mesh_owned_pad % 128 == 0 and l.mesh_owned_pad >= l.mesh_owned.size
    assert sum(l.recv_counts) == l.mesh_halo.size and l.recv_counts[l.rank] == 0
    assert np.array_equal(l.graph.m2g_receivers, np.repeat(np.arange(l.grid_owned.size), 3))
    assert l.graph.mesh_receivers.max() < l.mesh_owned.size            # receivers are owned rows
    assert l.graph.g2m_receivers.max() < l.mesh_owned.size
  for a in lgs:                       # what a sends to b is what b expects from a, in b's halo order
    off = 0
    for b in range(P):
      n = a.send_counts[b]
      rows = a.send_rows[off:off + n]; off += n
      if n:
        want = lgs[b].mesh_halo[sum(lgs[b].recv_counts[:a.rank]):sum(lgs[b].recv_counts[:a.rank + 1])]
        np.testing.assert_array_equal(a.mesh_owned[rows], want)

  def exchange(tables):
    """tables[r]: [n_mesh_local, B, D] with valid owned rows -> halo rows filled from the owners."""
    glob = torch.zeros((g.num_mesh_nodes,) + tuple(tables[0].shape[1:]), dtype=tables[0].dtype)
    for l, t in zip(lgs, tables):
      glob[torch.as_tensor(l.mesh_owned)] = t[:l.mesh_owned.size]
    out = []
    for l, t in zip(lgs, tables):
      t = t.clone()
      t[l.mesh_owned_pad:l.mesh_owned_pad + l.mesh_halo.size] = glob[torch.as_tensor(l.mesh_halo)]
      out.append(t)
    return out

  enc = [orc.encoder(l.graph.as_dict(), x[l.local_grid_ids]) for l in lgs]
  v = exchange([vm for vm, _ in enc])
  e = [orc.processor_embed(l.graph.as_dict(), x.shape[1]) for l in lgs]
  for k in range(orc.num_message_steps()):
    nxt = [orc.processor_step(l.graph.as_dict(), v[r], e[r], k) for r, l in enumerate(lgs)]
    e = [b for _, b in nxt]
    v = exchange([a for a, _ in nxt])
  y = np.zeros((g.num_grid_nodes, x.shape[1], 23))
  for r, l in enumerate(lgs):
    out = orc.decoder(l.graph.as_dict(), v[r], enc[r][1])
    y[l.grid_owned] = out[:l.grid_owned.size].numpy()
  return y, lgs


@pytest.mark.parametrize("P", [2, 4])
def test_partitioned_step_equals_the_full_step(P):
  g, params, x = _cases.small_case(c_in=31, n_out=23, msg_steps=3, batch=2)
  ref = oracle_gnn.Oracle(params, torch.float64).forward(g.as_dict(), x).numpy()
  y, lgs = _run_partitioned(g, params, x.astype(np.float64), P)
  assert np.abs(y - ref).max() / np.abs(ref).max() < 1e-12


def test_partition_sizes_at_the_benchmark_resolution():
  """0.25 deg / mesh 6, 8 ranks: balanced ownership, halo <= 1.3 MB of fp32 rows per step."""
  from graphcast_b200 import graph as graph_lib, synthetic
  lat, lon = synthetic.grid_coords(0.25)
  g = graph_lib.cached_static_graph(grid_lat=lat, grid_lon=lon, mesh_size=6,
                                    radius_query_fraction_edge_length=0.6)
  st = partitioned.plan_statistics(g, 8)
  assert sum(st["mesh_owned"]) == 40962 and max(st["mesh_owned"]) - min(st["mesh_owned"]) <= 1
  assert sum(st["grid_owned"]) == 1038240
  assert sum(st["mesh_edges"]) == 327660 and sum(st["g2m_edges"]) == 1618818
  assert max(st["halo_bytes_per_step"]) <= 1.4e6
  print(st)


# ---- the same step on two real processes (gloo), with the exchange PartitionedEngine issues ----------

def _gloo_worker(rank, world, port, q):
  import os
  import torch.distributed as dist
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    torch.set_num_threads(2)
    g, params, x = _cases.small_case(c_in=31, n_out=23, msg_steps=2, batch=1)
    x = x.astype(np.float64)
    orc = oracle_gnn.Oracle(params, torch.float64)
    lg = partitioned.build_local_graph(g, world, rank)
    gd = lg.graph.as_dict()
    send_rows = torch.as_tensor(lg.send_rows, dtype=torch.int64)
    n_halo = int(lg.mesh_halo.size)

    def exchange(table):
      """PartitionedEngine.exchange_halo with torch ops in place of the pack / unpack kernels:
      gather the owned boundary rows (grouped by peer), ONE all_to_all_single with the plan's split
      sizes, store the received rows behind the 128-aligned owned block."""
      flat = table.reshape(table.shape[0], -1)
      send = flat[send_rows].contiguous()
      recv = torch.empty([n_halo, flat.shape[1]], dtype=flat.dtype)
      dist.all_to_all_single(recv, send, output_split_sizes=lg.recv_counts,
                             input_split_sizes=lg.send_counts)
      out = table.clone()
      out[lg.mesh_owned_pad:lg.mesh_owned_pad + n_halo] = recv.reshape((n_halo,) + tuple(table.shape[1:]))
      return out

    vm, vg = orc.encoder(gd, x[lg.local_grid_ids])
    v = exchange(vm)
    e = orc.processor_embed(gd, x.shape[1])
    for k in range(orc.num_message_steps()):
      v, e = orc.processor_step(gd, v, e, k)
      v = exchange(v)
    out = orc.decoder(gd, v, vg)[:lg.grid_owned.size]
    q.put((rank, lg.grid_owned, out.numpy()))
  finally:
    dist.destroy_process_group()


def test_partitioned_step_on_two_gloo_ranks_equals_the_full_step():
  import socket
  import torch.multiprocessing as mp
  s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  parts = [q.get(timeout=600) for _ in procs]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  g, params, x = _cases.small_case(c_in=31, n_out=23, msg_steps=2, batch=1)
  ref = oracle_gnn.Oracle(params, torch.float64).forward(g.as_dict(), x.astype(np.float64)).numpy()
  y = np.full_like(ref, np.nan)
  for _, owned, out in parts:
    y[owned] = out
  assert np.isfinite(y).all()
  assert np.abs(y - ref).max() / np.abs(ref).max() < 1e-12
