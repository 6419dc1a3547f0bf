"""Bounded CPU sample of one GraphCast step on the REAL workload (test / measurement infrastructure).

TEST INFRASTRUCTURE -- see oracle/__init__.py: only tests, smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this.

A full fp32 oracle step at 0.25 degree is 29.3 TFLOP: minutes of host time.  `bench.py` needs a CPU
number that finishes in seconds per sample WITHOUT switching to a toy graph (round 1 extrapolated
from a 4 degree instance by FLOP ratio; two such extrapolations disagreed 4x).  This module times
every stage of the step -- the same Oracle.mlp / gather / segment_sum calls as `Oracle.forward`
(oracle/gnn.py, following weathernext1_graph/graphcast.py:550-678 and
utils/typed_graph_net.py:369-546) -- on a contiguous block of `fraction` of that stage's rows of the
real graph: real index arrays (so the gathers have the real locality), full-size gather tables and
aggregation targets, the real layer widths.  All stages are row-parallel (per-row MLPs; gathers and
the segment sum cost per edge), so   step time ~= sample time / fraction.   The latent tables the
gathers read are random (the arithmetic cost does not depend on the values).

`validate()` checks that premise where the full step is affordable: a full `Oracle.forward` against
sample / fraction on the same graph.
"""

from __future__ import annotations

import time
from typing import Dict, Mapping, Tuple

import numpy as np
import torch

from oracle import gnn


def _block(n: int, fraction: float) -> slice:
  k = max(1, int(round(n * fraction)))
  return slice(0, min(n, k))


class SampledStep:
  """Pre-allocates the full-size tables once; `run()` executes one sampled step."""

  def __init__(self, graph: Mapping[str, np.ndarray], params: gnn.Params, c_in: int,
               fraction: float, seed: int = 0):
    self.orc = gnn.Oracle(params, torch.float32)
    self.f = float(fraction)
    g = graph
    t = lambda a: torch.as_tensor(np.asarray(a)).to(torch.float32)
    idx = lambda a: torch.as_tensor(np.asarray(a)).to(torch.int64)
    gen = torch.Generator().manual_seed(seed)
    rnd = lambda *shape: torch.randn(*shape, generator=gen)
    self.ng, self.nm = g["grid_node_feats"].shape[0], g["mesh_node_feats"].shape[0]
    ng, nm = self.ng, self.nm
    self.bg, self.bm = _block(ng, fraction), _block(nm, fraction)
    # inputs of the sampled rows, with the structural features (graphcast.py:561-568, 573-583)
    self.grid_in = torch.cat([rnd(self.bg.stop, 1, c_in), t(g["grid_node_feats"])[self.bg, None, :]], -1)
    self.mesh_in = torch.cat([torch.zeros(self.bm.stop, 1, c_in), t(g["mesh_node_feats"])[self.bm, None, :]], -1)
    # full-size latent tables the gathers read / the segment sums write
    self.vg, self.vm = rnd(ng, 1, 512), rnd(nm, 1, 512)
    self.edges: Dict[str, Tuple] = {}
    for name in ("g2m", "mesh", "m2g"):
      s, r = idx(g[f"{name}_senders"]), idx(g[f"{name}_receivers"])
      b = _block(s.shape[0], fraction)
      self.edges[name] = (s[b], r[b], t(g[f"{name}_edge_feats"])[b, None, :], rnd(b.stop, 1, 512))
    self.steps = 0
    k = 0
    while gnn.mlp_name("mesh_gnn", f"processor_edges_{k}_", "mesh") + "_mlp/~/linear_0" in self.orc.p:
      k += 1
    self.steps = k

  def run(self) -> None:
    o, mn = self.orc, gnn.mlp_name
    # ---- grid2mesh_gnn (graphcast.py:550-604) ----
    g = "grid2mesh_gnn"
    o.mlp(mn(g, "encoder_nodes_", "grid_nodes"), [self.grid_in])
    o.mlp(mn(g, "encoder_nodes_", "mesh_nodes"), [self.mesh_in])
    s, r, feat, e = self.edges["g2m"]
    o.mlp(mn(g, "encoder_edges_", "grid2mesh"), [feat])
    m = o.mlp(mn(g, "processor_edges_0_", "grid2mesh"), [e, self.vg[s], self.vm[r]])
    agg = o.segment_sum(m, r, self.nm)
    o.mlp(mn(g, "processor_nodes_0_", "mesh_nodes"), [self.vm[self.bm], agg[self.bm]])
    o.mlp(mn(g, "processor_nodes_0_", "grid_nodes"), [self.vg[self.bg]])
    # ---- mesh_gnn (graphcast.py:606-639) ----
    g = "mesh_gnn"
    s, r, feat, e = self.edges["mesh"]
    o.mlp(mn(g, "encoder_edges_", "mesh"), [feat])
    for k in range(self.steps):
      m = o.mlp(mn(g, f"processor_edges_{k}_", "mesh"), [e, self.vm[s], self.vm[r]])
      agg = o.segment_sum(m, r, self.nm)
      o.mlp(mn(g, f"processor_nodes_{k}_", "mesh_nodes"), [self.vm[self.bm], agg[self.bm]])
      e = e + m
    # ---- mesh2grid_gnn (graphcast.py:641-678) ----
    g = "mesh2grid_gnn"
    s, r, feat, e = self.edges["m2g"]
    o.mlp(mn(g, "encoder_edges_", "mesh2grid"), [feat])
    m = o.mlp(mn(g, "processor_edges_0_", "mesh2grid"), [e, self.vm[s], self.vg[r]])
    agg = o.segment_sum(m, r, self.ng)
    v = self.vg[self.bg] + o.mlp(mn(g, "processor_nodes_0_", "grid_nodes"), [self.vg[self.bg], agg[self.bg]])
    o.mlp(mn(g, "decoder_nodes_", "grid_nodes"), [v], use_layer_norm=False)

  def time_one(self) -> float:
    t0 = time.perf_counter()
    self.run()
    return time.perf_counter() - t0


def validate(graph, params, c_in: int, fraction: float, reps: int = 2):
  """(full step seconds, sampled seconds / fraction): the premise of the sampling."""
  x = np.random.default_rng(0).standard_normal((graph["grid_node_feats"].shape[0], 1, c_in)).astype(np.float32)
  orc = gnn.Oracle(params, torch.float32)
  orc.forward(graph, x)
  t0 = time.perf_counter()
  orc.forward(graph, x)
  full = time.perf_counter() - t0
  samp = SampledStep(graph, params, c_in, fraction)
  samp.run()
  t = min(samp.time_one() for _ in range(reps))
  return full, t / fraction
