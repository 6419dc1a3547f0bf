"""The CUDA path against the EXECUTED REFERENCE, directly.

tests/golden/reference_gnn_forward_latent512.npz holds the output of the reference's own
`GraphCast._run_grid2mesh_gnn / _run_mesh_gnn / _run_mesh2grid_gnn` (run on numpy stand-ins for
jax / jraph / haiku, see tests/golden/make_golden.py) at the kernels' width, latent 512, on the
reference's own 10-degree graph: 684 grid nodes, 162 mesh nodes, 2 message-passing steps,
batch 2.  The 40 MB of parameters are not stored: they are regenerated from the stored seed and
creation manifest (`numpy_standins.regenerate`, checked bit-exact when the golden was written).

  * CPU: the oracle reproduces the executed reference (also validates the regeneration);
  * GPU: `engine.Engine` through the C ABI reproduces it within the parity bar (1e-4 for the
    bf16x3 mode, 1e-5 for the fp32 arm), with no oracle in between."""
import os
import sys

import numpy as np
import pytest
import torch

from graphcast_b200 import graph as graph_lib
from oracle import gnn as oracle_gnn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import numpy_standins  # noqa: E402  (parameter regeneration only)

GOLDEN = os.path.join(HERE, "golden", "reference_gnn_forward_latent512.npz")


@pytest.fixture(scope="module")
def ref():
  with np.load(GOLDEN) as z:
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def params(ref):
  manifest = [(p, k, a, b) for p, k, (a, b) in zip(ref["manifest_path"], ref["manifest_kind"],
                                                   ref["manifest_shape"])]
  return numpy_standins.regenerate(manifest, int(ref["seed"]))


def _static_graph(ref):
  f32 = lambda a: np.ascontiguousarray(a, np.float32)
  i32 = lambda a: np.ascontiguousarray(a, np.int32)
  return graph_lib.StaticGraph(
      num_grid_nodes=int(ref["grid_node_feats"].shape[0]),
      num_mesh_nodes=int(ref["mesh_node_feats"].shape[0]),
      grid_lat=f32(ref["grid_lat"]), grid_lon=f32(ref["grid_lon"]),
      grid_node_feats=f32(ref["grid_node_feats"]), mesh_node_feats=f32(ref["mesh_node_feats"]),
      g2m_senders=i32(ref["g2m_senders"]), g2m_receivers=i32(ref["g2m_receivers"]),
      g2m_edge_feats=f32(ref["g2m_edge_feats"]),
      mesh_senders=i32(ref["mesh_senders"]), mesh_receivers=i32(ref["mesh_receivers"]),
      mesh_edge_feats=f32(ref["mesh_edge_feats"]),
      m2g_senders=i32(ref["m2g_senders"]), m2g_receivers=i32(ref["m2g_receivers"]),
      m2g_edge_feats=f32(ref["m2g_edge_feats"]))


def _rel(a, b):
  a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
  return np.abs(a - b).max() / np.abs(b).max()


def test_oracle_matches_executed_reference_at_latent_512(ref, params):
  assert len(params) == 18 + 15 + 14
  w = params[oracle_gnn.mlp_name("mesh_gnn", "processor_edges_1_", "mesh") + "_mlp/~/linear_0"]["w"]
  assert w.shape == (1536, 512)
  orc = oracle_gnn.Oracle(params, torch.float32)
  out = orc.forward(_static_graph(ref).as_dict(), ref["grid_features"])
  assert _rel(out.numpy(), ref["output"]) < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("bf16x3", 1e-4), ("fp32_simt", 1e-5)])
def test_cuda_engine_matches_executed_reference(ref, params, precision, tol):
  from graphcast_b200 import engine
  g = _static_graph(ref)
  c_in, n_out = ref["grid_features"].shape[-1], ref["output"].shape[-1]
  eng = engine.Engine(g, params, c_in=c_in, n_out=n_out, msg_steps=int(ref["gnn_msg_steps"]),
                      precision=precision)
  y = eng.forward_features(torch.as_tensor(ref["grid_features"])).cpu().numpy()
  err = _rel(y, ref["output"])
  print(f"CUDA {precision} vs executed reference: max-abs relative error {err:.3e}")
  assert err <= tol
