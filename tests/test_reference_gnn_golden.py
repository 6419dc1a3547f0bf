"""The oracle's GNN forward against the REFERENCE'S OWN wiring, executed.

tests/golden/reference_gnn_forward.npz was produced by importing and running the reference's
`GraphCast.__init__ / _maybe_init / _run_grid2mesh_gnn / _run_mesh_gnn / _run_mesh2grid_gnn`
(with `DeepTypedGraphNet`, `typed_graph_net.InteractionNetwork / GraphMapFeatures`,
`typed_graph`) on numpy stand-ins for jax / jraph / haiku / chex (tests/golden/numpy_standins.py:
Linear, LayerNorm, swish, segment_sum, concatenated_args, pytree flatten - a few lines each).
So everything the reference itself decides - concat orders [edge | sender | receiver] and
[node | aggregate], gathers, which node sets are updated in which GNN, the residual
connections, PRE-update node features in the edge update, the dead sender aggregation, the
output MLP without LayerNorm, the Haiku-style parameter paths - is pinned here; only the
third-party primitives are restated."""
import os

import numpy as np
import pytest
import torch

from graphcast_b200 import graph as graph_lib
from oracle import gnn as oracle_gnn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_gnn_forward.npz")


@pytest.fixture(scope="module")
def ref():
  with np.load(GOLDEN) as z:
    return {k: z[k] for k in z.files}


def _params(ref):
  params = {}
  for k, v in ref.items():
    if k.startswith("param:"):
      _, path, leaf = k.split(":")
      params.setdefault(path, {})[leaf] = v
  return params


def _graph_from_reference(ref):
  return {
      "grid_node_feats": ref["grid_node_feats"], "mesh_node_feats": ref["mesh_node_feats"],
      "g2m_senders": ref["g2m_senders"], "g2m_receivers": ref["g2m_receivers"],
      "g2m_edge_feats": ref["g2m_edge_feats"],
      "mesh_senders": ref["mesh_senders"], "mesh_receivers": ref["mesh_receivers"],
      "mesh_edge_feats": ref["mesh_edge_feats"],
      "m2g_senders": ref["m2g_senders"], "m2g_receivers": ref["m2g_receivers"],
      "m2g_edge_feats": ref["m2g_edge_feats"],
  }


def _rel(a, b):
  a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
  return np.abs(a - b).max() / np.abs(b).max()


def test_parameter_paths_are_the_ones_the_oracle_expects(ref):
  params = _params(ref)
  assert len(params) == 53                      # 18 (grid2mesh) + 21 (mesh, 3 steps) + 14 (mesh2grid)
  g = "mesh_gnn"
  for k in range(3):
    assert oracle_gnn.mlp_name(g, f"processor_edges_{k}_", "mesh") + "_mlp/~/linear_0" in params
    assert oracle_gnn.mlp_name(g, f"processor_nodes_{k}_", "mesh_nodes") + "_layer_norm" in params
  # the output MLP has no LayerNorm; the reference also builds a (dead) mesh-node update in mesh2grid
  assert oracle_gnn.mlp_name("mesh2grid_gnn", "decoder_nodes_", "grid_nodes") + "_mlp/~/linear_1" in params
  assert oracle_gnn.mlp_name("mesh2grid_gnn", "decoder_nodes_", "grid_nodes") + "_layer_norm" not in params
  assert oracle_gnn.mlp_name("mesh2grid_gnn", "processor_nodes_0_", "mesh_nodes") + "_mlp/~/linear_0" in params
  # first edge-MLP layer: [edge | sender | receiver] = 3 x latent rows
  w = params[oracle_gnn.mlp_name(g, "processor_edges_0_", "mesh") + "_mlp/~/linear_0"]["w"]
  assert w.shape == (96, 32)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.float64, 2e-5)])
def test_oracle_forward_matches_executed_reference(ref, dtype, tol):
  orc = oracle_gnn.Oracle(_params(ref), dtype)
  out, inter = orc.forward(_graph_from_reference(ref), ref["grid_features"],
                           return_intermediates=True)
  assert tuple(out.shape) == ref["output"].shape == (684, 2, 3)
  assert _rel(inter["vm1"].numpy(), ref["latent_mesh_after_grid2mesh"]) < tol
  assert _rel(inter["vg1"].numpy(), ref["latent_grid_after_grid2mesh"]) < tol
  assert _rel(inter["v_mesh"].numpy(), ref["latent_mesh_after_mesh_gnn"]) < tol
  assert _rel(out.numpy(), ref["output"]) < tol


def test_oracle_on_this_repos_static_graph_matches_executed_reference(ref):
  """Same, but with the graph built by graphcast_b200.graph.build_static_graph: the whole
  chain (connectivity, features, forward) against the reference.  Edge order inside a
  receiver may differ (the sums are order-independent up to fp32 rounding)."""
  g = graph_lib.build_static_graph(grid_lat=ref["grid_lat"], grid_lon=ref["grid_lon"],
                                   mesh_size=int(ref["mesh_size"]),
                                   radius_query_fraction_edge_length=0.6)
  orc = oracle_gnn.Oracle(_params(ref), torch.float32)
  out = orc.forward(g.as_dict(), ref["grid_features"])
  assert _rel(out.numpy(), ref["output"]) < 5e-5


def _api_dataset(ref, prefix):
  from graphcast_b200 import xarray_shim as xs
  names = [k[len(prefix) + 1:] for k in ref if k.startswith(prefix + ":")]
  return xs.Dataset({n: (tuple(str(d) for d in ref[f"{prefix}_dims:{n}"]), ref[f"{prefix}:{n}"])
                     for n in names})


def test_inputs_to_grid_node_features_matches_executed_reference(ref):
  """[Ng, B, C] assembly of `GraphCast._inputs_to_grid_node_features` (graphcast.py:680-699):
  inputs then forcings on the channel axis, node = lat * n_lon + lon."""
  from graphcast_b200 import model_utils
  inputs, forcings = _api_dataset(ref, "api_in"), _api_dataset(ref, "api_forcing")
  stacked = np.concatenate([model_utils.dataset_to_stacked(inputs),
                            model_utils.dataset_to_stacked(forcings, inputs.sizes)], -1)
  b, la, lo, c = stacked.shape
  x = np.transpose(stacked, (1, 2, 0, 3)).reshape(la * lo, b, c)
  np.testing.assert_array_equal(x, ref["grid_features"])


def test_api_level_oracle_matches_the_reference_call(ref):
  """The helper that gates the GPU `GraphCast.__call__` test (tests/test_gpu_model.py:
  `_oracle_for_api`) against the reference's own `GraphCast.__call__`, executed end to end on
  stand-in datasets: Dataset -> features -> three GNNs -> Dataset."""
  import dataclasses
  import test_gpu_model
  from graphcast_b200 import model_utils
  inputs, forcings = _api_dataset(ref, "api_in"), _api_dataset(ref, "api_forcing")
  template = _api_dataset(ref, "api_out")
  g = dataclasses.make_dataclass("G", ["d"])(_graph_from_reference(ref))
  g.as_dict = lambda: g.d
  y = test_gpu_model._oracle_for_api(None, None, inputs, forcings, _params(ref), g)   # [B, lat, lon, n_out]
  got = model_utils.stacked_to_dataset(y, template)
  for name in template.data_vars.keys():
    want = ref[f"api_out:{name}"]
    assert got[name].dims == tuple(str(d) for d in ref[f"api_out_dims:{name}"])
    assert _rel(np.asarray(got[name].data), want) < 2e-5
