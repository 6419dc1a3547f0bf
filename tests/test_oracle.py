"""Oracle self-checks.  The GNN forward is pinned against the EXECUTED reference wiring in
tests/test_reference_gnn_golden.py and tests/test_reference_gnn_latent512.py; the tests here check
it against independent formulations, a committed regression vector, and place the reduced-precision
emulations (the product's "bf16" mode, the reference's own bf16 execution) against the exact result."""
import os

import numpy as np
import torch

import _cases
from oracle import gnn


def test_layer_norm_and_swish_definitions():
  x = torch.randn(7, 33, dtype=torch.float64)
  s, o = torch.rand(33, dtype=torch.float64) + 0.5, torch.randn(33, dtype=torch.float64)
  want = torch.nn.functional.layer_norm(x, (33,), s, o, eps=1e-5)
  np.testing.assert_allclose(gnn.layer_norm(x, s, o).numpy(), want.numpy(), atol=1e-12)
  np.testing.assert_allclose(gnn.swish(x).numpy(), torch.nn.functional.silu(x).numpy(), atol=1e-12)


def test_segment_sum_matches_loop():
  data = torch.randn(20, 2, 5, dtype=torch.float64)
  ids = torch.tensor(np.random.default_rng(0).integers(0, 6, 20))
  got = gnn.Oracle.segment_sum(data, ids, 6).numpy()
  want = np.zeros((6, 2, 5))
  for e in range(20):
    want[int(ids[e])] += data[e].numpy()
  np.testing.assert_allclose(got, want, atol=1e-12)


def test_edge_order_invariance_and_batch_independence():
  g, params, x = _cases.small_case(c_in=11, n_out=7, msg_steps=2, batch=2)
  gd = g.as_dict()
  o = gnn.Oracle(params, torch.float64)
  y = o.forward(gd, x).numpy()
  # permuting the edges of every edge set must not change the result
  rng = np.random.default_rng(1)
  gp = dict(gd)
  for pre in ("g2m", "mesh", "m2g"):
    p = rng.permutation(gd[f"{pre}_senders"].shape[0])
    for k in ("senders", "receivers", "edge_feats"):
      gp[f"{pre}_{k}"] = gd[f"{pre}_{k}"][p]
  np.testing.assert_allclose(o.forward(gp, x).numpy(), y, atol=1e-9)
  # batch elements are independent
  y0 = o.forward(gd, x[:, :1]).numpy()
  np.testing.assert_allclose(y[:, :1], y0, atol=1e-12)


def test_fp32_vs_fp64_and_regression_vector():
  g, params, x = _cases.small_case(c_in=11, n_out=7, msg_steps=2)
  y64 = gnn.Oracle(params, torch.float64).forward(g.as_dict(), x).numpy()
  y32 = gnn.Oracle(params, torch.float32).forward(g.as_dict(), x).numpy()
  assert np.abs(y32 - y64).max() / np.abs(y64).max() < 1e-5
  path = os.path.join(os.path.dirname(__file__), "golden", "oracle_small_case.npz")
  if not os.path.exists(path):          # first run creates the regression vector
    np.savez_compressed(path, rows=np.arange(0, y64.shape[0], 97), y=y64[::97])
  ref = np.load(path)
  np.testing.assert_allclose(y64[ref["rows"]], ref["y"], rtol=1e-9, atol=1e-9)


def test_param_inventory_matches_survey_appendix_b():
  p = gnn.init_params(c_in=471, n_out=227, msg_steps=16)
  n = sum(int(np.prod(a.shape)) for v in p.values() for a in v.values())
  assert n == 36348131          # SURVEY appendix B total (includes the dead mesh2grid mesh-node MLP)
  assert p["mesh_gnn/~_networks_builder/processor_edges_3_mesh_mlp/~/linear_0"]["w"].shape == (1536, 512)
  assert p["mesh2grid_gnn/~_networks_builder/decoder_nodes_grid_nodes_mlp/~/linear_1"]["w"].shape == (512, 227)
  assert "mesh2grid_gnn/~_networks_builder/decoder_nodes_grid_nodes_layer_norm" not in p


def test_identity_hooks_do_not_change_the_forward():
  """The rounding hooks of the base class are identities: a subclass that spells them out
  reproduces Oracle.forward bit for bit (the pinned goldens therefore cover the hooked code)."""
  class Spelled(gnn.Oracle):
    def bias_add(self, x, b): return x + b
    def activation(self, x): return x * torch.sigmoid(x)
    def normalize(self, x, scale, offset): return gnn.layer_norm(x, scale, offset)
    def add(self, a, b): return a + b
    def aggregate(self, name, data, ids, n): return gnn.Oracle.segment_sum(data, ids, n)
  g, params, x = _cases.small_case(c_in=11, n_out=7, msg_steps=2)
  a = gnn.Oracle(params, torch.float64).forward(g.as_dict(), x)
  b = Spelled(params, torch.float64).forward(g.as_dict(), x)
  assert torch.equal(a, b)


def test_bf16_mode_is_at_least_as_accurate_as_the_reference_bf16_execution():
  """`precision="bf16"` (bf16 operands, fp32 everything else: Bf16OperandOracle) against the
  reference's own execution under casting.Bfloat16Cast (every op result rounded to bfloat16,
  fp32 aggregation only in grid2mesh: ReferenceBf16Oracle), both measured against the exact fp64
  step: the product's mode must not be the less accurate of the two, and both stay far outside the
  1e-4 gate of the parity mode (which is why bf16x3 is the default)."""
  g, params, x = _cases.small_case(c_in=31, n_out=23, msg_steps=4, randomize_affine=False)
  gd = g.as_dict()
  exact = gnn.Oracle(params, torch.float64).forward(gd, x)
  scale = exact.abs().max().item()
  err = lambda y: (y.to(torch.float64) - exact).abs().max().item() / scale
  ours = err(gnn.Bf16OperandOracle(params, torch.float32).forward(gd, x))
  ref = gnn.ReferenceBf16Oracle(params)
  theirs = err(ref.forward(gd, x))
  print(f"bf16 mode {ours:.2e}  reference bf16 execution {theirs:.2e}")
  assert 1e-4 < ours < 3e-2 and 1e-4 < theirs < 1e-1
  assert ours <= theirs
  # every value the reference-style execution produces is a bfloat16 number
  y, inter = ref.forward(gd, x, return_intermediates=True)
  for name in ("vm1", "vg1", "v_mesh", "vg2"):
    t = inter[name]
    assert torch.equal(t, t.to(torch.bfloat16).to(torch.float32)), name
  assert torch.equal(y, y.to(torch.bfloat16).to(torch.float32))
