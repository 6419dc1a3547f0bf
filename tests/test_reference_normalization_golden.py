"""`normalization.InputsAndResiduals` (generic Dataset path) against the reference's own
`weathernext/utils/normalization.py`, which tests/golden/make_golden.py executed unmodified on
stand-in datasets (numpy arrays with named dimensions): what the wrapped predictor is given
(normalised inputs and forcings, level-wise statistics broadcast by name) and what comes out
(residual targets: y * diffs_std + last input frame; target-only variables: y * std + mean).
The fused CUDA path is checked against this generic path in tests/test_gpu_model.py."""
import os

import numpy as np
import pytest

from graphcast_b200 import normalization
from graphcast_b200 import xarray_shim as xs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_normalization.npz")


@pytest.fixture(scope="module")
def ref():
  with np.load(GOLDEN) as z:
    return {k: z[k] for k in z.files}


def _dims(ref, key):
  return tuple(str(d) for d in ref[key])


def _dataset(ref, prefix):
  names = [k[len(prefix) + 1:] for k in ref if k.startswith(prefix + ":")]
  return xs.Dataset({n: (_dims(ref, f"{prefix}_dims:{n}"), ref[f"{prefix}:{n}"]) for n in names})


def test_inputs_and_residuals_matches_executed_reference(ref):
  inputs, forcings = _dataset(ref, "in"), _dataset(ref, "forcing")
  names = [k[len("template_dims:"):] for k in ref if k.startswith("template_dims:")]
  template = xs.Dataset({n: (_dims(ref, f"template_dims:{n}"),
                             np.zeros(tuple(ref[f"template_shape:{n}"]), np.float32)) for n in names})
  stats = {s: _dataset(ref, s) for s in ("mean", "std", "diffs_std")}
  seen = {}

  def inner(norm_inputs, targets_template, forcings):
    seen["inputs"], seen["forcings"] = norm_inputs, forcings
    out = {}
    for name in targets_template.data_vars.keys():
      if name in norm_inputs:
        v = norm_inputs[name]
        out[name] = (v.dims, 0.5 * np.asarray(v.data)[:, -1:] + 0.25)
      else:
        t = targets_template[name]
        shape = tuple(np.asarray(t.data).shape)
        out[name] = (t.dims, np.full(shape, 0.125, np.float32) * (1 + np.arange(shape[-1], dtype=np.float32)))
    return xs.Dataset(out)

  wrapped = normalization.InputsAndResiduals(inner, stddev_by_level=stats["std"],
                                             mean_by_level=stats["mean"],
                                             diffs_stddev_by_level=stats["diffs_std"])
  got = wrapped(inputs, template, forcings)
  for name in inputs.data_vars.keys():
    np.testing.assert_allclose(np.asarray(seen["inputs"][name].data), ref[f"norm_in:{name}"],
                               rtol=1e-6, atol=1e-6)
  for name in forcings.data_vars.keys():
    np.testing.assert_allclose(np.asarray(seen["forcings"][name].data), ref[f"norm_forcing:{name}"],
                               rtol=1e-6, atol=1e-6)
  assert sorted(got.data_vars.keys()) == sorted(names)
  for name in names:
    assert got[name].dims == _dims(ref, f"out_dims:{name}")
    np.testing.assert_allclose(np.asarray(got[name].data), ref[f"out:{name}"], rtol=1e-6, atol=1e-6)
