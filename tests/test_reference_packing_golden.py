"""Channel packing against the reference's own `dataset_to_stacked` / `stacked_to_dataset`
(weathernext/utils/model_utils.py:645-776), which tests/golden/make_golden.py ran unmodified on a
numpy stand-in for the few xarray methods they use.  Pins the channel order of the model's
inputs and outputs (variables sorted by name; non-(batch, lat, lon) dims flattened in the
variable's own dim order; static variables broadcast over batch) for the host packing functions
and for the channel layout the CUDA pack / unpack kernels are driven by."""
import os

import numpy as np
import pytest

from graphcast_b200 import model_utils
from graphcast_b200 import xarray_shim as xs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_packing.npz")


@pytest.fixture(scope="module")
def ref():
  with np.load(GOLDEN) as z:
    return {k: z[k] for k in z.files}


def _dataset(ref, prefix):
  names = [k[len(prefix) + 1:] for k in ref if k.startswith(prefix + ":")]
  return xs.Dataset({n: (tuple(str(d) for d in ref[f"{prefix}_dims:{n}"]), ref[f"{prefix}:{n}"])
                     for n in names})


def test_dataset_to_stacked_matches_reference(ref):
  ds = _dataset(ref, "in")
  got = model_utils.dataset_to_stacked(ds)
  assert got.shape == ref["stacked_inputs"].shape
  np.testing.assert_array_equal(got, ref["stacked_inputs"])


def test_channel_layout_matches_reference_order(ref):
  ds = _dataset(ref, "in")
  slabs = model_utils.channel_layout(ds)
  assert [s.name for s in slabs] == sorted(n for n in ds.data_vars.keys())
  stacked = ref["stacked_inputs"]                       # [batch, lat, lon, channels]
  for s in slabs:
    var = ds.data_vars[s.name]
    planes = np.asarray(model_utils.variable_to_planes(var, ds.sizes))   # [B, nch, lat, lon]
    want = np.transpose(stacked[..., s.start:s.start + s.count], (0, 3, 1, 2))
    np.testing.assert_array_equal(planes, want)
  # time-major, level-minor inside a variable (reference :668-674)
  geo = next(s for s in slabs if s.name == "geopotential")
  assert geo.stack_dims == ("time", "level") and geo.count == 6


def test_stacked_to_dataset_matches_reference(ref):
  tmpl = _dataset(ref, "out")
  got = model_utils.stacked_to_dataset(ref["stacked_outputs"], tmpl)
  assert sorted(got.data_vars.keys()) == sorted(tmpl.data_vars.keys())
  for name in tmpl.data_vars.keys():
    assert got[name].dims == tuple(str(d) for d in ref[f"out_dims:{name}"])
    np.testing.assert_array_equal(np.asarray(got[name].data), ref[f"out:{name}"])
