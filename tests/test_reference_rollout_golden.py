"""`rollout.chunked_prediction_generator` / `_get_next_inputs` against the reference's own
`weathernext/utils/rollout.py`, executed unmodified on coordinate-aware stand-in datasets
(tests/golden/make_golden.py).  A recording predictor whose output depends on BOTH input frames,
on the forcing of the target time and on a static input makes the trajectory sensitive to the
feeding logic: which frames become the next inputs, which forcing slice each step receives,
static variables passing through, chunk-relative time coordinates for the predictor and absolute
ones on the yielded chunks."""
import os

import numpy as np
import pytest

from graphcast_b200 import rollout
from graphcast_b200 import xarray_shim as xs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_rollout.npz")
HOUR = np.timedelta64(1, "h")


@pytest.fixture(scope="module")
def ref():
  with np.load(GOLDEN) as z:
    return {k: z[k] for k in z.files}


def _dataset(ref, prefix, times):
  names = [k[len(prefix) + 1:] for k in ref if k.startswith(prefix + ":")]
  return xs.Dataset({n: (tuple(str(d) for d in ref[f"{prefix}_dims:{n}"]), ref[f"{prefix}:{n}"])
                     for n in names}, coords={"time": times})


@pytest.mark.parametrize("steps_per_chunk", [1])
def test_chunked_rollout_matches_executed_reference(ref, steps_per_chunk):
  in_times = ref["in_times"] * HOUR
  tgt_times = ref["target_times"] * HOUR
  inputs = _dataset(ref, "in", in_times)
  forcings = _dataset(ref, "forcing", tgt_times)
  template = _dataset(ref, "template", tgt_times)
  calls = []

  def predictor(rng, inputs, targets_template, forcings):
    calls.append((np.asarray(inputs.coords["time"][1]).copy(),
                  np.asarray(targets_template.coords["time"][1]).copy()))
    f = np.asarray(forcings["toa_incident_solar_radiation"].data)
    mask = np.asarray(inputs["land_sea_mask"].data)
    out = {}
    for name in targets_template.data_vars.keys():
      v = inputs[name]
      x = np.asarray(v.data)
      fb = f.reshape(f.shape[:2] + (1,) * (x.ndim - 4) + f.shape[2:])
      out[name] = (v.dims, 0.9 * x[:, 1:] + 0.1 * x[:, :1] + 0.05 * fb + 0.01 * mask)
    return xs.Dataset(out, coords={"time": targets_template.coords["time"]})

  chunks = list(rollout.chunked_prediction_generator(
      predictor, rng=0, inputs=inputs, targets_template=template,
      num_steps_per_chunk=steps_per_chunk, forcings=forcings))
  assert len(chunks) == 4
  for i, chunk in enumerate(chunks):
    for name in template.data_vars.keys():
      np.testing.assert_allclose(np.asarray(chunk[name].data), ref[f"chunk{i}:{name}"],
                                 rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(np.asarray(chunk.coords["time"][1]) // HOUR, ref[f"chunk{i}_time"])
    np.testing.assert_array_equal(calls[i][0] // HOUR, ref[f"call{i}_in_time"])
    np.testing.assert_array_equal(calls[i][1] // HOUR, ref[f"call{i}_target_time"])


ENSEMBLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                        "reference_rollout_ensemble.npz")


def _recording_predictor(rng, inputs, targets_template, forcings):
  f = np.asarray(forcings["toa_incident_solar_radiation"].data)
  mask = np.asarray(inputs["land_sea_mask"].data)
  out = {}
  for name in targets_template.data_vars.keys():
    v = inputs[name]
    x = np.asarray(v.data)
    fb = f.reshape(f.shape[:2] + (1,) * (x.ndim - 4) + f.shape[2:])
    out[name] = (v.dims, 0.9 * x[:, 1:] + 0.1 * x[:, :1] + 0.05 * fb + 0.01 * mask)
  return xs.Dataset(out, coords={"time": targets_template.coords["time"]})


def test_ensemble_driver_matches_executed_reference_and_shards_over_ranks():
  """`chunked_prediction_generator_multiple_runs` (reference rollout.py:158-306, non-pmap branch
  executed on stand-ins): member order, the "sample" slicing of inputs and forcings, the sample
  coordinate on every chunk.  With rank / world_size the same members are split over ranks with
  no communication; the union over ranks is the single-rank result."""
  with np.load(ENSEMBLE) as z:
    ref = {k: z[k] for k in z.files}
  ns = int(ref["num_samples"])
  in_times, tgt_times = ref["in_times"] * HOUR, ref["target_times"] * HOUR
  inputs = _dataset(ref, "in", in_times)
  forcings = _dataset(ref, "forcing", tgt_times)
  template = _dataset(ref, "template", tgt_times)
  kw = dict(rngs=np.arange(ns), inputs=inputs, targets_template=template, forcings=forcings,
            num_samples=ns, num_steps_per_chunk=1)

  def check(chunks, first_index):
    for j, chunk in enumerate(chunks):
      i = first_index + j
      assert int(np.asarray(chunk.coords["sample"][1])) == int(ref[f"chunk{i}_sample"])
      np.testing.assert_array_equal(np.asarray(chunk.coords["time"][1]) // HOUR, ref[f"chunk{i}_time"])
      for name in template.data_vars.keys():
        np.testing.assert_allclose(np.asarray(chunk[name].data), ref[f"chunk{i}:{name}"],
                                   rtol=1e-6, atol=1e-6)

  chunks = list(rollout.chunked_prediction_generator_multiple_runs(_recording_predictor, **kw))
  assert len(chunks) == ns * 4
  check(chunks, 0)
  # two ranks: members [0, 1] and [2]
  r0 = list(rollout.chunked_prediction_generator_multiple_runs(_recording_predictor, rank=0, world_size=2, **kw))
  r1 = list(rollout.chunked_prediction_generator_multiple_runs(_recording_predictor, rank=1, world_size=2, **kw))
  assert (len(r0), len(r1)) == (8, 4)
  check(r0, 0)
  check(r1, 8)
  with pytest.raises(ValueError, match="rngs"):
    list(rollout.chunked_prediction_generator_multiple_runs(
        _recording_predictor, **dict(kw, rngs=np.arange(ns + 1))))
  with pytest.raises(ValueError, match="pmap_devices"):
    list(rollout.chunked_prediction_generator_multiple_runs(
        _recording_predictor, pmap_devices=[0], **kw))
