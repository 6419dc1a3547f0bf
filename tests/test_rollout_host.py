"""Rollout driver host logic (no GPU): chunking, relative time coordinates,
next-input assembly (reference rollout.py:437-445, 453-457, 581-604) with a CPU
stand-in predictor, checked against a direct restatement."""
import numpy as np
import pytest

from graphcast_b200 import graphcast, rollout, synthetic
from graphcast_b200 import xarray_shim as xs


class PersistencePlusOne(graphcast.Predictor):
  """prediction = last input frame + 1 for every target (template decides names)."""

  def __init__(self):
    self.seen_times = []

  def __call__(self, inputs, targets_template, forcings, **kw):
    self.seen_times.append((tuple(inputs.coords["time"][1]), tuple(targets_template.coords["time"][1])))
    out = xs.Dataset(coords=targets_template.coords)
    for name, t in targets_template.data_vars.items():
      if name in inputs:
        last = inputs[name].isel(time=slice(-1, None))
        out[name] = xs.DataArray(np.asarray(last.values) + 1.0, last.dims).transpose(*t.dims)
      else:
        out[name] = xs.DataArray(np.full(t.shape, 5.0, np.float32), t.dims)
    return out


def _example(steps):
  inputs, template, forcings = synthetic.make_example(graphcast.TASK_13_PRECIP_OUT, 30.0,
                                                      num_target_steps=steps, seed=3)
  return inputs, template, forcings


def test_chunked_prediction_matches_direct_recursion():
  inputs, template, forcings = _example(4)
  pred = PersistencePlusOne()
  out = rollout.chunked_prediction(lambda rng, **kw: pred(**kw), rng=None, inputs=inputs,
                                   targets_template=template, forcings=forcings)
  assert out.sizes["time"] == 4
  np.testing.assert_array_equal(out.coords["time"][1], template.coords["time"][1])
  t0 = inputs.data_vars["2m_temperature"].values[:, -1]
  for k in range(4):
    np.testing.assert_allclose(out.data_vars["2m_temperature"].values[:, k], t0 + (k + 1), rtol=1e-6)
  # precipitation is a target but not an input in TASK_13_PRECIP_OUT -> predicted directly
  assert np.all(out.data_vars["total_precipitation_6hr"].values == 5.0)
  # every call saw the time coordinates of the first chunk (relative times)
  assert len(set(pred.seen_times)) == 1


def test_next_inputs_keep_last_two_frames_and_forcings():
  inputs, template, forcings = _example(2)
  frame = PersistencePlusOne()(inputs, template.isel(time=slice(0, 1)), None)
  nxt = rollout._get_next_inputs(inputs, frame.assign(forcings.isel(time=slice(0, 1))))
  assert set(nxt.keys()) == set(inputs.keys())
  g_prev = inputs.data_vars["geopotential"].values
  g_next = nxt.data_vars["geopotential"].values
  np.testing.assert_array_equal(g_next[:, 0], g_prev[:, 1])
  np.testing.assert_array_equal(g_next[:, 1], g_prev[:, 1] + 1.0)
  np.testing.assert_array_equal(nxt.data_vars["toa_incident_solar_radiation"].values[:, 1],
                                forcings.data_vars["toa_incident_solar_radiation"].values[:, 0])
  np.testing.assert_array_equal(nxt.data_vars["land_sea_mask"].values,
                                inputs.data_vars["land_sea_mask"].values)


def test_errors():
  inputs, template, forcings = _example(3)
  fn = lambda rng, **kw: PersistencePlusOne()(**kw)
  with pytest.raises(ValueError, match="evenly divide"):
    rollout.chunked_prediction(fn, None, inputs, template, forcings, num_steps_per_chunk=2)
  bad = template.assign_coords(time=np.array([6, 12, 24]) * np.timedelta64(1, "h"))
  with pytest.raises(ValueError, match="evenly spaced"):
    rollout.chunked_prediction(fn, None, inputs, bad, forcings)
  with pytest.raises(ValueError, match="replica_axis"):
    list(rollout.chunked_prediction_generator(fn, None, inputs, template, 1, forcings,
                                              pmap_devices=[0]))
  # an input with a time axis that is neither predicted nor forced
  inputs2 = inputs.copy()
  inputs2["mystery"] = inputs.data_vars["2m_temperature"]
  with pytest.raises(ValueError, match="not predicted or forced"):
    rollout.chunked_prediction(fn, None, inputs2, template, forcings)


def test_extend_targets_template():
  _, template, _ = _example(1)
  ext = rollout.extend_targets_template(template, 40)
  assert ext.sizes["time"] == 40
  assert ext.coords["time"][1][-1] == np.timedelta64(240, "h")
  assert ext.data_vars["temperature"].shape[1] == 40


def test_bfloat16_cast_wrapper_semantics():
  """casting.Bfloat16Cast: a GraphCast is returned itself, switched to the "bf16" mode (so that
  InputsAndResiduals still fuses); other predictors get a pass-through wrapper; enabled=False
  changes nothing (reference casting.py:31-65)."""
  from graphcast_b200 import casting, graphcast
  cfg = graphcast.ModelConfig(1.0, 5, 512, 16, 1, 0.6)
  model = graphcast.GraphCast(cfg, graphcast.TASK_13)
  assert model._precision == "bf16x3"
  assert casting.Bfloat16Cast(model, enabled=False) is model and model._precision == "bf16x3"
  wrapped = casting.Bfloat16Cast(model)
  assert wrapped is model and model._precision == "bf16"
  try:
    model.set_precision("fp8")
    assert False
  except ValueError:
    pass

  class Dummy(graphcast.Predictor):
    def __call__(self, inputs, targets_template, forcings, **kw):
      return ("called", inputs, targets_template, forcings, kw)

  w = casting.Bfloat16Cast(Dummy())
  assert isinstance(w, casting.Bfloat16Cast)
  assert w(1, 2, 3, flag=True) == ("called", 1, 2, 3, {"flag": True})


def test_autoregressive_predictor_unrolls_like_the_rollout():
  """autoregressive.Predictor (reference autoregressive.py:127-222): T target steps in one call =
  the chunked rollout with one step per chunk; validation errors as in the reference."""
  import pytest
  from graphcast_b200 import autoregressive, rollout
  from graphcast_b200 import xarray_shim as xs
  rng = np.random.default_rng(3)
  hour = np.timedelta64(6, "h")
  mk = lambda *s: rng.standard_normal(s).astype(np.float32)
  inputs = xs.Dataset({"t2m": (("batch", "time", "lat", "lon"), mk(1, 2, 3, 4)),
                       "toa": (("batch", "time", "lat", "lon"), mk(1, 2, 3, 4)),
                       "mask": (("lat", "lon"), mk(3, 4))},
                      coords={"time": np.array([-1, 0]) * hour})
  tt = (np.arange(3) + 1) * hour
  template = xs.Dataset({"t2m": (("batch", "time", "lat", "lon"), np.zeros((1, 3, 3, 4), np.float32))},
                        coords={"time": tt})
  forcings = xs.Dataset({"toa": (("batch", "time", "lat", "lon"), mk(1, 3, 3, 4))}, coords={"time": tt})

  class OneStep:
    def __call__(self, inputs, targets_template, forcings, scale=1.0):
      x = np.asarray(inputs["t2m"].data)
      f = np.asarray(forcings["toa"].data)
      y = scale * (0.8 * x[:, 1:] + 0.2 * x[:, :1]) + 0.1 * f + 0.01 * np.asarray(inputs["mask"].data)
      return xs.Dataset({"t2m": (inputs["t2m"].dims, y)}, coords={"time": targets_template.coords["time"]})

  ar = autoregressive.Predictor(OneStep(), gradient_checkpointing=True)
  got = ar(inputs, template, forcings, scale=0.5)
  want = rollout.chunked_prediction(
      lambda rng, inputs, targets_template, forcings: OneStep()(inputs, targets_template, forcings, scale=0.5),
      rng=None, inputs=inputs, targets_template=template, forcings=forcings)
  assert got["t2m"].dims == ("batch", "time", "lat", "lon") and got["t2m"].shape == (1, 3, 3, 4)
  np.testing.assert_array_equal(np.asarray(got["t2m"].data), np.asarray(want["t2m"].data))
  np.testing.assert_array_equal(np.asarray(got.coords["time"][1]), tt)
  with pytest.raises(ValueError, match="Time-dependent input variable"):
    ar(inputs, template, xs.Dataset(coords={"time": tt}))
  with pytest.raises(ValueError, match="both targets and forcings"):
    ar(inputs, template, xs.Dataset({"toa": forcings["toa"], "t2m": template["t2m"]}, coords={"time": tt}))


def test_extend_targets_template_keeps_batched_datetime():
  """`datetime` with dims (batch, time) -- what `generate_forcings` reads -- is extended along time and
  keeps its batch axis (utils/rollout.py:607-640 extends every time-indexed coordinate)."""
  template = xs.Dataset(
      {"t": xs.DataArray(np.zeros((1, 2, 3, 4), np.float32), ("batch", "time", "lat", "lon"))},
      coords={"time": (("time",), np.array([6, 12], "timedelta64[h]").astype("timedelta64[ns]")),
              "lat": (("lat",), np.arange(3.0)), "lon": (("lon",), np.arange(4.0))})
  t0 = np.datetime64("2021-03-17T06:00:00")
  dt = (t0 + np.asarray(template.coords["time"][1])).astype("datetime64[ns]")[None, :]
  template = template.assign_coords(datetime=(("batch", "time"), dt))
  out = rollout.extend_targets_template(template, 5)
  dims, vals = out.coords["datetime"]
  assert tuple(dims) == ("batch", "time") and np.asarray(vals).shape == (1, 5)
  step = np.timedelta64(6, "h")
  assert (np.diff(np.asarray(vals)[0]) == step).all() and np.asarray(vals)[0, 0] == t0 + step


def test_pinned_prediction_sink_refuses_host_resident_chunks():
  sink = rollout.PinnedPredictionSink(depth=3)
  chunk = xs.Dataset({"t": xs.DataArray(np.zeros((1, 1, 3, 4), np.float32), ("batch", "time", "lat", "lon"))})
  with pytest.raises(TypeError):
    sink(chunk)
  sink.wait()                       # nothing queued: a no-op
