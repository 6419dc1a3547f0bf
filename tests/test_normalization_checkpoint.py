"""Normalisation wrapper (generic path + the per-channel constants of the fused
path) and checkpoint round trip (reference checkpoint_test.py:65-121 style)."""
import dataclasses
import io
from typing import Any, Optional

import numpy as np
import pytest
import torch

from graphcast_b200 import checkpoint, graphcast, model_utils, normalization, synthetic
from graphcast_b200 import xarray_shim as xs


def _stats(task, seed):
  rng = np.random.default_rng(seed)
  levels = np.asarray(task.pressure_levels)
  ds = xs.Dataset(coords={"level": levels})
  for name in set(task.input_variables) | set(task.target_variables):
    if name in graphcast.variables.ALL_ATMOSPHERIC_VARS:
      ds[name] = xs.DataArray(rng.uniform(0.5, 2.0, len(levels)).astype(np.float32), ("level",))
    else:
      ds[name] = xs.DataArray(np.float32(rng.uniform(0.5, 2.0)), ())
  return ds


class Echo(graphcast.Predictor):
  """Returns normalised 'predictions' = 0.5 everywhere, records what it was given."""

  def __call__(self, inputs, targets_template, forcings, **kw):
    self.inputs, self.forcings = inputs, forcings
    out = xs.Dataset(coords=targets_template.coords)
    for name, t in targets_template.data_vars.items():
      out[name] = xs.DataArray(np.full(t.shape, 0.5, np.float32), t.dims)
    return out


def test_generic_path_semantics():
  task = graphcast.TASK_13_PRECIP_OUT
  inputs, template, forcings = synthetic.make_example(task, 30.0, seed=1)
  std, mean, dstd = _stats(task, 1), _stats(task, 2), _stats(task, 3)
  inner = Echo()
  out = normalization.InputsAndResiduals(inner, std, mean, dstd)(inputs, template, forcings)
  # inputs were normalised per level
  t_in = inputs.data_vars["temperature"].values
  want = (t_in - mean.data_vars["temperature"].values[None, None, :, None, None]) / \
      std.data_vars["temperature"].values[None, None, :, None, None]
  np.testing.assert_allclose(inner.inputs.data_vars["temperature"].values, want, rtol=1e-6)
  # residual target: 0.5 * diffs_std + last input
  want = 0.5 * dstd.data_vars["temperature"].values[None, None, :, None, None] + t_in[:, -1:]
  np.testing.assert_allclose(out.data_vars["temperature"].values, want, rtol=1e-6)
  # non-input target: 0.5 * std + mean
  want = 0.5 * std.data_vars["total_precipitation_6hr"].values + mean.data_vars["total_precipitation_6hr"].values
  np.testing.assert_allclose(out.data_vars["total_precipitation_6hr"].values, want, rtol=1e-6)
  two, _, _ = synthetic.make_example(task, 30.0, num_target_steps=2)
  _, template2, forcings2 = synthetic.make_example(task, 30.0, num_target_steps=2)
  with pytest.raises(ValueError, match="single timestep"):
    normalization.InputsAndResiduals(Echo(), std, mean, dstd)(inputs, template2, forcings2)


def test_fused_constants_agree_with_generic_arithmetic():
  task = graphcast.TASK_13_PRECIP_OUT
  inputs, template, forcings = synthetic.make_example(task, 30.0, seed=1)
  std, mean, dstd = _stats(task, 1), _stats(task, 2), _stats(task, 3)
  wrap = normalization.InputsAndResiduals(Echo(), std, mean, dstd)
  c = wrap._fused_constants(inputs, template, forcings, torch.device("cpu"))
  in_slabs = model_utils.channel_layout(inputs)
  n_in = sum(s.count for s in in_slabs)
  stacked_in = np.concatenate([model_utils.dataset_to_stacked(inputs),
                               model_utils.dataset_to_stacked(forcings, inputs.sizes)], -1)
  norm_in = np.concatenate([
      model_utils.dataset_to_stacked(normalization.normalize(inputs, std, mean)),
      model_utils.dataset_to_stacked(normalization.normalize(forcings, std, mean), inputs.sizes)], -1)
  fused = (stacked_in - c.in_mean.numpy()) / c.in_scale.numpy()
  np.testing.assert_allclose(fused, norm_in, rtol=1e-5, atol=1e-6)
  # output side: y*scale + offset + add_plane
  y = np.full((1, 7, 12, 83), 0.5, np.float32)
  add = np.where(c.add_plane_index.numpy() >= 0,
                 stacked_in[..., np.maximum(c.add_plane_index.numpy(), 0)], 0.0)
  fused_out = y * c.out_scale.numpy() + c.out_offset.numpy() + add
  want = model_utils.dataset_to_stacked(wrap(inputs, template, forcings))
  np.testing.assert_allclose(fused_out, want, rtol=1e-5, atol=1e-6)
  assert (c.add_plane_index.numpy() < 0).sum() == 1        # precipitation only


@dataclasses.dataclass
class Sub:
  a: int
  b: np.ndarray


@dataclasses.dataclass
class Tree:
  sub: Sub
  name: str
  items: tuple[int, ...]
  table: dict[str, Any]
  maybe: Optional[float] = None


def test_checkpoint_round_trip_and_format():
  t = Tree(sub=Sub(3, np.arange(6).reshape(2, 3)), name="x", items=(1, 2, 3),
           table={"p/q": {"w": np.ones(2)}}, maybe=None)
  buf = io.BytesIO()
  checkpoint.dump(buf, t)
  buf.seek(0)
  keys = set(np.load(buf).files)
  assert keys == {"sub:a", "sub:b", "name", "items:0", "items:1", "items:2", "table:p/q:w"}
  buf.seek(0)
  back = checkpoint.load(buf, Tree)
  assert back.sub.a == 3 and back.items == (1, 2, 3) and back.maybe is None and back.name == "x"
  np.testing.assert_array_equal(back.sub.b, t.sub.b)
  np.testing.assert_array_equal(back.table["p/q"]["w"], np.ones(2))
  with pytest.raises(ValueError, match="separator"):
    checkpoint.flatten({"a:b": 1})
  cp = graphcast.CheckPoint(params={"m": {"w": np.zeros(2, np.float32)}},
                            model_config=graphcast.ModelConfig(1.0, 5, 512, 16, 1, 0.6),
                            task_config=graphcast.TASK_13, description="d", license="l")
  buf = io.BytesIO(); checkpoint.dump(buf, cp); buf.seek(0)
  cp2 = checkpoint.load(buf, graphcast.CheckPoint)
  assert cp2.model_config == cp.model_config and cp2.task_config == cp.task_config


def test_checkpoint_written_by_the_reference_loads_and_round_trips():
  """tests/golden/reference_checkpoint.npz was written by the REFERENCE's
  `weathernext.utils.checkpoint.dump` (tests/golden/make_golden.py).  This repo's loader must
  read it into this repo's CheckPoint / ModelConfig / TaskConfig, and this repo's `dump` must
  produce the same flat keys and values."""
  import os
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_checkpoint.npz")
  with open(path, "rb") as f:
    ck = checkpoint.load(f, graphcast.CheckPoint)
  assert ck.model_config == graphcast.ModelConfig(1.0, 5, 512, 16, 1, 0.6, None)
  assert ck.task_config == graphcast.TaskConfig(
      input_variables=("2m_temperature", "geopotential"), target_variables=("2m_temperature",),
      forcing_variables=("toa_incident_solar_radiation",), pressure_levels=(50, 500, 1000),
      input_duration="12h")
  assert ck.license == "n/a" and ck.description.startswith("golden checkpoint")
  lin = ck.params["grid2mesh_gnn/~_networks_builder/encoder_nodes_grid_nodes_mlp/~/linear_0"]
  assert lin["w"].shape == (7, 4) and lin["w"].dtype == np.float32 and not lin["b"].any()
  ref = np.load(path)
  buf = io.BytesIO()
  checkpoint.dump(buf, ck)
  buf.seek(0)
  mine = np.load(buf)
  assert set(mine.files) == set(ref.files)
  for k in ref.files:
    np.testing.assert_array_equal(mine[k], ref[k])


def test_statistics_are_aligned_by_pressure_level_not_by_position():
  """The released `*_by_level` statistics have 37 levels and serve the 13-level models: the
  reference gets the right rows through xarray's label alignment (normalization.py:29-70)."""
  task13 = graphcast.TASK_13_PRECIP_OUT
  inputs, template, forcings = synthetic.make_example(task13, 30.0, seed=1)
  assert inputs["temperature"].index_labels("level") is not None
  std37, mean37, dstd37 = (_stats(graphcast.TASK, s) for s in (1, 2, 3))
  lv37 = list(graphcast.TASK.pressure_levels)
  pick = np.asarray([lv37.index(l) for l in task13.pressure_levels])
  sel = lambda ds: xs.Dataset(
      {k: (v.isel(level=pick) if "level" in v.dims else v) for k, v in ds.data_vars.items()},
      coords={"level": np.asarray(task13.pressure_levels)})
  std13, mean13, dstd13 = sel(std37), sel(mean37), sel(dstd37)
  # generic path
  a, b = Echo(), Echo()
  out37 = normalization.InputsAndResiduals(a, std37, mean37, dstd37)(inputs, template, forcings)
  out13 = normalization.InputsAndResiduals(b, std13, mean13, dstd13)(inputs, template, forcings)
  for name in ("temperature", "geopotential"):
    assert a.inputs.data_vars[name].shape == inputs.data_vars[name].shape
    np.testing.assert_array_equal(a.inputs.data_vars[name].values, b.inputs.data_vars[name].values)
    np.testing.assert_array_equal(out37.data_vars[name].values, out13.data_vars[name].values)
  # fused per-channel constants
  c37 = normalization.InputsAndResiduals(Echo(), std37, mean37, dstd37)._fused_constants(
      inputs, template, forcings, torch.device("cpu"))
  c13 = normalization.InputsAndResiduals(Echo(), std13, mean13, dstd13)._fused_constants(
      inputs, template, forcings, torch.device("cpu"))
  for f in ("in_mean", "in_scale", "out_scale", "out_offset"):
    np.testing.assert_array_equal(getattr(c37, f).numpy(), getattr(c13, f).numpy())
  # a level the statistics do not have is an error, not a silent misalignment
  bad = xs.Dataset({k: (v.isel(level=np.arange(12)) if "level" in v.dims else v)
                    for k, v in std13.data_vars.items()},
                   coords={"level": np.asarray(task13.pressure_levels[:12])})
  with pytest.raises(ValueError, match="lack level"):
    normalization.InputsAndResiduals(Echo(), bad, mean13, dstd13)._fused_constants(
        inputs, template, forcings, torch.device("cpu"))


def test_fused_path_rejects_targets_without_a_single_time_step():
  task = graphcast.TASK_13_PRECIP_OUT
  inputs, template, forcings = synthetic.make_example(task, 30.0, seed=1)
  std, mean, dstd = _stats(task, 1), _stats(task, 2), _stats(task, 3)
  no_time = xs.Dataset({k: v.isel(time=0) for k, v in template.data_vars.items()})
  with pytest.raises(ValueError, match="single timestep"):
    normalization.InputsAndResiduals(Echo(), std, mean, dstd)._fused_constants(
        inputs, no_time, forcings, torch.device("cpu"))
