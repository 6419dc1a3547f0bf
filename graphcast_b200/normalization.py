"""Input normalisation / residual un-normalisation wrapper -- mirror of
`weathernext/utils/normalization.py` (`normalize` :29-48, `unnormalize` :51-70,
`InputsAndResiduals` :73-160).

Semantics kept:
  model input      = (x - mean) / std                 for every input / forcing
  prediction       = y * diffs_std + last input frame for targets present in inputs
  prediction       = y * std + mean                   for targets not in inputs
  only single-step targets are supported (ValueError otherwise, :114-117).

When the wrapped predictor is the B200 `GraphCast`, the affine maps are not
applied to the Datasets at all: they are folded into per-channel vectors and
executed inside the pack / unpack kernels (`gcb_pack_grid_features`,
`gcb_unpack_grid_outputs`), which removes three full passes over the 0.25 degree
state per step.  Any other predictor gets the generic Dataset arithmetic.
"""

from __future__ import annotations

import logging
from typing import Optional

import numpy as np
import torch

from graphcast_b200 import graphcast
from graphcast_b200 import model_utils
from graphcast_b200 import xarray_shim as xs


def normalize(values: xs.Dataset, scales: xs.Dataset, locations: Optional[xs.Dataset]
              ) -> xs.Dataset:
  def one(array: xs.DataArray) -> xs.DataArray:
    if array.name is None:
      raise ValueError("Can't look up normalization constants because array has no name.")
    if locations is not None:
      if array.name in locations:
        array = array - locations[array.name].astype(array.dtype)
      else:
        logging.warning("No normalization location found for %s", array.name)
    if array.name in scales:
      array = array / scales[array.name].astype(array.dtype)
    else:
      logging.warning("No normalization scale found for %s", array.name)
    return array
  return values.map(one)


def unnormalize(values: xs.Dataset, scales: xs.Dataset, locations: Optional[xs.Dataset]
                ) -> xs.Dataset:
  def one(array: xs.DataArray) -> xs.DataArray:
    if array.name is None:
      raise ValueError("Can't look up normalization constants because array has no name.")
    if array.name in scales:
      array = array * scales[array.name].astype(array.dtype)
    else:
      logging.warning("No normalization scale found for %s", array.name)
    if locations is not None:
      if array.name in locations:
        array = array + locations[array.name].astype(array.dtype)
      else:
        logging.warning("No normalization location found for %s", array.name)
    return array
  return values.map(one)


def _per_channel(slab: model_utils.ChannelSlab, stats: Optional[xs.Dataset], default: float
                 ) -> np.ndarray:
  """Per-channel constant of one variable: stats[var] may be a scalar or vary
  along stacked dims (e.g. "level"); it is broadcast over the slab's stack dims."""
  out = np.full(slab.stack_sizes if slab.stack_sizes else (1,), default, np.float64)
  if stats is not None and slab.name in stats:
    v = stats.data_vars[slab.name]
    bad = [d for d in v.dims if d not in slab.stack_dims]
    if bad:
      raise ValueError(f"normalisation statistics of {slab.name!r} vary along {bad}, "
                       "which is not a channel dimension")
    arr = np.asarray(v.values, np.float64)
    shape = [v.sizes[d] if d in v.dims else 1 for d in slab.stack_dims] or [1]
    arr = np.transpose(arr, [v.dims.index(d) for d in slab.stack_dims if d in v.dims]) \
        if v.dims else arr
    out = out * 0 + arr.reshape(shape)
  return out.reshape(-1)


class InputsAndResiduals(graphcast.Predictor):
  """Normalises inputs and predicts normalised residuals (reference :73-160)."""

  def __init__(self, predictor: graphcast.Predictor, stddev_by_level: xs.Dataset,
               mean_by_level: xs.Dataset, diffs_stddev_by_level: xs.Dataset):
    self._predictor = predictor
    self._scales = xs.from_xarray(stddev_by_level)
    self._locations = xs.from_xarray(mean_by_level)
    self._residual_scales = xs.from_xarray(diffs_stddev_by_level)
    self._residual_locations = None
    self._fused_cache = None

  def _unnormalize_prediction_and_add_input(self, inputs, norm_prediction):
    if norm_prediction.sizes.get("time") != 1:
      raise ValueError("normalization.InputsAndResiduals only supports predicting a "
                       "single timestep.")
    if norm_prediction.name in inputs:
      prediction = unnormalize(xs.Dataset({norm_prediction.name: norm_prediction}),
                               self._residual_scales, self._residual_locations
                               )[norm_prediction.name]
      last_input = inputs[norm_prediction.name].isel(time=slice(-1, None))
      return (prediction.transpose(*last_input.dims) + last_input).transpose(
          *norm_prediction.dims)
    return unnormalize(xs.Dataset({norm_prediction.name: norm_prediction}),
                       self._scales, self._locations)[norm_prediction.name]

  # -- fused path --------------------------------------------------------------------
  def _fused_constants(self, inputs, targets_template, forcings, device):
    in_slabs = model_utils.channel_layout(inputs)
    n_in = sum(s.count for s in in_slabs)
    f_slabs = model_utils.channel_layout(forcings, start=n_in)
    key = (tuple((s.name, s.stack_sizes) for s in in_slabs + f_slabs),
           tuple(sorted(targets_template.data_vars.keys())), str(device))
    if self._fused_cache is not None and self._fused_cache[0] == key:
      return self._fused_cache[1]
    mean = np.concatenate([_per_channel(s, self._locations, 0.0) for s in in_slabs + f_slabs])
    scale = np.concatenate([_per_channel(s, self._scales, 1.0) for s in in_slabs + f_slabs])
    t_slabs = model_utils.channel_layout(targets_template)
    out_scale, out_offset, add_idx = [], [], []
    in_by_name = {s.name: s for s in in_slabs}
    for s in t_slabs:
      if s.stack_dims and s.stack_dims[0] == "time" and s.stack_sizes[0] != 1:
        raise ValueError("normalization.InputsAndResiduals only supports predicting a "
                         "single timestep.")
      if s.name in in_by_name:
        src = in_by_name[s.name]
        # channel of the LAST input frame with the same non-time indices
        n_time = src.stack_sizes[src.stack_dims.index("time")] if "time" in src.stack_dims else 1
        per_frame = src.count // n_time
        if "time" in src.stack_dims and src.stack_dims[0] != "time":
          raise ValueError(f"{s.name}: time must be the leading stacked dim of the inputs")
        if per_frame != s.count:
          raise ValueError(f"{s.name}: target has {s.count} channels per frame, input {per_frame}")
        out_scale.append(_per_channel(s, self._residual_scales, 1.0))
        out_offset.append(np.zeros([s.count]))
        add_idx.append(src.start + (n_time - 1) * per_frame + np.arange(s.count))
      else:
        out_scale.append(_per_channel(s, self._scales, 1.0))
        out_offset.append(_per_channel(s, self._locations, 0.0))
        add_idx.append(np.full([s.count], -1))
    t = lambda a, dt: torch.as_tensor(np.concatenate(a) if isinstance(a, list) else a).to(dt).to(device)
    consts = graphcast.FusedNormalization(
        in_mean=t(mean, torch.float32), in_scale=t(scale, torch.float32),
        out_scale=t(out_scale, torch.float32), out_offset=t(out_offset, torch.float32),
        add_plane_index=t(add_idx, torch.int32))
    self._fused_cache = (key, consts)
    return consts

  def __call__(self, inputs, targets_template, forcings, **kwargs):
    inputs, forcings = xs.from_xarray(inputs), xs.from_xarray(forcings)
    targets_template = xs.from_xarray(targets_template)
    if isinstance(self._predictor, graphcast.GraphCast):
      device = self._predictor._device or f"cuda:{torch.cuda.current_device()}"
      consts = self._fused_constants(inputs, targets_template, forcings, torch.device(device))
      return self._predictor._call(inputs, targets_template, forcings, norm=consts)
    norm_inputs = normalize(inputs, self._scales, self._locations)
    norm_forcings = normalize(forcings, self._scales, self._locations)
    norm_predictions = self._predictor(norm_inputs, targets_template, forcings=norm_forcings,
                                       **kwargs)
    return norm_predictions.map(
        lambda pred: self._unnormalize_prediction_and_add_input(inputs, pred))
