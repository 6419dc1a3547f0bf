"""Synthetic atmospheric states with the shapes of a GraphCast TaskConfig
(there is no network for ERA5 / checkpoints): N(0,1) model-space values, seeded.
Layout follows the reference's example batches (notebook cells 13-17): inputs
(batch, time=2, [level], lat, lon), static variables (lat, lon), time forcings
(batch, time) / (batch, time, lon), targets / forcings at one target time."""

from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from graphcast_b200 import graphcast
from graphcast_b200 import variables as V
from graphcast_b200 import xarray_shim as xs


def grid_coords(resolution: float) -> Tuple[np.ndarray, np.ndarray]:
  n_lat = int(round(180.0 / resolution)) + 1
  lat = np.linspace(-90.0, 90.0, n_lat).astype(np.float32)
  lon = (np.arange(int(round(360.0 / resolution))) * resolution).astype(np.float32)
  return lat, lon


def _var_dims(name: str, with_time: bool) -> Tuple[str, ...]:
  if name in V.STATIC_VARS:
    return ("lat", "lon")
  if name.startswith("year_progress"):
    return ("batch", "time")
  if name.startswith("day_progress"):
    return ("batch", "time", "lon")
  if name in V.ALL_ATMOSPHERIC_VARS:
    return ("batch", "time", "level", "lat", "lon")
  return ("batch", "time", "lat", "lon")


def make_example(task: graphcast.TaskConfig, resolution: float, *, batch: int = 1,
                 num_target_steps: int = 1, seed: int = 0, pinned: bool = False,
                 device: Optional[str] = None):
  """Returns (inputs, targets_template, forcings) Datasets.

  pinned=True allocates host arrays in page-locked memory (so the per-step
  host->device copies of the e2e path are real async DMA); device="cuda" creates
  the arrays on the GPU instead."""
  lat, lon = grid_coords(resolution)
  levels = np.asarray(task.pressure_levels, np.int32)
  sizes = {"batch": batch, "level": len(levels), "lat": len(lat), "lon": len(lon)}
  gen = torch.Generator(device="cpu")
  gen.manual_seed(seed)

  def rand(shape):
    if device is not None and device != "cpu":
      g = torch.Generator(device=device)
      g.manual_seed(int(torch.randint(0, 2**31 - 1, (1,), generator=gen)))
      return torch.randn(shape, generator=g, device=device, dtype=torch.float32)
    t = torch.empty(shape, dtype=torch.float32, pin_memory=pinned)
    t.normal_(generator=gen)
    return t.numpy()

  def dataset(names, n_time, time_values):
    coords = {"lat": lat, "lon": lon, "level": levels, "time": time_values}
    ds = xs.Dataset(coords=coords)
    for name in names:
      dims = _var_dims(name, True)
      shape = [n_time if d == "time" else sizes[d] for d in dims]
      ds[name] = xs.DataArray(rand(shape), dims)
    return ds

  six_h = np.timedelta64(6, "h")
  inputs = dataset(task.input_variables, 2, np.array([-six_h, 0 * six_h]))
  target_times = (np.arange(num_target_steps) + 1) * six_h
  forcings = dataset(task.forcing_variables, num_target_steps, target_times)
  # Template: only names / dims / coords matter -> zero-memory broadcast placeholders.
  template = xs.Dataset(coords={"lat": lat, "lon": lon, "level": levels, "time": target_times})
  for name in task.target_variables:
    dims = _var_dims(name, True)
    shape = [num_target_steps if d == "time" else sizes[d] for d in dims]
    template[name] = xs.DataArray(np.broadcast_to(np.float32(np.nan), shape), dims)
  return inputs, template, forcings


def num_input_channels(task: graphcast.TaskConfig, n_input_frames: int = 2) -> int:
  """C_in of SURVEY.md section 8 (without the 3 structural features)."""
  n = 0
  L = len(task.pressure_levels)
  for name in task.input_variables:
    if name in V.STATIC_VARS:
      n += 1
    elif name in V.ALL_ATMOSPHERIC_VARS:
      n += n_input_frames * L
    else:
      n += n_input_frames
  n += len(task.forcing_variables)          # target-time forcings
  return n
