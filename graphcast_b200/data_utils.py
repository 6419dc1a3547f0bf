"""Preparing model inputs from an example batch: mirror of the reference's
`weathernext/utils/data_utils.py` entry points `extract_input_target_times` (:214-277) and
`extract_inputs_targets_forcings` (:296-333), plus re-exports of the forcing generators
(`graphcast_b200/forcings.py`) under their reference names.

An example batch is a Dataset with dims (batch, time, [level,] lat, lon), a `time` coordinate of
timedeltas with a fixed step and a `datetime` coordinate of dims (batch, time).  No pandas:
durations are `numpy.timedelta64`, python `timedelta`, or strings in the pandas shorthand the
reference's callers use ("6h", "12h", "1 day", "5d12h", "24 hours").

Pinned against the reference functions, executed on stand-in datasets
(tests/golden/reference_data_utils.npz, tests/test_reference_forcings_golden.py)."""

from __future__ import annotations

import datetime as _dt
import re
from typing import Any, Sequence, Tuple, Union

import numpy as np

from graphcast_b200 import forcings
from graphcast_b200 import xarray_shim as xs
from graphcast_b200.forcings import (  # noqa: F401  (reference names)
    AVG_SEC_PER_YEAR, DAY_PROGRESS, SEC_PER_DAY, TISR, YEAR_PROGRESS, add_derived_vars,
    add_tisr_var, featurize_progress, get_day_progress, get_seconds_since_epoch,
    get_year_progress)

_DERIVED_VARS = {DAY_PROGRESS, f"{DAY_PROGRESS}_sin", f"{DAY_PROGRESS}_cos",
                 YEAR_PROGRESS, f"{YEAR_PROGRESS}_sin", f"{YEAR_PROGRESS}_cos"}

_UNIT_NS = {
    "w": 7 * 86400 * 10**9, "d": 86400 * 10**9, "day": 86400 * 10**9, "days": 86400 * 10**9,
    "h": 3600 * 10**9, "hr": 3600 * 10**9, "hour": 3600 * 10**9, "hours": 3600 * 10**9,
    "m": 60 * 10**9, "min": 60 * 10**9, "minute": 60 * 10**9, "minutes": 60 * 10**9,
    "t": 60 * 10**9, "s": 10**9, "sec": 10**9, "second": 10**9, "seconds": 10**9,
    "ms": 10**6, "us": 10**3, "ns": 1,
}


def to_timedelta(value: Any) -> np.timedelta64:
  """`pandas.Timedelta(value)` for the forms used with GraphCast, as a timedelta64[ns]."""
  if isinstance(value, np.timedelta64):
    return value.astype("timedelta64[ns]")
  if isinstance(value, _dt.timedelta):
    return np.timedelta64(value).astype("timedelta64[ns]")
  if isinstance(value, (int, np.integer)):
    return np.timedelta64(int(value), "ns")
  if isinstance(value, str):
    tokens = re.findall(r"([+-]?\d+(?:\.\d+)?)\s*([a-zA-Z]+)", value)
    if not tokens or "".join(a + b for a, b in tokens) != re.sub(r"\s+", "", value):
      raise ValueError(f"cannot parse timedelta {value!r}")
    total = 0
    for number, unit in tokens:
      unit = unit.lower()
      if unit not in _UNIT_NS:
        raise ValueError(f"unknown timedelta unit {unit!r} in {value!r}")
      total += int(round(float(number) * _UNIT_NS[unit]))
    return np.timedelta64(total, "ns")
  raise TypeError(f"cannot convert {type(value).__name__} to a timedelta")


TargetLeadTimes = Union[Any, Sequence[Any], slice]


def _process_target_lead_times_and_get_duration(target_lead_times):
  """(selection, duration of the last lead time); a slice has inclusive ends (:280-293)."""
  if isinstance(target_lead_times, slice):
    start = (np.timedelta64(1, "ns") if target_lead_times.start is None
             else to_timedelta(target_lead_times.start))
    stop = to_timedelta(target_lead_times.stop)
    step = None if target_lead_times.step is None else to_timedelta(target_lead_times.step)
    return slice(start, stop, step), stop
  if not isinstance(target_lead_times, (list, tuple, set)):
    target_lead_times = [target_lead_times]
  times = sorted(to_timedelta(x) for x in target_lead_times)
  return times, times[-1]


def _time_indices(time: np.ndarray, selection) -> np.ndarray:
  """Positions along `time` (timedelta64) of a label selection, like `Dataset.sel(time=...)`."""
  time = time.astype("timedelta64[ns]")
  if isinstance(selection, slice):
    idx = np.flatnonzero((time >= selection.start) & (time <= selection.stop))
    if selection.step is not None:
      if idx.size > 1:
        native = time[idx[1]] - time[idx[0]]
        stride, rem = divmod(selection.step.astype(np.int64), native.astype(np.int64))
        if rem != 0 or stride < 1:
          raise ValueError("slice step must be a multiple of the time resolution")
        idx = idx[::int(stride)]
    return idx
  out = []
  for t in selection:
    hit = np.flatnonzero(time == t)
    if hit.size != 1:
      raise KeyError(f"lead time {t} not found in the time coordinate")
    out.append(int(hit[0]))
  return np.asarray(out, dtype=np.int64)


def extract_input_target_times(dataset: xs.Dataset, input_duration, target_lead_times
                               ) -> Tuple[xs.Dataset, xs.Dataset]:
  """Splits along time into inputs (the `input_duration` ending at lead time 0) and targets
  (the requested lead times); the time coordinate becomes the lead time relative to the last
  input frame (:214-277)."""
  dataset = xs.from_xarray(dataset)
  selection, target_duration = _process_target_lead_times_and_get_duration(target_lead_times)
  time = np.asarray(dataset.coords["time"][1]).astype("timedelta64[ns]")
  dataset = dataset.assign_coords(time=time + target_duration - time[-1])
  time = np.asarray(dataset.coords["time"][1])
  targets = dataset.isel(time=_time_indices(time, selection))
  duration = to_timedelta(input_duration)
  zero = np.timedelta64(0, "ns")
  inputs = dataset.isel(time=_time_indices(time, slice(-duration + np.timedelta64(1, "ns"), zero, None)))
  return inputs, targets


def extract_inputs_targets_forcings(
    dataset: xs.Dataset, *, input_variables: Tuple[str, ...], target_variables: Tuple[str, ...],
    forcing_variables: Tuple[str, ...], pressure_levels: Tuple[int, ...], input_duration,
    target_lead_times) -> Tuple[xs.Dataset, xs.Dataset, xs.Dataset]:
  """(inputs, targets, forcings) of a task from an example batch (:296-333): select the task's
  pressure levels, generate derived / solar forcings if the task asks for them and they are
  missing, split in time, pick the variables."""
  dataset = xs.from_xarray(dataset).copy()
  if "level" in dataset.coords:
    levels = np.asarray(dataset.coords["level"][1])
    idx = []
    for lv in pressure_levels:
      hit = np.flatnonzero(levels == lv)
      if hit.size != 1:
        raise KeyError(f"pressure level {lv} not in the dataset")
      idx.append(int(hit[0]))
    dataset = dataset.isel(level=np.asarray(idx, dtype=np.int64))
  if set(forcing_variables) & _DERIVED_VARS:
    forcings.add_derived_vars(dataset)
  if set(forcing_variables) & {TISR}:
    forcings.add_tisr_var(dataset)
  dataset.coords.pop("datetime", None)
  inputs, targets = extract_input_target_times(dataset, input_duration=input_duration,
                                               target_lead_times=target_lead_times)
  if set(forcing_variables) & set(target_variables):
    raise ValueError(f"Forcing variables {forcing_variables} should not "
                     f"overlap with target variables {target_variables}.")
  return (inputs[list(input_variables)], targets[list(target_variables)],
          targets[list(forcing_variables)])
