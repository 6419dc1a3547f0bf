"""Autoregressive rollout driver -- mirror of `weathernext/utils/rollout.py`.

Public surface kept: `chunked_prediction` (:326-364),
`chunked_prediction_generator` (:367-565), `extend_targets_template` (:618-686),
the `PredictorFn(rng, inputs, targets_template, forcings, **kw)` protocol
(:78-87) and the error behaviour (uneven chunking :439-442, unevenly spaced
target times :444-445, inputs with a time axis that are neither predicted nor
forced :586-591).

Differences: predictions stay device-resident (torch CUDA tensors inside the
Dataset) between steps, so building the next inputs (`_get_next_inputs`
:581-604) is a device-side concatenate of the last `n_input` frames and never
round-trips through the host; `chunked_prediction` brings every chunk to host
memory exactly like the reference's `jax.device_get` (:362).  There is no JAX
PRNG: `rng` is passed through unchanged (GraphCast is deterministic).  The pmap /
replica arguments are accepted; ensemble members over several GPUs are driven by
one process per GPU (see bench.py / DESIGN.md) rather than `pmap`.
"""

from __future__ import annotations

import logging
from typing import Any, Callable, Iterator, Optional, Sequence

import numpy as np

from graphcast_b200 import xarray_shim as xs

PredictorFn = Callable[..., xs.Dataset]


def _to_host(ds: xs.Dataset) -> xs.Dataset:
  out = xs.Dataset(coords=ds.coords)
  for k, v in ds.data_vars.items():
    out[k] = xs.DataArray(np.asarray(v.values), v.dims)
  return out


class PinnedPredictionSink:
  """Device -> pinned-host copies of the chunks a rollout yields, overlapped with the next step.

  The reference's generator hands every chunk to the host before the next chunk starts
  (rollout.py:255-262 `jax.device_get`-style materialisation inside `chunked_prediction`).  Here the
  chunks stay on the device; a consumer that wants them on the host calls `sink(chunk)` per chunk:
  the copies (0.94 GB per 0.25 degree step) run on a side stream behind an event of the compute
  stream, so the next step's kernels -- launched by the following `next()` on the generator -- overlap
  them.  `depth` pinned buffer sets rotate; `wait()` drains.  Returns the pinned tensors of the chunk."""

  def __init__(self, depth: int = 2):
    import torch
    self._torch = torch
    self._bufs = [None] * max(1, int(depth))
    self._done = [None] * len(self._bufs)
    self._n = 0
    self._stream = None

  def __call__(self, chunk: xs.Dataset):
    torch = self._torch
    slot = self._n % len(self._bufs)
    self._n += 1
    tensors = {k: v.data for k, v in chunk.data_vars.items()}
    if any(not (isinstance(t, torch.Tensor) and t.is_cuda) for t in tensors.values()):
      raise TypeError("PinnedPredictionSink expects device-resident chunks")
    dev = next(iter(tensors.values())).device
    if self._stream is None:
      self._stream = torch.cuda.Stream(device=dev)
    if self._done[slot] is not None:
      self._done[slot].synchronize()               # the consumer had `depth` chunks to use the slot
    if self._bufs[slot] is None or any(self._bufs[slot][k].shape != t.shape for k, t in tensors.items()):
      self._bufs[slot] = {k: torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                          for k, t in tensors.items()}
    ready = torch.cuda.Event()
    ready.record(torch.cuda.current_stream(dev))
    with torch.cuda.stream(self._stream):
      self._stream.wait_event(ready)
      for k, t in tensors.items():
        self._bufs[slot][k].copy_(t, non_blocking=True)
        t.record_stream(self._stream)
      done = torch.cuda.Event()
      done.record(self._stream)
    self._done[slot] = done
    return self._bufs[slot]

  def wait(self) -> None:
    for ev in self._done:
      if ev is not None:
        ev.synchronize()


def chunked_prediction(predictor_fn: PredictorFn, rng: Any, inputs, targets_template,
                       forcings=None, num_steps_per_chunk: int = 1, **kwargs) -> xs.Dataset:
  """Full trajectory: concatenation in time of all predicted chunks (host arrays)."""
  chunks = []
  for chunk in chunked_prediction_generator(
      predictor_fn=predictor_fn, rng=rng, inputs=inputs,
      targets_template=targets_template, forcings=forcings,
      num_steps_per_chunk=num_steps_per_chunk, **kwargs):
    chunks.append(_to_host(chunk))
    del chunk
  return xs.concat_time(chunks)


def _slice_sample_if_present(inputs: xs.Dataset, forcings, sample_idx):
  """Member `sample_idx` of inputs / forcings that carry a "sample" dim (reference :310-322)."""
  if "sample" in inputs.dims:
    inputs = inputs.isel(sample=sample_idx, drop=True)
  if forcings is not None and "sample" in forcings.dims:
    forcings = forcings.isel(sample=sample_idx, drop=True)
  return inputs, forcings


def chunked_prediction_generator_multiple_runs(
    predictor_fn: PredictorFn, rngs, inputs, targets_template, forcings,
    num_samples: Optional[int], pmap_devices: Optional[Sequence[Any]] = None,
    rank: Optional[int] = None, world_size: Optional[int] = None,
    **chunked_prediction_kwargs) -> Iterator[xs.Dataset]:
  """Ensemble rollouts: all lead-time chunks of one member, then the next member
  (reference `chunked_prediction_generator_multiple_runs`, rollout.py:158-306).

  The reference spreads members over `pmap_devices` inside one process.  Here one process
  drives one GPU, so the equivalent is to give every rank its own block of members: pass
  `rank` / `world_size` (default: the initialised `torch.distributed` group, else a single
  rank) and this generator yields only the members of `parallel.members_for_rank`; there is
  no data-path collective.  Every yielded chunk carries `coords["sample"]` = member index, as
  in the reference's non-pmap branch.  `pmap_devices` must stay None."""
  from graphcast_b200 import parallel
  if pmap_devices is not None:
    raise ValueError("pmap_devices is not supported: run one process per GPU and pass rank / "
                     "world_size (or initialise torch.distributed)")
  inputs = xs.from_xarray(inputs)
  forcings = xs.from_xarray(forcings) if forcings is not None else None
  if num_samples is None:
    if "sample" not in inputs.dims:
      raise ValueError("The number of samples must be passed when `inputs` don't have a "
                       "`sample` dim.")
    num_samples = inputs.sizes["sample"]
  if "sample" in inputs.dims and num_samples != inputs.sizes["sample"]:
    raise ValueError("Inconsistent number of samples requested for inputs"
                     f"{num_samples} != {inputs.sizes['sample']}.")
  if num_samples != len(rngs):
    raise ValueError(f"Inconsistent number of rngs passed. {num_samples} != {len(rngs)}.")
  if forcings is not None and "sample" in forcings.dims and num_samples != forcings.sizes["sample"]:
    raise ValueError("Inconsistent number of samples requested for forcings"
                     f"{num_samples} != {forcings.sizes['sample']}.")
  if rank is None or world_size is None:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
      rank, world_size = dist.get_rank(), dist.get_world_size()
    else:
      rank, world_size = 0, 1
  for i in parallel.members_for_rank(num_samples, rank, world_size):
    logging.info("Sample %d/%d", i, num_samples)
    sample_inputs, sample_forcings = _slice_sample_if_present(inputs, forcings, i)
    for chunk in chunked_prediction_generator(
        predictor_fn, rngs[i], inputs=sample_inputs, targets_template=targets_template,
        forcings=sample_forcings, **chunked_prediction_kwargs):
      chunk.coords["sample"] = ((), np.asarray(i))
      yield chunk


def chunked_prediction_generator(
    predictor_fn: PredictorFn, rng: Any, inputs, targets_template,
    num_steps_per_chunk: int, forcings=None, verbose: bool = False,
    pmap_devices: Optional[Sequence[Any]] = None, replica_axis: Optional[str] = None,
    device_put_fn: Optional[Callable] = None, replicate_fn: Optional[Callable] = None,
    generate_forcings: Optional[Sequence[str]] = None,
) -> Iterator[xs.Dataset]:
  """Yields the predictions of each chunk (device-resident).

  `generate_forcings` (extension; the reference expects every forcing precomputed in `forcings`,
  data_utils.py:51-215): names of forcing variables to generate on the device for every chunk from
  the targets' `datetime` / `lat` / `lon` coordinates -- `toa_incident_solar_radiation` by the CUDA
  kernel, the year / day progress features on the host (KBs) -- so that a long rollout uploads no
  forcing fields at all."""
  if pmap_devices is not None and replica_axis is None:
    raise ValueError("Must provide replica_axis when pmap_devices is provided.")
  if (replicate_fn is None) ^ (replica_axis is None):
    raise ValueError("Must provide replicate_fn when replica_axis is provided.")

  inputs = xs.from_xarray(inputs)
  targets_template = xs.from_xarray(targets_template)
  if forcings is None:
    forcings = xs.Dataset(coords={n: c for n, c in targets_template.coords.items()
                                  if "time" in c[0]})
  forcings = xs.from_xarray(forcings)

  inputs, targets_template, forcings = inputs.copy(), targets_template.copy(), forcings.copy()
  if generate_forcings:
    if "datetime" not in targets_template.coords:
      raise ValueError("generate_forcings needs the `datetime` coordinate of the targets template")
    gen_dims, gen_datetime = targets_template.coords["datetime"]
    gen_lat = np.asarray(targets_template.coords["lat"][1])
    gen_lon = np.asarray(targets_template.coords["lon"][1])
  inputs.coords.pop("datetime", None)
  output_datetime = targets_template.coords.pop("datetime", None)
  forcings.coords.pop("datetime", None)

  num_target_steps = targets_template.sizes["time"]
  num_chunks, remainder = divmod(num_target_steps, num_steps_per_chunk)
  if remainder != 0:
    raise ValueError(
        f"The number of steps per chunk {num_steps_per_chunk} must "
        f"evenly divide the number of target steps {num_target_steps} ")
  target_times = np.asarray(targets_template.coords["time"][1])
  if len(np.unique(np.diff(target_times))) > 1:
    raise ValueError("The targets time coordinates must be evenly spaced")

  # Time coordinates relative to the chunk: every call sees those of the first chunk.
  chunk_inputs_time = np.asarray(inputs.coords["time"][1])
  chunk_targets_time = target_times[:num_steps_per_chunk]

  current_inputs = inputs
  if replicate_fn is not None:
    current_inputs = replicate_fn(current_inputs)
  if device_put_fn is not None:
    current_inputs = device_put_fn(current_inputs)
  del inputs

  for chunk_index in range(num_chunks):
    if verbose:
      logging.info("Chunk %d/%d", chunk_index, num_chunks)
    lo = num_steps_per_chunk * chunk_index
    target_slice = slice(lo, lo + num_steps_per_chunk)
    current_targets_template = targets_template.isel(time=target_slice)
    current_forcings = forcings.isel(time=target_slice)
    if generate_forcings:
      from graphcast_b200 import forcings as forcings_lib
      idx = tuple(target_slice if d == "time" else slice(None) for d in gen_dims)
      generated = forcings_lib.device_forcings(generate_forcings, np.asarray(gen_datetime)[idx], gen_dims,
                                               gen_lat, gen_lon)
      current_forcings = current_forcings.assign(generated)
    time_coords_to_override = {n: c for n, c in current_targets_template.coords.items()
                               if "time" in c[0]}
    if replicate_fn is not None:
      current_forcings = replicate_fn(current_forcings)
      current_targets_template = replicate_fn(current_targets_template)
    if device_put_fn is not None:
      current_forcings = device_put_fn(current_forcings)
      current_targets_template = device_put_fn(current_targets_template)

    current_inputs = current_inputs.assign_coords(time=chunk_inputs_time)
    current_forcings = current_forcings.assign_coords(time=chunk_targets_time)
    current_targets_template = current_targets_template.assign_coords(time=chunk_targets_time)
    predictions = predictor_fn(rng=rng, inputs=current_inputs,
                               targets_template=current_targets_template,
                               forcings=current_forcings)
    del current_targets_template

    if chunk_index == num_chunks - 1:
      current_inputs = None
    else:
      next_frame = predictions.assign(current_forcings)
      current_inputs = _get_next_inputs(current_inputs, next_frame)
      del next_frame
    del current_forcings

    predictions = predictions.assign_coords(time_coords_to_override)
    if output_datetime is not None:
      dims, vals = output_datetime
      idx = tuple(target_slice if d == "time" else slice(None) for d in dims)
      predictions.coords["datetime"] = (dims, np.asarray(vals)[idx])
    yield predictions


def _get_next_inputs(prev_inputs: xs.Dataset, next_frame: xs.Dataset) -> xs.Dataset:
  """Next inputs = last `n_input` frames of concat(prev inputs, new frame) for the
  input variables that are predicted or forced (reference :581-604)."""
  missing = set(prev_inputs.keys()) - set(next_frame.keys())
  for name in missing:
    if "time" in prev_inputs.data_vars[name].dims:
      raise ValueError("Found an input with a time index that is not predicted or forced.")
  next_keys = [k for k in prev_inputs.keys() if k in next_frame]
  num_inputs = prev_inputs.sizes["time"]
  merged = xs.concat_time([prev_inputs, _aligned(next_frame[next_keys], prev_inputs)])
  return merged.isel(time=slice(-num_inputs, None))


def _aligned(frame: xs.Dataset, like: xs.Dataset) -> xs.Dataset:
  """Variables of `frame` with the dim order of the same variables in `like`, plus
  the time-less variables of `like` (static inputs) so the concat keeps them."""
  out = xs.Dataset(coords={k: c for k, c in frame.coords.items()})
  for name, v in like.data_vars.items():
    if name in frame:
      out[name] = frame.data_vars[name].transpose(*v.dims)
    elif "time" not in v.dims:
      out[name] = v
  return out


def extend_targets_template(targets_template, required_num_steps: int,
                            value: Optional[float] = None) -> xs.Dataset:
  """Template with `required_num_steps` evenly spaced target times.  Values are
  placeholders (broadcast views, no memory), as only names/dims/coords are used."""
  targets_template = xs.from_xarray(targets_template)
  time = np.asarray(targets_template.coords["time"][1])
  timestep = time[0]
  if time.shape[0] > 1:
    assert np.all(timestep == time[1:] - time[:-1])
  extended_time = (np.arange(required_num_steps) + 1) * timestep
  coords = {k: c for k, c in targets_template.coords.items() if "time" not in c[0]}
  coords["time"] = (("time",), extended_time)
  if "datetime" in targets_template.coords:
    dims, vals = targets_template.coords["datetime"]
    vals = np.asarray(vals)
    if tuple(dims) == ("time",):
      coords["datetime"] = (("time",), (vals[0] - timestep) + extended_time)
    else:                                   # ([batch,] time): every batch element keeps its own start
      t_axis = tuple(dims).index("time")
      first = np.take(vals, [0], axis=t_axis)
      shape = [1] * vals.ndim
      shape[t_axis] = required_num_steps
      coords["datetime"] = (tuple(dims), (first - timestep) + extended_time.reshape(shape))
  out = xs.Dataset(coords=coords)
  fill = 0.0 if value is None else value
  for name, v in targets_template.data_vars.items():
    shape = list(v.shape)
    axis = v.dims.index("time")
    shape[axis] = required_num_steps
    dtype = v.data.dtype if isinstance(v.data, np.ndarray) else np.float32
    data = np.broadcast_to(np.asarray(fill, dtype=dtype), shape)
    out[name] = xs.DataArray(data, v.dims)
  return out
