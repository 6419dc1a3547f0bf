"""Forcing generation on the host: year / day progress features and top-of-atmosphere incident
solar radiation (TISR) - the inputs GraphCast needs at every target time of a rollout and that
are known in advance.

Mirror of `weathernext/utils/data_utils.py:51-190` (`get_year_progress`, `get_day_progress`,
`featurize_progress`, `add_derived_vars`, `add_tisr_var`) and of
`weathernext/utils/solar_radiation.py` (`get_tsi`, `get_toa_incident_solar_radiation`), numpy
only (no pandas / jax): timestamps are `numpy.datetime64`, all arithmetic is float64 and the
results are cast to float32 where the reference does.  Pinned against the reference's own
functions, executed (tests/golden/reference_forcings.npz, tests/test_reference_forcings_golden.py).

TISR follows the reference's scheme: the instantaneous flux `TSI * (1 / d_au)^2 * max(sin(alt), 0)`
with the orbital parameters of the GEM / IFS formulas it cites (solar_radiation.py:197-292),
integrated over `integration_period` (default 1 h, like ERA5's `tisr`) ending at each timestamp
with the trapezoidal rule on `num_integration_bins` bins.  The integration loop runs bin by bin
(361 passes over a [lat, lon] array) instead of materialising a [lat, lon, 361] tensor, so a
0.25 degree field needs 8 MB instead of 3 GB."""

from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from graphcast_b200 import xarray_shim as xs

SEC_PER_DAY = 24 * 3600
_AVG_DAY_PER_YEAR = 365.24219
AVG_SEC_PER_YEAR = SEC_PER_DAY * _AVG_DAY_PER_YEAR
DAY_PROGRESS = "day_progress"
YEAR_PROGRESS = "year_progress"
TISR = "toa_incident_solar_radiation"

_J2000_UNIX_DAYS = 10957.5              # 2000-01-01T12:00 TT epoch (JD 2451545.0) in days since 1970-01-01
_JULIAN_YEAR_DAYS = 365.25
_REFERENCE_TSI = 1361.0


# ---- progress features (data_utils.py:51-140) -------------------------------------------------
def get_year_progress(seconds_since_epoch: np.ndarray) -> np.ndarray:
  """Fraction of the (average) year elapsed, in [0, 1), float32."""
  years = np.asarray(seconds_since_epoch) / SEC_PER_DAY / np.float64(_AVG_DAY_PER_YEAR)
  return np.mod(years, 1.0).astype(np.float32)


def get_day_progress(seconds_since_epoch: np.ndarray, longitude: np.ndarray) -> np.ndarray:
  """Fraction of the local solar day elapsed at every longitude: [..., lon], float32."""
  greenwich = np.mod(np.asarray(seconds_since_epoch), SEC_PER_DAY) / SEC_PER_DAY
  offsets = np.deg2rad(np.asarray(longitude)) / (2 * np.pi)
  return np.mod(greenwich[..., np.newaxis] + offsets, 1.0).astype(np.float32)


def featurize_progress(name: str, dims: Sequence[str], progress: np.ndarray
                       ) -> Dict[str, Tuple[Tuple[str, ...], np.ndarray]]:
  """`name`, `name_sin`, `name_cos` as (dims, array) pairs."""
  if len(dims) != progress.ndim:
    raise ValueError(f"Number of feature dimensions ({len(dims)}) must be equal to the"
                     f" number of data dimensions: {progress.ndim}.")
  phase = progress * (2 * np.pi)
  dims = tuple(dims)
  return {name: (dims, progress), name + "_sin": (dims, np.sin(phase)),
          name + "_cos": (dims, np.cos(phase))}


def get_seconds_since_epoch(datetimes: np.ndarray) -> np.ndarray:
  return np.asarray(datetimes).astype("datetime64[s]").astype(np.int64)


def add_derived_vars(data: xs.Dataset) -> None:
  """Adds year / day progress (+ sin, cos) to `data` in place if missing; needs the `datetime`
  ([batch,] time) and `lon` coordinates (data_utils.py:143-183)."""
  for coord in ("datetime", "lon"):
    if coord not in data.coords:
      raise ValueError(f"'{coord}' must be in `data` coordinates.")
  dt_dims, dt = data.coords["datetime"]
  seconds = get_seconds_since_epoch(dt)
  batch_dim = ("batch",) if "batch" in data.sizes else ()
  if tuple(dt_dims) != batch_dim + ("time",):
    raise ValueError(f"`datetime` must have dims {batch_dim + ('time',)}, found {tuple(dt_dims)}")
  if YEAR_PROGRESS not in data:
    for k, v in featurize_progress(YEAR_PROGRESS, batch_dim + ("time",),
                                   get_year_progress(seconds)).items():
      data[k] = v
  if DAY_PROGRESS not in data:
    lon_dims, lon = data.coords["lon"]
    for k, v in featurize_progress(DAY_PROGRESS, batch_dim + ("time",) + tuple(lon_dims),
                                   get_day_progress(seconds, np.asarray(lon))).items():
      data[k] = v


# ---- total solar irradiance (solar_radiation.py:66-160) -----------------------------------------
def reference_tsi_data() -> Tuple[np.ndarray, np.ndarray]:
  """(years, W/m^2): one constant value."""
  return np.array([0.0]), np.array([_REFERENCE_TSI])


def era5_tsi_data() -> Tuple[np.ndarray, np.ndarray]:
  """(mid-year coordinates 1951.5 .. 2034.5, W/m^2): the yearly TSI series hard-coded in ECMWF's IFS
  cycle 41r2 as used for ERA5, scaled by 0.9965 (values: solar_radiation.py:83-115)."""
  years = np.arange(1951.5, 2035.5, 1.0)
  head = [  # 1951-1995
      1365.7765, 1365.7676, 1365.6284, 1365.6564, 1365.7773, 1366.3109, 1366.6681, 1366.6328,
      1366.3828, 1366.2767, 1365.9199, 1365.7484, 1365.6963, 1365.6976, 1365.7341, 1365.9178,
      1366.1143, 1366.1644, 1366.2476, 1366.2426, 1365.9580, 1366.0525, 1365.7991, 1365.7271,
      1365.5345, 1365.6453, 1365.8331, 1366.2747, 1366.6348, 1366.6482, 1366.6951, 1366.2859,
      1366.1992, 1365.8103, 1365.6416, 1365.6379, 1365.7899, 1366.0826, 1366.6479, 1366.5533,
      1366.4457, 1366.3021, 1366.0286, 1365.7971, 1365.6996]
  cycle = [  # 1996-2008, repeated for 2009-2021 and 2022-2034
      1365.6121, 1365.7399, 1366.1021, 1366.3851, 1366.6836, 1366.6022, 1366.6807, 1366.2300,
      1366.0480, 1365.8545, 1365.8107, 1365.7240, 1365.6918]
  return years, 0.9965 * np.array(head + 3 * cycle)


def get_tsi(timestamps: Sequence, tsi_data: Tuple[np.ndarray, np.ndarray]) -> np.ndarray:
  """TSI at each timestamp, linearly interpolated in fractional years."""
  ts = np.asarray(timestamps, dtype="datetime64[ns]")
  year_start = ts.astype("datetime64[Y]")
  next_year = year_start + np.timedelta64(1, "Y")
  day = np.timedelta64(1, "D")
  year_length = (next_year.astype("datetime64[D]") - year_start.astype("datetime64[D]")) / day
  elapsed = (ts - year_start.astype("datetime64[ns]")) / day.astype("timedelta64[ns]")
  fractional_year = year_start.astype(np.int64) + 1970 + elapsed / year_length
  years, values = tsi_data
  return np.interp(fractional_year, years, values)


# ---- solar geometry (solar_radiation.py:185-366) -------------------------------------------------
def _j2000_days(timestamps: np.ndarray) -> np.ndarray:
  ns = np.asarray(timestamps, dtype="datetime64[ns]").astype(np.int64)
  return ns / (SEC_PER_DAY * 1e9) - _J2000_UNIX_DAYS


def _orbital_parameters(j2000_days):
  """(rotational phase, sin / cos of the solar declination, equation of time [s], Earth-Sun
  distance [au]) at the given J2000 days (scalar or array)."""
  theta = j2000_days / _JULIAN_YEAR_DAYS
  rotational_phase = np.mod(j2000_days, 1.0)
  rel = 1.7535 + 6.283076 * theta
  rem = 6.240041 + 6.283020 * theta
  rlls = 4.8951 + 6.283076 * theta
  sin_rel, cos_rel = np.sin(rel), np.cos(rel)
  # ecliptic longitude of the Sun, declination
  rllls = (4.8952 + 6.283320 * theta - 0.0075 * sin_rel - 0.0326 * cos_rel
           - 0.0003 * np.sin(2.0 * rel) + 0.0002 * np.cos(2.0 * rel))
  sin_decl = np.sin(0.409093) * np.sin(rllls)
  cos_decl = np.sqrt(1.0 - sin_decl ** 2)
  sin_rem = np.sin(rem)
  eq_of_time_s = (591.8 * np.sin(2.0 * rlls) - 459.4 * sin_rem + 39.5 * sin_rem * np.cos(2.0 * rlls)
                  - 12.7 * np.sin(4.0 * rlls) - 4.8 * np.sin(2.0 * rem))
  distance_au = 1.0001 - 0.0163 * sin_rel + 0.0037 * cos_rel
  return rotational_phase, sin_decl, cos_decl, eq_of_time_s, distance_au


def _radiation_flux(j2000_days, sin_lat, cos_lat, lon, tsi):
  """Instantaneous TOA flux [W/m^2]; all arguments broadcast together."""
  rotational_phase, sin_decl, cos_decl, eq_of_time_s, distance_au = _orbital_parameters(j2000_days)
  solar_time = rotational_phase + eq_of_time_s / SEC_PER_DAY
  hour_angle = 2.0 * np.pi * solar_time + lon
  sin_alt = cos_lat * cos_decl * np.cos(hour_angle) + sin_lat * sin_decl
  return tsi * (1.0 / distance_au) ** 2 * np.maximum(sin_alt, 0.0)


def get_toa_incident_solar_radiation(
    timestamps: Sequence, latitude: np.ndarray, longitude: np.ndarray,
    tsi_data: Optional[Tuple[np.ndarray, np.ndarray]] = None,
    integration_period=np.timedelta64(1, "h"), num_integration_bins: int = 360) -> np.ndarray:
  """[time, lat, lon] float64: TOA incident solar radiation in J/m^2 integrated over
  `integration_period` up to each timestamp (solar_radiation.py:443-521)."""
  ts = np.asarray(timestamps, dtype="datetime64[ns]").reshape(-1)
  lat = np.radians(np.asarray(latitude, dtype=np.float64)).reshape(-1, 1)
  lon = np.radians(np.asarray(longitude, dtype=np.float64)).reshape(1, -1)
  sin_lat, cos_lat = np.sin(lat), np.cos(lat)
  tsi = get_tsi(ts, tsi_data if tsi_data is not None else era5_tsi_data())
  period_days = np.timedelta64(integration_period).astype("timedelta64[ns]").astype(np.int64) / (SEC_PER_DAY * 1e9)
  offsets = np.linspace(-period_days, 0.0, num_integration_bins + 1)
  dx = period_days * SEC_PER_DAY / num_integration_bins
  days = _j2000_days(ts)
  out = np.empty((ts.shape[0], lat.shape[0], lon.shape[1]), np.float64)
  for t in range(ts.shape[0]):
    # Per bin the orbital parameters are scalars: only cos(hour angle + lon) [lon], one outer
    # product, one add and one max touch the [lat, lon] field (same arithmetic, same order as
    # `_radiation_flux`).
    acc = np.zeros((lat.shape[0], lon.shape[1]), np.float64)
    for b, off in enumerate(offsets):
      weight = 0.5 if b in (0, num_integration_bins) else 1.0       # trapezoidal rule
      phase, sin_decl, cos_decl, eot, dist = _orbital_parameters(days[t] + off)
      cos_h = np.cos(2.0 * np.pi * (phase + eot / SEC_PER_DAY) + lon)            # [1, lon]
      sin_alt = (cos_lat * cos_decl) * cos_h + sin_lat * sin_decl              # [lat, lon]
      np.maximum(sin_alt, 0.0, out=sin_alt)
      acc += (weight * (tsi[t] * (1.0 / dist) ** 2)) * sin_alt
    out[t] = acc * dx
  return out


def add_tisr_var(data: xs.Dataset) -> None:
  """Adds `toa_incident_solar_radiation` ([batch,] time, lat, lon; float32) to `data` in place
  if missing; needs the `datetime`, `lat` and `lon` coordinates (data_utils.py:186-215)."""
  if TISR in data:
    return
  for coord in ("datetime", "lat", "lon"):
    if coord not in data.coords:
      raise ValueError(f"'{coord}' must be in `data` coordinates.")
  dt_dims, dt = data.coords["datetime"]
  lat = np.asarray(data.coords["lat"][1])
  lon = np.asarray(data.coords["lon"][1])
  dt = np.asarray(dt)
  flat = get_toa_incident_solar_radiation(dt.reshape(-1), lat, lon)
  tisr = flat.reshape(dt.shape + flat.shape[1:]).astype(np.float32)
  data[TISR] = (tuple(dt_dims) + ("lat", "lon"), tisr)


# ---- device version -------------------------------------------------------------------------------
def _integration_table(timestamps: np.ndarray, tsi_data, integration_period, num_integration_bins
                       ) -> np.ndarray:
  """[T, bins + 1, 5] float32 for `gcb_toa_incident_solar_radiation`: per integration bin the
  cos / sin of the solar declination, the cos / sin of the hour angle at longitude 0 and
  weight * TSI / d^2 * dx (all scalars per timestamp and bin, computed in float64)."""
  ts = np.asarray(timestamps, dtype="datetime64[ns]").reshape(-1)
  tsi = get_tsi(ts, tsi_data if tsi_data is not None else era5_tsi_data())
  period_days = np.timedelta64(integration_period).astype("timedelta64[ns]").astype(np.int64) / (SEC_PER_DAY * 1e9)
  offsets = np.linspace(-period_days, 0.0, num_integration_bins + 1)
  dx = period_days * SEC_PER_DAY / num_integration_bins
  weights = np.ones(num_integration_bins + 1)
  weights[[0, -1]] = 0.5
  days = _j2000_days(ts)[:, None] + offsets[None, :]                     # [T, bins + 1]
  phase, sin_decl, cos_decl, eot, dist = _orbital_parameters(days)
  h0 = 2.0 * np.pi * (phase + eot / SEC_PER_DAY)
  factor = weights[None, :] * tsi[:, None] * (1.0 / dist) ** 2 * dx
  return np.stack([cos_decl, sin_decl, np.cos(h0), np.sin(h0), factor], axis=-1).astype(np.float32)


def get_toa_incident_solar_radiation_device(
    timestamps: Sequence, latitude: np.ndarray, longitude: np.ndarray, device=None,
    tsi_data: Optional[Tuple[np.ndarray, np.ndarray]] = None,
    integration_period=np.timedelta64(1, "h"), num_integration_bins: int = 360):
  """Same quantity as `get_toa_incident_solar_radiation`, computed on the GPU
  (`gcb_toa_incident_solar_radiation`): returns a float32 torch tensor [time, lat, lon] on
  `device`.  The orbital scalars per (timestamp, bin) are prepared here in float64; the kernel
  does the [lat, lon] field work (0.25 degree: one pass, 4 MB written per timestamp)."""
  import torch
  from graphcast_b200 import _native
  lib = _native.lib()
  dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
  table = _integration_table(timestamps, tsi_data, integration_period, num_integration_bins)
  lat = np.radians(np.asarray(latitude, dtype=np.float64))
  lon = np.radians(np.asarray(longitude, dtype=np.float64))
  up = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
  t_tab, s_lat, c_lat, c_lon, s_lon = (up(table), up(np.sin(lat)), up(np.cos(lat)),
                                       up(np.cos(lon)), up(np.sin(lon)))
  out = torch.empty((table.shape[0], lat.shape[0], lon.shape[0]), dtype=torch.float32, device=dev)
  with torch.cuda.device(dev):
    _native.check(lib.gcb_toa_incident_solar_radiation(
        t_tab.data_ptr(), table.shape[0], table.shape[1], s_lat.data_ptr(), c_lat.data_ptr(),
        c_lon.data_ptr(), s_lon.data_ptr(), lat.shape[0], lon.shape[0], out.data_ptr(),
        torch.cuda.current_stream(dev).cuda_stream), "gcb_toa_incident_solar_radiation")
  return out


def device_forcings(names: Sequence[str], datetimes: np.ndarray, dt_dims: Sequence[str],
                    latitude: np.ndarray, longitude: np.ndarray, device=None) -> xs.Dataset:
  """The named forcing variables for the given target datetimes, generated ON THE DEVICE: the
  [lat, lon] field `toa_incident_solar_radiation` by the CUDA kernel (4 MB per timestamp at 0.25
  degree, nothing crosses PCIe), the progress features (a few scalars / one value per longitude
  per timestamp, data_utils.py:51-130) on the host and uploaded (KBs).  Used by
  `rollout.chunked_prediction_generator(..., generate_forcings=...)` so that a long rollout needs no
  per-step forcing transfer.  `datetimes` has dims `dt_dims` = ([batch,] time)."""
  import torch
  dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
  dt_dims = tuple(dt_dims)
  dts = np.asarray(datetimes)
  out = xs.Dataset()
  names = list(names)
  seconds = get_seconds_since_epoch(dts)
  progress: Dict[str, Tuple[Tuple[str, ...], np.ndarray]] = {}
  if any(n.startswith(YEAR_PROGRESS) for n in names):
    progress.update(featurize_progress(YEAR_PROGRESS, dt_dims, get_year_progress(seconds)))
  if any(n.startswith(DAY_PROGRESS) for n in names):
    progress.update(featurize_progress(DAY_PROGRESS, dt_dims + ("lon",),
                                       get_day_progress(seconds, np.asarray(longitude))))
  for n in names:
    if n == TISR:
      field = get_toa_incident_solar_radiation_device(dts.reshape(-1), latitude, longitude, device=dev)
      out[n] = xs.DataArray(field.reshape(tuple(dts.shape) + tuple(field.shape[1:])),
                            dt_dims + ("lat", "lon"))
    elif n in progress:
      dims, vals = progress[n]
      out[n] = xs.DataArray(torch.as_tensor(np.ascontiguousarray(vals, np.float32)).to(dev), dims)
    else:
      raise ValueError(f"cannot generate forcing variable {n!r} (known: {TISR}, "
                       f"{YEAR_PROGRESS}*, {DAY_PROGRESS}*)")
  return out
