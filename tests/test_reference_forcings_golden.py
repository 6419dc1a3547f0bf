"""Forcing generation (`graphcast_b200/forcings.py`) against the reference's own
`utils/data_utils.py` (year / day progress, featurize_progress) and `utils/solar_radiation.py`
(`get_tsi`, `get_toa_incident_solar_radiation`), executed unmodified by
tests/golden/make_golden.py (jnp = numpy, jax.jit = identity, pandas real)."""
import os

import numpy as np
import pytest

from graphcast_b200 import forcings
from graphcast_b200 import xarray_shim as xs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_forcings.npz")


@pytest.fixture(scope="module")
def ref():
  with np.load(GOLDEN) as z:
    return {k: z[k] for k in z.files}


def _stamps(ref):
  return ref["timestamps_ns"].astype("datetime64[ns]")


def test_progress_features(ref):
  seconds = forcings.get_seconds_since_epoch(_stamps(ref))
  yp = forcings.get_year_progress(seconds)
  dp = forcings.get_day_progress(seconds, ref["lon"])
  assert yp.dtype == np.float32 and dp.dtype == np.float32
  np.testing.assert_array_equal(yp, ref["year_progress"])
  np.testing.assert_array_equal(dp, ref["day_progress"])
  feats = forcings.featurize_progress("day_progress", ("time", "lon"), dp)
  for name, (dims, values) in feats.items():
    assert dims == tuple(str(d) for d in ref["feat_dims:" + name])
    np.testing.assert_array_equal(values, ref["feat:" + name])
  with pytest.raises(ValueError, match="feature dimensions"):
    forcings.featurize_progress("x", ("time",), dp)


def test_tsi_interpolation(ref):
  np.testing.assert_allclose(forcings.get_tsi(_stamps(ref), forcings.era5_tsi_data()), ref["tsi"],
                             rtol=1e-13, atol=0)


def test_toa_incident_solar_radiation(ref):
  got = forcings.get_toa_incident_solar_radiation(_stamps(ref), ref["lat"], ref["lon"])
  want = ref["tisr_1h_360"]
  assert got.shape == want.shape == (5, 7, 8)
  assert np.abs(got - want).max() <= 1e-9 * want.max()
  got6 = forcings.get_toa_incident_solar_radiation(
      _stamps(ref)[:2], ref["lat"], ref["lon"], tsi_data=forcings.reference_tsi_data(),
      integration_period=np.timedelta64(6, "h"), num_integration_bins=24)
  want6 = ref["tisr_6h_24_reference_tsi"]
  assert np.abs(got6 - want6).max() <= 1e-9 * want6.max()
  assert (got >= 0).all() and got[:, 0].max() < got.max()      # night side is exactly zero somewhere


def test_add_derived_vars_and_tisr_on_a_dataset(ref):
  stamps = _stamps(ref)[:3]
  ds = xs.Dataset({"t2m": (("batch", "time", "lat", "lon"), np.zeros((1, 3, 7, 8), np.float32))},
                  coords={"lat": ref["lat"], "lon": ref["lon"], "time": np.arange(3),
                          "datetime": (("batch", "time"), stamps[None])})
  forcings.add_derived_vars(ds)
  forcings.add_tisr_var(ds)
  assert ds["year_progress_sin"].dims == ("batch", "time")
  assert ds["day_progress_cos"].dims == ("batch", "time", "lon")
  np.testing.assert_array_equal(np.asarray(ds["day_progress"].data)[0], ref["day_progress"][:3])
  tisr = ds[forcings.TISR]
  assert tisr.dims == ("batch", "time", "lat", "lon") and np.asarray(tisr.data).dtype == np.float32
  np.testing.assert_allclose(np.asarray(tisr.data)[0], ref["tisr_1h_360"][:3], rtol=1e-6)
  with pytest.raises(ValueError, match="datetime"):
    forcings.add_derived_vars(xs.Dataset(coords={"lon": ref["lon"]}))


def test_device_table_reproduces_the_host_integral(ref):
  """The float64 -> float32 bin table of the CUDA version, evaluated with the kernel's own formula
  in numpy (float32 accumulation), gives the host / reference integral."""
  table = forcings._integration_table(_stamps(ref), None, np.timedelta64(1, "h"), 360)
  assert table.shape == (5, 361, 5) and table.dtype == np.float32
  lat, lon = np.radians(ref["lat"]), np.radians(ref["lon"])
  sl, cl = np.sin(lat).astype(np.float32)[:, None], np.cos(lat).astype(np.float32)[:, None]
  cw, sw = np.cos(lon).astype(np.float32)[None, :], np.sin(lon).astype(np.float32)[None, :]
  out = np.zeros((5, 7, 8), np.float32)
  for t in range(5):
    for b in range(361):
      cd, sd, ch, sh, f = table[t, b]
      sin_alt = (cl * cd) * (ch * cw - sh * sw) + sl * sd
      out[t] += f * np.maximum(sin_alt, np.float32(0))
  want = ref["tisr_1h_360"]
  assert np.abs(out - want).max() <= 2e-6 * want.max()


@pytest.mark.gpu
def test_tisr_cuda_kernel_matches_executed_reference(ref):
  import torch
  got = forcings.get_toa_incident_solar_radiation_device(_stamps(ref), ref["lat"], ref["lon"])
  torch.cuda.synchronize()
  want = ref["tisr_1h_360"]
  err = np.abs(got.cpu().numpy() - want).max() / want.max()
  print(f"TISR CUDA kernel vs executed reference: max-abs error / max {err:.2e}")
  assert got.shape == want.shape and err <= 1e-5
  # a larger, ragged grid against the host mirror (non-multiple-of-256 longitude count)
  lat, lon = np.linspace(-90, 90, 91), np.arange(0, 360, 0.7)
  stamps = _stamps(ref)[:2]
  host = forcings.get_toa_incident_solar_radiation(stamps, lat, lon)
  dev = forcings.get_toa_incident_solar_radiation_device(stamps, lat, lon).cpu().numpy()
  assert np.abs(dev - host).max() <= 1e-5 * host.max()
