"""`data_utils.extract_inputs_targets_forcings` / `extract_input_target_times` against the
reference's own functions (weathernext/utils/data_utils.py:214-333), executed unmodified with
real pandas Timedeltas on a stand-in example batch (tests/golden/make_golden.py)."""
import datetime
import os

import numpy as np
import pytest

from graphcast_b200 import data_utils
from graphcast_b200 import xarray_shim as xs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_data_utils.npz")
TASK = dict(input_variables=("2m_temperature", "geopotential", "toa_incident_solar_radiation", "land_sea_mask"),
            target_variables=("2m_temperature", "geopotential"),
            forcing_variables=("toa_incident_solar_radiation",), pressure_levels=(500, 1000),
            input_duration="12h")


@pytest.fixture(scope="module")
def ref():
  with np.load(GOLDEN) as z:
    return {k: z[k] for k in z.files}


def _example(ref):
  names = [k[3:] for k in ref if k.startswith("in:")]
  return xs.Dataset(
      {n: (tuple(str(d) for d in ref[f"in_dims:{n}"]), ref[f"in:{n}"]) for n in names},
      coords={"time": ref["time_ns"].astype("timedelta64[ns]"), "level": ref["level"],
              "lat": np.linspace(-45, 45, 3), "lon": np.arange(4) * 90.0,
              "datetime": (("batch", "time"), ref["datetime_ns"].astype("datetime64[ns]"))})


@pytest.mark.parametrize("tag,lead", [("slice", slice("6h", "18h")), ("list", ["12h"])])
def test_extract_inputs_targets_forcings_matches_executed_reference(ref, tag, lead):
  parts = data_utils.extract_inputs_targets_forcings(_example(ref), target_lead_times=lead, **TASK)
  for part_name, part in zip(("inputs", "targets", "forcings"), parts):
    assert sorted(part.data_vars.keys()) == [str(n) for n in ref[f"{tag}:{part_name}:names"]]
    assert "datetime" not in part.coords
    np.testing.assert_array_equal(
        np.asarray(part.coords["time"][1]).astype("timedelta64[ns]").astype(np.int64),
        ref[f"{tag}:{part_name}:time_ns"])
    for name in part.data_vars.keys():
      assert part[name].dims == tuple(str(d) for d in ref[f"{tag}:{part_name}_dims:{name}"])
      np.testing.assert_array_equal(np.asarray(part[name].data), ref[f"{tag}:{part_name}:{name}"])


def test_timedelta_parsing_and_errors(ref):
  h = np.timedelta64(3600 * 10**9, "ns")
  assert data_utils.to_timedelta("6h") == 6 * h
  assert data_utils.to_timedelta("1 day") == 24 * h
  assert data_utils.to_timedelta("5d12h") == 132 * h
  assert data_utils.to_timedelta("24 hours") == 24 * h
  assert data_utils.to_timedelta(datetime.timedelta(hours=3)) == 3 * h
  assert data_utils.to_timedelta(np.timedelta64(90, "m")) == np.timedelta64(5400 * 10**9, "ns")
  with pytest.raises(ValueError):
    data_utils.to_timedelta("six hours")
  with pytest.raises(ValueError, match="overlap"):
    data_utils.extract_inputs_targets_forcings(
        _example(ref), target_lead_times="6h",
        **dict(TASK, forcing_variables=("2m_temperature",)))
  with pytest.raises(KeyError):
    data_utils.extract_inputs_targets_forcings(
        _example(ref), target_lead_times="6h", **dict(TASK, pressure_levels=(123,)))


def test_generates_missing_forcings(ref):
  """A task asking for progress features / TISR that are not in the batch gets them generated
  (data_utils.py:313-316) before the split."""
  ex = _example(ref)
  ex = ex[[k for k in ex.data_vars.keys() if k != "toa_incident_solar_radiation"]]
  task = dict(TASK, input_variables=("2m_temperature", "geopotential", "toa_incident_solar_radiation",
                                     "year_progress_sin", "day_progress_cos"),
              forcing_variables=("toa_incident_solar_radiation", "year_progress_sin", "day_progress_cos"))
  inputs, targets, forcings = data_utils.extract_inputs_targets_forcings(ex, target_lead_times="6h", **task)
  assert forcings["toa_incident_solar_radiation"].dims == ("batch", "time", "lat", "lon")
  assert forcings["toa_incident_solar_radiation"].shape == (2, 1, 3, 4)
  assert inputs["day_progress_cos"].dims == ("batch", "time", "lon") and inputs["day_progress_cos"].shape == (2, 2, 4)
  assert forcings["year_progress_sin"].shape == (2, 1)
