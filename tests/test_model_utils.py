"""Structural features (closed form vs the oracle's scipy-Rotation restatement) and
channel packing order / round trip."""
import numpy as np
import pytest

from graphcast_b200 import graphcast, model_utils, synthetic
from graphcast_b200 import xarray_shim as xs
from oracle import graph_features as oracle_features


def test_bipartite_features_match_oracle():
  rng = np.random.default_rng(0)
  s_lat = rng.uniform(-90, 90, 50).astype(np.float32); s_lon = rng.uniform(0, 360, 50).astype(np.float32)
  r_lat = rng.uniform(-90, 90, 40).astype(np.float32); r_lon = rng.uniform(0, 360, 40).astype(np.float32)
  snd = rng.integers(0, 50, 300); rcv = rng.integers(0, 40, 300)
  got = model_utils.get_bipartite_graph_spatial_features(
      senders_node_lat=s_lat, senders_node_lon=s_lon, senders=snd,
      receivers_node_lat=r_lat, receivers_node_lon=r_lon, receivers=rcv)
  want = oracle_features.bipartite_features(s_lat, s_lon, r_lat, r_lon, snd, rcv)
  for a, b in zip(got, want):
    np.testing.assert_allclose(a, b, atol=1e-6)
  # explicit normalisation factor
  got = model_utils.get_bipartite_graph_spatial_features(
      senders_node_lat=s_lat, senders_node_lon=s_lon, senders=snd,
      receivers_node_lat=r_lat, receivers_node_lon=r_lon, receivers=rcv,
      edge_normalization_factor=0.5)
  want = oracle_features.bipartite_features(s_lat, s_lon, r_lat, r_lon, snd, rcv, 0.5)
  np.testing.assert_allclose(got[2], want[2], atol=1e-6)


def test_receiver_local_frame_properties():
  # In the receiver's frame the receiver sits at (1,0,0): relative position of a
  # sender at the same point is 0, and |rel| equals the chord length.
  lat = np.array([10., -35., 80.], np.float32); lon = np.array([20., 200., 355.], np.float32)
  snd = np.array([0, 1, 2, 0]); rcv = np.array([0, 1, 2, 1])
  _, e = model_utils.get_graph_spatial_features(node_lat=lat, node_lon=lon, senders=snd,
                                                receivers=rcv, edge_normalization_factor=1.0)
  np.testing.assert_allclose(e[:3], 0, atol=1e-6)
  p = np.stack(model_utils.spherical_to_cartesian(*model_utils.lat_lon_deg_to_spherical(lat, lon)), -1)
  np.testing.assert_allclose(e[3, 0], np.linalg.norm(p[0] - p[1]), rtol=1e-5)


def test_channel_order_matches_reference_contract():
  inputs, template, forcings = synthetic.make_example(graphcast.TASK_13, 10.0)
  slabs = model_utils.channel_layout(inputs)
  assert [s.name for s in slabs] == sorted(inputs.data_vars.keys())
  by = {s.name: s for s in slabs}
  assert by["geopotential"].stack_dims == ("time", "level") and by["geopotential"].count == 26
  assert by["land_sea_mask"].count == 1 and by["year_progress_sin"].count == 2
  assert sum(s.count for s in slabs) + 5 == synthetic.num_input_channels(graphcast.TASK_13) == 183
  stacked = model_utils.dataset_to_stacked(inputs)
  # time-major, level-minor inside a variable
  g = inputs.data_vars["geopotential"].values        # (batch,time,level,lat,lon)
  s = by["geopotential"].start
  np.testing.assert_array_equal(stacked[0, :, :, s + 1 * 13 + 4], g[0, 1, 4])
  # variables without lat are broadcast
  dp = inputs.data_vars["day_progress_sin"].values    # (batch,time,lon)
  np.testing.assert_array_equal(stacked[0, 3, :, by["day_progress_sin"].start], dp[0, 0])
  yp = inputs.data_vars["year_progress_cos"].values   # (batch,time)
  assert np.all(stacked[0, :, :, by["year_progress_cos"].start + 1] == yp[0, 1])


def test_stacked_round_trip_and_errors():
  _, template, _ = synthetic.make_example(graphcast.TASK_13, 10.0, batch=2)
  n = sum(s.count for s in model_utils.channel_layout(template))
  assert n == graphcast.num_outputs(graphcast.TASK_13) == 83
  stacked = np.random.default_rng(0).standard_normal((2, 19, 36, n)).astype(np.float32)
  ds = model_utils.stacked_to_dataset(stacked, template)
  assert ds.data_vars["temperature"].dims == ("batch", "time", "level", "lat", "lon")
  np.testing.assert_array_equal(model_utils.dataset_to_stacked(ds), stacked)
  with pytest.raises(ValueError, match="Expected 83 channels but found 82"):
    model_utils.stacked_to_dataset(stacked[..., :-1], template)
  bad = xs.Dataset({"x": xs.DataArray(np.zeros((2, 3)), ("batch", "time"))})
  with pytest.raises(ValueError, match="requires all Variables"):
    model_utils.stacked_to_dataset(stacked, bad)
