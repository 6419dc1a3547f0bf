"""Static-graph geometry and structural features against golden vectors computed by the
REFERENCE'S OWN numpy code (tests/golden/make_golden.py imports weathernext.utils.model_utils
and utils/legacy/grid_mesh_connectivity with jax / xarray / trimesh stubbed out; nothing of the
reference is re-implemented there).  Pins: grid coordinates, the grid2mesh radius query,
mesh lat/lon, node features and the receiver-local edge features of all three graphs - for the
product (`graphcast_b200.model_utils`, `grid_mesh_connectivity`, `graph.build_static_graph`)
and for the oracle's scipy-Rotation restatement (`oracle/graph_features.py`)."""
import os

import numpy as np
import pytest

from graphcast_b200 import graph as graph_lib
from graphcast_b200 import grid_mesh_connectivity as gm
from graphcast_b200 import icosahedral_mesh as im
from graphcast_b200 import model_utils
from oracle import graph_features as oracle_features

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_geometry.npz")


@pytest.fixture(scope="module")
def ref():
  with np.load(GOLDEN) as z:
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def finest():
  return im.get_hierarchy_of_triangular_meshes_for_sphere(splits=2)[-1]


def _grid_nodes(ref):
  lon2d, lat2d = np.meshgrid(ref["grid_lon"], ref["grid_lat"])
  return lat2d.reshape(-1).astype(np.float32), lon2d.reshape(-1).astype(np.float32)


def test_grid_coordinates(ref):
  got = gm._grid_lat_lon_to_coordinates(ref["grid_lat"], ref["grid_lon"])
  np.testing.assert_allclose(got, ref["grid_coordinates"], rtol=0, atol=1e-15)


def test_radius_query_same_pairs_same_order(ref, finest):
  g, m = gm.radius_query_indices(grid_latitude=ref["grid_lat"], grid_longitude=ref["grid_lon"],
                                 mesh=finest, radius=float(ref["radius"]))
  assert len(g) == len(ref["g2m_grid_indices"])
  want = set(zip(ref["g2m_grid_indices"].tolist(), ref["g2m_mesh_indices"].tolist()))
  assert set(zip(np.asarray(g).tolist(), np.asarray(m).tolist())) == want
  # grid index is non-decreasing in both (the reference appends per grid point)
  np.testing.assert_array_equal(g, ref["g2m_grid_indices"])


def test_mesh_lat_lon(ref, finest):
  v = finest.vertices
  phi, theta = model_utils.cartesian_to_spherical(v[:, 0], v[:, 1], v[:, 2])
  lat, lon = model_utils.spherical_to_lat_lon(phi=phi, theta=theta)
  np.testing.assert_allclose(lat.astype(np.float32), ref["mesh_lat"], rtol=0, atol=2e-5)
  # longitudes are only defined modulo 360 at the poles / date line
  d = np.abs((lon.astype(np.float64) - ref["mesh_lon"] + 180.0) % 360.0 - 180.0)
  assert d.max() < 2e-5


def _sorted_by_pair(senders, receivers, feats):
  order = np.lexsort((np.asarray(receivers), np.asarray(senders)))
  return np.asarray(feats)[order]


def test_grid2mesh_features_product_and_oracle(ref):
  glat, glon = _grid_nodes(ref)
  s, r = ref["g2m_grid_indices"], ref["g2m_mesh_indices"]
  sn, rn, ef = model_utils.get_bipartite_graph_spatial_features(
      senders_node_lat=glat, senders_node_lon=glon, senders=s,
      receivers_node_lat=ref["mesh_lat"], receivers_node_lon=ref["mesh_lon"], receivers=r)
  np.testing.assert_allclose(sn, ref["g2m_grid_node_feats"], rtol=0, atol=2e-7)
  np.testing.assert_allclose(rn, ref["g2m_mesh_node_feats"], rtol=0, atol=2e-7)
  np.testing.assert_allclose(ef, ref["g2m_edge_feats"], rtol=0, atol=2e-6)
  so, ro, eo = oracle_features.bipartite_features(glat, glon, ref["mesh_lat"], ref["mesh_lon"], s, r)
  np.testing.assert_allclose(so, ref["g2m_grid_node_feats"], rtol=0, atol=2e-7)
  np.testing.assert_allclose(ro, ref["g2m_mesh_node_feats"], rtol=0, atol=2e-7)
  np.testing.assert_allclose(eo, ref["g2m_edge_feats"], rtol=0, atol=2e-6)


def test_mesh_features(ref):
  nf, ef = model_utils.get_graph_spatial_features(
      node_lat=ref["mesh_lat"], node_lon=ref["mesh_lon"],
      senders=ref["mesh_senders"], receivers=ref["mesh_receivers"])
  np.testing.assert_allclose(nf, ref["mesh_node_feats"], rtol=0, atol=2e-7)
  np.testing.assert_allclose(ef, ref["mesh_edge_feats"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("tag,norm", [("", None), ("_norm2", 2.0)])
def test_mesh2grid_features(ref, tag, norm):
  glat, glon = _grid_nodes(ref)
  _, _, ef = model_utils.get_bipartite_graph_spatial_features(
      senders_node_lat=ref["mesh_lat"], senders_node_lon=ref["mesh_lon"],
      senders=ref["m2g_mesh_indices"], receivers_node_lat=glat, receivers_node_lon=glon,
      receivers=ref["m2g_grid_indices"], edge_normalization_factor=norm)
  np.testing.assert_allclose(ef, ref["m2g_edge_feats" + tag], rtol=0, atol=2e-6)


def test_build_static_graph_end_to_end(ref):
  g = graph_lib.build_static_graph(grid_lat=ref["grid_lat"], grid_lon=ref["grid_lon"], mesh_size=2,
                                   radius_query_fraction_edge_length=0.6)
  np.testing.assert_array_equal(g.g2m_senders, ref["g2m_grid_indices"])
  # the same pairs; within one grid point the reference's order is the KD-tree's
  assert (set(zip(g.g2m_senders.tolist(), g.g2m_receivers.tolist()))
          == set(zip(ref["g2m_grid_indices"].tolist(), ref["g2m_mesh_indices"].tolist())))
  np.testing.assert_allclose(g.grid_node_feats, ref["g2m_grid_node_feats"], rtol=0, atol=2e-7)
  np.testing.assert_allclose(g.mesh_node_feats, ref["g2m_mesh_node_feats"], rtol=0, atol=2e-7)
  np.testing.assert_allclose(
      _sorted_by_pair(g.g2m_senders, g.g2m_receivers, g.g2m_edge_feats),
      _sorted_by_pair(ref["g2m_grid_indices"], ref["g2m_mesh_indices"], ref["g2m_edge_feats"]),
      rtol=0, atol=2e-6)
  np.testing.assert_array_equal(g.mesh_senders, ref["mesh_senders"])
  np.testing.assert_array_equal(g.mesh_receivers, ref["mesh_receivers"])
  np.testing.assert_allclose(g.mesh_edge_feats, ref["mesh_edge_feats"], rtol=0, atol=2e-6)
  np.testing.assert_array_equal(g.m2g_receivers, ref["m2g_grid_indices"])
  np.testing.assert_array_equal(g.m2g_senders, ref["m2g_mesh_indices"])
  np.testing.assert_allclose(g.m2g_edge_feats, ref["m2g_edge_feats"], rtol=0, atol=2e-6)
