"""Grid->mesh connectivity at the BASELINE resolutions against the reference's own
`radius_query_indices` run on the reference's own mesh (tests/golden/make_golden.py ->
reference_connectivity.npz: edge count + sha256 of the index arrays at 1 deg / mesh 5 and
0.25 deg / mesh 6).  Index work must be bit-exact: a 1-ulp difference in the mesh vertices
is enough to flip radius-query ties at 0.25 deg (round 1 built 1 618 821 edges, the reference
1 618 818)."""
import hashlib
import os

import numpy as np
import pytest

from graphcast_b200 import graph as graph_lib
from graphcast_b200 import grid_mesh_connectivity as gm
from graphcast_b200 import icosahedral_mesh as im

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_connectivity.npz"))


def _sha(a):
  return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a, np.int64).tobytes()).digest(), np.uint8)


def _grid(res):
  lat = np.linspace(-90, 90, int(round(180 / res)) + 1).astype(np.float32)
  lon = (np.arange(int(round(360 / res))) * res).astype(np.float32)
  return lat, lon


@pytest.mark.parametrize("tag,res,splits", [("1deg_mesh5", 1.0, 5), ("0p25deg_mesh6", 0.25, 6)])
def test_radius_query_matches_the_reference_index_for_index(tag, res, splits):
  lat, lon = _grid(res)
  mesh = im.get_hierarchy_of_triangular_meshes_for_sphere(splits)[-1]
  radius = 0.6 * im.max_edge_length(mesh)
  assert radius == pytest.approx(float(GOLDEN[f"radius_{tag}"]), rel=1e-12)
  g, m = gm.radius_query_indices(grid_latitude=lat, grid_longitude=lon, mesh=mesh, radius=radius)
  assert g.shape[0] == int(GOLDEN[f"num_edges_{tag}"])
  np.testing.assert_array_equal(_sha(g), GOLDEN[f"grid_sha_{tag}"])
  np.testing.assert_array_equal(_sha(m), GOLDEN[f"mesh_sha_{tag}"])


@pytest.mark.parametrize("tag,res,splits", [("1deg_mesh5", 1.0, 5), ("0p25deg_mesh6", 0.25, 6)])
def test_cached_static_graph_has_the_reference_connectivity(tag, res, splits):
  # What bench.py / GraphCast actually load (possibly from .graph_cache).
  lat, lon = _grid(res)
  g = graph_lib.cached_static_graph(grid_lat=lat, grid_lon=lon, mesh_size=splits,
                                    radius_query_fraction_edge_length=0.6)
  assert g.g2m_senders.shape[0] == int(GOLDEN[f"num_edges_{tag}"])
  np.testing.assert_array_equal(_sha(g.g2m_senders), GOLDEN[f"grid_sha_{tag}"])
  np.testing.assert_array_equal(_sha(g.g2m_receivers), GOLDEN[f"mesh_sha_{tag}"])
