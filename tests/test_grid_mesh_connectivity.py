"""Connectivity: the reference's known answer for grid coordinates
(grid_mesh_connectivity_test.py:23-47), its two smoke tests, and the non-smoke
checks it leaves as TODO (brute-force radius query, closest-face geometry)."""
import numpy as np

from graphcast_b200 import grid_mesh_connectivity as gm
from graphcast_b200 import icosahedral_mesh as im


def test_grid_lat_lon_to_coordinates():
  lat = np.array([-45., 0., 45])
  lon = np.array([0., 90., 180., 270.])
  q = 1 / np.sqrt(2)
  expected = np.array([
      [[q, 0., -q], [0., q, -q], [-q, 0., -q], [0., -q, -q]],
      [[1., 0., 0.], [0., 1., 0.], [-1., 0., 0.], [0., -1., 0.]],
      [[q, 0., q], [0., q, q], [-q, 0., q], [0., -q, q]]])
  np.testing.assert_allclose(gm._grid_lat_lon_to_coordinates(lat, lon), expected, atol=1e-15)


def _setup():
  lat = np.linspace(-75, 75, 6)
  lon = np.arange(12) * 30.
  mesh = im.get_hierarchy_of_triangular_meshes_for_sphere(splits=3)[-1]
  return lat, lon, mesh


def test_radius_query_matches_brute_force():
  lat, lon, mesh = _setup()
  g, m = gm.radius_query_indices(grid_latitude=lat, grid_longitude=lon, mesh=mesh, radius=0.2)
  pts = gm._grid_lat_lon_to_coordinates(lat, lon).reshape(-1, 3)
  d = np.linalg.norm(pts[:, None, :] - mesh.vertices[None], axis=-1)
  eg, em = np.nonzero(d <= 0.2)          # row-major: grouped by grid, ascending mesh
  np.testing.assert_array_equal(g, eg)
  np.testing.assert_array_equal(m, em)


def test_radius_query_empty_and_ragged():
  lat, lon, mesh = _setup()
  g, m = gm.radius_query_indices(grid_latitude=lat, grid_longitude=lon, mesh=mesh, radius=1e-6)
  assert g.shape == (0,) and m.shape == (0,)
  g, m = gm.radius_query_indices(grid_latitude=lat, grid_longitude=lon, mesh=mesh, radius=0.6)
  counts = np.bincount(g, minlength=lat.size * lon.size)
  assert counts.min() > 8 and counts.max() > counts.min()      # exercises the k-widening loop


def test_in_mesh_triangle_is_closest_face():
  lat, lon, mesh = _setup()
  g, m = gm.in_mesh_triangle_indices(grid_latitude=lat, grid_longitude=lon, mesh=mesh)
  n = lat.size * lon.size
  assert g.shape == (3 * n,) and m.shape == (3 * n,)
  np.testing.assert_array_equal(g, np.repeat(np.arange(n), 3))
  pts = gm._grid_lat_lon_to_coordinates(lat, lon).reshape(-1, 3)
  tri = mesh.vertices.astype(np.float64)[mesh.faces]
  d2 = gm._point_triangle_sqdist(pts[:, None, :], tri[None, :, 0], tri[None, :, 1], tri[None, :, 2])
  chosen = m.reshape(n, 3)
  face_of = {tuple(f): i for i, f in enumerate(mesh.faces)}
  for i in range(n):
    fi = face_of[tuple(chosen[i])]
    assert d2[i, fi] <= d2[i].min() + 1e-12


def test_cached_static_graph_round_trip(tmp_path):
  """The on-disk cache keeps only the two spatial-query results; a graph rebuilt from it must
  equal a fresh build bit for bit, and a corrupt cache file must fall back to building."""
  import dataclasses
  import os
  from graphcast_b200 import graph as graph_lib
  lat = np.linspace(-90, 90, 19, dtype=np.float32)
  lon = np.arange(0, 360, 10, dtype=np.float32)
  kw = dict(grid_lat=lat, grid_lon=lon, mesh_size=2, radius_query_fraction_edge_length=0.6)
  fresh = graph_lib.build_static_graph(**kw)
  first = graph_lib.cached_static_graph(cache_dir=str(tmp_path), **kw)     # builds + writes
  files = os.listdir(tmp_path)
  assert len(files) == 1 and files[0].startswith("connectivity_")
  again = graph_lib.cached_static_graph(cache_dir=str(tmp_path), **kw)     # reads
  for f in dataclasses.fields(fresh):
    a, b, c = (getattr(x, f.name) for x in (fresh, first, again))
    if isinstance(a, np.ndarray):
      assert a.dtype == b.dtype == c.dtype
      np.testing.assert_array_equal(a, b)
      np.testing.assert_array_equal(a, c)
    else:
      assert a == b == c
  with open(os.path.join(tmp_path, files[0]), "wb") as fh:
    fh.write(b"not an npz")
  rebuilt = graph_lib.cached_static_graph(cache_dir=str(tmp_path), **kw)
  np.testing.assert_array_equal(rebuilt.m2g_senders, fresh.m2g_senders)
