"""Grid<->mesh connectivity (host side, one-time).

Same public surface as the reference's
`weathernext/utils/legacy/grid_mesh_connectivity.py`:
  radius_query_indices      (:40-86)   grid -> mesh edges of the encoder
  in_mesh_triangle_indices  (:89-134)  mesh -> grid edges of the decoder

Differences in construction (not in contract):
  * the radius query runs as chunked fixed-width k-NN queries bounded by the
    radius (no 1M Python lists), which yields the same pairs in the same order;
  * the reference delegates the containing-triangle query to
    `trimesh.nearest.on_surface`, which is not installed here.  We implement
    the same geometric definition -- the face of the flat triangle mesh that is
    closest (Euclidean) to the grid point -- with a KD-tree over face centroids
    to shortlist candidates and an exact point-to-triangle distance over the
    shortlist.  Grid points that are equidistant from two faces (points lying
    in a mesh symmetry plane) are resolved to the lowest face index; the
    reference's choice there depends on trimesh internals, so connectivity
    parity is checked geometrically (every chosen face is a closest face), not
    index-for-index.
"""

from __future__ import annotations

from typing import Tuple

import numpy as np
import scipy.spatial

from graphcast_b200 import icosahedral_mesh


def _grid_lat_lon_to_coordinates(grid_latitude: np.ndarray,
                                 grid_longitude: np.ndarray) -> np.ndarray:
  """[num_lat],[num_lon] degrees -> unit vectors [num_lat, num_lon, 3]
  (reference :22-37; known-answer test grid_mesh_connectivity_test.py:23-47)."""
  phi = np.deg2rad(grid_longitude)[None, :]
  theta = np.deg2rad(90 - grid_latitude)[:, None]
  sin_t = np.sin(theta)
  return np.stack([np.cos(phi) * sin_t,
                   np.sin(phi) * sin_t,
                   np.cos(theta) * np.ones_like(phi)], axis=-1)


def radius_query_indices(*, grid_latitude: np.ndarray,
                         grid_longitude: np.ndarray,
                         mesh: icosahedral_mesh.TriangularMesh,
                         radius: float) -> Tuple[np.ndarray, np.ndarray]:
  """All (grid, mesh) pairs with straight-line distance <= radius.

  Returns (grid_indices, mesh_indices), grouped by ascending grid index and,
  inside a grid point, ascending mesh index -- the order the reference produces
  (its cKDTree multi-point query sorts each neighbour list, :74-84).
  """
  grid_positions = _grid_lat_lon_to_coordinates(
      grid_latitude, grid_longitude).reshape([-1, 3])
  tree = scipy.spatial.cKDTree(mesh.vertices)
  num_grid = grid_positions.shape[0]
  num_mesh = mesh.vertices.shape[0]
  grid_parts, mesh_parts = [], []
  k = min(8, num_mesh)
  chunk = 1 << 16
  for lo in range(0, num_grid, chunk):
    pts = grid_positions[lo:lo + chunk]
    while True:
      # Fixed-width k-NN query bounded by the radius: missing neighbours come
      # back as index == num_mesh.  Widen k until the last column is empty.
      _, nbr = tree.query(pts, k=k, distance_upper_bound=radius)
      nbr = nbr.reshape(pts.shape[0], -1)
      if k >= num_mesh or np.all(nbr[:, -1] == num_mesh):
        break
      k = min(2 * k, num_mesh)
    nbr = np.sort(nbr, axis=1)                      # ascending mesh index; pads last
    valid = nbr < num_mesh
    rows = np.broadcast_to(np.arange(lo, lo + pts.shape[0])[:, None], nbr.shape)
    grid_parts.append(rows[valid])
    mesh_parts.append(nbr[valid])
  grid_indices = np.concatenate(grid_parts).astype(int)
  mesh_indices = np.concatenate(mesh_parts).astype(int)
  return grid_indices, mesh_indices


def _point_triangle_sqdist(p, a, b, c):
  """Squared distance from points p[...,3] to triangles (a,b,c)[...,3].

  Region-based closest point (Voronoi regions of the triangle), vectorised.
  """
  ab, ac, ap = b - a, c - a, p - a
  d1 = np.sum(ab * ap, -1)
  d2 = np.sum(ac * ap, -1)
  bp = p - b
  d3 = np.sum(ab * bp, -1)
  d4 = np.sum(ac * bp, -1)
  cp = p - c
  d5 = np.sum(ab * cp, -1)
  d6 = np.sum(ac * cp, -1)
  vc = d1 * d4 - d3 * d2
  vb = d5 * d2 - d1 * d6
  va = d3 * d6 - d5 * d4

  with np.errstate(divide="ignore", invalid="ignore"):
    # Interior (default): barycentric projection onto the face.
    denom = va + vb + vc
    v = vb / denom
    w = vc / denom
    closest = a + ab * v[..., None] + ac * w[..., None]
    # Edge BC.
    t_bc = (d4 - d3) / ((d4 - d3) + (d5 - d6))
    on_bc = (va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0)
    closest = np.where(on_bc[..., None], b + (c - b) * t_bc[..., None], closest)
    # Edge AC.
    t_ac = d2 / (d2 - d6)
    on_ac = (vb <= 0) & (d2 >= 0) & (d6 <= 0)
    closest = np.where(on_ac[..., None], a + ac * t_ac[..., None], closest)
    # Edge AB.
    t_ab = d1 / (d1 - d3)
    on_ab = (vc <= 0) & (d1 >= 0) & (d3 <= 0)
    closest = np.where(on_ab[..., None], a + ab * t_ab[..., None], closest)
  # Vertices (highest priority, checked last so they override).
  at_c = (d6 >= 0) & (d5 <= d6)
  closest = np.where(at_c[..., None], c, closest)
  at_b = (d3 >= 0) & (d4 <= d3)
  closest = np.where(at_b[..., None], b, closest)
  at_a = (d1 <= 0) & (d2 <= 0)
  closest = np.where(at_a[..., None], a, closest)
  diff = p - closest
  return np.sum(diff * diff, -1)


def closest_face_indices(points: np.ndarray,
                         mesh: icosahedral_mesh.TriangularMesh,
                         num_candidates: int = 8,
                         chunk: int = 1 << 15) -> np.ndarray:
  """Index of the mesh face closest to each point (ties -> lowest index)."""
  verts = mesh.vertices.astype(np.float64)
  faces = mesh.faces
  tri = verts[faces]                               # [F,3,3]
  centroids = tri.mean(axis=1)
  k = min(num_candidates, faces.shape[0])
  tree = scipy.spatial.cKDTree(centroids)
  out = np.empty([points.shape[0]], dtype=np.int64)
  for lo in range(0, points.shape[0], chunk):
    p = points[lo:lo + chunk].astype(np.float64)
    _, cand = tree.query(p, k=k)
    cand = cand.reshape(p.shape[0], k)
    cand = np.sort(cand, axis=1)                   # ties -> lowest face index
    t = tri[cand]                                  # [n,k,3,3]
    d2 = _point_triangle_sqdist(p[:, None, :], t[:, :, 0], t[:, :, 1], t[:, :, 2])
    # Treat distances equal up to rounding as ties.
    best = d2.min(axis=1, keepdims=True)
    is_best = d2 <= best + 1e-15
    out[lo:lo + chunk] = cand[np.arange(p.shape[0]), np.argmax(is_best, axis=1)]
  return out


def in_mesh_triangle_indices(*, grid_latitude: np.ndarray,
                             grid_longitude: np.ndarray,
                             mesh: icosahedral_mesh.TriangularMesh
                             ) -> Tuple[np.ndarray, np.ndarray]:
  """Edges from the 3 vertices of the closest mesh face to each grid point.

  Returns (grid_indices, mesh_indices) of length 3*num_grid_points, grouped by
  grid point (fixed fan-in 3), as the reference does (:121-134).
  """
  grid_positions = _grid_lat_lon_to_coordinates(
      grid_latitude, grid_longitude).reshape([-1, 3])
  face_idx = closest_face_indices(grid_positions, mesh)
  mesh_indices = mesh.faces[face_idx].reshape([-1]).astype(int)
  grid_indices = np.repeat(np.arange(grid_positions.shape[0]), 3).astype(int)
  return grid_indices, mesh_indices
