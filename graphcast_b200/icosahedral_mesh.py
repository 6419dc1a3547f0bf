"""Icosahedral multi-mesh construction (host side, one-time, numpy only).

Mirrors the public surface of the reference's
`weathernext/utils/icosahedral_mesh.py` (TriangularMesh:46, merge_meshes:79,
get_hierarchy_of_triangular_meshes_for_sphere:98, get_icosahedron:136,
faces_to_edges:366) so that vertex numbering, face order and edge order are
identical to the reference's -- released weights index the mesh by those
numbers.  The implementation is different: the 4-way face split is fully
vectorised (one `np.unique` over the 3F parent edges of a level instead of a
Python loop with a dict per face), which builds the level-6 mesh in ~50 ms
instead of several seconds.

Ordering contract reproduced here (reference `_two_split_unit_sphere_triangle_
faces`:225-263 and `_ChildVerticesBuilder`:321-363):
  * child vertices are appended after all parent vertices, in order of first
    appearance while walking faces in order and, inside a face (v1,v2,v3),
    the edges (v1,v2), (v2,v3), (v3,v1);
  * a child vertex is the float32 midpoint of its two parents, re-projected
    to the unit sphere;
  * each parent face yields, in order, [v1,m12,m31], [m12,v2,m23],
    [m31,m23,v3], [m12,m23,m31].
"""

from __future__ import annotations

from typing import List, NamedTuple, Sequence, Tuple

import numpy as np


class TriangularMesh(NamedTuple):
  """vertices [V,3] float32 on the unit sphere, faces [F,3] int32 (CCW)."""
  vertices: np.ndarray
  faces: np.ndarray


# The 20 faces of the base icosahedron for the vertex numbering produced by
# `get_icosahedron` (counter-clockwise seen from outside).  This table is data
# that fixes the mesh numbering used by released checkpoints
# (reference icosahedral_mesh.py:173-193).
_ICOSAHEDRON_FACES = np.array(
    [(0, 1, 2), (0, 6, 1), (8, 0, 2), (8, 4, 0), (3, 8, 2), (3, 2, 7),
     (7, 2, 1), (0, 4, 6), (4, 11, 6), (6, 11, 5), (1, 5, 7), (4, 10, 11),
     (4, 8, 10), (10, 8, 3), (10, 3, 9), (11, 10, 9), (11, 9, 5), (5, 9, 7),
     (9, 3, 7), (1, 6, 5)], dtype=np.int32)


def _rotation_about_y(angle: float) -> np.ndarray:
  """Active rotation matrix about +y (same matrix scipy's from_euler('y') gives)."""
  c, s = np.cos(angle), np.sin(angle)
  return np.array([[c, 0., s], [0., 1., 0.], [-s, 0., c]], dtype=np.float64)


def get_icosahedron(pole_parallel_faces: bool = True) -> TriangularMesh:
  """Regular icosahedron inscribed in the unit sphere.

  Vertex order follows reference icosahedral_mesh.py:161-170: for c1 in (1,-1),
  for c2 in (phi,-phi): (c1,c2,0), (0,c1,c2), (c2,0,c1); normalised in float32.
  With `pole_parallel_faces` the solid is rotated about y so that the top and
  bottom faces are parallel to the x-y plane (no vertex at the poles)
  (reference :195-219).
  """
  phi = (1.0 + np.sqrt(5.0)) / 2.0
  rows = []
  for c1 in (1.0, -1.0):
    for c2 in (phi, -phi):
      rows += [(c1, c2, 0.0), (0.0, c1, c2), (c2, 0.0, c1)]
  vertices = np.array(rows, dtype=np.float32)
  vertices /= np.linalg.norm([1.0, phi])
  if pole_parallel_faces:
    angle_between_faces = 2.0 * np.arcsin(phi / np.sqrt(3.0))
    rot = _rotation_about_y((np.pi - angle_between_faces) / 2.0)
    # Row vectors times matrix, as in the reference (`np.dot(vertices, R)`).
    vertices = np.dot(vertices, rot)
  return TriangularMesh(vertices=vertices.astype(np.float32),
                        faces=_ICOSAHEDRON_FACES.copy())


def _split_faces_once(mesh: TriangularMesh) -> TriangularMesh:
  """One 4-way split of every face, vectorised; ordering as documented above."""
  v = mesh.vertices
  f = mesh.faces.astype(np.int64)
  num_parent = v.shape[0]
  # The 3F directed parent edges in walk order: per face (v1,v2),(v2,v3),(v3,v1).
  a = f[:, [0, 1, 2]].reshape(-1)
  b = f[:, [1, 2, 0]].reshape(-1)
  lo = np.minimum(a, b)
  hi = np.maximum(a, b)
  key = lo * num_parent + hi
  uniq, first_pos, inverse = np.unique(key, return_index=True,
                                       return_inverse=True)
  # Child index = rank of the edge's first appearance in walk order.
  order = np.argsort(first_pos, kind="stable")
  rank = np.empty_like(order)
  rank[order] = np.arange(order.shape[0])
  child_index = (num_parent + rank[inverse]).reshape(-1, 3)   # [F,3]: m12,m23,m31
  # Positions: parents listed in the order of the first appearance (a then b),
  # float32 arithmetic as in the reference (mean of two float32 rows, then
  # division by the float32 norm).
  pa = a[first_pos[order]]
  pb = b[first_pos[order]]
  mid = (v[pa] + v[pb]) / np.float32(2.0)
  # The reference divides every child by `np.linalg.norm(child)` (:344), i.e. sqrt(BLAS sdot)
  # of a 3-vector; a vectorised sum of squares rounds differently in the last bit for some
  # vertices, which is enough to flip radius-query ties at 0.25 degree.  Same call, per vertex.
  norm = np.empty([mid.shape[0]], np.float32)
  for i in range(mid.shape[0]):
    norm[i] = np.linalg.norm(mid[i])
  mid = (mid / norm[:, None]).astype(np.float32)
  vertices = np.concatenate([v, mid], axis=0)

  v1, v2, v3 = f[:, 0], f[:, 1], f[:, 2]
  m12, m23, m31 = child_index[:, 0], child_index[:, 1], child_index[:, 2]
  faces = np.stack([
      np.stack([v1, m12, m31], -1),
      np.stack([m12, v2, m23], -1),
      np.stack([m31, m23, v3], -1),
      np.stack([m12, m23, m31], -1),
  ], axis=1).reshape(-1, 3).astype(np.int32)
  return TriangularMesh(vertices=vertices, faces=faces)


def get_hierarchy_of_triangular_meshes_for_sphere(
    splits: int, pole_parallel_faces: bool = True) -> List[TriangularMesh]:
  """Meshes M0 (icosahedron) ... M_splits, coarse to fine (reference :98-133)."""
  meshes = [get_icosahedron(pole_parallel_faces=pole_parallel_faces)]
  for _ in range(splits):
    meshes.append(_split_faces_once(meshes[-1]))
  return meshes


def get_last_triangular_mesh_for_sphere(splits: int) -> TriangularMesh:
  return get_hierarchy_of_triangular_meshes_for_sphere(splits)[-1]


def merge_meshes(mesh_list: Sequence[TriangularMesh]) -> TriangularMesh:
  """Multi-mesh: finest vertices, faces of every level concatenated coarse->fine
  (reference :79-95)."""
  for coarse, fine in zip(mesh_list[:-1], mesh_list[1:]):
    n = coarse.vertices.shape[0]
    if not np.allclose(coarse.vertices, fine.vertices[:n]):
      raise ValueError("meshes are not nested: coarse vertices must be a prefix")
  return TriangularMesh(
      vertices=mesh_list[-1].vertices,
      faces=np.concatenate([m.faces for m in mesh_list], axis=0))


def faces_to_edges(faces: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
  """Directed edges of all faces: all v0->v1, then all v1->v2, then all v2->v0
  (reference :366-388; order pinned by its test :72-94)."""
  if faces.ndim != 2 or faces.shape[-1] != 3:
    raise ValueError("faces must be [F,3]")
  senders = faces.T.reshape(-1)            # f[:,0] | f[:,1] | f[:,2]
  receivers = np.roll(faces, -1, axis=1).T.reshape(-1)   # f[:,1] | f[:,2] | f[:,0]
  return senders.copy(), receivers.copy()


def max_edge_length(mesh: TriangularMesh) -> np.float32:
  """Longest edge of the mesh in R^3 (reference graphcast.py:733-737).  Returned as the
  numpy float32 scalar the reference's `_get_max_edge_distance` yields, so that
  `max_edge * radius_query_fraction_edge_length` (graphcast.py:266-267) goes through the
  same numpy promotion (float32 under NEP 50) and the query radius is bit-identical."""
  s, r = faces_to_edges(mesh.faces)
  return np.linalg.norm(mesh.vertices[s] - mesh.vertices[r], axis=-1).max()
