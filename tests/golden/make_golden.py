"""Generates tests/golden/*.npz by IMPORTING THE REFERENCE (run in the build
container only: PYTHONPATH=/root/reference python tests/golden/make_golden.py).

The reference's JAX model stack is not installable here, but
`weathernext.utils.icosahedral_mesh` is pure numpy/scipy and imports fine; its
outputs pin the static-graph half of the oracle / product:
  * mesh hierarchy (vertices, faces) for splits 0..3, full arrays;
  * splits 4..6: faces hashed (sha256 of int32 bytes) + vertex checksums;
  * multi-mesh `faces_to_edges` senders/receivers for splits 3 (full) and 6 (hash).
"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
from weathernext.utils import icosahedral_mesh as ref  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))
out = {}
for splits in range(0, 7):
  meshes = ref.get_hierarchy_of_triangular_meshes_for_sphere(splits)
  m = meshes[-1]
  faces = np.ascontiguousarray(m.faces, np.int32)
  verts = np.ascontiguousarray(m.vertices, np.float32)
  if splits <= 3:
    out[f"vertices_{splits}"] = verts
    out[f"faces_{splits}"] = faces
  out[f"faces_sha_{splits}"] = np.frombuffer(hashlib.sha256(faces.tobytes()).digest(), np.uint8)
  out[f"vertex_sum_{splits}"] = verts.astype(np.float64).sum(0)
  out[f"vertex_abs_sum_{splits}"] = np.abs(verts.astype(np.float64)).sum(0)
  merged = ref.merge_meshes(meshes)
  s, r = ref.faces_to_edges(merged.faces)
  s, r = np.ascontiguousarray(s, np.int32), np.ascontiguousarray(r, np.int32)
  if splits == 3:
    out["multimesh_senders_3"], out["multimesh_receivers_3"] = s, r
  out[f"multimesh_edges_sha_{splits}"] = np.frombuffer(
      hashlib.sha256(s.tobytes() + r.tobytes()).digest(), np.uint8)
np.savez_compressed(os.path.join(here, "icosahedral_mesh_reference.npz"), **out)
print("wrote", os.path.join(here, "icosahedral_mesh_reference.npz"))
