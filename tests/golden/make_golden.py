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


# ---- checkpoint format (weathernext/utils/checkpoint.py is pure numpy and imports fine) ----
# A GraphCast-CheckPoint-shaped tree (graphcast.py:115-151: params, model_config, task_config,
# description, license) written by the REFERENCE's `checkpoint.dump`; the test loads it with
# this repo's `checkpoint.load` into this repo's dataclasses and re-dumps it.
import dataclasses  # noqa: E402
from typing import Any, Optional  # noqa: E402

from weathernext.utils import checkpoint as ref_ckpt  # noqa: E402


@dataclasses.dataclass(frozen=True)
class TaskConfig:
  input_variables: tuple[str, ...]      # annotations as in weathernext/utils/task.py:21-28
  target_variables: tuple[str, ...]
  forcing_variables: tuple[str, ...]
  pressure_levels: tuple[int, ...]
  input_duration: str


@dataclasses.dataclass(frozen=True)
class ModelConfig:
  resolution: float
  mesh_size: int
  latent_size: int
  gnn_msg_steps: int
  hidden_layers: int
  radius_query_fraction_edge_length: float
  mesh2grid_edge_normalization_factor: Optional[float] = None


@dataclasses.dataclass(frozen=True)
class CheckPoint:
  params: dict[str, Any]
  model_config: ModelConfig
  task_config: TaskConfig
  description: str
  license: str


rng = np.random.default_rng(5)
ck = CheckPoint(
    params={
        "grid2mesh_gnn/~_networks_builder/encoder_nodes_grid_nodes_mlp/~/linear_0":
            {"w": rng.standard_normal((7, 4)).astype(np.float32), "b": np.zeros(4, np.float32)},
        "grid2mesh_gnn/~_networks_builder/encoder_nodes_grid_nodes_layer_norm":
            {"scale": np.ones(4, np.float32), "offset": rng.standard_normal(4).astype(np.float32)},
    },
    model_config=ModelConfig(1.0, 5, 512, 16, 1, 0.6, None),
    task_config=TaskConfig(("2m_temperature", "geopotential"), ("2m_temperature",),
                           ("toa_incident_solar_radiation",), (50, 500, 1000), "12h"),
    description="golden checkpoint written by the reference's checkpoint.dump",
    license="n/a")
path = os.path.join(here, "reference_checkpoint.npz")
with open(path, "wb") as f:
  ref_ckpt.dump(f, ck)
with open(path, "rb") as f:                       # the reference reads its own file back
  back = ref_ckpt.load(f, CheckPoint)
assert back.model_config == ck.model_config and back.task_config == ck.task_config
print("wrote", path)
