"""Cross-check of the stand-in goldens against the REAL reference stack (run it wherever
`weathernext` + jax + dm-haiku + jraph + xarray ARE installed; they are not in this image).

It rebuilds the tiny case of tests/golden/reference_gnn_forward.npz (10 degree grid, mesh 2,
latent 32, 3 message steps, batch 2) with the real `GraphCast` under `hk.transform`, loads the
parameters stored in that file (they use the real Haiku module paths), runs the three GNNs and
prints the deviation from the stored outputs, which were produced by the same reference code
on numpy stand-ins (tests/golden/numpy_standins.py).  A deviation above fp32 rounding would mean
one of the stand-in primitives misstates its library.

  python tools/dump_reference.py [path/to/reference_gnn_forward.npz]
"""
import os
import sys
import types

import numpy as np


def main():
  here = os.path.dirname(os.path.abspath(__file__))
  path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
      here, "..", "tests", "golden", "reference_gnn_forward.npz")
  z = np.load(path)
  import haiku as hk                      # real packages from here on
  import jax
  from weathernext.weathernext1_graph import graphcast as gc

  params = {}
  for k in z.files:
    if k.startswith("param:"):
      _, module, leaf = k.split(":")
      params.setdefault(module, {})[leaf] = z[k]
  task = gc.TaskConfig(
      input_variables=("2m_temperature", "geopotential", "toa_incident_solar_radiation"),
      target_variables=("2m_temperature", "geopotential"),
      forcing_variables=("toa_incident_solar_radiation",), pressure_levels=(500, 850),
      input_duration="12h")
  cfg = gc.ModelConfig(resolution=10.0, mesh_size=int(z["mesh_size"]), latent_size=32,
                       gnn_msg_steps=int(z["gnn_msg_steps"]), hidden_layers=1,
                       radius_query_fraction_edge_length=0.6)
  sample = types.SimpleNamespace(lat=z["grid_lat"], lon=z["grid_lon"])

  @hk.transform
  def forward(x):
    model = gc.GraphCast(cfg, task)
    model._maybe_init(sample)
    latent_mesh, latent_grid = model._run_grid2mesh_gnn(x)
    updated = model._run_mesh_gnn(latent_mesh)
    return latent_mesh, latent_grid, updated, model._run_mesh2grid_gnn(updated, latent_grid)

  got = forward.apply(params, jax.random.PRNGKey(0), z["grid_features"])
  names = ("latent_mesh_after_grid2mesh", "latent_grid_after_grid2mesh",
           "latent_mesh_after_mesh_gnn", "output")
  worst = 0.0
  for name, y in zip(names, got):
    want = z[name]
    err = float(np.abs(np.asarray(y, np.float64) - want).max() / np.abs(want).max())
    worst = max(worst, err)
    print(f"{name:32s} max-abs relative deviation {err:.3e}")
  print("OK" if worst < 1e-4 else "MISMATCH: a stand-in primitive disagrees with its library")


if __name__ == "__main__":
  main()
