"""Premise of bench.py's CPU sample (oracle/sampled_step.py): full fp32 oracle step vs row-sampled
step / fraction, on this host.

  python tests/tools/validate_cpu_sample.py [threads] [resolution]     # resolution 1.0 (default) or 0.25

1.0: BASELINE config 1 (1 deg, mesh 5, 13 levels) in full, 3 repetitions per fraction.
0.25: ONE full pass of the benchmark workload (0.25 deg, mesh 6, 37 levels: 29.3 TFLOP) against the
sample bench.py times (fractions 1/32 and 1/16)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from graphcast_b200 import graph as graph_lib, graphcast, synthetic
from oracle import gnn, sampled_step
threads = int(sys.argv[1]) if len(sys.argv) > 1 else 16
res = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
mesh, task = (5, graphcast.TASK_13) if res == 1.0 else (6, graphcast.TASK)
lat, lon = synthetic.grid_coords(res)
g = graph_lib.cached_static_graph(grid_lat=lat, grid_lon=lon, mesh_size=mesh, radius_query_fraction_edge_length=0.6)
c_in = synthetic.num_input_channels(task); n_out = graphcast.num_outputs(task)
params = gnn.init_params(c_in=c_in, n_out=n_out, msg_steps=16, seed=1)
torch.set_num_threads(threads)
print("threads", torch.get_num_threads(), "cores", os.cpu_count(), "resolution", res)
if res == 1.0:
  for f in (1.0, 0.25, 0.0625):
    full, est = sampled_step.validate(g.as_dict(), params, c_in, f, reps=3)
    print(f"1 deg: fraction {f}: full oracle step {full:.2f} s, sampled / fraction {est:.2f} s, ratio {est / full:.2f}", flush=True)
else:
  gd = g.as_dict()
  x = np.random.default_rng(0).standard_normal((g.num_grid_nodes, 1, c_in)).astype(np.float32)
  orc = gnn.Oracle(params, torch.float32)
  t0 = time.perf_counter()
  orc.forward(gd, x)
  full = time.perf_counter() - t0
  print(f"0.25 deg: ONE full fp32 oracle step (cold: includes first-touch page faults): {full:.1f} s", flush=True)
  for f in (1.0 / 32, 1.0 / 16):
    s = sampled_step.SampledStep(gd, params, c_in, f)
    s.run()
    t = min(s.time_one() for _ in range(3))
    print(f"0.25 deg: fraction {f:.4f}: sample {t:.2f} s, / fraction {t / f:.1f} s, ratio to the full step {t / f / full:.2f}", flush=True)
