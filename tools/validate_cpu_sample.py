"""Premise of bench.py's CPU sample (oracle/sampled_step.py): full fp32 oracle step vs row-sampled
step / fraction on BASELINE config 1 (1 deg, mesh 5, 13 levels), on this host."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphcast_b200 import graph as graph_lib, graphcast, synthetic
from oracle import gnn, sampled_step
lat, lon = synthetic.grid_coords(1.0)
g = graph_lib.cached_static_graph(grid_lat=lat, grid_lon=lon, mesh_size=5, radius_query_fraction_edge_length=0.6)
task = graphcast.TASK_13
c_in = synthetic.num_input_channels(task); n_out = graphcast.num_outputs(task)
params = gnn.init_params(c_in=c_in, n_out=n_out, msg_steps=16, seed=1)
torch.set_num_threads(int(sys.argv[1]) if len(sys.argv) > 1 else 16)
print("threads", torch.get_num_threads(), "cores", os.cpu_count())
for f in (1.0, 0.25, 0.0625):
  full, est = sampled_step.validate(g.as_dict(), params, c_in, f, reps=3)
  print(f"1 deg: fraction {f}: full oracle step {full:.2f} s, sampled / fraction {est:.2f} s, ratio {est / full:.2f}", flush=True)
