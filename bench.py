#!/usr/bin/env python
"""Benchmark of the GraphCast 6 h step on B200 (contract: see DESIGN.md section 6).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # CPU baseline arm

A "step" is one 6 h forecast step of GraphCast 0.25 deg (721x1440, 37 levels,
mesh 6, latent 512, 16 message steps) on synthetic N(0,1) inputs with
Haiku-default random weights.  Prints ONE JSON line on rank 0.

  value    : steps/s with inputs resident in HBM (pack -> step -> unpack), CUDA
             events, max over ranks; N>1 = one independent forecast (ensemble
             member) per GPU, no data-path collective ("weak" scaling).
  e2e      : steps/s through the public API (GraphCast.__call__) with pinned HOST
             inputs: per step H2D of inputs+forcings and D2H of the predictions.
  roofline : the dominant kernel (the tcgen05 fused MLP layer) -- algorithmic
             FLOPs of all its launches in a step / their summed CUDA-event time.
  cpu_baseline : the fp32 CPU oracle (torch-CPU, all host threads) on a bounded
             sample, scaled by the algorithmic FLOP ratio (stated in `sample`).
"""

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv:
  # The CPU arm uses every host thread; torchrun exports OMP_NUM_THREADS=1 to its workers,
  # which would silently make it single-threaded.  Must happen before numpy / torch load.
  for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {
    # name: (resolution, mesh_size, task name)
    "graphcast_0.25deg_37lvl": (0.25, 6, "TASK"),
    "graphcast_operational_0.25deg_13lvl": (0.25, 6, "TASK_13_PRECIP_OUT"),
    "graphcast_small_1deg_13lvl": (1.0, 5, "TASK_13"),
    "sample_2deg_13lvl": (2.0, 4, "TASK_13"),
    "tiny_4deg_13lvl": (4.0, 3, "TASK_13"),
}
DEFAULT_WORKLOAD = "graphcast_0.25deg_37lvl"
REFERENCE_BUDGET_S = 150.0   # wall-clock target of a whole `--impl reference` run


def algorithmic_flops(ng, nm, e1, e2, e3, c_in, n_out, steps, d=512):
  """2*MAC of every MLP in one step, reference dataflow."""
  mlp = lambda rows, d_in, d_out: rows * (d_in * d + d * d_out)
  mac = (mlp(ng, c_in + 3, d) + mlp(nm, c_in + 3, d) + mlp(e1, 4, d) + mlp(e1, 3 * d, d)
         + mlp(nm, 2 * d, d) + mlp(ng, d, d)
         + mlp(e2, 4, d) + steps * (mlp(e2, 3 * d, d) + mlp(nm, 2 * d, d))
         + mlp(e3, 4, d) + mlp(e3, 3 * d, d) + mlp(ng, 2 * d, d) + mlp(ng, d, n_out))
  return 2.0 * mac


class ClockSampler:
  """nvidia-smi clocks / throttle reasons sampled during the timed region."""

  def __init__(self, gpu_index=0):
    self.rows = []
    self.proc = None
    self.gpu_index = gpu_index

  def start(self):
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
      self.proc = subprocess.Popen(
          ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
           "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except OSError:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([x.strip() for x in line.split(",")])

  def stop(self):
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except Exception:
      self.proc.kill()
    sm, smax, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for r in self.rows:
      try:
        sm.append(float(r[0])); smax.append(float(r[1]))
      except (ValueError, IndexError):
        continue
      for name, v in zip(names, r[3:7]):
        if v.lower().startswith("active"):
          reasons.add(name)
    # "under load": samples with clocks above idle
    load = [x for x in sm if x > 500] or sm
    return {"sm_mhz": float(np.median(load)) if load else None,
            "sm_max_mhz": max(smax) if smax else None,
            "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  return rank, world, local


def run_reference(args):
  """CPU arm: the fp32 oracle (restatement of the reference; its JAX stack cannot
  be installed here) on the host cores, bounded sample scaled by FLOPs."""
  rank, world, _ = dist_env()
  if rank != 0:
    return
  import torch
  from graphcast_b200 import graph as graph_lib, graphcast, synthetic
  from oracle import gnn as oracle_gnn
  cores = os.cpu_count() or 1
  res, mesh, task_name = WORKLOADS[args.workload]
  task = getattr(graphcast, task_name)
  # Bounded sample: every step is one full oracle pass over a reduced-resolution instance of
  # the same model, scaled by algorithmic FLOPs.  The whole --steps/--warmup run must end
  # within a few minutes, so fall back to the smaller sample when the requested one would
  # not fit REFERENCE_BUDGET_S (an oracle pass over the 2 degree sample takes 4.5 s with 32 threads on a GPU box's host, 30 s on an 8-core VM).
  cpu_sample = args.cpu_sample
  if cpu_sample == "sample_2deg_13lvl" and (args.steps + args.warmup + 2) * 8.0 > REFERENCE_BUDGET_S:
    cpu_sample = "tiny_4deg_13lvl"
  args.cpu_sample = cpu_sample
  s_res, s_mesh, s_task_name = WORKLOADS[args.cpu_sample]
  s_task = getattr(graphcast, s_task_name)
  lat, lon = synthetic.grid_coords(s_res)
  g = graph_lib.cached_static_graph(grid_lat=lat, grid_lon=lon, mesh_size=s_mesh,
                                    radius_query_fraction_edge_length=0.6)
  c_in = synthetic.num_input_channels(s_task)
  n_out = graphcast.num_outputs(s_task)
  params = oracle_gnn.init_params(c_in=c_in, n_out=n_out, msg_steps=16, seed=1)
  x = np.random.default_rng(0).standard_normal((g.num_grid_nodes, 1, c_in)).astype(np.float32)
  orc = oracle_gnn.Oracle(params, torch.float32)
  gd = g.as_dict()
  sample_flops = algorithmic_flops(g.num_grid_nodes, g.num_mesh_nodes, len(g.g2m_senders),
                                   len(g.mesh_senders), len(g.m2g_senders), c_in, n_out, 16)
  # full-workload FLOPs from the known sizes of the named config
  full = full_workload_sizes(args.workload)
  full_flops = algorithmic_flops(*full)
  scale = full_flops / sample_flops
  # Thread count: the oracle's torch ops regress badly with 128 threads on the GPU boxes' hosts
  # (64 s per 2-degree pass against 4.5 s with 32), so calibrate on one pass each, keep the fastest.
  best = None
  for nthreads in thread_candidates(cores):
    torch.set_num_threads(nthreads)
    tc = time.perf_counter()
    orc.forward(gd, x)
    tc = time.perf_counter() - tc
    if best is None or tc < best[0]:
      best = (tc, nthreads)
  cores = best[1]
  torch.set_num_threads(cores)
  for _ in range(args.warmup):
    orc.forward(gd, x)
  t0 = time.perf_counter()
  for _ in range(args.steps):
    orc.forward(gd, x)
  dt = (time.perf_counter() - t0) / args.steps
  ms = dt * scale * 1e3
  value = 1e3 / ms
  sample = (f"{args.cpu_sample}: full fp32 oracle step ({sample_flops/1e12:.2f} TFLOP, {dt:.2f} s "
            f"measured) scaled x{scale:.2f} by algorithmic FLOPs to {args.workload}")
  line = {
      "impl": "reference", "metric": "6h-step forecasts/sec", "value": value, "unit": "steps/s",
      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
      "data": "synthetic", "config": {"workload": args.workload, "cpu_sample": args.cpu_sample},
      "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port",
                       "sample": sample},
      "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
  }
  print(json.dumps(line), flush=True)


def full_workload_sizes(name):
  """(Ng, Nm, E_g2m, E_mesh, E_m2g, c_in, n_out, steps) of a named workload
  (edge counts: SURVEY.md section 8 table, measured with this repo's builder)."""
  from graphcast_b200 import graphcast, synthetic
  res, mesh, task_name = WORKLOADS[name]
  task = getattr(graphcast, task_name)
  n_lat, n_lon = int(round(180 / res)) + 1, int(round(360 / res))
  ng = n_lat * n_lon
  nm = 10 * 4 ** mesh + 2
  e2 = sum(60 * 4 ** l for l in range(mesh + 1))
  e1 = {(0.25, 6): 1618818, (1.0, 5): 101892}.get((res, mesh))
  if e1 is None:
    e1 = int(1.56 * ng)
  return (ng, nm, e1, e2, 3 * ng, synthetic.num_input_channels(task),
          graphcast.num_outputs(task), 16)


def thread_candidates(cores):
  """Thread counts worth trying for the torch-CPU oracle, most promising first."""
  c = [min(cores, 32), min(cores, 16)]
  if cores <= 64:
    c.append(cores)
  return sorted(set(c), reverse=True)


def cpu_baseline_sample(args, torch):
  """Bounded CPU sample on rank 0 (reported beside the GPU number)."""
  from graphcast_b200 import graph as graph_lib, graphcast, synthetic
  from oracle import gnn as oracle_gnn
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  s_res, s_mesh, s_task_name = WORKLOADS[args.cpu_sample]
  s_task = getattr(graphcast, s_task_name)
  lat, lon = synthetic.grid_coords(s_res)
  g = graph_lib.cached_static_graph(grid_lat=lat, grid_lon=lon, mesh_size=s_mesh,
                                    radius_query_fraction_edge_length=0.6)
  c_in = synthetic.num_input_channels(s_task)
  n_out = graphcast.num_outputs(s_task)
  params = oracle_gnn.init_params(c_in=c_in, n_out=n_out, msg_steps=16, seed=1)
  x = np.random.default_rng(0).standard_normal((g.num_grid_nodes, 1, c_in)).astype(np.float32)
  orc = oracle_gnn.Oracle(params, torch.float32)
  gd = g.as_dict()
  sample_flops = algorithmic_flops(g.num_grid_nodes, g.num_mesh_nodes, len(g.g2m_senders),
                                   len(g.mesh_senders), len(g.m2g_senders), c_in, n_out, 16)
  full_flops = algorithmic_flops(*full_workload_sizes(args.workload))
  # One full pass per candidate thread count (the first doubles as the page-fault warm-up);
  # report the fastest.  128 threads are not tried: 64 s per pass against 4.5 s with 32.
  dt = None
  for nthreads in thread_candidates(cores):
    torch.set_num_threads(nthreads)
    t0 = time.perf_counter()
    orc.forward(gd, x)
    t = time.perf_counter() - t0
    if dt is None or t < dt:
      dt, used = t, nthreads
  cores = used
  scale = full_flops / sample_flops
  return {"value": 1.0 / (dt * scale), "unit": "steps/s", "cores": cores, "kind": "port",
          "sample": (f"{args.cpu_sample}: full fp32 oracle step ({sample_flops/1e12:.2f} TFLOP, "
                     f"{dt:.2f} s measured) scaled x{scale:.2f} by algorithmic FLOPs to "
                     f"{args.workload}")}


def run_b200(args):
  import torch
  import torch.distributed as dist
  from graphcast_b200 import _native, graphcast, synthetic

  rank, world, local = dist_env()
  if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
  torch.cuda.set_device(local)
  dev = torch.device(f"cuda:{local}")
  lib = _native.lib()
  if args.cluster:
    _native.check(lib.gcb_set_cluster_size(args.cluster), "gcb_set_cluster_size")
  if os.environ.get("GCB_DEBUG_FLAGS"):   # kernel experiment switches (gcb_debug_flags); not for results
    _native.check(lib.gcb_debug_flags(int(os.environ["GCB_DEBUG_FLAGS"])), "gcb_debug_flags")

  res, mesh, task_name = WORKLOADS[args.workload]
  task = getattr(graphcast, task_name)
  cfg = graphcast.ModelConfig(resolution=res, mesh_size=mesh, latent_size=512, gnn_msg_steps=16,
                              hidden_layers=1, radius_query_fraction_edge_length=0.6)
  c_in = synthetic.num_input_channels(task)
  t_setup = time.perf_counter()
  inputs, template, forcings = synthetic.make_example(task, res, batch=1, seed=rank,
                                                      pinned=True)
  params = graphcast.init_params(cfg, task, c_in, seed=1)
  model = graphcast.GraphCast(cfg, task, params=params, precision=args.precision, device=dev,
                              pregather=args.pregather, fuse=args.fuse, chain_lag=args.chain_lag,
                              image_residual=args.image_residual)
  # First call builds the static graph, uploads weights, allocates the workspace.
  pred = model(inputs, template, forcings)
  torch.cuda.synchronize()
  eng = model.engine
  setup_s = time.perf_counter() - t_setup
  n_out = eng.n_out
  h2d = sum(int(np.prod(v.shape)) * 4 for ds in (inputs, forcings) for v in ds.data_vars.values())
  d2h = n_out * eng.num_grid * 4

  # ---------------- device-resident timed region --------------------------------
  planes_in = model._planes_in[0]
  planes_out = torch.empty([n_out, eng.num_grid], dtype=torch.float32, device=dev)

  def one_step():
    eng.pack_inputs(planes_in)
    eng.step()
    eng.unpack_outputs(planes_out)

  for _ in range(max(args.warmup, 3)):
    one_step()
  torch.cuda.synchronize()
  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  cap = 256 * args.steps
  lib.gcb_profile_begin()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  ev0.record()
  for _ in range(args.steps):
    one_step()
  ev1.record()
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  elapsed_ms = ev0.elapsed_time(ev1)
  kinds = (C.c_int32 * cap)(); ms = (C.c_float * cap)()
  flops = (C.c_double * cap)(); nbytes = (C.c_double * cap)(); cnt = C.c_int32(0)
  _native.check(lib.gcb_profile_end(cap, kinds, ms, flops, nbytes, C.byref(cnt)), "profile_end")
  n_launch = min(cnt.value, cap)
  per_step_launches = cnt.value // args.steps
  clocks = sampler.stop() if rank == 0 else None

  t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  ms_per_step = float(t.item()) / args.steps
  value = world * 1e3 / ms_per_step

  if args.dump_launches and rank == 0:
    with open(args.dump_launches, "w") as f:
      f.write("idx,kind,ms,gflop,gbyte\n")
      lo = (args.steps - 1) * per_step_launches
      for i in range(lo, min(lo + per_step_launches, n_launch)):
        f.write(f"{i - lo},{kinds[i]},{ms[i]:.4f},{flops[i] / 1e9:.2f},{nbytes[i] / 1e9:.4f}\n")
  # per-kind aggregation (this rank)
  kind_names = {0: "mlp_layer_tc", 1: "segment_sum", 2: "pack", 3: "unpack", 4: "mlp_layer_simt",
                5: "rows_to_image", 6: "mlp_layer_tc"}   # 6 = fused chain launches of the same kernel family
  agg = {}
  for i in range(n_launch):
    a = agg.setdefault(kind_names[kinds[i]], [0.0, 0.0, 0.0, 0])
    a[0] += ms[i]; a[1] += flops[i]; a[2] += nbytes[i]; a[3] += 1
  m = eng._model
  alg_flops = algorithmic_flops(m.num_grid, m.num_mesh, m.e_g2m, m.e_mesh, m.e_m2g, c_in, n_out, 16)
  tc = agg.get("mlp_layer_tc", agg.get("mlp_layer_simt", [1e-9, 0, 0, 0]))
  tc_ms_per_step = tc[0] / args.steps
  achieved_tflops = alg_flops / (tc_ms_per_step * 1e-3) / 1e12
  peaks = {}
  try:
    peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
  except Exception:
    pass
  peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
  peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PF"
  products = {"bf16x3": 3, "bf16": 1, "fp32_simt": 1}[args.precision]
  # MACs the kernel really issues (the split edge layers execute fewer than the reference
  # dataflow the algorithmic figure is defined on), times the products per MAC.
  executed_tflops = (tc[1] / args.steps) * products / (tc_ms_per_step * 1e-3) / 1e12
  # DRAM bytes of this kernel's launches in one step, from the committed ncu pass over the same
  # build and workload (dram__bytes_read.sum + dram__bytes_write.sum, `profiles/r01_launches_v11_ncu.csv`:
  # 110.9 GB read + 110.3 GB written; the algorithmic figure is 254.8 GB, the L2 absorbs part of
  # the gathers).  Only meaningful for the configuration it was captured on.
  traffic, traffic_src = None, None
  if args.workload == DEFAULT_WORKLOAD and args.precision == "bf16x3" and args.pregather:
    traffic, traffic_src = 221.2e9, "ncu, profiles/r01_launches_v11_ncu.csv"
  roofline = {
      "kernel": "gcb::mlp_layer_tc_kernel", "bound": "tensor",
      "achieved": achieved_tflops, "peak": peak_tf, "unit": "TFLOP/s",
      "frac": achieved_tflops / peak_tf, "traffic": traffic, "traffic_unit": "bytes per step (all launches of this kernel)",
      "traffic_source": traffic_src, "peak_source": peak_src,
      "launches_per_step": tc[3] // args.steps, "kernel_ms_per_step": tc_ms_per_step,
      "kernel_share_of_step": tc_ms_per_step / (elapsed_ms / args.steps),
      "algorithmic_tflop_per_step": alg_flops / 1e12,
      "tensor_products_per_mac": products,
      "executed_tflop_per_step": tc[1] / args.steps * products / 1e12,
      "executed_tensor_tflops": executed_tflops,
      "tensor_pipe_frac": executed_tflops / peak_tf,
      "other_kernels_ms_per_step": {k: v[0] / args.steps for k, v in agg.items() if k != "mlp_layer_tc"},
      "hbm": {k: {"GB_per_step": v[2] / args.steps / 1e9,
                  "GBps": (v[2] / 1e9) / (v[0] * 1e-3) if v[0] > 0 else None}
              for k, v in agg.items()},
  }

  # ---------------- end-to-end through the public API ---------------------------
  # Serving loop: every step uploads its inputs from pinned host memory (GraphCast.__call__
  # stages them on its own copy stream, double buffered) and downloads its predictions to
  # pinned host memory on a second copy stream; the host only synchronises at the end, so
  # the transfers of neighbouring steps overlap the kernels.  All K uploads, K steps and K
  # downloads are inside the timed region.
  host_out = [{name: torch.empty(v.shape, dtype=torch.float32, pin_memory=True)
               for name, v in pred.data_vars.items()} for _ in range(2)]
  d2h_stream = torch.cuda.Stream(device=dev)
  compute = torch.cuda.current_stream(dev)

  def e2e_step(i):
    p = model(inputs, template, forcings)           # H2D of every input inside
    done = torch.cuda.Event()
    done.record(compute)
    with torch.cuda.stream(d2h_stream):
      d2h_stream.wait_event(done)
      for name, v in p.data_vars.items():
        v.data.record_stream(d2h_stream)
        host_out[i % 2][name].copy_(v.data, non_blocking=True)   # D2H of the predictions

  e2e_steps = max(2, min(args.steps, args.e2e_steps))
  e2e_step(0)
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  t0 = time.perf_counter()
  for i in range(e2e_steps):
    e2e_step(i)
  torch.cuda.synchronize()
  e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
  t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
  e2e_value = world * 1e3 / float(t.item())

  if rank == 0:
    cpu = None
    if world == 1 and not args.skip_cpu_baseline:
      cpu = cpu_baseline_sample(args, torch)
    line = {
        "metric": "6h-step forecasts/sec", "value": value, "unit": "steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16x3": "bf16x3 (3 bf16 tensor-core products, fp32 accumulate; parity mode)",
                  "bf16": "bf16", "fp32_simt": "f32"}[args.precision],
        "data": "synthetic",
        "config": {"workload": args.workload, "resolution_deg": res, "mesh_size": mesh,
                   "levels": len(task.pressure_levels), "latent": 512, "msg_steps": 16,
                   "batch": 1, "precision": args.precision, "cluster": args.cluster or "default(2)",
                   "pregather": bool(args.pregather), "fuse": bool(args.fuse), "chain_lag": args.chain_lag,
                   "image_residual": bool(args.image_residual),
                   "parallelism": "1 forecast per GPU (ensemble members), no collective",
                   "l2_policy": "working set per step (>20 GB) far exceeds the 126 MB L2; no flush needed",
                   "setup_s": setup_s},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": float(t.item()), "steps": e2e_steps},
        "gpu_launches": per_step_launches * args.steps,
        "gpu_launches_per_step": per_step_launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.destroy_process_group()


def main():
  # Exactly one JSON line may reach stdout: libraries (NCCL's version banner, warnings) are
  # diverted to stderr by pointing fd 1 at fd 2 for the duration of the run.
  real_stdout = os.dup(1)
  os.dup2(2, 1)
  sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
  os.environ["NCCL_DEBUG"] = os.environ.get("GCB_NCCL_DEBUG", "WARN")
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=10)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
  ap.add_argument("--workload", choices=sorted(WORKLOADS), default=DEFAULT_WORKLOAD)
  ap.add_argument("--precision", choices=["bf16x3", "bf16", "fp32_simt"], default="bf16x3")
  ap.add_argument("--cpu-sample", dest="cpu_sample", choices=sorted(WORKLOADS),
                  default="sample_2deg_13lvl",
                  help="bounded CPU sample (a few seconds per step), scaled by algorithmic FLOPs")
  ap.add_argument("--e2e-steps", dest="e2e_steps", type=int, default=10)
  ap.add_argument("--skip-cpu-baseline", action="store_true")
  ap.add_argument("--cluster", type=int, default=0, help="CTAs per cluster (0 = library default)")
  ap.add_argument("--dump-launches", default="", help="write per-launch (kind, ms, GFLOP, GB) of the last timed step to this file")
  ap.add_argument("--no-fuse", dest="fuse", action="store_false",
                  help="one launch per linear layer (hidden activations through HBM)")
  ap.add_argument("--chain-lag", dest="chain_lag", type=int, default=0)
  ap.add_argument("--image-residual", dest="image_residual", action="store_true",
                  help="latent streams as operand images only (no fp32 masters)")
  ap.add_argument("--no-pregather", dest="pregather", action="store_false",
                  help="evaluate the first edge-MLP layer over the concatenated K=1536 input")
  args = ap.parse_args()
  if args.impl == "reference":
    run_reference(args)
  else:
    run_b200(args)


if __name__ == "__main__":
  main()
