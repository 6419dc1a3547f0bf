"""First-contact GPU diagnostics: layer self-tests + small-model parity vs oracle."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
import torch

from graphcast_b200 import _native, engine
from oracle import gnn as oracle_gnn
import _cases

lib = _native.lib()
print("sm_count", lib.gcb_sm_count(0), torch.cuda.get_device_name(0), flush=True)
for prec_name in ("fp32_simt", "bf16x3", "bf16"):
  for rows, k, n in ((128, 16, 256), (128, 512, 512), (1000, 1536, 512), (40000, 480, 512), (777, 512, 256)):
    err = C.c_float(-1)
    t = time.time()
    rc = lib.gcb_selftest_layer(rows, k, n, _native.PRECISIONS[prec_name], C.byref(err))
    print(f"selftest {prec_name:9s} rows={rows:6d} k={k:5d} n={n}: rc={rc} rel_err={err.value:.3e} "
          f"({time.time()-t:.2f}s) {lib.gcb_last_error().decode() if rc else ''}", flush=True)

g, params, x = _cases.small_case()
ref64 = oracle_gnn.Oracle(params, torch.float64).forward(g.as_dict(), x).numpy()
ref32 = oracle_gnn.Oracle(params, torch.float32).forward(g.as_dict(), x).numpy()
print("oracle fp32 vs fp64", np.abs(ref32 - ref64).max() / np.abs(ref64).max(), flush=True)
for prec_name in ("fp32_simt", "bf16x3", "bf16"):
  eng = engine.Engine(g, params, c_in=x.shape[-1], n_out=ref64.shape[-1], msg_steps=3,
                      precision=prec_name)
  y = eng.forward_features(torch.as_tensor(x)).cpu().numpy()
  torch.cuda.synchronize()
  err = np.abs(y - ref64).max() / np.abs(ref64).max()
  print(f"model {prec_name:9s}: max-abs rel err vs fp64 oracle = {err:.3e}, launches={eng.launches_per_step}, "
        f"finite={np.isfinite(y).all()}", flush=True)
