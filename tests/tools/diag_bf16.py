"""Diagnostic: what does precision="bf16" compute?  Layer level (device vs bf16-rounded-operand
fp64 product) and step level (device vs oracle.gnn.Bf16OperandOracle, several engine layouts)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np, torch
import _cases, test_gpu_chain as tc
from graphcast_b200 import _native, engine
from oracle import gnn
lib = _native.lib()
g = torch.Generator().manual_seed(1)
rows = 128 * 40 + 3
x = torch.randn(rows, 512, generator=g)
l0 = tc.Layer(lib, 512, 512, g, ln=False)
out = torch.full((rows, 512), float("nan"), device="cuda:0")
xd = x.to("cuda:0")
img = tc._image(lib, xd, rows, 512)
tc._layer_forward(lib, "bf16", rows, [tc._seg_img(img, 512)], l0, act=False, out=out)
torch.cuda.synchronize()
r = lambda t: t.to(torch.bfloat16).double()
want = r(x) @ r(l0.w) + l0.bias.double()
exact = x.double() @ l0.w.double() + l0.bias.double()
got = out.cpu().double()
print("layer: device vs bf16-operand product", float((got - want).abs().max() / want.abs().max()),
      " vs exact", float((got - exact).abs().max() / exact.abs().max()))
gr, params, xx = _cases.small_case(c_in=31, n_out=23, msg_steps=3, batch=1)
emu = gnn.Bf16OperandOracle(params, torch.float64).forward(gr.as_dict(), xx).numpy()
ref = gnn.Oracle(params, torch.float64).forward(gr.as_dict(), xx).numpy()
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
rms = lambda a, b: float(np.sqrt(np.mean((a - b) ** 2)) / np.sqrt(np.mean(b ** 2)))
for kw in (dict(), dict(deep_chains=False), dict(image_residual=False), dict(fuse=False, image_residual=False),
           dict(fuse=False, image_residual=False, pregather=False)):
  eng = engine.Engine(gr, params, c_in=31, n_out=23, msg_steps=3, precision="bf16", **kw)
  y = eng.forward_features(torch.as_tensor(xx)).cpu().numpy()
  print(kw, "vs emulation max", rel(y, emu), "rms", rms(y, emu), "| vs exact max", rel(y, ref), "rms", rms(y, ref))
print("emulation vs exact max", rel(emu, ref), "rms", rms(emu, ref))
