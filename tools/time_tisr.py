"""Times the TISR CUDA kernel (gcb_toa_incident_solar_radiation) at 0.25 degrees."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphcast_b200 import _native, forcings
lib = _native.lib()
dev = torch.device("cuda:0")
T = 8
stamps = (np.datetime64("2020-06-01T00:00", "ns") + np.arange(T) * np.timedelta64(6, "h"))
lat, lon = np.linspace(-90, 90, 721), np.arange(0, 360, 0.25)
table = forcings._integration_table(stamps, None, np.timedelta64(1, "h"), 360)
up = lambda a: torch.as_tensor(np.ascontiguousarray(a, np.float32)).to(dev)
la, lo = np.radians(lat), np.radians(lon)
t_tab, s_lat, c_lat, c_lon, s_lon = up(table), up(np.sin(la)), up(np.cos(la)), up(np.cos(lo)), up(np.sin(lo))
out = torch.empty((T, 721, 1440), dtype=torch.float32, device=dev)
call = lambda: _native.check(lib.gcb_toa_incident_solar_radiation(
    t_tab.data_ptr(), T, table.shape[1], s_lat.data_ptr(), c_lat.data_ptr(), c_lon.data_ptr(),
    s_lon.data_ptr(), 721, 1440, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "tisr")
for _ in range(3): call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): call()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
host = forcings.get_toa_incident_solar_radiation(stamps[:1], lat, lon)
err = np.abs(out[0].cpu().numpy() - host[0]).max() / host.max()
print(f"TISR kernel, {T} timestamps x 721 x 1440, 361 bins: {ms:.3f} ms per call = {ms / T * 1e3:.1f} us per field, "
      f"{T * 721 * 1440 * 4 / ms / 1e6:.0f} GB/s written, {T*721*1440*361*7/ms/1e9:.1f} TFLOP/s (fp32, 7 flop per bin); vs host mirror {err:.1e}")
