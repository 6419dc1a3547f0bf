"""The C-ABI library loads and exports every symbol include/graphcast_b200.h
declares; host-only entry points and argument validation (no GPU compute)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from graphcast_b200 import _native

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                      "include", "graphcast_b200.h")


def _declared_functions():
  text = open(HEADER).read()
  text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
  return sorted(set(re.findall(r"\b(gcb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
  lib = _native.lib()
  names = _declared_functions()
  assert len(names) >= 12
  for name in names:
    assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert name in _native.EXPORTS, f"{name} has no ctypes signature in _native.EXPORTS"
  assert lib.gcb_abi_version() == _native.GCB_ABI_VERSION == 2


def test_missing_library_fails_loudly(monkeypatch):
  monkeypatch.setattr(_native, "_lib", None)
  monkeypatch.setattr(_native, "_LIB_PATH", "/nonexistent/libgraphcast_b200.so")
  with pytest.raises(_native.NativeLibraryError, match="no CPU / PyTorch fallback"):
    _native.lib()


def _bf16_to_f32(u16):
  return (u16.astype(np.uint32) << 16).view(np.float32)


def test_pack_weight_host_layout_and_split_accuracy():
  lib = _native.lib()
  k_real, n_real, k, n = 20, 227, 32, 256
  w = np.random.default_rng(0).standard_normal((k_real, n_real)).astype(np.float32)
  assert lib.gcb_packed_weight_bytes(k, n) == k * n * 4
  img = np.zeros(k * n * 4, np.uint8)
  assert lib.gcb_pack_weight_host(w.ctypes.data, k_real, n_real, k, n, img.ctypes.data) == 0
  im = img.view(np.uint16).reshape(k // 16, n // 256, 2, 2, 256, 8)   # kstep, block, hi|lo, chunk, row, j
  hi, lo = _bf16_to_f32(im[:, :, 0]), _bf16_to_f32(im[:, :, 1])        # [ks, h, c, r, j]
  rec = (hi + lo).transpose(0, 2, 4, 1, 3).reshape(k, n)               # [ks, c, j, h, r] -> [k, n]
  np.testing.assert_allclose(rec[:k_real, :n_real], w, rtol=2 ** -16, atol=1e-30)
  assert np.all(rec[k_real:] == 0) and np.all(rec[:, n_real:] == 0)
  # hi is the round-to-nearest-even bf16 of w
  want_hi = (((w.view(np.uint32) + 0x7fff + ((w.view(np.uint32) >> 16) & 1)) >> 16) << 16).view(np.float32)
  np.testing.assert_array_equal(hi.transpose(0, 2, 4, 1, 3).reshape(k, n)[:k_real, :n_real], want_hi)
  # 512-wide layer: two 256-column blocks per K-step
  w2 = np.random.default_rng(1).standard_normal((32, 512)).astype(np.float32)
  img2 = np.zeros(32 * 512 * 4, np.uint8)
  assert lib.gcb_pack_weight_host(w2.ctypes.data, 32, 512, 32, 512, img2.ctypes.data) == 0
  im2 = img2.view(np.uint16).reshape(2, 2, 2, 2, 256, 8)
  rec2 = (_bf16_to_f32(im2[:, :, 0]) + _bf16_to_f32(im2[:, :, 1])).transpose(0, 2, 4, 1, 3).reshape(32, 512)
  np.testing.assert_allclose(rec2, w2, rtol=2 ** -16, atol=1e-30)
  assert lib.gcb_pack_weight_host(w.ctypes.data, k_real, n_real, 24, n, img.ctypes.data) == -1
  assert b"multiple of 16" in lib.gcb_last_error()


def test_layer_validation_rejects_bad_descriptors():
  lib = _native.lib()
  d = _native.LayerDesc()
  d.rows, d.n, d.n_valid, d.nseg = 10, 300, 300, 1
  assert lib.gcb_layer_forward(C.byref(d), None) == -1
  assert b"n must be 256 or 512" in lib.gcb_last_error()
  d.n, d.n_valid = 512, 512
  d.seg[0].table, d.seg[0].ld, d.seg[0].k, d.seg[0].k_valid, d.seg[0].fan = 256, 8, 20, 8, 1
  assert lib.gcb_layer_forward(C.byref(d), None) == -1
  assert b"multiple of 16" in lib.gcb_last_error()
  with pytest.raises(ValueError):
    _native.check(-1, "x")


def test_struct_layouts_match_the_header():
  """ctypes mirrors vs the C structs: sizes computed by a C compiler from the header itself."""
  import subprocess
  import tempfile
  src = ('#include <stdio.h>\n#include "graphcast_b200.h"\n'
         'int main(void){printf("%zu %zu %zu %zu %zu\\n", sizeof(gcb_segment), sizeof(gcb_layer_desc),'
         ' sizeof(gcb_chain_layer), sizeof(gcb_chain_desc), sizeof(gcb_model)); return 0;}\n')
  with tempfile.TemporaryDirectory() as d:
    c = os.path.join(d, "s.c")
    open(c, "w").write(src)
    exe = os.path.join(d, "s")
    subprocess.run(["gcc", "-I", os.path.dirname(HEADER), c, "-o", exe], check=True)
    sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
  assert sizes == [C.sizeof(_native.Segment), C.sizeof(_native.LayerDesc), C.sizeof(_native.ChainLayer),
                   C.sizeof(_native.ChainDesc), C.sizeof(_native.Model)]
