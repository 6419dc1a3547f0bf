"""Minimal labelled-array containers standing in for `xarray`.

The reference's public API (`GraphCast.__call__`, `rollout.chunked_prediction`,
`normalization.InputsAndResiduals`) speaks `xarray.Dataset`.  xarray is not
installed in this image, so the host-side mirror uses these two small classes,
which implement exactly the subset of the xarray surface the hot path touches
(reference graphcast.py:680-723, model_utils.py:645-776, rollout.py:416-604,
normalization.py:29-160): named dims, dims-aware broadcasting arithmetic,
`isel`, `transpose`, `assign_coords`, `assign`, `copy`, and time-axis concat.
Array payloads may be numpy arrays (host) or torch tensors (device-resident
state during a rollout); nothing here touches the GPU itself.

If a real `xarray.Dataset` is passed to the public entry points it is converted
with `from_xarray` (duck-typed: `.data_vars`, `.dims`, `.values`, `.coords`).
"""

from __future__ import annotations

from typing import Any, Dict, Iterable, Mapping, Optional, Sequence, Tuple

import numpy as np

try:  # torch is optional for the shim itself.
  import torch
except Exception:  # pragma: no cover
  torch = None


def _is_torch(x) -> bool:
  return torch is not None and isinstance(x, torch.Tensor)


def _permute(x, order: Sequence[int]):
  return x.permute(*order) if _is_torch(x) else np.transpose(x, order)


def _expand_dims(x, axis: int):
  return x.unsqueeze(axis) if _is_torch(x) else np.expand_dims(x, axis)


class DataArray:
  """N-d array with named dims and (optional) coordinates."""

  __array_priority__ = 100

  def __init__(self, data, dims: Sequence[str],
               coords: Optional[Mapping[str, Any]] = None,
               name: Optional[str] = None):
    if not _is_torch(data):
      data = np.asarray(data)
    dims = tuple(dims)
    if len(dims) != data.ndim:
      raise ValueError(f"dims {dims} do not match data of rank {data.ndim}")
    self.data = data
    self.dims = dims
    self.name = name
    self.coords: Dict[str, Tuple[Tuple[str, ...], np.ndarray]] = {}
    for k, v in (coords or {}).items():
      self.coords[k] = _as_coord(k, v)

  # -- basic properties ------------------------------------------------------
  @property
  def shape(self):
    return tuple(self.data.shape)

  @property
  def ndim(self):
    return len(self.dims)

  @property
  def dtype(self):
    return self.data.dtype

  @property
  def sizes(self) -> Dict[str, int]:
    return dict(zip(self.dims, self.shape))

  @property
  def values(self) -> np.ndarray:
    if _is_torch(self.data):
      return self.data.detach().cpu().numpy()
    return self.data

  def __repr__(self):
    return f"DataArray(name={self.name!r}, dims={self.dims}, shape={self.shape})"

  def copy(self, data=None) -> "DataArray":
    return DataArray(self.data if data is None else data, self.dims,
                     dict(self.coords), self.name)

  def astype(self, dtype) -> "DataArray":
    if _is_torch(self.data):
      return self.copy(self.data.to(dtype))
    return self.copy(self.data.astype(dtype))

  # -- indexing / layout -----------------------------------------------------
  def isel(self, indexers: Optional[Mapping[str, Any]] = None, drop: bool = False,
           **kw) -> "DataArray":
    indexers = dict(indexers or {}, **kw)
    index = [slice(None)] * self.ndim
    kept_dims = list(self.dims)
    for dim, sel in indexers.items():
      if dim not in self.dims:
        continue
      index[self.dims.index(dim)] = sel
      if isinstance(sel, (int, np.integer)):
        kept_dims.remove(dim)
    data = self.data[tuple(index)]
    coords = {}
    for k, (cdims, cval) in self.coords.items():
      cidx = tuple(indexers.get(d, slice(None)) for d in cdims)
      new_cdims = tuple(d for d in cdims
                        if not isinstance(indexers.get(d, slice(None)),
                                          (int, np.integer)))
      cnew = cval[cidx] if cdims else cval
      if not new_cdims and cdims and drop:
        continue
      coords[k] = (new_cdims, np.asarray(cnew))
    return DataArray(data, kept_dims, coords, self.name)

  def transpose(self, *dims) -> "DataArray":
    if Ellipsis in dims:
      i = dims.index(Ellipsis)
      rest = [d for d in self.dims if d not in dims]
      dims = tuple(dims[:i]) + tuple(rest) + tuple(dims[i + 1:])
    if not dims:
      dims = tuple(reversed(self.dims))
    if set(dims) != set(self.dims) or len(dims) != len(self.dims):
      raise ValueError(f"transpose dims {dims} are not a permutation of {self.dims}")
    order = [self.dims.index(d) for d in dims]
    return DataArray(_permute(self.data, order), dims, dict(self.coords), self.name)

  def assign_coords(self, coords: Optional[Mapping[str, Any]] = None, **kw
                    ) -> "DataArray":
    out = self.copy()
    for k, v in dict(coords or {}, **kw).items():
      out.coords[k] = _as_coord(k, v)
    return out

  # -- dims-aware arithmetic ---------------------------------------------------
  def index_labels(self, dim: str) -> Optional[np.ndarray]:
    """Labels of `dim` (its 1-d index coordinate), or None when the dim is unlabelled."""
    c = self.coords.get(dim)
    if c is not None and c[0] == (dim,) and c[1].shape[0] == self.sizes.get(dim, -1):
      return c[1]
    return None

  def _binary(self, other, op):
    if isinstance(other, DataArray):
      # Label alignment like xarray's arithmetic (inner join on every shared, labelled
      # dim): the released 37-level `*_by_level` statistics must combine with 13-level
      # inputs by pressure level, not by position (reference normalization.py:29-70
      # relies on this).
      lhs, rhs = self, other
      for d in self.dims:
        if d not in other.dims:
          continue
        la, lb = self.index_labels(d), other.index_labels(d)
        if la is None or lb is None:
          continue
        if la.shape == lb.shape and np.array_equal(la, lb):
          continue
        pos_b = {v.item() if hasattr(v, "item") else v: i for i, v in enumerate(lb)}
        keep_a = [i for i, v in enumerate(la) if (v.item() if hasattr(v, "item") else v) in pos_b]
        keep_b = [pos_b[(la[i].item() if hasattr(la[i], "item") else la[i])] for i in keep_a]
        lhs = lhs.isel({d: np.asarray(keep_a, np.int64)})
        rhs = rhs.isel({d: np.asarray(keep_b, np.int64)})
      if lhs is not self or rhs is not other:
        return lhs._binary_aligned(rhs, op)
      return self._binary_aligned(other, op)
    return DataArray(op(self.data, other), self.dims, dict(self.coords), self.name)

  def _binary_aligned(self, other, op):
    if isinstance(other, DataArray):
      out_dims = list(self.dims) + [d for d in other.dims if d not in self.dims]
      a = _broadcast_to_dims(self, out_dims)
      b = _broadcast_to_dims(other, out_dims)
      if _is_torch(a) != _is_torch(b):
        if _is_torch(a):
          b = torch.as_tensor(np.ascontiguousarray(b), device=a.device)
        else:
          a = torch.as_tensor(np.ascontiguousarray(a), device=b.device)
      coords = dict(other.coords)
      coords.update(self.coords)
      return DataArray(op(a, b), out_dims, coords, self.name)
    return DataArray(op(self.data, other), self.dims, dict(self.coords), self.name)

  def __add__(self, o): return self._binary(o, lambda a, b: a + b)
  def __sub__(self, o): return self._binary(o, lambda a, b: a - b)
  def __mul__(self, o): return self._binary(o, lambda a, b: a * b)
  def __truediv__(self, o): return self._binary(o, lambda a, b: a / b)
  __radd__ = __add__
  __rmul__ = __mul__


def _as_coord(name: str, value) -> Tuple[Tuple[str, ...], np.ndarray]:
  if isinstance(value, tuple) and len(value) == 2 and (
      isinstance(value[0], (str, tuple, list))):
    dims, arr = value
    dims = (dims,) if isinstance(dims, str) else tuple(dims)
    return dims, np.asarray(arr)
  if isinstance(value, DataArray):
    return value.dims, np.asarray(value.values)
  arr = np.asarray(value)
  return ((name,) if arr.ndim == 1 else ()), arr


def _broadcast_to_dims(arr: DataArray, out_dims: Sequence[str]):
  data = arr.data
  present = [d for d in out_dims if d in arr.dims]
  data = _permute(data, [arr.dims.index(d) for d in present])
  for axis, d in enumerate(out_dims):
    if d not in arr.dims:
      data = _expand_dims(data, axis)
  return data


class Dataset:
  """Ordered mapping name -> DataArray sharing coordinates."""

  def __init__(self, data_vars: Optional[Mapping[str, Any]] = None,
               coords: Optional[Mapping[str, Any]] = None):
    self._vars: Dict[str, DataArray] = {}
    self.coords: Dict[str, Tuple[Tuple[str, ...], np.ndarray]] = {}
    for k, v in (coords or {}).items():
      self.coords[k] = _as_coord(k, v)
    for name, v in (data_vars or {}).items():
      self[name] = v

  # -- mapping protocol --------------------------------------------------------
  def __setitem__(self, name: str, value):
    if isinstance(value, tuple):
      dims, data = value
      value = DataArray(data, dims)
    if not isinstance(value, DataArray):
      raise TypeError("Dataset values must be DataArray or (dims, data)")
    value = value.copy()
    value.name = name
    for k, c in value.coords.items():
      self.coords.setdefault(k, c)
    self._vars[name] = value

  def __getitem__(self, key):
    if isinstance(key, (list, tuple)):
      return Dataset({k: self._vars[k] for k in key}, self.coords)
    v = self._vars[key]
    out = v.copy()
    out.coords = {k: c for k, c in self.coords.items()
                  if all(d in v.dims for d in c[0])}
    return out

  def __contains__(self, key):
    return key in self._vars

  def __iter__(self):
    return iter(self._vars)

  def __len__(self):
    return len(self._vars)

  def keys(self):
    return self._vars.keys()

  def items(self):
    return ((k, self[k]) for k in self._vars)

  @property
  def data_vars(self) -> Dict[str, DataArray]:
    return self._vars

  @property
  def variables(self) -> Dict[str, DataArray]:
    return self._vars

  def __getattr__(self, name):
    # coordinate access as attribute, e.g. ds.lat, ds.lon
    coords = self.__dict__.get("coords", {})
    if name in coords:
      dims, arr = coords[name]
      return DataArray(arr, dims, name=name)
    raise AttributeError(name)

  def __repr__(self):
    body = ", ".join(f"{k}{v.dims}" for k, v in self._vars.items())
    return f"Dataset({body}; sizes={self.sizes})"

  # -- structure ---------------------------------------------------------------
  @property
  def sizes(self) -> Dict[str, int]:
    out: Dict[str, int] = {}
    for dims, arr in self.coords.values():
      for d, n in zip(dims, arr.shape):
        out.setdefault(d, n)
    for v in self._vars.values():
      for d, n in v.sizes.items():
        if out.setdefault(d, n) != n:
          raise ValueError(f"conflicting sizes for dim {d!r}")
    return out

  @property
  def dims(self) -> Dict[str, int]:
    return self.sizes

  def copy(self) -> "Dataset":
    return Dataset(dict(self._vars), dict(self.coords))

  def isel(self, indexers: Optional[Mapping[str, Any]] = None, drop: bool = False,
           **kw) -> "Dataset":
    indexers = dict(indexers or {}, **kw)
    out = Dataset()
    for k, (cdims, cval) in self.coords.items():
      cidx = tuple(indexers.get(d, slice(None)) for d in cdims)
      new_cdims = tuple(d for d in cdims
                        if not isinstance(indexers.get(d, slice(None)),
                                          (int, np.integer)))
      out.coords[k] = (new_cdims, np.asarray(cval[cidx] if cdims else cval))
    for name, v in self._vars.items():
      sub = v.isel(indexers, drop=drop)
      sub.coords = {}
      out._vars[name] = sub
    return out

  def assign_coords(self, coords: Optional[Mapping[str, Any]] = None, **kw
                    ) -> "Dataset":
    out = self.copy()
    for k, v in dict(coords or {}, **kw).items():
      out.coords[k] = _as_coord(k, v)
    return out

  def assign(self, other: Optional[Mapping[str, Any]] = None, **kw) -> "Dataset":
    out = self.copy()
    src = dict(other.items()) if isinstance(other, Dataset) else dict(other or {})
    src.update(kw)
    for k, v in src.items():
      out[k] = v
    if isinstance(other, Dataset):
      for k, c in other.coords.items():
        out.coords.setdefault(k, c)
    return out

  def compute(self) -> "Dataset":
    return self

  def map(self, fn) -> "Dataset":
    return Dataset({k: fn(self[k]) for k in self._vars}, self.coords)


def concat_time(datasets: Iterable[Dataset]) -> Dataset:
  """Concatenate datasets along "time" (variables without a time dim are taken
  from the first dataset), cf. `xarray.concat(..., dim="time")`."""
  datasets = list(datasets)
  first = datasets[0]
  out = Dataset(coords={k: c for k, c in first.coords.items()
                        if "time" not in c[0]})
  for k, (cdims, _) in first.coords.items():
    if "time" in cdims:
      axis = cdims.index("time")
      out.coords[k] = (cdims, np.concatenate(
          [np.asarray(d.coords[k][1]) for d in datasets], axis=axis))
  for name, v in first.data_vars.items():
    if "time" not in v.dims:
      out._vars[name] = v.copy()
      continue
    axis = v.dims.index("time")
    parts = [d.data_vars[name].data for d in datasets]
    if any(_is_torch(p) for p in parts):
      dev = next(p.device for p in parts if _is_torch(p))
      parts = [p if _is_torch(p) else torch.as_tensor(p, device=dev) for p in parts]
      data = torch.cat(parts, dim=axis)
    else:
      data = np.concatenate(parts, axis=axis)
    out._vars[name] = DataArray(data, v.dims, name=name)
  return out


def from_xarray(ds) -> Dataset:
  """Duck-typed conversion of a real xarray.Dataset (if one is ever passed)."""
  if isinstance(ds, Dataset):
    return ds
  coords = {k: (tuple(c.dims), np.asarray(c.values)) for k, c in ds.coords.items()}
  return Dataset({k: DataArray(np.asarray(v.values), tuple(v.dims))
                  for k, v in ds.data_vars.items()}, coords)
