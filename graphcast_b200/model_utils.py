"""Structural graph features and channel packing (host side).

Public surface mirrors the pieces of the reference's
`weathernext/utils/model_utils.py` that the GraphCast path uses:
  get_graph_spatial_features            (:29-152)
  get_bipartite_graph_spatial_features  (:406-544)
  lat_lon_deg_to_spherical / spherical_to_cartesian / ... (:180-234)
  dataset_to_stacked / stacked_to_dataset (:645-776) -- via `channel_layout`.

Construction differs from the reference: the receiver-local rotation
R = Ry(pi/2 - theta_r) . Rz(-phi_r)  (what `Rotation.from_euler("zy", ...)`
builds, reference :369-380) is applied in closed form per edge instead of
materialising one 3x3 scipy rotation matrix per node and an einsum, which is
what makes the 3.1 M-edge mesh2grid graph cheap to build.  Only the GraphCast
configuration of the feature switches is implemented
(`add_node_positions=False, add_node_latitude=True, add_node_longitude=True,
add_relative_positions=True, relative_*_local_coordinates=True`,
reference graphcast.py:186-193).
"""

from __future__ import annotations

from typing import List, Mapping, Optional, Tuple

import numpy as np

from graphcast_b200 import xarray_shim as xs


# -- spherical helpers (reference :180-234) -----------------------------------
def lat_lon_deg_to_spherical(node_lat, node_lon):
  return np.deg2rad(node_lon), np.deg2rad(90 - node_lat)


def spherical_to_lat_lon(phi, theta):
  return 90 - np.rad2deg(theta), np.mod(np.rad2deg(phi), 360)


def cartesian_to_spherical(x, y, z):
  with np.errstate(invalid="ignore"):
    return np.arctan2(y, x), np.arccos(z)


def spherical_to_cartesian(phi, theta):
  return (np.cos(phi) * np.sin(theta), np.sin(phi) * np.sin(theta), np.cos(theta))


def cartesian_to_lat_lon(x, y, z):
  return spherical_to_lat_lon(*cartesian_to_spherical(x, y, z))


def _node_features(phi, theta) -> np.ndarray:
  """[cos(theta)=sin(lat), cos(lon), sin(lon)] (reference :91-104)."""
  return np.stack([np.cos(theta), np.cos(phi), np.sin(phi)], axis=-1)


def _relative_positions_receiver_local(s_phi, s_theta, r_phi, r_theta,
                                       senders, receivers) -> np.ndarray:
  """R_recv . (p_sender - p_receiver) per edge, float64, [E,3].

  The node positions are evaluated in the dtype of the angles (float32 in
  GraphCast, as in the reference :270-271/:588-594), the rotation in float64
  (scipy builds float64 matrices, :378-380).
  """
  s_pos = np.stack(spherical_to_cartesian(s_phi, s_theta), axis=-1)   # [Ns,3]
  r_pos = np.stack(spherical_to_cartesian(r_phi, r_theta), axis=-1)   # [Nr,3]
  # Per-receiver rotation coefficients (float64), gathered per edge below.
  az = -r_phi.astype(np.float64)                    # about z
  pol = -r_theta.astype(np.float64) + np.pi / 2     # about y
  ca, sa, cb, sb = np.cos(az), np.sin(az), np.cos(pol), np.sin(pol)

  num_edges = senders.shape[0]
  out = np.empty([num_edges, 3], dtype=np.float64)
  chunk = 1 << 17                                   # bounds temporary memory
  for lo in range(0, num_edges, chunk):
    snd = senders[lo:lo + chunk]
    rcv = receivers[lo:lo + chunk]
    # Rotation is linear: R.(ps) - R.(pr), each rotated separately as the
    # reference does (:621-642) so rounding matches.
    res = None
    for pos, sign in ((s_pos[snd], 1.0), (r_pos[rcv], -1.0)):
      x = pos[:, 0].astype(np.float64)
      y = pos[:, 1].astype(np.float64)
      z = pos[:, 2].astype(np.float64)
      x1 = ca[rcv] * x - sa[rcv] * y                # Rz(az)
      y1 = sa[rcv] * x + ca[rcv] * y
      x2 = cb[rcv] * x1 + sb[rcv] * z               # Ry(pol)
      z2 = -sb[rcv] * x1 + cb[rcv] * z
      rot = np.stack([x2, y1, z2], axis=-1)
      res = rot if res is None else res - rot
    out[lo:lo + chunk] = res
  return out


def _edge_features(rel: np.ndarray, edge_normalization_factor: Optional[float]
                   ) -> np.ndarray:
  """[|d|, dx, dy, dz] / norm  (reference :121-133 / :524-537)."""
  dist = np.linalg.norm(rel, axis=-1, keepdims=True)
  if edge_normalization_factor is None:
    edge_normalization_factor = dist.max()
  return np.concatenate([dist, rel], axis=-1) / edge_normalization_factor


def get_graph_spatial_features(*, node_lat: np.ndarray, node_lon: np.ndarray,
                               senders: np.ndarray, receivers: np.ndarray,
                               edge_normalization_factor: Optional[float] = None,
                               ) -> Tuple[np.ndarray, np.ndarray]:
  """Node [N,3] and edge [E,4] structural features of a homogeneous graph."""
  phi, theta = lat_lon_deg_to_spherical(node_lat, node_lon)
  rel = _relative_positions_receiver_local(phi, theta, phi, theta,
                                           senders, receivers)
  return _node_features(phi, theta), _edge_features(rel, edge_normalization_factor)


def get_bipartite_graph_spatial_features(
    *, senders_node_lat: np.ndarray, senders_node_lon: np.ndarray,
    senders: np.ndarray, receivers_node_lat: np.ndarray,
    receivers_node_lon: np.ndarray, receivers: np.ndarray,
    edge_normalization_factor: Optional[float] = None,
) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
  """(sender node feats, receiver node feats, edge feats) of a bipartite graph."""
  if senders_node_lat.dtype != receivers_node_lat.dtype:
    raise ValueError("sender/receiver coordinates must share a dtype")
  s_phi, s_theta = lat_lon_deg_to_spherical(senders_node_lat, senders_node_lon)
  r_phi, r_theta = lat_lon_deg_to_spherical(receivers_node_lat, receivers_node_lon)
  rel = _relative_positions_receiver_local(s_phi, s_theta, r_phi, r_theta,
                                           senders, receivers)
  return (_node_features(s_phi, s_theta), _node_features(r_phi, r_theta),
          _edge_features(rel, edge_normalization_factor))


# -- channel packing (reference :645-776, graphcast.py:680-723) ---------------
_PRESERVED = ("batch", "lat", "lon")


class ChannelSlab:
  """One variable's place in the packed channel axis."""

  def __init__(self, name: str, start: int, stack_dims: Tuple[str, ...],
               stack_sizes: Tuple[int, ...], var_dims: Tuple[str, ...],
               stack_labels: Optional[Mapping[str, np.ndarray]] = None):
    self.name = name
    self.start = start
    self.stack_dims = stack_dims
    self.stack_sizes = stack_sizes
    self.var_dims = var_dims
    # index labels of the stacked dims that have them (e.g. "level" -> pressure levels)
    self.stack_labels = dict(stack_labels or {})

  @property
  def count(self) -> int:
    return int(np.prod(self.stack_sizes, dtype=np.int64)) if self.stack_sizes else 1

  def __repr__(self):
    return (f"ChannelSlab({self.name!r}, [{self.start}:{self.start + self.count}), "
            f"stack={dict(zip(self.stack_dims, self.stack_sizes))})")


def channel_layout(dataset: xs.Dataset, start: int = 0) -> List[ChannelSlab]:
  """Channel slabs in `dataset_to_stacked` order: variables sorted by name, each
  flattened over its non-(batch,lat,lon) dims in the variable's own dim order
  (reference :668-674, :699-703)."""
  slabs = []
  offset = start
  for name in sorted(dataset.data_vars.keys()):
    var = dataset.data_vars[name]
    stack_dims = tuple(d for d in var.dims if d not in _PRESERVED)
    stack_sizes = tuple(var.sizes[d] for d in stack_dims)
    labelled = dataset[name]          # with the Dataset's coordinates attached
    labels = {d: labelled.index_labels(d) for d in stack_dims
              if labelled.index_labels(d) is not None}
    slab = ChannelSlab(name, offset, stack_dims, stack_sizes, var.dims, labels)
    slabs.append(slab)
    offset += slab.count
  return slabs


def variable_to_planes(var: xs.DataArray, sizes: Mapping[str, int]):
  """DataArray -> array [batch, channels, lat, lon] (channel-major planes),
  broadcasting any missing batch/lat/lon dims (reference variable_to_stacked
  :645-674 produces the same values in (batch,lat,lon,channels) order)."""
  stack_dims = [d for d in var.dims if d not in _PRESERVED]
  lead = [d for d in ("batch",) if d in var.dims]
  tail = [d for d in ("lat", "lon") if d in var.dims]
  arr = var.transpose(*lead, *stack_dims, *tail).data
  nch = int(np.prod([var.sizes[d] for d in stack_dims], dtype=np.int64)) \
      if stack_dims else 1
  b = var.sizes.get("batch", 1)
  la = var.sizes.get("lat", 1)
  lo = var.sizes.get("lon", 1)
  # Insert singleton axes for the missing preserved dims, then broadcast.
  shape = [b if "batch" in var.dims else 1, nch,
           la if "lat" in var.dims else 1, lo if "lon" in var.dims else 1]
  arr = arr.reshape(shape)
  target = (sizes["batch"], nch, sizes["lat"], sizes["lon"])
  if tuple(arr.shape) == target:
    return arr
  if xs._is_torch(arr):
    return arr.expand(*target)
  return np.broadcast_to(arr, target)


def dataset_to_stacked(dataset: xs.Dataset,
                       sizes: Optional[Mapping[str, int]] = None) -> np.ndarray:
  """Host reference implementation of the packing: [batch, lat, lon, channels]."""
  sizes = sizes or dataset.sizes
  planes = [np.asarray(variable_to_planes(dataset.data_vars[s.name], sizes))
            for s in channel_layout(dataset)]
  stacked = np.concatenate(planes, axis=1)            # [B, C, lat, lon]
  return np.transpose(stacked, (0, 2, 3, 1))


def stacked_to_dataset(stacked, template: xs.Dataset) -> xs.Dataset:
  """Inverse of dataset_to_stacked for a [batch, lat, lon, channels] array
  (numpy or torch); variables/dims/coords follow `template`
  (reference :713-776, including its two ValueErrors)."""
  for name in sorted(template.data_vars.keys()):
    tv = template.data_vars[name]
    if not all(d in tv.dims for d in _PRESERVED):
      raise ValueError(
          f"stacked_to_dataset requires all Variables to have {_PRESERVED} "
          f"dimensions, but found only {tv.dims}.")
  slabs = channel_layout(template)
  expected = sum(s.count for s in slabs)
  found = stacked.shape[-1]
  if expected != found:
    raise ValueError(
        f"Expected {expected} channels but found {found}, when trying to "
        f"convert a stacked array of shape {tuple(stacked.shape)} to a dataset "
        f"of shape {template}.")
  out = xs.Dataset(coords=template.coords)
  for s in slabs:
    piece = stacked[..., s.start:s.start + s.count]
    piece = piece.reshape(tuple(piece.shape[:3]) + s.stack_sizes)
    da = xs.DataArray(piece, _PRESERVED + s.stack_dims)
    out[s.name] = da.transpose(*s.var_dims)
  return out
