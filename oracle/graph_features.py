"""Oracle: structural node / edge features, restated the way the reference
computes them (scipy Rotation matrices per receiver node + einsum).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Used to cross-check the
product's closed-form implementation in graphcast_b200/model_utils.py.

Follows utils/model_utils.py:
  lat_lon_deg_to_spherical :180-186, spherical_to_cartesian :209-216,
  get_rotation_matrices_to_local_coordinates :322-398 (the "zy" branch used by
  GraphCast, :375-380), rotate_with_matrices :401-403,
  get_bipartite_relative_position_in_receiver_local_coordinates :547-642,
  get_bipartite_graph_spatial_features :406-544 (a homogeneous graph is the
  special case senders-set == receivers-set, :29-152).
"""

from __future__ import annotations

import numpy as np
from scipy.spatial import transform


def bipartite_features(s_lat, s_lon, r_lat, r_lon, senders, receivers,
                       edge_normalization_factor=None):
  s_phi, s_theta = np.deg2rad(s_lon), np.deg2rad(90 - s_lat)
  r_phi, r_theta = np.deg2rad(r_lon), np.deg2rad(90 - r_lat)
  s_feats = np.stack([np.cos(s_theta), np.cos(s_phi), np.sin(s_phi)], axis=-1)
  r_feats = np.stack([np.cos(r_theta), np.cos(r_phi), np.sin(r_phi)], axis=-1)

  s_pos = np.stack([np.cos(s_phi) * np.sin(s_theta),
                    np.sin(s_phi) * np.sin(s_theta), np.cos(s_theta)], axis=-1)
  r_pos = np.stack([np.cos(r_phi) * np.sin(r_theta),
                    np.sin(r_phi) * np.sin(r_theta), np.cos(r_theta)], axis=-1)
  rot = transform.Rotation.from_euler(
      "zy", np.stack([-r_phi, -r_theta + np.pi / 2], axis=1)).as_matrix()
  edge_rot = rot[receivers]
  r_rot = np.einsum("...ji,...i->...j", edge_rot, r_pos[receivers])
  s_rot = np.einsum("...ji,...i->...j", edge_rot, s_pos[senders])
  rel = s_rot - r_rot
  dist = np.linalg.norm(rel, axis=-1, keepdims=True)
  if edge_normalization_factor is None:
    edge_normalization_factor = dist.max()
  e_feats = np.concatenate([dist / edge_normalization_factor,
                            rel / edge_normalization_factor], axis=-1)
  return s_feats, r_feats, e_feats
