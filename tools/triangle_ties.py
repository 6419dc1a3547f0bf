#!/usr/bin/env python
"""How many grid points lie ON a mesh edge / vertex, i.e. where the containing-triangle lookup
(utils/legacy/grid_mesh_connectivity.py:89-134, trimesh in the reference) has to break a tie between
adjacent triangles.  Prints counts and where they are.  python tools/triangle_ties.py"""
import sys, numpy as np
sys.path.insert(0,'/root/repo')
from graphcast_b200 import graph as graph_lib, synthetic, partitioned
for res, mesh in ((1.0,5),(0.25,6)):
    lat, lon = synthetic.grid_coords(res)
    g = graph_lib.cached_static_graph(grid_lat=lat, grid_lon=lon, mesh_size=mesh, radius_query_fraction_edge_length=0.6)
    xyz_m = partitioned.mesh_xyz(g)
    f = g.grid_node_feats.astype(np.float64)
    cl = np.sqrt(np.maximum(0,1-f[:,0]**2))
    p = np.stack([cl*f[:,1], cl*f[:,2], f[:,0]],1)
    tri = g.m2g_senders.reshape(-1,3)
    a,b,c = xyz_m[tri[:,0]], xyz_m[tri[:,1]], xyz_m[tri[:,2]]
    # barycentric coordinates of the central projection of p onto the triangle's plane
    n = np.cross(b-a, c-a)
    t = (np.einsum('ij,ij->i', a, n) / np.einsum('ij,ij->i', p, n))[:,None]
    q = p*t
    def area(u,v,w): return np.einsum('ij,ij->i', np.cross(v-u, w-u), n)
    tot = area(a,b,c)
    w0, w1, w2 = area(q,b,c)/tot, area(a,q,c)/tot, area(a,b,q)/tot
    m = np.minimum(np.minimum(w0,w1),w2)
    for eps in (1e-12, 1e-9, 1e-7, 1e-6):
        print(res, "grid points with a barycentric weight <", eps, ":", int((m<eps).sum()), "of", m.size, " min", m.min())
    idx = np.flatnonzero(m < 1e-7)
    la = np.degrees(np.arcsin(np.clip(f[idx,0],-1,1))); lo = np.degrees(np.arctan2(f[idx,2], f[idx,1])) % 360
    import collections
    print(" poles:", int((np.abs(la) > 89.99).sum()), " equator:", int((np.abs(la) < 1e-3).sum()))
    c = collections.Counter(np.round(lo[np.abs(la) <= 89.99], 2))
    print(" most common longitudes:", c.most_common(8))
