"""CPU oracle for the GraphCast hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (numpy / torch-CPU, float32 or float64), the
arithmetic of the reference's GraphCast single 6 h step and the host logic
around it.  Every function cites the reference file:line it follows
(paths relative to /root/reference).

Who may import it: `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs -- as the checker or the reported CPU
baseline, never as the product.  Nothing under `graphcast_b200/` imports it;
the product path fails loudly when its CUDA library is missing.

Pinning status:
  * static graph (icosahedral mesh, faces->edges order, grid coordinates):
    PINNED against the reference's own known-answer tests
    (icosahedral_mesh_test.py:36-94, grid_mesh_connectivity_test.py:23-47) and
    against golden vectors generated here by importing the reference's
    `icosahedral_mesh` module (tests/golden/, script tests/golden/make_golden.py).
  * structural features (node features, receiver-local edge features of the three
    graphs; `oracle/graph_features.py`) and the grid2mesh radius query: PINNED
    against golden vectors computed by the reference's own numpy code
    (`model_utils.py`, `legacy/grid_mesh_connectivity.py`, imported with jax /
    xarray / trimesh stubbed out; tests/test_reference_geometry_golden.py).
  * channel packing order and checkpoint format (product-side host logic):
    PINNED against the reference's `dataset_to_stacked` / `stacked_to_dataset`
    and `checkpoint.dump` (tests/golden/reference_packing.npz,
    reference_checkpoint.npz).
  * GNN forward (gather / MLP / LayerNorm / segment_sum / residuals; `oracle/gnn.py`),
    normalisation wrapper and rollout: PARITY UNPINNED -- the reference
    has no test or golden vector for them and its JAX/haiku/jraph/xarray stack
    cannot be installed in this image (no network, not in /opt/wheelhouse), so
    these are restatements reviewed line by line against the cited code, not
    outputs of the executed reference.
"""
