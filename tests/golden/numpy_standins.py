"""numpy stand-ins for the third-party packages the reference's GNN code imports (jax, jax.numpy,
jraph, haiku, chex), so that `weathernext/weathernext1_graph/graphcast.py`,
`utils/legacy/deep_typed_graph_net.py` and `utils/typed_graph_net.py` can be EXECUTED here
(tests/golden/make_golden.py) although their real dependencies cannot be installed.

What runs for real is the reference's own code: graph construction, which features are
concatenated in which order, the sender / receiver gathers, the aggregation per receiver, the
residual connections, which node set gets which update, the three-GNN composition.  What is
restated here are only the primitives of the missing libraries, each a few lines, following
their published semantics:

  jax.numpy           -> numpy (plus `repeat(..., total_repeat_length=)`)
  jax.nn.swish        -> x * sigmoid(x)
  jax.tree / tree_util-> flatten / unflatten / map over None, tuples, lists, namedtuples and dicts
                         (dict keys in sorted order, None = empty node; as in JAX)
  jraph.segment_sum   -> zeros + np.add.at
  jraph.concatenated_args -> concatenate the flattened (args, kwargs) leaves on the last axis
  hk.Module           -> base class that records the name of the module whose __call__ is active
  hk.nets.MLP         -> Linear (x @ w + b) stack, activation between layers, none after the last
  hk.LayerNorm        -> (x - mean) * rsqrt(var_biased + 1e-5) * scale + offset over the last axis
  hk.Sequential       -> function composition;  hk.remat -> identity
  chex.dataclass      -> dataclasses.dataclass;  chex.Array -> np.ndarray

Parameters are created on first use from a seeded generator and recorded in PARAMS under the
Haiku-style path `<gnn name>/~_networks_builder/<module name>[/~/linear_<i>]` with entries
{w, b} / {scale, offset}, i.e. the layout of the released checkpoints.
Biases, scales and offsets are randomised (Haiku's defaults 0 / 1 / 0 would hide wiring errors).
"""
import collections
import dataclasses
import sys
import types

import numpy as np

PARAMS = collections.OrderedDict()
MANIFEST = []         # (path, leaf, shape) in creation order: enough to regenerate PARAMS from the seed
SEED = 20240917
_RNG = np.random.default_rng(SEED)
_ACTIVE = []          # names of hk.Modules whose __call__ is executing (innermost last)
DTYPE = np.float32


class _DummyMeta(type):
  """Dummy types whose attributes / subscripts are dummy types again (`hk.initializers.Initializer`,
  `jraph.GraphsTuple`, `X[int]`, `X | None` all evaluate)."""

  def __getattr__(cls, name):
    if name.startswith("__"):
      raise AttributeError(name)
    return _DummyMeta(name, (), {})

  def __getitem__(cls, item):
    return cls


class _Stub(types.ModuleType):
  """Module whose unknown attributes are dummy types (enough for annotations)."""

  def __getattr__(self, name):
    if name.startswith("__"):
      raise AttributeError(name)
    return _DummyMeta(name, (), {})


# ---- pytrees ---------------------------------------------------------------------------------
def _is_namedtuple(x):
  return isinstance(x, tuple) and hasattr(x, "_fields")


def tree_flatten(tree):
  leaves = []

  def walk(x):
    if x is None:
      return ("none",)
    if _is_namedtuple(x):
      return ("namedtuple", type(x), [walk(v) for v in x])
    if isinstance(x, (tuple, list)):
      return (type(x).__name__, [walk(v) for v in x])
    if isinstance(x, dict):
      keys = sorted(x.keys())
      return ("dict", type(x), keys, [walk(x[k]) for k in keys])
    leaves.append(x)
    return ("leaf",)

  return leaves, walk(tree)


def tree_unflatten(treedef, leaves):
  it = iter(leaves)

  def build(d):
    kind = d[0]
    if kind == "none":
      return None
    if kind == "leaf":
      return next(it)
    if kind == "namedtuple":
      return d[1](*[build(c) for c in d[2]])
    if kind == "tuple":
      return tuple(build(c) for c in d[1])
    if kind == "list":
      return [build(c) for c in d[1]]
    if kind == "dict":
      return {k: build(c) for k, c in zip(d[2], d[3])}
    raise ValueError(kind)

  return build(treedef)


def tree_leaves(tree):
  return tree_flatten(tree)[0]


def tree_map(f, tree, *rest):
  leaves, treedef = tree_flatten(tree)
  others = [tree_flatten(r)[0] for r in rest]
  return tree_unflatten(treedef, [f(*xs) for xs in zip(leaves, *others)])


# ---- jax / jax.numpy -------------------------------------------------------------------------
def _repeat(a, repeats, axis=None, total_repeat_length=None):
  out = np.repeat(a, repeats, axis=axis)
  if total_repeat_length is not None:
    assert out.shape[axis or 0] == total_repeat_length
  return out


def _swish(x):
  with np.errstate(over="ignore"):       # exp(-x) -> inf for very negative x: x / inf = -0, as wanted
    return x / (1.0 + np.exp(-x))        # x * sigmoid(x)


def _make_jax():
  jnp = _Stub("jax.numpy")
  for k in dir(np):
    if not k.startswith("_"):
      setattr(jnp, k, getattr(np, k))
  jnp.repeat = _repeat
  tree_util = _Stub("jax.tree_util")
  tree_mod = _Stub("jax.tree")
  for m in (tree_util, tree_mod):
    m.tree_flatten = m.flatten = tree_flatten
    m.tree_unflatten = m.unflatten = tree_unflatten
    m.tree_leaves = m.leaves = tree_leaves
    m.tree_map = m.map = tree_map
  nn = _Stub("jax.nn")
  nn.swish = _swish
  jax = _Stub("jax")
  jax.numpy, jax.tree_util, jax.tree, jax.nn = jnp, tree_util, tree_mod, nn
  return {"jax": jax, "jax.numpy": jnp, "jax.tree_util": tree_util, "jax.tree": tree_mod,
          "jax.nn": nn}


# ---- jraph ------------------------------------------------------------------------------------
def segment_sum(data, segment_ids, num_segments=None, indices_are_sorted=False,
                unique_indices=False):
  del indices_are_sorted, unique_indices
  if num_segments is None:
    num_segments = int(np.max(segment_ids)) + 1
  out = np.zeros((int(num_segments),) + tuple(data.shape[1:]), dtype=data.dtype)
  np.add.at(out, np.asarray(segment_ids), data)
  return out


def concatenated_args(update=None, *, axis=-1):
  def curry(fn):
    def wrapper(*args, **kwargs):
      combined = tree_flatten((args, kwargs))[0]
      return fn(np.concatenate(combined, axis=axis))
    return wrapper
  return curry if update is None else curry(update)


def _make_jraph():
  jraph = _Stub("jraph")
  jraph.segment_sum = segment_sum
  jraph.concatenated_args = concatenated_args
  return {"jraph": jraph}


# ---- haiku -------------------------------------------------------------------------------------
class Module:

  def __init__(self, name=None):
    self.name = name or type(self).__name__

  def __init_subclass__(cls, **kw):
    super().__init_subclass__(**kw)
    call = cls.__dict__.get("__call__")
    if call is not None:
      def wrapped(self, *a, _orig=call, **k):
        _ACTIVE.append(self.name)
        try:
          return _orig(self, *a, **k)
        finally:
          _ACTIVE.pop()
      cls.__call__ = wrapped


def _scope(name):
  if not _ACTIVE:
    raise RuntimeError(f"module {name!r} used outside of an hk.Module call")
  return f"{_ACTIVE[-1]}/~_networks_builder/{name}"


def _truncated_normal(shape, stddev, rng=None):
  rng = rng or _RNG
  x = rng.standard_normal(shape)
  bad = np.abs(x) > 2.0
  while bad.any():
    x[bad] = rng.standard_normal(int(bad.sum()))
    bad = np.abs(x) > 2.0
  return (x * stddev).astype(DTYPE)


def _new_linear(fan_in, size, rng=None):
  rng = rng or _RNG
  return {"w": _truncated_normal((fan_in, size), 1.0 / np.sqrt(fan_in), rng),
          "b": (0.1 * rng.standard_normal(size)).astype(DTYPE)}


def _new_layer_norm(n, rng=None):
  rng = rng or _RNG
  return {"scale": (1.0 + 0.1 * rng.standard_normal(n)).astype(DTYPE),
          "offset": (0.1 * rng.standard_normal(n)).astype(DTYPE)}


def reset(seed):
  """Fresh parameter store and generator (for a second model in the same process)."""
  global _RNG, SEED
  SEED = seed
  _RNG = np.random.default_rng(seed)
  PARAMS.clear()
  del MANIFEST[:]


def regenerate(manifest, seed):
  """The PARAMS a run with this seed created, from its manifest alone (no reference needed):
  entries are drawn from the same generator in the same order with the same shapes."""
  rng = np.random.default_rng(int(seed))
  out = collections.OrderedDict()
  for path, kind, a, b in manifest:
    path, kind = str(path), str(kind)
    out[path] = _new_linear(int(a), int(b), rng) if kind == "linear" else _new_layer_norm(int(a), rng)
  return out


class MLP:
  """hk.nets.MLP: Linear layers with `activation` between them (activate_final=False)."""

  def __init__(self, output_sizes, name=None, activation=None, **unused):
    self.output_sizes = list(output_sizes)
    self.name = name
    self.activation = activation

  def __call__(self, x):
    path = _scope(self.name)
    for i, size in enumerate(self.output_sizes):
      key = f"{path}/~/linear_{i}"
      if key not in PARAMS:
        fan_in = x.shape[-1]
        PARAMS[key] = _new_linear(fan_in, size)
        MANIFEST.append((key, "linear", fan_in, size))
      p = PARAMS[key]
      if i > 0:
        x = self.activation(x)
      x = x @ p["w"] + p["b"]
    return x


class LayerNorm:

  def __init__(self, axis, create_scale, create_offset, eps=1e-5, name=None, **unused):
    assert axis == -1 and create_scale and create_offset
    self.eps = eps
    self.name = name

  def __call__(self, x):
    key = _scope(self.name)
    if key not in PARAMS:
      n = x.shape[-1]
      PARAMS[key] = _new_layer_norm(n)
      MANIFEST.append((key, "layer_norm", n, 0))
    p = PARAMS[key]
    mean = x.mean(axis=-1, keepdims=True)
    var = x.var(axis=-1, keepdims=True)              # biased, like hk.LayerNorm
    inv = p["scale"] / np.sqrt(var + self.eps)
    return inv * (x - mean) + p["offset"]


class Sequential:

  def __init__(self, layers, name=None):
    self.layers = list(layers)

  def __call__(self, x, *args, **kwargs):
    for i, f in enumerate(self.layers):
      x = f(x, *args, **kwargs) if i == 0 else f(x)
    return x


def _make_haiku():
  hk = _Stub("haiku")
  hk.Module = Module
  nets = _Stub("haiku.nets")
  nets.MLP = MLP
  hk.nets = nets
  hk.LayerNorm = LayerNorm
  hk.Sequential = Sequential
  hk.remat = lambda f, **kw: f
  hk.name_like = lambda method_name: (lambda f: f)        # decorator used by utils/dense.py
  return {"haiku": hk, "haiku.nets": nets}


# ---- chex --------------------------------------------------------------------------------------
def _chex_dataclass(cls=None, **kw):
  kw = {k: v for k, v in kw.items() if k in ("frozen", "eq", "init", "repr", "order")}
  return dataclasses.dataclass(cls, **kw) if cls is not None else (
      lambda c: dataclasses.dataclass(c, **kw))


def _make_chex():
  chex = _Stub("chex")
  chex.dataclass = _chex_dataclass
  chex.Array = np.ndarray
  return {"chex": chex}


def install(extra_stubs=("xarray", "xarray.ufuncs", "xarray_jax", "trimesh", "absl", "absl.logging",
                         "dask", "dask.array", "tree")):
  """Registers the stand-ins (and empty stubs for packages the executed code never calls)."""
  mods = {}
  for make in (_make_jax, _make_jraph, _make_haiku, _make_chex):
    mods.update(make())
  for name in extra_stubs:
    mods.setdefault(name, _Stub(name))
  for name, m in mods.items():
    sys.modules[name] = m
  for name in mods:                      # wire submodules as attributes of their parents
    if "." in name:
      parent, child = name.rsplit(".", 1)
      if parent in mods:
        setattr(mods[parent], child, mods[name])
  return mods
