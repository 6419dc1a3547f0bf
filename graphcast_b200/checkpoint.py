"""Checkpoint (de)serialisation -- same on-disk format as the reference's
`weathernext/utils/checkpoint.py` (`dump` :26-39, `load` :42-54): one `.npz`
whose keys are the ':'-joined paths of a tree of dicts / dataclasses / lists /
tuples (:57-82), reloaded through a dataclass schema (:98-170), so released
`GraphCast*.npz` files load directly into `graphcast.CheckPoint`."""

from __future__ import annotations

import dataclasses
import io
import types
import typing
from typing import Any, BinaryIO, Dict

import numpy as np

_SEP = ":"


def _children(node):
  if dataclasses.is_dataclass(node) and not isinstance(node, type):
    return {f.name: getattr(node, f.name) for f in dataclasses.fields(node)
            if getattr(node, f.name) is not None}
  if isinstance(node, (list, tuple)):
    return {str(i): v for i, v in enumerate(node)}
  if isinstance(node, dict):
    return {str(k): v for k, v in node.items()}
  return None


def flatten(tree: Any) -> Dict[str, Any]:
  flat: Dict[str, Any] = {}

  def walk(prefix, node):
    kids = _children(node)
    if kids is None:
      if node is None:
        raise ValueError(f"None leaf at {prefix!r} (None is only allowed as a dataclass field)")
      flat[prefix] = node
      return
    for k, v in kids.items():
      if _SEP in k:
        raise ValueError(f"key {k!r} contains the separator {_SEP!r}")
      walk(k if not prefix else f"{prefix}{_SEP}{k}", v)

  kids = _children(tree)
  if kids is None:
    raise TypeError("top-level value must be a dict, dataclass, list or tuple")
  walk("", tree)
  return flat


def unflatten(flat: Dict[str, Any]) -> Dict[str, Any]:
  tree: Dict[str, Any] = {}
  for path, value in flat.items():
    *parents, leaf = path.split(_SEP)
    node = tree
    for p in parents:
      node = node.setdefault(p, {})
    node[leaf] = value
  return tree


def dump(dest: BinaryIO, value: Any) -> None:
  buf = io.BytesIO()
  np.savez(buf, **flatten(value))
  dest.write(buf.getvalue())


def load(source: BinaryIO, typ: type) -> Any:
  with np.load(source) as z:
    flat = {k: z[k] for k in z.files}
  return _as_type(typ, unflatten(flat))


def _ordered(values: Dict[str, Any]):
  return [v for _, v in sorted(values.items(), key=lambda kv: int(kv[0]))]


def _as_type(typ, value):
  if typ in (Any, ...):
    return value
  if typ in (int, float, str, bool):
    return typ(value)
  if typ is np.ndarray:
    if not isinstance(value, np.ndarray):
      raise TypeError("expected an array")
    return value
  if dataclasses.is_dataclass(typ):
    hints = typing.get_type_hints(typ)
    kwargs = {}
    for f in dataclasses.fields(typ):
      ftype = hints.get(f.name, f.type)
      origin = typing.get_origin(ftype)
      if origin in (typing.Union, types.UnionType):
        options = [a for a in typing.get_args(ftype) if a is not type(None)]
        if len(options) != 1:
          raise TypeError("Optional works, Union with anything except None doesn't")
        if f.name not in value:
          kwargs[f.name] = None
          continue
        ftype = options[0]
      if f.name not in value:
        raise ValueError(f"Missing value: {f.name}")
      kwargs[f.name] = _as_type(ftype, value[f.name])
    return typ(**kwargs)
  origin = typing.get_origin(typ)
  args = typing.get_args(typ)
  if origin is dict:
    kt, vt = args
    return {_as_type(kt, k): _as_type(vt, v) for k, v in value.items()}
  if origin is list:
    return [_as_type(args[0], v) for v in _ordered(value)]
  if origin is tuple:
    if len(args) == 2 and args[1] is ...:
      return tuple(_as_type(args[0], v) for v in _ordered(value))
    if len(args) != len(value):
      raise ValueError("tuple length mismatch")
    return tuple(_as_type(t, v) for t, v in zip(args, _ordered(value)))
  try:
    return typ(value)
  except TypeError as e:
    raise TypeError(f"cannot convert checkpoint value to {typ}") from e
