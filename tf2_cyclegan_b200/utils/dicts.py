"""Tiny dict helpers (cyclegan/utils.py:101-109)."""
from __future__ import annotations


def append_dict(dict1: dict, dict2: dict, replace: bool = False) -> None:
  """Append items in dict2 to lists in dict1 (or overwrite when ``replace``)."""
  for key, value in dict2.items():
    if replace:
      dict1[key] = value
    else:
      dict1.setdefault(key, []).append(value)
