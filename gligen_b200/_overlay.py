"""Overlaying this repo's drop-in modules on a reference checkout.

The documented integration is a PYTHONPATH overlay: this repo FIRST, the reference checkout after
it.  Two things make that work although both trees hold packages called `ldm` / `grounding_input`:

  * `extend(__path__, __name__)` in every drop-in package __init__: the package's search path also covers the
    same-named directory of every later sys.path entry, so modules this repo does not provide
    (ldm.models.autoencoder, ldm.modules.diffusionmodules.model, ldm.modules.encoders, ...) still import from the
    reference (the reference's `ldm` is a namespace package; a regular package would otherwise shadow it completely);
  * `fallback(__name__, __file__)` as the module-level __getattr__ of every drop-in MODULE: a name the drop-in does
    not define (e.g. `checkpoint`, `conv_nd` of ldm.modules.diffusionmodules.util, used by reference modules outside
    the hot path) resolves to the module of the same dotted name in the reference portion, loaded once under a
    private alias.  Without a reference checkout on the path the lookup fails with the usual AttributeError.

Reference: the imports at gligen_inference.py:6-12 and the dotted class names stored in checkpoints
(`config_dict`, gligen_inference.py:70-86, resolved by ldm/util.py:71-86).
"""
from __future__ import annotations

import importlib.util
import os
import pkgutil
import sys

_shadow_cache = {}


def extend(path, name):
    return pkgutil.extend_path(path, name)


def _shadowed_module(module_name: str, module_file: str):
    if module_name in _shadow_cache:
        return _shadow_cache[module_name]
    found = None
    parent_name, _, leaf = module_name.rpartition(".")
    search = sys.modules[parent_name].__path__ if parent_name else sys.path
    here = os.path.dirname(os.path.abspath(module_file))
    for d in search:
        if not isinstance(d, str) or os.path.abspath(d) == here:
            continue
        cand = os.path.join(d, leaf + ".py")
        if os.path.isfile(cand):
            spec = importlib.util.spec_from_file_location("_gligen_b200_shadowed." + module_name, cand)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[spec.name] = mod
            _shadow_cache[module_name] = mod           # before exec: a cycle sees the partially initialised module
            spec.loader.exec_module(mod)
            found = mod
            break
    _shadow_cache[module_name] = found
    return found


def fallback(module_name: str, module_file: str):
    def __getattr__(name: str):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        ref = _shadowed_module(module_name, module_file)
        if ref is not None and hasattr(ref, name):
            return getattr(ref, name)
        raise AttributeError(f"module {module_name!r} (gligen_b200 drop-in) has no attribute {name!r}"
                             + ("" if ref is not None else " and no reference checkout follows it on sys.path"))
    return __getattr__
