"""Run the reference's own, unmodified Python package (``peapods``: ``Ising``, ``run_sweep``, the ``peapods`` CLI) on this engine.

The reference's Python layer reaches native code through exactly one name, ``peapods._core.IsingSimulation``
(python/peapods/spin_models.py:3, src/lib.rs:636-640).  ``install()`` registers a module of that name whose class is
``peapods_b200._core.IsingSimulation`` (same constructor, ``sample`` keywords, ``get_spins`` / ``reset``), so with the reference's
``python/`` directory on ``sys.path``

    import peapods_b200.dropin as dropin
    dropin.install("/path/to/peapods/python")   # or have it on PYTHONPATH already
    from peapods import Ising, run_sweep        # the reference's classes, sampling on the GPU
    from peapods.cli import main                # the reference's CLI

works without touching the reference's sources.  (The Rust-side integration, where the crate itself calls the C ABI, is in
INTEGRATION.md.)"""
from __future__ import annotations

import sys
import types


def install(reference_python_dir: str | None = None, backend=None) -> types.ModuleType:
    """Register ``peapods._core`` backed by this engine; optionally put the reference's ``python/`` directory on ``sys.path``.
    Returns the registered module.  Must run before the first ``import peapods``.  ``backend``: another class with the
    ``IsingSimulation`` interface (the test suite wires a CPU stand-in through the same door where there is no GPU)."""
    if backend is None:
        from ._core import IsingSimulation
    else:
        IsingSimulation = backend

    if reference_python_dir is not None and reference_python_dir not in sys.path:
        sys.path.insert(0, reference_python_dir)
    if "peapods" in sys.modules and not isinstance(sys.modules.get("peapods._core"), types.ModuleType):
        raise RuntimeError("peapods was imported before dropin.install(): import order matters")
    core = types.ModuleType("peapods._core")
    core.IsingSimulation = IsingSimulation
    core.__doc__ = "peapods_b200 engine behind the reference's extension-module name"
    sys.modules["peapods._core"] = core
    return core
