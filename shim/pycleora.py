"""Replacement for the reference's native extension module ``pycleora.pycleora`` (``src/lib.rs:490-495``, imported at
``pycleora/__init__.py:4`` as ``from .pycleora import SparseMatrix``).

Install: put this file into the reference's ``pycleora/`` package directory in place of the compiled
``pycleora.*.so`` (with ``cleora_b200`` importable).  Every unmodified reference ``.py`` file -- ``embed()``, the CLI
(``cli.py:137-157``), the benchmark helpers, ``hetero.py`` -- then runs on top of the CUDA path: ``SparseMatrix`` here is
the ctypes mirror of the pyo3 class bound to ``libcleora_b200.so`` (same names, defaults and exception types; see
``cleora_b200/pycleora.py``).

Note: on this import path the reference's pure-Python ``embed()`` drives the loop one iteration at a time through
``left_markov_propagate`` (two n x d host copies per iteration, exactly as with the Rust module).  The device-resident
whole-loop call is ``cleora_b200.embed()`` -- same signature -- or, to get it without touching call sites, rebind
``pycleora.embed = cleora_b200.embed`` after import.
"""
from cleora_b200.pycleora import SparseMatrix

__all__ = ["SparseMatrix"]
