"""ternary_birch_b200 -- B200 (sm_100a) implementation of ternary-birch's p-neighbor
Hecke hot path behind the reference's Genus / hecke_matrix / isometry_sequence API.

The compute lives in ``libtbirch.so`` (hand-written CUDA, C ABI in
``include/tbirch.h``); this package is the Python host side: a ctypes binding
(``cabi``), the genus-table container (``table``) and twins of the reference's
``Genus<R>`` / ``IsometrySequence`` C++ classes and ``BirchGenus`` Python class
(``genus``).  There is no CPU fallback: importing works anywhere, computing
needs the CUDA library and a GPU.
"""
from .table import (GenusTable, build_genus_table, load_genus_table, prime_symbols, quad_form,  # noqa: F401
                    save_genus_table)
from .genus import Genus, IsometrySequence, BirchGenus, EigenvectorManager  # noqa: F401
from . import cabi  # noqa: F401

__all__ = ["GenusTable", "build_genus_table", "quad_form", "prime_symbols", "load_genus_table", "save_genus_table", "Genus", "IsometrySequence",
           "BirchGenus", "EigenvectorManager", "cabi"]
