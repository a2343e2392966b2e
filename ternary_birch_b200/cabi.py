"""ctypes binding of libtbirch.so (C ABI: include/tbirch.h).

Error translation follows the Cython ``except +`` table the reference relies on
(SURVEY.md 8b): invalid_argument / domain_error -> ValueError, overflow_error ->
OverflowError, runtime_error -> RuntimeError, each carrying the reference's
``what()`` text.
"""
from __future__ import annotations

import ctypes
from pathlib import Path

import os

LIB_PATH = Path(os.environ.get("TBIRCH_LIB", Path(__file__).resolve().parent / "libtbirch.so"))

TB_OK, TB_ERR_INVALID_ARGUMENT, TB_ERR_OVERFLOW, TB_ERR_DOMAIN, TB_ERR_RUNTIME = range(5)

_lib = None


class GenusDesc(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int64), ("K", ctypes.c_int64), ("seed", ctypes.c_uint64),
                ("disc", ctypes.c_int64), ("primes", ctypes.c_void_p), ("conductors", ctypes.c_void_p),
                ("dims", ctypes.c_void_p), ("q", ctypes.c_void_p), ("s", ctypes.c_void_p),
                ("sinv", ctypes.c_void_p), ("scalar", ctypes.c_void_p), ("lut", ctypes.c_void_p)]


class PrimeSymbol(ctypes.Structure):
    _fields_ = [("p", ctypes.c_int64), ("power", ctypes.c_int32), ("ramified", ctypes.c_int32)]


def lib() -> ctypes.CDLL:
    """Load libtbirch.so; fail loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                           f"g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
    L = ctypes.CDLL(str(LIB_PATH))
    vp, i64, u64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int
    L.tb_abi_version.restype = i32
    L.tb_last_error.restype = ctypes.c_char_p
    L.tb_launch_count.restype = u64
    L.tb_last_kernel_time.argtypes = [vp, vp, vp, vp]
    L.tb_last_int_mode.argtypes = [vp]
    L.tb_hecke_copy_wait.argtypes = [vp]
    L.tb_hecke_copy_sparse_all.argtypes = [vp, vp, vp, vp]
    L.tb_hecke_transfer_bytes.argtypes = [vp, i64]
    L.tb_hecke_transfer_bytes.restype = i64
    L.tb_host_widen_probe.argtypes = [i64, vp, vp, vp]
    L.tb_readback_stats.argtypes = [vp, vp, vp, vp, vp]
    L.tb_hecke_packed_bytes.argtypes = [vp, vp]
    L.tb_hecke_packed_bytes.restype = i64
    L.tb_hecke_pack.argtypes = [vp, vp, vp]
    L.tb_unpack_rows.argtypes = [vp, i32, i64, vp, vp, vp, vp, vp, vp]
    L.tb_modkernel_csr.argtypes = [vp, vp, vp, i64, i64, ctypes.c_uint32, vp]
    L.tb_modkernel_dense.argtypes = [vp, i64, i64, ctypes.c_uint32, vp]
    L.tb_modkernel_dim.argtypes = [vp]
    L.tb_modkernel_dim.restype = i64
    L.tb_modkernel_cols.argtypes = [vp]
    L.tb_modkernel_cols.restype = i64
    L.tb_modkernel_free_cols.argtypes = [vp, vp]
    L.tb_modkernel_basis.argtypes = [vp, vp]
    L.tb_modkernel_ms.argtypes = [vp]
    L.tb_modkernel_ms.restype = ctypes.c_float
    L.tb_modkernel_destroy.argtypes = [vp]
    L.tb_modkernel_destroy.restype = None
    L.tb_genus_create.argtypes = [ctypes.POINTER(GenusDesc), ctypes.POINTER(vp)]
    L.tb_genus_destroy.argtypes = [vp]
    L.tb_genus_destroy.restype = None
    L.tb_genus_set_stream.argtypes = [vp, vp]
    L.tb_genus_upload.argtypes = [vp, ctypes.POINTER(GenusDesc)]
    for name in ("tb_genus_size", "tb_genus_num_conductors"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = i64
    for name in ("tb_genus_conductor", "tb_genus_dim"):
        getattr(L, name).argtypes = [vp, i64]
        getattr(L, name).restype = i64
    L.tb_hecke_compute.argtypes = [vp, u64, i32, ctypes.POINTER(vp)]
    L.tb_hecke_compute_rows.argtypes = [vp, u64, i32, i64, i64, ctypes.POINTER(vp)]
    L.tb_hecke_num_conductors.argtypes = [vp]
    L.tb_hecke_num_conductors.restype = i64
    for name in ("tb_hecke_conductor", "tb_hecke_dim", "tb_hecke_rows", "tb_hecke_row_begin", "tb_hecke_nnz"):
        getattr(L, name).argtypes = [vp, i64]
        getattr(L, name).restype = i64
    L.tb_hecke_copy_sparse.argtypes = [vp, i64, vp, vp, vp]
    L.tb_hecke_copy_sparse_async.argtypes = [vp, i64, vp, vp, vp, vp]
    L.tb_hecke_copy_dense.argtypes = [vp, i64, vp]
    L.tb_hecke_destroy.argtypes = [vp]
    L.tb_hecke_destroy.restype = None
    L.tb_eigenvalues.argtypes = [vp, u64, i32, i64, vp, vp, vp, vp]
    L.tb_isoseq_open.argtypes = [vp, u64, i32, ctypes.POINTER(vp)]
    L.tb_isoseq_done.argtypes = [vp]
    L.tb_isoseq_next.argtypes = [vp, vp, vp, vp, vp]
    L.tb_isoseq_read.argtypes = [vp, i64, vp, vp]
    L.tb_isoseq_close.argtypes = [vp]
    L.tb_isoseq_close.restype = None
    L.tb_neighbor_records.argtypes = [vp, u64, i32, i64, i64, vp]
    L.tb_int_peak.argtypes = [vp, vp, vp]
    L.tb_int_peak_detail.argtypes = [vp]
    L.tb_quad_form.argtypes = [ctypes.POINTER(PrimeSymbol), i64, vp]
    L.tb_genus_enumerate.argtypes = [vp, ctypes.POINTER(PrimeSymbol), i64, u64, ctypes.POINTER(vp)]
    L.tb_table_desc.argtypes = [vp, ctypes.POINTER(GenusDesc), ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp)]
    L.tb_table_destroy.argtypes = [vp]
    L.tb_table_destroy.restype = None
    if L.tb_abi_version() != 2:
        raise RuntimeError("libtbirch.so ABI version mismatch")
    _lib = L
    return L


EXPORTS = [
    "tb_abi_version", "tb_last_error", "tb_launch_count", "tb_last_kernel_time", "tb_last_int_mode",
    "tb_genus_create", "tb_genus_destroy", "tb_genus_set_stream", "tb_genus_upload", "tb_genus_size",
    "tb_genus_num_conductors", "tb_genus_conductor", "tb_genus_dim",
    "tb_hecke_compute", "tb_hecke_compute_rows", "tb_hecke_num_conductors", "tb_hecke_conductor",
    "tb_hecke_dim", "tb_hecke_rows", "tb_hecke_row_begin", "tb_hecke_nnz", "tb_hecke_copy_sparse",
    "tb_hecke_copy_sparse_async", "tb_hecke_copy_sparse_all", "tb_hecke_copy_wait", "tb_hecke_transfer_bytes", "tb_host_widen_probe", "tb_readback_stats", "tb_hecke_packed_bytes", "tb_hecke_pack", "tb_unpack_rows",
    "tb_hecke_copy_dense", "tb_hecke_destroy", "tb_eigenvalues", "tb_isoseq_open", "tb_isoseq_done",
    "tb_isoseq_next", "tb_isoseq_read", "tb_isoseq_close", "tb_neighbor_records", "tb_int_peak", "tb_int_peak_detail",
    "tb_quad_form", "tb_genus_enumerate", "tb_table_desc", "tb_table_destroy",
    "tb_modkernel_csr", "tb_modkernel_dense", "tb_modkernel_dim", "tb_modkernel_cols", "tb_modkernel_free_cols",
    "tb_modkernel_basis", "tb_modkernel_ms", "tb_modkernel_destroy",
]


def check(status: int) -> None:
    """Raise the Python exception Cython's `except +` would raise for the reference's C++ one."""
    if status == TB_OK:
        return
    msg = lib().tb_last_error().decode()
    if status in (TB_ERR_INVALID_ARGUMENT, TB_ERR_DOMAIN):
        raise ValueError(msg)
    if status == TB_ERR_OVERFLOW:
        raise OverflowError(msg)
    raise RuntimeError(msg)
