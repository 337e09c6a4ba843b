"""ctypes binding of the C ABI in include/di_b200.h (built from csrc/ into libdi_b200.so).

This is the only door between the Python classes that mirror the reference's API and the
CUDA kernels. There is deliberately no CPU fallback: if the shared library is missing, or no
CUDA device is present, every compute call raises.
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
REPO_DIR = PKG_DIR.parent
import os
# DI_B200_LIB lets a tuning run point at an alternative build of the SAME sources (e.g. other block size)
LIB_PATH = Path(os.environ.get("DI_B200_LIB", PKG_DIR / "libdi_b200.so"))
SOURCES = [PKG_DIR / "csrc" / n for n in
           ("di_b200.cu", "collection.cu", "run_io.cu", "common.cuh", "scan_sort.cuh", "build.cuh", "select.cuh", "score_tile.cuh", "search.cuh")] + [REPO_DIR / "include" / "di_b200.h"]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

OOV = 0xFFFFFFFF
ALL_DOCS = 0xFFFFFFFF
INDEX_NO_SEEDS, INDEX_PER_TILE, INDEX_TILE_BOUNDS = 1, 2, 4

_u8p = ctypes.POINTER(ctypes.c_uint8)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u64p = ctypes.POINTER(ctypes.c_uint64)
_f64p = ctypes.POINTER(ctypes.c_double)
_vp = ctypes.c_void_p


class IndexParams(ctypes.Structure):
    _fields_ = [("tile_docs", ctypes.c_uint32), ("dense_ratio", ctypes.c_uint32),
                ("cand_slack", ctypes.c_uint32), ("flags", ctypes.c_uint32)]


class IndexInfo(ctypes.Structure):
    _fields_ = [("n_postings", ctypes.c_uint64), ("payload_bytes", ctypes.c_uint64),
                ("table_bytes", ctypes.c_uint64), ("n_dense_segments", ctypes.c_uint64),
                ("n_sparse_segments", ctypes.c_uint64), ("n_dense_postings", ctypes.c_uint64),
                ("n_terms", ctypes.c_uint32), ("doc_lo", ctypes.c_uint32), ("doc_hi", ctypes.c_uint32),
                ("n_tiles", ctypes.c_uint32), ("tile_docs", ctypes.c_uint32),
                ("max_docid_plus1", ctypes.c_uint32)]


class Timings(ctypes.Structure):
    _fields_ = [("score_ms", ctypes.c_float), ("finalize_ms", ctypes.c_float), ("total_ms", ctypes.c_float),
                ("score_launches", ctypes.c_uint32), ("other_launches", ctypes.c_uint32),
                ("lanes", ctypes.c_uint32), ("acc32", ctypes.c_uint32), ("tiles_skipped", ctypes.c_uint64)]


# name -> (restype, argtypes); must list every symbol include/di_b200.h declares
SIGNATURES = {
    "di_last_error": (ctypes.c_char_p, []),
    "di_version": (ctypes.c_int, []),
    "di_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "di_set_device": (ctypes.c_int, [ctypes.c_int]),
    "di_find_max_f64": (ctypes.c_int, [_vp, ctypes.c_int64, _f64p]),
    "di_quantize_f64": (ctypes.c_int, [_vp, ctypes.c_int64, ctypes.c_double, _vp]),
    "di_find_max_f64_dev": (ctypes.c_int, [_vp, ctypes.c_int64, _vp, _vp]),
    "di_quantize_f64_dev": (ctypes.c_int, [_vp, ctypes.c_int64, ctypes.c_double, _vp, _vp]),
    "di_invert": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_uint64, ctypes.c_uint32, _vp, _vp, _vp]),
    "di_invert_dev": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, _vp, _vp, _vp, _vp, _vp]),
    "di_index_create_docmajor_dev": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64,
                                                    ctypes.c_uint32, ctypes.POINTER(IndexParams), ctypes.POINTER(_vp)]),
    "di_serialize": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_uint32, _vp, _vp]),
    "di_serialize_dev": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_uint32, ctypes.c_uint64, _vp, _vp, _vp]),
    "di_index_create_csr": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                           ctypes.POINTER(IndexParams), ctypes.POINTER(_vp)]),
    "di_index_create_csr_dev": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint32,
                                               ctypes.c_uint32, ctypes.POINTER(IndexParams), ctypes.POINTER(_vp)]),
    "di_index_create_files": (ctypes.c_int, [_vp, ctypes.c_uint64, _vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                             ctypes.POINTER(IndexParams), ctypes.POINTER(_vp)]),
    "di_index_destroy": (None, [_vp]),
    "di_index_get_info": (ctypes.c_int, [_vp, ctypes.POINTER(IndexInfo)]),
    "di_index_export_seed_hist_dev": (ctypes.c_int, [_vp, _vp, _vp]),
    "di_index_import_seed_hist_dev": (ctypes.c_int, [_vp, _vp, _vp]),
    "di_index_set_sorted_prefix": (ctypes.c_int, [_vp, ctypes.c_uint32]),
    "di_index_term_df": (ctypes.c_int, [_vp, _vp, ctypes.c_uint64, _vp]),
    "di_search": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_uint32, ctypes.c_uint32, _vp, _vp, _vp]),
    "di_search_dev": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, _vp, _vp, _vp, _vp]),
    "di_unpack_keys_dev": (ctypes.c_int, [_vp, ctypes.c_uint64, _vp, _vp, _vp]),
    "di_merge_topk_dev": (ctypes.c_int, [_vp, _vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                         _vp, _vp, _vp, _vp]),
    "di_merge_pull_dev": (ctypes.c_int, [_vp, _vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                         ctypes.c_uint32, ctypes.c_uint32, _vp, _vp, _vp, _vp]),
    "di_shared_alloc": (ctypes.c_int, [ctypes.c_uint64, ctypes.POINTER(_vp), ctypes.c_char_p]),
    "di_shared_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "di_shared_close": (ctypes.c_int, [_vp]),
    "di_shared_free": (ctypes.c_int, [_vp]),
    "di_peer_barrier_dev": (ctypes.c_int, [_vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, _vp]),
    "di_host_alloc": (ctypes.c_int, [ctypes.c_uint64, ctypes.POINTER(_vp)]),
    "di_host_free": (ctypes.c_int, [_vp]),
    "di_run_writer_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "di_run_writer_submit": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_uint32, ctypes.c_uint32]),
    "di_run_writer_close": (ctypes.c_int, [_vp]),
    "di_write_run_file": (ctypes.c_int, [ctypes.c_char_p, _vp, _vp, _vp, _vp, _vp, ctypes.c_uint32, ctypes.c_uint32]),
    "di_eval_ranks_dev": (ctypes.c_int, [_vp, _vp, ctypes.c_uint32, ctypes.c_uint32, _vp, _vp, _vp, ctypes.c_uint32, _vp, _vp, _vp]),
    "di_get_timings": (ctypes.c_int, [_vp, ctypes.POINTER(Timings)]),
    "di_collection_parse": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(_vp)]),
    "di_collection_free": (None, [_vp]),
    "di_collection_info": (ctypes.c_int, [_vp, _u64p, _u64p, _u32p]),
    "di_collection_arrays": (ctypes.c_int, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp),
                                            ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    "di_collection_write_quantized": (ctypes.c_int, [_vp, _vp, ctypes.c_char_p]),
}

_ERRORS = {1: "CUDA", 2: "ARG", 3: "RANGE", 4: "NOMEM", 5: "NODEVICE", 6: "FORMAT", 7: "UNSUPPORTED"}
ERR_FORMAT, ERR_UNSUPPORTED = 6, 7
_lib = None


class NativeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"di_b200 error {code} ({_ERRORS.get(code, '?')}): {message}")
        self.code = code


def build(force: bool = False, verbose: bool = False) -> Path:
    """nvcc-compile csrc/ into libdi_b200.so for sm_100a (cross-compiles without a GPU)."""
    newest = max(p.stat().st_mtime for p in SOURCES)
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < newest:
        cmd = ["nvcc", *NVCC_FLAGS, "-o", str(LIB_PATH), str(PKG_DIR / "csrc" / "di_b200.cu"),
               str(PKG_DIR / "csrc" / "collection.cu"), str(PKG_DIR / "csrc" / "run_io.cu")]
        if verbose:
            print(' '.join(cmd))
        subprocess.run(cmd, check=True)
    return LIB_PATH


def lib():
    """Load libdi_b200.so (never builds implicitly, never falls back)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        L = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise NativeError(rc, lib().di_last_error().decode(errors="replace"))


def device_count() -> int:
    n = ctypes.c_int(0)
    check(lib().di_device_count(ctypes.byref(n)))
    return n.value


def set_device(dev: int):
    check(lib().di_set_device(dev))


def ptr(a) -> int:
    """Address of a numpy array / torch tensor / raw int."""
    if a is None:
        return 0
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()  # torch tensor


def np_c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)
