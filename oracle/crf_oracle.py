"""ctypes binding of oracle/crf_oracle.c plus a tiny numpy twin.

TEST INFRASTRUCTURE — see the header of ``crf_oracle.c``.  The numpy twin (`chain_marginals_numpy`)
is an independent second restatement of SURVEY.md Appendix B used to cross-check the C code on small
cases; the C library is what the parity tests and the CPU baseline run.
"""

from __future__ import annotations

import ctypes
import os
import pathlib
import subprocess
from typing import Optional, Tuple

import numpy

_HERE = pathlib.Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libcrf_oracle.so"
_lib: Optional[ctypes.CDLL] = None


def build(force: bool = False) -> pathlib.Path:
    """Compile the oracle with the committed Makefile (gcc only)."""
    src = _HERE / "crf_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s", "-B"], check=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(_LIB_PATH))
        dp = ctypes.POINTER(ctypes.c_double)
        i64p = ctypes.POINTER(ctypes.c_int64)
        i32p = ctypes.POINTER(ctypes.c_int32)
        L.oracle_chain_marginals.restype = ctypes.c_int
        L.oracle_chain_marginals.argtypes = [dp, ctypes.c_int32, ctypes.c_int32, dp, i64p, i32p,
                                             ctypes.c_int64, ctypes.c_int64, dp]
        L.oracle_marginals_windowed.restype = ctypes.c_int
        L.oracle_marginals_windowed.argtypes = [dp, ctypes.c_int32, ctypes.c_int32, dp, ctypes.c_int32,
                                                i64p, i64p, i32p, ctypes.c_int64, ctypes.c_int32,
                                                ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, dp, i64p]
        _lib = L
    return _lib


def _ptr(a: numpy.ndarray, ctype):
    return a.ctypes.data_as(ctypes.POINTER(ctype))


def _prep(state_w, trans_w, gene_ptr, attr_idx):
    state_w = numpy.ascontiguousarray(state_w, dtype=numpy.float64)
    trans_w = numpy.ascontiguousarray(trans_w, dtype=numpy.float64)
    gene_ptr = numpy.ascontiguousarray(gene_ptr, dtype=numpy.int64)
    attr_idx = numpy.ascontiguousarray(attr_idx, dtype=numpy.int32)
    if attr_idx.size == 0:
        attr_idx = numpy.zeros(1, dtype=numpy.int32)
    return state_w, trans_w, gene_ptr, attr_idx


def chain_marginals(state_w, trans_w, gene_ptr, attr_idx, row_begin=0, row_end=None) -> numpy.ndarray:
    """Marginals [T][L] of one sequence = CSR rows [row_begin, row_end) (predict_marginals_single)."""
    state_w, trans_w, gene_ptr, attr_idx = _prep(state_w, trans_w, gene_ptr, attr_idx)
    if row_end is None:
        row_end = len(gene_ptr) - 1
    A, L = state_w.shape
    out = numpy.empty((row_end - row_begin, L), dtype=numpy.float64)
    rc = lib().oracle_chain_marginals(_ptr(state_w, ctypes.c_double), A, L, _ptr(trans_w, ctypes.c_double),
                                      _ptr(gene_ptr, ctypes.c_int64), _ptr(attr_idx, ctypes.c_int32),
                                      row_begin, row_end, _ptr(out, ctypes.c_double))
    if rc != 0:
        raise RuntimeError(f"oracle_chain_marginals failed ({rc})")
    return out


def marginals_windowed(state_w, trans_w, pos_label, contig_ptr, gene_ptr, attr_idx, window, step=1,
                       pad=True, nthreads=1) -> Tuple[numpy.ndarray, int]:
    """Per-gene P(pos_label) through the reference's window/pad/max loop; NaN = skipped contig."""
    state_w, trans_w, gene_ptr, attr_idx = _prep(state_w, trans_w, gene_ptr, attr_idx)
    contig_ptr = numpy.ascontiguousarray(contig_ptr, dtype=numpy.int64)
    A, L = state_w.shape
    G = len(gene_ptr) - 1
    out = numpy.zeros(max(G, 1), dtype=numpy.float64)
    windows = ctypes.c_int64(0)
    rc = lib().oracle_marginals_windowed(_ptr(state_w, ctypes.c_double), A, L, _ptr(trans_w, ctypes.c_double),
                                         int(pos_label), _ptr(contig_ptr, ctypes.c_int64),
                                         _ptr(gene_ptr, ctypes.c_int64), _ptr(attr_idx, ctypes.c_int32),
                                         len(contig_ptr) - 1, int(window), int(step), int(bool(pad)),
                                         int(nthreads), _ptr(out, ctypes.c_double), ctypes.byref(windows))
    if rc != 0:
        raise ValueError(f"oracle_marginals_windowed failed ({rc})")
    return out[:G], int(windows.value)


# ---------------------------------------------------------------------------------------------
# numpy twin (small cases only)
# ---------------------------------------------------------------------------------------------


def chain_marginals_numpy(state_w, trans_w, items) -> numpy.ndarray:
    """SURVEY.md Appendix B, written independently of the C file.  ``items`` = list of id lists."""
    state_w = numpy.asarray(state_w, dtype=numpy.float64)
    M = numpy.exp(numpy.asarray(trans_w, dtype=numpy.float64))
    A, L = state_w.shape
    T = len(items)
    E = numpy.empty((T, L))
    for t, item in enumerate(items):
        s = numpy.zeros(L)
        for a in item:
            if 0 <= a < A:
                s = s + state_w[a]
        E[t] = numpy.exp(s)
    alpha = numpy.empty((T, L)); beta = numpy.empty((T, L)); c = numpy.empty(T)
    for t in range(T):
        a = E[0].copy() if t == 0 else (alpha[t - 1] @ M) * E[t]
        tot = a.sum()
        c[t] = 1.0 / tot if tot != 0 else 1.0
        alpha[t] = a * c[t]
    beta[T - 1] = c[T - 1]
    for t in range(T - 2, -1, -1):
        beta[t] = c[t] * (M @ (E[t + 1] * beta[t + 1]))
    return alpha * beta / c[:, None]


def marginals_windowed_numpy(state_w, trans_w, pos_label, contig_ptr, gene_ptr, attr_idx, window,
                             step=1, pad=True) -> numpy.ndarray:
    """gecco/crf/__init__.py:209-258 with lists and the numpy twin above."""
    G = len(gene_ptr) - 1
    out = numpy.zeros(G)
    for c in range(len(contig_ptr) - 1):
        g0, g1 = int(contig_ptr[c]), int(contig_ptr[c + 1])
        feats = [list(attr_idx[gene_ptr[g]:gene_ptr[g + 1]]) for g in range(g0, g1)]
        n = len(feats)
        delta = 0
        if n < window:
            if not pad:
                out[g0:g1] = numpy.nan
                continue
            delta = window - n
            feats = [[] for _ in range(delta // 2)] + feats + [[] for _ in range((delta + 1) // 2)]
        prob = numpy.zeros(max(n, window))
        for i in range(0, len(feats) + 1 - window, step):
            m = chain_marginals_numpy(state_w, trans_w, feats[i:i + window])[:, pos_label]
            numpy.maximum(prob[i:i + window], m, out=prob[i:i + window])
        out[g0:g1] = prob[delta // 2:][:n]
    return out
