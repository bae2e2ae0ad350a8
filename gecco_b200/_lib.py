"""ctypes binding of ``libgecco_crf_b200.so`` (C ABI: ``include/gecco_crf_b200.h``).

There is deliberately no fallback: if the shared library is missing or no B200 is visible, the calls
raise.  The CPU oracle under ``oracle/`` is test infrastructure and is never imported from here.
"""

from __future__ import annotations

import ctypes
import os
import pathlib
from typing import List, Optional, Tuple

import numpy

from .model_io import CRFWeights

__all__ = ["CRFEngine", "GcrfError", "WireBatch", "load_library", "library_path", "EXPORTED_SYMBOLS"]

GCRF_FLAG_DEVICE_PTRS = 0x1
GCRF_FLAG_OUT_F32 = 0x2
GCRF_FLAG_PTR64 = 0x4
GCRF_FLAG_PROB_F32 = 0x8
GCRF_FLAG_RESET_PER_CONTIG = 0x10
GCRF_FLAG_IDX_U16 = 0x20
GCRF_FLAG_ACCESSIONS = 0x40
GCRF_FLAG_F64 = 0x80
GCRF_FLAG_MULTICAST = 0x100

# every symbol include/gecco_crf_b200.h declares (tests/test_abi.py checks the header against this)
EXPORTED_SYMBOLS = (
    "gcrf_version",
    "gcrf_last_error",
    "gcrf_device_count",
    "gcrf_model_create",
    "gcrf_model_destroy",
    "gcrf_model_set_stream",
    "gcrf_model_synchronize",
    "gcrf_marginals_windowed",
    "gcrf_marginals_windowed_peers",
    "gcrf_marginals_chain",
    "gcrf_model_set_vocabulary",
    "gcrf_features_from_accessions",
    "gcrf_segments",
    "gcrf_wire_last_error",
    "gcrf_wire_encode",
    "gcrf_wire_destroy",
    "gcrf_wire_bytes",
    "gcrf_wire_contigs",
    "gcrf_wire_genes",
    "gcrf_wire_ids",
    "gcrf_wire_decode_host",
    "gcrf_marginals_windowed_wire",
    "gcrf_host_alloc",
    "gcrf_host_free",
    "gcrf_max_window",
    "gcrf_model_launch_count",
    "gcrf_model_set_timing",
    "gcrf_model_last_kernel_ms",
    "gcrf_table_last_error",
    "gcrf_table_load",
    "gcrf_table_parse",
    "gcrf_table_destroy",
    "gcrf_table_contigs",
    "gcrf_table_genes",
    "gcrf_table_domains",
    "gcrf_table_contig_id",
    "gcrf_table_gene_id",
    "gcrf_table_contig_ptr",
    "gcrf_table_annotated",
    "gcrf_table_gene_coordinates",
    "gcrf_table_pack",
    "gcrf_table_pack_accessions",
    "gcrf_table_row_gene",
    "gcrf_table_gene_probabilities",
    "gcrf_table_write_genes",
    "gcrf_table_write_features",
    "gcrf_table_write_clusters",
)

_STATUS = {0: "GCRF_OK", -1: "GCRF_EINVAL", -2: "GCRF_ENODEVICE", -3: "GCRF_ECUDA", -4: "GCRF_ENOMEM",
           -5: "GCRF_EUNSUPPORTED"}


class GcrfError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{_STATUS.get(status, status)}: {message}")
        self.status = status


def library_path() -> pathlib.Path:
    # GCRF_TUNING_LIB=1 selects the build with phase timers / ablation switches (tools/quick_kernel_time.py)
    name = "libgecco_crf_b200_tuning.so" if os.environ.get("GCRF_TUNING_LIB") == "1" else "libgecco_crf_b200.so"
    name = os.environ.get("GCRF_LIB_NAME", name)  # A/B builds of single kernels (tools/ab_libs.sh)
    return pathlib.Path(__file__).resolve().parent / name


_lib: Optional[ctypes.CDLL] = None


def load_library() -> ctypes.CDLL:
    """Load the CUDA library; raises if it was not built (``python -c "import __graft_entry__ as g; g.build()"``)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.exists():
        raise RuntimeError(
            f"{path} is missing: build it with `make -C gecco_b200/csrc` (or __graft_entry__.build()); "
            "gecco_b200 has no CPU fallback"
        )
    lib = ctypes.CDLL(str(path))
    vp, i32, i64, u32, u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint64
    lib.gcrf_version.restype = ctypes.c_int
    lib.gcrf_version.argtypes = []
    lib.gcrf_last_error.restype = ctypes.c_char_p
    lib.gcrf_last_error.argtypes = []
    lib.gcrf_device_count.restype = ctypes.c_int
    lib.gcrf_device_count.argtypes = []
    lib.gcrf_model_create.restype = ctypes.c_int
    lib.gcrf_model_create.argtypes = [vp, i32, i32, vp, i32, i32, ctypes.POINTER(vp)]
    lib.gcrf_model_destroy.restype = None
    lib.gcrf_model_destroy.argtypes = [vp]
    lib.gcrf_model_set_stream.restype = ctypes.c_int
    lib.gcrf_model_set_stream.argtypes = [vp, vp]
    lib.gcrf_model_synchronize.restype = ctypes.c_int
    lib.gcrf_model_synchronize.argtypes = [vp]
    lib.gcrf_marginals_windowed.restype = ctypes.c_int
    lib.gcrf_marginals_windowed.argtypes = [vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, vp, u32]
    lib.gcrf_marginals_windowed_peers.restype = ctypes.c_int
    lib.gcrf_marginals_windowed_peers.argtypes = [vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, vp, vp, i32, i64, u32]
    lib.gcrf_marginals_chain.restype = ctypes.c_int
    lib.gcrf_marginals_chain.argtypes = [vp, vp, vp, vp, i64, i64, i64, vp, u32]
    lib.gcrf_model_set_vocabulary.restype = ctypes.c_int
    lib.gcrf_model_set_vocabulary.argtypes = [vp, vp, i32]
    lib.gcrf_features_from_accessions.restype = ctypes.c_int
    lib.gcrf_features_from_accessions.argtypes = [vp, vp, vp, i64, i64, vp, u32]
    lib.gcrf_segments.restype = ctypes.c_int
    lib.gcrf_segments.argtypes = [vp, vp, vp, vp, i64, i64, ctypes.c_double, i32, i32, i32, vp, vp, vp, vp, vp, vp, i64,
                                  ctypes.POINTER(i64), u32]
    lib.gcrf_wire_last_error.restype = ctypes.c_char_p
    lib.gcrf_wire_last_error.argtypes = []
    lib.gcrf_wire_encode.restype = ctypes.c_int
    lib.gcrf_wire_encode.argtypes = [vp, vp, vp, i64, i64, i64, i32, u32, ctypes.POINTER(vp)]
    lib.gcrf_wire_destroy.restype = None
    lib.gcrf_wire_destroy.argtypes = [vp]
    for name in ("gcrf_wire_bytes", "gcrf_wire_contigs", "gcrf_wire_genes", "gcrf_wire_ids"):
        getattr(lib, name).restype = i64
        getattr(lib, name).argtypes = [vp]
    lib.gcrf_wire_decode_host.restype = ctypes.c_int
    lib.gcrf_wire_decode_host.argtypes = [vp, vp, vp]
    lib.gcrf_marginals_windowed_wire.restype = ctypes.c_int
    lib.gcrf_marginals_windowed_wire.argtypes = [vp, vp, i32, i32, i32, vp, u32]
    lib.gcrf_host_alloc.restype = ctypes.c_int
    lib.gcrf_host_alloc.argtypes = [ctypes.POINTER(vp), u64]
    lib.gcrf_host_free.restype = ctypes.c_int
    lib.gcrf_host_free.argtypes = [vp]
    lib.gcrf_max_window.restype = i32
    lib.gcrf_max_window.argtypes = [vp, i32]
    lib.gcrf_model_launch_count.restype = i64
    lib.gcrf_model_launch_count.argtypes = [vp]
    lib.gcrf_model_set_timing.restype = ctypes.c_int
    lib.gcrf_model_set_timing.argtypes = [vp, i32]
    lib.gcrf_model_last_kernel_ms.restype = ctypes.c_double
    lib.gcrf_model_last_kernel_ms.argtypes = [vp]
    # tables (host code)
    cp, dbl = ctypes.c_char_p, ctypes.c_double
    lib.gcrf_table_last_error.restype = cp
    lib.gcrf_table_last_error.argtypes = []
    lib.gcrf_table_load.restype = ctypes.c_int
    lib.gcrf_table_load.argtypes = [cp, ctypes.POINTER(cp), i32, dbl, dbl, ctypes.POINTER(vp)]
    lib.gcrf_table_parse.restype = ctypes.c_int
    lib.gcrf_table_parse.argtypes = [cp, u64, ctypes.POINTER(cp), ctypes.POINTER(u64), i32, dbl, dbl, ctypes.POINTER(vp)]
    lib.gcrf_table_destroy.restype = None
    lib.gcrf_table_destroy.argtypes = [vp]
    for name in ("gcrf_table_contigs", "gcrf_table_genes", "gcrf_table_domains"):
        getattr(lib, name).restype = i64
        getattr(lib, name).argtypes = [vp]
    lib.gcrf_table_contig_id.restype = cp
    lib.gcrf_table_contig_id.argtypes = [vp, i64]
    lib.gcrf_table_gene_id.restype = cp
    lib.gcrf_table_gene_id.argtypes = [vp, i64]
    lib.gcrf_table_contig_ptr.restype = vp
    lib.gcrf_table_contig_ptr.argtypes = [vp]
    lib.gcrf_table_annotated.restype = vp
    lib.gcrf_table_annotated.argtypes = [vp]
    lib.gcrf_table_gene_coordinates.restype = ctypes.c_int
    lib.gcrf_table_gene_coordinates.argtypes = [vp, vp, vp]
    lib.gcrf_table_pack.restype = ctypes.c_int
    lib.gcrf_table_pack.argtypes = [vp, ctypes.POINTER(cp), i32, i32, ctypes.POINTER(vp), ctypes.POINTER(vp),
                                    ctypes.POINTER(vp), ctypes.POINTER(i64), ctypes.POINTER(i64)]
    lib.gcrf_table_pack_accessions.restype = ctypes.c_int
    lib.gcrf_table_pack_accessions.argtypes = [vp, i32, i32, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp),
                                               ctypes.POINTER(i64), ctypes.POINTER(i64)]
    lib.gcrf_table_row_gene.restype = vp
    lib.gcrf_table_row_gene.argtypes = [vp]
    lib.gcrf_table_gene_probabilities.restype = ctypes.c_int
    lib.gcrf_table_gene_probabilities.argtypes = [vp, vp, vp, vp]
    lib.gcrf_table_write_genes.restype = ctypes.c_int
    lib.gcrf_table_write_genes.argtypes = [vp, vp, cp]
    lib.gcrf_table_write_features.restype = ctypes.c_int
    lib.gcrf_table_write_features.argtypes = [vp, vp, cp]
    lib.gcrf_table_write_clusters.restype = ctypes.c_int
    lib.gcrf_table_write_clusters.argtypes = [vp, vp, vp, vp, vp, vp, i64, cp]
    _lib = lib
    return lib


def _check(lib: ctypes.CDLL, status: int) -> None:
    if status != 0:
        raise GcrfError(status, lib.gcrf_last_error().decode("utf-8", "replace"))


class PinnedArray:
    """A numpy view over page-locked host memory from ``gcrf_host_alloc``."""

    def __init__(self, shape, dtype):
        self._lib = load_library()
        self.dtype = numpy.dtype(dtype)
        n = int(numpy.prod(shape))
        ptr = ctypes.c_void_p()
        _check(self._lib, self._lib.gcrf_host_alloc(ctypes.byref(ptr), max(1, n * self.dtype.itemsize)))
        self._ptr = ptr
        buf = (ctypes.c_char * max(1, n * self.dtype.itemsize)).from_address(ptr.value)
        self.array = numpy.frombuffer(buf, dtype=self.dtype, count=n).reshape(shape)

    def free(self) -> None:
        if self._ptr is not None:
            self.array = None
            self._lib.gcrf_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class WireBatch:
    """A CSR batch in the compact wire format (``gcrf_wire_encode``): one page-locked block holding ``contig_ptr``, per-gene
    id / byte counts and the genes' sorted attribute ids as Rice-coded deltas — what a host-buffer call has to move over
    PCIe, at ~1.06 bytes per id instead of 4.  Encoding is host code (no GPU needed)."""

    def __init__(self, contig_ptr, gene_ptr, attr_idx, num_attrs: int):
        self._lib = load_library()
        contig_ptr = numpy.ascontiguousarray(contig_ptr, dtype=numpy.int32)
        gene_ptr = numpy.asarray(gene_ptr)
        flags = 0
        if gene_ptr.dtype == numpy.int64:
            gene_ptr = numpy.ascontiguousarray(gene_ptr)
            flags = GCRF_FLAG_PTR64
        else:
            gene_ptr = numpy.ascontiguousarray(gene_ptr, dtype=numpy.int32)
        attr_idx = numpy.ascontiguousarray(attr_idx, dtype=numpy.int32)
        handle = ctypes.c_void_p()
        rc = self._lib.gcrf_wire_encode(contig_ptr.ctypes.data, gene_ptr.ctypes.data, attr_idx.ctypes.data if len(attr_idx) else None,
                                        len(contig_ptr) - 1, len(gene_ptr) - 1, len(attr_idx), int(num_attrs), flags,
                                        ctypes.byref(handle))
        if rc != 0:
            raise GcrfError(rc, self._lib.gcrf_wire_last_error().decode("utf-8", "replace"))
        self._h = handle

    @property
    def nbytes(self) -> int:
        """Size of the block = host-to-device bytes of one call."""
        return int(self._lib.gcrf_wire_bytes(self._h))

    @property
    def C(self) -> int:
        return int(self._lib.gcrf_wire_contigs(self._h))

    @property
    def G(self) -> int:
        return int(self._lib.gcrf_wire_genes(self._h))

    @property
    def nnz(self) -> int:
        return int(self._lib.gcrf_wire_ids(self._h))

    def decode(self):
        """``(gene_ptr, attr_idx)`` of the batch the device will see: ids sorted per gene, unknown ids = ``num_attrs``."""
        gene_ptr = numpy.zeros(self.G + 1, dtype=numpy.int32)
        attr_idx = numpy.zeros(self.nnz, dtype=numpy.int32)
        rc = self._lib.gcrf_wire_decode_host(self._h, gene_ptr.ctypes.data, attr_idx.ctypes.data if self.nnz else None)
        if rc != 0:
            raise GcrfError(rc, self._lib.gcrf_wire_last_error().decode("utf-8", "replace"))
        return gene_ptr, attr_idx

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            self._lib.gcrf_wire_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Segments:
    """Clusters found by ``gcrf_segments``: parallel arrays, one entry per cluster, in the reference's order."""

    def __init__(self, capacity: int):
        n = max(1, int(capacity))
        self.contig = numpy.zeros(n, dtype=numpy.int32)
        self.begin = numpy.zeros(n, dtype=numpy.int32)
        self.end = numpy.zeros(n, dtype=numpy.int32)
        self.ordinal = numpy.zeros(n, dtype=numpy.int32)
        self.average_p = numpy.zeros(n, dtype=numpy.float64)
        self.max_p = numpy.zeros(n, dtype=numpy.float64)

    def truncated(self, n: int) -> "Segments":
        for name in ("contig", "begin", "end", "ordinal", "average_p", "max_p"):
            setattr(self, name, getattr(self, name)[:n])
        return self

    def __len__(self) -> int:
        return len(self.contig)


class CRFEngine:
    """One model handle on one B200 (``gcrf_model``).  Not thread-safe; use one engine per thread/device."""

    def __init__(self, weights: CRFWeights, device: int = 0, positive_label: str = "1"):
        self._lib = load_library()
        self.weights = weights
        self.device = int(device)
        state_w = numpy.ascontiguousarray(weights.state_w, dtype=numpy.float64)
        trans_w = numpy.ascontiguousarray(weights.trans_w, dtype=numpy.float64)
        A, L = state_w.shape
        # gecco/crf/__init__.py:253 asks the tagger for p['1']: resolve the label id by name
        pos = weights.label_id(positive_label)
        handle = ctypes.c_void_p()
        _check(self._lib, self._lib.gcrf_model_create(state_w.ctypes.data, A, L, trans_w.ctypes.data, pos,
                                                     self.device, ctypes.byref(handle)))
        self._handle = handle
        self.has_vocabulary = False
        self.vocabulary_digits = 0
        # Pfam-style attribute names ("PF" + a fixed number of ASCII digits, so that number <-> name is one to one):
        # hand the integer vocabulary to the device so that feature extraction (accession -> attribute id, repeats
        # inside a gene dropped) can run there
        accs = []
        widths = set()
        for name in weights.attrs:
            if len(name) > 2 and name[:2] == "PF" and name[2:].isascii() and name[2:].isdigit() and len(name) <= 11:
                accs.append(int(name[2:]))
                widths.add(len(name) - 2)
            else:
                accs = None
                break
        if accs and len(widths) == 1 and len(set(accs)) == len(accs):
            self.vocabulary_digits = widths.pop()
            arr = numpy.asarray(accs, dtype=numpy.int32)
            _check(self._lib, self._lib.gcrf_model_set_vocabulary(self._handle, arr.ctypes.data if len(arr) else None, len(arr)))
            self.has_vocabulary = True

    # ------------------------------------------------------------------ lifetime
    def close(self) -> None:
        if getattr(self, "_handle", None) is not None:
            self._lib.gcrf_model_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _host_csr(contig_ptr, gene_ptr, attr_idx):
        contig_ptr = numpy.ascontiguousarray(contig_ptr, dtype=numpy.int32)
        gene_ptr = numpy.asarray(gene_ptr)
        if gene_ptr.dtype == numpy.int64 and (gene_ptr.size == 0 or int(gene_ptr[-1]) > 0x7FFFFFFF):
            gene_ptr = numpy.ascontiguousarray(gene_ptr, dtype=numpy.int64)
            flags = GCRF_FLAG_PTR64
        else:
            gene_ptr = numpy.ascontiguousarray(gene_ptr, dtype=numpy.int32)
            flags = 0
        if isinstance(attr_idx, numpy.ndarray) and attr_idx.dtype == numpy.uint16:
            # compact ids (0xFFFF = unknown attribute): half the bytes over PCIe, widened on the device
            attr_idx = numpy.ascontiguousarray(attr_idx)
            flags |= GCRF_FLAG_IDX_U16
        else:
            attr_idx = numpy.ascontiguousarray(attr_idx, dtype=numpy.int32)
        return contig_ptr, gene_ptr, attr_idx, flags

    # ------------------------------------------------------------------ host-pointer calls
    def marginals_windowed(self, contig_ptr, gene_ptr, attr_idx, *, window: Optional[int] = None,
                           step: Optional[int] = None, pad: bool = True, out: Optional[numpy.ndarray] = None,
                           f32: bool = False, accessions: bool = False, f64_arith: bool = False) -> numpy.ndarray:
        """Per-gene cluster probability (``gcrf_marginals_windowed``, host buffers, blocking).  ``accessions``:
        ``attr_idx`` holds integer domain accessions, one row per domain in domain-start order; they are mapped to
        attribute ids and de-duplicated per gene on the device (``GCRF_FLAG_ACCESSIONS``).  ``f64_arith``: compute in
        the reference's own f64 arithmetic (``GCRF_FLAG_F64``) instead of FP32."""
        contig_ptr, gene_ptr, attr_idx, flags = self._host_csr(contig_ptr, gene_ptr, attr_idx)
        if f64_arith:
            flags |= GCRF_FLAG_F64
        if accessions:
            if not self.has_vocabulary:
                raise ValueError("the model's attributes are not Pfam-style accessions; pack attribute ids on the host")
            flags |= GCRF_FLAG_ACCESSIONS
        C, G, nnz = len(contig_ptr) - 1, len(gene_ptr) - 1, len(attr_idx)
        dtype = numpy.float32 if f32 else numpy.float64
        if out is None:
            out = numpy.empty(G, dtype=dtype)
        elif out.dtype != dtype or out.size != G or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous array of G elements of the requested dtype")
        if f32:
            flags |= GCRF_FLAG_OUT_F32
        window = self.weights.window_size if window is None else window
        step = self.weights.window_step if step is None else step
        _check(self._lib, self._lib.gcrf_marginals_windowed(
            self._handle, contig_ptr.ctypes.data, gene_ptr.ctypes.data, attr_idx.ctypes.data if nnz else None,
            C, G, nnz, int(window), int(step), int(bool(pad)), out.ctypes.data if G else None, flags))
        return out

    def marginals_windowed_wire(self, wire: "WireBatch", *, window: Optional[int] = None, step: Optional[int] = None,
                                pad: bool = True, out: Optional[numpy.ndarray] = None, f32: bool = False,
                                f64_arith: bool = False) -> numpy.ndarray:
        """``gcrf_marginals_windowed_wire``: one H2D copy of the compact block, decode + marginals on the device, D2H."""
        dtype = numpy.float32 if f32 else numpy.float64
        G = wire.G
        if out is None:
            out = numpy.empty(G, dtype=dtype)
        elif out.dtype != dtype or out.size != G or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous array of G elements of the requested dtype")
        flags = (GCRF_FLAG_OUT_F32 if f32 else 0) | (GCRF_FLAG_F64 if f64_arith else 0)
        window = self.weights.window_size if window is None else window
        step = self.weights.window_step if step is None else step
        _check(self._lib, self._lib.gcrf_marginals_windowed_wire(self._handle, wire._h, int(window), int(step), int(bool(pad)),
                                                                 out.ctypes.data if G else None, flags))
        return out

    def marginals_chain(self, contig_ptr, gene_ptr, attr_idx, *, out: Optional[numpy.ndarray] = None,
                        f32: bool = False, accessions: bool = False) -> numpy.ndarray:
        """Whole-contig marginals (``gcrf_marginals_chain``): one chain per contig, no windows."""
        contig_ptr, gene_ptr, attr_idx, flags = self._host_csr(contig_ptr, gene_ptr, attr_idx)
        if accessions:
            flags |= GCRF_FLAG_ACCESSIONS
        C, G, nnz = len(contig_ptr) - 1, len(gene_ptr) - 1, len(attr_idx)
        dtype = numpy.float32 if f32 else numpy.float64
        if out is None:
            out = numpy.empty(G, dtype=dtype)
        if f32:
            flags |= GCRF_FLAG_OUT_F32
        _check(self._lib, self._lib.gcrf_marginals_chain(
            self._handle, contig_ptr.ctypes.data, gene_ptr.ctypes.data, attr_idx.ctypes.data if nnz else None,
            C, G, nnz, out.ctypes.data if G else None, flags))
        return out

    def features_from_accessions(self, accession, gene_ptr, *, ptr64: bool = False) -> numpy.ndarray:
        """Integer domain accessions (row by row, domain-start order) -> attribute ids, -1 for unknown accessions
        and for repeats inside a gene (``gcrf_features_from_accessions``, host buffers, blocking).  ``ptr64``
        passes 64-bit row pointers even when 32 bits would do."""
        accession = numpy.ascontiguousarray(accession, dtype=numpy.int32)
        gene_ptr = numpy.asarray(gene_ptr)
        flags = 0
        if ptr64 or (gene_ptr.dtype == numpy.int64 and gene_ptr.size and int(gene_ptr[-1]) > 0x7FFFFFFF):
            gene_ptr = numpy.ascontiguousarray(gene_ptr, dtype=numpy.int64)
            flags = GCRF_FLAG_PTR64
        else:
            gene_ptr = numpy.ascontiguousarray(gene_ptr, dtype=numpy.int32)
        out = numpy.empty(len(accession), dtype=numpy.int32)
        _check(self._lib, self._lib.gcrf_features_from_accessions(
            self._handle, accession.ctypes.data if len(accession) else None, gene_ptr.ctypes.data,
            len(gene_ptr) - 1, len(accession), out.ctypes.data if len(out) else None, flags))
        return out

    def segments(self, contig_ptr, prob, annotated, *, threshold: float = 0.8, n_cds: int = 5, edge_distance: int = 0,
                 trim: bool = True, reset_per_contig: bool = False, capacity: Optional[int] = None) -> "Segments":
        """Threshold + contiguous-segment extraction (``gcrf_segments``; ``gecco/refine.py:120-200``, criterion
        "gecco") on per-gene probabilities (NaN = no probability) and annotation marks; host buffers, blocking.
        ``reset_per_contig``: one ``iter_clusters`` call per contig, as ``gecco run`` does (``_common.py:616-618``)."""
        contig_ptr = numpy.ascontiguousarray(contig_ptr, dtype=numpy.int32)
        prob = numpy.asarray(prob)
        flags = GCRF_FLAG_RESET_PER_CONTIG if reset_per_contig else 0
        if prob.dtype == numpy.float32:
            prob = numpy.ascontiguousarray(prob)
            flags |= GCRF_FLAG_PROB_F32
        else:
            prob = numpy.ascontiguousarray(prob, dtype=numpy.float64)
        annotated = numpy.ascontiguousarray(numpy.asarray(annotated) != 0, dtype=numpy.uint8)
        C, G = len(contig_ptr) - 1, len(prob)
        if len(annotated) != G:
            raise ValueError("prob and annotated must have one entry per gene")
        cap = int(capacity) if capacity is not None else max(1024, G // (8 * max(1, int(n_cds))))
        while True:
            seg = Segments(cap)
            n = ctypes.c_int64(0)
            _check(self._lib, self._lib.gcrf_segments(
                self._handle, contig_ptr.ctypes.data, prob.ctypes.data if G else None,
                annotated.ctypes.data if G else None, C, G, float(threshold), int(n_cds), int(edge_distance),
                int(bool(trim)), seg.contig.ctypes.data, seg.begin.ctypes.data, seg.end.ctypes.data,
                seg.ordinal.ctypes.data, seg.average_p.ctypes.data, seg.max_p.ctypes.data, cap, ctypes.byref(n), flags))
            if n.value <= cap:
                return seg.truncated(n.value)
            cap = int(n.value)  # more clusters than the arrays hold: once more with room for all of them

    def extract_segments(self, contig_ptr, prob, annotated, **kwargs) -> List[Tuple[int, int, int]]:
        """``[(contig, first_gene, last_gene + 1), ...]`` in the reference's order (see ``segments``)."""
        seg = self.segments(contig_ptr, prob, annotated, **kwargs)
        return list(zip(seg.contig.tolist(), seg.begin.tolist(), seg.end.tolist()))

    # ------------------------------------------------------------------ device-pointer calls
    def set_stream(self, cuda_stream: Optional[int]) -> None:
        """Run later calls on the given ``cudaStream_t`` (an int; ``None`` = the engine's own stream)."""
        _check(self._lib, self._lib.gcrf_model_set_stream(self._handle, ctypes.c_void_p(cuda_stream or 0)))

    def marginals_windowed_device(self, contig_ptr: int, gene_ptr: int, attr_idx: int, C: int, G: int, nnz: int,
                                  out: int, *, window: Optional[int] = None, step: Optional[int] = None,
                                  pad: bool = True, f32: bool = False, ptr64: bool = False, flags: int = 0) -> None:
        """Enqueue on device-resident arrays given as raw device addresses; returns without syncing."""
        flags |= GCRF_FLAG_DEVICE_PTRS | (GCRF_FLAG_OUT_F32 if f32 else 0) | (GCRF_FLAG_PTR64 if ptr64 else 0)
        window = self.weights.window_size if window is None else window
        step = self.weights.window_step if step is None else step
        _check(self._lib, self._lib.gcrf_marginals_windowed(
            self._handle, contig_ptr, gene_ptr, attr_idx, C, G, nnz, int(window), int(step), int(bool(pad)), out, flags))

    def marginals_windowed_peers(self, contig_ptr: int, gene_ptr: int, attr_idx: int, C: int, G: int, nnz: int, out: Optional[int],
                                 peer_out: List[int], out_offset: int, *, window: Optional[int] = None,
                                 step: Optional[int] = None, pad: bool = True, f32: bool = False, ptr64: bool = False,
                                 multicast: bool = False) -> None:
        """``gcrf_marginals_windowed_peers``: the windowed kernel on this GPU's shard, every result stored straight into the
        output arrays of the peer GPUs (raw device addresses mapped into this process; with ``multicast`` one NVLS multicast
        address) at ``out_offset + gene`` — the gather of a contig-sharded batch fused into the kernel.  Enqueues only."""
        flags = GCRF_FLAG_DEVICE_PTRS | (GCRF_FLAG_OUT_F32 if f32 else 0) | (GCRF_FLAG_PTR64 if ptr64 else 0) | \
            (GCRF_FLAG_MULTICAST if multicast else 0)
        window = self.weights.window_size if window is None else window
        step = self.weights.window_step if step is None else step
        arr = (ctypes.c_void_p * max(1, len(peer_out)))(*[ctypes.c_void_p(int(p)) for p in peer_out])
        _check(self._lib, self._lib.gcrf_marginals_windowed_peers(
            self._handle, contig_ptr, gene_ptr, attr_idx, C, G, nnz, int(window), int(step), int(bool(pad)),
            ctypes.c_void_p(out or 0), arr, len(peer_out), int(out_offset), flags))

    def features_from_accessions_device(self, accession: int, gene_ptr: int, G: int, nnz: int, out: int, *,
                                        ptr64: bool = False) -> None:
        """``gcrf_features_from_accessions`` on device-resident arrays (raw device addresses); returns without syncing."""
        flags = GCRF_FLAG_DEVICE_PTRS | (GCRF_FLAG_PTR64 if ptr64 else 0)
        _check(self._lib, self._lib.gcrf_features_from_accessions(self._handle, accession, gene_ptr, G, nnz, out, flags))

    def marginals_chain_device(self, contig_ptr: int, gene_ptr: int, attr_idx: int, C: int, G: int, nnz: int,
                               out: int, *, f32: bool = False, ptr64: bool = False) -> None:
        flags = GCRF_FLAG_DEVICE_PTRS | (GCRF_FLAG_OUT_F32 if f32 else 0) | (GCRF_FLAG_PTR64 if ptr64 else 0)
        _check(self._lib, self._lib.gcrf_marginals_chain(self._handle, contig_ptr, gene_ptr, attr_idx, C, G, nnz, out, flags))

    def synchronize(self) -> None:
        _check(self._lib, self._lib.gcrf_model_synchronize(self._handle))

    def max_window(self, f64_arith: bool = False) -> int:
        """Largest ``window`` the device path accepts for this model (``gcrf_max_window``)."""
        return int(self._lib.gcrf_max_window(self._handle, int(bool(f64_arith))))

    @property
    def launch_count(self) -> int:
        return int(self._lib.gcrf_model_launch_count(self._handle))

    def set_timing(self, enable: bool) -> None:
        """Bracket the kernels of later calls with CUDA events (needed by ``last_kernel_ms``; off by default)."""
        _check(self._lib, self._lib.gcrf_model_set_timing(self._handle, int(bool(enable))))

    def last_kernel_ms(self) -> float:
        return float(self._lib.gcrf_model_last_kernel_ms(self._handle))
