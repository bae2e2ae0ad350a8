"""`gecco predict`'s table path without per-row Python objects.

The reference reads ``*.genes.tsv`` / ``*.features.tsv`` with polars and rebuilds one ``Gene`` / ``Protein`` /
``Domain`` object per row before the CRF ever runs (``gecco/cli/commands/predict.py:62-88``), then converts
the objects back into tables to write the results (``_common.py:47-76``).  ``FeatureTables`` keeps everything in
native arrays owned by ``libgecco_crf_b200.so`` (``csrc/gcrf_tables.cpp``): load + annotate + sort + filter,
pack to the CSR batch of the marginal kernels, write the result tables from the probability array.

    tables = FeatureTables.load("x.genes.tsv", ["x.features.tsv"], p_filter=1e-9)
    crf = ClusterCRF.trained()
    prob = tables.predict(crf)                 # float64 per packed row, on the B200
    tables.write_genes("out/x.genes.tsv", prob)
    tables.write_features("out/x.features.tsv", prob)
    clusters = tables.segments(crf, prob, threshold=0.8, n_cds=3)
"""

from __future__ import annotations

import bz2
import collections.abc
import ctypes
import gzip
import lzma
import math
import os
from typing import Iterable, List, Optional, Sequence, Union

import numpy

from ._lib import GcrfError, load_library
from .packer import PackedGenes

__all__ = ["FeatureTables", "predict_tables"]

PathLike = Union[str, "os.PathLike[str]"]
_MAGIC = ((b"\x1f\x8b", gzip.open), (b"BZh", bz2.open), (b"\xfd7zXZ", lzma.open))


def _read(path: PathLike) -> Optional[bytes]:
    """Bytes of a compressed table (what ``gecco._meta.zopen`` sniffs, ``gecco/_meta.py:169-185``), else None."""
    with open(path, "rb") as f:
        head = f.read(6)
    for magic, opener in _MAGIC:
        if head.startswith(magic):
            with opener(path, "rb") as f:
                return f.read()
    return None


def _nan_if_none(x: Optional[float]) -> float:
    return math.nan if x is None else float(x)


class _LazyNames(collections.abc.Sequence):
    """Names held by the native table, fetched when somebody looks: a metagenome table has a million contig ids, and
    ``pack`` hands them along with every batch without anybody reading them on the bulk path."""

    def __init__(self, n: int, fetch):
        self._n, self._fetch, self._all = n, fetch, None

    def __len__(self) -> int:
        return self._n

    def _list(self) -> List[str]:
        if self._all is None:
            self._all = [self._fetch(i).decode() for i in range(self._n)]
        return self._all

    def __getitem__(self, i):
        if isinstance(i, slice) or self._all is not None:
            return self._list()[i]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        return self._fetch(i).decode()

    def __iter__(self):
        return iter(self._list())

    def __eq__(self, other):
        return self._list() == list(other) if isinstance(other, (list, tuple, _LazyNames)) else NotImplemented

    def __repr__(self) -> str:
        return repr(self._list())


class FeatureTables:
    """A genes table plus its feature tables, annotated, sorted and filtered like ``gecco predict`` does."""

    def __init__(self, handle: ctypes.c_void_p):
        self._lib = load_library()
        self._h = handle
        self._packed: Optional[PackedGenes] = None
        self.feature_type: Optional[str] = None
        self._contig_ids: Optional[_LazyNames] = None
        self._gene_ids: Optional[_LazyNames] = None

    # ------------------------------------------------------------------ construction
    @classmethod
    def load(cls, genes: PathLike, features: Union[PathLike, Iterable[PathLike]], *, e_filter: Optional[float] = None,
             p_filter: Optional[float] = 1e-9) -> "FeatureTables":
        """``p_filter=1e-9`` is the default of ``gecco predict`` (``gecco/cli/commands/_parser.py:171-175``)."""
        lib = load_library()
        if isinstance(features, (str, os.PathLike)):
            features = [features]
        paths = [os.fspath(genes)] + [os.fspath(p) for p in features]
        blobs = [_read(p) for p in paths]
        handle = ctypes.c_void_p()
        if any(b is not None for b in blobs):  # a compressed table: hand the library memory buffers
            data = [b if b is not None else open(p, "rb").read() for p, b in zip(paths, blobs)]
            return cls.parse(data[0], data[1:], e_filter=e_filter, p_filter=p_filter)
        arr = (ctypes.c_char_p * max(1, len(paths) - 1))(*[p.encode() for p in paths[1:]])
        rc = lib.gcrf_table_load(paths[0].encode(), arr, len(paths) - 1, _nan_if_none(e_filter), _nan_if_none(p_filter),
                                 ctypes.byref(handle))
        if rc != 0:
            raise _table_error(lib, rc)
        return cls(handle)

    @classmethod
    def parse(cls, genes: bytes, features: Sequence[bytes], *, e_filter: Optional[float] = None,
              p_filter: Optional[float] = 1e-9) -> "FeatureTables":
        lib = load_library()
        handle = ctypes.c_void_p()
        n = len(features)
        bufs = (ctypes.c_char_p * max(1, n))(*features)
        lens = (ctypes.c_uint64 * max(1, n))(*[len(b) for b in features])
        rc = lib.gcrf_table_parse(genes, len(genes), bufs, lens, n, _nan_if_none(e_filter), _nan_if_none(p_filter),
                                  ctypes.byref(handle))
        if rc != 0:
            raise _table_error(lib, rc)
        return cls(handle)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.gcrf_table_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ shape
    @property
    def contigs(self) -> int:
        return int(self._lib.gcrf_table_contigs(self._h))

    @property
    def genes(self) -> int:
        return int(self._lib.gcrf_table_genes(self._h))

    @property
    def domains(self) -> int:
        return int(self._lib.gcrf_table_domains(self._h))

    @property
    def contig_ids(self) -> Sequence[str]:
        if self._contig_ids is None:
            self._contig_ids = _LazyNames(self.contigs, lambda c: self._lib.gcrf_table_contig_id(self._h, c))
        return self._contig_ids

    @property
    def gene_ids(self) -> Sequence[str]:
        if self._gene_ids is None:
            self._gene_ids = _LazyNames(self.genes, lambda g: self._lib.gcrf_table_gene_id(self._h, g))
        return self._gene_ids

    def _view(self, ptr: Optional[int], n: int, dtype) -> numpy.ndarray:
        if not ptr or n == 0:
            return numpy.zeros(n, dtype=dtype)
        ctype = {numpy.int32: ctypes.c_int32, numpy.uint8: ctypes.c_uint8}[dtype]
        return numpy.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctype)), shape=(n,))

    @property
    def contig_ptr(self) -> numpy.ndarray:
        """``int32[C+1]`` into genes."""
        if self.genes == 0:
            return numpy.zeros(1, dtype=numpy.int32)
        return self._view(self._lib.gcrf_table_contig_ptr(self._h), self.contigs + 1, numpy.int32)

    @property
    def annotated(self) -> numpy.ndarray:
        """``uint8[G]``: the gene kept at least one domain (``gene.protein.domains`` is non-empty)."""
        return self._view(self._lib.gcrf_table_annotated(self._h), self.genes, numpy.uint8)

    def gene_coordinates(self):
        start = numpy.zeros(self.genes, dtype=numpy.int64)
        end = numpy.zeros(self.genes, dtype=numpy.int64)
        rc = self._lib.gcrf_table_gene_coordinates(self._h, start.ctypes.data, end.ctypes.data)
        if rc != 0:
            raise _table_error(self._lib, rc)
        return start, end

    # ------------------------------------------------------------------ features
    def pack(self, attrs: Optional[Sequence[str]], feature_type: str = "protein", *, accessions: bool = False,
             digits: int = 5) -> PackedGenes:
        """CSR batch for the marginal kernels (views into the table: valid until the next ``pack`` / ``close``).

        ``accessions=True`` (``gcrf_table_pack_accessions``): nothing is looked up or de-duplicated on the host; the
        batch holds the Pfam number of every kept domain row and the device does the feature extraction
        (``GCRF_FLAG_ACCESSIONS``; for models whose attributes are all ``PF`` + ``digits`` digits — only domain names
        of exactly that form count, the reference compares names — ``attrs`` is not needed)."""
        if feature_type not in ("protein", "domain"):
            raise ValueError(f"invalid feature type: {feature_type!r}")
        cp, rp, ai = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        rows, nnz = ctypes.c_int64(), ctypes.c_int64()
        if accessions:
            rc = self._lib.gcrf_table_pack_accessions(self._h, 0 if feature_type == "protein" else 1, int(digits), ctypes.byref(cp),
                                                      ctypes.byref(rp), ctypes.byref(ai), ctypes.byref(rows), ctypes.byref(nnz))
        else:
            names = (ctypes.c_char_p * max(1, len(attrs)))(*[a.encode() for a in attrs])
            rc = self._lib.gcrf_table_pack(self._h, names, len(attrs), 0 if feature_type == "protein" else 1, ctypes.byref(cp),
                                           ctypes.byref(rp), ctypes.byref(ai), ctypes.byref(rows), ctypes.byref(nnz))
        if rc != 0:
            raise _table_error(self._lib, rc)
        C = self.contigs
        packed = PackedGenes(self._view(cp.value, C + 1, numpy.int32) if self.genes else numpy.zeros(1, dtype=numpy.int32),
                             self._view(rp.value, rows.value + 1, numpy.int32), self._view(ai.value, nnz.value, numpy.int32),
                             contig_ids=self.contig_ids, accessions=accessions)
        self._packed, self.feature_type = packed, feature_type
        return packed

    @property
    def row_gene(self) -> numpy.ndarray:
        n = self._packed.G if self._packed is not None else 0
        return self._view(self._lib.gcrf_table_row_gene(self._h), n, numpy.int32)

    def predict(self, crf, *, pad: bool = True) -> numpy.ndarray:
        """Pack with the model's vocabulary and feature type, run the marginal kernels: one value per packed row."""
        # Pfam-only vocabularies (the shipped model): the table hands over raw accession numbers and the feature
        # extraction runs on the device; GECCO_B200_HOST_FEATURES=1 keeps the host packer (A/B, other vocabularies)
        engine = crf._get_engine()
        on_device = os.environ.get("GECCO_B200_HOST_FEATURES", "0") != "1" and engine.has_vocabulary
        packed = self.pack(crf._weights.attrs, crf.feature_type, accessions=on_device,
                           digits=getattr(engine, "vocabulary_digits", 0))
        return crf.marginals(packed, pad=pad) if packed.G else numpy.zeros(0)

    # ------------------------------------------------------------------ results
    def gene_probabilities(self, row_prob: numpy.ndarray):
        """``(average_p, max_p)`` per gene (``Gene.average_probability`` / ``maximum_probability``)."""
        prob = numpy.ascontiguousarray(row_prob, dtype=numpy.float64)
        avg = numpy.empty(self.genes)
        mx = numpy.empty(self.genes)
        rc = self._lib.gcrf_table_gene_probabilities(self._h, prob.ctypes.data, avg.ctypes.data, mx.ctypes.data)
        if rc != 0:
            raise _table_error(self._lib, rc)
        return avg, mx

    def _prob_ptr(self, row_prob):
        if row_prob is None:
            return None, None
        if self._packed is None:
            raise ValueError("pack() or predict() first")
        prob = numpy.ascontiguousarray(row_prob, dtype=numpy.float64)
        if prob.shape != (self._packed.G,):
            raise ValueError("gene and probability lists don't have the same length")  # features.py:93-94
        return prob, prob.ctypes.data

    def write_genes(self, path: PathLike, row_prob: Optional[numpy.ndarray] = None) -> None:
        """``GeneTable.from_genes(genes).dump(path)`` (``gecco/cli/commands/_common.py:47-60``)."""
        keep, ptr = self._prob_ptr(row_prob)
        rc = self._lib.gcrf_table_write_genes(self._h, ptr, os.fspath(path).encode())
        if rc != 0:
            raise _table_error(self._lib, rc)

    def write_features(self, path: PathLike, row_prob: Optional[numpy.ndarray] = None) -> None:
        """``FeatureTable.from_genes(genes).dump(path)`` (``gecco/cli/commands/_common.py:63-76``)."""
        keep, ptr = self._prob_ptr(row_prob)
        rc = self._lib.gcrf_table_write_features(self._h, ptr, os.fspath(path).encode())
        if rc != 0:
            raise _table_error(self._lib, rc)

    def find_segments(self, crf, row_prob: numpy.ndarray, *, threshold: float = 0.8, n_cds: int = 3, edge_distance: int = 0,
                      trim: bool = True):
        """``gcrf_segments`` on this table's genes: the raw ``Segments`` arrays (see :meth:`segments`)."""
        avg, _ = self.gene_probabilities(row_prob)
        return crf._get_engine().segments(self.contig_ptr, avg, self.annotated, threshold=threshold, n_cds=n_cds,
                                          edge_distance=edge_distance, trim=trim, reset_per_contig=True)

    def write_clusters(self, path: PathLike, row_prob: numpy.ndarray, seg) -> None:
        """``ClusterTable.from_clusters(clusters).dump(path)`` for clusters without a predicted type
        (``gecco/cli/commands/_common.py:79-92``, ``gecco/model.py:735-771``); ``seg`` comes from :meth:`find_segments`."""
        keep, ptr = self._prob_ptr(row_prob)
        arrays = [numpy.ascontiguousarray(a, dtype=numpy.int32) for a in (seg.contig, seg.begin, seg.end, seg.ordinal)]
        rc = self._lib.gcrf_table_write_clusters(self._h, ptr, *[a.ctypes.data for a in arrays], len(arrays[0]),
                                                 os.fspath(path).encode())
        if rc != 0:
            raise _table_error(self._lib, rc)

    def segments(self, crf, row_prob: numpy.ndarray, *, threshold: float = 0.8, n_cds: int = 3, edge_distance: int = 0,
                 trim: bool = True):
        """Cluster segments straight from the arrays (``ClusterRefiner.iter_clusters``, ``gecco/refine.py:118-134``;
        defaults of ``gecco predict``: ``--threshold 0.8 --cds 3``).  Returns rows
        ``(contig_id, first_gene_id, last_gene_id, first_gene, last_gene + 1, average_p, max_p)``."""
        seg = self.find_segments(crf, row_prob, threshold=threshold, n_cds=n_cds, edge_distance=edge_distance, trim=trim)
        ids, contigs = self.gene_ids, self.contig_ids
        return [(contigs[int(c)], ids[int(b)], ids[int(e) - 1], int(b), int(e), float(a), float(m))
                for c, b, e, a, m in zip(seg.contig, seg.begin, seg.end, seg.average_p, seg.max_p)]


def _table_error(lib, rc: int) -> Exception:
    message = lib.gcrf_table_last_error().decode("utf-8", "replace")
    # the reference raises ValueError for every inconsistency in the tables (_common.py:217-249)
    return ValueError(message) if rc == -1 else GcrfError(rc, message)


def predict_tables(genes: PathLike, features: Union[PathLike, Iterable[PathLike]], output_dir: PathLike, *, model=None,
                   base: Optional[str] = None, e_filter: Optional[float] = None, p_filter: Optional[float] = 1e-9,
                   pad: bool = True, threshold: float = 0.8, cds: int = 3, edge_distance: int = 0, trim: bool = True,
                   clusters: bool = True):
    """The table-to-table part of ``gecco predict`` (``predict.py:62-110``): load, annotate, sort, filter, CRF
    marginals on the B200, write ``{base}.genes.tsv`` and ``{base}.features.tsv``; then threshold + segment
    extraction on the device and ``{base}.clusters.tsv`` (untyped: the type classifier is not part of this package;
    like the reference, no clusters table is written when nothing is found).  Returns ``(tables, row_prob)``."""
    from .crf import ClusterCRF

    crf = model if isinstance(model, ClusterCRF) else ClusterCRF.trained(model)
    tables = FeatureTables.load(genes, features, e_filter=e_filter, p_filter=p_filter)
    prob = tables.predict(crf, pad=pad)
    os.makedirs(output_dir, exist_ok=True)
    if base is None:
        base = os.path.basename(os.fspath(genes))
        for suffix in (".gz", ".bz2", ".xz", ".tsv", ".genes"):
            if base.endswith(suffix):
                base = base[: -len(suffix)]
    tables.write_genes(os.path.join(output_dir, f"{base}.genes.tsv"), prob)
    tables.write_features(os.path.join(output_dir, f"{base}.features.tsv"), prob)
    if clusters and tables.genes:
        seg = tables.find_segments(crf, prob, threshold=threshold, n_cds=cds, edge_distance=edge_distance, trim=trim)
        if len(seg.contig):
            tables.write_clusters(os.path.join(output_dir, f"{base}.clusters.tsv"), prob, seg)
    return tables, prob
