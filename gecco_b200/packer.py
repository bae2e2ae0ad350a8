"""Gene records -> ragged CSR arrays, with the reference's ordering and set semantics.

What this restates (``gecco/crf/__init__.py:199-206`` and ``gecco/crf/features.py:13-35``):

* genes are sorted by ``(source.id, start)`` (stable) and grouped by contig;
* inside every gene the domains are sorted by ``start`` — **in place, on the caller's objects**, like
  the reference does (``:200-201``);
* a gene's features are a *dict keyed by domain name*: duplicates collapse, first occurrence keeps its
  place; genes without domains are kept as empty rows (``empty=True``);
* names the model does not know are dropped (python-crfsuite ignores unknown attribute strings).
"""

from __future__ import annotations

import operator
from dataclasses import dataclass, field
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy

__all__ = ["PackedGenes", "pack_genes", "pack_records", "pack_arrays", "pfam_lut", "compact_ids"]


@dataclass
class PackedGenes:
    contig_ptr: numpy.ndarray  # int32 [C+1]
    gene_ptr: numpy.ndarray    # int32 [G+1]
    attr_idx: numpy.ndarray    # int32 [nnz]
    contig_ids: List[Any] = field(default_factory=list)
    order: Optional[numpy.ndarray] = None  # packed position -> index in the caller's sequence
    gene_ids: List[Any] = field(default_factory=list)
    # attr_idx holds integer domain accessions (every domain row, unknown names -1, repeats kept): feature extraction
    # is left to the device (GCRF_FLAG_ACCESSIONS)
    accessions: bool = False

    @property
    def C(self) -> int:
        return len(self.contig_ptr) - 1

    @property
    def G(self) -> int:
        return len(self.gene_ptr) - 1

    @property
    def nnz(self) -> int:
        return len(self.attr_idx)


def _finish(contig_lens: List[int], gene_lens: List[int], attrs: List[int], **kw) -> PackedGenes:
    contig_ptr = numpy.zeros(len(contig_lens) + 1, dtype=numpy.int64)
    numpy.cumsum(contig_lens, out=contig_ptr[1:])
    gene_ptr = numpy.zeros(len(gene_lens) + 1, dtype=numpy.int64)
    numpy.cumsum(gene_lens, out=gene_ptr[1:])
    if gene_ptr[-1] > 0x7FFFFFFF or contig_ptr[-1] > 0x7FFFFFFF:
        raise ValueError("batch too large for int32 row pointers; shard it")
    return PackedGenes(contig_ptr.astype(numpy.int32), gene_ptr.astype(numpy.int32),
                       numpy.asarray(attrs, dtype=numpy.int32), **kw)


def pack_records(records: Iterable[Tuple[Any, int, Any, Sequence[str]]], attr_index: Dict[str, int]) -> PackedGenes:
    """``records`` = ``(contig_id, gene_start, gene_id, [domain names ordered by domain start])``."""
    records = list(records)
    order = sorted(range(len(records)), key=lambda i: (records[i][0], records[i][1]))
    contig_ids: List[Any] = []
    contig_lens: List[int] = []
    gene_lens: List[int] = []
    attrs: List[int] = []
    for i in order:
        cid, _start, _gid, names = records[i]
        if not contig_ids or contig_ids[-1] != cid:
            contig_ids.append(cid)
            contig_lens.append(0)
        contig_lens[-1] += 1
        n0 = len(attrs)
        for name in dict.fromkeys(names):  # dict keys: unique, first occurrence first
            a = attr_index.get(name)
            if a is not None:
                attrs.append(a)
        gene_lens.append(len(attrs) - n0)
    return _finish(contig_lens, gene_lens, attrs, contig_ids=contig_ids,
                   order=numpy.asarray(order, dtype=numpy.int64), gene_ids=[records[i][2] for i in order])


def pack_genes(genes: Iterable[Any], attr_index: Dict[str, int], feature_type: str = "protein"
               ) -> Tuple[PackedGenes, List[Any], List[Tuple[int, int]]]:
    """Duck-typed ``gecco.model.Gene`` objects -> CSR.

    Returns ``(packed, sorted_genes, contig_slices)``; ``contig_slices[c] = (first, last+1)`` indexes
    ``sorted_genes``.  In ``"protein"`` mode there is one row per gene; in ``"domain"`` mode one row
    per domain and one empty row per domain-less gene (``features.py:38-48``).
    """
    if feature_type not in ("protein", "domain"):
        raise ValueError(f"invalid feature type: {feature_type!r}")
    genes = sorted(genes, key=operator.attrgetter("source.id", "start"))
    for gene in genes:
        gene.protein.domains.sort(key=operator.attrgetter("start"))  # mutates, like the reference
    contig_ids: List[Any] = []
    contig_lens: List[int] = []
    slices: List[Tuple[int, int]] = []
    gene_lens: List[int] = []
    attrs: List[int] = []
    for k, gene in enumerate(genes):
        cid = gene.source.id
        if not contig_ids or contig_ids[-1] != cid:
            contig_ids.append(cid)
            contig_lens.append(0)
            slices.append((k, k))
        slices[-1] = (slices[-1][0], k + 1)
        domains = gene.protein.domains
        if feature_type == "protein":
            contig_lens[-1] += 1
            n0 = len(attrs)
            for name in dict.fromkeys(d.name for d in domains):
                a = attr_index.get(name)
                if a is not None:
                    attrs.append(a)
            gene_lens.append(len(attrs) - n0)
        elif domains:
            for d in domains:
                contig_lens[-1] += 1
                a = attr_index.get(d.name)
                if a is not None:
                    attrs.append(a)
                    gene_lens.append(1)
                else:
                    gene_lens.append(0)
        else:
            contig_lens[-1] += 1
            gene_lens.append(0)
    return _finish(contig_lens, gene_lens, attrs, contig_ids=contig_ids), genes, slices


def compact_ids(attr_idx: numpy.ndarray, num_attrs: int) -> numpy.ndarray:
    """``int32`` attribute ids -> ``uint16`` with ``0xFFFF`` for every id outside ``[0, num_attrs)`` — the layout of
    ``GCRF_FLAG_IDX_U16``: the ids are what a host-buffer call moves over PCIe, this halves them."""
    if num_attrs >= 0xFFFF:
        raise ValueError("compact ids need a model with fewer than 65535 attributes")
    attr_idx = numpy.asarray(attr_idx)
    out = attr_idx.astype(numpy.uint16)
    out[(attr_idx < 0) | (attr_idx >= num_attrs)] = 0xFFFF
    return out


def pfam_lut(attrs: Sequence[str]) -> numpy.ndarray:
    """``lut[n]`` = attribute id of Pfam accession ``PF{n:05d}`` or -1."""
    nums = []
    for name in attrs:
        if not (len(name) == 7 and name.startswith("PF") and name[2:].isdigit()):
            raise ValueError(f"attribute {name!r} is not a Pfam accession; use the name-based packer")
        nums.append(int(name[2:]))
    lut = numpy.full(max(nums, default=0) + 1, -1, dtype=numpy.int32)
    lut[numpy.asarray(nums, dtype=numpy.int64)] = numpy.arange(len(nums), dtype=numpy.int32)
    return lut


def pack_arrays(gene_contig: numpy.ndarray, dom_ptr: numpy.ndarray, dom_pfam: numpy.ndarray, weights) -> PackedGenes:
    """Bulk (vectorised) packer for table-shaped input.

    ``gene_contig[g]`` is the contig number of gene ``g`` (genes already in the reference's order, so
    it is non-decreasing), ``dom_pfam[dom_ptr[g]:dom_ptr[g+1]]`` the Pfam accession numbers of its
    domain rows ordered by domain start.  Unknown accessions are dropped and repeats inside a gene
    collapse onto their first occurrence.
    """
    gene_contig = numpy.asarray(gene_contig, dtype=numpy.int64)
    dom_ptr = numpy.asarray(dom_ptr, dtype=numpy.int64)
    dom_pfam = numpy.asarray(dom_pfam, dtype=numpy.int64)
    G = len(gene_contig)
    if len(dom_ptr) != G + 1:
        raise ValueError("dom_ptr must have G+1 entries")
    if G and numpy.any(numpy.diff(gene_contig) < 0):
        raise ValueError("genes must be ordered by contig")
    lut = pfam_lut(weights.attrs)
    ids = numpy.full(len(dom_pfam), -1, dtype=numpy.int64)
    ok = (dom_pfam >= 0) & (dom_pfam < len(lut))
    ids[ok] = lut[dom_pfam[ok]]
    gene_of = numpy.repeat(numpy.arange(G, dtype=numpy.int64), numpy.diff(dom_ptr))
    keep = ids >= 0
    key = gene_of * (len(weights.attrs) + 1) + ids
    srt = numpy.argsort(key, kind="stable")
    dup = numpy.zeros(len(key), dtype=bool)
    dup[srt[1:]] = key[srt[1:]] == key[srt[:-1]]
    keep &= ~dup
    counts = numpy.bincount(gene_of[keep], minlength=G)
    gene_ptr = numpy.zeros(G + 1, dtype=numpy.int64)
    numpy.cumsum(counts, out=gene_ptr[1:])
    if G:
        change = numpy.flatnonzero(numpy.diff(gene_contig)) + 1
        contig_ptr = numpy.concatenate([[0], change, [G]])
        contig_ids = gene_contig[contig_ptr[:-1]].tolist()
    else:
        contig_ptr = numpy.zeros(1, dtype=numpy.int64)
        contig_ids = []
    if gene_ptr[-1] > 0x7FFFFFFF:
        raise ValueError("batch too large for int32 row pointers; shard it")
    return PackedGenes(contig_ptr.astype(numpy.int32), gene_ptr.astype(numpy.int32),
                       ids[keep].astype(numpy.int32), contig_ids=contig_ids)
