"""Synthetic ragged contig batches of the BASELINE.json configs (shapes/seeds: SURVEY.md §8(d)).

Everything is produced directly in the CSR layout the C ABI takes:
``contig_ptr[C+1]`` -> genes, ``gene_ptr[G+1]`` -> ``attr_idx[nnz]`` (ids unique and sorted inside a
gene, i.e. what the packer emits after the reference's set semantics, ``gecco/crf/features.py:32``).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy

__all__ = ["CsrBatch", "make_batch", "config2", "config3_ecoli_like", "config4", "config4_lengths", "config4_chunked",
           "config5", "algorithmic_bytes", "concat_batches"]


@dataclass
class CsrBatch:
    contig_ptr: numpy.ndarray  # int32 [C+1]
    gene_ptr: numpy.ndarray    # int32 or int64 [G+1]
    attr_idx: numpy.ndarray    # int32 [nnz]
    name: str = ""

    @property
    def C(self) -> int:
        return len(self.contig_ptr) - 1

    @property
    def G(self) -> int:
        return len(self.gene_ptr) - 1

    @property
    def nnz(self) -> int:
        return len(self.attr_idx)

    def windows(self, window: int, step: int = 1, pad: bool = True) -> int:
        """Number of windows the reference loop evaluates (``gecco/crf/__init__.py:239`` for step 1)."""
        n = numpy.diff(self.contig_ptr).astype(numpy.int64)
        long = n[n >= window]
        total = int(((long - window) // step + 1).sum())
        if pad:
            total += int((n < window).sum())
        return total

    def slice_contigs(self, c0: int, c1: int) -> "CsrBatch":
        """Contigs [c0, c1) as an independent batch (pointers rebased) — the sharding primitive."""
        g0, g1 = int(self.contig_ptr[c0]), int(self.contig_ptr[c1])
        p0, p1 = int(self.gene_ptr[g0]), int(self.gene_ptr[g1])
        return CsrBatch(
            (self.contig_ptr[c0:c1 + 1] - g0).astype(numpy.int32),
            (self.gene_ptr[g0:g1 + 1] - p0).astype(self.gene_ptr.dtype),
            self.attr_idx[p0:p1],
            name=f"{self.name}[{c0}:{c1}]",
        )


def algorithmic_bytes(C: int, G: int, nnz: int, out_itemsize: int = 8) -> int:
    """4*nnz (attr_idx) + 4*(G+1) (gene_ptr) + 4*(C+1) (contig_ptr) + 8*G (out f64)."""
    return 4 * nnz + 4 * (G + 1) + 4 * (C + 1) + out_itemsize * G


def make_batch(rng: numpy.random.Generator, genes_per_contig: numpy.ndarray, mean_domains: float, num_attrs: int,
               unknown_fraction: float = 0.0, name: str = "", zipf: float = 0.0) -> CsrBatch:
    """k_g ~ Poisson(mean_domains) ids uniform over the vocabulary (``zipf`` > 0: attribute of rank r with probability
    ~ r^-zipf, ranks assigned by a random permutation — the frequency skew of real Pfam annotations),
    made unique and sorted per gene."""
    genes_per_contig = numpy.asarray(genes_per_contig, dtype=numpy.int64)
    contig_ptr = numpy.zeros(len(genes_per_contig) + 1, dtype=numpy.int64)
    numpy.cumsum(genes_per_contig, out=contig_ptr[1:])
    G = int(contig_ptr[-1])
    k = rng.poisson(mean_domains, size=G).astype(numpy.int64)
    raw = int(k.sum())
    if zipf > 0:
        cdf = numpy.cumsum(numpy.arange(1, num_attrs + 1, dtype=numpy.float64) ** -zipf)
        ranks = numpy.searchsorted(cdf, rng.random(raw) * cdf[-1], side="right").clip(0, num_attrs - 1)
        ids = rng.permutation(num_attrs)[ranks].astype(numpy.int64)
    else:
        ids = rng.integers(0, num_attrs, size=raw, dtype=numpy.int64)
    gene_of = numpy.repeat(numpy.arange(G, dtype=numpy.int64), k)
    key = gene_of * (1 << 20) + ids  # vocabularies stay far below 2^20
    key.sort()
    keep = numpy.ones(raw, dtype=bool)
    keep[1:] = key[1:] != key[:-1]
    key = key[keep]
    gene_of = key >> 20
    attr = (key & ((1 << 20) - 1)).astype(numpy.int32)
    counts = numpy.bincount(gene_of, minlength=G)
    gene_ptr = numpy.zeros(G + 1, dtype=numpy.int64)
    numpy.cumsum(counts, out=gene_ptr[1:])
    if unknown_fraction > 0 and len(attr):
        unknown = rng.random(len(attr)) < unknown_fraction
        attr[unknown] = -1
    ptr_dtype = numpy.int32 if int(gene_ptr[-1]) <= 0x7FFFFFFF else numpy.int64
    return CsrBatch(contig_ptr.astype(numpy.int32), gene_ptr.astype(ptr_dtype), attr, name=name)


def config2(num_attrs: int = 2659, seed: int = 2, contigs: int = 10_000, mean_genes: float = 200.0,
            mean_domains: float = 25.0, unknown_fraction: float = 0.05, zipf: float = 0.0) -> CsrBatch:
    """10k contigs x Poisson(200) genes x Poisson(25) domains — the headline 1xB200 workload."""
    rng = numpy.random.default_rng(seed)
    n = numpy.maximum(1, rng.poisson(mean_genes, size=contigs))
    return make_batch(rng, n, mean_domains, num_attrs, unknown_fraction, name=f"config2(seed={seed})", zipf=zipf)


def config3_ecoli_like(num_attrs: int = 2659, seed: int = 3, genes: int = 4300, mean_domains: float = 1.4) -> CsrBatch:
    """One contig of 4,300 genes, Poisson(1.4) domains — stand-in for the absent E. coli table."""
    rng = numpy.random.default_rng(seed)
    return make_batch(rng, numpy.array([genes]), mean_domains, num_attrs, 0.0, name=f"config3(seed={seed})")


def config4(num_attrs: int = 2659, seed: int = 4, contigs: int = 1_000_000, mean_domains: float = 25.0,
            unknown_fraction: float = 0.05) -> CsrBatch:
    """Metagenome: lognormal contig lengths (mean ~40 genes, ~31 % shorter than W=20)."""
    rng = numpy.random.default_rng(seed)
    n = numpy.maximum(1, numpy.rint(rng.lognormal(mean=numpy.log(40.0) - 0.32, sigma=0.8, size=contigs))).astype(numpy.int64)
    return make_batch(rng, n, mean_domains, num_attrs, unknown_fraction, name=f"config4(seed={seed})")


def config4_lengths(seed: int = 4, contigs: int = 1_000_000) -> numpy.ndarray:
    """Genes per contig of config 4 — the first draw of ``config4``'s generator, so G is the same (39,973,222 for seed 4)."""
    rng = numpy.random.default_rng(seed)
    return numpy.maximum(1, numpy.rint(rng.lognormal(mean=numpy.log(40.0) - 0.32, sigma=0.8, size=contigs))).astype(numpy.int64)


def concat_batches(parts) -> CsrBatch:
    """Concatenate independent batches (pointers rebased) — the inverse of ``slice_contigs``."""
    parts = list(parts)
    g = numpy.cumsum([0] + [p.G for p in parts])
    z = numpy.cumsum([0] + [p.nnz for p in parts])
    ptr_dtype = numpy.int32 if int(z[-1]) <= 0x7FFFFFFF else numpy.int64
    contig_ptr = numpy.concatenate([parts[0].contig_ptr[:1]] + [p.contig_ptr[1:] + g[i] for i, p in enumerate(parts)]).astype(numpy.int32)
    gene_ptr = numpy.concatenate([numpy.zeros(1, dtype=ptr_dtype)] + [p.gene_ptr[1:].astype(ptr_dtype) + ptr_dtype(z[i]) for i, p in enumerate(parts)])
    return CsrBatch(contig_ptr, gene_ptr, numpy.concatenate([p.attr_idx for p in parts]), name="+".join(p.name for p in parts[:2]) + "...")


def config4_chunked(num_attrs: int = 2659, seed: int = 4, contigs: int = 1_000_000, mean_domains: float = 25.0,
                    unknown_fraction: float = 0.05, contig_range=None, chunk: int = 8192, threads: int = 0) -> CsrBatch:
    """Config 4 at sizes where one global sort of a billion (gene, id) keys would take minutes: the same contig lengths
    as ``config4`` (so the same G), the attribute ids drawn per chunk of ``chunk`` contigs from ``default_rng([seed, j])``
    — chunks are independent, run on ``threads`` host threads (0 = all), and a rank that owns contigs
    ``contig_range = (c0, c1)`` generates only the chunks it needs.  Not the same ids as ``config4`` (same distribution)."""
    import concurrent.futures
    import os

    lens = config4_lengths(seed, contigs)
    c0, c1 = (0, contigs) if contig_range is None else contig_range
    if c1 <= c0:
        return CsrBatch(numpy.zeros(1, dtype=numpy.int32), numpy.zeros(1, dtype=numpy.int32), numpy.zeros(0, dtype=numpy.int32))
    first, last = c0 // chunk, (c1 - 1) // chunk

    def one(j: int) -> CsrBatch:
        a, b = max(c0, j * chunk), min(c1, (j + 1) * chunk)
        rng = numpy.random.default_rng([seed, j])
        whole = make_batch(rng, lens[j * chunk:min(contigs, (j + 1) * chunk)], mean_domains, num_attrs, unknown_fraction)
        return whole if (a, b) == (j * chunk, min(contigs, (j + 1) * chunk)) else whole.slice_contigs(a - j * chunk, b - j * chunk)

    n = threads or len(os.sched_getaffinity(0))
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, n)) as pool:
        parts = list(pool.map(one, range(first, last + 1)))
    out = concat_batches(parts)
    out.name = f"config4_chunked(seed={seed})[{c0}:{c1}]"
    return out


def config5(num_attrs: int = 2659, seed: int = 5, contigs: int = 100, genes: int = 5000, mean_domains: float = 25.0) -> CsrBatch:
    """100 long contigs x 5,000 genes (deep-chain stress)."""
    rng = numpy.random.default_rng(seed)
    return make_batch(rng, numpy.full(contigs, genes), mean_domains, num_attrs, 0.0, name=f"config5(seed={seed})")


def ragged_edge_cases(num_attrs: int = 2659, seed: int = 7, window: int = 20) -> CsrBatch:
    """Small batch that hits every branch: 1-gene contigs, n = W-1, W, W+1, tile-straddling contigs, empty genes."""
    rng = numpy.random.default_rng(seed)
    lens = [1, 1, 2, window - 1, window, window + 1, 3, 700, 5, 19, 21, 236, 237, 1, 40, 7, 255, 256, 257, 12, 1000, 2, 2, 2]
    lens += list(rng.integers(1, 60, size=200))
    batch = make_batch(rng, numpy.array(lens), 3.0, num_attrs, 0.1, name="ragged_edge_cases")
    return batch
