"""Multi-GPU: contigs are independent, so the path shards by contig with no collective inside the math
(SURVEY.md §8(e)).  One process per GPU (``torch.distributed``); NCCL over NVLink only moves CSR shards out
(optional — each rank can also ingest its own shard from host memory, which is what saturates 8 PCIe links
instead of one GPU's NVLink egress) and per-gene marginals back.

``partition_contigs`` is pure numpy; the collectives work on whatever backend the process group uses
(``nccl`` with CUDA tensors on the GPU box, ``gloo`` with CPU tensors in the CPU test-suite).
"""

from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy

from .synth import CsrBatch

__all__ = ["partition_contigs", "scatter_batch", "gather_marginals", "predict_sharded"]


def partition_contigs(contig_ptr: numpy.ndarray, gene_ptr: numpy.ndarray, n_shards: int, window: int,
                      kappa: float = 2.0) -> List[Tuple[int, int]]:
    """Split contigs into ``n_shards`` contiguous ranges of roughly equal cost.

    Cost of a contig = its attribute ids (the bytes streamed from HBM) + ``kappa * windows * window`` (the
    dynamic-programming steps it triggers).  Contiguous ranges keep every shard a plain slice of the CSR arrays.
    Returns ``[(c_begin, c_end), ...]``; trailing shards may be empty when there are fewer contigs than shards.
    """
    contig_ptr = numpy.asarray(contig_ptr, dtype=numpy.int64)
    gene_ptr = numpy.asarray(gene_ptr, dtype=numpy.int64)
    C = len(contig_ptr) - 1
    if n_shards <= 0:
        raise ValueError("n_shards must be positive")
    if C == 0:
        return [(0, 0)] * n_shards
    n = numpy.diff(contig_ptr)
    nnz = gene_ptr[contig_ptr[1:]] - gene_ptr[contig_ptr[:-1]]
    windows = numpy.maximum(n, window) - window + 1
    cost = nnz + kappa * windows * window
    cum = numpy.concatenate([[0.0], numpy.cumsum(cost, dtype=numpy.float64)])
    targets = cum[-1] * numpy.arange(1, n_shards) / n_shards
    cuts = numpy.searchsorted(cum, targets, side="left")
    cuts = numpy.clip(cuts, 0, C)
    bounds = numpy.concatenate([[0], numpy.maximum.accumulate(cuts), [C]])
    return [(int(bounds[i]), int(bounds[i + 1])) for i in range(n_shards)]


def _dist():
    import torch.distributed as dist

    return dist


def scatter_batch(batch: Optional[CsrBatch], window: int, src: int = 0, device=None) -> CsrBatch:
    """Rank ``src`` holds the whole batch; every rank returns its own contiguous shard.

    Grouped point-to-point sends (NCCL has no scatterv): sizes first, then the three arrays.
    """
    import torch

    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = device if device is not None else torch.device("cpu")
    if rank == src:
        assert batch is not None
        parts = partition_contigs(batch.contig_ptr, batch.gene_ptr, world, window)
        shards = [batch.slice_contigs(c0, c1) for c0, c1 in parts]
        meta = [[s.C, s.G, s.nnz, int(s.gene_ptr.dtype == numpy.int64)] for s in shards]
    else:
        shards, meta = None, None
    meta_list = [meta]
    dist.broadcast_object_list(meta_list, src=src)
    meta = meta_list[0]
    C, G, nnz, p64 = meta[rank]
    ptr_dtype = torch.int64 if p64 else torch.int32
    if rank == src:
        reqs = []
        for r, s in enumerate(shards):
            if r == src:
                continue
            for arr in (s.contig_ptr, s.gene_ptr, s.attr_idx):
                if arr.size:
                    reqs.append(dist.isend(torch.from_numpy(numpy.ascontiguousarray(arr)).to(dev), dst=r))
        for q in reqs:
            q.wait()
        return shards[src]
    out = []
    for size, dtype in ((C + 1, torch.int32), (G + 1, ptr_dtype), (nnz, torch.int32)):
        t = torch.empty(size, dtype=dtype, device=dev)
        if size:
            dist.recv(t, src=src)
        out.append(t.cpu().numpy())
    return CsrBatch(out[0], out[1], out[2], name=f"shard{rank}")


def gather_marginals(local, genes_per_rank: Sequence[int]):
    """All-gather the per-gene marginals of every rank (shards padded to the largest one); returns the
    concatenation in rank order as a tensor on ``local``'s device."""
    import torch

    dist = _dist()
    world = dist.get_world_size()
    gmax = max(int(g) for g in genes_per_rank) if len(genes_per_rank) else 0
    padded = torch.zeros(max(gmax, 1), dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([p[: int(g)] for p, g in zip(parts, genes_per_rank)])


def predict_sharded(engine, shard: CsrBatch, *, window: int, step: int = 1, pad: bool = True, device=None):
    """Run this rank's shard through ``engine`` and gather everybody's marginals (rank order = contig order).

    ``engine`` is a ``CRFEngine`` (or anything with its ``marginals_windowed``).  Returns a float64 numpy array
    with the marginals of ALL shards on every rank.
    """
    import torch

    dist = _dist()
    dev = device if device is not None else torch.device("cpu")
    local = engine.marginals_windowed(shard.contig_ptr, shard.gene_ptr, shard.attr_idx, window=window, step=step, pad=pad)
    sizes = [None] * dist.get_world_size()
    dist.all_gather_object(sizes, int(shard.G))
    out = gather_marginals(torch.from_numpy(numpy.ascontiguousarray(local)).to(dev), sizes)
    return out.cpu().numpy()
