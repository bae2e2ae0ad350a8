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

__all__ = ["partition_contigs", "scatter_batch", "gather_marginals", "predict_sharded", "DeviceShard", "FusedGather"]


def partition_contigs(contig_ptr: numpy.ndarray, gene_ptr: Optional[numpy.ndarray], n_shards: int, window: int,
                      kappa: float = 2.0, contig_nnz: Optional[numpy.ndarray] = None) -> List[Tuple[int, int]]:
    """Split contigs into ``n_shards`` contiguous ranges of roughly equal cost.

    Cost of a contig = its attribute ids (the bytes streamed from HBM) + ``kappa * windows * window`` (the
    dynamic-programming steps it triggers).  Contiguous ranges keep every shard a plain slice of the CSR arrays.
    Returns ``[(c_begin, c_end), ...]``; trailing shards may be empty when there are fewer contigs than shards.
    ``contig_nnz`` (ids per contig, exact or expected) replaces ``gene_ptr`` when the row pointers are not at hand.
    """
    contig_ptr = numpy.asarray(contig_ptr, dtype=numpy.int64)
    C = len(contig_ptr) - 1
    if n_shards <= 0:
        raise ValueError("n_shards must be positive")
    if C == 0:
        return [(0, 0)] * n_shards
    n = numpy.diff(contig_ptr)
    if contig_nnz is not None:
        nnz = numpy.asarray(contig_nnz, dtype=numpy.float64)
    else:
        gene_ptr = numpy.asarray(gene_ptr, dtype=numpy.int64)
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
    concatenation in rank order as a tensor on ``local``'s device.  One collective on the caller's stream
    (``all_gather_into_tensor``: NCCL on GPUs, gloo on CPU tensors); nothing touches the host."""
    import torch

    dist = _dist()
    world = dist.get_world_size()
    gmax = max(int(g) for g in genes_per_rank) if len(genes_per_rank) else 0
    gmax = max(gmax, 1)
    if local.numel() == gmax:
        padded = local
    else:
        padded = torch.zeros(gmax, dtype=local.dtype, device=local.device)
        padded[: local.numel()] = local
    gathered = torch.empty(world * gmax, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded)
    if all(int(g) == gmax for g in genes_per_rank):
        return gathered
    return torch.cat([gathered[r * gmax: r * gmax + int(g)] for r, g in enumerate(genes_per_rank)])


class DeviceShard:
    """This rank's CSR shard resident in HBM (uploaded once; ``predict_sharded`` then runs kernel + gather only)."""

    def __init__(self, shard: CsrBatch, device):
        import torch

        self.C, self.G, self.nnz = shard.C, shard.G, shard.nnz
        self.ptr64 = shard.gene_ptr.dtype == numpy.int64
        self.contig_ptr = torch.from_numpy(numpy.ascontiguousarray(shard.contig_ptr)).to(device)
        self.gene_ptr = torch.from_numpy(numpy.ascontiguousarray(shard.gene_ptr)).to(device)
        # the streaming kernel's bulk copies read whole 16-byte units: keep slack behind the last id
        self.attr_idx = torch.full((shard.nnz + 16,), -1, dtype=torch.int32, device=device)
        self.attr_idx[: shard.nnz] = torch.from_numpy(numpy.ascontiguousarray(shard.attr_idx)).to(device)


def predict_sharded(engine, shard, *, window: int, step: int = 1, pad: bool = True, device=None,
                    genes_per_rank: Optional[Sequence[int]] = None, to_host: bool = True):
    """Run this rank's shard through ``engine`` and gather everybody's marginals (rank order = contig order).

    CUDA device (``device.type == "cuda"``, NCCL process group): ``shard`` is a ``CsrBatch`` (uploaded here) or a
    ``DeviceShard`` (already resident); the kernel writes straight into the padded send buffer of ONE
    ``all_gather_into_tensor`` on the same stream — no host round trip between kernel and collective.  Returns the
    float64 marginals of ALL shards on every rank: a device tensor when ``to_host`` is false, else a numpy array.
    CPU (gloo, the test-suite): ``engine.marginals_windowed`` on host arrays, gather on CPU tensors.
    """
    import torch

    dist = _dist()
    world = dist.get_world_size()
    if genes_per_rank is None:
        genes_per_rank = [None] * world
        dist.all_gather_object(genes_per_rank, int(shard.G))
    genes_per_rank = [int(g) for g in genes_per_rank]
    gmax = max(max(genes_per_rank), 1)
    if device is not None and torch.device(device).type == "cuda":
        dev = torch.device(device)
        ds = shard if isinstance(shard, DeviceShard) else DeviceShard(shard, dev)
        current = torch.cuda.current_stream(dev)
        # the engine runs on the stream torch's collectives order themselves against; stream handle 0 (the legacy
        # default stream) would mean "the engine's own stream", so work moves to a side stream in that case
        side = torch.cuda.Stream(dev) if current.cuda_stream == 0 else current
        side.wait_stream(current)
        with torch.cuda.stream(side):
            engine.set_stream(side.cuda_stream)
            local = torch.empty(gmax, dtype=torch.float64, device=dev)
            if ds.G < gmax:
                local[ds.G:].zero_()
            if ds.G:
                engine.marginals_windowed_device(ds.contig_ptr.data_ptr(), ds.gene_ptr.data_ptr(), ds.attr_idx.data_ptr(),
                                                 ds.C, ds.G, ds.nnz, local.data_ptr(), window=window, step=step, pad=pad,
                                                 ptr64=ds.ptr64)
            out = gather_marginals(local, genes_per_rank)
            engine.set_stream(None)
        current.wait_stream(side)
        if not to_host:
            return out
        return out.cpu().numpy()
    local = engine.marginals_windowed(shard.contig_ptr, shard.gene_ptr, shard.attr_idx, window=window, step=step, pad=pad)
    dev = device if device is not None else torch.device("cpu")
    out = gather_marginals(torch.from_numpy(numpy.ascontiguousarray(local)).to(dev), genes_per_rank)
    return out.cpu().numpy() if to_host else out


class FusedGather:
    """The gather of a contig-sharded batch fused into the marginal kernel (``gcrf_marginals_windowed_peers``).

    Every rank owns a symmetric buffer of ALL genes' marginals (``torch.distributed._symmetric_memory``: one allocation per
    GPU, mapped into every process of the node over NVLink / NVSwitch).  ``predict`` runs the windowed kernel on this
    rank's shard and the kernel itself stores every result into all ranks' buffers at the shard's gene offset — peer
    stores, or ONE ``multimem.st`` to the NVLS multicast address when the fabric offers it — so the transfer overlaps the
    computation tile by tile and no collective follows, only a device-side barrier before anybody reads.

    Needs the streaming kernel's peer-store variant (window 5 or 20, FP32 arithmetic); callers fall back to ``predict_sharded`` otherwise.
    """

    def __init__(self, genes_per_rank: Sequence[int], device, *, f32: bool = False, multicast: Optional[bool] = None, group=None):
        import torch
        import torch.distributed._symmetric_memory as symm_mem

        dist = _dist()
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.sizes = [int(g) for g in genes_per_rank]
        if len(self.sizes) != self.world:
            raise ValueError("one gene count per rank")
        self.offsets = numpy.concatenate([[0], numpy.cumsum(self.sizes)]).astype(numpy.int64)
        self.total = int(self.offsets[-1])
        self.dtype = torch.float32 if f32 else torch.float64
        self.f32 = f32
        self.buffer = symm_mem.empty(max(self.total, 1), dtype=self.dtype, device=device)
        self.handle = symm_mem.rendezvous(self.buffer, self.group.group_name)
        can_mc = bool(getattr(self.handle, "has_multicast_support", False)) and int(getattr(self.handle, "multicast_ptr", 0) or 0) != 0
        self.multicast = can_mc if multicast is None else (bool(multicast) and can_mc)
        if self.world > 8 and not self.multicast:
            raise ValueError("at most 8 peer output arrays per call without NVLS multicast")

    def predict(self, engine, shard: "DeviceShard", *, window: int, step: int = 1, pad: bool = True, barrier_before: bool = True):
        """Enqueue kernel + barrier on the current stream; returns the symmetric buffer (all genes, this rank's copy)."""
        import torch

        stream = torch.cuda.current_stream(self.buffer.device)
        if stream.cuda_stream == 0:
            raise RuntimeError("run FusedGather.predict on a non-default CUDA stream (torch.cuda.stream(...))")
        engine.set_stream(stream.cuda_stream)
        if barrier_before:
            self.handle.barrier(channel=0)  # nobody still reads the previous result while peers overwrite it
        if shard.G:
            peers = [int(self.handle.multicast_ptr)] if self.multicast else [int(p) for p in self.handle.buffer_ptrs]
            engine.marginals_windowed_peers(shard.contig_ptr.data_ptr(), shard.gene_ptr.data_ptr(), shard.attr_idx.data_ptr(),
                                            shard.C, shard.G, shard.nnz, None, peers, int(self.offsets[self.rank]),
                                            window=window, step=step, pad=pad, f32=self.f32, ptr64=shard.ptr64,
                                            multicast=self.multicast)
        self.handle.barrier(channel=1)  # every rank's kernel has finished storing into this rank's buffer
        return self.buffer[: self.total]
