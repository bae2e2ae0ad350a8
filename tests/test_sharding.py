"""Contig sharding: the partitioner, and the scatter -> per-rank inference -> gather path on gloo (2 ranks)."""
import os
import socket

import numpy
import pytest

from gecco_b200 import sharding, synth


def test_partition_is_contiguous_balanced_and_complete(weights):
    batch = synth.config4(len(weights.attrs), contigs=3000, mean_domains=5.0)
    for n in (1, 2, 3, 8):
        parts = sharding.partition_contigs(batch.contig_ptr, batch.gene_ptr, n, 20)
        assert len(parts) == n and parts[0][0] == 0 and parts[-1][1] == batch.C
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        cost = []
        for c0, c1 in parts:
            s = batch.slice_contigs(c0, c1)
            cost.append(s.nnz + 2.0 * 20 * s.windows(20))
        assert max(cost) <= 1.05 * (sum(cost) / n) + 20000


def test_partition_edge_cases():
    cp = numpy.array([0, 3], dtype=numpy.int32)
    gp = numpy.array([0, 1, 1, 2], dtype=numpy.int32)
    assert sharding.partition_contigs(cp, gp, 4, 20) == [(0, 0), (0, 0), (0, 0), (0, 1)] or \
        sum(b - a for a, b in sharding.partition_contigs(cp, gp, 4, 20)) == 1
    assert sharding.partition_contigs(numpy.zeros(1), numpy.zeros(1), 2, 20) == [(0, 0), (0, 0)]
    with pytest.raises(ValueError):
        sharding.partition_contigs(cp, gp, 0, 20)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, result_dir):
    import sys
    import pathlib

    root = pathlib.Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root))
    sys.path.insert(0, str(root / "tests"))
    import torch.distributed as dist
    from gecco_b200 import model_io
    from test_crf_dropin import OracleEngine

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    weights = model_io.load_tsv_model(model_io.bundled_model_dir())
    batch = synth.ragged_edge_cases(len(weights.attrs)) if rank == 0 else None
    shard = sharding.scatter_batch(batch, 20, src=0)
    out = sharding.predict_sharded(OracleEngine(weights), shard, window=20, step=1, pad=True)
    numpy.save(os.path.join(result_dir, f"rank{rank}.npy"), out)
    numpy.save(os.path.join(result_dir, f"genes{rank}.npy"), numpy.array([shard.G, shard.C]))
    dist.destroy_process_group()


def test_scatter_infer_gather_two_ranks_gloo(tmp_path, weights):
    import torch.multiprocessing as mp
    from oracle import crf_oracle

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    batch = synth.ragged_edge_cases(len(weights.attrs))
    want, _ = crf_oracle.marginals_windowed(weights.state_w, weights.trans_w, 1, batch.contig_ptr, batch.gene_ptr,
                                            batch.attr_idx, 20, 1, True)
    r0, r1 = numpy.load(tmp_path / "rank0.npy"), numpy.load(tmp_path / "rank1.npy")
    assert numpy.array_equal(r0, r1) and numpy.array_equal(r0, want)
    g0, g1 = numpy.load(tmp_path / "genes0.npy"), numpy.load(tmp_path / "genes1.npy")
    assert g0[0] + g1[0] == batch.G and g0[1] + g1[1] == batch.C and min(g0[0], g1[0]) > 0
