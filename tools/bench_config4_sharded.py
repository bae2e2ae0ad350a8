"""BASELINE config 4 (metagenome: lognormal contig lengths, ~31 % shorter than the window) sharded by contig over
the ranks of a torchrun job: strong scaling, no data-path collective, NCCL only to collect the
timings.  Every rank builds the same synthetic batch from the seed, keeps its cost-balanced contiguous slice
(gecco_b200.sharding.partition_contigs), runs it device-resident and reports max-over-ranks time.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 \
        tools/bench_config4_sharded.py [--scale 0.2]
"""
import argparse
import json
import os
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy
import torch
import torch.distributed as dist

from gecco_b200 import model_io, sharding, synth
from gecco_b200._lib import CRFEngine


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.2, help="fraction of config 4's 1M contigs")
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = model_io.load_tsv_model(model_io.bundled_model_dir())
    batch = synth.config4(len(w.attrs), contigs=int(1e6 * args.scale))
    c0, c1 = sharding.partition_contigs(batch.contig_ptr, batch.gene_ptr, world, 20)[rank]
    shard = batch.slice_contigs(c0, c1)
    eng = CRFEngine(w, device=local)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    cp, gp, ai = (torch.from_numpy(a).to(dev) for a in (shard.contig_ptr, shard.gene_ptr, shard.attr_idx))
    out = torch.empty(shard.G, dtype=torch.float64, device=dev)
    ptr64 = shard.gene_ptr.dtype == numpy.int64

    def step():
        eng.marginals_windowed_device(cp.data_ptr(), gp.data_ptr(), ai.data_ptr(), shard.C, shard.G, shard.nnz, out.data_ptr(), ptr64=ptr64)

    for _ in range(3):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps, float(shard.G), float(shard.nnz)], dtype=torch.float64, device=dev)
    if world > 1:
        parts = [torch.empty_like(ms) for _ in range(world)]
        dist.all_gather(parts, ms)
    else:
        parts = [ms]
    if rank == 0:
        t = max(float(p[0]) for p in parts)
        print(json.dumps({"workload": f"config 4, {batch.C} contigs / {batch.G} genes / {batch.nnz} ids, contig-sharded x{world}",
                          "n_gpus": world, "ms_per_step_max_over_ranks": t, "genes_per_s": batch.G / (t * 1e-3),
                          "per_rank": [{"ms": float(p[0]), "genes": int(p[1]), "nnz": int(p[2])} for p in parts]}))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
