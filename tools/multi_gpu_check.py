"""Multi-GPU path on real GPUs: rank 0 holds a batch, NCCL scatters contig-aligned CSR shards,
every rank runs its shard on its own B200, NCCL all-gathers the per-gene marginals; rank 0 checks the result against
the single-GPU result (bit-identical: contigs are independent) and the CPU oracle (<= 1e-5), and prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py [--contigs 4000]
"""
import argparse
import json
import os
import pathlib
import sys
import time

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy
import torch
import torch.distributed as dist

from gecco_b200 import model_io, sharding, synth
from gecco_b200._lib import CRFEngine


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--contigs", type=int, default=4000)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    w = model_io.load_tsv_model(model_io.bundled_model_dir())
    eng = CRFEngine(w, device=local)
    batch = synth.config4(len(w.attrs), contigs=args.contigs, mean_domains=25.0) if rank == 0 else None

    def sync():
        torch.cuda.synchronize(dev)
        dist.barrier()

    sharding.scatter_batch(batch, 20, src=0, device=dev)  # warm-up: NCCL channels, allocator
    sync()
    t0 = time.perf_counter()
    shard = sharding.scatter_batch(batch, 20, src=0, device=dev)
    sync()
    t1 = time.perf_counter()
    out = sharding.predict_sharded(eng, shard, window=20, step=1, pad=True, device=dev)
    sync()
    t2 = time.perf_counter()
    if rank == 0:
        from oracle import crf_oracle

        single = eng.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx)
        sub = batch.slice_contigs(0, min(batch.C, 400))
        want, _ = crf_oracle.marginals_windowed(w.state_w, w.trans_w, w.label_id("1"), sub.contig_ptr, sub.gene_ptr, sub.attr_idx,
                                                20, 1, True, nthreads=8)
        print(json.dumps({
            "world": world, "contigs": batch.C, "genes": batch.G, "nnz": batch.nnz, "shard0_genes": shard.G,
            "scatter_ms": 1e3 * (t1 - t0), "infer_and_gather_ms": 1e3 * (t2 - t1),
            "identical_to_single_gpu": bool(numpy.array_equal(out, single, equal_nan=True)),
            "max_abs_err_vs_oracle": float(numpy.nanmax(numpy.abs(out[:sub.G] - want))),
        }))
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
