"""Kernel-only throughput + parity spot checks on the BASELINE.json config shapes.

    python tools/bench_configs.py [--scale4 0.2]

Config 2 is bench.py's headline; this script adds the sparse (config 3-like), metagenome (config 4, scaled) and
long-contig (config 5) shapes so that no regime is pathological.  Prints one JSON line per config.
"""
import argparse
import json
import pathlib
import sys
import time

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy
import torch

from gecco_b200 import model_io, synth
from gecco_b200._lib import CRFEngine
from oracle import crf_oracle


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale4", type=float, default=0.2, help="fraction of config 4's 1M contigs")
    args = ap.parse_args()
    w = model_io.load_tsv_model(model_io.bundled_model_dir())
    A = len(w.attrs)
    dev = torch.device("cuda:0")
    eng = CRFEngine(w, 0)
    eng.set_timing(True)
    configs = [
        ("config2 dense 10k contigs", lambda: synth.config2(A)),
        ("config2 shape, 1.4 domains/gene (real-data density)", lambda: synth.config2(A, mean_domains=1.4)),
        ("config3 E.coli-like 1 contig x 4300 genes", lambda: synth.config3_ecoli_like(A)),
        (f"config4 metagenome {int(1e6 * args.scale4)} contigs, 25 domains", lambda: synth.config4(A, contigs=int(1e6 * args.scale4))),
        (f"config4 metagenome {int(1e6 * args.scale4)} contigs, 1.4 domains", lambda: synth.config4(A, contigs=int(1e6 * args.scale4), mean_domains=1.4)),
        ("config5 100 contigs x 5000 genes", lambda: synth.config5(A)),
    ]
    for name, make in configs:
        t0 = time.time()
        b = make()
        cp = torch.from_numpy(b.contig_ptr).to(dev)
        gp = torch.from_numpy(b.gene_ptr).to(dev)
        ai = torch.from_numpy(b.attr_idx).to(dev)
        out = torch.empty(b.G, dtype=torch.float64, device=dev)
        ptr64 = b.gene_ptr.dtype == numpy.int64
        ts = []
        for it in range(13):
            eng.marginals_windowed_device(cp.data_ptr(), gp.data_ptr(), ai.data_ptr(), b.C, b.G, b.nnz, out.data_ptr(), ptr64=ptr64)
            if it >= 3:
                ts.append(eng.last_kernel_ms())
        ms = sorted(ts)[len(ts) // 2]
        # parity on the first contigs (about 50k genes)
        c1 = int(numpy.searchsorted(b.contig_ptr, 50_000))
        c1 = max(1, min(b.C, c1))
        sub = b.slice_contigs(0, c1)
        want, _ = crf_oracle.marginals_windowed(w.state_w, w.trans_w, 1, sub.contig_ptr, sub.gene_ptr, sub.attr_idx, 20, 1, True, nthreads=8)
        got = out[:sub.G].cpu().numpy()
        err = float(numpy.nanmax(numpy.abs(got - want))) if sub.G else 0.0
        algo = synth.algorithmic_bytes(b.C, b.G, b.nnz)
        print(json.dumps({"config": name, "contigs": b.C, "genes": b.G, "nnz": b.nnz, "windows": b.windows(20),
                          "kernel_ms": ms, "genes_per_s": b.G / (ms * 1e-3), "algorithmic_GBps": algo / (ms * 1e-3) / 1e9,
                          "parity_max_abs_err": err, "parity_genes": sub.G, "gen_s": round(time.time() - t0, 1)}), flush=True)
        if name.startswith("config5"):
            # (ii) the deep-chain primitive: one 5,000-gene chain per contig (gcrf_marginals_chain, f64 2x2 scan)
            ts = []
            for it in range(13):
                eng.marginals_chain_device(cp.data_ptr(), gp.data_ptr(), ai.data_ptr(), b.C, b.G, b.nnz, out.data_ptr(), ptr64=ptr64)
                if it >= 3:
                    ts.append(eng.last_kernel_ms())
            ms = sorted(ts)[len(ts) // 2]
            sub = b.slice_contigs(0, 4)
            got = out[:sub.G].cpu().numpy()
            err = 0.0
            for c in range(sub.C):
                g0, g1 = int(sub.contig_ptr[c]), int(sub.contig_ptr[c + 1])
                want = crf_oracle.chain_marginals(w.state_w, w.trans_w, sub.gene_ptr, sub.attr_idx, g0, g1)[:, 1]
                err = max(err, float(numpy.abs(got[g0:g1] - want).max()))
            print(json.dumps({"config": name + " — whole-contig chains (gcrf_marginals_chain)", "contigs": b.C, "genes": b.G,
                              "nnz": b.nnz, "kernel_ms": ms, "genes_per_s": b.G / (ms * 1e-3),
                              "algorithmic_GBps": algo / (ms * 1e-3) / 1e9, "parity_max_abs_err": err, "parity_genes": sub.G}),
                  flush=True)
        del cp, gp, ai, out


if __name__ == "__main__":
    main()
