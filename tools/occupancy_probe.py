"""Does the streaming kernel gain from more resident CTAs?  (tuning probe, B200 only)

The production geometry (128 threads x 4 CTAs per SM) is capped by shared memory: every CTA stages its own copy of
the 10.6 KB delta table.  With a truncated model (PROBE_ATTRS attributes, ids drawn from them) five or six CTAs fit,
so variants built with -DGCRF_STREAM_MINB=5/6 (tools/build_variant.sh) can be timed at the higher occupancy before
anybody restructures the kernel around a table shared per SM.  Prints per-call / back-to-back kernel times.

    GCRF_LIB_NAME=lib_nt128b5.so PROBE_ATTRS=100 python tools/occupancy_probe.py
"""
import dataclasses
import os
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy
import torch
from gecco_b200 import model_io, synth
from gecco_b200._lib import CRFEngine

A = int(os.environ.get("PROBE_ATTRS", "100"))
w = model_io.load_tsv_model(model_io.bundled_model_dir())
# the A attributes with the largest |delta|: the spread of unary odds stays realistic
order = numpy.argsort(-numpy.abs(w.state_w[:, 1] - w.state_w[:, 0]))[:A]
order.sort()
small = dataclasses.replace(w, attrs=[w.attrs[i] for i in order], state_w=w.state_w[order].copy(), state_mask=w.state_mask[order].copy())
dev = torch.device("cuda:0")
eng = CRFEngine(small, 0)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

for name, d in (("dense d=25", 25.0), ("sparse d=1.4", 1.4)):
    rng = numpy.random.default_rng(2)
    n = numpy.maximum(1, rng.poisson(200, size=10_000))
    # ids with replacement would collapse to < 25 unique ones out of 100: draw from a wide id space, fold afterwards
    b = synth.make_batch(rng, n, d, 2659, 0.05)
    b.attr_idx = numpy.where(b.attr_idx >= 0, b.attr_idx % A, -1).astype(numpy.int32)
    cp = torch.from_numpy(b.contig_ptr).to(dev); gp = torch.from_numpy(b.gene_ptr).to(dev)
    ai = torch.full((b.nnz + 16,), -1, dtype=torch.int32, device=dev); ai[: b.nnz] = torch.from_numpy(b.attr_idx).to(dev)
    out = torch.empty(b.G, dtype=torch.float64, device=dev)

    def call():
        eng.marginals_windowed_device(cp.data_ptr(), gp.data_ptr(), ai.data_ptr(), b.C, b.G, b.nnz, out.data_ptr(), window=20)

    for _ in range(3):
        call()
    eng.set_timing(True)
    ts = []
    for _ in range(20):
        if d < 5:
            flush.zero_()
        call()
        ts.append(eng.last_kernel_ms())
    eng.set_timing(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(20):
        call()
    e1.record(stream)
    torch.cuda.synchronize()
    print(f"{os.environ.get('GCRF_LIB_NAME', 'production'):22s} A={A} {name:13s} G={b.G} nnz={b.nnz}: per-call min {min(ts)*1e3:6.1f} us "
          f"median {sorted(ts)[10]*1e3:6.1f} us, back-to-back {e0.elapsed_time(e1) / 20 * 1e3:6.1f} us", flush=True)
