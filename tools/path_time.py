"""Kernel-only timing of the windowed device path on the synthetic shapes (tuning aid, B200 only).

    python tools/path_time.py [config2|sparse|config4|config5 ...]

Per shape: per-call time (events inside the ABI around the kernel) and back-to-back time (20 calls between two
stream events), for the fused streaming kernel and the generic kernel.
"""
import os
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy
import torch
from gecco_b200 import model_io, synth
from gecco_b200._lib import CRFEngine

w = model_io.load_tsv_model(model_io.bundled_model_dir())
A = len(w.attrs)
shapes = {
    "config2": lambda: synth.config2(A),
    "sparse": lambda: synth.config2(A, mean_domains=1.4),
    "config4": lambda: synth.config4(A, contigs=200_000),
    "config5": lambda: synth.config5(A),
}
dev = torch.device("cuda:0")
eng = CRFEngine(w, 0)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)


def measure(b, path):
    os.environ["GCRF_FORCE_GENERIC"] = "1" if path == "generic" else "0"
    cp = torch.from_numpy(b.contig_ptr).to(dev); gp = torch.from_numpy(b.gene_ptr).to(dev); ai = torch.from_numpy(b.attr_idx).to(dev)
    out = torch.full((b.G,), -1.0, dtype=torch.float64, device=dev)

    def call():
        eng.marginals_windowed_device(cp.data_ptr(), gp.data_ptr(), ai.data_ptr(), b.C, b.G, b.nnz, out.data_ptr())

    for _ in range(3):
        call()
    eng.set_timing(True)
    ts = []
    for _ in range(20):
        call()
        ts.append(eng.last_kernel_ms())
    eng.set_timing(False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(20):
        call()
    e1.record(stream)
    torch.cuda.synchronize()
    return min(ts), sorted(ts)[len(ts) // 2], e0.elapsed_time(e1) / 20, out.cpu().numpy()


for name in (sys.argv[1:] or ["config2"]):
    b = shapes[name]()
    print(f"== {name}: C={b.C} G={b.G} nnz={b.nnz}", flush=True)
    ref = None
    for label in ("fused", "generic"):
        lo, med, b2b, got = measure(b, label)
        if ref is None:
            ref = got
        err = float(numpy.nanmax(numpy.abs(got - ref)))
        print(f"{label:8s} per-call min {lo*1e3:7.1f} us  median {med*1e3:7.1f} us  back-to-back {b2b*1e3:7.1f} us  "
              f"max|dp vs fused| {err:.1e}", flush=True)
