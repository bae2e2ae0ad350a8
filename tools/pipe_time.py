"""Kernel-only timing of the W=20 device paths on synthetic batches (tuning aid, B200 only).

    python tools/pipe_time.py [config2|sparse|config4|config5 ...]

For every requested shape: the fused kernel, then the two-kernel pipeline over a grid of GCRF_PIPE_CTAS x
GCRF_PIPE_STAGES (+ GCRF_PIPE_SERIAL=1 for the overlap A/B).  Prints per-call time (events inside the ABI) and
back-to-back time (20 calls between two stream events), and the max |dp| between the paths.
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


def measure(b, env):
    for k in ("GCRF_PATH", "GCRF_PIPE_CTAS", "GCRF_PIPE_STAGES", "GCRF_PIPE_SERIAL"):
        os.environ.pop(k, None)
    os.environ.update(env)
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
    lo, med, b2b, ref = measure(b, {"GCRF_PATH": "fused"})
    print(f"fused                      per-call min {lo*1e3:7.1f} us  median {med*1e3:7.1f} us  back-to-back {b2b*1e3:7.1f} us", flush=True)
    grid = [(c, s, 0) for c in (1, 2, 3, 4) for s in (1, 2)] + [(2, 2, 1), (3, 1, 1)]
    for ctas, stages, serial in grid:
        env = {"GCRF_PATH": "pipeline", "GCRF_PIPE_CTAS": str(ctas), "GCRF_PIPE_STAGES": str(stages), "GCRF_PIPE_SERIAL": str(serial)}
        lo, med, b2b, got = measure(b, env)
        ok = numpy.array_equal(numpy.isnan(got), numpy.isnan(ref))
        err = float(numpy.nanmax(numpy.abs(got - ref))) if ok else float("inf")
        print(f"pipeline ctas={ctas} stages={stages} serial={serial}  per-call min {lo*1e3:7.1f} us  median {med*1e3:7.1f} us  "
              f"back-to-back {b2b*1e3:7.1f} us  max|dp vs fused| {err:.1e}", flush=True)
