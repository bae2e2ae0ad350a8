"""Kernel-only timing of the windowed path on the config-2 batch (tuning aid).
   GCRF_PHASE_PROFILE / GCRF_DEBUG_SKIP / GCRF_STREAM_THREADS are read by the library."""
import os
import sys

os.environ.setdefault("GCRF_TUNING_LIB", "1")
import pathlib

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from gecco_b200 import model_io, synth
from gecco_b200._lib import CRFEngine

w = model_io.load_tsv_model(model_io.bundled_model_dir())
mean_domains = float(os.environ.get("QK_DOMAINS", "25"))
window = int(os.environ.get("QK_WINDOW", "20"))
b = synth.config2(len(w.attrs), mean_domains=mean_domains, contigs=int(os.environ.get("QK_CONTIGS", "10000")))
dev = torch.device("cuda:0")
eng = CRFEngine(w, 0)
eng.set_timing(True)
cp = torch.from_numpy(b.contig_ptr).to(dev); gp = torch.from_numpy(b.gene_ptr).to(dev); ai = torch.from_numpy(b.attr_idx).to(dev)
out = torch.empty(b.G, dtype=torch.float64, device=dev)


def run(n):
    ts = []
    for _ in range(n):
        eng.marginals_windowed_device(cp.data_ptr(), gp.data_ptr(), ai.data_ptr(), b.C, b.G, b.nnz, out.data_ptr(), window=window)
        ts.append(eng.last_kernel_ms())
    return ts


run(3)
if os.environ.get("QK_PHASES", "1") == "1":
    os.environ["GCRF_PHASE_PROFILE"] = "1"
    run(1)
    os.environ["GCRF_PHASE_PROFILE"] = "0"
for skip in os.environ.get("QK_SKIPS", "0").split(","):
    os.environ["GCRF_DEBUG_SKIP"] = skip
    run(2)
    ts = run(20)
    print("skip=%s d=%g kernel ms: min %.4f median %.4f" % (skip, mean_domains, min(ts), sorted(ts)[len(ts) // 2]))

if os.environ.get("QK_PHASE_SKIPS"):
    for skip in os.environ["QK_PHASE_SKIPS"].split(","):
        os.environ["GCRF_DEBUG_SKIP"] = skip
        os.environ["GCRF_PHASE_PROFILE"] = "1"
        print("phases with skip=" + skip, flush=True)
        run(1)
        os.environ["GCRF_PHASE_PROFILE"] = "0"
