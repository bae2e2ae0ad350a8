"""A few calls of the windowed device path on one synthetic shape (for ncu captures).

    python tools/run_once.py [config2|sparse|config4|config5] [calls] [f64]
"""
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from gecco_b200 import model_io, synth
from gecco_b200._lib import CRFEngine

w = model_io.load_tsv_model(model_io.bundled_model_dir())
shape = sys.argv[1] if len(sys.argv) > 1 else "config2"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
flags = 0x80 if len(sys.argv) > 3 and sys.argv[3] == "f64" else 0  # GCRF_FLAG_F64
A = len(w.attrs)
b = {"config2": lambda: synth.config2(A), "sparse": lambda: synth.config2(A, mean_domains=1.4),
     "config4": lambda: synth.config4(A, contigs=200_000), "config5": lambda: synth.config5(A)}[shape]()
dev = torch.device("cuda:0")
eng = CRFEngine(w, 0)
cp = torch.from_numpy(b.contig_ptr).to(dev); gp = torch.from_numpy(b.gene_ptr).to(dev); ai = torch.from_numpy(b.attr_idx).to(dev)
out = torch.empty(b.G, dtype=torch.float64, device=dev)
for _ in range(calls):
    eng.marginals_windowed_device(cp.data_ptr(), gp.data_ptr(), ai.data_ptr(), b.C, b.G, b.nnz, out.data_ptr(), flags=flags)
eng.synchronize()
print("done", shape, b.G)
