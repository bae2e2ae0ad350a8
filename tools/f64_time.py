import sys, os
sys.path.insert(0, "/root/repo")
import numpy
from gecco_b200 import model_io, synth
from gecco_b200._lib import CRFEngine
import torch
w = model_io.load_tsv_model(model_io.bundled_model_dir())
b = synth.config2(len(w.attrs), seed=1, contigs=10000)
eng = CRFEngine(w, device=0)
eng.set_timing(True)
ts = []
for _ in range(6):
    out = eng.marginals_windowed(b.contig_ptr, b.gene_ptr, b.attr_idx, f64_arith=True)
    ts.append(eng.last_kernel_ms())
import hashlib
print(os.environ.get("GCRF_LIB_NAME", "default"), "f64 kernels ms", min(ts), "sha", hashlib.sha256(out.tobytes()).hexdigest()[:16])
