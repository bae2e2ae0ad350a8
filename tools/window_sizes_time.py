import sys, pathlib
sys.path.insert(0, '.')
import torch, numpy
from gecco_b200 import model_io, synth
from gecco_b200._lib import CRFEngine
w = model_io.load_tsv_model(model_io.bundled_model_dir())
dev = torch.device("cuda:0")
eng = CRFEngine(w, 0)
for name, b in (("config2", synth.config2(len(w.attrs))), ("sparse", synth.config2(len(w.attrs), mean_domains=1.4))):
    cp = torch.from_numpy(b.contig_ptr).to(dev); gp = torch.from_numpy(b.gene_ptr).to(dev); ai = torch.from_numpy(b.attr_idx).to(dev)
    out = torch.empty(b.G, dtype=torch.float64, device=dev)
    for W in (5, 10, 20, 25, 30, 40, 50, 64):
        eng.set_timing(True)
        ts = []
        for it in range(13):
            eng.marginals_windowed_device(cp.data_ptr(), gp.data_ptr(), ai.data_ptr(), b.C, b.G, b.nnz, out.data_ptr(), window=W)
            if it >= 3: ts.append(eng.last_kernel_ms())
        print(name, "W=%d" % W, "kernel us %.1f" % (1e3 * sorted(ts)[5]), "G genes/s %.1f" % (b.G / sorted(ts)[5] / 1e6), flush=True)
