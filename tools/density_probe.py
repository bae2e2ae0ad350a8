import sys
sys.path.insert(0, '.')
import torch, numpy
from gecco_b200 import model_io, synth
from gecco_b200._lib import CRFEngine
w = model_io.load_tsv_model(model_io.bundled_model_dir())
dev = torch.device("cuda:0")
eng = CRFEngine(w, 0)
for d in (20, 25, 27, 28, 30, 40, 60):
    b = synth.config2(len(w.attrs), mean_domains=float(d), contigs=5000)
    cp = torch.from_numpy(b.contig_ptr).to(dev); gp = torch.from_numpy(b.gene_ptr).to(dev)
    ai = torch.full((b.nnz + 16,), -1, dtype=torch.int32, device=dev); ai[:b.nnz] = torch.from_numpy(b.attr_idx).to(dev)
    out = torch.empty(b.G, dtype=torch.float64, device=dev)
    eng.set_timing(True)
    ts = []
    for it in range(8):
        eng.marginals_windowed_device(cp.data_ptr(), gp.data_ptr(), ai.data_ptr(), b.C, b.G, b.nnz, out.data_ptr(), window=20)
        if it >= 3: ts.append(eng.last_kernel_ms())
    t = sorted(ts)[2]
    print(f"d={d}: G={b.G} nnz={b.nnz} kernel {t*1e3:.1f} us, {b.G/t/1e6:.1f} G genes/s, {(4*b.nnz+12*b.G)/t/1e6:.0f} GB/s", flush=True)
