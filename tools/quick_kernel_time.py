import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy, torch
from gecco_b200 import model_io, synth
from gecco_b200._lib import CRFEngine
w = model_io.load_tsv_model(model_io.bundled_model_dir())
b = synth.config2(len(w.attrs))
dev = torch.device('cuda:0')
eng = CRFEngine(w, 0)
cp = torch.from_numpy(b.contig_ptr).to(dev); gp = torch.from_numpy(b.gene_ptr).to(dev); ai = torch.from_numpy(b.attr_idx).to(dev)
out = torch.empty(b.G, dtype=torch.float64, device=dev)
def run(n):
    ts=[]
    for _ in range(n):
        eng.marginals_windowed_device(cp.data_ptr(), gp.data_ptr(), ai.data_ptr(), b.C, b.G, b.nnz, out.data_ptr())
        ts.append(eng.last_kernel_ms())
    return ts
run(3)
os.environ['GCRF_PHASE_PROFILE']='1'
run(2)
os.environ['GCRF_PHASE_PROFILE']='0'
ts = run(20)
print("kernel ms: min %.4f median %.4f" % (min(ts), sorted(ts)[len(ts)//2]))
