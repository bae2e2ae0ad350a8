import os, sys
sys.path.insert(0, "/root/repo")
import numpy
from gecco_b200 import model_io, synth
from gecco_b200._lib import CRFEngine, PinnedArray, WireBatch
weights = model_io.load_tsv_model(model_io.bundled_model_dir())
batch = synth.config2(len(weights.attrs), seed=1, contigs=10000)
engine = CRFEngine(weights, device=0)
pout = PinnedArray((batch.G,), numpy.float64)
wire = WireBatch(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, len(weights.attrs))
for _ in range(3):
    engine.marginals_windowed_wire(wire, window=20, step=1, pad=True, out=pout.array)
    print("--", file=sys.stderr)
