"""Kernel time of the wire call against the batch size (B200 only): python tools/wire_small_probe.py"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy

from gecco_b200 import model_io, synth
from gecco_b200._lib import CRFEngine, PinnedArray, WireBatch

os.environ["GCRF_WIRE_SLICES"] = "1"
weights = model_io.load_tsv_model(model_io.bundled_model_dir())
engine = CRFEngine(weights, device=0)
for contigs in (10000, 5000, 2500, 1250, 700, 100):
    batch = synth.config2(len(weights.attrs), seed=1, contigs=contigs)
    pout = PinnedArray((batch.G,), numpy.float64)
    wire = WireBatch(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, len(weights.attrs))
    engine.set_timing(True)
    ks = []
    for _ in range(5):
        engine.marginals_windowed_wire(wire, window=20, step=1, pad=True, out=pout.array)
        ks.append(engine.last_kernel_ms())
    engine.set_timing(False)
    print(f"{batch.G} genes: kernels of the wire call {min(ks) * 1e3:.1f} us (median {sorted(ks)[2] * 1e3:.1f})")
    wire.close()
