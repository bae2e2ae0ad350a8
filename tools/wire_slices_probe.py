"""e2e of gcrf_marginals_windowed_wire against the number of slices (B200 only).

    python tools/wire_slices_probe.py [sparse]
"""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy

from gecco_b200 import model_io, synth
from gecco_b200._lib import CRFEngine, PinnedArray, WireBatch

weights = model_io.load_tsv_model(model_io.bundled_model_dir())
sparse = len(sys.argv) > 1 and sys.argv[1] == "sparse"
batch = synth.config2(len(weights.attrs), seed=1, contigs=10000, mean_domains=1.4 if sparse else 25.0)
engine = CRFEngine(weights, device=0)
ref = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, window=20, step=1, pad=True)
pout = PinnedArray((batch.G,), numpy.float64)
for n in ("1", "2", "3", "4", "6", "8", ""):
    if n:
        os.environ["GCRF_WIRE_SLICES"] = n
    else:
        os.environ.pop("GCRF_WIRE_SLICES", None)
    wire = WireBatch(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, len(weights.attrs))
    call = lambda: engine.marginals_windowed_wire(wire, window=20, step=1, pad=True, out=pout.array)
    call(); call()
    t0 = time.perf_counter()
    for _ in range(10):
        call()
    ms = (time.perf_counter() - t0) / 10 * 1e3
    same = bool(numpy.array_equal(pout.array, ref))
    engine.set_timing(True)
    call()
    kms = engine.last_kernel_ms()
    engine.set_timing(False)
    print(f"slices {n or 'default':>7}: {ms:.3f} ms per call = {batch.G / ms / 1e3:.0f} M genes/s, {wire.nbytes / 1e6:.1f} MB in, "
          f"identical {same}, kernels alone (one slice) {kms:.3f} ms")
    wire.close()
