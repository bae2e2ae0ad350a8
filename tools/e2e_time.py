"""End-to-end time of gcrf_marginals_windowed with pinned HOST buffers on config 2, with and without the
overlapped slices (GCRF_HOST_SLICES=1 turns them off).  B200 only."""
import os
import pathlib
import sys
import time

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy
from gecco_b200 import model_io, synth
from gecco_b200._lib import CRFEngine, PinnedArray

w = model_io.load_tsv_model(model_io.bundled_model_dir())
b = synth.config2(len(w.attrs))
eng = CRFEngine(w, 0)
pins = [PinnedArray(a.shape, a.dtype) for a in (b.contig_ptr, b.gene_ptr, b.attr_idx)]
for pin, a in zip(pins, (b.contig_ptr, b.gene_ptr, b.attr_idx)):
    pin.array[...] = a
pout = PinnedArray((b.G,), numpy.float64)
ref = None
for slices in ("1", "2", "4", "6", "8", ""):
    if slices:
        os.environ["GCRF_HOST_SLICES"] = slices
    else:
        os.environ.pop("GCRF_HOST_SLICES", None)
    for _ in range(3):
        eng.marginals_windowed(pins[0].array, pins[1].array, pins[2].array, out=pout.array)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        eng.marginals_windowed(pins[0].array, pins[1].array, pins[2].array, out=pout.array)
        ts.append(time.perf_counter() - t0)
    if ref is None:
        ref = pout.array.copy()
    same = numpy.array_equal(ref, pout.array)
    print(f"slices={slices or 'default':8s} min {min(ts)*1e3:.3f} ms  median {sorted(ts)[5]*1e3:.3f} ms  "
          f"{b.G/min(ts)/1e6:.0f} M genes/s  identical={same}", flush=True)
