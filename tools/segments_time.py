"""Kernel time of gcrf_segments (threshold + segment extraction, N1) on the marginals of config 2 / config 4. B200 only."""
import pathlib
import sys
import time

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy
from gecco_b200 import model_io, synth
from gecco_b200._lib import CRFEngine

w = model_io.load_tsv_model(model_io.bundled_model_dir())
eng = CRFEngine(w, 0)
for name, b in (("config2", synth.config2(len(w.attrs))), ("config4 200k", synth.config4(len(w.attrs), contigs=200_000, mean_domains=6.0))):
    p = eng.marginals_windowed(b.contig_ptr, b.gene_ptr, b.attr_idx)
    ann = (numpy.diff(b.gene_ptr) > 0).astype(numpy.uint8)
    eng.set_timing(True)
    ts = []
    for it in range(8):
        t0 = time.perf_counter()
        seg = eng.segments(b.contig_ptr, p, ann, threshold=0.8, n_cds=3, reset_per_contig=True)
        host = time.perf_counter() - t0
        if it >= 2:
            ts.append((eng.last_kernel_ms(), host * 1e3))
    eng.set_timing(False)
    k = sorted(t[0] for t in ts)[len(ts) // 2]
    h = sorted(t[1] for t in ts)[len(ts) // 2]
    print(f"{name}: G={b.G} clusters={len(seg.contig)} kernels {k*1e3:.1f} us, host call (H2D 18 B/gene + kernels + D2H) {h:.2f} ms, "
          f"{b.G/(k*1e-3)/1e9:.1f} G genes/s on device", flush=True)
