"""Where the time of one host-buffer call goes (B200 only): raw pinned copies of the same sizes timed with CUDA events,
the kernels of the call (gcrf_model_set_timing), and the wall time of the call itself.

    python tools/e2e_probe.py
"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy
import torch

from gecco_b200 import model_io, synth
from gecco_b200._lib import CRFEngine, PinnedArray, WireBatch

weights = model_io.load_tsv_model(model_io.bundled_model_dir())
batch = synth.config2(len(weights.attrs), seed=1, contigs=10000)
engine = CRFEngine(weights, device=0)
wire = WireBatch(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, len(weights.attrs))
pout = PinnedArray((batch.G,), numpy.float64)
dev = torch.device("cuda:0")


def ev_ms(fn, n=10):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for nbytes, label in ((wire.nbytes, "H2D wire block"), (batch.G * 8, "D2H float64 marginals")):
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    devt = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    if label.startswith("H2D"):
        ms = ev_ms(lambda: devt.copy_(host, non_blocking=True))
    else:
        ms = ev_ms(lambda: host.copy_(devt, non_blocking=True))
    print(f"{label}: {nbytes / 1e6:.1f} MB raw copy {ms:.3f} ms = {nbytes / ms / 1e6:.1f} GB/s")


def wall(fn, n=10):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e3


call = lambda: engine.marginals_windowed_wire(wire, window=20, step=1, pad=True, out=pout.array)
print(f"wire call wall {wall(call):.3f} ms")
engine.set_timing(True)
call()
print(f"kernels of the wire call (decode + marginals) {engine.last_kernel_ms():.3f} ms")
engine.set_timing(False)
pins = [PinnedArray(a.shape, a.dtype) for a in (batch.contig_ptr, batch.gene_ptr, batch.attr_idx)]
for pin, a in zip(pins, (batch.contig_ptr, batch.gene_ptr, batch.attr_idx)):
    pin.array[...] = a
csr = lambda: engine.marginals_windowed(pins[0].array, pins[1].array, pins[2].array, window=20, step=1, pad=True, out=pout.array)
print(f"int32 CSR call wall {wall(csr):.3f} ms ({sum(p.array.nbytes for p in pins) / 1e6:.1f} MB in)")
t0 = time.perf_counter()
ok = bool((numpy.diff(pins[1].array) >= 0).all())
print(f"numpy monotonicity pass over gene_ptr for scale: {(time.perf_counter() - t0) * 1e3:.3f} ms")

# do the two copy directions overlap on this host?  68.7 MB in on one stream, 16 MB out on another
host_in = torch.empty(wire.nbytes, dtype=torch.uint8).pin_memory()
dev_in = torch.empty(wire.nbytes, dtype=torch.uint8, device=dev)
host_out = torch.empty(batch.G * 8, dtype=torch.uint8).pin_memory()
dev_out = torch.empty(batch.G * 8, dtype=torch.uint8, device=dev)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    with torch.cuda.stream(s_in):
        dev_in.copy_(host_in, non_blocking=True)
    with torch.cuda.stream(s_out):
        host_out.copy_(dev_out, non_blocking=True)


def wall_sync(fn, n=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def pieces(k):
    cut = [wire.nbytes * i // k for i in range(k + 1)]

    def run():
        with torch.cuda.stream(s_in):
            for i in range(k):
                dev_in[cut[i]:cut[i + 1]].copy_(host_in[cut[i]:cut[i + 1]], non_blocking=True)
    return run


print(f"H2D alone (wall, sync each) {wall_sync(pieces(1)):.3f} ms; in 4 pieces {wall_sync(pieces(4)):.3f} ms; in 12 pieces {wall_sync(pieces(12)):.3f} ms")
print(f"H2D 68.7 MB and D2H 16 MB on two streams at once {wall_sync(both):.3f} ms")
