"""Kernel time of on-device feature extraction (gcrf_features_from_accessions) on config-2-like accession rows. B200 only."""
import pathlib
import sys
import time

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy
import torch
from gecco_b200 import model_io, synth
from gecco_b200._lib import CRFEngine, GCRF_FLAG_DEVICE_PTRS

w = model_io.load_tsv_model(model_io.bundled_model_dir())
eng = CRFEngine(w, 0)
nums = numpy.array([int(a[2:]) for a in w.attrs], dtype=numpy.int32)
dev = torch.device("cuda:0")
for name, b in (("config2", synth.config2(len(w.attrs))), ("sparse", synth.config2(len(w.attrs), mean_domains=1.4))):
    ids = b.attr_idx.copy()
    acc = numpy.where(ids >= 0, nums[numpy.clip(ids, 0, len(nums) - 1)], 99999).astype(numpy.int32)
    d_acc = torch.from_numpy(acc).to(dev)
    d_ptr = torch.from_numpy(b.gene_ptr).to(dev)
    d_out = torch.empty(len(acc), dtype=torch.int32, device=dev)
    lib = eng._lib
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    ts = []
    for it in range(8):
        e0.record(stream)
        rc = lib.gcrf_features_from_accessions(eng._handle, d_acc.data_ptr(), d_ptr.data_ptr(), b.G, len(acc), d_out.data_ptr(), GCRF_FLAG_DEVICE_PTRS)
        e1.record(stream)
        torch.cuda.synchronize()
        assert rc == 0
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    out = d_out.cpu().numpy()
    ok = numpy.array_equal(out[ids >= 0], ids[ids >= 0]) and bool((out[ids < 0] == -1).all())
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name}: {len(acc)} rows, {b.G} genes: {ms*1e3:.1f} us ({len(acc)*8/ms/1e6:.0f} GB/s of 8 B/row), matches packer ids: {ok}", flush=True)
