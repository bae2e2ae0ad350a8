#!/usr/bin/env python
"""bench.py — genes/s of ClusterCRF marginal inference on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--contigs C]

A *step* is one pass of the hot path (``gcrf_marginals_windowed``, W=20, step=1, pad=1, shipped GECCO
weights) over one synthetic batch of BASELINE config 2 (10k contigs x Poisson(200) genes x Poisson(25)
domains, 5 % unknown ids).  With N GPUs every rank owns its own batch of that shape
(contigs are independent, so the path shards with no data-path collective: weak scaling).

* ``value``     genes/s, inputs resident in HBM, CUDA events around the K steps, max over ranks
* ``e2e``       the same metric through the C ABI with HOST (pinned) buffers: H2D of the CSR batch,
                kernel, D2H of the marginals inside the timed region
* ``roofline``  algorithmic bytes (4 nnz + 4 (G+1) + 4 (C+1) + 8 G) / kernel time vs the
                measured HBM copy bandwidth in MEASURED_PEAKS.json
* ``cpu_baseline``  the CPU oracle (oracle/crf_oracle.c, a port of the reference's CPU path) on this
                box's host cores, same workload

``--impl reference`` times that CPU port alone (the reference's own tagger, python-crfsuite, is not
installable here: no network, not in the image — DESIGN.md "Oracle").
"""

from __future__ import annotations

import argparse
import json
import os
import pathlib
import statistics
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
# stdout carries exactly ONE JSON line.  Libraries print there too (NCCL's version banner at NCCL_DEBUG=WARN/VERSION):
# file descriptor 1 is pointed at stderr for the life of the process and the JSON line goes to the saved descriptor.
sys.stdout.flush()
_JSON_FD = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_JSON_FD, (json.dumps(line) + "\n").encode())


import numpy  # noqa: E402

METRIC = "genes/sec CRF marginal inference"
UNIT = "genes/s"
WINDOW, STEP, PAD = 20, 1, True
KERNEL_NAME = "gcrf::stream_kernel<20,128,4,int,256,false>"


def load_weights():
    from gecco_b200 import model_io

    return model_io.load_tsv_model(model_io.bundled_model_dir())


def make_batch(weights, contigs: int, seed: int):
    from gecco_b200 import synth

    return synth.config2(len(weights.attrs), seed=seed, contigs=contigs)


def workload_config(batch, n_gpus: int, contigs: int) -> dict:
    from gecco_b200 import synth

    return {
        "workload": f"BASELINE config 2: synthetic {contigs} contigs x Poisson(200) genes x Poisson(25) domains per GPU, "
                    f"5% unknown ids, shipped GECCO v0.11.0 weights, window 20 step 1 pad",
        "contigs_per_gpu": batch.C,
        "genes_per_gpu": batch.G,
        "nnz_per_gpu": batch.nnz,
        "windows_per_gpu": batch.windows(WINDOW, STEP, PAD),
        "algorithmic_bytes_per_gpu": synth.algorithmic_bytes(batch.C, batch.G, batch.nnz),
        "l2": "inputs+outputs per step exceed the 126 MB L2 (no flush needed)" if synth.algorithmic_bytes(batch.C, batch.G, batch.nnz) > 130e6
              else "L2 flushed between steps",
        "parallelism": f"contig-sharded x{n_gpus}, no data-path collective",
    }


# ------------------------------------------------------------------------------------------------
# CPU port (oracle) timing — used by cpu_baseline and by --impl reference
# ------------------------------------------------------------------------------------------------


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def time_cpu_port(weights, batch, threads: int, repeats: int):
    from oracle import crf_oracle

    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        crf_oracle.marginals_windowed(weights.state_w, weights.trans_w, weights.label_id("1"), batch.contig_ptr,
                                      batch.gene_ptr, batch.attr_idx, WINDOW, STEP, PAD, nthreads=threads)
        times.append(time.perf_counter() - t0)
    return times


def run_reference(args) -> None:
    """The reference arm: CPU port of the path, all host threads, bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    weights = load_weights()
    cores = host_cores()
    # calibrate on 200 contigs, then size the per-step sample so that the whole run stays ~<= 150 s
    probe = make_batch(weights, 200, seed=2)
    t_probe = min(time_cpu_port(weights, probe, cores, 2))
    per_contig = t_probe / probe.C
    budget = 150.0 / max(1, args.steps + args.warmup)
    contigs = int(max(200, min(args.contigs, budget / per_contig)))
    batch = make_batch(weights, contigs, seed=2) if contigs != probe.C else probe
    time_cpu_port(weights, batch, cores, args.warmup)
    times = time_cpu_port(weights, batch, cores, args.steps)
    total = sum(times)
    value = batch.G * args.steps / total
    cfg = workload_config(batch, 1, contigs)
    sample = f"{batch.C} contigs / {batch.G} genes of config 2 per step (full step = {args.contigs} contigs)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "restated CRFsuite forward-backward + the reference's window loop in C (oracle/crf_oracle.c); "
                                 "python-crfsuite itself is not installable in this image"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        if self.thread is not None:
            self.thread.join(timeout=5)

    def summary(self, t0: float, t1: float) -> dict:
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(rows[0][2]) if rows[0][2].isdigit() else None,
                "power_w_max": max((float(r[3]) for r in rows if r[3].replace(".", "").isdigit()), default=None),
                "samples": len(rows), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------


def measured_peak_gbs():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        try:
            return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except (ValueError, KeyError):
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def gpu_cpu_affinity(gpu_index: int):
    """Host cores NVML reports as local to the GPU (its NUMA node), or None."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        return [c for c in cpus if c in os.sched_getaffinity(0)] or None
    except Exception:
        return None


class bound_to_gpu:
    """Run the enclosed host work (pinned allocations are first-touched here, copies are driven from here) on this
    rank's slice of the cores local to its GPU; restores the previous affinity on exit."""

    def __init__(self, gpu_index: int, rank_slot: int, slots: int):
        self.gpu, self.slot, self.slots, self.prev, self.cpus = gpu_index, rank_slot, slots, None, None

    def __enter__(self):
        local = gpu_cpu_affinity(self.gpu)
        if local:
            # ranks whose GPUs share a node split its cores so that their copy threads do not stack on one core
            per = max(1, len(local) // max(1, self.slots))
            mine = local[self.slot * per:(self.slot + 1) * per] or local
            try:
                self.prev = os.sched_getaffinity(0)
                os.sched_setaffinity(0, mine)
                self.cpus = mine
            except OSError:
                self.prev = None
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            os.sched_setaffinity(0, self.prev)


def oracle_full(weights, batch, window=WINDOW, step=STEP, pad=PAD):
    from oracle import crf_oracle

    p, _ = crf_oracle.marginals_windowed(weights.state_w, weights.trans_w, weights.label_id("1"), batch.contig_ptr,
                                         batch.gene_ptr, batch.attr_idx, window, step, pad, nthreads=host_cores())
    return p


def max_abs_err(got, want) -> float:
    ok = ~numpy.isnan(want)
    if not numpy.array_equal(numpy.isnan(got), ~ok):
        return float("inf")
    return float(numpy.abs(got[ok] - want[ok]).max()) if ok.any() else 0.0


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from gecco_b200 import sharding, synth
    from gecco_b200._lib import GCRF_FLAG_F64, CRFEngine, PinnedArray

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    weights = load_weights()
    batch = make_batch(weights, args.contigs, seed=2 + rank)
    engine = CRFEngine(weights, device=local_rank)
    # a dedicated non-default stream: torch events only see the stream they are recorded on, and the
    # default stream's handle (0) would make the engine fall back to its own stream
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    engine.set_stream(stream.cuda_stream)
    peak, peak_src = measured_peak_gbs()
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > the 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def reduce_ranks(x: float, op) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x: float) -> float:
        return reduce_ranks(x, dist.ReduceOp.MAX) if world > 1 else x

    def sum_over_ranks(x: float) -> float:
        return reduce_ranks(x, dist.ReduceOp.SUM) if world > 1 else x

    def upload(b):
        cp = torch.from_numpy(b.contig_ptr).to(dev)
        gp = torch.from_numpy(b.gene_ptr).to(dev)
        ai = torch.full((b.nnz + 16,), -1, dtype=torch.int32, device=dev)  # readable slack behind the last id
        ai[: b.nnz] = torch.from_numpy(b.attr_idx).to(dev)
        return cp, gp, ai

    def kernel_times(launch, n: int, flush: bool):
        """Average device time of the kernels of one call (events inside the ABI around the kernels only)."""
        engine.set_timing(True)
        ms = []
        for it in range(2 + n):
            if flush:
                flush_buf.zero_()
            launch()
            t = engine.last_kernel_ms()
            if it >= 2:
                ms.append(t)
        engine.set_timing(False)
        return sum(ms) / len(ms), min(ms)

    # ---- device-resident inputs: the headline loop (BASELINE config 2, FP32 arithmetic)
    cp, gp, ai = upload(batch)
    out = torch.empty(batch.G, dtype=torch.float64, device=dev)

    def step_device(flags: int = 0):
        engine.marginals_windowed_device(cp.data_ptr(), gp.data_ptr(), ai.data_ptr(), batch.C, batch.G, batch.nnz,
                                         out.data_ptr(), window=WINDOW, step=STEP, pad=PAD, flags=flags)

    for _ in range(args.warmup):
        step_device()
    launches0 = engine.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    barrier()
    t_wall1 = time.perf_counter()
    launches = engine.launch_count - launches0
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    total_genes = sum_over_ranks(float(batch.G))
    value = total_genes * args.steps / (dev_ms * 1e-3)

    # ---- per-launch kernel time, separate pass (inputs + outputs exceed L2: no flush needed for config 2)
    algo_bytes = synth.algorithmic_bytes(batch.C, batch.G, batch.nnz)
    kernel_ms_avg, kernel_ms_min = kernel_times(step_device, min(args.steps, 10), flush=algo_bytes < 130e6)
    if rank == 0:
        time.sleep(0.2)
        sampler.stop()
    clocks = sampler.summary(t_wall0, t_wall1) if rank == 0 else None
    out_f32_arith = out.cpu().numpy()

    # ---- CPU baseline beside it (rank 0, N=1 only) — its output is the whole-batch parity reference
    cpu, want = None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        times = time_cpu_port(weights, batch, cores, 3)
        few = batch.slice_contigs(0, max(1, batch.C // 20))
        one = time_cpu_port(weights, few, 1, 1)[0]
        cpu = {"value": batch.G / min(times), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"the whole batch ({batch.C} contigs / {batch.G} genes), best of 3 passes, {cores} threads over contigs",
               "single_thread_value": few.G / one}
    if rank == 0:
        want = oracle_full(weights, batch)  # every gene of the timed batch, not a sample
    parity = max_abs_err(out_f32_arith, want) if rank == 0 else None

    # ---- the same batch in the reference's own arithmetic (GCRF_FLAG_F64): CRFsuite's f64 recursion, op by op
    f64_ms_avg, f64_ms_min = kernel_times(lambda: step_device(GCRF_FLAG_F64), 5, flush=False)
    f64_line = None
    if rank == 0:
        got64 = out.cpu().numpy()
        f64_line = {"dtype": "f64", "value": batch.G / (f64_ms_avg * 1e-3), "unit": UNIT, "kernel_ms_avg": f64_ms_avg,
                    "kernels": "gcrf::exact_unary_kernel + gcrf::exact_window_kernel<20>",
                    "roofline": {"bound": "hbm", "achieved": algo_bytes / (f64_ms_avg * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                 "frac": algo_bytes / (f64_ms_avg * 1e-3) / 1e9 / peak},
                    "parity_max_abs_err_vs_oracle": max_abs_err(got64, want),
                    "bit_identical_to_oracle_frac": float((got64 == want).mean()),
                    "note": "n_gpus=1 figure of rank 0; CRFsuite's scaled forward-backward in f64, no FMA contraction, "
                            "correctly rounded exp (gcrf_exact.cu)"}
    step_device()  # leave the FP32 result in `out` for the bit-identity checks below
    torch.cuda.synchronize(dev)

    # ---- end to end through the C ABI with pinned HOST buffers (H2D + kernels + D2H per step), three wire layouts
    from gecco_b200.packer import compact_ids

    engine.set_stream(None)
    slots = max(1, world)
    with bound_to_gpu(local_rank, rank_slot=local_rank, slots=slots) as binding:
        small = compact_ids(batch.attr_idx, len(weights.attrs))
        pins = [PinnedArray(a.shape, a.dtype) for a in (batch.contig_ptr, batch.gene_ptr, batch.attr_idx, small)]
        for pin, a in zip(pins, (batch.contig_ptr, batch.gene_ptr, batch.attr_idx, small)):
            pin.array[...] = a
        pout = PinnedArray((batch.G,), numpy.float64)
        pout32 = PinnedArray((batch.G,), numpy.float32)
        e2e_steps = max(3, min(args.steps, 10))

        def e2e_leg(one, h2d: int, out_pin, genes=None):
            for _ in range(2):
                one()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                one()
            torch.cuda.synchronize(dev)
            mine = time.perf_counter() - t0
            secs = max_over_ranks(mine)
            barrier()
            d2h = int(out_pin.array.nbytes)
            return {"value": (total_genes if genes is None else genes) * e2e_steps / secs, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": 1e3 * secs / e2e_steps,
                    "this_rank_pcie_gbs": (h2d + d2h) * e2e_steps / mine / 1e9}

        def csr_call(ids_pin, out_pin, f32):
            return lambda: engine.marginals_windowed(pins[0].array, pins[1].array, ids_pin.array, window=WINDOW, step=STEP,
                                                     pad=PAD, out=out_pin.array, f32=f32)

        def csr_bytes(ids_pin):
            return pins[0].array.nbytes + pins[1].array.nbytes + ids_pin.array.nbytes

        # headline: the compact wire format (gcrf_wire_encode: sorted ids as Rice-coded deltas, one-byte row lengths, one
        # page-locked block — what the packers hand to a bulk call), float64 marginals back
        from gecco_b200._lib import WireBatch

        WireBatch(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, len(weights.attrs)).close()  # first call: page-locking warms up
        t_enc = time.perf_counter()
        wire = WireBatch(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, len(weights.attrs))
        t_enc = time.perf_counter() - t_enc
        e2e = e2e_leg(lambda: engine.marginals_windowed_wire(wire, window=WINDOW, step=STEP, pad=PAD, out=pout.array), wire.nbytes, pout)
        e2e["layout"] = "gcrf_wire block (int32 contig_ptr, uint8 ids/bytes per gene, Rice-coded deltas of the sorted ids; copy in / kernels / copy back pipelined over 4 slices), float64 marginals"
        e2e["bit_identical_to_device_path"] = bool(numpy.array_equal(pout.array, out_f32_arith))
        e2e["wire_encode_ms_once_outside_the_timed_region"] = 1e3 * t_enc  # host threads, CSR arrays -> page-locked block
        e2e_u16 = e2e_leg(csr_call(pins[3], pout, False), csr_bytes(pins[3]), pout)
        e2e_u16["layout"] = "int32 row pointers, uint16 attribute ids, float64 marginals"
        e2e_u16["bit_identical_to_device_path"] = bool(numpy.array_equal(pout.array, out_f32_arith))
        e2e_i32 = e2e_leg(csr_call(pins[2], pout, False), csr_bytes(pins[2]), pout)
        e2e_i32["layout"] = "int32 row pointers, int32 attribute ids, float64 marginals (the CSR layout of SURVEY.md 8(b))"
        e2e_i32["bit_identical_to_device_path"] = bool(numpy.array_equal(pout.array, out_f32_arith))
        compact = e2e_leg(lambda: engine.marginals_windowed_wire(wire, window=WINDOW, step=STEP, pad=PAD, out=pout32.array, f32=True),
                          wire.nbytes, pout32)
        compact["layout"] = "gcrf_wire block, float32 marginals (the FP32 results un-widened)"
        compact["identical_to_float32_of_device_path"] = bool(numpy.array_equal(pout32.array, out_f32_arith.astype(numpy.float32)))
        wire.close()
        # the same call on the real annotation density (1.4 domains per gene): what a metagenome table looks like
        real = None
        if world == 1:
            from gecco_b200 import synth as synth_mod

            sparse = synth_mod.config2(len(weights.attrs), seed=2, contigs=args.contigs, mean_domains=1.4, unknown_fraction=0.0)
            wire_s = WireBatch(sparse.contig_ptr, sparse.gene_ptr, sparse.attr_idx, len(weights.attrs))
            pout_s = PinnedArray((sparse.G,), numpy.float64)
            want_s = engine.marginals_windowed(sparse.contig_ptr, sparse.gene_ptr, sparse.attr_idx, window=WINDOW, step=STEP, pad=PAD)
            real = e2e_leg(lambda: engine.marginals_windowed_wire(wire_s, window=WINDOW, step=STEP, pad=PAD, out=pout_s.array), wire_s.nbytes, pout_s,
                           genes=float(sparse.G))
            real["layout"] = "gcrf_wire block, float64 marginals; config 2 at 1.4 domains per gene"
            real["genes"] = int(sparse.G)
            real["bit_identical_to_device_path"] = bool(numpy.array_equal(pout_s.array, want_s))
            wire_s.close()
            del sparse, want_s, pout_s
        per_rank_gbs = [e2e["this_rank_pcie_gbs"]]
        if world > 1:
            t = torch.tensor([e2e["this_rank_pcie_gbs"]], dtype=torch.float64, device=dev)
            allv = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(allv, t)
            per_rank_gbs = [float(v.item()) for v in allv]
        e2e["per_rank_pcie_gbs"] = per_rank_gbs
        e2e["host_binding"] = {"cpus_of_rank0": binding.cpus, "policy": "each rank on its slice of the cores NVML reports local to its GPU; "
                               "pinned staging first-touched there"}

    # ---- the stage in front of the marginals: accession -> attribute id on device, same batch.
    features_stage = None
    if rank == 0 and world == 1 and engine.has_vocabulary:  # like cpu_baseline: on the N=1 line only
        nums = numpy.array([int(a[2:]) for a in weights.attrs], dtype=numpy.int32)
        acc_host = numpy.where(batch.attr_idx >= 0, nums[numpy.clip(batch.attr_idx, 0, len(nums) - 1)], 99_999).astype(numpy.int32)
        d_acc = torch.from_numpy(acc_host).to(dev)
        d_ids = torch.empty(batch.nnz, dtype=torch.int32, device=dev)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        engine.set_stream(stream.cuda_stream)
        f_ms = []
        for it in range(3 + 8):
            f0.record(stream)
            engine.features_from_accessions_device(d_acc.data_ptr(), gp.data_ptr(), batch.G, batch.nnz, d_ids.data_ptr())
            f1.record(stream)
            torch.cuda.synchronize(dev)
            if it >= 3:
                f_ms.append(f0.elapsed_time(f1))
        f_ms_avg = sum(f_ms) / len(f_ms)
        f_bytes = 8 * batch.nnz + 4 * (batch.G + 1)  # accession in, id out, row pointers
        ids_equal = bool(torch.equal(d_ids, ai[: batch.nnz]))
        # accessions in -> marginals out in ONE call on device-resident arrays (GCRF_FLAG_ACCESSIONS)
        from gecco_b200._lib import GCRF_FLAG_ACCESSIONS

        def step_acc():
            engine.marginals_windowed_device(cp.data_ptr(), gp.data_ptr(), d_acc.data_ptr(), batch.C, batch.G, batch.nnz,
                                             out.data_ptr(), window=WINDOW, step=STEP, pad=PAD, flags=GCRF_FLAG_ACCESSIONS)
        acc_ms_avg, _ = kernel_times(step_acc, 8, flush=False)
        acc_equal = bool(numpy.array_equal(out.cpu().numpy(), out_f32_arith))
        features_stage = {"kernel_ms_avg": f_ms_avg, "rows": int(batch.nnz), "algorithmic_bytes_per_launch": int(f_bytes),
                          "achieved_gbs": f_bytes / (f_ms_avg * 1e-3) / 1e9, "ids_equal_packer": ids_equal,
                          "accessions_to_marginals": {"kernel_ms_avg": acc_ms_avg, "genes_per_s": batch.G / (acc_ms_avg * 1e-3),
                                                      "roofline_frac": algo_bytes / (acc_ms_avg * 1e-3) / 1e9 / peak,
                                                      "bit_identical_to_ids_path": acc_equal}}
        engine.set_stream(None)
        del d_acc, d_ids

    # ---- the other BASELINE configs on the same line (rank 0, N=1): kernel time, roofline fraction, whole-batch parity
    configs = None
    big = None  # config 4 at its full size, kept for the sharded leg
    engine.set_stream(stream.cuda_stream)
    if rank == 0 and world == 1 and not args.no_configs:
        configs = []
        A = len(weights.attrs)

        def config_entry(name, b, *, chain=False, note=None):
            dcp, dgp, dai = upload(b)
            dout = torch.empty(b.G, dtype=torch.float64, device=dev)
            if chain:
                def go():
                    engine.marginals_chain_device(dcp.data_ptr(), dgp.data_ptr(), dai.data_ptr(), b.C, b.G, b.nnz, dout.data_ptr(),
                                                  ptr64=b.gene_ptr.dtype == numpy.int64)
            else:
                def go():
                    engine.marginals_windowed_device(dcp.data_ptr(), dgp.data_ptr(), dai.data_ptr(), b.C, b.G, b.nnz, dout.data_ptr(),
                                                     window=WINDOW, step=STEP, pad=PAD, ptr64=b.gene_ptr.dtype == numpy.int64)
            ab = synth.algorithmic_bytes(b.C, b.G, b.nnz)
            ms_avg, ms_min = kernel_times(go, 6, flush=ab < 400e6)
            got = dout.cpu().numpy()
            if chain:
                from oracle import crf_oracle

                err = 0.0
                for c in range(b.C):
                    g0, g1 = int(b.contig_ptr[c]), int(b.contig_ptr[c + 1])
                    ref = crf_oracle.chain_marginals(weights.state_w, weights.trans_w, b.gene_ptr, b.attr_idx, g0, g1)[:, 1]
                    err = max(err, float(numpy.abs(got[g0:g1] - ref).max()))
            else:
                err = max_abs_err(got, oracle_full(weights, b))
            entry = {"config": name, "contigs": b.C, "genes": b.G, "nnz": b.nnz, "kernel_ms_avg": ms_avg, "kernel_ms_min": ms_min,
                     "value": b.G / (ms_avg * 1e-3), "unit": UNIT, "algorithmic_bytes": ab,
                     "roofline_frac": ab / (ms_avg * 1e-3) / 1e9 / peak, "parity_max_abs_err_vs_oracle": err,
                     "parity_genes_checked": b.G, "l2_flushed_between_launches": ab < 400e6}
            if note:
                entry["note"] = note
            configs.append(entry)
            del dcp, dgp, dai, dout

        config_entry("2 at real annotation density (10k contigs x Poisson(200) genes x Poisson(1.4) domains)",
                     synth.config2(A, seed=2, mean_domains=1.4, unknown_fraction=0.0))
        config_entry("2 with Zipf(1.1) attribute frequencies (SURVEY 8(d): the skew of real Pfam annotations; repeats inside a gene are dropped, "
                     "so rows are shorter: see nnz)", synth.config2(A, seed=2, zipf=1.1))
        config_entry("3 E. coli-like stand-in (1 contig x 4,300 genes x Poisson(1.4) domains; the real table is not in the reference)",
                     synth.config3_ecoli_like(A))
        big = synth.config4_chunked(A, seed=4)
        config_entry("4 metagenome, full size (1M contigs, lognormal lengths, ~31% shorter than the window, Poisson(25) domains)", big)
        config_entry("4 metagenome, full size, real annotation density (Poisson(1.4) domains)", synth.config4_chunked(A, seed=4, mean_domains=1.4))
        five = synth.config5(A)
        config_entry("5 long contigs (100 x 5,000 genes x Poisson(25) domains), GECCO semantics W=20", five)
        config_entry("5 long contigs, deep-chain primitive (one 5,000-item chain per contig, gcrf_marginals_chain, f64)", five, chain=True,
                     note="tolerance 1e-9 (f64 2x2 scan vs the oracle's sequential scaled recursion)")

    # ---- ONE batch sharded by contig over the ranks (BASELINE config 4 at full size): each rank ingests its own
    #      shard from pinned host memory, runs the kernel, and an NCCL all-gather of the padded f64 shards puts every
    #      gene's marginal on every rank — kernel + collective inside the timed region (strong scaling)
    sharded = None
    if not args.no_sharded:
        A = len(weights.attrs)
        lens = synth.config4_lengths(4)
        cptr = numpy.concatenate([[0], numpy.cumsum(lens)])
        parts = sharding.partition_contigs(cptr, None, world, WINDOW, contig_nnz=25.0 * lens)  # expected ids per contig
        c0, c1 = parts[rank]
        if big is not None and world == 1:
            shard = big
        else:
            shard = synth.config4_chunked(A, seed=4, contig_range=(c0, c1), threads=max(1, host_cores() // max(1, world)))
        big = None
        sizes = [int(cptr[b] - cptr[a]) for a, b in parts]
        gmax = max(sizes)
        ds = sharding.DeviceShard(shard, dev)
        local = torch.zeros(gmax, dtype=torch.float64, device=dev)
        gathered = torch.empty(world * gmax, dtype=torch.float64, device=dev) if world > 1 else local
        ptr64 = shard.gene_ptr.dtype == numpy.int64

        def shard_kernel():
            engine.marginals_windowed_device(ds.contig_ptr.data_ptr(), ds.gene_ptr.data_ptr(), ds.attr_idx.data_ptr(), ds.C, ds.G,
                                             ds.nnz, local.data_ptr(), window=WINDOW, step=STEP, pad=PAD, ptr64=ptr64)

        def shard_gather():
            if world > 1:
                dist.all_gather_into_tensor(gathered, local)

        sh_steps = max(3, min(args.steps, 10))
        for _ in range(3):
            shard_kernel()
            shard_gather()
        barrier()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2 * sh_steps + 1)]
        e[0].record(stream)
        for k in range(sh_steps):
            shard_kernel()
            e[2 * k + 1].record(stream)
            shard_gather()
            e[2 * k + 2].record(stream)
        barrier()
        ms_total = max_over_ranks(e[0].elapsed_time(e[-1])) / sh_steps
        ms_kernel = max_over_ranks(sum(e[2 * k].elapsed_time(e[2 * k + 1]) for k in range(sh_steps)) / sh_steps)
        ms_gather = max_over_ranks(sum(e[2 * k + 1].elapsed_time(e[2 * k + 2]) for k in range(sh_steps)) / sh_steps)
        # parity: every rank checks ITS contigs' slice of the gathered array against the oracle on its shard
        got = gathered[rank * gmax: rank * gmax + shard.G].cpu().numpy() if world > 1 else local[: shard.G].cpu().numpy()
        err = max_over_ranks(max_abs_err(got, oracle_full(weights, shard)))
        # ---- the gather fused into the kernel: results stored straight into every rank's output array over NVLink
        #      (peer stores, or one multimem.st per result to the NVLS multicast address), a device-side barrier instead
        #      of a collective (gcrf_marginals_windowed_peers, gecco_b200/sharding.py FusedGather)
        fused = {}
        if world > 1:
            offs = numpy.concatenate([[0], numpy.cumsum(sizes)])
            for name, mc in (("peer_stores", False), ("nvls_multicast", True)):
                try:
                    fg = sharding.FusedGather(sizes, dev, multicast=mc)
                    if mc and not fg.multicast:
                        fused[name] = {"unavailable": "no NVLS multicast support reported by torch symmetric memory"}
                        continue
                    for _ in range(3):
                        res = fg.predict(engine, ds, window=WINDOW, step=STEP, pad=PAD)
                    barrier()
                    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    f0.record(stream)
                    for _ in range(sh_steps):
                        res = fg.predict(engine, ds, window=WINDOW, step=STEP, pad=PAD)
                    f1.record(stream)
                    barrier()
                    ms_f = max_over_ranks(f0.elapsed_time(f1)) / sh_steps
                    same = all(bool(torch.equal(res[int(offs[r]):int(offs[r + 1])], gathered[r * gmax: r * gmax + sizes[r]])) for r in range(world))
                    same = reduce_ranks(1.0 if same else 0.0, dist.ReduceOp.MIN) == 1.0
                    fused[name] = {"value": int(cptr[-1]) / (ms_f * 1e-3), "unit": UNIT, "ms_per_step": ms_f,
                                   "equal_to_nccl_all_gather_on_every_rank": same,
                                   "note": "two device-side barriers per step inside the timed region (before: nobody still reads the "
                                           "previous result; after: every rank's stores have landed)"}
                    del fg, res
                except Exception as err:  # symmetric memory is a torch-private API: report, do not fail the bench
                    fused[name] = {"unavailable": f"{type(err).__name__}: {err}"[:300]}
            engine.set_stream(stream.cuda_stream)
        # the same pass with the shard coming from (pinned) host memory every step: per-rank H2D inside the timed region
        with bound_to_gpu(local_rank, rank_slot=local_rank, slots=slots):
            hp = [PinnedArray(a.shape, a.dtype) for a in (shard.contig_ptr, shard.gene_ptr, shard.attr_idx)]
            for pin, a in zip(hp, (shard.contig_ptr, shard.gene_ptr, shard.attr_idx)):
                pin.array[...] = a
            srcs = [torch.from_numpy(pin.array) for pin in hp]
            dsts = [ds.contig_ptr, ds.gene_ptr, ds.attr_idx[: shard.nnz]]

            def ingest():
                for d, h in zip(dsts, srcs):
                    d.copy_(h, non_blocking=True)

            for _ in range(2):
                ingest()
                shard_kernel()
                shard_gather()
            barrier()
            i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            i0.record(stream)
            for _ in range(sh_steps):
                ingest()
                shard_kernel()
                shard_gather()
            i1.record(stream)
            barrier()
            ms_ingest_total = max_over_ranks(i0.elapsed_time(i1)) / sh_steps
            for pin in hp:
                pin.free()
        G_all = int(cptr[-1])
        sharded = {"workload": "BASELINE config 4 at full size: 1,000,000 contigs / %d genes, Poisson(25) domains, one batch "
                               "cost-partitioned by contig over the ranks" % G_all,
                   "scaling": "strong", "n_gpus": world, "genes": G_all, "genes_per_rank": sizes,
                   "value": G_all / (ms_total * 1e-3), "unit": UNIT, "ms_per_step": ms_total, "ms_kernel": ms_kernel,
                   "ms_gather": ms_gather, "gather_bytes": int(8 * gmax * world) if world > 1 else 0,
                   "collective": "ncclAllGather (torch.distributed all_gather_into_tensor) on the kernel's stream" if world > 1 else None,
                   "steps": sh_steps, "parity_max_abs_err_vs_oracle": err, "parity_genes_checked": G_all,
                   "fused_gather": fused or None,
                   "with_host_ingest": {"value": G_all / (ms_ingest_total * 1e-3), "ms_per_step": ms_ingest_total,
                                        "h2d_bytes_this_rank": int(shard.contig_ptr.nbytes + shard.gene_ptr.nbytes + shard.attr_idx.nbytes),
                                        "note": "every rank copies its own shard from pinned host memory each step, then kernel + all-gather"}}
        del ds, local, gathered

    if rank == 0:
        traffic = None
        tpath = ROOT / "profiles" / "traffic_bytes_per_launch.json"
        if tpath.exists():
            try:
                t = json.loads(tpath.read_text())
                if int(t.get("genes", -1)) == batch.G:
                    traffic = t.get("dram_bytes_per_launch")
            except ValueError:
                pass
        achieved = algo_bytes / (kernel_ms_avg * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(batch, world, args.contigs),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": "ncu --set full capture of this kernel on this workload, committed under profiles/",
                         "peak_source": peak_src, "kernel": KERNEL_NAME,
                         "kernel_ms_avg": kernel_ms_avg, "kernel_ms_min": kernel_ms_min, "algorithmic_bytes_per_launch": algo_bytes,
                         "note": "kernel_ms_avg brackets single launches with events; back-to-back launches (ms_per_step) overlap "
                                 "their prologues through programmatic dependent launch and are slightly faster"},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "e2e_uint16_ids": e2e_u16,
            "e2e_int32_ids": e2e_i32,
            "e2e_compact": compact,
            "e2e_real_density": real,
            "f64": f64_line,
            "configs": configs,
            "sharded": sharded,
            "features_stage": features_stage,
            "gpu_launches": launches,
            "clocks": clocks,
            "parity_max_abs_err_vs_oracle": parity,
            "parity_genes_checked": batch.G,
        }
        emit(line)
    for p_ in pins + [pout, pout32]:
        p_.free()
    engine.close()
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--contigs", type=int, default=10_000, help="contigs per GPU (config 2 = 10000)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs (N=1 line)")
    ap.add_argument("--no-sharded", action="store_true", help="skip the strong-scaling leg (config 4 sharded by contig)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
