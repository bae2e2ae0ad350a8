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


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from gecco_b200 import synth
    from gecco_b200._lib import CRFEngine, PinnedArray

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

    # ---- device-resident inputs
    cp = torch.from_numpy(batch.contig_ptr).to(dev)
    gp = torch.from_numpy(batch.gene_ptr).to(dev)
    ai = torch.from_numpy(batch.attr_idx).to(dev)
    out = torch.empty(batch.G, dtype=torch.float64, device=dev)

    def step_device():
        engine.marginals_windowed_device(cp.data_ptr(), gp.data_ptr(), ai.data_ptr(), batch.C, batch.G, batch.nnz,
                                         out.data_ptr(), window=WINDOW, step=STEP, pad=PAD)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

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

    # ---- per-launch kernel time (events inside the ABI around the kernel only), separate pass
    kernel_ms = []
    engine.set_timing(True)
    for _ in range(min(args.steps, 10)):
        step_device()
        kernel_ms.append(engine.last_kernel_ms())
    kernel_ms_avg = sum(kernel_ms) / len(kernel_ms)
    engine.set_timing(False)
    if rank == 0:
        time.sleep(0.2)
        sampler.stop()
    clocks = sampler.summary(t_wall0, t_wall1) if rank == 0 else None

    # ---- parity spot check of what was timed (oracle on the first contigs; rank 0 only)
    parity = None
    if rank == 0:
        from oracle import crf_oracle

        sub = batch.slice_contigs(0, min(batch.C, 64))
        want, _ = crf_oracle.marginals_windowed(weights.state_w, weights.trans_w, weights.label_id("1"), sub.contig_ptr,
                                                sub.gene_ptr, sub.attr_idx, WINDOW, STEP, PAD, nthreads=4)
        got = out[:sub.G].cpu().numpy()
        parity = float(numpy.abs(got - want).max())

    # ---- end to end through the C ABI with pinned HOST buffers (H2D + kernel + D2H per step)
    engine.set_stream(None)
    pins = [PinnedArray(a.shape, a.dtype) for a in (batch.contig_ptr, batch.gene_ptr, batch.attr_idx)]
    for pin, a in zip(pins, (batch.contig_ptr, batch.gene_ptr, batch.attr_idx)):
        pin.array[...] = a
    pout = PinnedArray((batch.G,), numpy.float64)

    def step_host():
        engine.marginals_windowed(pins[0].array, pins[1].array, pins[2].array, window=WINDOW, step=STEP, pad=PAD,
                                  out=pout.array)

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = total_genes * e2e_steps / e2e_s
    h2d = int(batch.contig_ptr.nbytes + batch.gene_ptr.nbytes + batch.attr_idx.nbytes)
    d2h = int(batch.G * 8)
    e2e_equal = bool(numpy.array_equal(pout.array, out.cpu().numpy()))

    # ---- the same call with compact buffers: uint16 ids (GCRF_FLAG_IDX_U16, widened on the device) and float32
    #      marginals — a separate variant with its own byte counts, not the headline
    from gecco_b200.packer import compact_ids

    small = compact_ids(batch.attr_idx, len(weights.attrs))
    pin16 = PinnedArray(small.shape, small.dtype)
    pin16.array[...] = small
    pout32 = PinnedArray((batch.G,), numpy.float32)

    def step_host_compact():
        engine.marginals_windowed(pins[0].array, pins[1].array, pin16.array, window=WINDOW, step=STEP, pad=PAD,
                                  out=pout32.array, f32=True)

    for _ in range(2):
        step_host_compact()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host_compact()
    torch.cuda.synchronize(dev)
    compact_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    compact = {"value": total_genes * e2e_steps / compact_s, "unit": UNIT,
               "h2d_bytes_per_step": int(batch.contig_ptr.nbytes + batch.gene_ptr.nbytes + small.nbytes),
               "d2h_bytes_per_step": int(batch.G * 4), "ms_per_step": 1e3 * compact_s / e2e_steps,
               "layout": "uint16 attribute ids, float32 marginals",
               "identical_to_float32_of_device_path": bool(numpy.array_equal(pout32.array, out.cpu().numpy().astype(numpy.float32)))}

    # ---- the stage in front of the marginals: accession -> attribute id on device, same batch.
    # Outside the headline's timed region; reported so that "accessions in, marginals out" has a measured number.
    features_stage = None
    if rank == 0 and world == 1 and engine.has_vocabulary:  # like cpu_baseline: on the N=1 line only
        nums = numpy.array([int(a[2:]) for a in weights.attrs], dtype=numpy.int32)
        acc_host = numpy.where(batch.attr_idx >= 0, nums[numpy.clip(batch.attr_idx, 0, len(nums) - 1)], 99_999).astype(numpy.int32)
        d_acc = torch.from_numpy(acc_host).to(dev)
        d_ids = torch.empty(batch.nnz, dtype=torch.int32, device=dev)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        engine.set_stream(stream.cuda_stream)  # the host-buffer legs above run on the engine's own stream
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
        features_stage = {"kernel": "gcrf::features_kernel<int,8,8,512,true>", "kernel_ms_avg": f_ms_avg,
                          "rows": int(batch.nnz), "algorithmic_bytes_per_launch": int(f_bytes),
                          "achieved_gbs": f_bytes / (f_ms_avg * 1e-3) / 1e9,
                          "ids_equal_packer": bool(torch.equal(d_ids, ai)),
                          "genes_per_s_features_plus_marginals": batch.G / ((f_ms_avg + kernel_ms_avg) * 1e-3)}
        engine.set_stream(None)
        del d_acc, d_ids
        # ... and end to end from raw accessions (GCRF_FLAG_ACCESSIONS): H2D + features + marginals + D2H per step
        pin_acc = PinnedArray(acc_host.shape, acc_host.dtype)
        pin_acc.array[...] = acc_host

        def step_host_accessions():
            engine.marginals_windowed(pins[0].array, pins[1].array, pin_acc.array, window=WINDOW, step=STEP, pad=PAD,
                                      out=pout.array, accessions=True)

        for _ in range(2):
            step_host_accessions()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_host_accessions()
        torch.cuda.synchronize(dev)
        acc_s = time.perf_counter() - t0
        features_stage["e2e_from_accessions"] = {
            "value": batch.G * e2e_steps / acc_s, "unit": UNIT, "ms_per_step": 1e3 * acc_s / e2e_steps,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "n_gpus": 1,
            "bit_identical_to_device_path": bool(numpy.array_equal(pout.array, out.cpu().numpy()))}
        pin_acc.free()

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        sample = batch if batch.C <= 10_000 else batch.slice_contigs(0, 10_000)
        times = time_cpu_port(weights, sample, cores, 3)
        one = time_cpu_port(weights, sample.slice_contigs(0, max(1, sample.C // 20)), 1, 1)[0]
        cpu = {"value": sample.G / min(times), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{sample.C} contigs / {sample.G} genes of the same batch, best of 3 passes, {cores} threads over contigs",
               "single_thread_value": (sample.slice_contigs(0, max(1, sample.C // 20)).G) / one}

    if rank == 0:
        algo_bytes = synth.algorithmic_bytes(batch.C, batch.G, batch.nnz)
        peak, peak_src = measured_peak_gbs()
        achieved = algo_bytes / (kernel_ms_avg * 1e-3) / 1e9
        traffic = None
        tpath = ROOT / "profiles" / "traffic_bytes_per_launch.json"
        if tpath.exists():
            try:
                t = json.loads(tpath.read_text())
                if int(t.get("genes", -1)) == batch.G:
                    traffic = t.get("dram_bytes_per_launch")
            except ValueError:
                pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(batch, world, args.contigs),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "gcrf::stream_kernel<20,128,4,int>",
                         "kernel_ms_avg": kernel_ms_avg, "algorithmic_bytes_per_launch": algo_bytes},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps, "bit_identical_to_device_path": e2e_equal},
            "e2e_compact": compact,
            "features_stage": features_stage,
            "gpu_launches": launches,
            "clocks": clocks,
            "parity_max_abs_err_vs_oracle": parity,
        }
        emit(line)
    for p in pins + [pout, pin16, pout32]:
        p.free()
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
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
