"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures.

Tolerance: the north star asks for |dp| <= 1e-5 against Tagger.marginal(); the device computes in
FP32, the oracle in f64.  TOL below is the bar written into the test.
"""
import numpy
import pytest

from conftest import pack_case

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(autouse=True, params=["fused", "generic"])
def device_path(request, monkeypatch):
    """Every parity test runs through both device paths for W=20: the fused streaming kernel (default) and the
    generic kernel that serves every other window size (the library reads GCRF_FORCE_GENERIC at call time)."""
    monkeypatch.setenv("GCRF_FORCE_GENERIC", "1" if request.param == "generic" else "0")
    return request.param


def oracle(weights, batch, window=20, step=1, pad=True, nthreads=8):
    import os

    from oracle import crf_oracle

    if nthreads <= 0:  # whole-batch checks: every host core
        nthreads = len(os.sched_getaffinity(0))

    p, _ = crf_oracle.marginals_windowed(weights.state_w, weights.trans_w, weights.label_id("1"), batch.contig_ptr,
                                         batch.gene_ptr, batch.attr_idx, window, step, pad, nthreads=nthreads)
    return p


def assert_close(got, want, tol=TOL, what=""):
    assert got.shape == want.shape
    assert numpy.array_equal(numpy.isnan(got), numpy.isnan(want)), what
    ok = ~numpy.isnan(want)
    if ok.any():
        err = numpy.abs(got[ok] - want[ok]).max()
        assert err <= tol, f"{what}: max |dp| = {err:.3e}"
        return err
    return 0.0


def test_golden_bgc0001866(engine, bgc, weights):
    from test_oracle import _bgc_csr

    packed, genes = _bgc_csr(bgc, weights)
    p = engine.marginals_windowed(packed.contig_ptr, packed.gene_ptr, packed.attr_idx)
    golden = numpy.array([g["average_p"] for g in genes])[packed.order]
    assert_close(p, golden, what="python-crfsuite golden")


def test_reference_loop_cases(engine, ref_cases, weights):
    for case in ref_cases:
        packed = pack_case(case, weights)
        p = engine.marginals_windowed(packed.contig_ptr, packed.gene_ptr, packed.attr_idx, window=case["window"],
                                      step=case["step"], pad=case["pad"])
        want = numpy.array([numpy.nan if e["p"] is None else e["p"] for e in case["expected"]])
        assert_close(p, want, what=case["name"])


def test_mibig_real_features(engine, mibig, weights):
    from gecco_b200.packer import pack_arrays

    packed = pack_arrays(mibig["gene_contig"], mibig["dom_ptr"], mibig["dom_pfam"], weights)
    p = engine.marginals_windowed(packed.contig_ptr, packed.gene_ptr, packed.attr_idx)
    err = assert_close(p, mibig["ref_loop_prob"], what="mibig vs reference loop")
    print(f"mibig max |dp| = {err:.3e}")


@pytest.mark.parametrize("window,step,pad", [(20, 1, True), (20, 1, False), (20, 3, True), (5, 1, True), (5, 2, False),
                                             (10, 1, True), (10, 3, False), (7, 7, True), (1, 1, True), (33, 4, True),
                                             (64, 1, True), (128, 5, True), (15, 1, True), (25, 2, False), (30, 1, True),
                                             (40, 1, True), (40, 7, False), (50, 3, True), (64, 9, False)])
def test_ragged_edge_cases(engine, weights, window, step, pad):
    from gecco_b200 import synth

    batch = synth.ragged_edge_cases(len(weights.attrs))
    want = oracle(weights, batch, window, step, pad)
    got = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, window=window, step=step, pad=pad)
    assert_close(got, want, what=f"W={window} step={step} pad={pad}")


@pytest.mark.parametrize("window", [5, 10, 15, 25, 30, 40, 50, 64])
def test_other_streaming_windows_on_dense_and_short_contigs(engine, weights, window):
    """The streaming kernel is also compiled for W = 5 (`gecco train`'s default) and a spread of other sizes: odd and
    even meeting points of the two DP chains, on the dense shape and on the metagenome shape (padding and skipping)."""
    from gecco_b200 import synth

    batch = synth.config2(len(weights.attrs), contigs=200)
    assert_close(engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, window=window),
                 oracle(weights, batch, window=window), what=f"config2 W={window}")
    batch = synth.config4(len(weights.attrs), contigs=3000, mean_domains=4.0)
    for pad in (True, False):
        for step in (1, 2):
            got = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, window=window, step=step, pad=pad)
            assert_close(got, oracle(weights, batch, window=window, step=step, pad=pad), what=f"config4 W={window} step={step} pad={pad}")


def test_random_shapes_fuzz(engine, weights):
    """Seeded fuzz over everything that shapes the control flow: window, step, padding, contig lengths from one gene
    to several tiles, domain density from none to rows longer than the fixed-point guard, unknown ids."""
    from gecco_b200 import synth

    rng = numpy.random.default_rng(2024)
    for case in range(36):
        window = int(rng.choice([5, 10, 20, 20, 20, 15, 25, 30, 40, 50, 64, rng.integers(1, 41)]))
        step = int(rng.integers(1, window + 1)) if rng.random() < 0.4 else 1
        pad = bool(rng.random() < 0.7)
        kind = case % 4
        if kind == 0:
            lens = rng.integers(1, 2 * window + 2, size=int(rng.integers(1, 400)))
        elif kind == 1:
            lens = numpy.maximum(1, rng.poisson(rng.choice([30, 250, 900]), size=int(rng.integers(1, 40))))
        elif kind == 2:
            lens = numpy.concatenate([rng.integers(1, 4, size=300), [int(rng.integers(500, 3000))], rng.integers(1, 60, size=50)])
            rng.shuffle(lens)
        else:
            lens = numpy.array([int(rng.integers(1, 5000))])
        density = float(rng.choice([0.0, 0.3, 1.4, 8.0, 25.0, 45.0]))
        batch = synth.make_batch(rng, lens, density, len(weights.attrs), float(rng.choice([0.0, 0.05, 0.5])))
        got = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, window=window, step=step, pad=pad)
        want = oracle(weights, batch, window, step, pad)
        assert_close(got, want, what=f"case {case}: W={window} step={step} pad={pad} C={batch.C} G={batch.G} d={density}")


def test_full_size_config2_properties(engine, weights, monkeypatch):
    """BASELINE config 2 at its full size (10,000 contigs, 2.0 M genes, 49.8 M ids): (1) every gene against the oracle
    (see also tests/test_gpu_fullsize.py); (2) contigs are independent: any contiguous slice of the batch, run alone,
    reproduces its part bit for bit; (3) every probability is a probability."""
    from gecco_b200 import synth

    batch = synth.config2(len(weights.attrs))
    assert batch.G == 2_000_810 and batch.nnz == 49_793_052  # SURVEY.md §8(d), seed 2
    fast = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx)
    assert numpy.isfinite(fast).all() and fast.min() >= 0.0 and fast.max() <= 1.0
    assert_close(fast, oracle(weights, batch, nthreads=0), what="config 2, all genes")
    for c0, c1 in ((0, 1), (4321, 4400), (9000, 10000)):
        part = batch.slice_contigs(c0, c1)
        got = engine.marginals_windowed(part.contig_ptr, part.gene_ptr, part.attr_idx)
        g0, g1 = int(batch.contig_ptr[c0]), int(batch.contig_ptr[c1])
        assert numpy.array_equal(got, fast[g0:g1])


def test_metagenome_scale_properties(engine, weights, monkeypatch):
    """BASELINE config 4 at a fifth of its size (200,000 contigs, 8.0 M genes, 31 % of the contigs shorter than the
    window): the two device implementations agree; without padding exactly the genes of short contigs are NaN and
    every other gene keeps the padded run's value bit for bit; the oracle on a strided sample of contigs."""
    from gecco_b200 import synth

    batch = synth.config4(len(weights.attrs), contigs=200_000, mean_domains=6.0)
    lens = numpy.diff(batch.contig_ptr)
    short_gene = numpy.repeat(lens < 20, lens)
    assert 0.25 < (lens < 20).mean() < 0.4
    monkeypatch.setenv("GCRF_FORCE_GENERIC", "0")
    padded = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, pad=True)
    skipped = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, pad=False)
    monkeypatch.setenv("GCRF_FORCE_GENERIC", "1")
    generic = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, pad=True)
    assert numpy.isfinite(padded).all() and padded.min() >= 0.0 and padded.max() <= 1.0
    assert numpy.abs(padded - generic).max() <= TOL
    assert numpy.array_equal(numpy.isnan(skipped), short_gene)
    assert numpy.array_equal(skipped[~short_gene], padded[~short_gene])
    monkeypatch.setenv("GCRF_FORCE_GENERIC", "0")
    for c in range(0, batch.C, 9973):
        one = batch.slice_contigs(c, c + 1)
        g0, g1 = int(batch.contig_ptr[c]), int(batch.contig_ptr[c + 1])
        assert_close(padded[g0:g1], oracle(weights, one, nthreads=1), what=f"contig {c} ({g1 - g0} genes)")


def test_f32_output_and_int64_pointers(engine, weights):
    from gecco_b200 import synth

    batch = synth.ragged_edge_cases(len(weights.attrs))
    want = oracle(weights, batch)
    got32 = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, f32=True)
    assert got32.dtype == numpy.float32
    assert_close(got32.astype(numpy.float64), want, what="f32 out")
    lib_flags_ptr64 = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr.astype(numpy.int64), batch.attr_idx)
    assert_close(lib_flags_ptr64, want, what="int64-able gene_ptr")


def test_dense_config2_slice(engine, weights):
    """BASELINE config 2 shape (Poisson(200) genes x Poisson(25) domains, 5 % unknown ids), 300 contigs."""
    from gecco_b200 import synth

    batch = synth.config2(len(weights.attrs), contigs=300)
    want = oracle(weights, batch)
    got = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx)
    err = assert_close(got, want, what="config2 slice")
    print(f"config2 slice ({batch.G} genes, nnz {batch.nnz}) max |dp| = {err:.3e}")


def test_tiles_larger_than_one_staging_round(engine, weights):
    """~40 domains per gene: a tile's ids exceed the streaming kernel's staging buffer -> direct path."""
    from gecco_b200 import synth

    batch = synth.config2(len(weights.attrs), contigs=40, mean_domains=40.0)
    assert_close(engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx), oracle(weights, batch),
                 what="oversized tiles")


def test_tiles_full_of_tiny_contigs(engine, weights):
    """Contigs of 1-3 genes: every 236-gene tile holds > 127 contigs (the streaming kernel's wide contig slice and
    unbounded contig search), every window is a padded one (or skipped without padding)."""
    from gecco_b200 import synth

    rng = numpy.random.default_rng(21)
    lens = rng.integers(1, 4, size=3000)
    lens[::97] = 25  # a few regular contigs in between
    batch = synth.make_batch(rng, lens, 3.0, len(weights.attrs), 0.05)
    for pad in (True, False):
        assert_close(engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, pad=pad),
                     oracle(weights, batch, pad=pad), what=f"tiny contigs pad={pad}")


def test_sparse_ecoli_like(engine, weights):
    from gecco_b200 import synth

    batch = synth.config3_ecoli_like(len(weights.attrs))
    assert_close(engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx), oracle(weights, batch),
                 what="config3")


def test_metagenome_short_contigs(engine, weights):
    """BASELINE config 4 shape: lognormal contig lengths, ~31 % shorter than the window -> padding path."""
    from gecco_b200 import synth

    batch = synth.config4(len(weights.attrs), contigs=4000, mean_domains=6.0)
    for pad in (True, False):
        assert_close(engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, pad=pad),
                     oracle(weights, batch, pad=pad), what=f"config4 pad={pad}")


def test_long_contigs_windowed_and_chain(engine, weights):
    """BASELINE config 5 shape (scaled): windows are per-window parallel; chain = whole-contig marginals."""
    from gecco_b200 import synth
    from oracle import crf_oracle

    batch = synth.config5(len(weights.attrs), contigs=6, genes=5000)
    assert_close(engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx), oracle(weights, batch),
                 what="config5 windowed")
    got = engine.marginals_chain(batch.contig_ptr, batch.gene_ptr, batch.attr_idx)
    for c in range(batch.C):
        g0, g1 = int(batch.contig_ptr[c]), int(batch.contig_ptr[c + 1])
        want = crf_oracle.chain_marginals(weights.state_w, weights.trans_w, batch.gene_ptr, batch.attr_idx, g0, g1)[:, 1]
        assert_close(got[g0:g1], want, tol=1e-9, what=f"chain contig {c}")


def test_chain_ragged(engine, weights):
    from gecco_b200 import synth
    from oracle import crf_oracle

    batch = synth.ragged_edge_cases(len(weights.attrs))
    got = engine.marginals_chain(batch.contig_ptr, batch.gene_ptr, batch.attr_idx)
    for c in range(batch.C):
        g0, g1 = int(batch.contig_ptr[c]), int(batch.contig_ptr[c + 1])
        want = crf_oracle.chain_marginals(weights.state_w, weights.trans_w, batch.gene_ptr, batch.attr_idx, g0, g1)[:, 1]
        assert_close(got[g0:g1], want, tol=1e-9, what=f"chain contig {c} (n={g1 - g0})")


@pytest.mark.parametrize("team", ["warp", "block"])
def test_chain_teams_agree(engine, weights, monkeypatch, team):
    """The chain primitive runs one warp or one CTA per contig (chosen by mean contig length): both on both shapes."""
    from gecco_b200 import synth
    from oracle import crf_oracle

    monkeypatch.setenv("GCRF_CHAIN_TEAM", team)
    rng = numpy.random.default_rng(17)
    lens = numpy.array([1, 2, 31, 32, 33, 255, 256, 257, 1000, 5003, 8, 700])
    for batch in (synth.ragged_edge_cases(len(weights.attrs)), synth.make_batch(rng, lens, 6.0, len(weights.attrs), 0.05)):
        got = engine.marginals_chain(batch.contig_ptr, batch.gene_ptr, batch.attr_idx)
        for c in range(batch.C):
            g0, g1 = int(batch.contig_ptr[c]), int(batch.contig_ptr[c + 1])
            want = crf_oracle.chain_marginals(weights.state_w, weights.trans_w, batch.gene_ptr, batch.attr_idx, g0, g1)[:, 1]
            assert_close(got[g0:g1], want, tol=1e-9, what=f"{team} team, contig {c} (n={g1 - g0})")


def test_window_equal_to_contig_is_the_chain(engine, weights):
    """Size-independent property: one window covering a whole contig == the chain primitive."""
    from gecco_b200 import synth

    rng = numpy.random.default_rng(3)
    batch = synth.make_batch(rng, numpy.full(50, 20), 5.0, len(weights.attrs))
    a = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, window=20)
    b = engine.marginals_chain(batch.contig_ptr, batch.gene_ptr, batch.attr_idx)
    assert numpy.abs(a - b).max() <= TOL


def test_extreme_unaries_stay_finite(engine, weights):
    """Genes with every strongly positive / strongly negative attribute: |delta| far beyond the FP32 clamp."""
    from gecco_b200.synth import CsrBatch

    d = weights.state_w[:, 1] - weights.state_w[:, 0]
    # strongest attributes whose summed |scores| stay below ~300, i.e. inside f64 exp() range (CRFsuite
    # itself overflows beyond 709) but an order of magnitude past the device's FP32 clamp
    order = numpy.argsort(-d)
    pos = order[:int(numpy.searchsorted(numpy.cumsum(numpy.abs(weights.state_w[order]).max(axis=1)), 300.0))].astype(numpy.int32)
    order = numpy.argsort(d)
    neg = order[:int(numpy.searchsorted(numpy.cumsum(numpy.abs(weights.state_w[order]).max(axis=1)), 300.0))].astype(numpy.int32)
    assert d[pos].sum() > 150 and d[neg].sum() < -60
    rows = []
    for k in range(60):
        rows.append(pos if k % 2 == 0 else neg)  # adversarial alternation
    for k in range(40):
        rows.append(pos)
    for k in range(40):
        rows.append(numpy.zeros(0, dtype=numpy.int32))
    gene_ptr = numpy.cumsum([0] + [len(r) for r in rows]).astype(numpy.int32)
    batch = CsrBatch(numpy.array([0, len(rows)], dtype=numpy.int32), gene_ptr, numpy.concatenate(rows))
    want = oracle(weights, batch)
    got = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx)
    assert numpy.isfinite(got).all()
    assert_close(got, want, what="extreme unaries")


def test_rows_longer_than_the_fixed_point_guard(engine, weights):
    """Genes with hundreds of distinct domains (the reference fixture has 834 raw rows in one gene): the
    streaming kernel's wrapping int32 row sums hand such rows to its float path."""
    from gecco_b200.synth import CsrBatch

    rng = numpy.random.default_rng(9)
    # weakest attributes first, so that even 1,200 of them keep |score| inside f64 exp() range (CRFsuite
    # itself returns NaN beyond that)
    weakest = numpy.argsort(numpy.abs(weights.state_w).max(axis=1))
    rows = []
    for k in range(300):
        n = [0, 3, 25, 105, 106, 107, 150, 400, 834, 1200][k % 10]
        pool = weakest[:max(n, 60) + 40] if n > 25 else numpy.arange(len(weights.attrs))
        rows.append(numpy.sort(rng.choice(pool, size=n, replace=False)).astype(numpy.int32))
    gene_ptr = numpy.cumsum([0] + [len(r) for r in rows]).astype(numpy.int32)
    batch = CsrBatch(numpy.array([0, 120, 300], dtype=numpy.int32), gene_ptr, numpy.concatenate(rows))
    want = oracle(weights, batch)
    got = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx)
    assert_close(got, want, what="huge rows")


def test_empty_batch_and_argument_errors(engine, weights):
    from gecco_b200._lib import GcrfError

    z = numpy.zeros(1, dtype=numpy.int32)
    out = engine.marginals_windowed(z, z, numpy.zeros(0, dtype=numpy.int32))
    assert out.shape == (0,)
    one = numpy.array([0, 1], dtype=numpy.int32)
    for window, step in [(0, 1), (5, 0), (5, 6)]:
        with pytest.raises(GcrfError):
            engine.marginals_windowed(one, numpy.array([0, 0], dtype=numpy.int32), numpy.zeros(0, dtype=numpy.int32),
                                      window=window, step=step)
    with pytest.raises(GcrfError):  # empty contig
        engine.marginals_windowed(numpy.array([0, 0, 1], dtype=numpy.int32), numpy.array([0, 0], dtype=numpy.int32),
                                  numpy.zeros(0, dtype=numpy.int32))


def test_device_pointer_mode_matches_host_mode(engine, weights):
    import torch
    from gecco_b200 import synth

    batch = synth.config2(len(weights.attrs), contigs=50)
    host = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx)
    dev = torch.device("cuda:0")
    cp = torch.from_numpy(batch.contig_ptr).to(dev)
    gp = torch.from_numpy(batch.gene_ptr).to(dev)
    ai = torch.from_numpy(batch.attr_idx).to(dev)
    out = torch.empty(batch.G, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    assert engine.last_kernel_ms() < 0  # timing is off by default
    engine.set_timing(True)
    engine.marginals_windowed_device(cp.data_ptr(), gp.data_ptr(), ai.data_ptr(), batch.C, batch.G, batch.nnz, out.data_ptr())
    engine.synchronize()
    assert numpy.array_equal(out.cpu().numpy(), host)
    assert engine.last_kernel_ms() > 0
    engine.set_timing(False)
    assert engine.launch_count > 0


def test_compact_uint16_ids_are_bit_identical(engine, weights):
    """GCRF_FLAG_IDX_U16: uint16 ids (0xFFFF = unknown) widened on the device give the same bits, host and device mode."""
    import torch
    from gecco_b200 import synth
    from gecco_b200._lib import GCRF_FLAG_IDX_U16
    from gecco_b200.packer import compact_ids

    for batch in (synth.config2(len(weights.attrs), contigs=120), synth.ragged_edge_cases(len(weights.attrs)),
                  synth.config2(len(weights.attrs), contigs=3, mean_genes=2.0, mean_domains=1.0)):
        want = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx)
        small = compact_ids(batch.attr_idx, len(weights.attrs))
        assert small.dtype == numpy.uint16 and small.nbytes * 2 == batch.attr_idx.nbytes
        assert numpy.array_equal(engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, small), want, equal_nan=True)
        assert numpy.array_equal(engine.marginals_chain(batch.contig_ptr, batch.gene_ptr, small),
                                 engine.marginals_chain(batch.contig_ptr, batch.gene_ptr, batch.attr_idx), equal_nan=True)
        dev = torch.device("cuda:0")
        cp, gp = torch.from_numpy(batch.contig_ptr).to(dev), torch.from_numpy(batch.gene_ptr).to(dev)
        ai = torch.from_numpy(small.view(numpy.int16)).to(dev)
        out = torch.empty(batch.G, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        engine.marginals_windowed_device(cp.data_ptr(), gp.data_ptr(), ai.data_ptr(), batch.C, batch.G, batch.nnz, out.data_ptr(),
                                         flags=GCRF_FLAG_IDX_U16)
        engine.synchronize()
        assert numpy.array_equal(out.cpu().numpy(), want, equal_nan=True)


def test_overlapped_host_slices_are_bit_identical(engine, weights, monkeypatch):
    """The host-buffer path cuts PCIe-bound batches into contig-aligned slices (copies of both directions overlap);
    forced here on small batches: same bits as the one-launch path, for every output type and the skip path."""
    from gecco_b200 import synth

    rng = numpy.random.default_rng(4)
    lens = rng.integers(1, 60, size=700)
    batches = [synth.config2(len(weights.attrs), contigs=300), synth.make_batch(rng, lens, 4.0, len(weights.attrs), 0.05),
               synth.make_batch(rng, numpy.array([5000, 3, 1, 900]), 12.0, len(weights.attrs), 0.0)]
    for batch in batches:
        for kw in (dict(), dict(pad=False), dict(f32=True), dict(window=7, step=2)):
            monkeypatch.setenv("GCRF_HOST_SLICES", "1")
            whole = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, **kw)
            for slices in ("2", "5", "8"):
                monkeypatch.setenv("GCRF_HOST_SLICES", slices)
                cut = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, **kw)
                assert numpy.array_equal(whole, cut, equal_nan=True), (slices, kw)
        monkeypatch.setenv("GCRF_HOST_SLICES", "3")
        p64 = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr.astype(numpy.int64), batch.attr_idx)
        monkeypatch.setenv("GCRF_HOST_SLICES", "1")
        assert numpy.array_equal(p64, engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx), equal_nan=True)


def test_feature_extraction_on_device(engine, mibig, weights):
    """Accession -> attribute id on device with the reference's set semantics: running the marginals on its output
    (unknown / repeated rows marked -1, row pointers untouched) equals the host packer + oracle on the real fixture."""
    from gecco_b200.packer import pack_arrays

    assert engine.has_vocabulary
    dom_ptr, dom_pfam = mibig["dom_ptr"], mibig["dom_pfam"]
    ids = engine.features_from_accessions(dom_pfam, dom_ptr)
    packed = pack_arrays(mibig["gene_contig"], dom_ptr, dom_pfam, weights)
    assert int((ids >= 0).sum()) == packed.nnz == 16261  # 25,446 rows, 19,083 in the vocabulary, 16,261 after the per-gene set
    # same sets per gene, in the same order
    keep = ids >= 0
    gene_of = numpy.repeat(numpy.arange(len(dom_ptr) - 1), numpy.diff(dom_ptr))
    assert numpy.array_equal(ids[keep], packed.attr_idx)
    assert numpy.array_equal(numpy.bincount(gene_of[keep], minlength=len(dom_ptr) - 1), numpy.diff(packed.gene_ptr))
    p = engine.marginals_windowed(packed.contig_ptr, dom_ptr, ids)
    assert_close(p, mibig["ref_loop_prob"], what="device features + marginals vs reference loop")
    # the same in ONE call (GCRF_FLAG_ACCESSIONS): accessions in, marginals out, host and device buffers
    p1 = engine.marginals_windowed(packed.contig_ptr, dom_ptr, dom_pfam, accessions=True)
    assert numpy.array_equal(p1, p)
    assert numpy.array_equal(engine.marginals_chain(packed.contig_ptr, dom_ptr, dom_pfam, accessions=True),
                             engine.marginals_chain(packed.contig_ptr, dom_ptr, ids))
    import torch
    from gecco_b200._lib import GCRF_FLAG_ACCESSIONS

    dev = torch.device("cuda", engine.device)
    with torch.cuda.device(dev):
        cp = torch.from_numpy(numpy.ascontiguousarray(packed.contig_ptr, dtype=numpy.int32)).to(dev)
        gp = torch.from_numpy(numpy.ascontiguousarray(dom_ptr, dtype=numpy.int32)).to(dev)
        ac = torch.from_numpy(numpy.ascontiguousarray(dom_pfam, dtype=numpy.int32)).to(dev)
        out = torch.empty(len(dom_ptr) - 1, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        engine.marginals_windowed_device(cp.data_ptr(), gp.data_ptr(), ac.data_ptr(), len(packed.contig_ptr) - 1, len(dom_ptr) - 1,
                                         len(dom_pfam), out.data_ptr(), flags=GCRF_FLAG_ACCESSIONS)
        engine.synchronize()
        assert numpy.array_equal(out.cpu().numpy(), p)
    with pytest.raises(Exception, match="exclude each other"):
        engine.marginals_windowed(packed.contig_ptr, dom_ptr, dom_pfam.astype(numpy.uint16), accessions=True)
    # repeats and unknown accessions
    acc = numpy.array([109, 109, 99999, 5, 109, 5, 2801], dtype=numpy.int32)
    out = engine.features_from_accessions(acc, numpy.array([0, 5, 7], dtype=numpy.int32))
    ix = weights.attr_index
    assert out.tolist() == [ix["PF00109"], -1, -1, ix["PF00005"], -1, ix["PF00005"], ix["PF02801"]]


def _features_restated(accession, gene_ptr, lut):
    """features.py:13-35 on arrays: a gene's features are a dict keyed by domain name, so the first row of each
    accession keeps the id; the tagger drops names its dictionary does not hold."""
    out = numpy.full(len(accession), -1, dtype=numpy.int32)
    for g in range(len(gene_ptr) - 1):
        seen = set()
        for p in range(int(gene_ptr[g]), int(gene_ptr[g + 1])):
            a = int(accession[p])
            if 0 <= a < len(lut) and lut[a] >= 0 and a not in seen:
                out[p] = lut[a]
                seen.add(a)
    return out


@pytest.mark.parametrize("mean_rows", [0.7, 1.4, 5.0, 9.0, 25.0])
def test_feature_extraction_fuzz(engine, weights, mean_rows, monkeypatch):
    """Ragged genes with repeats (inside one 32-row step, across steps and across batches), unknown and negative
    accessions, empty genes, genes of hundreds of rows: both bitmap kernels (sparse / dense grouping), both pointer
    widths and the gene-by-gene kernel against the row-by-row restatement."""
    from gecco_b200.packer import pfam_lut

    lut = pfam_lut(weights.attrs)
    known = numpy.flatnonzero(lut >= 0)
    rng = numpy.random.default_rng(int(mean_rows * 10))
    for G in (1, 31, 32, 33, 700, 5003):
        rows = rng.poisson(mean_rows, size=G)
        rows[rng.random(G) < 0.02] = rng.integers(33, 700)  # a few long genes
        if G > 40:
            rows[7:19] = 0
        gene_ptr = numpy.concatenate([[0], numpy.cumsum(rows)]).astype(numpy.int64)
        n = int(gene_ptr[-1])
        acc = rng.choice(known, size=n).astype(numpy.int32)
        small = rng.choice(known, size=12)  # a small pool makes repeats common
        pick = rng.random(n) < 0.3
        acc[pick] = rng.choice(small, size=int(pick.sum()))
        acc[rng.random(n) < 0.05] = 3  # PF00003 is not in the model
        acc[rng.random(n) < 0.01] = -7
        acc[rng.random(n) < 0.01] = len(lut) + 11
        want = _features_restated(acc, gene_ptr, lut)
        for simple in ("0", "1", "2", "3", "4"):  # default, gene by gene, table in global / shared memory, exact sparse bitmaps
            monkeypatch.setenv("GCRF_FEATURES_SIMPLE", simple)
            for ptr64 in (False, True):
                got = engine.features_from_accessions(acc, gene_ptr, ptr64=ptr64)
                assert numpy.array_equal(got, want), (G, simple, ptr64, numpy.flatnonzero(got != want)[:10])
        monkeypatch.setenv("GCRF_FEATURES_SIMPLE", "0")
        # a second call on the same handle: the kernels leave their bitmaps clean
        assert numpy.array_equal(engine.features_from_accessions(acc, gene_ptr), want)


@pytest.mark.parametrize("feature_type", ["protein", "domain"])
def test_table_path_with_feature_extraction_on_device(weights, mibig, feature_type, monkeypatch):
    """FeatureTables.predict hands raw Pfam numbers to the device (GCRF_FLAG_ACCESSIONS) when the vocabulary allows:
    same probabilities, bit for bit, as packing attribute ids on the host — on a synthetic table with repeated,
    unknown and foreign domain names."""
    import random

    from gecco_b200.crf import ClusterCRF
    from gecco_b200.tables import FeatureTables

    rng = random.Random(5)
    names = list(weights.attrs[:200]) + ["PF99999", "TIGR00001"]
    glines, flines = ["sequence_id\tprotein_id\tstart\tend\tstrand"], [
        "sequence_id\tprotein_id\tstart\tend\tstrand\tdomain\thmm\ti_evalue\tpvalue\tdomain_start\tdomain_end"]
    for c in range(40):
        pos = 1
        for k in range(rng.randrange(1, 90)):
            start, end = pos + 10, pos + 10 + 3 * rng.randrange(30, 300)
            pos = end
            head = f"c{c:03d}\tc{c:03d}_{k + 1}\t{start}\t{end}\t+"
            glines.append(head)
            for _ in range(rng.choice([0, 1, 1, 2, 5])):
                ds = rng.randrange(1, 200)
                flines.append(f"{head}\t{rng.choice(names)}\tPfam\t1e-20\t1e-22\t{ds}\t{ds + 50}")
    crf = ClusterCRF.trained(None)
    crf.feature_type = feature_type
    with FeatureTables.parse(("\n".join(glines) + "\n").encode(), [("\n".join(flines) + "\n").encode()]) as tables:
        monkeypatch.setenv("GECCO_B200_HOST_FEATURES", "1")
        host = tables.predict(crf)
        assert not tables._packed.accessions
        monkeypatch.setenv("GECCO_B200_HOST_FEATURES", "0")
        device = tables.predict(crf)
        assert tables._packed.accessions
    assert numpy.array_equal(host, device, equal_nan=True)


def test_peer_output_arrays_single_gpu(engine, weights, device_path):
    """gcrf_marginals_windowed_peers with the "peer" arrays on this GPU: the local array (optional) and every peer array
    receive the result at out_offset; paths other than the streaming kernel refuse (callers then gather with a collective)."""
    import torch

    from gecco_b200 import synth
    from gecco_b200._lib import GcrfError

    batch = synth.config4(len(weights.attrs), contigs=3000, mean_domains=6.0)
    want = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx)
    dev = torch.device("cuda", 0)
    cp, gp = torch.from_numpy(batch.contig_ptr).to(dev), torch.from_numpy(batch.gene_ptr).to(dev)
    ai = torch.full((batch.nnz + 16,), -1, dtype=torch.int32, device=dev)
    ai[: batch.nnz] = torch.from_numpy(batch.attr_idx).to(dev)
    off = 1234
    for f32 in (False, True):
        dt = torch.float32 if f32 else torch.float64
        local = torch.full((batch.G,), -1.0, dtype=dt, device=dev)
        peers = [torch.full((batch.G + off + 7,), -1.0, dtype=dt, device=dev) for _ in range(3)]
        call = lambda out, plist: engine.marginals_windowed_peers(cp.data_ptr(), gp.data_ptr(), ai.data_ptr(), batch.C, batch.G, batch.nnz,
                                                                   out, [p.data_ptr() for p in plist], off, window=20, f32=f32)
        if device_path == "generic":
            with pytest.raises(GcrfError, match="streaming kernel"):
                call(local.data_ptr(), peers)
            continue
        call(local.data_ptr(), peers)
        engine.synchronize()
        ref = want.astype(numpy.float32) if f32 else want
        assert numpy.array_equal(local.cpu().numpy(), ref)
        for p in peers:
            got = p.cpu().numpy()
            assert numpy.array_equal(got[off:off + batch.G], ref) and (got[:off] == -1).all() and (got[off + batch.G:] == -1).all()
        # no local array at all
        peers[0].fill_(-1.0)
        call(None, peers[:1])
        engine.synchronize()
        assert numpy.array_equal(peers[0].cpu().numpy()[off:off + batch.G], ref)


@pytest.mark.parametrize("slots", ["128", "64"])
def test_half_and_quarter_tiles(engine, weights, slots, monkeypatch, device_path):
    """The streaming kernel's half- and quarter-tile variants (chosen by density; forced here): every control-flow shape
    of the full-tile kernel again, plus batches dense enough to need them."""
    from gecco_b200 import synth

    if device_path == "generic":
        pytest.skip("streaming kernel only")
    monkeypatch.setenv("GCRF_STREAM_SLOTS", slots)
    batch = synth.ragged_edge_cases(len(weights.attrs))
    for window, step, pad in ((20, 1, True), (20, 3, False), (5, 1, True), (5, 2, False)):
        got = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, window=window, step=step, pad=pad)
        assert_close(got, oracle(weights, batch, window, step, pad), what=f"slots={slots} W={window} step={step} pad={pad}")
    for shape in (synth.config2(len(weights.attrs), contigs=150), synth.config4(len(weights.attrs), contigs=3000, mean_domains=6.0)):
        assert_close(engine.marginals_windowed(shape.contig_ptr, shape.gene_ptr, shape.attr_idx), oracle(weights, shape), what=f"slots={slots}")
    rng = numpy.random.default_rng(8)
    tiny = synth.make_batch(rng, rng.integers(1, 4, size=2000), 3.0, len(weights.attrs), 0.05)
    assert_close(engine.marginals_windowed(tiny.contig_ptr, tiny.gene_ptr, tiny.attr_idx), oracle(weights, tiny), what="tiny contigs")


def test_dense_batches_pick_smaller_tiles(engine, weights, monkeypatch, device_path):
    """40 and 90 ids per gene on average: a full tile's ids no longer fit one staging round; the plan switches to half /
    quarter tiles by itself and the result is the oracle's."""
    from gecco_b200 import synth

    monkeypatch.delenv("GCRF_STREAM_SLOTS", raising=False)
    for d in (40.0, 90.0):
        batch = synth.config2(len(weights.attrs), contigs=60, mean_domains=d)
        assert_close(engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx), oracle(weights, batch), what=f"d={d}")
