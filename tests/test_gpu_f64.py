"""GCRF_FLAG_F64 — the reference's own arithmetic on the device (gcrf_exact.cu): CRFsuite's scaled forward-backward in
f64, operation by operation.  Bars: bit-identical to python-crfsuite on the reference's golden fixture; <= 1e-12 against
the CPU oracle everywhere (what is left is the host libm's rounding of exp: the device exponential is correctly
rounded, glibc's is not always)."""
import numpy
import pytest

from conftest import pack_case

pytestmark = pytest.mark.gpu

TOL64 = 1e-12


def oracle(weights, batch, window=20, step=1, pad=True):
    from oracle import crf_oracle

    p, _ = crf_oracle.marginals_windowed(weights.state_w, weights.trans_w, weights.label_id("1"), batch.contig_ptr,
                                         batch.gene_ptr, batch.attr_idx, window, step, pad, nthreads=8)
    return p


def close(got, want, tol=TOL64):
    assert numpy.array_equal(numpy.isnan(got), numpy.isnan(want))
    ok = ~numpy.isnan(want)
    return (float(numpy.abs(got[ok] - want[ok]).max()) if ok.any() else 0.0) <= tol


def test_golden_bgc0001866_bit_identical(engine, bgc, weights):
    from test_oracle import _bgc_csr

    packed, genes = _bgc_csr(bgc, weights)
    p = engine.marginals_windowed(packed.contig_ptr, packed.gene_ptr, packed.attr_idx, f64_arith=True)
    golden = numpy.array([g["average_p"] for g in genes])[packed.order]
    assert numpy.array_equal(p, golden), numpy.abs(p - golden).max()


def test_reference_loop_cases(engine, ref_cases, weights):
    for case in ref_cases:
        packed = pack_case(case, weights)
        p = engine.marginals_windowed(packed.contig_ptr, packed.gene_ptr, packed.attr_idx, window=case["window"],
                                      step=case["step"], pad=case["pad"], f64_arith=True)
        want = numpy.array([numpy.nan if e["p"] is None else e["p"] for e in case["expected"]])
        assert close(p, want), case["name"]


def test_mibig_real_features(engine, mibig, weights):
    from gecco_b200.packer import pack_arrays

    packed = pack_arrays(mibig["gene_contig"], mibig["dom_ptr"], mibig["dom_pfam"], weights)
    p = engine.marginals_windowed(packed.contig_ptr, packed.gene_ptr, packed.attr_idx, f64_arith=True)
    want = mibig["ref_loop_prob"]
    assert close(p, want)
    print(f"mibig: bit-identical to the reference loop on {(p == want).mean():.2%} of {len(p)} genes, max|dp| {numpy.abs(p - want).max():.2e}")


@pytest.mark.parametrize("window,step,pad", [(20, 1, True), (20, 1, False), (20, 3, True), (5, 1, True), (5, 2, False),
                                             (10, 3, False), (7, 7, True), (1, 1, True), (33, 4, True), (64, 1, True),
                                             (128, 5, True), (200, 1, True), (700, 13, False), (1000, 1, True)])
def test_ragged_edge_cases_any_window(engine, weights, window, step, pad):
    """Every branch of the window logic, incl. windows far beyond the FP32 kernels' limit of 128 (the reference accepts
    any window_size >= 1, gecco/crf/__init__.py:134-137)."""
    from gecco_b200 import synth

    batch = synth.ragged_edge_cases(len(weights.attrs))
    got = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, window=window, step=step, pad=pad,
                                    f64_arith=True)
    assert close(got, oracle(weights, batch, window, step, pad)), (window, step, pad)


def test_float_output_int64_pointers_and_device_pointers(engine, weights):
    import torch

    from gecco_b200 import synth
    from gecco_b200._lib import GCRF_FLAG_F64

    batch = synth.ragged_edge_cases(len(weights.attrs))
    want = oracle(weights, batch)
    got32 = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, f32=True, f64_arith=True)
    assert got32.dtype == numpy.float32 and numpy.array_equal(got32, want.astype(numpy.float32), equal_nan=True)
    assert close(engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr.astype(numpy.int64), batch.attr_idx, f64_arith=True), want)
    dev = torch.device("cuda", 0)
    cp, gp, ai = (torch.from_numpy(numpy.ascontiguousarray(a)).to(dev) for a in (batch.contig_ptr, batch.gene_ptr, batch.attr_idx))
    out = torch.full((batch.G,), -1.0, dtype=torch.float64, device=dev)
    engine.marginals_windowed_device(cp.data_ptr(), gp.data_ptr(), ai.data_ptr(), batch.C, batch.G, batch.nnz, out.data_ptr(),
                                     window=20, step=1, pad=True, flags=GCRF_FLAG_F64)
    engine.synchronize()
    assert close(out.cpu().numpy(), want)


def test_extreme_unaries(engine, weights):
    """State scores far outside what FP32 odds can hold: in f64 nothing is clamped, the result is the oracle's."""
    from gecco_b200.synth import CsrBatch

    order = numpy.argsort(weights.state_w[:, 1] - weights.state_w[:, 0])
    neg, pos = order[:60].astype(numpy.int32), order[-60:].astype(numpy.int32)
    rows = [pos if g % 7 == 0 else neg if g % 5 == 0 else pos[:3] for g in range(120)]
    gene_ptr = numpy.cumsum([0] + [len(r) for r in rows]).astype(numpy.int32)
    batch = CsrBatch(numpy.array([0, 50, 120], dtype=numpy.int32), gene_ptr, numpy.concatenate(rows).astype(numpy.int32))
    got = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, f64_arith=True)
    assert close(got, oracle(weights, batch))


def test_device_exp_is_correctly_rounded(engine, weights):
    """The double-double exponential of gcrf_exact.cu against Python's decimal module (60 digits) on the state scores
    of a real batch: the unary pass leaves exp(s_0), exp(s_1) of every gene behind; window = 1 turns them into the
    marginal E_1 / (E_0 + E_1) of a one-item chain, which the oracle computes from libm's exp."""
    from gecco_b200 import synth

    batch = synth.config2(len(weights.attrs), contigs=20)
    got = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, window=1, f64_arith=True)
    want = oracle(weights, batch, window=1)
    assert close(got, want, 1e-15)
    assert (got == want).mean() > 0.98  # glibc's exp misses the correctly rounded value on < 0.1 % of arguments
