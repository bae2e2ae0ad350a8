"""BASELINE configs at their FULL sizes, every gene against the CPU oracle (no sampling): the oracle runs at 10-20 M
genes/s on the box's host cores, so even the 40 M-gene metagenome is seconds of CPU time.  FP32 device arithmetic
against the 1e-5 bar of the north star, and the reference-order f64 path (GCRF_FLAG_F64) against 1e-12."""
import os

import numpy
import pytest

pytestmark = pytest.mark.gpu

TOL, TOL64 = 1e-5, 1e-12


def oracle(weights, batch, window=20, step=1, pad=True):
    from oracle import crf_oracle

    p, _ = crf_oracle.marginals_windowed(weights.state_w, weights.trans_w, weights.label_id("1"), batch.contig_ptr,
                                         batch.gene_ptr, batch.attr_idx, window, step, pad,
                                         nthreads=len(os.sched_getaffinity(0)))
    return p


@pytest.fixture(scope="module")
def metagenome(weights):
    from gecco_b200 import synth

    return synth.config4(len(weights.attrs), mean_domains=1.4)


def max_err(got, want):
    assert got.shape == want.shape
    assert numpy.array_equal(numpy.isnan(got), numpy.isnan(want))
    ok = ~numpy.isnan(want)
    return float(numpy.abs(got[ok] - want[ok]).max()) if ok.any() else 0.0


def test_config2_full_batch_f32_and_f64(engine, weights):
    from gecco_b200 import synth

    batch = synth.config2(len(weights.attrs))
    assert (batch.C, batch.G, batch.nnz) == (10_000, 2_000_810, 49_793_052)
    want = oracle(weights, batch)
    e32 = max_err(engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx), want)
    got64 = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, f64_arith=True)
    e64 = max_err(got64, want)
    print(f"config 2 full: f32 max|dp| {e32:.3e}, f64 max|dp| {e64:.3e}, f64 bit-identical genes {(got64 == want).mean():.4%}")
    assert e32 <= TOL and e64 <= TOL64


def test_config3_ecoli_like_full(engine, weights):
    from gecco_b200 import synth

    batch = synth.config3_ecoli_like(len(weights.attrs))
    want = oracle(weights, batch)
    assert max_err(engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx), want) <= TOL
    assert max_err(engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, f64_arith=True), want) <= TOL64


@pytest.mark.parametrize("pad", [True, False])
def test_config4_metagenome_full_batch(engine, weights, metagenome, pad):
    """1,000,000 contigs / 39,973,222 genes / ~31 % of the contigs shorter than the window (seed 4), at the
    real annotation density (1.4 domains per gene) so that the id array stays a few hundred MB; the dense variant
    (25 domains per gene, 4 GB of ids) runs in bench.py's `configs` leg."""
    batch = metagenome
    assert (batch.C, batch.G) == (1_000_000, 39_973_222)
    lens = numpy.diff(batch.contig_ptr)
    assert 0.25 < (lens < 20).mean() < 0.4
    want = oracle(weights, batch, pad=pad)
    got = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, pad=pad)
    e32 = max_err(got, want)
    print(f"config 4 full (pad={pad}): f32 max|dp| {e32:.3e}")
    assert e32 <= TOL
    if pad:
        e64 = max_err(engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, f64_arith=True), want)
        print(f"config 4 full: f64 max|dp| {e64:.3e}")
        assert e64 <= TOL64


def test_config5_long_contigs_full_batch(engine, weights):
    """100 contigs x 5,000 genes: (i) GECCO semantics, W = 20, every gene vs the oracle; (ii) the deep-chain primitive
    (one 5,000-item chain per contig) vs the oracle's chain marginals, every gene."""
    from gecco_b200 import synth
    from oracle import crf_oracle

    batch = synth.config5(len(weights.attrs))
    assert (batch.C, batch.G) == (100, 500_000)
    want = oracle(weights, batch)
    assert max_err(engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx), want) <= TOL
    assert max_err(engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, f64_arith=True), want) <= TOL64
    got = engine.marginals_chain(batch.contig_ptr, batch.gene_ptr, batch.attr_idx)
    worst = 0.0
    for c in range(batch.C):
        g0, g1 = int(batch.contig_ptr[c]), int(batch.contig_ptr[c + 1])
        ref = crf_oracle.chain_marginals(weights.state_w, weights.trans_w, batch.gene_ptr, batch.attr_idx, g0, g1)[:, 1]
        worst = max(worst, float(numpy.abs(got[g0:g1] - ref).max()))
    print(f"config 5 chains: max|dp| {worst:.3e}")
    assert worst <= 1e-9
