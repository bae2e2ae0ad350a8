"""The compact wire format (gcrf_wire_*): host encoder / decoder on the CPU, the device decoder + marginals on the GPU."""
import numpy
import pytest

from gecco_b200 import synth
from gecco_b200._lib import WireBatch


def sorted_reference(batch, A):
    """What the wire holds: every gene's ids ascending, ids outside [0, A) replaced by A."""
    ids = numpy.where((batch.attr_idx >= 0) & (batch.attr_idx < A), batch.attr_idx, A).astype(numpy.int64)
    gene_of = numpy.repeat(numpy.arange(batch.G, dtype=numpy.int64), numpy.diff(batch.gene_ptr))
    order = numpy.lexsort((ids, gene_of))
    return ids[order].astype(numpy.int32)


@pytest.mark.parametrize("make", ["ragged", "dense", "empty_genes", "long_rows", "nothing"])
def test_encode_decode_round_trip(weights, make):
    A = len(weights.attrs)
    rng = numpy.random.default_rng(3)
    if make == "ragged":
        batch = synth.ragged_edge_cases(A)
    elif make == "dense":
        batch = synth.config2(A, contigs=60)
    elif make == "empty_genes":
        batch = synth.make_batch(rng, numpy.array([5, 1, 30]), 0.2, A, 0.3)
    elif make == "long_rows":  # rows beyond 255 ids: the length arrays switch to uint16
        batch = synth.make_batch(rng, numpy.array([3, 40]), 400.0, A, 0.05)
        assert numpy.diff(batch.gene_ptr).max() > 255
    else:
        batch = synth.CsrBatch(numpy.zeros(1, dtype=numpy.int32), numpy.zeros(1, dtype=numpy.int32), numpy.zeros(0, dtype=numpy.int32))
    wire = WireBatch(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, A)
    assert (wire.C, wire.G, wire.nnz) == (batch.C, batch.G, batch.nnz)
    gene_ptr, attr_idx = wire.decode()
    assert numpy.array_equal(gene_ptr, batch.gene_ptr)
    assert numpy.array_equal(attr_idx, sorted_reference(batch, A))
    if make == "dense":
        plain = batch.contig_ptr.nbytes + batch.gene_ptr.nbytes + batch.attr_idx.nbytes
        assert wire.nbytes < 0.36 * plain  # ~1.06 bytes per id + 2.5 per gene against 4 + 4
    # int64 row pointers encode to the same block
    again = WireBatch(batch.contig_ptr, batch.gene_ptr.astype(numpy.int64), batch.attr_idx, A)
    assert again.nbytes == wire.nbytes and numpy.array_equal(again.decode()[1], attr_idx)


def test_encoder_rejects_malformed_row_pointers(weights):
    from gecco_b200._lib import GcrfError

    with pytest.raises(GcrfError, match="non-decreasing"):
        WireBatch(numpy.array([0, 2]), numpy.array([0, 3, 2]), numpy.array([1, 2], dtype=numpy.int32), 10)
    with pytest.raises(GcrfError, match="start at 0 and end at nnz"):
        WireBatch(numpy.array([0, 1]), numpy.array([0, 1]), numpy.array([1, 2], dtype=numpy.int32), 10)


@pytest.mark.gpu
@pytest.mark.parametrize("slices", ["1", "3", "8"])
@pytest.mark.parametrize("make", ["ragged", "dense", "sparse", "long_rows"])
def test_wire_call_equals_the_plain_call(engine, weights, make, slices, monkeypatch):
    """Bit for bit in FP32 arithmetic (exact integer row sums: the order inside a row does not matter); the f64 path
    sums in the sorted order, i.e. within a few ulps of the first-occurrence order."""
    A = len(weights.attrs)
    rng = numpy.random.default_rng(5)
    batch = {"ragged": lambda: synth.ragged_edge_cases(A), "dense": lambda: synth.config2(A, contigs=300),
             "sparse": lambda: synth.config2(A, contigs=300, mean_domains=1.4),
             "long_rows": lambda: synth.make_batch(rng, numpy.array([3, 40, 25]), 300.0, A, 0.05)}[make]()
    # unsorted rows on the plain side: the wire sorts, the result must not care
    shuffled = batch.attr_idx.copy()
    for g in range(0, batch.G, 3):
        a, b = int(batch.gene_ptr[g]), int(batch.gene_ptr[g + 1])
        shuffled[a:b] = shuffled[a:b][::-1]
    # the block is cut into contig-aligned slices when it is encoded (the call overlaps their copies): any cut, same result
    monkeypatch.setenv("GCRF_WIRE_SLICES", slices)
    wire = WireBatch(batch.contig_ptr, batch.gene_ptr, shuffled, A)
    for window, step, pad in ((20, 1, True), (5, 2, False), (7, 1, True)):
        plain = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, shuffled, window=window, step=step, pad=pad)
        got = engine.marginals_windowed_wire(wire, window=window, step=step, pad=pad)
        # rows past the fixed-point guard take the float path, and so does every row of the first-generation kernel
        # (window 7 has no streaming instantiation): same value up to summation order
        if make == "long_rows" or window == 7:
            assert numpy.allclose(got, plain, rtol=0, atol=5e-6 if window == 7 else 1e-6, equal_nan=True)
        else:
            assert numpy.array_equal(got, plain, equal_nan=True)
        got32 = engine.marginals_windowed_wire(wire, window=window, step=step, pad=pad, f32=True)
        assert got32.dtype == numpy.float32 and numpy.array_equal(got32, got.astype(numpy.float32), equal_nan=True)
    exact = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, shuffled, f64_arith=True)
    assert numpy.allclose(engine.marginals_windowed_wire(wire, f64_arith=True), exact, rtol=0, atol=1e-12, equal_nan=True)


@pytest.mark.parametrize("slices", ["1", "3", "8"])
@pytest.mark.parametrize("make", ["ragged", "dense", "long_rows"])
def test_round_trip_of_a_sliced_block(weights, make, slices, monkeypatch):
    """The block holds one section per slice (cut at contig starts when it is encoded): the host decoder walks them."""
    A = len(weights.attrs)
    rng = numpy.random.default_rng(11)
    batch = {"ragged": lambda: synth.ragged_edge_cases(A), "dense": lambda: synth.config2(A, contigs=90),
             "long_rows": lambda: synth.make_batch(rng, numpy.array([3, 40, 7, 1]), 400.0, A, 0.05)}[make]()
    monkeypatch.setenv("GCRF_WIRE_SLICES", "1")
    whole = WireBatch(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, A)
    monkeypatch.setenv("GCRF_WIRE_SLICES", slices)
    wire = WireBatch(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, A)
    gene_ptr, attr_idx = wire.decode()
    assert numpy.array_equal(gene_ptr, batch.gene_ptr)
    assert numpy.array_equal(attr_idx, sorted_reference(batch, A))
    assert wire.nbytes <= whole.nbytes + 64 * int(slices) + 16 * (batch.G // 512 + int(slices) + 1)


@pytest.mark.gpu
def test_a_gene_larger_than_the_decoder_staging_area(engine, weights):
    """Rows of 9,000 - 20,000 ids (the decoder stages 8,192 ids / 12 KB of stream per round in shared memory) next to
    ordinary ones."""
    A = len(weights.attrs)
    rng = numpy.random.default_rng(13)
    sizes = numpy.array([3, 9000, 0, 25, 20000, 8192, 8193, 1, 30] + [20] * 40)
    gene_ptr = numpy.concatenate([[0], numpy.cumsum(sizes)]).astype(numpy.int32)
    attr_idx = rng.integers(0, A, size=int(gene_ptr[-1])).astype(numpy.int32)
    contig_ptr = numpy.array([0, 5, 9, len(sizes)], dtype=numpy.int32)
    wire = WireBatch(contig_ptr, gene_ptr, attr_idx, A)
    plain = engine.marginals_windowed(contig_ptr, gene_ptr, attr_idx, window=5, step=1, pad=True)
    got = engine.marginals_windowed_wire(wire, window=5, step=1, pad=True)
    assert numpy.allclose(got, plain, rtol=0, atol=1e-6)
    exact = engine.marginals_windowed(contig_ptr, gene_ptr, attr_idx, window=5, f64_arith=True)
    # (scores of thousands of summed weights overflow exp() in the reference's arithmetic: NaN on both sides)
    assert numpy.allclose(engine.marginals_windowed_wire(wire, window=5, f64_arith=True), exact, rtol=0, atol=1e-12, equal_nan=True)


@pytest.mark.gpu
def test_sliced_wire_call_with_the_kernels_being_timed(engine, weights, monkeypatch):
    """gcrf_model_set_timing: the slices run their three stages one after the other on one stream — same numbers."""
    A = len(weights.attrs)
    batch = synth.config2(A, contigs=200)
    monkeypatch.setenv("GCRF_WIRE_SLICES", "4")
    wire = WireBatch(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, A)
    plain = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx)
    engine.set_timing(True)
    try:
        got = engine.marginals_windowed_wire(wire)
        assert engine.last_kernel_ms() > 0
    finally:
        engine.set_timing(False)
    assert numpy.array_equal(got, plain)
    assert numpy.array_equal(engine.marginals_windowed_wire(wire), plain)


@pytest.mark.gpu
def test_wire_call_with_a_vocabulary_past_three_byte_deltas(weights):
    """A model with more than 2^21 attributes: deltas past the unary part of the code (the 24-bit escape)."""
    import dataclasses

    from gecco_b200._lib import CRFEngine

    A = (1 << 21) + 1000
    rng = numpy.random.default_rng(17)
    state_w = (rng.standard_normal((A, 2)) * 0.3).astype(numpy.float64)
    big = dataclasses.replace(weights, attrs=[f"A{i}" for i in range(A)], state_w=state_w, state_mask=numpy.ones((A, 2), dtype=bool))
    eng = CRFEngine(big, device=0)
    sizes = rng.poisson(6.0, size=84)
    gene_ptr = numpy.concatenate([[0], numpy.cumsum(sizes)]).astype(numpy.int32)
    batch = synth.CsrBatch(numpy.array([0, 30, 34, 84], dtype=numpy.int32), gene_ptr,
                           rng.integers(0, A, size=int(gene_ptr[-1])).astype(numpy.int32))
    for g in range(0, batch.G, 7):  # deltas of four bytes: a first id (delta from 0) and a jump at or above 2^21
        a, b = int(batch.gene_ptr[g]), int(batch.gene_ptr[g + 1])
        if b - a >= 2:
            batch.attr_idx[a], batch.attr_idx[a + 1] = (1 << 21) + 7 + g, 3
        elif b - a == 1:
            batch.attr_idx[a] = (1 << 21) + g
    wire = WireBatch(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, A)
    # (a vocabulary of this size is past the FP32 kernels' shared-memory table: the f64 path takes it)
    plain = eng.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, window=5, step=1, pad=True, f64_arith=True)
    got = eng.marginals_windowed_wire(wire, window=5, step=1, pad=True, f64_arith=True)
    assert numpy.allclose(got, plain, rtol=0, atol=1e-12)
    gene_ptr, attr_idx = wire.decode()
    assert numpy.array_equal(attr_idx, sorted_reference(batch, A))
    eng.close()


@pytest.mark.gpu
def test_wire_call_fuzz(engine, weights, monkeypatch):
    """Random shapes: densities from almost empty to rows past 255 ids (the uint8 / uint16 switch sits inside the
    range), one to hundreds of contigs, every slice count — the FP32 marginals of the wire call equal the plain call's."""
    A = len(weights.attrs)
    rng = numpy.random.default_rng(23)
    for trial in range(36):
        contigs = int(rng.integers(1, 400))
        density = float(rng.choice([0.2, 1.4, 8.0, 25.0, 60.0, 180.0]))
        lens = numpy.maximum(1, rng.poisson(float(rng.choice([1.0, 6.0, 40.0])), size=contigs))
        if density >= 60:
            lens = lens[:40]
        batch = synth.make_batch(rng, lens, density, A, float(rng.choice([0.0, 0.05, 0.5])))
        monkeypatch.setenv("GCRF_WIRE_SLICES", str(int(rng.integers(1, 9))))
        wire = WireBatch(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, A)
        gene_ptr, attr_idx = wire.decode()
        assert numpy.array_equal(gene_ptr, batch.gene_ptr) and numpy.array_equal(attr_idx, sorted_reference(batch, A))
        window = int(rng.choice([5, 20]))
        pad = bool(rng.integers(0, 2))
        plain = engine.marginals_windowed(batch.contig_ptr, batch.gene_ptr, batch.attr_idx, window=window, step=1, pad=pad)
        got = engine.marginals_windowed_wire(wire, window=window, step=1, pad=pad)
        if density >= 60:  # rows past the fixed-point guard: float row sums, order-dependent in the last bit
            assert numpy.allclose(got, plain, rtol=0, atol=1e-6, equal_nan=True), (trial, contigs, density)
        else:
            assert numpy.array_equal(got, plain, equal_nan=True), (trial, contigs, density)
        wire.close()
