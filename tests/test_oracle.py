"""Pins the CPU oracle: the reference's golden vector, the survey's KATs, the reference's own loop."""
import numpy
import pytest

from oracle import crf_oracle
from conftest import pack_case


def _bgc_csr(bgc, weights, n_genes=None):
    from gecco_b200.packer import pack_records

    doms = {}
    for d in bgc["domains"]:
        if d["pvalue"] < 1e-9:
            doms.setdefault(d["protein_id"], []).append((d["domain_start"], d["domain"]))
    genes = bgc["genes"][:n_genes]
    records = [(g["sequence_id"], g["start"], g["protein_id"], [n for _, n in sorted(doms.get(g["protein_id"], []))])
               for g in genes]
    return pack_records(records, weights.attr_index), genes


def test_golden_bgc0001866(bgc, weights):
    """KAT E: python-crfsuite's own numbers (tests/test_cli/data/BGC0001866.genes.tsv)."""
    packed, genes = _bgc_csr(bgc, weights)
    assert packed.G == 23 and packed.nnz == 35  # 37 rows, PF00550 x3 in one gene collapses (set semantics)
    p, windows = crf_oracle.marginals_windowed(weights.state_w, weights.trans_w, weights.label_id("1"),
                                               packed.contig_ptr, packed.gene_ptr, packed.attr_idx, 20, 1, True)
    assert windows == 4
    golden = numpy.array([g["average_p"] for g in genes])[packed.order]
    assert numpy.abs(p - golden).max() < 1e-12
    # the per-domain column of features.tsv repeats the same numbers
    by_gene = {g["protein_id"]: x for g, x in zip([genes[i] for i in packed.order], p)}
    for d in bgc["domains"]:
        assert abs(by_gene[d["protein_id"]] - d["cluster_probability"]) < 1e-12


def test_kat_a_chain_primitive(weights):
    ix = weights.attr_index
    items = [[ix["PF00109"], ix["PF02801"]], [], [ix["PF00005"]]]
    ptr = numpy.cumsum([0] + [len(i) for i in items])
    idx = numpy.array([a for i in items for a in i], dtype=numpy.int32)
    m = crf_oracle.chain_marginals(weights.state_w, weights.trans_w, ptr, idx)
    expect = [0.7084717024510778, 0.7053529511853454, 0.7028351630136186]
    assert numpy.abs(m[:, 1] - expect).max() < 1e-13
    assert numpy.abs(m.sum(axis=1) - 1).max() < 1e-13
    twin = crf_oracle.chain_marginals_numpy(weights.state_w, weights.trans_w, items)
    assert numpy.abs(twin - m).max() < 1e-14


def test_kat_b_padding(bgc, weights):
    packed, _ = _bgc_csr(bgc, weights, n_genes=7)
    p, windows = crf_oracle.marginals_windowed(weights.state_w, weights.trans_w, 1, packed.contig_ptr,
                                               packed.gene_ptr, packed.attr_idx, 20, 1, True)
    expect = [0.6744886046494636, 0.6758759268440484, 0.6778253989121013, 0.680357377868751,
              0.6834983034536436, 0.6872809742232185, 0.679995045588538]
    assert windows == 1
    assert numpy.abs(p - expect).max() < 1e-13


def test_kat_c_d_empty_chains(weights):
    ptr = numpy.zeros(21, dtype=numpy.int64)
    m = crf_oracle.chain_marginals(weights.state_w, weights.trans_w, ptr, numpy.zeros(0, dtype=numpy.int32))
    assert abs(m[0, 1] - 0.15497426949187523) < 1e-13
    assert abs(m[1, 1] - 0.15149818778666058) < 1e-13
    assert abs(m[2, 1] - 0.14849454367489034) < 1e-13
    assert abs(m[10, 1] - 0.13861536218626613) < 1e-13
    one = crf_oracle.chain_marginals(weights.state_w, weights.trans_w, numpy.array([0, 0]), numpy.zeros(0, dtype=numpy.int32))
    assert abs(one[0, 1] - 0.5) < 1e-15
    a = weights.attr_index["PF00109"]
    one = crf_oracle.chain_marginals(weights.state_w, weights.trans_w, numpy.array([0, 1]), numpy.array([a], dtype=numpy.int32))
    assert abs(one[0, 1] - 0.6584367528343128) < 1e-13


def test_kat_f_full_chain_is_not_the_windowed_result(bgc, weights):
    packed, genes = _bgc_csr(bgc, weights)
    full = crf_oracle.chain_marginals(weights.state_w, weights.trans_w, packed.gene_ptr, packed.attr_idx)[:, 1]
    win, _ = crf_oracle.marginals_windowed(weights.state_w, weights.trans_w, 1, packed.contig_ptr, packed.gene_ptr,
                                           packed.attr_idx, 20, 1, True)
    d = numpy.abs(full - win).max()
    assert 0 < d < 1e-9


def test_reference_loop_cases(ref_cases, weights):
    """Oracle vs the REFERENCE'S OWN ClusterCRF.predict_probabilities (tools/make_golden.py harness)."""
    for case in ref_cases:
        packed = pack_case(case, weights)
        p, _ = crf_oracle.marginals_windowed(weights.state_w, weights.trans_w, 1, packed.contig_ptr, packed.gene_ptr,
                                             packed.attr_idx, case["window"], case["step"], case["pad"])
        expected = case["expected"]
        assert [packed.gene_ids[i] for i in range(packed.G)] == [e["id"] for e in expected], case["name"]
        for x, e in zip(p, expected):
            if e["p"] is None:
                assert numpy.isnan(x), case["name"]
            else:
                assert abs(x - e["p"]) < 1e-12, case["name"]


def test_reference_loop_mibig(mibig, weights):
    """15,158 real genes: oracle vs the reference's loop driven through the fake tagger."""
    from gecco_b200.packer import pack_arrays

    packed = pack_arrays(mibig["gene_contig"], mibig["dom_ptr"], mibig["dom_pfam"], weights)
    assert packed.C == 18 and packed.G == 15158
    p, windows = crf_oracle.marginals_windowed(weights.state_w, weights.trans_w, 1, packed.contig_ptr, packed.gene_ptr,
                                               packed.attr_idx, 20, 1, True, nthreads=4)
    assert windows == 14816
    assert numpy.abs(p - mibig["ref_loop_prob"]).max() < 1e-12


def test_numpy_twin_agrees_on_ragged_batch(weights):
    from gecco_b200 import synth

    batch = synth.ragged_edge_cases(len(weights.attrs))
    sub = batch.slice_contigs(0, 40)
    for window, step, pad in [(20, 1, True), (20, 1, False), (5, 2, True), (7, 7, True)]:
        a, _ = crf_oracle.marginals_windowed(weights.state_w, weights.trans_w, 1, sub.contig_ptr, sub.gene_ptr,
                                             sub.attr_idx, window, step, pad)
        b = crf_oracle.marginals_windowed_numpy(weights.state_w, weights.trans_w, 1, sub.contig_ptr, sub.gene_ptr,
                                                sub.attr_idx, window, step, pad)
        assert numpy.array_equal(numpy.isnan(a), numpy.isnan(b))
        assert numpy.nanmax(numpy.abs(a - b)) < 1e-13


def test_threads_do_not_change_results(weights):
    from gecco_b200 import synth

    batch = synth.ragged_edge_cases(len(weights.attrs))
    a, wa = crf_oracle.marginals_windowed(weights.state_w, weights.trans_w, 1, batch.contig_ptr, batch.gene_ptr,
                                          batch.attr_idx, 20, 1, True, nthreads=1)
    b, wb = crf_oracle.marginals_windowed(weights.state_w, weights.trans_w, 1, batch.contig_ptr, batch.gene_ptr,
                                          batch.attr_idx, 20, 1, True, nthreads=5)
    assert wa == wb == batch.windows(20)
    assert numpy.array_equal(a, b)


def test_window_rules(weights):
    z = numpy.zeros(2, dtype=numpy.int64)
    for window, step in [(0, 1), (5, 0), (5, 6)]:
        with pytest.raises(ValueError):
            crf_oracle.marginals_windowed(weights.state_w, weights.trans_w, 1, z, z, numpy.zeros(0, dtype=numpy.int32), window, step, True)
