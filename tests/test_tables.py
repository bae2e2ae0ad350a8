"""The native table path (csrc/gcrf_tables.cpp): loading, annotating, sorting and
filtering the genes / features tables, packing them for the kernels and writing the result tables — against the
row-by-row Python restatement (oracle/tables_oracle.py) and the reference's own committed result tables."""
import gzip
import hashlib
import math
import pathlib
import random

import numpy
import pytest

from oracle import tables_oracle

REFERENCE = pathlib.Path("/root/reference/tests/test_cli/data")


def bgc_tables(bgc, *, with_probabilities=False, shuffle=None, crlf=False, reorder=False):
    """The fixture's two tables as text, from tests/golden/bgc0001866.json."""
    genes = {g["protein_id"]: g for g in bgc["genes"]}
    gcols = ["sequence_id", "protein_id", "start", "end", "strand"] + (["average_p", "max_p"] if with_probabilities else [])
    fcols = ["sequence_id", "protein_id", "start", "end", "strand", "domain", "hmm", "i_evalue", "pvalue", "domain_start",
             "domain_end"] + (["cluster_probability"] if with_probabilities else [])
    grows = [{c: g[c] for c in gcols} for g in bgc["genes"]]
    frows = []
    for d in bgc["domains"]:
        g = genes[d["protein_id"]]
        row = {c: g[c] for c in ("sequence_id", "protein_id", "start", "end", "strand")}
        row.update({c: d[c] for c in fcols if c in d})
        frows.append(row)
    if shuffle is not None:
        rng = random.Random(shuffle)
        rng.shuffle(grows)
        rng.shuffle(frows)
    if reorder:
        gcols = gcols[::-1]
        fcols = fcols[3:] + fcols[:3]
    nl = "\r\n" if crlf else "\n"

    def fmt(x):
        return repr(x) if isinstance(x, float) else str(x)

    gtext = nl.join(["\t".join(gcols)] + ["\t".join(fmt(r[c]) for c in gcols) for r in grows]) + nl
    ftext = nl.join(["\t".join(fcols)] + ["\t".join(fmt(r[c]) for c in fcols) for r in frows]) + nl
    return gtext, ftext


def native(gtext, ftexts, **kw):
    from gecco_b200.tables import FeatureTables

    return FeatureTables.parse(gtext.encode(), [f.encode() for f in ftexts], **kw)


def assert_accession_pack(tables, genes, weights, feature_type):
    """gcrf_table_pack_accessions + the device's part restated (vocabulary lookup, first of equal names per row) ==
    the row-by-row restatement's pack; written results must not depend on which pack ran last."""
    cp, rp, ai, rg = tables_oracle.pack(genes, weights.attr_index, feature_type)
    acc = tables.pack(None, feature_type, accessions=True)
    assert acc.accessions and acc.contig_ptr.tolist() == cp and tables.row_gene.tolist() == rg
    number = {int(a[2:]): i for i, a in enumerate(weights.attrs)}
    rows, ptr = [], [0]
    for r in range(acc.G):
        seen = []
        for a in acc.attr_idx[acc.gene_ptr[r]:acc.gene_ptr[r + 1]].tolist():
            if a in number and number[a] not in seen:
                seen.append(number[a])
        rows.extend(seen)
        ptr.append(len(rows))
    assert ptr == rp and rows == ai
    return acc


def assert_same_pack(tables, genes, weights, feature_type):
    packed = tables.pack(weights.attrs, feature_type)
    cp, rp, ai, rg = tables_oracle.pack(genes, weights.attr_index, feature_type)
    assert packed.contig_ptr.tolist() == cp
    assert packed.gene_ptr.tolist() == rp
    assert packed.attr_idx.tolist() == ai
    assert tables.row_gene.tolist() == rg
    return packed


def test_result_tables_are_byte_identical_to_the_reference(bgc, weights, tmp_path):
    """Golden probabilities in -> the reference's committed genes.tsv / features.tsv out (SHA-256 with "\\n" ends)."""
    gtext, ftext = bgc_tables(bgc)
    with native(gtext, [ftext]) as tables:
        assert (tables.contigs, tables.genes, tables.domains) == (1, 23, 37)
        assert tables.gene_ids == [g["protein_id"] for g in bgc["genes"]]
        packed = tables.pack(weights.attrs)
        assert packed.nnz == 35  # 37 rows, PF00550 three times in gene _23
        prob = numpy.array([g["average_p"] for g in bgc["genes"]])
        tables.write_genes(tmp_path / "genes.tsv", prob)
        tables.write_features(tmp_path / "features.tsv", prob)
        for name, key in (("genes.tsv", "genes_tsv"), ("features.tsv", "features_tsv")):
            data = (tmp_path / name).read_bytes()
            assert hashlib.sha256(data).hexdigest() == bgc["sha256_lf"][key], name
            if (REFERENCE / f"BGC0001866.{name}").exists():
                assert data == (REFERENCE / f"BGC0001866.{name}").read_bytes().replace(b"\r\n", b"\n")
        avg, mx = tables.gene_probabilities(prob)
        assert numpy.array_equal(avg, prob) and numpy.array_equal(mx, prob)


def test_clusters_table_matches_the_reference_row(bgc, weights, tmp_path):
    """Golden probabilities -> segments (CPU oracle of gcrf_segments here, the device in the gpu test) -> the
    reference's committed clusters.tsv row, in every column that does not come from the type classifier."""
    from gecco_b200._lib import Segments
    from oracle import refine_oracle

    gtext, ftext = bgc_tables(bgc, shuffle=9)
    with native(gtext, [ftext]) as tables:
        tables.pack(weights.attrs)
        prob = numpy.array([g["average_p"] for g in bgc["genes"]])
        rows = refine_oracle.extract_clusters(tables.contig_ptr, prob, tables.annotated, threshold=0.8, n_cds=3,
                                              edge_distance=0, trim=True, reset_per_contig=True)
        assert len(rows) == 1  # tests/test_cli/test_run.py:68-70
        seg = Segments(len(rows)).truncated(len(rows))
        for i, (c, b, e, o, avg, mx) in enumerate(rows):
            seg.contig[i], seg.begin[i], seg.end[i], seg.ordinal[i] = c, b, e, o
        tables.write_clusters(tmp_path / "clusters.tsv", prob, seg)
    got = tables_oracle.read_table((tmp_path / "clusters.tsv").read_text())
    assert len(got) == 1 and got[0]["type"] == "Unknown"
    assert list(got[0]) == ["sequence_id", "cluster_id", "start", "end", "average_p", "max_p", "proteins", "domains", "type"]
    want = bgc["clusters"][0]
    for key in ("sequence_id", "cluster_id", "average_p", "max_p"):
        assert got[0][key] == want[key], key  # strings: the floats are the reference's digits exactly
    # v0.11.0 writes sorted(protein ids) and sorted(domain names, repeats included) (gecco/model.py:751-758); the
    # committed fixture predates that: proteins in gene order, each domain name once
    assert got[0]["proteins"] == ";".join(sorted(want["proteins"].split(";")))
    assert got[0]["domains"] == ";".join(sorted(d["domain"] for d in bgc["domains"]))
    assert sorted(set(got[0]["domains"].split(";"))) == want["domains"].split(";")
    assert (int(got[0]["start"]), int(got[0]["end"])) == (want["start"], want["end"])


@pytest.mark.parametrize("feature_type", ["protein", "domain"])
@pytest.mark.parametrize("variant", [dict(), dict(shuffle=1, crlf=True), dict(shuffle=2, reorder=True, with_probabilities=True)])
def test_load_matches_the_row_by_row_restatement(bgc, weights, feature_type, variant, tmp_path):
    gtext, ftext = bgc_tables(bgc, **variant)
    for filters in (dict(), dict(p_filter=1e-20), dict(p_filter=None, e_filter=1e-15), dict(p_filter=1e-12, e_filter=1e-10)):
        genes = tables_oracle.load(gtext.replace("\r\n", "\n"), [ftext.replace("\r\n", "\n")], **filters)
        with native(gtext, [ftext], **filters) as tables:
            assert tables.gene_ids == [g["protein_id"] for g in genes]
            assert tables.domains == sum(len(g["domains"]) for g in genes)
            assert tables.annotated.tolist() == [int(bool(g["domains"])) for g in genes]
            packed = assert_same_pack(tables, genes, weights, feature_type)
            rng = numpy.random.default_rng(7)
            prob = rng.random(packed.G)
            prob[rng.random(packed.G) < 0.2] = numpy.nan
            tables.write_genes(tmp_path / "g.tsv", prob)
            tables.write_features(tmp_path / "f.tsv", prob)
            assert (tmp_path / "g.tsv").read_text() == tables_oracle.dump_genes(genes, prob.tolist(), feature_type)
            assert (tmp_path / "f.tsv").read_text() == tables_oracle.dump_features(genes, prob.tolist(), feature_type)
            # no probabilities at all: the columns are left out
            tables.write_genes(tmp_path / "g0.tsv")
            tables.write_features(tmp_path / "f0.tsv")
            assert (tmp_path / "g0.tsv").read_text() == tables_oracle.dump_genes(genes)
            assert (tmp_path / "f0.tsv").read_text() == tables_oracle.dump_features(genes)


def synthetic_tables(seed, contigs, genes_per_contig, names):
    rng = random.Random(seed)
    grows, frows = [], []
    for c in range(contigs):
        cid = f"ctg{rng.randrange(10**6):06d}.{c}"
        pos = 1
        for k in range(rng.randrange(1, genes_per_contig * 2)):
            start = pos + rng.randrange(0, 300)
            end = start + 3 * rng.randrange(30, 900)
            pos = start + rng.randrange(0, 50) if rng.random() < 0.1 else end  # a few overlapping / equal starts
            gid, strand = f"{cid}_{k + 1}", rng.choice("+-")
            grows.append((cid, gid, start, end, strand))
            for _ in range(rng.choice([0, 0, 1, 1, 2, 3, 8])):
                ds = rng.randrange(1, 400)
                pv = 10 ** rng.uniform(-30, -5)
                frows.append((cid, gid, start, end, strand, rng.choice(names), "Pfam", pv * 2766, pv, ds,
                              ds + rng.randrange(5, 200)))
    rng.shuffle(grows)
    rng.shuffle(frows)
    gtext = "sequence_id\tprotein_id\tstart\tend\tstrand\n" + "".join("\t".join(map(str, r)) + "\n" for r in grows)
    head = "sequence_id\tprotein_id\tstart\tend\tstrand\tdomain\thmm\ti_evalue\tpvalue\tdomain_start\tdomain_end\n"
    half = len(frows) // 2
    ftexts = [head + "".join("\t".join(repr(x) if isinstance(x, float) else str(x) for x in r) + "\n" for r in part)
              for part in (frows[:half], frows[half:])]
    return gtext, ftexts


def test_larger_shuffled_tables_in_two_feature_files(weights, tmp_path):
    """~6k genes / ~12k domain rows, unknown and repeated domain names, rows shuffled, two feature files, gzip."""
    names = list(weights.attrs[:300]) + ["PF99999", "TIGR00001"]
    gtext, ftexts = synthetic_tables(11, 60, 100, names)
    genes = tables_oracle.load(gtext, ftexts)
    with native(gtext, ftexts) as tables:
        assert tables.gene_ids == [g["protein_id"] for g in genes]
        assert tables.contig_ids == list(dict.fromkeys(g["sequence_id"] for g in genes))
        for feature_type in ("protein", "domain"):
            assert_same_pack(tables, genes, weights, feature_type)
            assert_accession_pack(tables, genes, weights, feature_type)
        start, end = tables.gene_coordinates()
        assert start.tolist() == [g["start"] for g in genes] and end.tolist() == [g["end"] for g in genes]
    # the same through files, one of them gzipped (gecco._meta.zopen sniffs the magic bytes)
    from gecco_b200.tables import FeatureTables

    (tmp_path / "x.genes.tsv").write_text(gtext)
    (tmp_path / "a.features.tsv").write_text(ftexts[0])
    with gzip.open(tmp_path / "b.features.tsv.gz", "wt") as f:
        f.write(ftexts[1])
    for paths in ([tmp_path / "a.features.tsv", tmp_path / "b.features.tsv.gz"],):
        with FeatureTables.load(tmp_path / "x.genes.tsv", paths) as tables:
            assert tables.gene_ids == [g["protein_id"] for g in genes]
            assert_same_pack(tables, genes, weights, "protein")
    with FeatureTables.load(tmp_path / "x.genes.tsv", tmp_path / "a.features.tsv", p_filter=None) as tables:
        assert tables.domains == ftexts[0].count("\n") - 1


def test_random_small_tables_property(weights, tmp_path):
    """Hypothesis: random tiny tables — equal starts, shared ids across files, NaN / empty numeric fields, unknown and
    repeated domain names, every filter combination — native result == row-by-row restatement, both feature types."""
    from hypothesis import HealthCheck, given, settings, strategies as st

    names = list(weights.attrs[:6]) + ["PF99999", "X"]
    gene = st.tuples(st.sampled_from(["b", "a", "a.1", "c10", "c9"]), st.integers(1, 60), st.integers(1, 40), st.sampled_from("+-"))
    number = st.one_of(st.just(""), st.just("nan"), st.floats(1e-30, 1e-3).map(repr), st.just("0.0"), st.just("1e-09"))
    domain = st.tuples(st.integers(0, 11), st.sampled_from(names), number, number, st.integers(1, 30), st.integers(1, 30))

    @settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.too_slow])
    @given(st.lists(gene, min_size=0, max_size=12), st.lists(domain, max_size=30), st.sampled_from([None, 1e-9, 1e-5]),
           st.sampled_from([None, 1e-6]), st.integers(0, 3))
    def check(genes, domains, p_filter, e_filter, split):
        rows = [(seq, f"{seq}_g{k}", start, start + 3 * length, strand) for k, (seq, start, length, strand) in enumerate(genes)]
        gtext = "sequence_id\tprotein_id\tstart\tend\tstrand\n" + "".join("\t".join(map(str, r)) + "\n" for r in rows)
        head = "sequence_id\tprotein_id\tstart\tend\tstrand\tdomain\thmm\ti_evalue\tpvalue\tdomain_start\tdomain_end\n"
        frows = []
        for g, name, ev, pv, ds, length in domains:
            if rows:
                r = rows[g % len(rows)]
                frows.append("\t".join(map(str, (*r, name, "Pfam", ev, pv, ds, ds + length))) + "\n")
        cut = min(split, len(frows))
        ftexts = [head + "".join(frows[:cut]), head + "".join(frows[cut:])]
        want = tables_oracle.load(gtext, ftexts, e_filter=e_filter, p_filter=p_filter)
        with native(gtext, ftexts, e_filter=e_filter, p_filter=p_filter) as tables:
            assert tables.gene_ids == [g["protein_id"] for g in want]
            assert tables.annotated.tolist() == [int(bool(g["domains"])) for g in want]
            for feature_type in ("protein", "domain"):
                packed = assert_same_pack(tables, want, weights, feature_type)
                if (len(genes) + len(domains)) % 2:  # the writers below then run behind the accession pack
                    assert assert_accession_pack(tables, want, weights, feature_type).G == packed.G
                prob = [0.25 + 0.5 * ((7 * k) % 11) / 11 for k in range(packed.G)]
                tables.write_genes(tmp_path / "g.tsv", numpy.array(prob))
                tables.write_features(tmp_path / "f.tsv", numpy.array(prob))
                assert (tmp_path / "g.tsv").read_text() == tables_oracle.dump_genes(want, prob, feature_type)
                assert (tmp_path / "f.tsv").read_text() == tables_oracle.dump_features(want, prob, feature_type)

    check()


def test_table_errors_are_the_reference_s_value_errors(bgc):
    gtext, ftext = bgc_tables(bgc)
    lines = gtext.splitlines()
    with pytest.raises(ValueError, match="Duplicate gene names"):
        native("\n".join(lines + [lines[1]]) + "\n", [ftext])
    with pytest.raises(ValueError, match="no column 'strand'"):
        native(gtext.replace("strand", "direction"), [ftext])
    flines = ftext.splitlines()
    bad = flines[1].split("\t")
    bad[2] = str(int(bad[2]) + 3)
    with pytest.raises(ValueError, match="Mismatched gene"):
        native(gtext, ["\n".join([flines[0], "\t".join(bad)]) + "\n"])
    bad = flines[1].split("\t")
    bad[1] = "nobody_1"
    with pytest.raises(ValueError, match="nobody_1"):
        native(gtext, ["\n".join([flines[0], "\t".join(bad)]) + "\n"])
    with pytest.raises(FileNotFoundError):
        from gecco_b200.tables import FeatureTables

        FeatureTables.load("/nonexistent/genes.tsv", [])
    # empty tables are fine
    with native("sequence_id\tprotein_id\tstart\tend\tstrand\n", []) as tables:
        assert (tables.contigs, tables.genes, tables.domains) == (0, 0, 0)
        assert tables.pack(["PF00001"]).G == 0


def test_float_layout_is_python_repr(tmp_path):
    """Shortest round-trip digits, repr() layout: 1e-05 not 1e-5, 1e+16, 123456789012345.0, 0.0001."""
    values = [0.0, 1.0, 0.5, 1e-4, 9.999e-5, 1e-5, 1e15, 1e16, 123456789012345.0, 1.7976931348623157e308, 5e-324,
              0.1 + 0.2, 2.262067179461254e-08, 1 / 3, 100.0, 1234.5, 0.9998703656415205]
    rng = random.Random(3)
    values += [rng.random() * 10 ** rng.randrange(-30, 30) for _ in range(2000)]
    gtext = "sequence_id\tprotein_id\tstart\tend\tstrand\n" + "".join(f"c\tg{i}\t{1 + 10 * i}\t{9 + 10 * i}\t+\n" for i in range(len(values)))
    with native(gtext, []) as tables:
        tables.pack([])
        tables.write_genes(tmp_path / "g.tsv", numpy.array(values))
        got = [line.split("\t")[5] for line in (tmp_path / "g.tsv").read_text().splitlines()[1:]]
    assert got == [repr(v) for v in values]


@pytest.mark.parametrize("feature_type", ["protein", "domain"])
def test_accession_pack_is_the_id_pack_before_feature_extraction(weights, feature_type):
    """gcrf_table_pack_accessions hands the device every kept domain row as a Pfam number; mapping those through the
    vocabulary and keeping the first of equal names per row (features.py:13-35, what gcrf::features_kernel does) must
    give the batch gcrf_table_pack builds on the host — same rows, same contigs, same row -> gene map."""
    names = list(weights.attrs[:40]) + ["PF99999", "TIGR00001", "PFAM1", "PF", "PF12x", "PF" + weights.attrs[0][2:].lstrip("0"),
                                        "PF0" + weights.attrs[1][2:]]  # the last two: right number, wrong NAME
    gtext, ftexts = synthetic_tables(23, 40, 30, names)
    with native(gtext, ftexts) as tables:
        ids = tables.pack(weights.attrs, feature_type)
        want = (ids.contig_ptr.copy(), ids.gene_ptr.copy(), ids.attr_idx.copy(), tables.row_gene.copy())
        acc = tables.pack(None, feature_type, accessions=True)
        assert acc.accessions and not ids.accessions
        assert numpy.array_equal(acc.contig_ptr, want[0])
        assert numpy.array_equal(tables.row_gene, want[3])
        assert acc.G == len(want[1]) - 1
        number = {int(a[2:]): i for i, a in enumerate(weights.attrs)}
        rows, ptr = [], [0]
        for r in range(acc.G):
            seen = []
            for a in acc.attr_idx[acc.gene_ptr[r]:acc.gene_ptr[r + 1]].tolist():
                if a in number and number[a] not in seen:
                    seen.append(number[a])
            rows.extend(seen)
            ptr.append(len(rows))
        assert ptr == want[1].tolist() and rows == want[2].tolist()
        assert (acc.attr_idx == -1).sum() > 0 and 99999 in acc.attr_idx  # foreign names -> -1, unknown Pfam numbers kept
    with native("sequence_id\tprotein_id\tstart\tend\tstrand\n", []) as tables:
        empty = tables.pack(None, feature_type, accessions=True)
        assert (empty.C, empty.G, empty.nnz) == (0, 0, 0)


def test_concurrent_name_index_finds_duplicates_and_every_gene(monkeypatch):
    """120k genes: the gene-name index is filled by several threads (one compare-and-swap per slot); a name that occurs
    twice, far apart, is still the reference's "Duplicate gene names" error (_common.py:217-219), and without it every
    feature row finds its gene."""
    monkeypatch.setenv("GCRF_TABLE_THREADS", "8")
    n = 120_000
    head = "sequence_id\tprotein_id\tstart\tend\tstrand\n"
    lines = [f"c{i // 50:05d}\tc{i // 50:05d}_{i % 50 + 1}\t{1 + 100 * (i % 50)}\t{90 + 100 * (i % 50)}\t+\n" for i in range(n)]
    fhead = "sequence_id\tprotein_id\tstart\tend\tstrand\tdomain\thmm\ti_evalue\tpvalue\tdomain_start\tdomain_end\n"
    frows = [lines[i][:-1] + "\tPF00005\tPfam\t1e-20\t1e-22\t1\t20\n" for i in range(0, n, 7)]
    with native(head + "".join(lines), [fhead + "".join(frows)]) as tables:
        assert (tables.genes, tables.domains) == (n, len(frows))
        assert int(tables.annotated.sum()) == len(frows)
    dup = list(lines)
    dup[n - 5] = dup[n - 5].replace(f"c{(n - 5) // 50:05d}_{(n - 5) % 50 + 1}\t", "c00003_4\t")
    with pytest.raises(ValueError, match="Duplicate gene names"):
        native(head + "".join(dup), [])


@pytest.mark.parametrize("threads", ["8", "3", "1"])
def test_many_contigs_are_ranked_by_id_on_all_threads(threads, monkeypatch):
    """170k contigs whose ids arrive in no particular order: the distinct ids are sorted in slices by all threads and
    merged pairwise; contigs come out in id order (predict.py:81), genes inside a contig by start."""
    monkeypatch.setenv("GCRF_TABLE_THREADS", threads)
    n = 170_000
    head = "sequence_id\tprotein_id\tstart\tend\tstrand\n"
    order = [(i * 7919) % n for i in range(n)]  # a permutation: 7919 is prime and does not divide n
    lines = [f"k{c:06d}\tk{c:06d}_1\t10\t400\t+\n" for c in order]
    lines += [f"k{c:06d}\tk{c:06d}_0\t5\t9\t-\n" for c in order[:1000]]  # a second, earlier gene for some
    with native(head + "".join(lines), []) as tables:
        assert (tables.contigs, tables.genes) == (n, n + 1000)
        ids = tables.contig_ids
        assert ids[0] == "k000000" and ids[n - 1] == f"k{n - 1:06d}" and list(ids) == sorted(ids)
        first = sorted(order[:1000])[0]
        cp = tables.contig_ptr
        assert cp[first + 1] - cp[first] == 2
        assert tables.gene_ids[int(cp[first])] == f"k{first:06d}_0" and tables.gene_ids[int(cp[first]) + 1] == f"k{first:06d}_1"


def test_writers_are_independent_of_threads_and_chunks(weights, tmp_path, monkeypatch):
    """~30k genes / ~60k domain rows: the writers cut the genes into chunks that threads format and place with pwrite at
    chained offsets; one thread or eight, the files are the same bytes, and every row is the restatement's row."""
    names = list(weights.attrs[:300]) + ["PF99999"]
    gtext, ftexts = synthetic_tables(5, 300, 100, names)
    genes = tables_oracle.load(gtext, ftexts)
    rng = numpy.random.default_rng(0)
    outs = {}
    for threads in ("1", "8", "3"):
        monkeypatch.setenv("GCRF_TABLE_THREADS", threads)
        with native(gtext, ftexts) as tables:
            packed = tables.pack(weights.attrs)
            prob = rng.random(packed.G) if threads == "1" else prob
            tables.write_genes(tmp_path / f"g{threads}.tsv", prob)
            tables.write_features(tmp_path / f"f{threads}.tsv", prob)
        outs[threads] = ((tmp_path / f"g{threads}.tsv").read_bytes(), (tmp_path / f"f{threads}.tsv").read_bytes())
    assert outs["1"] == outs["8"] == outs["3"]
    assert outs["1"][0].decode() == tables_oracle.dump_genes(genes, prob.tolist())
    assert outs["1"][1].decode() == tables_oracle.dump_features(genes, prob.tolist())


@pytest.mark.skipif(not (REFERENCE / "mibig-2.0.proG2.features.tsv").exists(), reason="reference checkout not mounted")
def test_reference_mibig_tables(weights, mibig):
    """The reference's own 15,158-gene / 123,031-row training tables: same CSR as the golden arrays."""
    from gecco_b200.packer import pack_arrays
    from gecco_b200.tables import FeatureTables

    with FeatureTables.load(REFERENCE / "mibig-2.0.proG2.genes.tsv", REFERENCE / "mibig-2.0.proG2.features.tsv") as tables:
        assert (tables.contigs, tables.genes, tables.domains) == (18, 15158, 25446)
        packed = tables.pack(weights.attrs)
        want = pack_arrays(mibig["gene_contig"], mibig["dom_ptr"], mibig["dom_pfam"], weights)
        assert numpy.array_equal(packed.contig_ptr, want.contig_ptr)
        assert numpy.array_equal(packed.gene_ptr, want.gene_ptr)
        assert numpy.array_equal(packed.attr_idx, want.attr_idx)


@pytest.mark.gpu
def test_predict_tables_on_device(bgc, tmp_path):
    """Tables in, tables out through the B200: the reference's golden numbers within the north-star tolerance."""
    from gecco_b200.crf import ClusterCRF
    from gecco_b200.tables import predict_tables

    gtext, ftext = bgc_tables(bgc, shuffle=5)
    (tmp_path / "BGC0001866.genes.tsv").write_text(gtext)
    (tmp_path / "BGC0001866.features.tsv").write_text(ftext)
    crf = ClusterCRF.trained()
    tables, prob = predict_tables(tmp_path / "BGC0001866.genes.tsv", tmp_path / "BGC0001866.features.tsv", tmp_path / "out", model=crf)
    golden = numpy.array([g["average_p"] for g in bgc["genes"]])
    assert numpy.abs(prob - golden).max() <= 1e-5
    rows = tables_oracle.read_table((tmp_path / "out" / "BGC0001866.genes.tsv").read_text())
    assert [r["protein_id"] for r in rows] == [g["protein_id"] for g in bgc["genes"]]
    assert max(abs(float(r["average_p"]) - g["average_p"]) for r, g in zip(rows, bgc["genes"])) <= 1e-5
    frows = tables_oracle.read_table((tmp_path / "out" / "BGC0001866.features.tsv").read_text())
    assert len(frows) == 37 and all(abs(float(r["cluster_probability"]) - d["cluster_probability"]) <= 1e-5
                                    for r, d in zip(frows, bgc["domains"]))
    # the reference finds exactly one cluster on this genome (tests/test_cli/test_run.py:68-70)
    clusters = tables.segments(crf, prob, threshold=0.8, n_cds=3)
    assert len(clusters) == 1 and clusters[0][0] == "BGC0001866.1"
    crows = tables_oracle.read_table((tmp_path / "out" / "BGC0001866.clusters.tsv").read_text())
    want = bgc["clusters"][0]
    assert len(crows) == 1 and all(crows[0][k] == want[k] for k in ("sequence_id", "cluster_id"))
    assert crows[0]["proteins"] == ";".join(sorted(want["proteins"].split(";")))
    assert sorted(set(crows[0]["domains"].split(";"))) == want["domains"].split(";")
    assert (int(crows[0]["start"]), int(crows[0]["end"])) == (want["start"], want["end"])
    assert abs(float(crows[0]["average_p"]) - float(want["average_p"])) <= 1e-5
    tables.close()


@pytest.mark.gpu
def test_predict_tables_reproduces_the_reference_tables_byte_for_byte(bgc, tmp_path):
    """Tables in -> CRF on the B200 in the reference's own arithmetic (GCRF_FLAG_F64, the drop-in's default) -> result
    tables whose SHA-256 equals that of the reference's committed BGC0001866.genes.tsv / .features.tsv (python-crfsuite
    output, 16-17 digits per probability), from the DEVICE'S OWN numbers; the clusters row's average_p / max_p strings
    equal the reference's as well."""
    from gecco_b200.crf import ClusterCRF
    from gecco_b200.tables import predict_tables

    gtext, ftext = bgc_tables(bgc, shuffle=3)
    (tmp_path / "BGC0001866.genes.tsv").write_text(gtext)
    (tmp_path / "BGC0001866.features.tsv").write_text(ftext)
    crf = ClusterCRF.trained()
    assert crf.arithmetic == "f64"
    tables, prob = predict_tables(tmp_path / "BGC0001866.genes.tsv", tmp_path / "BGC0001866.features.tsv", tmp_path / "out", model=crf)
    tables.close()
    assert prob.tolist() == [g["average_p"] for g in bgc["genes"]]
    for name, key in (("genes.tsv", "genes_tsv"), ("features.tsv", "features_tsv")):
        data = (tmp_path / "out" / f"BGC0001866.{name}").read_bytes()
        assert hashlib.sha256(data).hexdigest() == bgc["sha256_lf"][key], name
    crows = tables_oracle.read_table((tmp_path / "out" / "BGC0001866.clusters.tsv").read_text())
    assert len(crows) == 1 and all(crows[0][k] == bgc["clusters"][0][k] for k in ("average_p", "max_p", "cluster_id"))


def test_the_abi_from_plain_c(bgc, tmp_path):
    """examples/tables_roundtrip.c, compiled as C99 against include/gecco_crf_b200.h and linked with the shared library
    (no Python, no torch in that process): loads the reference fixture's tables, packs the accession batch and writes
    the genes table — the header is valid C and the table half of the ABI needs no GPU."""
    import shutil
    import subprocess

    root = pathlib.Path(__file__).resolve().parent.parent
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = tmp_path / "tables_roundtrip"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", f"-I{root / 'include'}", str(root / "examples" / "tables_roundtrip.c"),
                    f"-L{root / 'gecco_b200'}", "-lgecco_crf_b200", f"-Wl,-rpath,{root / 'gecco_b200'}", "-lm", "-o", str(exe)],
                   check=True, capture_output=True)
    gtext, ftext = bgc_tables(bgc)
    (tmp_path / "x.genes.tsv").write_text(gtext)
    (tmp_path / "x.features.tsv").write_text(ftext)
    run = subprocess.run([str(exe), str(tmp_path / "x.genes.tsv"), str(tmp_path / "x.features.tsv"), str(tmp_path / "o.genes.tsv")],
                         check=True, capture_output=True, text=True)
    fields = dict(kv.split("=") for kv in run.stdout.split())
    with native(gtext, [ftext]) as tables:
        assert (int(fields["contigs"]), int(fields["genes"]), int(fields["domains"])) == (tables.contigs, tables.genes, tables.domains)
        assert int(fields["rows"]) == tables.genes and int(fields["nnz"]) == tables.domains == int(fields["last_row_end"])
        assert fields["first_contig"] == tables.contig_ids[0]
        tables.write_genes(tmp_path / "p.genes.tsv")
    assert (tmp_path / "o.genes.tsv").read_bytes() == (tmp_path / "p.genes.tsv").read_bytes()


class StandInEngine:
    """What FeatureTables / predict_tables ask of CRFEngine, answered on the CPU: marginals from the oracle (after the
    device's feature extraction restated on arrays when the batch holds accessions), segments from the refine oracle."""

    has_vocabulary = True
    vocabulary_digits = 5

    def __init__(self, weights):
        self.weights = weights
        self.accession_batches = 0

    def marginals_windowed(self, contig_ptr, gene_ptr, attr_idx, *, window, step, pad, accessions=False, f64_arith=False):
        from oracle import crf_oracle

        gene_ptr, attr_idx = numpy.asarray(gene_ptr), numpy.asarray(attr_idx)
        if accessions:
            self.accession_batches += 1
            number = {int(a[2:]): i for i, a in enumerate(self.weights.attrs)}
            ids, ptr = [], [0]
            for r in range(len(gene_ptr) - 1):
                seen = []
                for a in attr_idx[gene_ptr[r]:gene_ptr[r + 1]].tolist():
                    if a in number and number[a] not in seen:
                        seen.append(number[a])
                ids.extend(seen)
                ptr.append(len(ids))
            gene_ptr, attr_idx = numpy.array(ptr, dtype=numpy.int32), numpy.array(ids, dtype=numpy.int32)
        p, _ = crf_oracle.marginals_windowed(self.weights.state_w, self.weights.trans_w, self.weights.label_id("1"),
                                             numpy.asarray(contig_ptr), gene_ptr, attr_idx, window, step, pad)
        return p

    def segments(self, contig_ptr, prob, annotated, *, threshold, n_cds, edge_distance, trim, reset_per_contig):
        from gecco_b200._lib import Segments
        from oracle import refine_oracle

        found = refine_oracle.extract_clusters(contig_ptr, prob, annotated, threshold, n_cds, edge_distance, trim, reset_per_contig)
        seg = Segments(len(found)).truncated(len(found))
        for k, (c, b, e, o, a, m) in enumerate(found):
            seg.contig[k], seg.begin[k], seg.end[k], seg.ordinal[k], seg.average_p[k], seg.max_p[k] = c, b, e, o, a, m
        return seg


@pytest.mark.parametrize("host_features", ["0", "1"])
def test_predict_tables_plumbing_on_the_cpu(bgc, weights, tmp_path, monkeypatch, host_features):
    """predict_tables end to end with the device replaced by the oracles: tables in, the reference's golden genes /
    features / clusters tables out — through the accession batch (feature extraction left to the engine) and through
    the host packer."""
    from gecco_b200.crf import ClusterCRF
    from gecco_b200.tables import predict_tables

    monkeypatch.setenv("GECCO_B200_HOST_FEATURES", host_features)
    gtext, ftext = bgc_tables(bgc, shuffle=5)
    (tmp_path / "BGC0001866.genes.tsv").write_text(gtext)
    (tmp_path / "BGC0001866.features.tsv").write_text(ftext)
    crf = ClusterCRF.trained()
    crf._engine = StandInEngine(weights)
    tables, prob = predict_tables(tmp_path / "BGC0001866.genes.tsv", tmp_path / "BGC0001866.features.tsv", tmp_path / "out", model=crf)
    assert crf._engine.accession_batches == (1 if host_features == "0" else 0)
    golden = numpy.array([g["average_p"] for g in bgc["genes"]])
    assert numpy.abs(prob - golden).max() <= 1e-12
    rows = tables_oracle.read_table((tmp_path / "out" / "BGC0001866.genes.tsv").read_text())
    assert [r["protein_id"] for r in rows] == [g["protein_id"] for g in bgc["genes"]]
    frows = tables_oracle.read_table((tmp_path / "out" / "BGC0001866.features.tsv").read_text())
    assert len(frows) == 37 and all(abs(float(r["cluster_probability"]) - d["cluster_probability"]) <= 1e-12
                                    for r, d in zip(frows, bgc["domains"]))
    clusters = tables.segments(crf, prob, threshold=0.8, n_cds=3)
    assert len(clusters) == 1 and clusters[0][0] == "BGC0001866.1"  # tests/test_cli/test_run.py:68-70
    crows = tables_oracle.read_table((tmp_path / "out" / "BGC0001866.clusters.tsv").read_text())
    want = bgc["clusters"][0]
    assert len(crows) == 1 and all(crows[0][k] == want[k] for k in ("sequence_id", "cluster_id"))
    assert (int(crows[0]["start"]), int(crows[0]["end"])) == (want["start"], want["end"])
    assert abs(float(crows[0]["average_p"]) - float(want["average_p"])) <= 1e-12
    tables.close()
