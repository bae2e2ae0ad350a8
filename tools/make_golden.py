#!/usr/bin/env python
"""Generate the committed fixtures from the read-only reference checkout (run in the build container).

    python tools/make_golden.py [/root/reference]

Outputs (all small, all derived — no reference source is copied):

* ``gecco_b200/data/gecco-0.11.0/model.{state,trans,meta}.tsv`` — the weight tables decoded from the
  CRFsuite blob inside ``gecco/crf/model.pkl`` (same three-column format ``gecco train`` writes,
  ``gecco/cli/commands/train.py:66-85``); floats are written with ``repr`` so they round-trip exactly.
* ``tests/golden/bgc0001866.json`` — the 23 genes / 37 domain rows of the reference's CLI fixture
  together with the golden ``average_p`` column that real python-crfsuite produced
  (``tests/test_cli/data/BGC0001866.{genes,features}.tsv``; the same files the Galaxy tool test diffs).
* ``tests/golden/mibig_proG2.npz`` — the 18-contig mibig fixture after the default p<1e-9 domain filter,
  as integer arrays (contig of each gene, gene start, Pfam accession number of each domain row in the
  reference's order), plus ``ref_loop_prob``: the output of the REFERENCE'S OWN
  ``gecco.crf.ClusterCRF.predict_probabilities`` imported from the checkout, with the absent
  third-party tagger replaced by a fake ``sklearn_crfsuite.CRF`` that answers
  ``predict_marginals_single`` from the oracle's chain primitive.  This pins the window / pad / max-pool
  loop of the oracle against the reference's real Python code.
* ``tests/golden/ref_loop_cases.json`` — the same harness on small edge cases (short contig with and
  without padding, several contigs, unknown domains, duplicated domains).
"""

from __future__ import annotations

import csv
import itertools
import json
import pathlib
import pickle
import sys
import types
import warnings

import numpy

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from gecco_b200 import model_io  # noqa: E402
from oracle import crf_oracle  # noqa: E402


def read_tsv(path):
    with open(path, newline="") as f:
        return list(csv.DictReader(f, dialect="excel-tab"))


# ------------------------------------------------------------------------------------------------
# reference import harness
# ------------------------------------------------------------------------------------------------


def import_reference_crf(ref: pathlib.Path):
    """Import gecco.crf from the checkout with stub Bio modules (Biopython is not installed)."""
    bio = types.ModuleType("Bio"); bio.__version__ = "1.83"; bio.__path__ = []
    seq = types.ModuleType("Bio.Seq"); seq.Seq = type("Seq", (str,), {})
    rec = types.ModuleType("Bio.SeqRecord")

    class SeqRecord:
        def __init__(self, seq=None, id="<unknown id>", name="", **kw):
            self.seq, self.id, self.name = seq, id, name

    rec.SeqRecord = SeqRecord
    feat = types.ModuleType("Bio.SeqFeature")
    for name in ("SeqFeature", "FeatureLocation", "CompoundLocation", "Reference"):
        setattr(feat, name, type(name, (), {"__init__": lambda self, *a, **k: None}))
    sys.modules.update({"Bio": bio, "Bio.Seq": seq, "Bio.SeqRecord": rec, "Bio.SeqFeature": feat})
    sys.path.insert(0, str(ref))
    import gecco.crf  # noqa: F401
    import gecco.model

    return gecco.crf, gecco.model, SeqRecord


class OracleBackedCRF:
    """Fake ``sklearn_crfsuite.CRF``: the two members the reference touches on the hot path."""

    def __init__(self, weights: model_io.CRFWeights):
        self.w = weights
        self.state_features_ = weights.state_features_
        self.calls = 0

    def predict_marginals_single(self, xseq):
        index = self.w.attr_index
        ptr = [0]; idx = []
        for item in xseq:
            for name, value in item.items():
                assert value is True
                a = index.get(name)
                if a is not None:  # unknown attributes are dropped by the tagger
                    idx.append(a)
            ptr.append(len(idx))
        m = crf_oracle.chain_marginals(self.w.state_w, self.w.trans_w, numpy.array(ptr), numpy.array(idx, dtype=numpy.int32))
        self.calls += 1
        return [{label: float(m[t, l]) for l, label in enumerate(self.w.labels)} for t in range(len(xseq))]


def reference_predict(crfmod, modelmod, SeqRecord, weights, contigs, pad=True, window=None, step=None):
    """contigs: list of (contig_id, [(gene_id, start, [(domain_name, domain_start), ...]), ...])."""
    crf = crfmod.ClusterCRF(weights.feature_type, window_size=window or weights.window_size,
                            window_step=step or weights.window_step)
    crf.model = OracleBackedCRF(weights)
    genes = []
    for cid, cgenes in contigs:
        src = SeqRecord(id=cid)
        for gid, start, doms in cgenes:
            domains = [modelmod.Domain(n, s, s + 10, "Pfam", 1e-20, 1e-20) for n, s in doms]
            genes.append(modelmod.Gene(src, start, start + 99, modelmod.Strand.Coding,
                                       modelmod.Protein(gid, None, domains)))
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        out = crf.predict_probabilities(genes, pad=pad)
    return out, [str(w.message) for w in caught]


# ------------------------------------------------------------------------------------------------


def make_bgc_golden(ref_dir: str = "/root/reference") -> None:
    """tests/golden/bgc0001866.json: the reference's CLI fixture (inputs + python-crfsuite golden probabilities) and the
    SHA-256 of its two result tables with "\n" line ends, for the byte-exactness test of the native table writers."""
    import hashlib

    data = pathlib.Path(ref_dir) / "tests" / "test_cli" / "data"
    golden = ROOT / "tests" / "golden"
    genes = read_tsv(data / "BGC0001866.genes.tsv")
    feats = read_tsv(data / "BGC0001866.features.tsv")

    def sha_lf(path):
        return hashlib.sha256(path.read_bytes().replace(b"\r\n", b"\n")).hexdigest()

    doc = {
        "source": "zellerlab/GECCO v0.11.0 tests/test_cli/data/BGC0001866.{genes,features}.tsv",
        "note": "average_p was produced by real python-crfsuite (reference golden); p-values kept for the default p<1e-9 filter",
        "sha256_lf": {"genes_tsv": sha_lf(data / "BGC0001866.genes.tsv"), "features_tsv": sha_lf(data / "BGC0001866.features.tsv")},
        "clusters": [
            {k: (int(c[k]) if k in ("start", "end") else c[k])
             for k in ("sequence_id", "cluster_id", "start", "end", "average_p", "max_p", "proteins", "domains")}
            for c in read_tsv(data / "BGC0001866.clusters.tsv")
        ],
        "genes": [
            {"sequence_id": g["sequence_id"], "protein_id": g["protein_id"], "start": int(g["start"]),
             "end": int(g["end"]), "strand": g["strand"], "average_p": float(g["average_p"]),
             "max_p": float(g["max_p"])}
            for g in genes
        ],
        "domains": [
            {"protein_id": f["protein_id"], "domain": f["domain"], "hmm": f["hmm"], "domain_start": int(f["domain_start"]),
             "domain_end": int(f["domain_end"]), "i_evalue": float(f["i_evalue"]), "pvalue": float(f["pvalue"]),
             "cluster_probability": float(f["cluster_probability"])}
            for f in feats
        ],
    }
    (golden / "bgc0001866.json").write_text(json.dumps(doc, indent=1) + "\n")
    print(f"bgc0001866: {len(doc['genes'])} genes, {len(doc['domains'])} domain rows")


def main(ref_dir: str = "/root/reference") -> None:
    ref = pathlib.Path(ref_dir)
    golden = ROOT / "tests" / "golden"
    golden.mkdir(parents=True, exist_ok=True)

    # (1) weight tables
    weights = model_io.load_pickled_model(ref / "gecco" / "crf")
    model_io.save_tsv_model(weights, model_io.bundled_model_dir())
    back = model_io.load_tsv_model(model_io.bundled_model_dir())
    assert back.attrs == weights.attrs and back.labels == weights.labels
    assert numpy.array_equal(back.state_w, weights.state_w) and numpy.array_equal(back.trans_w, weights.trans_w)
    assert numpy.array_equal(back.state_mask, weights.state_mask)
    print(f"model: {len(weights.attrs)} attrs, {int(weights.state_mask.sum())} state features")

    data = ref / "tests" / "test_cli" / "data"
    # (2) BGC0001866 golden
    make_bgc_golden(ref_dir)

    # (3) mibig through the reference's own loop
    crfmod, modelmod, SeqRecord = import_reference_crf(ref)
    mgenes = read_tsv(data / "mibig-2.0.proG2.genes.tsv")
    contig_names = []
    contig_of = {}
    gene_rows = {}
    for g in mgenes:
        if g["sequence_id"] not in contig_of:
            contig_of[g["sequence_id"]] = len(contig_names)
            contig_names.append(g["sequence_id"])
        gene_rows[g["protein_id"]] = (contig_of[g["sequence_id"]], int(g["start"]), [])
    kept = 0
    with open(data / "mibig-2.0.proG2.features.tsv", newline="") as f:
        for row in csv.DictReader(f, dialect="excel-tab"):
            if float(row["pvalue"]) < 1e-9:  # default filter, gecco/cli/commands/_parser.py:171-175
                gene_rows[row["protein_id"]][2].append((row["domain"], int(row["domain_start"])))
                kept += 1
    contigs = [(name, []) for name in contig_names]
    for gid, (c, start, doms) in gene_rows.items():
        contigs[c][1].append((gid, start, doms))
    out, warns = reference_predict(crfmod, modelmod, SeqRecord, weights, contigs)
    assert not warns, warns
    # the reference returns genes sorted by (contig id, start); store inputs in that order too
    order = {g.protein.id: k for k, g in enumerate(out)}
    gene_contig_name = [g.source.id for g in out]
    sorted_names = sorted(set(gene_contig_name))
    g_contig = numpy.array([sorted_names.index(n) for n in gene_contig_name], dtype=numpy.int32)
    g_start = numpy.array([g.start for g in out], dtype=numpy.int64)
    dom_ptr = [0]; dom_acc = []
    for g in out:
        for d in g.protein.domains:  # already sorted by domain start by the reference (:200-201)
            assert d.name.startswith("PF") and len(d.name) == 7
            dom_acc.append(int(d.name[2:]))
        dom_ptr.append(len(dom_acc))
    prob = numpy.array([g.average_probability for g in out], dtype=numpy.float64)
    numpy.savez_compressed(
        golden / "mibig_proG2.npz",
        contig_names=numpy.array(sorted_names),
        gene_contig=g_contig, gene_start=g_start,
        dom_ptr=numpy.array(dom_ptr, dtype=numpy.int64), dom_pfam=numpy.array(dom_acc, dtype=numpy.int32),
        ref_loop_prob=prob,
    )
    print(f"mibig: {len(out)} genes, {kept} domain rows kept, {len(sorted_names)} contigs; order check {len(order)}")

    # (4) small edge cases through the same harness
    names = weights.attrs
    rng = numpy.random.default_rng(11)

    def rand_contig(cid, n, dmean, unknown=False, dup=False):
        genes = []
        for k in range(n):
            nd = int(rng.poisson(dmean))
            doms = [(names[int(rng.integers(0, len(names)))], int(rng.integers(1, 500))) for _ in range(nd)]
            if unknown and nd:
                doms.append(("PF99999", 3))
            if dup and doms:
                doms.append((doms[0][0], doms[0][1] + 50))
            genes.append((f"{cid}_{k + 1}", 100 + 1000 * k, doms))
        return (cid, genes)

    cases = []
    specs = [
        ("short_pad", [rand_contig("ctgA", 7, 2.0)], True, None, None),
        ("short_nopad", [rand_contig("ctgA", 7, 2.0), rand_contig("ctgB", 25, 1.5)], False, None, None),
        ("single_gene", [rand_contig("ctgA", 1, 3.0)], True, None, None),
        ("exact_window", [rand_contig("ctgA", 20, 1.0)], True, None, None),
        ("multi_contig", [rand_contig("zz", 31, 1.5), rand_contig("aa", 19, 4.0), rand_contig("mm", 64, 0.7)], True, None, None),
        ("unknown_and_dups", [rand_contig("ctgA", 40, 3.0, unknown=True, dup=True)], True, None, None),
        ("window5_step2", [rand_contig("ctgA", 23, 2.0), rand_contig("ctgB", 3, 2.0)], True, 5, 2),
        ("window7_step7", [rand_contig("ctgA", 30, 2.0)], True, 7, 7),
        ("dense", [rand_contig("ctgA", 45, 25.0)], True, None, None),
    ]
    for name, ctgs, pad, window, step in specs:
        out, warns = reference_predict(crfmod, modelmod, SeqRecord, weights, ctgs, pad=pad, window=window, step=step)
        cases.append({
            "name": name, "pad": pad, "window": window or weights.window_size, "step": step or weights.window_step,
            "contigs": [{"id": cid, "genes": [{"id": gid, "start": s, "domains": [[n, ds] for n, ds in doms]}
                                              for gid, s, doms in cg]} for cid, cg in ctgs],
            "warnings": warns,
            "expected": [{"id": g.protein.id, "contig": g.source.id,
                          "p": g.average_probability,
                          "weights": [d.cluster_weight for d in g.protein.domains]} for g in out],
        })
    (golden / "ref_loop_cases.json").write_text(json.dumps({"cases": cases}, indent=0) + "\n")
    print(f"ref_loop_cases: {len(cases)} cases")


def make_refine_golden(ref_dir: str = "/root/reference") -> None:
    """Next row N1: the REFERENCE'S OWN ``gecco.refine.ClusterRefiner`` on random gene tables
    (probabilities incl. missing ones, annotated / unannotated genes) -> ``tests/golden/refine_cases.json``."""
    ref = pathlib.Path(ref_dir)
    crfmod, modelmod, SeqRecord = import_reference_crf(ref)
    import gecco.refine

    rng = numpy.random.default_rng(31)
    cases = []
    settings = [dict(threshold=0.8, n_cds=3, edge_distance=0, trim=True),   # the CLI defaults (--cds 3)
                dict(threshold=0.8, n_cds=5, edge_distance=0, trim=True),   # the class defaults
                dict(threshold=0.5, n_cds=1, edge_distance=0, trim=False),
                dict(threshold=0.8, n_cds=3, edge_distance=2, trim=True),
                dict(threshold=0.3, n_cds=2, edge_distance=5, trim=False),
                dict(threshold=0.6, n_cds=4, edge_distance=1, trim=True)]
    for k in range(24):
        kw = settings[k % len(settings)]
        contigs = []
        genes = []
        for c in range(int(rng.integers(1, 6))):
            cid = f"ctg{int(rng.integers(0, 1000)):03d}_{c}"
            src = SeqRecord(id=cid)
            n = int(rng.integers(1, 60))
            # runs of high / low probability so that clusters actually appear
            probs = []
            state = rng.random() < 0.4
            for _ in range(n):
                if rng.random() < 0.12:
                    state = not state
                probs.append(float(numpy.clip(rng.normal(0.93 if state else 0.2, 0.12), 0, 1)))
            none_mode = rng.random()
            cg = []
            for i in range(n):
                p = probs[i]
                if none_mode < 0.25 and rng.random() < 0.15:
                    p = None
                if none_mode > 0.9:
                    p = None  # a contig skipped by predict_probabilities(pad=False)
                annotated = bool(rng.random() < 0.6)
                doms = [modelmod.Domain("PF00001", 1, 10, "Pfam", 1e-20, 1e-20)] if annotated else []
                gene = modelmod.Gene(src, 100 + 1000 * i, 900 + 1000 * i, modelmod.Strand.Coding,
                                     modelmod.Protein(f"{cid}_{i + 1}", None, doms), _probability=p)
                genes.append(gene)
                cg.append({"id": gene.id, "p": p, "annotated": annotated})
            contigs.append({"id": cid, "genes": cg})
        shuffled = list(genes)
        rng.shuffle(shuffled)
        refiner = gecco.refine.ClusterRefiner(criterion="gecco", **kw)
        clusters = list(refiner.iter_clusters(shuffled))

        def dump(cl):
            return {"id": cl.id, "genes": [g.id for g in cl.genes], "average_p": cl.average_probability,
                    "max_p": cl.maximum_probability}

        # the pipeline calls iter_clusters once per contig (gecco/cli/commands/_common.py:616-618): a fresh
        # GeneGrouper each time, so the in-cluster state does not leak from one contig into the next
        ordered = sorted(genes, key=lambda g: (g.source.id, g.start))
        per_contig = []
        for _, group in itertools.groupby(ordered, key=lambda g: g.source.id):
            per_contig.extend(refiner.iter_clusters(list(group)))
        cases.append({"settings": kw, "contigs": contigs, "clusters": [dump(cl) for cl in clusters],
                      "clusters_per_contig_call": [dump(cl) for cl in per_contig]})
    out = ROOT / "tests" / "golden" / "refine_cases.json"
    out.write_text(json.dumps({"cases": cases}, indent=0) + "\n")
    print(f"refine_cases: {len(cases)} cases, {sum(len(c['clusters']) for c in cases)} clusters")


def make_antismash_golden(ref_dir: str = "/root/reference") -> None:
    """``criterion="antismash"`` of the reference's ``ClusterRefiner`` (``gecco/refine.py:157-163``): the list of Pfam
    accessions antiSMASH counts as biosynthetic (data of the reference, ``:20-58``) -> ``gecco_b200/data/bio_pfams.txt``,
    and the class itself on random tables with named domains -> ``tests/golden/refine_antismash_cases.json``."""
    ref = pathlib.Path(ref_dir)
    crfmod, modelmod, SeqRecord = import_reference_crf(ref)
    import gecco.refine

    bio = sorted(gecco.refine.BIO_PFAMS)
    (ROOT / "gecco_b200" / "data" / "bio_pfams.txt").write_text("\n".join(bio) + "\n")
    rng = numpy.random.default_rng(47)
    other = [f"PF{n:05d}" for n in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12)]
    assert not set(other) & set(bio)
    settings = [dict(threshold=0.8, n_cds=5, n_biopfams=5, average_threshold=0.6, trim=True),   # the class defaults
                dict(threshold=0.8, n_cds=3, n_biopfams=2, average_threshold=0.9, trim=True),
                dict(threshold=0.5, n_cds=1, n_biopfams=1, average_threshold=0.6, trim=False),
                dict(threshold=0.6, n_cds=4, n_biopfams=3, average_threshold=0.85, trim=True),
                dict(threshold=0.3, n_cds=2, n_biopfams=0, average_threshold=0.0, trim=False)]
    cases = []
    for k in range(20):
        kw = settings[k % len(settings)]
        contigs, genes = [], []
        for c in range(int(rng.integers(1, 5))):
            cid = f"ctg{int(rng.integers(0, 1000)):03d}_{c}"
            src = SeqRecord(id=cid)
            n = int(rng.integers(1, 70))
            state = rng.random() < 0.4
            cg = []
            for i in range(n):
                if rng.random() < 0.1:
                    state = not state
                p = float(numpy.clip(rng.normal(0.9 if state else 0.2, 0.12), 0, 1))
                names = []
                for _ in range(int(rng.poisson(1.2))):
                    pool = bio[:12] if rng.random() < 0.5 else other
                    names.append(pool[int(rng.integers(0, len(pool)))])
                doms = [modelmod.Domain(a, 1 + 10 * j, 9 + 10 * j, "Pfam", 1e-20, 1e-20) for j, a in enumerate(names)]
                gene = modelmod.Gene(src, 100 + 1000 * i, 900 + 1000 * i, modelmod.Strand.Coding,
                                     modelmod.Protein(f"{cid}_{i + 1}", None, doms), _probability=p)
                genes.append(gene)
                cg.append({"id": gene.id, "p": p, "domains": names})
            contigs.append({"id": cid, "genes": cg})
        shuffled = list(genes)
        rng.shuffle(shuffled)
        refiner = gecco.refine.ClusterRefiner(criterion="antismash", **kw)
        clusters = [{"id": cl.id, "genes": [g.id for g in cl.genes]} for cl in refiner.iter_clusters(shuffled)]
        cases.append({"settings": kw, "contigs": contigs, "clusters": clusters})
    out = ROOT / "tests" / "golden" / "refine_antismash_cases.json"
    out.write_text(json.dumps({"cases": cases}, indent=0) + "\n")
    print(f"refine_antismash_cases: {len(cases)} cases, {sum(len(c['clusters']) for c in cases)} clusters, {len(bio)} biosynthetic Pfams")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "refine":
        make_refine_golden(*sys.argv[2:])
    elif len(sys.argv) > 1 and sys.argv[1] == "antismash":
        make_antismash_golden(*sys.argv[2:])
    else:
        main(*sys.argv[1:])
        make_refine_golden(*sys.argv[1:])
        make_antismash_golden(*sys.argv[1:])
