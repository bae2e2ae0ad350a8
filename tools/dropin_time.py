"""Throughput of the three entry levels on the real-data shape (the 15,158-gene mibig fixture from tests/golden):
the List[Gene] drop-in (gecco_b200.crf.ClusterCRF.predict_probabilities, Python-object bound), the table path
(gecco_b200.tables, native load -> pack -> kernel -> write) and the array call.  B200 only."""
import pathlib
import sys
import tempfile
import time

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy
from fake_model import Domain, Gene, Protein, Source
from gecco_b200 import ClusterCRF
from gecco_b200.packer import pack_arrays
from gecco_b200.tables import FeatureTables

m = numpy.load(ROOT / "tests" / "golden" / "mibig_proG2.npz")
crf = ClusterCRF.trained()
w = crf._weights
gene_contig, dom_ptr, dom_pfam = m["gene_contig"], m["dom_ptr"], m["dom_pfam"]
G = len(gene_contig)


def make_genes():
    sources = {}
    genes = []
    for g in range(G):
        src = sources.setdefault(int(gene_contig[g]), Source(f"contig{int(gene_contig[g]):03d}"))
        doms = [Domain(f"PF{int(a):05d}", 10 * k + 1, 10 * k + 9) for k, a in enumerate(dom_pfam[dom_ptr[g]:dom_ptr[g + 1]])]
        genes.append(Gene(src, 1000 * g + 1, 1000 * g + 900, 1, Protein(f"g{g}", None, doms)))
    return genes


def best(fn, n=5):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


genes = make_genes()
crf.predict_probabilities(genes)  # warm-up: engine creation
t_obj = best(lambda: crf.predict_probabilities(genes), 3)
packed = pack_arrays(gene_contig, dom_ptr, dom_pfam, w)
t_arr = best(lambda: crf.marginals(packed))
with tempfile.TemporaryDirectory() as tmp:
    tmp = pathlib.Path(tmp)
    gt = "sequence_id\tprotein_id\tstart\tend\tstrand\n" + "".join(
        f"{g.source.id}\t{g.protein.id}\t{g.start}\t{g.end}\t+\n" for g in genes)
    ft = "sequence_id\tprotein_id\tstart\tend\tstrand\tdomain\thmm\ti_evalue\tpvalue\tdomain_start\tdomain_end\n" + "".join(
        f"{g.source.id}\t{g.protein.id}\t{g.start}\t{g.end}\t+\t{d.name}\tPfam\t1e-20\t1e-24\t{d.start}\t{d.end}\n"
        for g in genes for d in g.protein.domains)
    (tmp / "x.genes.tsv").write_text(gt)
    (tmp / "x.features.tsv").write_text(ft)

    def table_path():
        with FeatureTables.load(tmp / "x.genes.tsv", tmp / "x.features.tsv") as t:
            p = t.predict(crf)
            t.write_genes(tmp / "o.genes.tsv", p)
            t.write_features(tmp / "o.features.tsv", p)
            return p

    p_tab = table_path()
    t_tab = best(table_path)
p_obj = numpy.array([g.average_probability for g in crf.predict_probabilities(genes)])
assert numpy.array_equal(p_obj, p_tab) and numpy.array_equal(p_obj, crf.marginals(packed))
print(f"{G} genes, {len(dom_pfam)} domain rows (mibig fixture)")
print(f"List[Gene] drop-in (pack + kernel + with_probability/with_cluster_weight): {t_obj*1e3:8.1f} ms  {G/t_obj/1e3:9.0f} k genes/s")
print(f"tables: load + annotate + sort + filter + pack + kernel + write 2 tables:  {t_tab*1e3:8.1f} ms  {G/t_tab/1e3:9.0f} k genes/s")
print(f"array call (host buffers, one launch):                                     {t_arr*1e3:8.3f} ms  {G/t_arr/1e3:9.0f} k genes/s")
