"""Phase times of the native table path (gcrf_tables.cpp) on a synthetic genes / features table pair: load + annotate +
sort + filter, pack to CSR, write the genes and features tables.  CPU only (no kernel call); GCRF_TABLE_TIMING=1 adds the
reader's own phase times on stderr.

    python tools/tables_time.py [genes] [rows_per_gene] [contigs]
"""
import os
import pathlib
import sys
import tempfile
import time

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy
from gecco_b200 import model_io
from gecco_b200.tables import FeatureTables


def write_tables(tmp: pathlib.Path, genes: int, rows_per_gene: float, contigs: int, attrs, seed: int = 7):
    rng = numpy.random.default_rng(seed)
    contig_of = numpy.sort(rng.integers(0, contigs, size=genes))
    k = rng.poisson(rows_per_gene, size=genes)
    start = numpy.zeros(genes, dtype=numpy.int64)
    gl, fl = ["sequence_id\tprotein_id\tstart\tend\tstrand\n"], [
        "sequence_id\tprotein_id\tstart\tend\tstrand\tdomain\thmm\ti_evalue\tpvalue\tdomain_start\tdomain_end\n"]
    names = numpy.array(attrs)
    ordinal = 0
    last = -1
    pos = 0
    doms = rng.integers(0, len(attrs), size=int(k.sum()))
    pv = (10.0 ** rng.uniform(-30, -3, size=len(doms))).tolist()
    ds = rng.integers(1, 300, size=len(doms))
    d = 0
    for g in range(genes):
        c = int(contig_of[g])
        if c != last:
            ordinal, pos, last = 0, 0, c
        ordinal += 1
        pos += int(rng.integers(50, 400))
        s, e = pos, pos + int(rng.integers(200, 3000))
        pos = e
        strand = "+" if (g & 1) else "-"
        head = f"contig_{c:07d}\tcontig_{c:07d}_{ordinal}\t{s}\t{e}\t{strand}"
        gl.append(head + "\n")
        for _ in range(int(k[g])):
            fl.append(f"{head}\t{names[doms[d]]}\tPfam\t{pv[d] * 2766!r}\t{pv[d]!r}\t{ds[d]}\t{ds[d] + 80}\n")
            d += 1
    (tmp / "genes.tsv").write_text("".join(gl))
    (tmp / "features.tsv").write_text("".join(fl))
    return tmp / "genes.tsv", tmp / "features.tsv"


def main():
    genes = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    rows = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
    contigs = int(sys.argv[3]) if len(sys.argv) > 3 else max(1, genes // 40)
    w = model_io.load_tsv_model(model_io.bundled_model_dir())
    tmp = pathlib.Path(os.environ.get("GCRF_TABLES_TMP", tempfile.mkdtemp(prefix="gcrf_tables_")))
    tmp.mkdir(parents=True, exist_ok=True)
    gpath, fpath = tmp / "genes.tsv", tmp / "features.tsv"
    if not (gpath.exists() and fpath.exists()):
        t0 = time.perf_counter()
        write_tables(tmp, genes, rows, contigs, w.attrs)
        print(f"generated in {time.perf_counter() - t0:.1f} s: {gpath.stat().st_size / 1e6:.0f} MB genes, "
              f"{fpath.stat().st_size / 1e6:.0f} MB features", flush=True)
    for rep in range(2):
        t0 = time.perf_counter()
        t = FeatureTables.load(gpath, fpath)
        t1 = time.perf_counter()
        ta = time.perf_counter()
        t.pack(None, accessions=True)
        tb = time.perf_counter()
        packed = t.pack(w.attrs)
        t2 = time.perf_counter() - (tb - ta)
        prob = numpy.random.default_rng(1).random(packed.G)
        t.write_genes(tmp / "out.genes.tsv", prob)
        t3 = time.perf_counter() - (tb - ta)
        t.write_features(tmp / "out.features.tsv", prob)
        t4 = time.perf_counter() - (tb - ta)
        print(f"pass {rep}: {t.genes} genes / {t.domains} rows: load {t1 - t0:.3f} s, pack ids {t2 - t1:.3f} s / accessions {tb - ta:.3f} s, "
              f"write genes {t3 - t2:.3f} s, write features {t4 - t3:.3f} s", flush=True)
        t.close()


if __name__ == "__main__":
    main()
