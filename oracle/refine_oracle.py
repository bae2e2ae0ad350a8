"""Array-form restatement of ``gecco.refine.ClusterRefiner`` (criterion "gecco") — TEST INFRASTRUCTURE.

Follows ``gecco/refine.py``:

* ``GeneGrouper`` (:51-64): a gene is in a cluster when its probability exceeds the threshold; a gene WITHOUT a
  probability keeps the state of the gene before it — and the grouper object lives across contigs (:190), so the
  state leaks from the last gene of one contig into the next one;
* ``_iter_clusters`` (:183-200): contigs in sorted order, maximal runs of in-cluster genes inside a contig;
* ``_trim_cluster`` (:167-180): genes without domain annotation are dropped from both ends;
* ``_validate_cluster`` (:139-165): at least ``n_cds`` annotated genes, and at least ``n_cds`` genes (annotated
  or not) outside the contig's first / last ``edge_distance`` annotated genes.

``reset_per_contig`` restates the pipeline's use of the class — one ``iter_clusters`` call, hence a fresh grouper,
per contig (``gecco/cli/commands/_common.py:616-618``).  ``extract_clusters`` also returns the per-contig ordinal of
the raw run (cluster ids are numbered before validation, :199-200) and the cluster's mean / max probability
(``gecco/model.py:443-454``).

Pinned on ``tests/golden/refine_cases.json`` (outputs of the reference class itself, ``tools/make_golden.py refine``).
"""

from typing import List, Tuple

import numpy


def extract_segments(contig_ptr, prob, annotated, threshold=0.8, n_cds=5, edge_distance=0, trim=True,
                     reset_per_contig=False) -> List[Tuple[int, int, int]]:
    """Returns ``[(contig, first_gene, last_gene + 1), ...]`` (global gene indices) in the reference's order."""
    return [seg[:3] for seg in extract_clusters(contig_ptr, prob, annotated, threshold, n_cds, edge_distance, trim,
                                                reset_per_contig)]


def extract_clusters(contig_ptr, prob, annotated, threshold=0.8, n_cds=5, edge_distance=0, trim=True,
                     reset_per_contig=False) -> List[Tuple[int, int, int, int, float, float]]:
    """Returns ``[(contig, first_gene, last_gene + 1, ordinal, average_p, max_p), ...]``."""
    contig_ptr = numpy.asarray(contig_ptr, dtype=numpy.int64)
    prob = numpy.asarray(prob, dtype=numpy.float64)
    annotated = numpy.asarray(annotated, dtype=bool)
    out = []
    state = False
    for c in range(len(contig_ptr) - 1):
        g0, g1 = int(contig_ptr[c]), int(contig_ptr[c + 1])
        if reset_per_contig:
            state = False
        ordinal = 0
        flags = []
        for g in range(g0, g1):
            if not numpy.isnan(prob[g]):
                state = bool(prob[g] > threshold)
            flags.append(state)
        ann_ids = [g for g in range(g0, g1) if annotated[g]]
        edge = set(ann_ids[:edge_distance]) | set(ann_ids[-edge_distance:]) if edge_distance > 0 else set()
        g = g0
        while g < g1:
            if not flags[g - g0]:
                g += 1
                continue
            e = g
            while e < g1 and flags[e - g0]:
                e += 1
            b, t = g, e  # run [g, e)
            ordinal += 1
            if trim:
                while b < t and not annotated[b]:
                    b += 1
                while t > b and not annotated[t - 1]:
                    t -= 1
            n_annot = int(annotated[b:t].sum())
            n_inner = sum(1 for x in range(b, t) if x not in edge)
            if n_annot >= n_cds and n_inner >= n_cds:
                have = prob[b:t][~numpy.isnan(prob[b:t])]
                out.append((c, b, t, ordinal, float(have.mean()) if len(have) else float("nan"),
                            float(have.max()) if len(have) else float("nan")))
            g = e
    return out
