"""Drop-in replacement for ``gecco.refine.ClusterRefiner`` whose segmentation runs on a B200 (next row N1).

The reference walks every gene in Python (``gecco/refine.py:183-200``: sort, ``itertools.groupby`` with a stateful
``GeneGrouper``), trims un-annotated genes from the cluster edges (``:167-180``) and validates (``:139-165``).  Here
the genes are reduced to three arrays (contig pointers, probabilities, "has a domain" marks), one call into
``libgecco_crf_b200.so`` (``gcrf_segments``) returns the valid clusters in the reference's order, and only those are
turned back into ``Cluster`` objects.  ``Gene`` / ``Cluster`` are duck-typed like in :mod:`gecco_b200.crf`.

Both criteria take their segments from the device.  ``criterion="gecco"`` — the default of the class and of
``gecco run`` — is validated there too (annotated genes, distance to the contig edge).  ``criterion="antismash"``
(``:157-163``) needs the domain *names* of a cluster: the device returns every (trimmed) segment, and the three tests
— mean probability, distinct biosynthetic Pfams, number of genes — run here over the few segments that exist.
"""

from __future__ import annotations

import operator
from typing import Any, Iterator, List, Optional

import numpy

__all__ = ["ClusterRefiner", "BIO_PFAMS"]


def _load_bio_pfams() -> frozenset:
    """The Pfam accessions antiSMASH counts as biosynthetic, as GECCO ships them (``gecco/refine.py:20-58``; data,
    extracted by ``tools/make_golden.py antismash``)."""
    import pathlib

    path = pathlib.Path(__file__).resolve().parent / "data" / "bio_pfams.txt"
    return frozenset(line.strip() for line in path.read_text().splitlines() if line.strip())


BIO_PFAMS = _load_bio_pfams()


class _Cluster:
    """Stand-in for ``gecco.model.Cluster`` (``gecco/model.py:390-454``) when GECCO itself is not importable."""

    def __init__(self, id: str, genes: Optional[List[Any]] = None, type: Any = None, type_probabilities: Any = None):
        self.id = id
        self.genes = genes or []
        self.type = type
        self.type_probabilities = type_probabilities or {}


def _cluster_type():
    try:
        from gecco.model import Cluster  # type: ignore

        return Cluster
    except Exception:
        return _Cluster


class ClusterRefiner:
    """A post-processor to extract contiguous clusters from CRF predictions (``gecco/refine.py:67-118``)."""

    def __init__(self, *, threshold: float = 0.8, criterion: str = "gecco", n_cds: int = 5, n_biopfams: int = 5,
                 average_threshold: float = 0.6, edge_distance: int = 0, trim: bool = True, engine: Any = None,
                 device: int = 0) -> None:
        self.threshold = threshold
        self.criterion = criterion
        self.n_cds = n_cds
        self.n_biopfams = n_biopfams
        self.average_threshold = average_threshold
        self.edge_distance = edge_distance
        self.trim = trim
        self._engine = engine
        self._device = device

    def _get_engine(self):
        if self._engine is None:
            from ._lib import CRFEngine  # raises if the CUDA library is missing: there is no CPU fallback
            from .model_io import CRFWeights

            # gcrf_segments only needs the handle's device, stream and scratch buffers: a one-attribute null model
            blank = CRFWeights(attrs=["-"], labels=["0", "1"], state_w=numpy.zeros((1, 2)),
                               state_mask=numpy.zeros((1, 2), dtype=bool), trans_w=numpy.zeros((2, 2)))
            self._engine = CRFEngine(blank, device=self._device)
        return self._engine

    def iter_clusters(self, genes: List[Any]) -> Iterator[Any]:
        """Find all clusters in a table of CRF predictions (``gecco/refine.py:120-137``)."""
        if self.criterion not in ("gecco", "antismash"):
            raise ValueError(f"Unknown cluster filtering criterion: {self.criterion}")  # :164-165
        antismash = self.criterion == "antismash"

        # :193-195 — stable sort by contig id, then by coordinates inside each contig
        ordered = sorted(genes, key=operator.attrgetter("source.id"))
        contig_ptr = [0]
        table: List[Any] = []
        i = 0
        while i < len(ordered):
            j = i
            sid = ordered[i].source.id
            while j < len(ordered) and ordered[j].source.id == sid:
                j += 1
            table.extend(sorted(ordered[i:j], key=operator.attrgetter("start", "end")))
            contig_ptr.append(j)
            i = j
        if not table:
            return
        prob = numpy.array([numpy.nan if (p := g.average_probability) is None else p for g in table], dtype=numpy.float64)
        annotated = numpy.array([1 if g.protein.domains else 0 for g in table], dtype=numpy.uint8)

        # "antismash": every segment comes back (no annotated-gene or edge test on the device), validated below
        seg = self._get_engine().segments(numpy.asarray(contig_ptr, dtype=numpy.int32), prob, annotated,
                                          threshold=self.threshold, n_cds=0 if antismash else self.n_cds,
                                          edge_distance=0 if antismash else self.edge_distance, trim=self.trim)
        Cluster = _cluster_type()
        for c, b, e, k in zip(seg.contig.tolist(), seg.begin.tolist(), seg.end.tolist(), seg.ordinal.tolist()):
            members = table[b:e]
            if antismash:
                # :157-163 — a segment trimmed down to nothing has no mean probability and fails the first test
                if not members:
                    continue
                domains = {d.name for gene in members for d in gene.protein.domains}
                p_crit = numpy.mean([gene.average_probability for gene in members]) >= self.average_threshold
                bio_crit = len(domains & BIO_PFAMS) >= self.n_biopfams
                cds_crit = len(members) >= self.n_cds
                if not (p_crit and bio_crit and cds_crit):
                    continue
            seq_id = table[contig_ptr[c]].source.id
            yield Cluster(f"{seq_id}_cluster_{k}", members)  # :199-200 names clusters before validation
