"""Drop-in replacement for ``gecco.refine.ClusterRefiner`` whose segmentation runs on a B200 (next row N1).

The reference walks every gene in Python (``gecco/refine.py:183-200``: sort, ``itertools.groupby`` with a stateful
``GeneGrouper``), trims un-annotated genes from the cluster edges (``:167-180``) and validates (``:139-165``).  Here
the genes are reduced to three arrays (contig pointers, probabilities, "has a domain" marks), one call into
``libgecco_crf_b200.so`` (``gcrf_segments``) returns the valid clusters in the reference's order, and only those are
turned back into ``Cluster`` objects.  ``Gene`` / ``Cluster`` are duck-typed like in :mod:`gecco_b200.crf`.

Only ``criterion="gecco"`` — the default of the class and of ``gecco run`` — runs on the device; the
``"antismash"`` criterion needs the domain *names* of every cluster (``:157-163``) and is delegated to the reference
class when GECCO is importable.
"""

from __future__ import annotations

import operator
from typing import Any, Iterator, List, Optional

import numpy

__all__ = ["ClusterRefiner"]


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
        if self.criterion == "antismash":
            try:
                import gecco.refine  # type: ignore
            except ImportError as err:
                raise NotImplementedError("criterion 'antismash' is delegated to gecco.refine, which is not installed") from err
            yield from gecco.refine.ClusterRefiner(
                threshold=self.threshold, criterion=self.criterion, n_cds=self.n_cds, n_biopfams=self.n_biopfams,
                average_threshold=self.average_threshold, edge_distance=self.edge_distance, trim=self.trim,
            ).iter_clusters(genes)
            return
        if self.criterion != "gecco":
            raise ValueError(f"Unknown cluster filtering criterion: {self.criterion}")  # :164-165

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

        seg = self._get_engine().segments(numpy.asarray(contig_ptr, dtype=numpy.int32), prob, annotated,
                                          threshold=self.threshold, n_cds=self.n_cds, edge_distance=self.edge_distance,
                                          trim=self.trim)
        Cluster = _cluster_type()
        for c, b, e, k in zip(seg.contig.tolist(), seg.begin.tolist(), seg.end.tolist(), seg.ordinal.tolist()):
            seq_id = table[contig_ptr[c]].source.id
            yield Cluster(f"{seq_id}_cluster_{k}", table[b:e])  # :199-200 names clusters before validation
