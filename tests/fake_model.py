"""Minimal stand-ins for ``gecco.model.{Domain,Protein,Gene}`` (which need Biopython to import).

Only the members the CRF path touches are provided, with the same copy-on-write behaviour
(``gecco/model.py:110-375``): frozen dataclasses whose ``with_*`` methods return new objects.
"""
from dataclasses import dataclass, field
from typing import Any, Iterable, List, Optional


@dataclass(frozen=True)
class Source:
    id: str


@dataclass(frozen=True)
class Domain:
    name: str
    start: int
    end: int
    hmm: str = "Pfam"
    i_evalue: float = 0.0
    pvalue: float = 0.0
    probability: Optional[float] = None
    cluster_weight: Optional[float] = None

    def with_probability(self, probability):
        return Domain(self.name, self.start, self.end, self.hmm, self.i_evalue, self.pvalue, probability, self.cluster_weight)

    def with_cluster_weight(self, cluster_weight):
        return Domain(self.name, self.start, self.end, self.hmm, self.i_evalue, self.pvalue, self.probability, cluster_weight)


@dataclass(frozen=True)
class Protein:
    id: str
    seq: Any = None
    domains: List[Domain] = field(default_factory=list)

    def with_domains(self, domains: Iterable[Domain]):
        return Protein(self.id, self.seq, list(domains))


@dataclass(frozen=True)
class Gene:
    source: Source
    start: int
    end: int
    strand: int
    protein: Protein
    _probability: Optional[float] = None

    @property
    def id(self):
        return self.protein.id

    @property
    def average_probability(self):
        if self._probability is not None:
            return self._probability
        p = [d.probability for d in self.protein.domains if d.probability is not None]
        return sum(p) / len(p) if p else None

    def with_protein(self, protein):
        return Gene(self.source, self.start, self.end, self.strand, protein, self._probability)

    def with_probability(self, probability):
        return Gene(self.source, self.start, self.end, self.strand,
                    self.protein.with_domains([d.with_probability(probability) for d in self.protein.domains]), probability)


def genes_of_case(case, shuffle_seed=None):
    """Gene objects of one tests/golden/ref_loop_cases.json case (domains deliberately out of order)."""
    import random

    genes = []
    for contig in case["contigs"]:
        src = Source(contig["id"])
        for g in contig["genes"]:
            doms = [Domain(n, s, s + 10) for n, s in g["domains"]]
            genes.append(Gene(src, g["start"], g["start"] + 99, 1, Protein(g["id"], None, doms)))
    if shuffle_seed is not None:
        random.Random(shuffle_seed).shuffle(genes)
    return genes
