"""CPU oracle for the table path (TEST INFRASTRUCTURE — never imported by gecco_b200/).

A plain-Python restatement of what ``gecco predict`` does to its input tables before the CRF runs and to its
result tables afterwards, row by row like the reference:

* ``GeneTable.load`` / ``FeatureTable.load``           gecco/_base.py:119-131, gecco/model.py:621-637, 773-789
* ``annotate_genes``                                   gecco/cli/commands/_common.py:211-262
* the coordinate sorts                                 gecco/cli/commands/predict.py:81-83
* ``filter_domains``                                   gecco/cli/commands/_common.py:419-448
* ``extract_features_protein`` / ``_domain``           gecco/crf/features.py:13-48
* ``GeneTable.from_genes().dump`` / ``FeatureTable.from_genes().dump``   gecco/model.py:644-670, 791-813,
  gecco/_base.py:133-151 — floats as ``repr()``, the layout of the reference's committed result tables.

Pinned on the reference's own result tables (tests/golden/bgc0001866.json carries their SHA-256).
"""

from __future__ import annotations

import csv
import io
import math
import statistics
from typing import Dict, List, Optional, Sequence


def read_table(text: str) -> List[Dict[str, str]]:
    return list(csv.DictReader(io.StringIO(text, newline=""), dialect="excel-tab"))


def _f(x: str) -> float:
    return math.nan if x == "" else float(x)


def load(genes_text: str, features_texts: Sequence[str], e_filter: Optional[float] = None, p_filter: Optional[float] = 1e-9):
    """-> list of genes (dicts with a ``domains`` list), ordered and filtered like ``gecco predict``."""
    genes = [dict(sequence_id=r["sequence_id"], protein_id=r["protein_id"], start=int(r["start"]), end=int(r["end"]),
                  strand="+" if r["strand"] == "+" else "-", domains=[]) for r in read_table(genes_text)]
    index = {g["protein_id"]: g for g in genes}
    if len(index) < len(genes):
        raise ValueError("Duplicate gene names in input genes")
    for text in features_texts:
        for r in read_table(text):
            g = index[r["protein_id"]]
            if g["sequence_id"] != r["sequence_id"] or g["start"] != int(r["start"]) or g["end"] != int(r["end"]) \
                    or g["strand"] != r["strand"]:
                raise ValueError(f"Mismatched gene for {r['protein_id']!r}")
            g["domains"].append(dict(name=r["domain"], hmm=r["hmm"], i_evalue=_f(r["i_evalue"]), pvalue=_f(r["pvalue"]),
                                     start=int(r["domain_start"]), end=int(r["domain_end"])))
    genes = list(index.values())
    genes.sort(key=lambda g: (g["sequence_id"], g["start"], g["end"]))
    for g in genes:
        g["domains"].sort(key=lambda d: (d["start"], d["end"]))
        if e_filter is not None:
            g["domains"] = [d for d in g["domains"] if d["i_evalue"] < e_filter]
        if p_filter is not None:
            g["domains"] = [d for d in g["domains"] if d["pvalue"] < p_filter]
    return genes


def pack(genes, attr_index: Dict[str, int], feature_type: str = "protein"):
    """-> (contig_ptr, row_ptr, attr_idx, row_gene) as Python lists."""
    contig_ptr, row_ptr, attr_idx, row_gene = [0], [0], [], []
    for k, g in enumerate(genes):
        if k and g["sequence_id"] != genes[k - 1]["sequence_id"]:
            contig_ptr.append(len(row_ptr) - 1)
        if feature_type == "protein":
            for name in dict.fromkeys(d["name"] for d in g["domains"]):
                if name in attr_index:
                    attr_idx.append(attr_index[name])
            row_ptr.append(len(attr_idx))
            row_gene.append(k)
        else:
            for d in g["domains"] or [None]:
                if d is not None and d["name"] in attr_index:
                    attr_idx.append(attr_index[d["name"]])
                row_ptr.append(len(attr_idx))
                row_gene.append(k)
    if genes:
        contig_ptr.append(len(row_ptr) - 1)
    return contig_ptr, row_ptr, attr_idx, row_gene


def _cell(x: float) -> str:
    return "" if x is None or math.isnan(x) else repr(float(x))


def gene_probabilities(genes, row_prob, feature_type: str = "protein"):
    avg, mx, it = [], [], iter(row_prob)
    for g in genes:
        if feature_type == "protein" or not g["domains"]:
            p = next(it)
            avg.append(p)
            mx.append(p)
        else:
            ps = [p for p in (next(it) for _ in g["domains"]) if not math.isnan(p)]
            avg.append(statistics.mean(ps) if ps else math.nan)
            mx.append(max(ps) if ps else math.nan)
    return avg, mx


def dump_genes(genes, row_prob=None, feature_type: str = "protein") -> str:
    nan = [math.nan] * len(genes)
    avg, mx = gene_probabilities(genes, row_prob, feature_type) if row_prob is not None else (nan, nan)
    has_avg, has_max = any(not math.isnan(x) for x in avg), any(not math.isnan(x) for x in mx)
    out = ["\t".join(["sequence_id", "protein_id", "start", "end", "strand"] + ["average_p"] * has_avg + ["max_p"] * has_max)]
    for g, a, m in zip(genes, avg, mx):
        row = [g["sequence_id"], g["protein_id"], str(g["start"]), str(g["end"]), g["strand"]]
        if has_avg:
            row.append(_cell(a))
        if has_max:
            row.append(_cell(m))
        out.append("\t".join(row))
    return "\n".join(out) + "\n"


def dump_features(genes, row_prob=None, feature_type: str = "protein") -> str:
    it = iter(row_prob) if row_prob is not None else None
    rows, any_p = [], False
    for g in genes:
        if it is None:
            ps = [math.nan] * len(g["domains"])
        elif feature_type == "protein" or not g["domains"]:
            ps = [next(it)] * len(g["domains"])
        else:
            ps = [next(it) for _ in g["domains"]]
        for d, p in zip(g["domains"], ps):
            any_p |= not math.isnan(p)
            rows.append(([g["sequence_id"], g["protein_id"], str(g["start"]), str(g["end"]), g["strand"], d["name"], d["hmm"],
                          _cell(d["i_evalue"]), _cell(d["pvalue"]), str(d["start"]), str(d["end"])], p))
    head = ["sequence_id", "protein_id", "start", "end", "strand", "domain", "hmm", "i_evalue", "pvalue", "domain_start",
            "domain_end"] + ["cluster_probability"] * any_p
    out = ["\t".join(head)] + ["\t".join(r + ([_cell(p)] if any_p else [])) for r, p in rows]
    return "\n".join(out) + "\n"
