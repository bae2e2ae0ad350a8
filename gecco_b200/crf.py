"""Drop-in replacement for ``gecco.crf.ClusterCRF`` whose inference runs on a B200.

GECCO injects the CRF class everywhere through ``gecco.cli.main(..., crf_type=...)``
(``gecco/cli/commands/__init__.py:127-168``); the pipeline only calls ``crf_type.trained(model)`` and
``.predict_probabilities(genes, pad=, progress=)`` (``gecco/cli/commands/_common.py:565-592``).  This class
keeps those two entry points, their argument meaning, warnings and exceptions, and replaces the per-window
Python loop around the third-party tagger (``gecco/crf/__init__.py:244-258``) by ONE call into
``libgecco_crf_b200.so``.

``Gene`` / ``Protein`` / ``Domain`` objects are duck-typed (``gecco.model`` needs Biopython at import time and
is never imported here): what is used is ``gene.source.id``, ``gene.start``, ``gene.protein.domains``
(``.name``, ``.start``), and the copy-on-write helpers ``with_probability`` / ``with_protein`` /
``with_domains`` / ``with_cluster_weight`` (``gecco/model.py:110-375``).

Differences from the reference, all deliberate:

* ``predict_probabilities`` on an unfitted object raises ``NotFittedError`` (a ``ValueError`` subclass; the
  sklearn one when scikit-learn is importable).  The reference names ``NotFittedError`` without importing it
  and dies with ``NameError`` (``gecco/crf/__init__.py:186``).
* ``feature_type="domain"`` sizes the probability vector by the number of feature rows.  The reference sizes it
  by genes and fails with a shape error as soon as a gene has two domains (``:251`` vs ``:254``).
* ``progress`` is called with ``(0, total)`` and ``(total, total)`` — the windows are not evaluated one by one.
* arithmetic: the drop-in computes in f64 like the reference's tagger (``GCRF_FLAG_F64``; results within 1e-12 of
  python-crfsuite's, so thresholded cluster calls cannot flip) — the Python object handling around the call costs
  orders of magnitude more than the wider arithmetic.  ``GECCO_B200_ARITHMETIC=f32`` (or ``crf.arithmetic = "f32"``)
  selects the FP32 streaming kernels of the bulk array API (within 1e-5).
* any ``window_size`` works, like in the reference (``:134-137``): sizes beyond what the FP32 kernels hold in shared
  memory (``CRFEngine.max_window``: 128 for Pfam-sized models) run the f64 path.
* training (``fit`` / ``save``) is delegated to the reference class when GECCO and sklearn-crfsuite are
  installed; it is out of scope here.
"""

from __future__ import annotations

import os
import typing
import warnings
from typing import Any, Callable, Dict, FrozenSet, Iterable, List, Optional, Tuple, Union

import numpy

from . import model_io
from .model_io import CRFWeights
from .packer import PackedGenes, pack_genes

__all__ = ["ClusterCRF", "NotFittedError"]

try:  # same exception type a scikit-learn user would catch
    from sklearn.exceptions import NotFittedError  # type: ignore
except Exception:  # pragma: no cover - scikit-learn is optional

    class NotFittedError(ValueError, AttributeError):  # type: ignore
        """Raised when predicting with a ``ClusterCRF`` that holds no model."""


class _TaggerView:
    """What GECCO reads from ``ClusterCRF.model`` besides the marginals: the feature dictionaries
    (``gecco/crf/__init__.py:264``, ``gecco/cli/commands/train.py:73,84``)."""

    def __init__(self, weights: CRFWeights):
        self.weights = weights

    @property
    def state_features_(self) -> Dict[Tuple[str, str], float]:
        return self.weights.state_features_

    @property
    def transition_features_(self) -> Dict[Tuple[str, str], float]:
        return self.weights.transition_features_

    @property
    def attributes_(self) -> List[str]:
        return list(self.weights.attrs)

    @property
    def classes_(self) -> List[str]:
        return list(self.weights.labels)


class ClusterCRF(object):
    """A linear-chain CRF over genes, evaluated by hand-written sm_100a kernels."""

    _FILENAME = "model.pkl"

    # ------------------------------------------------------------------ construction
    @classmethod
    def trained(cls, model_path: Union[Any, str, None] = None) -> "ClusterCRF":
        """Create a pre-trained instance (``gecco/crf/__init__.py:61-99``).

        ``model_path`` is a directory holding ``model.pkl`` + ``model.pkl.md5`` (as written by
        ``gecco train``; MD5-checked, ``ValueError`` on mismatch) or the ``model.state.tsv`` /
        ``model.trans.tsv`` tables; ``None`` uses the model embedded in an installed GECCO, else the
        bundled tables derived from the v0.11.0 model.
        """
        weights = model_io.load_model(model_path)
        self = cls(weights.feature_type, window_size=weights.window_size, window_step=weights.window_step)
        self._set_weights(weights)

        extra = weights.extra or {}
        self.algorithm = extra.get("algorithm") or self.algorithm
        self.significance = extra.get("significance")
        self.significant_features = extra.get("significant_features")
        return self

    def __init__(self, feature_type: str = "protein", algorithm: str = "lbfgs", window_size: int = 5,
                 window_step: int = 1, **kwargs: Any) -> None:
        # gecco/crf/__init__.py:132-137
        if feature_type not in {"protein", "domain"}:
            raise ValueError(f"invalid feature type: {feature_type!r}")
        if window_size <= 0:
            raise ValueError("Window size must be strictly positive")
        if window_step <= 0 or window_step > window_size:
            raise ValueError("Window step must be strictly positive and under `window_size`")
        self.feature_type = feature_type
        self.window_size = window_size
        self.window_step = window_step
        self.algorithm = algorithm
        self.significance: Optional[Dict[str, float]] = None
        self.significant_features: Optional[FrozenSet[str]] = None
        self.model: Optional[_TaggerView] = None
        self._options = {"algorithm": algorithm, **kwargs}
        self._weights: Optional[CRFWeights] = None
        self._engine = None
        self.device = int(os.environ.get("GECCO_B200_DEVICE", "0"))
        self.arithmetic = os.environ.get("GECCO_B200_ARITHMETIC", "f64")
        if self.arithmetic not in ("f32", "f64"):
            raise ValueError(f"invalid arithmetic: {self.arithmetic!r} (expected 'f32' or 'f64')")

    def _set_weights(self, weights: CRFWeights) -> None:
        self._weights = weights
        self.model = _TaggerView(weights)
        self._engine = None

    # ------------------------------------------------------------------ device
    def _get_engine(self):
        if self._engine is None:
            from ._lib import CRFEngine  # raises if the CUDA library is missing: there is no CPU fallback

            assert self._weights is not None
            self._engine = CRFEngine(self._weights, device=self.device)
            self._check_window(self._engine)
        return self._engine

    def _check_window(self, engine) -> None:
        # the reference accepts any window_size >= 1 (gecco/crf/__init__.py:134-137) and so does the f64 path; the
        # FP32 kernels keep a window's state in shared memory and stop at engine.max_window() (128 for Pfam-sized models)
        if self.arithmetic == "f32" and self.window_size > engine.max_window(False):
            warnings.warn(f"window_size {self.window_size} exceeds the FP32 kernels' limit ({engine.max_window(False)}); "
                          "using the f64 path")
            self.arithmetic = "f64"

    def marginals(self, packed: PackedGenes, *, pad: bool = True) -> numpy.ndarray:
        """Bulk entry point: per-row cluster probability of an already packed batch (NaN = skipped contig)."""
        if self.model is None:
            raise NotFittedError("This ClusterCRF instance is not fitted yet.")
        extra = {"accessions": True} if getattr(packed, "accessions", False) else {}
        engine = self._get_engine()
        if self.arithmetic == "f64":
            extra["f64_arith"] = True
        return engine.marginals_windowed(packed.contig_ptr, packed.gene_ptr, packed.attr_idx,
                                                     window=self.window_size, step=self.window_step, pad=pad, **extra)

    # ------------------------------------------------------------------ the hot path
    def predict_probabilities(self, genes: Iterable[Any], *, pad: bool = True,
                              progress: Optional[Callable[[int, int], None]] = None) -> List[Any]:
        """Predict how likely each given gene is part of a gene cluster (``gecco/crf/__init__.py:148-273``).

        Returns new ``Gene`` objects, ordered by contig id then start, with their probability set and every
        domain annotated with its CRF state weight.
        """
        _progress = progress or (lambda x, y: None)
        if self.model is None:
            raise NotFittedError("This ClusterCRF instance is not fitted yet.")
        if self.feature_type not in ("protein", "domain"):
            raise ValueError(f"invalid feature type: {self.feature_type!r}")
        assert self._weights is not None
        W = self.window_size

        # :199-206 — sort genes, sort each gene's domains in place, group by contig; :209-213 — features
        packed, genes, slices = pack_genes(genes, self._weights.attr_index, self.feature_type)

        # :216-236 — contigs shorter than the window: pad with a warning, or skip with a warning
        rows = numpy.diff(packed.contig_ptr)
        skipped = numpy.zeros(packed.C, dtype=bool)
        total = 0
        for c in numpy.flatnonzero(rows < W).tolist():
            first, last = slices[c]
            n_rows, n_genes = int(rows[c]), last - first
            if pad:
                unit = self.feature_type if W - n_rows == 1 else f"{self.feature_type}s"
                warnings.warn(
                    f"Contig {genes[first].source.id!r} does not contain enough"
                    f" {self.feature_type}s ({n_genes}) for sliding window"
                    f" of size {W}, padding with"
                    f" {W - n_rows} {unit}"
                )
            else:
                warnings.warn(
                    f"Contig {genes[first].source.id!r} does not contain enough"
                    f" {self.feature_type}s ({n_genes}) for sliding window"
                    f" of size {W}"
                )
                skipped[c] = True
        # :239 — the reference counts len(feats) - W + 1 windows per contig (also when step > 1)
        total = int(numpy.maximum(rows[~skipped], W).sum() - (W - 1) * int((~skipped).sum()))
        _progress(0, total)

        # :244-256 — every window of every contig, max-pooled: one library call
        prob = self.marginals(packed, pad=pad) if packed.G else numpy.zeros(0)
        _progress(total, total)

        # :258 — write the probabilities back
        predicted: List[Any] = []
        for c, (first, last) in enumerate(slices):
            contig = genes[first:last]
            if skipped[c]:
                predicted.extend(contig)  # :246-248 — returned without a probability
                continue
            p = prob[packed.contig_ptr[c]:packed.contig_ptr[c + 1]]
            if self.feature_type == "protein":
                # gecco/crf/features.py:74-96
                if len(p) != len(contig):
                    raise ValueError("gene and probability lists don't have the same length")
                predicted.extend(gene.with_probability(float(x)) for gene, x in zip(contig, p))
            else:
                # gecco/crf/features.py:99-120
                it = iter(p.tolist())
                for gene in contig:
                    if gene.protein.domains:
                        predicted.append(gene.with_protein(gene.protein.with_domains(
                            [domain.with_probability(next(it)) for domain in gene.protein.domains])))
                    else:
                        predicted.append(gene.with_probability(next(it)))
                if next(it, None) is not None:
                    raise ValueError("gene and probability lists don't have the same length")

        # :261-269 — label domains with their weight for the positive label (None when the model has none)
        weights = self.model.state_features_
        return [
            gene.with_protein(gene.protein.with_domains(
                domain.with_cluster_weight(weights.get((domain.name, "1"))) for domain in gene.protein.domains
            ))
            for gene in predicted
        ]

    # ------------------------------------------------------------------ training: delegated
    def _reference(self):
        try:
            import gecco.crf  # type: ignore
            import sklearn_crfsuite  # type: ignore  # noqa: F401
        except ImportError as err:
            raise NotImplementedError(
                "training is not part of the B200 inference engine; install gecco-tool and sklearn-crfsuite to "
                "fit models, then load them with ClusterCRF.trained(path)"
            ) from err
        ref = gecco.crf.ClusterCRF(self.feature_type, window_size=self.window_size, window_step=self.window_step,
                                   **self._options)
        return ref

    def fit(self, genes: Iterable[Any], *, select: Optional[float] = None, shuffle: bool = True,
            cpus: Optional[int] = None, correction_method: Optional[str] = None) -> None:
        """Fit with the reference implementation (``gecco/crf/__init__.py:275-378``), then adopt its weights."""
        import tempfile

        ref = self._reference()
        ref.fit(genes, select=select, shuffle=shuffle, cpus=cpus, correction_method=correction_method)
        self.significance = ref.significance
        self.significant_features = ref.significant_features
        with tempfile.TemporaryDirectory() as tmp:
            ref.save(tmp)
            self._fitted_reference = ref
            self._set_weights(model_io.load_pickled_model(tmp))

    def save(self, model_path: "os.PathLike[str]") -> None:
        """``model.pkl`` + MD5 through the reference class when this object was fitted by it
        (``gecco/crf/__init__.py:380-402``); the weight tables otherwise."""
        ref = getattr(self, "_fitted_reference", None)
        if ref is not None:
            ref.save(model_path)
            return
        if self._weights is None:
            raise NotFittedError("This ClusterCRF instance is not fitted yet.")
        weights = self._weights
        weights.feature_type, weights.window_size, weights.window_step = self.feature_type, self.window_size, self.window_step
        model_io.save_tsv_model(weights, model_path)


def _predict_tables_main(argv: List[str]) -> int:
    """``python -m gecco_b200 predict-tables``: the table-to-table part of ``gecco predict`` without GECCO's Python
    object model (same option names and defaults as ``gecco/cli/commands/_parser.py:171-337``)."""
    import argparse

    from .tables import predict_tables

    ap = argparse.ArgumentParser(prog="python -m gecco_b200 predict-tables",
                                 description="genes + features tables -> CRF marginals on the B200 -> genes / features / clusters tables")
    ap.add_argument("--genes", required=True, help="a genes table (*.genes.tsv, optionally compressed)")
    ap.add_argument("--features", required=True, nargs="+", help="one or more features tables (*.features.tsv)")
    ap.add_argument("-o", "--output-dir", default=".", help="where to write {base}.genes.tsv / .features.tsv / .clusters.tsv")
    ap.add_argument("--base", default=None, help="output file prefix (default: derived from the genes table's name)")
    ap.add_argument("--model", default=None, help="directory with model.pkl (+ .md5) or model.state.tsv / model.trans.tsv")
    ap.add_argument("-e", "--e-filter", type=float, default=None, help="drop domains with an i-evalue over this")
    ap.add_argument("-p", "--p-filter", type=float, default=1e-9, help="drop domains with a p-value over this")
    ap.add_argument("--no-pad", dest="pad", action="store_false", help="skip contigs shorter than the window instead of padding")
    ap.add_argument("-m", "--threshold", type=float, default=0.8, help="probability threshold of cluster membership")
    ap.add_argument("-c", "--cds", type=int, default=3, help="minimum number of annotated genes in a cluster")
    ap.add_argument("-E", "--edge-distance", type=int, default=0, help="annotated genes this close to a contig edge do not count")
    ap.add_argument("--no-trim", dest="trim", action="store_false", help="keep un-annotated genes at cluster edges")
    args = ap.parse_args(argv)
    tables, prob = predict_tables(args.genes, args.features, args.output_dir, model=args.model, base=args.base,
                                  e_filter=args.e_filter, p_filter=args.p_filter, pad=args.pad, threshold=args.threshold,
                                  cds=args.cds, edge_distance=args.edge_distance, trim=args.trim)
    print(f"{tables.genes} genes on {tables.contigs} contigs, {tables.domains} domains -> {args.output_dir}")
    tables.close()
    return 0


def main(argv: Optional[List[str]] = None) -> int:
    """``python -m gecco_b200 ...`` = ``gecco ...`` with the CRF swapped for this one; ``predict-tables`` is this
    package's own table-to-table command."""
    import sys

    argv = list(sys.argv[1:] if argv is None else argv)
    if argv and argv[0] == "predict-tables":
        return _predict_tables_main(argv[1:])
    try:
        from gecco.cli import main as gecco_main  # type: ignore
    except ImportError as err:  # pragma: no cover - GECCO is not installed in the build image
        raise SystemExit(f"gecco-tool is not installed ({err}); use gecco_b200.ClusterCRF from Python instead")
    return gecco_main(argv, crf_type=ClusterCRF)
