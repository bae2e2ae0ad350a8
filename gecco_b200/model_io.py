"""Model loading for the B200 ClusterCRF engine — no sklearn-crfsuite / python-crfsuite needed.

The reference stores its CRF as a pickled ``gecco.crf.ClusterCRF`` whose ``.model`` is a
``sklearn_crfsuite.CRF`` holding the raw CRFsuite binary model
(reference: ``gecco/crf/__init__.py:61-99`` for the MD5-checked load, ``:380-402`` for the save).
Neither third-party package exists on the GPU box, so this module

* verifies the MD5 side-file exactly like the reference (case-insensitive hex compare),
* unpickles with stub classes (nothing from the pickle is ever executed),
* decodes the CRFsuite ``lCRF``/``FOMC`` container into dense tables.

It also reads the ``model.state.tsv`` / ``model.trans.tsv`` pair that ``gecco train`` writes
(reference: ``gecco/cli/commands/train.py:66-85``), which is the format of the weight table bundled
under ``gecco_b200/data`` (a derived table of the v0.11.0 model; see ``tools/make_golden.py``).
"""

from __future__ import annotations

import csv
import hashlib
import io
import os
import pathlib
import pickle
import struct
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import numpy

__all__ = [
    "CRFWeights",
    "parse_crfsuite_model",
    "load_pickled_model",
    "load_tsv_model",
    "load_model",
    "bundled_model_dir",
]


@dataclass
class CRFWeights:
    """Dense view of a first-order CRFsuite model.

    ``state_w[a, l]`` is the weight of state feature (attribute ``a`` -> label ``l``), 0 where the
    model has no such feature; ``state_mask`` remembers which entries exist so that
    ``state_features_`` (``gecco/crf/__init__.py:264``) can answer ``None`` for absent ones.
    ``trans_w[i, j]`` is the weight of transition ``i -> j``.
    """

    attrs: List[str]
    labels: List[str]
    state_w: numpy.ndarray
    state_mask: numpy.ndarray
    trans_w: numpy.ndarray
    feature_type: str = "protein"
    window_size: int = 20
    window_step: int = 1
    extra: Dict[str, Any] = field(default_factory=dict)

    @property
    def attr_index(self) -> Dict[str, int]:
        idx = self.__dict__.get("_attr_index")
        if idx is None:
            idx = {name: i for i, name in enumerate(self.attrs)}
            self.__dict__["_attr_index"] = idx
        return idx

    def label_id(self, label: str) -> int:
        """Label ids are looked up by *name*, as ``Tagger.marginal('1', t)`` does."""
        return self.labels.index(label)

    @property
    def state_features_(self) -> Dict[Tuple[str, str], float]:
        """Same mapping as ``sklearn_crfsuite.CRF.state_features_``: only stored features."""
        d = self.__dict__.get("_state_features")
        if d is None:
            d = {}
            aa, ll = numpy.nonzero(self.state_mask)
            for a, l in zip(aa.tolist(), ll.tolist()):
                d[(self.attrs[a], self.labels[l])] = float(self.state_w[a, l])
            self.__dict__["_state_features"] = d
        return d

    @property
    def transition_features_(self) -> Dict[Tuple[str, str], float]:
        return {
            (self.labels[i], self.labels[j]): float(self.trans_w[i, j])
            for i in range(len(self.labels))
            for j in range(len(self.labels))
        }


# --------------------------------------------------------------------------------------------
# CRFsuite binary container
# --------------------------------------------------------------------------------------------


def _parse_cqdb(blob: bytes, off: int) -> List[str]:
    """Decode one CQDB string<->id dictionary; returns keys ordered by id."""
    magic, size, _flag, byteorder, bwd_size, bwd_offset = struct.unpack_from("<4sIIIII", blob, off)
    if magic != b"CQDB":
        raise ValueError("bad CQDB chunk magic")
    if byteorder != 0x62445371:
        raise ValueError("unsupported CQDB byte order")
    # 256 hash-table references follow the 24-byte header; the key/value records start after them
    # and stop where the first hash table begins.
    refs = struct.unpack_from("<512I", blob, off + 24)
    table_offsets = [refs[2 * i] for i in range(256) if refs[2 * i + 1] != 0 and refs[2 * i] != 0]
    end = min(table_offsets) if table_offsets else size
    if bwd_offset:
        end = min(end, bwd_offset) if bwd_offset > 24 + 2048 else end
    pos = 24 + 2048
    by_id: Dict[int, str] = {}
    while pos + 8 <= end:
        ident, ksize = struct.unpack_from("<II", blob, off + pos)
        key = blob[off + pos + 8 : off + pos + 8 + ksize]
        if ksize == 0 or len(key) != ksize:
            break
        by_id[ident] = key.rstrip(b"\0").decode("utf-8")
        pos += 8 + ksize
    n = len(by_id)
    if sorted(by_id) != list(range(n)):
        raise ValueError("CQDB ids are not a dense range")
    if bwd_size and bwd_size != n:
        # the backward array can be larger than the number of keys only in corrupt files
        raise ValueError("CQDB backward array size does not match the key count")
    return [by_id[i] for i in range(n)]


def parse_crfsuite_model(blob: bytes) -> CRFWeights:
    """Decode a CRFsuite 1st-order model file (magic ``lCRF``, type ``FOMC``, version 100)."""
    if len(blob) < 48:
        raise ValueError("CRFsuite model too short")
    (magic, size, typ, version, _num_features, num_labels, num_attrs,
     off_features, off_labels, off_attrs, _off_labelrefs, _off_attrrefs) = struct.unpack_from(
        "<4sI4sIIIIIIIII", blob, 0
    )
    if magic != b"lCRF" or typ != b"FOMC":
        raise ValueError("not a CRFsuite first-order model")
    if version != 100:
        raise ValueError(f"unsupported CRFsuite model version {version}")
    if size != len(blob):
        raise ValueError("CRFsuite model size field does not match the data length")

    labels = _parse_cqdb(blob, off_labels)
    attrs = _parse_cqdb(blob, off_attrs)
    if len(labels) != num_labels or len(attrs) != num_attrs:
        raise ValueError("CRFsuite dictionaries do not match the header counts")

    fmagic, _chunk, nfeat = struct.unpack_from("<4sII", blob, off_features)
    if fmagic != b"FEAT":
        raise ValueError("bad feature chunk magic")
    rec = numpy.frombuffer(
        blob,
        dtype=numpy.dtype([("type", "<u4"), ("src", "<u4"), ("dst", "<u4"), ("w", "<f8")]),
        count=nfeat,
        offset=off_features + 12,
    )
    state_w = numpy.zeros((num_attrs, num_labels), dtype=numpy.float64)
    state_mask = numpy.zeros((num_attrs, num_labels), dtype=bool)
    trans_w = numpy.zeros((num_labels, num_labels), dtype=numpy.float64)
    st = rec[rec["type"] == 0]
    tr = rec[rec["type"] == 1]
    if len(st) + len(tr) != nfeat:
        raise ValueError("unknown feature type in CRFsuite model")
    # CRFsuite sums every matching feature; duplicates do not occur in practice but add up if so
    numpy.add.at(state_w, (st["src"].astype(numpy.int64), st["dst"].astype(numpy.int64)), st["w"])
    state_mask[st["src"], st["dst"]] = True
    numpy.add.at(trans_w, (tr["src"].astype(numpy.int64), tr["dst"].astype(numpy.int64)), tr["w"])
    return CRFWeights(attrs=attrs, labels=labels, state_w=state_w, state_mask=state_mask, trans_w=trans_w)


# --------------------------------------------------------------------------------------------
# pickled ClusterCRF
# --------------------------------------------------------------------------------------------


class _Stub:
    """Stand-in for every class named by the pickle; only absorbs state."""

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        pass

    def __setstate__(self, state: Any) -> None:
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state


class _StubUnpickler(pickle.Unpickler):
    _SAFE_BUILTINS = {"frozenset", "set", "dict", "list", "tuple", "bytes", "bytearray", "str", "int",
                      "float", "bool", "complex", "slice", "range", "object"}

    def find_class(self, module: str, name: str) -> Any:
        if module == "builtins" and name in self._SAFE_BUILTINS:
            return super().find_class(module, name)
        if module == "collections" and name == "OrderedDict":
            return super().find_class(module, name)
        return type(name, (_Stub,), {"__module__": module})


def _md5_matches(pkl: bytes, signature: str) -> bool:
    # gecco/crf/__init__.py:96 — upper-cased compare of the hex digest and the stripped signature
    return hashlib.md5(pkl).hexdigest().upper() == signature.strip().upper()


def _open_child(base: Any, name: str, mode: str):
    if hasattr(base, "joinpath"):
        return base.joinpath(name).open(mode)
    return open(os.path.join(os.fspath(base), name), mode)


def load_pickled_model(model_dir: Any, filename: str = "model.pkl") -> CRFWeights:
    """Load ``model.pkl`` (+ ``model.pkl.md5``) written by ``ClusterCRF.save``.

    Raises ``ValueError("MD5 hash of model data does not match signature")`` like the reference
    (``gecco/crf/__init__.py:96-97``).
    """
    with _open_child(model_dir, filename, "rb") as f:
        data = f.read()
    with _open_child(model_dir, f"{filename}.md5", "r") as f:
        signature = f.read()
    if not _md5_matches(data, signature):
        raise ValueError("MD5 hash of model data does not match signature")
    obj = _StubUnpickler(io.BytesIO(data)).load()
    state = obj.__dict__
    crf = state.get("model")
    if crf is None:
        raise ValueError("pickled ClusterCRF holds no fitted model")
    resource = crf.__dict__.get("modelfile")
    blob = None if resource is None else resource.__dict__.get("__FILE_RESOURCE_DATA__")
    if not isinstance(blob, (bytes, bytearray)):
        raise ValueError("pickled CRF does not embed a CRFsuite model")
    weights = parse_crfsuite_model(bytes(blob))
    weights.feature_type = state.get("feature_type", "protein")
    weights.window_size = int(state.get("window_size", 5))
    weights.window_step = int(state.get("window_step", 1))
    weights.extra = {
        "algorithm": state.get("algorithm"),
        "significance": state.get("significance"),
        "significant_features": state.get("significant_features"),
        "c1": crf.__dict__.get("c1"),
        "c2": crf.__dict__.get("c2"),
    }
    return weights


# --------------------------------------------------------------------------------------------
# `gecco train` weight tables
# --------------------------------------------------------------------------------------------


def load_tsv_model(model_dir: Any) -> CRFWeights:
    """Load ``model.state.tsv`` + ``model.trans.tsv`` (``gecco/cli/commands/train.py:66-85``).

    An optional ``model.meta.tsv`` (key/value) carries ``feature_type``, ``window_size``,
    ``window_step`` and the label order; without it the GECCO defaults of the shipped model apply
    and labels are ordered as first seen in the transition table.
    """
    meta: Dict[str, str] = {}
    try:
        with _open_child(model_dir, "model.meta.tsv", "r") as f:
            for row in csv.reader(f, dialect="excel-tab"):
                if len(row) >= 2 and not row[0].startswith("#"):
                    meta[row[0]] = row[1]
    except FileNotFoundError:
        pass

    labels: List[str] = meta["labels"].split(",") if "labels" in meta else []
    trans: List[Tuple[str, str, float]] = []
    with _open_child(model_dir, "model.trans.tsv", "r") as f:
        rows = csv.reader(f, dialect="excel-tab")
        header = next(rows)
        if header[:3] != ["from", "to", "weight"]:
            raise ValueError("unexpected header in model.trans.tsv")
        for src, dst, w in rows:
            trans.append((src, dst, float(w)))
            for l in (src, dst):
                if l not in labels:
                    labels.append(l)

    attrs: List[str] = []
    index: Dict[str, int] = {}
    entries: List[Tuple[int, str, float]] = []
    with _open_child(model_dir, "model.state.tsv", "r") as f:
        rows = csv.reader(f, dialect="excel-tab")
        header = next(rows)
        if header[:3] != ["attr", "label", "weight"]:
            raise ValueError("unexpected header in model.state.tsv")
        for attr, label, w in rows:
            a = index.get(attr)
            if a is None:
                a = index[attr] = len(attrs)
                attrs.append(attr)
            if label not in labels:
                labels.append(label)
            entries.append((a, label, float(w)))

    L = len(labels)
    state_w = numpy.zeros((len(attrs), L), dtype=numpy.float64)
    state_mask = numpy.zeros((len(attrs), L), dtype=bool)
    for a, label, w in entries:
        l = labels.index(label)
        state_w[a, l] += w
        state_mask[a, l] = True
    trans_w = numpy.zeros((L, L), dtype=numpy.float64)
    for src, dst, w in trans:
        trans_w[labels.index(src), labels.index(dst)] += w
    return CRFWeights(
        attrs=attrs,
        labels=labels,
        state_w=state_w,
        state_mask=state_mask,
        trans_w=trans_w,
        feature_type=meta.get("feature_type", "protein"),
        window_size=int(meta.get("window_size", 20)),
        window_step=int(meta.get("window_step", 1)),
    )


def save_tsv_model(weights: CRFWeights, model_dir: Any) -> None:
    """Write the three-table form read by `load_tsv_model` (floats as ``repr`` => exact round trip)."""
    out = pathlib.Path(os.fspath(model_dir))
    out.mkdir(parents=True, exist_ok=True)
    with open(out / "model.trans.tsv", "w", newline="") as f:
        w = csv.writer(f, dialect="excel-tab")
        w.writerow(["from", "to", "weight"])
        for (src, dst), weight in weights.transition_features_.items():
            w.writerow([src, dst, repr(weight)])
    with open(out / "model.state.tsv", "w", newline="") as f:
        w = csv.writer(f, dialect="excel-tab")
        w.writerow(["attr", "label", "weight"])
        for (attr, label), weight in weights.state_features_.items():
            w.writerow([attr, label, repr(weight)])
    with open(out / "model.meta.tsv", "w", newline="") as f:
        w = csv.writer(f, dialect="excel-tab")
        w.writerow(["feature_type", weights.feature_type])
        w.writerow(["window_size", weights.window_size])
        w.writerow(["window_step", weights.window_step])
        w.writerow(["labels", ",".join(weights.labels)])


def bundled_model_dir() -> pathlib.Path:
    """Directory of the weight tables derived from the model shipped with GECCO v0.11.0."""
    return pathlib.Path(__file__).resolve().parent / "data" / "gecco-0.11.0"


def load_model(model_path: Any = None) -> CRFWeights:
    """Resolve a model like ``ClusterCRF.trained`` does (``gecco/crf/__init__.py:78-82``).

    ``None`` -> the model embedded in an installed ``gecco`` package if importable, else the bundled
    derived tables; a directory with ``model.pkl`` -> MD5-checked pickle; a directory with
    ``model.state.tsv`` -> weight tables.
    """
    if model_path is None:
        try:
            from importlib.resources import files

            embedded = files("gecco.crf")
            if embedded.joinpath("model.pkl").is_file():
                return load_pickled_model(embedded)
        except (ImportError, ModuleNotFoundError, FileNotFoundError, TypeError):
            pass
        return load_tsv_model(bundled_model_dir())
    has_pkl = False
    try:
        if hasattr(model_path, "joinpath") and not isinstance(model_path, (str, os.PathLike)):
            has_pkl = model_path.joinpath("model.pkl").is_file()
        else:
            model_path = pathlib.Path(os.fspath(model_path))
            has_pkl = (model_path / "model.pkl").is_file()
    except OSError:
        has_pkl = False
    if has_pkl:
        return load_pickled_model(model_path)
    return load_tsv_model(model_path)
