import hashlib
import io
import pathlib
import pickle
import struct

import numpy
import pytest

from gecco_b200 import model_io

REFERENCE = pathlib.Path("/root/reference/gecco/crf")


def test_bundled_tables(weights):
    assert len(weights.attrs) == 2659 and weights.labels == ["0", "1"]
    assert int(weights.state_mask.sum()) == 4211
    assert weights.window_size == 20 and weights.window_step == 1 and weights.feature_type == "protein"
    t = weights.transition_features_
    assert t[("0", "0")] == 2.669891070463728 and t[("0", "1")] == -2.599571900486168
    assert t[("1", "0")] == -2.6019205422130995 and t[("1", "1")] == 2.5683226020688488
    sf = weights.state_features_
    assert sf[("PF00109", "1")] == 0.3281678016907602 and sf[("PF00109", "0")] == -0.32816780168964776
    assert sum(1 for k in sf if k[1] == "1") == 1822
    assert ("PF99999", "1") not in sf


# ---- a synthetic CRFsuite file + pickle, so that the binary parser is tested without the reference


def _cqdb(keys):
    body = b""
    offsets = []
    for i, k in enumerate(keys):
        offsets.append(24 + 2048 + len(body))
        kb = k.encode() + b"\0"
        body += struct.pack("<II", i, len(kb)) + kb
    tables_at = 24 + 2048 + len(body)
    # one (unused by our reader) hash table so that "records end where the first table begins"
    table = struct.pack("<II", 0, 0) * 2
    refs = [(0, 0)] * 256
    refs[0] = (tables_at, 2)
    bwd_at = tables_at + len(table)
    bwd = b"".join(struct.pack("<I", o) for o in offsets)
    size = bwd_at + len(bwd)
    head = struct.pack("<4sIIIII", b"CQDB", size, 0, 0x62445371, len(keys), bwd_at)
    return head + b"".join(struct.pack("<II", *r) for r in refs) + body + table + bwd


def make_crfsuite_blob(attrs, labels, state, trans):
    feats = [(0, a, l, w) for (a, l), w in state.items()] + [(1, i, j, w) for (i, j), w in trans.items()]
    feat = struct.pack("<4sII", b"FEAT", 12 + 20 * len(feats), len(feats))
    feat += b"".join(struct.pack("<IIId", *f) for f in feats)
    off_features = 48
    off_labels = off_features + len(feat)
    lab = _cqdb(labels)
    off_attrs = off_labels + len(lab)
    att = _cqdb(attrs)
    total = off_attrs + len(att)
    head = struct.pack("<4sI4sIIIIIIIII", b"lCRF", total, b"FOMC", 100, 0, len(labels), len(attrs), off_features,
                       off_labels, off_attrs, 0, 0)
    return head + feat + lab + att


class FileResource:  # pickled under the third-party module names below
    pass


class CRF:
    pass


class ClusterCRF:
    pass


def _write_model_dir(tmp_path, blob, corrupt_md5=False):
    import sys
    import types

    mods = {}
    for mod, cls in (("sklearn_crfsuite._fileresource", FileResource), ("sklearn_crfsuite.estimator", CRF),
                     ("gecco.crf", ClusterCRF)):
        m = types.ModuleType(mod)
        cls.__module__ = mod
        setattr(m, cls.__name__, cls)
        mods[mod] = m
    for parent in ("gecco", "sklearn_crfsuite"):
        pm = types.ModuleType(parent)
        pm.__path__ = []
        mods[parent] = pm
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        res = FileResource(); res.__dict__.update({"suffix": ".crfsuite", "__FILE_RESOURCE_DATA__": blob})
        crf = CRF(); crf.__dict__.update({"c1": 0.4, "c2": 0.0, "modelfile": res})
        obj = ClusterCRF(); obj.__dict__.update({"feature_type": "protein", "window_size": 7, "window_step": 2,
                                                 "algorithm": "lbfgs", "model": crf})
        data = pickle.dumps(obj, protocol=4)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    (tmp_path / "model.pkl").write_bytes(data)
    digest = hashlib.md5(data).hexdigest()
    (tmp_path / "model.pkl.md5").write_text(("0" * 32 if corrupt_md5 else digest.upper()) + "\n")


def test_crfsuite_parser_and_stub_unpickling(tmp_path):
    attrs = ["PF00001", "PF00002", "weird attr"]
    labels = ["1", "0"]  # label id 0 is '1' here: ids must be resolved by name
    state = {(0, 0): 1.5, (0, 1): -1.5, (2, 1): 0.25}
    trans = {(0, 0): 2.0, (0, 1): -1.0, (1, 0): -1.25, (1, 1): 2.5}
    blob = make_crfsuite_blob(attrs, labels, state, trans)
    w = model_io.parse_crfsuite_model(blob)
    assert w.attrs == attrs and w.labels == labels
    assert w.state_w[0, 0] == 1.5 and w.state_w[2, 1] == 0.25 and w.state_w[1].tolist() == [0, 0]
    assert w.state_mask.sum() == 3 and w.trans_w[1, 0] == -1.25
    assert w.label_id("1") == 0
    assert w.state_features_ == {("PF00001", "1"): 1.5, ("PF00001", "0"): -1.5, ("weird attr", "0"): 0.25}

    _write_model_dir(tmp_path, blob)
    loaded = model_io.load_model(tmp_path)
    assert loaded.window_size == 7 and loaded.window_step == 2 and loaded.attrs == attrs
    assert numpy.array_equal(loaded.state_w, w.state_w)

    _write_model_dir(tmp_path, blob, corrupt_md5=True)
    with pytest.raises(ValueError, match="MD5 hash of model data does not match signature"):
        model_io.load_model(tmp_path)

    with pytest.raises(ValueError):
        model_io.parse_crfsuite_model(blob[:-4])
    with pytest.raises(ValueError):
        model_io.parse_crfsuite_model(b"lCRX" + blob[4:])


def test_unpickler_never_executes_pickled_globals(tmp_path):
    class Evil:
        def __reduce__(self):
            return (eval, ("1/0",))

    data = pickle.dumps(Evil(), protocol=4)
    obj = model_io._StubUnpickler(io.BytesIO(data)).load()  # builds a stub instead of calling eval
    assert type(obj).__name__ == "eval"


def test_tsv_round_trip(tmp_path, weights):
    model_io.save_tsv_model(weights, tmp_path)
    back = model_io.load_tsv_model(tmp_path)
    assert back.attrs == weights.attrs and numpy.array_equal(back.state_w, weights.state_w)
    assert numpy.array_equal(back.trans_w, weights.trans_w) and numpy.array_equal(back.state_mask, weights.state_mask)


@pytest.mark.skipif(not REFERENCE.exists(), reason="reference checkout not mounted (GPU box)")
def test_bundled_tables_equal_the_reference_pickle(weights):
    ref = model_io.load_pickled_model(REFERENCE)
    assert ref.attrs == weights.attrs and ref.labels == weights.labels
    assert numpy.array_equal(ref.state_w, weights.state_w) and numpy.array_equal(ref.trans_w, weights.trans_w)
    assert numpy.array_equal(ref.state_mask, weights.state_mask)
    assert (ref.window_size, ref.window_step, ref.feature_type) == (20, 1, "protein")
