"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference)
on the seeded inputs of cases.py.  Run in the build container only (the GPU box has no
/root/reference); the outputs are committed.

    python tests/golden/make_golden.py

Files written:
  sort_golden.npz    per case: input digest, the metric function's full output (bytes of the
                     reference's own dtype), and hyperDB_ranking_algorithm_sort's (indices, scores);
                     euclidean cases also hold euclidean_metric(..., get_similarity_score=False)
  pokemon_c1.npz     BASELINE config C1: demo/pokemon_hyperdb.pickle vectors (fp32, exact) + the
                     reference's cosine top-5 for several stored rows used as queries
  hyperdb_tail.npz   HyperDB.query brute-force tail run through the real HyperDB class (third-party
                     imports stubbed as in SURVEY.md appendix A): recency, skip_doc, metadata filter
"""
from __future__ import annotations

import importlib.util
import io
import contextlib
import json
import os
import pickle
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases as C  # noqa: E402

REF_ROOT = "/root/reference"


def load_ranking():
    spec = importlib.util.spec_from_file_location("ref_ranking", f"{REF_ROOT}/hyperdb/ranking_algorithm.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_hyperdb_class():
    """stub the four absent third-party imports; no model download."""
    if not hasattr(np, "float_"):
        np.float_ = np.float64

    def _mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    _mod("onnxruntime", set_default_logger_severity=lambda *a, **k: None,
         InferenceSession=type("InferenceSession", (), {}))
    _mod("pympler", asizeof=_mod("pympler.asizeof", asizeof=lambda o: sys.getsizeof(o)))
    _mod("fast_sentence_transformers", FastSentenceTransformer=type("FastSentenceTransformer", (), {}))

    class AnnoyIndex:
        def __init__(self, *a, **k): pass
        def add_item(self, *a): pass
        def build(self, *a, **k): pass
        def save(self, *a): pass
        def load(self, *a, **k): pass
        def unload(self, *a, **k): pass

    _mod("annoy", AnnoyIndex=AnnoyIndex)
    sys.path.insert(0, REF_ROOT)
    import hyperdb.hyperdb as H
    H.HyperDB.initialize_model = lambda self: None
    return H


def gen_sort(ref):
    out, meta = {}, []
    fn = {m: getattr(ref, m) for m in C.METRICS}
    for i, case in enumerate(C.sort_cases()):
        V, q, ts = C.make_inputs(case)
        rec = dict(case)
        rec["digest"] = C.digest(V, q, ts)
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
            sims = np.asarray(fn[case["metric"]](V.copy(), q.copy()))
            idx, sc = ref.hyperDB_ranking_algorithm_sort(
                V.copy(), q.copy(), top_k=case["k"], metric=case["metric"],
                timestamps=ts, recency_bias=case["bias"] if ts is not None else 0)
        rec["sims_dtype"] = str(sims.dtype)
        out[f"sims_{i}"] = np.frombuffer(np.ascontiguousarray(sims).tobytes(), np.uint8)
        if case["metric"] == "euclidean_metric":         # the distance form, ranking_algorithm.py:49-52 (same dtype as sims)
            with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
                dist = np.asarray(ref.euclidean_metric(V.copy(), q.copy(), get_similarity_score=False))
            assert dist.dtype == sims.dtype
            out[f"dist_{i}"] = np.frombuffer(np.ascontiguousarray(dist).tobytes(), np.uint8)
        out[f"idx_{i}"] = np.asarray(idx, np.int64)
        out[f"sc_{i}"] = np.asarray(sc, np.float64).reshape(-1)
        meta.append(rec)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "sort_golden.npz"), **out)
    print("sort cases:", len(meta))


def gen_pokemon(ref):
    db = pickle.load(open(f"{REF_ROOT}/demo/pokemon_hyperdb.pickle", "rb"))
    V64 = np.asarray(db["vectors"])
    V32 = V64.astype(np.float32)
    assert np.array_equal(V32.astype(np.float64), V64)
    names = np.array([d["name"] for d in db["documents"]])
    out = dict(vectors=V32, names=names)
    for row in (0, 25, 77, 150):
        for tag, V, q in (("f32", V32, V32[row]), ("f64", V64, V64[row]), ("mixed", V64, V32[row])):
            with contextlib.redirect_stdout(io.StringIO()):
                idx, sc = ref.hyperDB_ranking_algorithm_sort(V.copy(), q.copy(), top_k=5, metric="cosine_similarity")
            out[f"idx_{tag}_{row}"] = np.asarray(idx, np.int64)
            out[f"sc_{tag}_{row}"] = np.asarray(sc, np.float64)
    np.savez_compressed(os.path.join(HERE, "pokemon_c1.npz"), **out)
    print("pokemon: ", out["idx_f32_0"], out["sc_f32_0"])


def gen_tail():
    H = load_hyperdb_class()
    rng = np.random.default_rng(77)
    n, d = 120, 32
    V = rng.standard_normal((n, d)).astype(np.float32)
    ts = (1.7e9 + rng.uniform(0, 4.0, n)).round(3)
    docs = [{"id": i, "group": ["a", "b", "c"][i % 3], "timestamp": float(ts[i])} for i in range(n)]
    queries = rng.standard_normal((4, d)).astype(np.float32)

    def embed(documents, **kw):          # embedding_function contract: (vectors, source_indices, split_info)
        ids = [doc["id"] for doc in documents]
        return V[ids], list(ids), {i: 1 for i in ids}

    with contextlib.redirect_stdout(io.StringIO()):
        db = H.HyperDB(documents=None, embedding_function=embed, ann_metric="hamming",
                       metadata_keys=["group", "timestamp"], fp_precision="float32")
        db.add(docs)
    assert db.vectors.shape == (n, d), db.vectors.shape
    out = dict(V=V, ts=ts, queries=queries, groups=np.array([doc["group"] for doc in docs]))
    runs = []
    specs = [
        dict(metric="cosine_similarity", top_k=7, recency_bias=0, filters=None),
        dict(metric="dot_product", top_k=5, recency_bias=0.5, filters=None),
        dict(metric="euclidean_metric", top_k=6, recency_bias=0.25, filters=[("skip_doc", 10)]),
        dict(metric="manhattan_distance", top_k=6, recency_bias=0, filters=[("skip_doc", -15)]),
        dict(metric="cosine_similarity", top_k=9, recency_bias=0.3, filters=[("metadata", {"group": "b"})]),
        dict(metric="dot_product", top_k=200, recency_bias=0, filters=[("metadata", {"group": "c"})]),
    ]
    for si, spec in enumerate(specs):
        for qi in range(len(queries)):
            with contextlib.redirect_stdout(io.StringIO()):
                res = db.query(queries[qi], timestamp_key="timestamp" if spec["recency_bias"] else None, **spec)
            out[f"ids_{si}_{qi}"] = np.array([doc["id"] for doc, _s, _i in res], np.int64)
            out[f"sc_{si}_{qi}"] = np.array([float(np.asarray(s).reshape(-1)[0]) for _d, s, _i in res], np.float64)
        runs.append({k: v for k, v in spec.items()})
    out["meta"] = np.frombuffer(json.dumps(runs).encode(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "hyperdb_tail.npz"), **out)
    print("tail specs:", len(runs), "first:", out["ids_0_0"], out["sc_0_0"])


if __name__ == "__main__":
    ref = load_ranking()
    gen_sort(ref)
    gen_pokemon(ref)
    gen_tail()
