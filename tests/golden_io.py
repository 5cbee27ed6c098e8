"""Loader for the committed golden vectors (see tests/golden/make_golden.py)."""
import json
import os

import numpy as np

import cases as C

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_sort_golden():
    z = np.load(os.path.join(GOLD, "sort_golden.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    out = []
    for i, case in enumerate(meta):
        sims = np.frombuffer(bytes(z[f"sims_{i}"]), dtype=np.dtype(case["sims_dtype"]))
        out.append((case, sims, z[f"idx_{i}"], z[f"sc_{i}"]))
    return out


def load_dist_golden():
    """[(case, distances)] for the euclidean cases: euclidean_metric(V, q, get_similarity_score=False) of the real reference."""
    z = np.load(os.path.join(GOLD, "sort_golden.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return [(case, np.frombuffer(bytes(z[f"dist_{i}"]), dtype=np.dtype(case["sims_dtype"])))
            for i, case in enumerate(meta) if f"dist_{i}" in z.files]


def case_id(case):
    return f"{case['metric'][:4]}-{case['vdt']}x{case['qdt']}-{case['n']}x{case['d']}-{case['kind']}-k{case['k']}" + ("-ts" if case["ts"] else "")


def inputs(case):
    V, q, ts = C.make_inputs(case)
    assert C.digest(V, q, ts) == case["digest"], "NumPy Generator stream drifted: regenerate golden vectors"
    return V, q, ts


def canon_order(scores, k):
    """north_star tie rule applied to a full score vector: (score desc, index asc)."""
    n = len(scores)
    k = max(0, min(k, n))
    return np.lexsort((np.arange(n), -scores))[:k]


def load_tail():
    z = np.load(os.path.join(GOLD, "hyperdb_tail.npz"))
    return z, json.loads(bytes(z["meta"]).decode())


def load_pokemon():
    return np.load(os.path.join(GOLD, "pokemon_c1.npz"))
