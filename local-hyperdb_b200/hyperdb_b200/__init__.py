"""hyperdb_b200 -- B200-native brute-force ranking engine behind local-hyperDB's own API.

Drop-in for `hyperdb/ranking_algorithm.py` (module `hyperdb_b200.ranking_algorithm`) and for the
brute-force branch of `HyperDB.query` (`hyperdb_b200.hyperdb.HyperDB`).  All arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI of include/hyperdb_b200.h
(lib/libhyperdb_b200.so, bound with ctypes in `_native`).  There is no CPU fallback: importing the
package works anywhere, computing requires the built library and a CUDA device.
"""
from . import ranking_algorithm  # noqa: F401
from .device_matrix import DeviceMatrix  # noqa: F401
from .ranking_algorithm import (  # noqa: F401
    cosine_similarity,
    custom_ranking_algorithm_sort,
    dot_product,
    euclidean_metric,
    get_norm_vector,
    hamming_distance,
    hyperDB_ranking_algorithm_sort,
    jaccard_similarity,
    manhattan_distance,
    pearson_correlation,
)

__all__ = [
    "DeviceMatrix", "ranking_algorithm", "cosine_similarity", "dot_product", "euclidean_metric",
    "manhattan_distance", "hamming_distance", "jaccard_similarity", "pearson_correlation", "get_norm_vector", "hyperDB_ranking_algorithm_sort",
    "custom_ranking_algorithm_sort",
]
