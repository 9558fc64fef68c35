"""psi_b200 — B200-native server evaluation of the batched-FHE private indexed equality check
(BatchedFHEHIPPIE::run of SAP/nested-hashing-psi) behind the reference's PIE call interface.

The directory name carries the reference's name and is not a Python identifier; import it through
the repo-root shim:  ``import psi_b200``.

Python is only the test / bench harness language here: every operation is a call into the C-ABI
library ``libpsi_b200.so`` (include/psi_b200.h).  There is no CPU fallback: if the library or a
CUDA device is missing the calls raise.
"""
from .capi import (PsiError, PsiParams, lib, lib_path, build_library, params_generate, PLAINTEXT_MODULUS,
                   depth_for_E, MAX_LIMBS)
from .pie import (CryptoContext, MultiContext, PublicKey, TabulationHashing, HierarchicalCuckooHashTable, BatchedFHEHIPPIE,
                  RandomDataInput, client_table, hash_index, FHEHIPPIE, FHEHIPPIECollection,
                  eval_sum_indices, rotation_index)
from .sharding import bin_shard, ShardedPIE, QueryDistributor, query_slice
from .client_query import build_query_slots, extract_intersection

__all__ = [
    "PsiError", "PsiParams", "lib", "lib_path", "build_library", "params_generate", "PLAINTEXT_MODULUS",
    "depth_for_E", "MAX_LIMBS", "CryptoContext", "MultiContext", "PublicKey", "TabulationHashing", "HierarchicalCuckooHashTable",
    "BatchedFHEHIPPIE", "FHEHIPPIE", "FHEHIPPIECollection", "eval_sum_indices", "rotation_index", "RandomDataInput", "client_table", "hash_index", "bin_shard", "ShardedPIE", "QueryDistributor", "query_slice", "build_query_slots", "extract_intersection",
]
