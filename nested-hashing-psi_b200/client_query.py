"""Plain-integer restatement of what the CLIENT puts on the wire for one query and how it reads the
answer (out of scope for the GPU, but the server path needs inputs of exactly this shape):

    build_query_slots      BatchedFHEPSIClient.cpp:114-152   one-hot index matrix + (-x) vector
    extract_intersection   BatchedFHEPSIClient.cpp:178-192   zero test over the b decrypted results

Slot s = outerHf * e + outerPos (client `currentTableCount`, :125).  Encryption/decryption of these
slot vectors is the client's business (OpenFHE there, the oracle in this repo's tests).
"""
import numpy as np

from .pie import hash_index


def build_query_slots(hashfunction, client_cells, K, E):
    """client_cells: [k][e] uint64 (0 = empty).  Returns (idx_slots [K][E][k*e], minus_slots [k*e]) int64."""
    k, e = client_cells.shape
    flat = np.ascontiguousarray(client_cells, dtype=np.uint64).reshape(-1)
    n = k * e
    idx = np.zeros((K, E, n), dtype=np.int64)
    minus = np.ones(n, dtype=np.int64)                    # empty slot -> +1   (:128-131)
    occupied = np.nonzero(flat)[0]
    minus[occupied] = -flat[occupied].astype(np.int64)    # -x                 (:138)
    for hf in range(K):                                   # hash ids k .. k+K-1 (:140-147)
        pos = hash_index(hashfunction, flat[occupied], k + hf, E).astype(np.int64)
        idx[hf, pos, occupied] = 1
    return idx, minus


def extract_intersection(client_cells, decrypted):
    """decrypted: [b][>= k*e] slot values of the b result ciphertexts.  A client slot is in the
    intersection iff some bin decrypts to 0 there (:178-192)."""
    flat = np.ascontiguousarray(client_cells, dtype=np.uint64).reshape(-1)
    hit = (np.asarray(decrypted)[:, :flat.shape[0]] == 0).any(axis=0)
    return flat[hit]
