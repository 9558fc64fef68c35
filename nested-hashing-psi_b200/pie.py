"""Host-side mirror of the reference's objects around the batched FHE PIE, over the C ABI.

    BatchedFHEHIPPIE(cryptoContext, pK, hct)   src/Common/Crypto/PrivateIndexedEqualityCheck/BatchedFHEHIPPIE.hpp:30-31
      .setIndex(indexMatrix)                   :40-43   K x E ciphertexts
      .setMinusCompareElement(ct)              :45-48
      .run()                                   :33      BatchedFHEHIPPIE.cpp:88-129
      .getResultList()                         :35-38   b ciphertexts
    HierarchicalCuckooHashTable(hashfunction, eachSimpleTableSize, eachCuckooTableSize, serverStashSize,
                                numberOfSimpleHashFunctions, numberOfCuckooHashFunctions, simpleMultiTable,
                                cuckooMultiTable, maxItemsPerPosition)   src/Common/Hashing/HierarchicalCuckooHashTable.cpp:16-53
    TabulationHashing(seed, numberOfHashfunctions)                      src/Common/Hashing/TabulationHashing.cpp:16-36
    RandomDataInput(serverSetSize, clientSetSize, intersectionSetSize, seed, bitSize)  src/Common/DataInput/RandomDataInput.cpp:10-29

Ciphertexts are numpy uint64 arrays [2][L][N] (the limbs DCRTPoly::GetElementAtIndex(l).GetValues()
exposes), EVALUATION format.  Errors follow the reference: invalid arguments raise ValueError with the
reference's message (std::invalid_argument there), cuckoo insertion failure raises RuntimeError.
"""
import ctypes

import numpy as np

from . import capi
from .capi import PsiError, check, lib


def _u64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))


def _i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def _raise_like_reference(err):
    """std::invalid_argument -> ValueError, std::runtime_error (hashing) -> RuntimeError."""
    if err.status == capi.PSI_ERR_INVALID:
        raise ValueError(err.message) from None
    raise err


SEED_RANDOM = 0xFFFFFFFFFFFFFFFF   # PSI_SEED_RANDOM


def _seed(s):
    return SEED_RANDOM if s is None else int(s)


class PublicKey:
    """Stored but never used by run() (BatchedFHEHIPPIE.hpp:22); kept for signature parity."""


def _evk_shape(params):
    """psi_set_relin_key layouts: BV [L][L][N] (digit, limb); HYBRID [numPartQ][L + Lk][N]."""
    if params.ks_technique == capi.KS_HYBRID:
        return (params.ks_num_parts, params.L + params.Lk, params.N)
    return (params.L, params.L, params.N)


class CryptoContext:
    """What the PIE needs of lbcrypto::CryptoContext<DCRTPoly>: the BFV-RNS parameter tables, the
    relinearisation key (DeserializeEvalMultKey, BatchedFHEPSIServer.cpp:49) and the device evaluator."""

    def __init__(self, params, device=0):
        if not isinstance(params, capi.PsiParams):   # any byte-compatible mirror of struct psi_params
            if ctypes.sizeof(params) != ctypes.sizeof(capi.PsiParams):
                raise ValueError("params is not a struct psi_params")
            params = capi.PsiParams.from_buffer_copy(bytes(params))
        self.params = params
        self.device = device
        self.N, self.L, self.Lp, self.t = params.N, params.L, params.Lp, params.t
        h = ctypes.c_void_p()
        check(lib().psi_ctx_create(ctypes.byref(params), device, ctypes.byref(h)))
        self._h = h

    @classmethod
    def generate(cls, N, t, depth, device=0, L=0):
        return cls(capi.params_generate(N, t, depth, L), device)

    def GetPlaintextModulus(self):
        return self.t

    def GetRingDimension(self):
        return self.N

    def InsertEvalMultKey(self, evk_b, evk_a):
        (evk_b, pb), (evk_a, pa) = _u64(evk_b), _u64(evk_a)
        assert evk_b.shape == _evk_shape(self.params) and evk_a.shape == evk_b.shape
        check(lib().psi_set_relin_key(self._h, pb, pa))

    # --- raw C-ABI level (the PIE class sits on top of these) ----------------------------------
    def db_load_limbs(self, pt, mask):
        (pt, pp), (mask, pm) = _u64(pt), _u64(mask)
        K, b, E = pt.shape[:3]
        assert pt.shape == (K, b, E, self.L, self.N) and mask.shape == (b, self.L, self.N)
        check(lib().psi_db_load_limbs(self._h, K, b, E, pp, pm))
        self._dims = (K, b, E)

    def db_encode_slots(self, slots, mask_slots):
        (slots, ps), (mask_slots, pm) = _i64(slots), _i64(mask_slots)
        K, b, E, n = slots.shape
        assert mask_slots.shape == (b, n)
        try:
            check(lib().psi_db_encode_slots(self._h, K, b, E, n, ps, pm))
        except PsiError as e:
            _raise_like_reference(e)
        self._dims = (K, b, E)

    def set_encode_lift(self, mode):
        """0 = PSI_ENCODE_LIFT_PLAIN, 1 = PSI_ENCODE_LIFT_CENTRED; before the encoding db_* calls."""
        check(lib().psi_set_encode_lift(self._h, mode))

    def set_packing_cofactor(self, mode):
        """0 = PSI_PACK_COFACTOR_3, 1 = PSI_PACK_COFACTOR_CONJ; before the encoding db_* calls."""
        check(lib().psi_set_packing_cofactor(self._h, mode))

    def db_load_limbs_shard(self, pt, mask, bin_begin, bin_end):
        """Full database arrays; bins [bin_begin, bin_end) become resident."""
        (pt, pp), (mask, pm) = _u64(pt), _u64(mask)
        K, b, E = pt.shape[:3]
        check(lib().psi_db_load_limbs_shard(self._h, K, b, bin_begin, bin_end, E, pp, pm))
        self._dims = (K, bin_end - bin_begin, E)

    def db_encode_slots_shard(self, slots, mask_slots, bin_begin, bin_end):
        (slots, ps), (mask_slots, pm) = _i64(slots), _i64(mask_slots)
        K, b, E, n = slots.shape
        try:
            check(lib().psi_db_encode_slots_shard(self._h, K, b, bin_begin, bin_end, E, n, ps, pm))
        except PsiError as e:
            _raise_like_reference(e)
        self._dims = (K, bin_end - bin_begin, E)

    def db_build_from_items_shard(self, hashfunction, k, e, K, E, b, items, bin_begin, bin_end, evictionSeed=0x5EED,
                                  shuffleSeed=1, maskSeed=2):
        """One shard of the device-built database; all shards of one database share the three seeds."""
        (items, pi) = _u64(items)
        try:
            check(lib().psi_db_build_from_items_shard(self._h, hashfunction.seed, k, e, K, E, b, evictionSeed, pi,
                                                      items.shape[0], _seed(shuffleSeed), _seed(maskSeed), bin_begin, bin_end))
        except PsiError as err:
            if err.status == capi.PSI_ERR_STATE:
                raise RuntimeError(err.message) from None
            _raise_like_reference(err)
        self._dims = (K, bin_end - bin_begin, E)

    def db_get_bin_limbs(self, bin_):
        """(pt [K][E][L][N], mask [L][N]) of one resident bin."""
        K, b, E = self._dims
        pt = np.empty((K, E, self.L, self.N), dtype=np.uint64)
        mask = np.empty((self.L, self.N), dtype=np.uint64)
        p = ctypes.POINTER(ctypes.c_uint64)
        check(lib().psi_db_get_bin_limbs(self._h, bin_, pt.ctypes.data_as(p), mask.ctypes.data_as(p)))
        return pt, mask

    def db_get_limbs(self):
        K, b, E = self._dims
        pt = np.empty((K, b, E, self.L, self.N), dtype=np.uint64)
        mask = np.empty((b, self.L, self.N), dtype=np.uint64)
        check(lib().psi_db_get_limbs(self._h, pt.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                                     mask.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))))
        return pt, mask

    def hct_build_device(self, hashfunction, k, e, K, E, b, items, evictionSeed=0x5EED):
        """Nested cuckoo table built on this context's GPU; cells [k][e][K][b][E], identical to the host build."""
        (items, pi) = _u64(items)
        cells = np.empty((k, e, K, b, E), dtype=np.uint64)
        try:
            check(lib().psi_hct_build_device(self._h, hashfunction.seed, k, e, K, E, b, evictionSeed, pi, items.shape[0],
                                             cells.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))))
        except PsiError as err:
            if err.status == capi.PSI_ERR_STATE:
                raise RuntimeError(err.message) from None
            _raise_like_reference(err)
        return cells

    def db_build_from_items(self, hashfunction, k, e, K, E, b, items, evictionSeed=0x5EED, shuffleSeed=None,
                            maskSeed=None):
        """The PIE constructor's whole data path on the device (table build, shuffle, transposition, encode).
        shuffleSeed / maskSeed: None = fresh std::random_device seeds (the reference's behaviour; the masks are
        security-critical), explicit values only for reproducible tests."""
        (items, pi) = _u64(items)
        shuffleSeed, maskSeed = _seed(shuffleSeed), _seed(maskSeed)
        try:
            check(lib().psi_db_build_from_items(self._h, hashfunction.seed, k, e, K, E, b, evictionSeed, pi, items.shape[0],
                                                shuffleSeed, maskSeed))
        except PsiError as err:
            if err.status == capi.PSI_ERR_STATE:
                raise RuntimeError(err.message) from None
            _raise_like_reference(err)
        self._dims = (K, b, E)

    def query_set(self, idx, minus, stream=None):
        (idx, pi), (minus, pm) = _u64(idx), _u64(minus)
        K, b, E = self._dims
        assert idx.shape == (K, E, 2, self.L, self.N) and minus.shape == (2, self.L, self.N)
        check(lib().psi_query_set(self._h, pi, pm, stream))
        self._keep = (idx, minus)  # the copy is asynchronous: keep the host buffers alive

    def query_set_ptr(self, idx_ptr, minus_ptr, stream=None):
        """Same, from raw (pinned) host addresses."""
        check(lib().psi_query_set(self._h, ctypes.cast(idx_ptr, ctypes.POINTER(ctypes.c_uint64)),
                                  ctypes.cast(minus_ptr, ctypes.POINTER(ctypes.c_uint64)), stream))

    def query_upload_ptr(self, idx_ptr, minus_ptr, stream=None):
        """H2D of the next query into landing buffers only (overlaps a running evaluation)."""
        check(lib().psi_query_upload(self._h, ctypes.cast(idx_ptr, ctypes.POINTER(ctypes.c_uint64)),
                                     ctypes.cast(minus_ptr, ctypes.POINTER(ctypes.c_uint64)), stream))

    def query_commit(self, stream=None):
        check(lib().psi_query_commit(self._h, stream))

    def query_upload_limbs(self, idx_vectors, minus_vectors, stream=None):
        """Upload from K*E*2*L + 2*L separately allocated uint64[N] arrays ([hf][pos][comp][limb] order); accepts
        prepared ctypes pointer arrays too (bench.py builds them once)."""
        ai = idx_vectors if isinstance(idx_vectors, ctypes.Array) else MultiContext._ptr_array(idx_vectors)
        am = minus_vectors if isinstance(minus_vectors, ctypes.Array) else MultiContext._ptr_array(minus_vectors)
        check(lib().psi_query_upload_limbs(self._h, ai, am, stream))

    def result_get_limbs(self, out_vectors=None, stream=None):
        """Results into b*2*L separate uint64[N] arrays ([bin][comp][limb] order); synchronous."""
        K, b, E = self._dims
        vecs = None
        if out_vectors is None:
            vecs = [np.empty(self.N, dtype=np.uint64) for _ in range(b * 2 * self.L)]
            out_vectors = MultiContext._ptr_array(vecs)
        elif not isinstance(out_vectors, ctypes.Array):
            vecs, out_vectors = out_vectors, MultiContext._ptr_array(out_vectors)
        check(lib().psi_result_get_limbs(self._h, out_vectors, stream))
        return vecs

    def query_run_streamed_limbs(self, idx_vectors, minus_vectors, out_vectors=None, stream=None):
        """One query from K*E*2*L + 2*L separate limb vectors into b*2*L result vectors with gather / upload /
        evaluation / download / scatter overlapped inside the query (psi_query_run_streamed_limbs); synchronous."""
        K, b, E = self._dims
        ai = idx_vectors if isinstance(idx_vectors, ctypes.Array) else MultiContext._ptr_array(idx_vectors)
        am = minus_vectors if isinstance(minus_vectors, ctypes.Array) else MultiContext._ptr_array(minus_vectors)
        vecs = None
        if out_vectors is None:
            vecs = [np.empty(self.N, dtype=np.uint64) for _ in range(b * 2 * self.L)]
            out_vectors = MultiContext._ptr_array(vecs)
        elif not isinstance(out_vectors, ctypes.Array):
            vecs, out_vectors = out_vectors, MultiContext._ptr_array(out_vectors)
        check(lib().psi_query_run_streamed_limbs(self._h, ai, am, out_vectors, stream))
        return vecs

    def set_host_threads(self, n):
        check(lib().psi_set_host_threads(self._h, n))

    def set_tuning(self, mac_variant=-1, phase2_groups=-1):
        """psi_debug_set_tuning: force the inner-product bin-block width / the number of phase-2 bin groups."""
        check(lib().psi_debug_set_tuning(self._h, mac_variant, phase2_groups))

    def query_landing_ptrs(self, which):
        """(idx_ptr, idx_bytes, minus_ptr, minus_bytes) of device landing buffer `which` (0 or 1)."""
        pi, pm, ni, nm = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_size_t()
        check(lib().psi_query_landing_ptr(self._h, which, ctypes.byref(pi), ctypes.byref(ni), ctypes.byref(pm),
                                          ctypes.byref(nm)))
        return pi.value, ni.value, pm.value, nm.value

    def query_next_landing(self):
        """Landing buffer the next query goes to (they are used in turn)."""
        w = ctypes.c_uint32()
        check(lib().psi_query_next_landing(self._h, ctypes.byref(w)))
        return w.value

    def query_uploaded(self, which):
        """Declares landing buffer `which` filled by the caller (sharding.QueryDistributor); query_commit follows."""
        check(lib().psi_query_uploaded(self._h, which))

    def query_run_streamed_ptr(self, idx_ptr, minus_ptr, out_ptr, stream=None):
        """One query host -> host with upload slices / evaluation / download groups overlapped (raw addresses)."""
        p = ctypes.POINTER(ctypes.c_uint64)
        check(lib().psi_query_run_streamed(self._h, ctypes.cast(idx_ptr, p), ctypes.cast(minus_ptr, p), ctypes.cast(out_ptr, p), stream))

    def query_run_streamed(self, idx, minus, stream=None):
        (idx, pi), (minus, pm) = _u64(idx), _u64(minus)
        K, b, E = self._dims
        assert idx.shape == (K, E, 2, self.L, self.N) and minus.shape == (2, self.L, self.N)
        out = np.empty((b, 2, self.L, self.N), dtype=np.uint64)
        check(lib().psi_query_run_streamed(self._h, pi, pm, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), stream))
        check(lib().psi_stream_sync(stream))
        return out

    def run(self, stream=None, phases=3):
        """psi_run / psi_run_phases: 1 = inner products only, 2 = ct x ct + mask only, 3 = all."""
        check(lib().psi_run_phases(self._h, phases, stream))

    def result_get(self, out=None, stream=None, sync=True):
        K, b, E = self._dims
        if out is None:
            out = np.empty((b, 2, self.L, self.N), dtype=np.uint64)
        check(lib().psi_result_get(self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), stream))
        if sync:
            check(lib().psi_stream_sync(stream))
        return out

    def result_get_ptr(self, out_ptr, stream=None):
        check(lib().psi_result_get(self._h, ctypes.cast(out_ptr, ctypes.POINTER(ctypes.c_uint64)), stream))

    def sync(self, stream=None):
        check(lib().psi_stream_sync(stream))

    def run_launch_count(self):
        n = ctypes.c_uint32()
        check(lib().psi_run_launch_count(self._h, ctypes.byref(n)))
        return n.value

    def result_device_ptr(self):
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        check(lib().psi_result_device_ptr(self._h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def debug_ntt(self, polys, moduli, inverse=False):
        polys = np.array(polys, dtype=np.uint64, order="C", copy=True)
        moduli = np.ascontiguousarray(moduli, dtype=np.uint32)
        assert polys.ndim == 2 and polys.shape[1] == self.N and moduli.shape == (polys.shape[0],)
        check(lib().psi_debug_ntt(self._h, polys.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                                  moduli.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), polys.shape[0],
                                  1 if inverse else 0))
        return polys

    def debug_mul_ctct(self, ct1, ct2):
        (ct1, p1), (ct2, p2) = _u64(ct1), _u64(ct2)
        out = np.empty((2, self.L, self.N), dtype=np.uint64)
        check(lib().psi_debug_mul_ctct(self._h, p1, p2, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))))
        return out

    # --- non-batched FHEHIPPIE (FHEHIPPIE.cpp, SimpleFHEPSIServer.cpp) ---------------------------
    def InsertEvalAutomorphismKeys(self, indices, key_b, key_a):
        """DeserializeEvalSumKey + DeserializeEvalAutomorphismKey (SimpleFHEPSIServer.cpp:45-62):
        indices [n] automorphism indices, key_b / key_a [n][L][L][N] (BV) or [n][numPartQ][L + Lk][N] (HYBRID)."""
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        (key_b, pb), (key_a, pa) = _u64(key_b), _u64(key_a)
        assert key_b.shape == (len(idx),) + _evk_shape(self.params) and key_a.shape == key_b.shape
        check(lib().psi_nb_set_automorphism_keys(self._h, len(idx), idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), pb, pa))

    def nb_db_load_limbs(self, pt, mask, merge_pt):
        (pt, pp), (mask, pm), (merge_pt, pg) = _u64(pt), _u64(mask), _u64(merge_pt)
        n_pie, K, b = pt.shape[:3]
        assert pt.shape == (n_pie, K, b, self.L, self.N) and mask.shape == (n_pie, K, self.L, self.N)
        assert merge_pt.shape == (self.L, self.N)
        try:
            check(lib().psi_nb_db_load_limbs(self._h, n_pie, K, b, pp, pm, pg))
        except PsiError as e:
            _raise_like_reference(e)
        self._nb_dims = (n_pie, K, b)

    def nb_db_encode_slots(self, slots, mask_slots):
        """slots [n_pie][K][b][nslots] (plainVec of FHEHIPPIE.cpp:44-50), mask_slots [n_pie][K][b]."""
        (slots, ps), (mask_slots, pm) = _i64(slots), _i64(mask_slots)
        n_pie, K, b, n = slots.shape
        assert mask_slots.shape == (n_pie, K, b)
        try:
            check(lib().psi_nb_db_encode_slots(self._h, n_pie, K, b, n, ps, pm))
        except PsiError as e:
            _raise_like_reference(e)
        self._nb_dims = (n_pie, K, b)

    def nb_db_get_limbs(self):
        n_pie, K, b = self._nb_dims
        pt = np.empty((n_pie, K, b, self.L, self.N), dtype=np.uint64)
        mask = np.empty((n_pie, K, self.L, self.N), dtype=np.uint64)
        merge = np.empty((self.L, self.N), dtype=np.uint64)
        p = ctypes.POINTER(ctypes.c_uint64)
        check(lib().psi_nb_db_get_limbs(self._h, pt.ctypes.data_as(p), mask.ctypes.data_as(p), merge.ctypes.data_as(p)))
        return pt, mask, merge

    def nb_run(self, idx, pie_begin=0, pie_end=None, stream=None):
        """idx [n][K][2][L][N] for PIEs pie_begin .. pie_end-1 -> results [n][K][2][L][N] (natural hf order)."""
        n_pie, K, b = self._nb_dims
        pie_end = n_pie if pie_end is None else pie_end
        idx, pi = _u64(idx)
        assert idx.shape == (pie_end - pie_begin, K, 2, self.L, self.N)
        out = np.empty_like(idx)
        check(lib().psi_nb_run(self._h, pie_begin, pie_end, pi, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), stream))
        return out

    def nb_launch_count(self):
        n = ctypes.c_uint32()
        check(lib().psi_nb_launch_count(self._h, ctypes.byref(n)))
        return n.value

    def close(self):
        if getattr(self, "_h", None):
            lib().psi_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiContext:
    """CryptoContext over a device list: the single-process multi-device evaluator (psi_multi_*).  Same raw calls as
    CryptoContext; the arrays are always the FULL database / query / result, sharding happens inside the library."""

    def __init__(self, params, devices=(0,)):
        if not isinstance(params, capi.PsiParams):
            if ctypes.sizeof(params) != ctypes.sizeof(capi.PsiParams):
                raise ValueError("params is not a struct psi_params")
            params = capi.PsiParams.from_buffer_copy(bytes(params))
        self.params = params
        self.devices = list(devices)
        self.N, self.L, self.Lp, self.t = params.N, params.L, params.Lp, params.t
        h = ctypes.c_void_p()
        arr = (ctypes.c_int * len(self.devices))(*self.devices)
        check(lib().psi_multi_create(ctypes.byref(params), arr, len(self.devices), ctypes.byref(h)))
        self._h = h

    def GetPlaintextModulus(self):
        return self.t

    def set_encode_lift(self, mode):
        check(lib().psi_multi_set_encode_lift(self._h, mode))

    def InsertEvalMultKey(self, evk_b, evk_a):
        (evk_b, pb), (evk_a, pa) = _u64(evk_b), _u64(evk_a)
        assert evk_b.shape == _evk_shape(self.params) and evk_a.shape == evk_b.shape
        check(lib().psi_multi_set_relin_key(self._h, pb, pa))

    def db_load_limbs(self, pt, mask):
        (pt, pp), (mask, pm) = _u64(pt), _u64(mask)
        K, b, E = pt.shape[:3]
        assert pt.shape == (K, b, E, self.L, self.N) and mask.shape == (b, self.L, self.N)
        try:
            check(lib().psi_multi_db_load_limbs(self._h, K, b, E, pp, pm))
        except PsiError as e:
            _raise_like_reference(e)
        self._dims = (K, b, E)

    def db_encode_slots(self, slots, mask_slots):
        (slots, ps), (mask_slots, pm) = _i64(slots), _i64(mask_slots)
        K, b, E, n = slots.shape
        assert mask_slots.shape == (b, n)
        try:
            check(lib().psi_multi_db_encode_slots(self._h, K, b, E, n, ps, pm))
        except PsiError as e:
            _raise_like_reference(e)
        self._dims = (K, b, E)

    def db_build_from_items(self, hashfunction, k, e, K, E, b, items, evictionSeed=0x5EED, shuffleSeed=None, maskSeed=None):
        (items, pi) = _u64(items)
        try:
            check(lib().psi_multi_db_build_from_items(self._h, hashfunction.seed, k, e, K, E, b, evictionSeed, pi,
                                                      items.shape[0], _seed(shuffleSeed), _seed(maskSeed)))
        except PsiError as err:
            if err.status == capi.PSI_ERR_STATE:
                raise RuntimeError(err.message) from None
            _raise_like_reference(err)
        self._dims = (K, b, E)

    def bin_ranges(self):
        out = []
        for d in range(len(self.devices)):
            b0, b1 = ctypes.c_uint32(), ctypes.c_uint32()
            check(lib().psi_multi_bin_range(self._h, d, ctypes.byref(b0), ctypes.byref(b1)))
            out.append((b0.value, b1.value))
        return out

    def query_set(self, idx, minus):
        (idx, pi), (minus, pm) = _u64(idx), _u64(minus)
        K, b, E = self._dims
        assert idx.shape == (K, E, 2, self.L, self.N) and minus.shape == (2, self.L, self.N)
        check(lib().psi_multi_query_set(self._h, pi, pm))
        self._keep = (idx, minus)

    def query_set_ptr(self, idx_ptr, minus_ptr):
        check(lib().psi_multi_query_set(self._h, ctypes.cast(idx_ptr, ctypes.POINTER(ctypes.c_uint64)),
                                        ctypes.cast(minus_ptr, ctypes.POINTER(ctypes.c_uint64))))

    @staticmethod
    def _ptr_array(vectors):
        P64 = ctypes.POINTER(ctypes.c_uint64)
        arr = (P64 * len(vectors))()
        for i, v in enumerate(vectors):
            assert v.dtype == np.uint64 and v.flags["C_CONTIGUOUS"]
            arr[i] = v.ctypes.data_as(P64)
        return arr

    def query_set_limbs(self, idx_vectors, minus_vectors):
        """idx_vectors: K*E*2*L separately allocated uint64[N] arrays in [hf][pos][comp][limb] order (what a
        deserialised OpenFHE query holds), minus_vectors: 2*L of them."""
        K, b, E = self._dims
        assert len(idx_vectors) == K * E * 2 * self.L and len(minus_vectors) == 2 * self.L
        ai, am = self._ptr_array(idx_vectors), self._ptr_array(minus_vectors)
        check(lib().psi_multi_query_set_limbs(self._h, ai, am))

    def run(self):
        check(lib().psi_multi_run(self._h))

    def result_get(self, out=None, sync=True):
        K, b, E = self._dims
        if out is None:
            out = np.empty((b, 2, self.L, self.N), dtype=np.uint64)
        check(lib().psi_multi_result_get(self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))))
        if sync:
            check(lib().psi_multi_sync(self._h))
        return out

    def result_get_ptr(self, out_ptr):
        check(lib().psi_multi_result_get(self._h, ctypes.cast(out_ptr, ctypes.POINTER(ctypes.c_uint64))))

    def result_get_limbs(self):
        """b*2*L freshly allocated uint64[N] arrays in [bin][comp][limb] order."""
        K, b, E = self._dims
        vecs = [np.empty(self.N, dtype=np.uint64) for _ in range(b * 2 * self.L)]
        arr = self._ptr_array(vecs)
        check(lib().psi_multi_result_get_limbs(self._h, arr))
        return vecs

    def query_run_limbs(self, idx_vectors, minus_vectors):
        """psi_multi_query_run_limbs: one query, limb vectors in, b*2*L freshly allocated result vectors out."""
        K, b, E = self._dims
        assert len(idx_vectors) == K * E * 2 * self.L and len(minus_vectors) == 2 * self.L
        vecs = [np.empty(self.N, dtype=np.uint64) for _ in range(b * 2 * self.L)]
        check(lib().psi_multi_query_run_limbs(self._h, self._ptr_array(idx_vectors), self._ptr_array(minus_vectors),
                                              self._ptr_array(vecs)))
        return vecs

    def sync(self):
        check(lib().psi_multi_sync(self._h))

    def run_launch_count(self):
        n = ctypes.c_uint32()
        check(lib().psi_multi_run_launch_count(self._h, ctypes.byref(n)))
        return n.value

    # --- non-batched FHEHIPPIE collection over the devices -----------------------------------------
    def InsertEvalAutomorphismKeys(self, indices, key_b, key_a):
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        (key_b, pb), (key_a, pa) = _u64(key_b), _u64(key_a)
        check(lib().psi_multi_nb_set_automorphism_keys(self._h, len(idx), idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), pb, pa))

    def nb_db_load_limbs(self, pt, mask, merge_pt):
        (pt, pp), (mask, pm), (merge_pt, pg) = _u64(pt), _u64(mask), _u64(merge_pt)
        n_pie, K, b = pt.shape[:3]
        check(lib().psi_multi_nb_db_load_limbs(self._h, n_pie, K, b, pp, pm, pg))
        self._nb_dims = (n_pie, K, b)

    def nb_db_encode_slots(self, slots, mask_slots):
        (slots, ps), (mask_slots, pm) = _i64(slots), _i64(mask_slots)
        n_pie, K, b, n = slots.shape
        try:
            check(lib().psi_multi_nb_db_encode_slots(self._h, n_pie, K, b, n, ps, pm))
        except PsiError as e:
            _raise_like_reference(e)
        self._nb_dims = (n_pie, K, b)

    def nb_pie_ranges(self):
        out = []
        for i in range(len(self.devices)):
            a, b = ctypes.c_uint32(), ctypes.c_uint32()
            check(lib().psi_multi_nb_pie_range(self._h, i, ctypes.byref(a), ctypes.byref(b)))
            out.append((a.value, b.value))
        return out

    def nb_run(self, idx, pie_begin=0, pie_end=None):
        n_pie, K, b = self._nb_dims
        assert pie_begin == 0 and pie_end in (None, n_pie), "the multi-device collection runs as a whole"
        idx, pi = _u64(idx)
        assert idx.shape == (n_pie, K, 2, self.L, self.N)
        out = np.empty_like(idx)
        check(lib().psi_multi_nb_run(self._h, pi, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))))
        return out

    def close(self):
        if getattr(self, "_h", None):
            lib().psi_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TabulationHashing:
    """TabulationHashing(seed, numberOfHashfunctions) — tables are derived inside the library from
    (seed, count), so this object only carries the two numbers."""

    def __init__(self, seed=342797434736, numberOfHashfunctions=3):
        self.seed = int(seed)
        self.numberOfHashfunctions = int(numberOfHashfunctions)


def hash_index(hashfunction, items, hfInd, tableSize):
    """calculateHashIndex (HashUtils.cpp:34-37), vectorised over items."""
    (items, pi) = _u64(np.atleast_1d(items))
    out = np.empty(items.shape[0], dtype=np.uint64)
    check(lib().psi_hash_index(hashfunction.seed, hashfunction.numberOfHashfunctions, pi, items.shape[0], hfInd,
                               tableSize, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))))
    return out


class HierarchicalCuckooHashTable:
    def __init__(self, hashfunction, eachSimpleTableSize, eachCuckooTableSize, serverStashSize=0,
                 numberOfSimpleHashFunctions=2, numberOfCuckooHashFunctions=2, simpleMultiTable=False,
                 cuckooMultiTable=True, maxItemsPerPosition=1, evictionSeed=0x5EED):
        if hashfunction.numberOfHashfunctions < numberOfSimpleHashFunctions + numberOfCuckooHashFunctions:
            raise ValueError("hash function object provides too few hash functions")
        self.hashfunction = hashfunction
        self.k, self.e = numberOfSimpleHashFunctions, eachSimpleTableSize
        self.K, self.E, self.b = numberOfCuckooHashFunctions, eachCuckooTableSize, maxItemsPerPosition
        self.stash = serverStashSize
        h = ctypes.c_void_p()
        try:
            check(lib().psi_hct_create(hashfunction.seed, self.k, self.e, self.K, self.E, self.b, serverStashSize,
                                       int(simpleMultiTable), int(cuckooMultiTable), evictionSeed, ctypes.byref(h)))
        except PsiError as e:
            _raise_like_reference(e)
        self._h = h

    def insertAll(self, elements):
        (elements, pe) = _u64(elements)
        try:
            check(lib().psi_hct_insert_all(self._h, pe, elements.shape[0]))
        except PsiError as e:
            if e.status == capi.PSI_ERR_STATE:
                raise RuntimeError(e.message) from None
            _raise_like_reference(e)

    def cells(self):
        """[k][e][K][b][E] — hierarchicalCuckooTable[outerHf][outerPos].cuckooTable[innerHf][bin][pos]."""
        out = np.empty((self.k, self.e, self.K, self.b, self.E), dtype=np.uint64)
        check(lib().psi_hct_get_cells(self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))))
        return out

    def getServerStashSize(self):
        return self.stash

    def getEachBinSize(self):
        return self.b

    def getEachCuckooTableSize(self):
        return self.E

    def getNumberOfCuckooHashFunctions(self):
        return self.K

    def getNumberOfSimpleTables(self):
        return self.k

    def getEachSimpleTableSize(self):
        return self.e

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().psi_hct_destroy(self._h)
                self._h = None
        except Exception:
            pass


def client_table(hashfunction, k, e, K, items, evictionSeed=0xC11E):
    """The client's cuckoo table (BatchedFHEPSIClient.cpp:97-99,109): [k][e], 0 = empty."""
    (items, pi) = _u64(items)
    out = np.empty((k, e), dtype=np.uint64)
    try:
        check(lib().psi_client_table(hashfunction.seed, k, e, K, pi, items.shape[0], evictionSeed,
                                     out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))))
    except PsiError as err:
        if err.status == capi.PSI_ERR_STATE:
            raise RuntimeError(err.message) from None
        _raise_like_reference(err)
    return out


class RandomDataInput:
    def __init__(self, serverSetSize, clientSetSize, intersectionSetSize, setGenerationSeed=123456789, bitSize=32):
        self.serverSet = np.empty(serverSetSize, dtype=np.uint64)
        self.clientSet = np.empty(clientSetSize, dtype=np.uint64)
        self.intersectionSet = np.empty(intersectionSetSize, dtype=np.uint64)
        p = ctypes.POINTER(ctypes.c_uint64)
        try:
            check(lib().psi_random_data_input(serverSetSize, clientSetSize, intersectionSetSize, setGenerationSeed,
                                              bitSize, self.serverSet.ctypes.data_as(p),
                                              self.clientSet.ctypes.data_as(p), self.intersectionSet.ctypes.data_as(p)))
        except PsiError as e:
            _raise_like_reference(e)

    def getServerSet(self):
        return self.serverSet

    def getClientSet(self):
        return self.clientSet

    def getIntersectionSet(self):
        return self.intersectionSet


class BatchedFHEHIPPIE:
    """The reference's operator, same five entry points.  The constructor mutates `hct` (bin shuffle),
    exactly as BatchedFHEHIPPIE.cpp:28-35 does."""

    def __init__(self, cryptoContext, pK, hct, shuffleSeed=None, maskSeed=None, keepSlots=False):
        """shuffleSeed / maskSeed: None (default) draws them from std::random_device like the reference
        (BatchedFHEHIPPIE.cpp:25-26); the masks are security-critical, explicit seeds are for tests only."""
        self.cryptoContext = cryptoContext
        self.pK = pK
        h = ctypes.c_void_p()
        create = lib().psi_pie_create_multi if isinstance(cryptoContext, MultiContext) else lib().psi_pie_create
        try:
            check(create(cryptoContext._h, ctypes.byref(cryptoContext.params), hct._h, _seed(shuffleSeed),
                         _seed(maskSeed), int(keepSlots), ctypes.byref(h)))
        except PsiError as e:
            _raise_like_reference(e)
        self._h = h
        K, b, E, n = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        check(lib().psi_pie_dims(h, ctypes.byref(K), ctypes.byref(b), ctypes.byref(E), ctypes.byref(n)))
        self.K, self.b, self.E, self.batchSize = K.value, b.value, E.value, n.value
        cryptoContext._dims = (self.K, self.b, self.E)
        self.indexMatrix = None
        self.minusCompareElement = None
        self._uploaded = False
        self._resultList = None

    @classmethod
    def fromServerSet(cls, cryptoContext, pK, hashfunction, eachSimpleTableSize, eachCuckooTableSize,
                      numberOfSimpleHashFunctions, numberOfCuckooHashFunctions, maxItemsPerPosition, serverSet,
                      evictionSeed=0x5EED, shuffleSeed=None, maskSeed=None):
        """Offline phase entirely on the device: HierarchicalCuckooHashTable::insertAll + the PIE constructor
        (table build, bin shuffle, transposition, MakePackedPlaintext, SetFormat) without the table ever visiting
        the host.  Same database as BatchedFHEHIPPIE(ctx, pk, host-built table) with the same seeds."""
        self = cls.__new__(cls)
        self.cryptoContext, self.pK, self._h = cryptoContext, pK, None
        k, e, K, E, b = (numberOfSimpleHashFunctions, eachSimpleTableSize, numberOfCuckooHashFunctions,
                         eachCuckooTableSize, maxItemsPerPosition)
        cryptoContext.db_build_from_items(hashfunction, k, e, K, E, b, serverSet, evictionSeed, shuffleSeed, maskSeed)
        self.K, self.b, self.E, self.batchSize = K, b, E, k * e
        self.indexMatrix = self.minusCompareElement = self._resultList = None
        self._uploaded = False
        return self

    def setIndex(self, indexMatrix):
        """indexMatrix: K x E ciphertexts (nested lists of [2][L][N] arrays, or one [K][E][2][L][N] array)."""
        self.indexMatrix = indexMatrix
        self._uploaded = False

    def setMinusCompareElement(self, minusCompareElement):
        self.minusCompareElement = minusCompareElement
        self._uploaded = False

    def _upload(self):
        if self._uploaded:
            return
        cc = self.cryptoContext
        if self.indexMatrix is None or self.minusCompareElement is None:
            raise ValueError("setIndex and setMinusCompareElement must be called before run()")
        idx = np.ascontiguousarray(np.asarray(self.indexMatrix, dtype=np.uint64))
        if idx.shape != (self.K, self.E, 2, cc.L, cc.N):
            raise ValueError("indexMatrix must hold K x E ciphertexts of [2][L][N] limbs")
        (idx, pi), (minus, pm) = _u64(idx), _u64(self.minusCompareElement)
        if minus.shape != (2, cc.L, cc.N):
            raise ValueError("minusCompareElement must be one ciphertext of [2][L][N] limbs")
        try:
            if self._h is None:       # device-built database: the raw C-ABI calls
                cc.query_set(idx, minus)
                cc.sync()
            else:
                check(lib().psi_pie_set_query(self._h, pi, pm))
        except PsiError as e:
            _raise_like_reference(e)
        self._uploaded = True

    def run(self):
        self._upload()
        if self._h is None:
            self.cryptoContext.run()
        else:
            check(lib().psi_pie_run(self._h))
        self._resultList = None

    def getResultList(self):
        if self._resultList is None:
            cc = self.cryptoContext
            out = np.empty((self.b, 2, cc.L, cc.N), dtype=np.uint64)
            if self._h is None:
                cc.result_get(out)
            else:
                check(lib().psi_pie_get_result_list(self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))))
            self._resultList = out
        return self._resultList

    def slots(self):
        """(slots [K][b][E][batch], mask slots [b][batch]) the constructor encoded; needs keepSlots."""
        s = np.empty((self.K, self.b, self.E, self.batchSize), dtype=np.int64)
        m = np.empty((self.b, self.batchSize), dtype=np.int64)
        check(lib().psi_pie_get_slots(self._h, s.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                      m.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))))
        return s, m

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().psi_pie_destroy(self._h)
                self._h = None
        except Exception:
            pass


def eval_sum_indices(N, batch_size):
    """Automorphism indices of EvalSum(ct, batch_size) (EvalSum_2n): what EvalSumKeyGen must have produced."""
    buf = np.zeros(32, dtype=np.uint64)
    n = ctypes.c_uint32()
    check(lib().psi_nb_eval_sum_indices(N, batch_size, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), ctypes.byref(n)))
    return [int(v) for v in buf[:n.value]]


def rotation_index(N, i):
    """FindAutomorphismIndex2n(i): the automorphism index of EvalAtIndex(ct, i)."""
    g = ctypes.c_uint64()
    check(lib().psi_nb_rotation_index(N, i, ctypes.byref(g)))
    return g.value


class FHEHIPPIE:
    """The reference's non-batched operator (FHEHIPPIE.hpp:18-50): one PIE over one inner cuckoo table, entry points
    setIndex / run / getResultList.  PIEs live in a FHEHIPPIECollection, whose database is ONE device array and whose
    runAll() evaluates every PIE in lock step; run() on a single PIE evaluates just that one."""

    def __init__(self, collection, number):
        self._c, self._n = collection, number
        self.indexMatrix = None

    def setIndex(self, indexMatrix):
        """K ciphertexts [2][L][N] (one per cuckoo hash function, SimpleFHEPSIServer.cpp:162-176)."""
        self.indexMatrix = indexMatrix
        self._c._results[self._n] = None

    def run(self):
        self._c._run_range(self._n, self._n + 1)

    def getResultList(self):
        r = self._c._results[self._n]
        if r is None:
            raise ValueError("run() has not been called")
        return r


class FHEHIPPIECollection:
    """PIECollection.hpp FHEHIPPIECollection: addPIE(inner cuckoo table) ..., runAll().
    The per-PIE constructor work (FHEHIPPIE.cpp:9-59: checks, bin permutation, plainVec with the trailing 1, random
    non-zero masks, result permutation) happens in addPIE; the encoding happens on the device at the first run."""

    def __init__(self, cryptoContext, pK, seed=None):
        self.cryptoContext, self.pK = cryptoContext, pK
        self._rng = np.random.default_rng(seed)   # None: OS entropy, like the reference's std::random_device
        self.myPIEs = []
        self._slots, self._masks, self._perm = [], [], []
        self._results = []
        self._encoded = False

    def addPIE(self, table, stash_size=0):
        """table: the cells [K][b][E] of one inner CuckooHashTable (HierarchicalCuckooHashTable.cells()[i][j])."""
        table = np.asarray(table)
        K, b, E = table.shape
        if b != E:
            raise ValueError("Error, for FHE PIE the size of a cuckoo bin has to be equal than the number of bins per hash function.")
        if stash_size != 0:
            raise ValueError("Error, FHE PIE does not support a stash (yet).")
        if self._slots and self._slots[0].shape != (K, b, E + 1):
            raise ValueError("all PIEs of a collection must have the same table shape")
        t = int(self.cryptoContext.t)
        perm2 = self._rng.permutation(b)          # permVec2: hides the bin index (FHEHIPPIE.cpp:29)
        slots = np.zeros((K, b, E + 1), dtype=np.int64)
        slots[:, perm2, :E] = table.astype(np.int64)
        slots[:, :, E] = 1                        # exponent slot of the "minus client" element (:49)
        self._slots.append(slots)
        self._masks.append(self._rng.integers(1, t, size=(K, b), dtype=np.int64))   # without 0 (:54)
        self._perm.append(self._rng.permutation(K))                                  # permutationVector
        self._results.append(None)
        self._encoded = False
        pie = FHEHIPPIE(self, len(self.myPIEs))
        self.myPIEs.append(pie)
        return pie

    def _encode(self):
        if not self._encoded:
            self.cryptoContext.nb_db_encode_slots(np.stack(self._slots), np.stack(self._masks))
            self._encoded = True

    def _run_range(self, p0, p1):
        self._encode()
        cc = self.cryptoContext
        K = self._slots[0].shape[0]
        idx = np.empty((p1 - p0, K, 2, cc.L, cc.N), dtype=np.uint64)
        for p in range(p0, p1):
            im = self.myPIEs[p].indexMatrix
            if im is None:
                raise ValueError("setIndex must be called before run()")
            idx[p - p0] = np.asarray(im, dtype=np.uint64)
        try:
            out = cc.nb_run(idx, p0, p1)
        except PsiError as e:
            _raise_like_reference(e)
        for p in range(p0, p1):
            shuffled = np.empty_like(out[p - p0])
            shuffled[self._perm[p]] = out[p - p0]   # shuffledResultList[permutationVector[hf]] = result (:74)
            self._results[p] = shuffled

    def runAll(self):
        self._run_range(0, len(self.myPIEs))
