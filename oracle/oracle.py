"""
ORACLE (test infrastructure) — ctypes/numpy binding of oracle/libpsi_oracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.  See oracle/psi_oracle.c for the reference line map and the "parity unpinned"
statement.
"""
import ctypes
import os
import subprocess

import numpy as np

from .params_ref import PsiParams

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_u64p = ctypes.POINTER(ctypes.c_uint64)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    so = os.path.join(_HERE, "libpsi_oracle.so")
    src = os.path.join(_HERE, "psi_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libpsi_oracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        L.orc_create.restype = ctypes.c_void_p
        L.orc_create.argtypes = [ctypes.POINTER(PsiParams)]
        L.orc_destroy.argtypes = [ctypes.c_void_p]
        L.orc_ntt.argtypes = [ctypes.c_void_p, _u64p, ctypes.c_int, ctypes.c_int]
        L.orc_pack.argtypes = [ctypes.c_void_p, _i64p, ctypes.c_int, _u64p]
        L.orc_unpack.argtypes = [ctypes.c_void_p, _u64p, _i64p]
        L.orc_encode.argtypes = [ctypes.c_void_p, _i64p, ctypes.c_int, _u64p]
        L.orc_set_encode_lift.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.orc_set_packing_cofactor.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.orc_keygen.argtypes = [ctypes.c_void_p, ctypes.c_uint64, _u64p, _u64p, _u64p]
        L.orc_encrypt_sk.argtypes = [ctypes.c_void_p, _u64p, _i64p, ctypes.c_int, ctypes.c_uint64, _u64p]
        L.orc_decrypt.argtypes = [ctypes.c_void_p, _u64p, _u64p, ctypes.c_int, _i64p,
                                  ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)]
        L.orc_mac_bin.argtypes = [ctypes.c_void_p, ctypes.c_int, _u64p, _u64p, _u64p, _u64p]
        L.orc_mul_ctpt.argtypes = [ctypes.c_void_p, _u64p, _u64p, _u64p]
        L.orc_mul_ctct.argtypes = [ctypes.c_void_p, _u64p, _u64p, _u64p, _u64p, _u64p]
        L.orc_mul_core.argtypes = [ctypes.c_void_p, _u64p, _u64p, _u64p]
        L.orc_relin.argtypes = [ctypes.c_void_p, _u64p, _u64p, _u64p, _u64p]
        L.orc_relin_hybrid.argtypes = [ctypes.c_void_p, _u64p, _u64p, _u64p, _u64p]
        L.orc_keygen_hybrid.argtypes = [ctypes.c_void_p, ctypes.c_uint64, _u64p, _u64p, _u64p]
        L.orc_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _u64p, _u64p, _u64p,
                              _u64p, _u64p, _u64p, _u64p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.orc_run_cold.argtypes = L.orc_run.argtypes
        L.orc_max_threads.restype = ctypes.c_int
        # non-batched FHEHIPPIE
        L.orc_automorphism_eval.argtypes = [ctypes.c_void_p, _u64p, ctypes.c_uint64, _u64p]
        L.orc_automorphism_coeff.argtypes = [ctypes.c_void_p, _u64p, ctypes.c_uint64, ctypes.c_int, _u64p]
        L.orc_find_automorphism_index.restype = ctypes.c_uint64
        L.orc_find_automorphism_index.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        L.orc_eval_sum_indices.argtypes = [ctypes.c_void_p, ctypes.c_int, _u64p]
        L.orc_auto_keygen.argtypes = [ctypes.c_void_p, _u64p, ctypes.c_uint64, ctypes.c_uint64, _u64p, _u64p]
        L.orc_auto_keygen_hybrid.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, _u64p, _u64p]
        L.orc_eval_automorphism.argtypes = [ctypes.c_void_p, _u64p, ctypes.c_uint64, _u64p, _u64p, _u64p]
        L.orc_nb_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _u64p, _u64p, _u64p, _u64p, ctypes.c_int,
                                 _u64p, _u64p, _u64p, _u64p]
        _LIB = L
    return _LIB


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


def _pi(a):
    assert a.dtype == np.int64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_i64p)


class Oracle:
    """CPU restatement of the BFV-RNS context + the BatchedFHEPIE server evaluation."""

    def __init__(self, params_struct):
        if not isinstance(params_struct, PsiParams):  # e.g. the product's own ctypes mirror
            assert ctypes.sizeof(params_struct) == ctypes.sizeof(PsiParams)
            params_struct = PsiParams.from_buffer_copy(bytes(params_struct))
        self.params = params_struct
        self.N, self.L, self.Lp, self.t = params_struct.N, params_struct.L, params_struct.Lp, params_struct.t
        self._h = lib().orc_create(ctypes.byref(params_struct))
        assert self._h

    def __del__(self):
        try:
            if self._h:
                lib().orc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- transforms ---------------------------------------------------------------
    def ntt(self, poly, mod_index, inverse=False):
        out = np.ascontiguousarray(poly, dtype=np.uint64).copy()
        lib().orc_ntt(self._h, _p(out), mod_index, 1 if inverse else 0)
        return out

    def pack(self, slots):
        slots = np.ascontiguousarray(slots, dtype=np.int64)
        out = np.empty(self.N, dtype=np.uint64)
        rc = lib().orc_pack(self._h, _pi(slots), len(slots), _p(out))
        if rc:
            raise ValueError("slot value out of range of the plaintext modulus")
        return out

    def unpack(self, coeff):
        out = np.empty(self.N, dtype=np.int64)
        lib().orc_unpack(self._h, _p(np.ascontiguousarray(coeff, dtype=np.uint64)), _pi(out))
        return out

    def set_encode_lift(self, mode):
        """0 = plain (default), 1 = centred; see orc_encode."""
        lib().orc_set_encode_lift(self._h, mode)

    def set_packing_cofactor(self, mode):
        """0 = co-factor 3 (default), 1 = 2N - 1; see orc_set_packing_cofactor."""
        lib().orc_set_packing_cofactor(self._h, mode)

    def encode(self, slots):
        """MakePackedPlaintext + SetFormat(EVALUATION): [L][N]."""
        slots = np.ascontiguousarray(slots, dtype=np.int64)
        out = np.empty((self.L, self.N), dtype=np.uint64)
        rc = lib().orc_encode(self._h, _pi(slots), len(slots), _p(out))
        if rc:
            raise ValueError("slot value out of range of the plaintext modulus")
        return out

    # -- client side --------------------------------------------------------------
    @property
    def hybrid(self):
        return self.params.ks_technique == 1

    def evk_shape(self):
        """BV: [L][L][N] (digit, limb); HYBRID: [numPartQ][L + Lk][N] (digit, limb of Q then of the special basis)."""
        if self.hybrid:
            return (self.params.ks_num_parts, self.L + self.params.Lk, self.N)
        return (self.L, self.L, self.N)

    def keygen(self, seed):
        sk = np.empty((self.L, self.N), dtype=np.uint64)
        evk_b = np.empty(self.evk_shape(), dtype=np.uint64)
        evk_a = np.empty(self.evk_shape(), dtype=np.uint64)
        (lib().orc_keygen_hybrid if self.hybrid else lib().orc_keygen)(self._h, seed, _p(sk), _p(evk_b), _p(evk_a))
        return sk, evk_b, evk_a

    def encrypt(self, sk, slots, seed):
        slots = np.ascontiguousarray(slots, dtype=np.int64)
        ct = np.empty((2, self.L, self.N), dtype=np.uint64)
        rc = lib().orc_encrypt_sk(self._h, _p(sk), _pi(slots), len(slots), seed, _p(ct))
        if rc:
            raise ValueError("slot value out of range of the plaintext modulus")
        return ct

    def decrypt(self, sk, ct):
        """-> (slots[N] centred int64, ambiguous roundings, noise budget bits)."""
        ct = np.ascontiguousarray(ct, dtype=np.uint64)
        out = np.empty(self.N, dtype=np.int64)
        amb = ctypes.c_int(0)
        nb = ctypes.c_double(0)
        lib().orc_decrypt(self._h, _p(sk), _p(ct), ct.shape[0], _pi(out), ctypes.byref(amb), ctypes.byref(nb))
        return out, amb.value, nb.value

    # -- server side --------------------------------------------------------------
    def mac_bin(self, idx, pt, minus):
        """idx [E][2][L][N], pt [E][L][N], minus [2][L][N] -> [2][L][N]."""
        E = idx.shape[0]
        out = np.empty((2, self.L, self.N), dtype=np.uint64)
        lib().orc_mac_bin(self._h, E, _p(np.ascontiguousarray(idx)), _p(np.ascontiguousarray(pt)),
                          _p(np.ascontiguousarray(minus)), _p(out))
        return out

    def mul_ctpt(self, ct, pt):
        out = np.empty((2, self.L, self.N), dtype=np.uint64)
        lib().orc_mul_ctpt(self._h, _p(np.ascontiguousarray(ct)), _p(np.ascontiguousarray(pt)), _p(out))
        return out

    def mul_ctct(self, ct1, ct2, evk_b, evk_a):
        out = np.empty((2, self.L, self.N), dtype=np.uint64)
        lib().orc_mul_ctct(self._h, _p(np.ascontiguousarray(ct1)), _p(np.ascontiguousarray(ct2)),
                           _p(evk_b), _p(evk_a), _p(out))
        return out

    def mul_core(self, ct1, ct2):
        """Tensor + scale-and-round only: [3][L][N], COEFFICIENT format, basis Q."""
        out = np.empty((3, self.L, self.N), dtype=np.uint64)
        lib().orc_mul_core(self._h, _p(np.ascontiguousarray(ct1)), _p(np.ascontiguousarray(ct2)), _p(out))
        return out

    def relin(self, res, evk_b, evk_a):
        out = np.empty((2, self.L, self.N), dtype=np.uint64)
        assert evk_b.shape == self.evk_shape() and evk_a.shape == self.evk_shape()
        (lib().orc_relin_hybrid if self.hybrid else lib().orc_relin)(self._h, _p(np.ascontiguousarray(res)), _p(evk_b),
                                                                    _p(evk_a), _p(out))
        return out

    def run(self, pt, mask, idx, minus, evk_b, evk_a, bin_begin=0, bin_end=None, nthreads=1, out=None):
        """BatchedFHEHIPPIE::run.  pt [K][b][E][L][N], mask [b][L][N], idx [K][E][2][L][N],
        minus [2][L][N] -> [b][2][L][N] (only bins bin_begin..bin_end-1 are written)."""
        K, b, E = pt.shape[0], pt.shape[1], pt.shape[2]
        if bin_end is None:
            bin_end = b
        if out is None:
            out = np.zeros((b, 2, self.L, self.N), dtype=np.uint64)
        lib().orc_run(self._h, K, b, E, _p(pt), _p(mask), _p(idx), _p(minus), _p(evk_b), _p(evk_a), _p(out),
                      bin_begin, bin_end, nthreads)
        return out

    # -- non-batched FHEHIPPIE (FHEHIPPIE.cpp:61-77) -------------------------------
    def automorphism_eval(self, poly, g):
        """AutomorphismTransform(g) of an EVALUATION polynomial [L][N]."""
        poly = np.ascontiguousarray(poly, dtype=np.uint64)
        out = np.empty_like(poly)
        lib().orc_automorphism_eval(self._h, _p(poly), g, _p(out))
        return out

    def automorphism_coeff(self, limb, g, mod_index):
        """a(X) -> a(X^g) on one COEFFICIENT limb [N] (the definition the EVALUATION permutation is checked against)."""
        limb = np.ascontiguousarray(limb, dtype=np.uint64)
        out = np.empty_like(limb)
        lib().orc_automorphism_coeff(self._h, _p(limb), g, mod_index, _p(out))
        return out

    def find_automorphism_index(self, i):
        return int(lib().orc_find_automorphism_index(self._h, i))

    def eval_sum_indices(self, batch_size):
        buf = np.zeros(32, dtype=np.uint64)
        n = lib().orc_eval_sum_indices(self._h, batch_size, _p(buf))
        return [int(v) for v in buf[:n]]

    def auto_keygen(self, sk, seed, indices, key_seed=None):
        """EvalSumKeyGen / EvalRotateKeyGen: one key per automorphism index -> (key_b, key_a), [n] + evk_shape().
        HYBRID contexts need key_seed (the seed keygen() was called with): the secret is re-derived over the special primes."""
        n = len(indices)
        kb = np.empty((n,) + self.evk_shape(), dtype=np.uint64)
        ka = np.empty_like(kb)
        for i, g in enumerate(indices):
            if self.hybrid:
                assert key_seed is not None
                lib().orc_auto_keygen_hybrid(self._h, key_seed, seed, int(g), _p(kb[i]), _p(ka[i]))
            else:
                lib().orc_auto_keygen(self._h, _p(sk), seed, int(g), _p(kb[i]), _p(ka[i]))
        return kb, ka

    def eval_automorphism(self, ct, g, key_b, key_a):
        ct = np.ascontiguousarray(ct, dtype=np.uint64)
        out = np.empty_like(ct)
        lib().orc_eval_automorphism(self._h, _p(ct), int(g), _p(key_b), _p(key_a), _p(out))
        return out

    def nb_run(self, idx, pt, merge_pt, mask, key_index, key_b, key_a):
        """FHEHIPPIE::run for one PIE.  idx [K][2][L][N], pt [K][b][L][N], merge_pt [L][N], mask [K][L][N],
        key_index [n] (automorphism indices), key_b / key_a [n][L][L][N] -> [K][2][L][N] (unpermuted)."""
        K, b = pt.shape[0], pt.shape[1]
        out = np.zeros((K, 2, self.L, self.N), dtype=np.uint64)
        ki = np.ascontiguousarray(key_index, dtype=np.uint64)
        rc = lib().orc_nb_run(self._h, K, b, _p(np.ascontiguousarray(idx)), _p(np.ascontiguousarray(pt)),
                              _p(np.ascontiguousarray(merge_pt)), _p(np.ascontiguousarray(mask)), len(ki), _p(ki),
                              _p(np.ascontiguousarray(key_b)), _p(np.ascontiguousarray(key_a)), _p(out))
        if rc:
            raise KeyError("automorphism key missing")
        return out


def run_cold(o, pt_coeff, mask_coeff, idx, minus, evk_b, evk_a, nthreads=1, out=None):
    """orc_run_cold: pt_coeff [K][b][E][N], mask_coeff [b][N] packed coefficients mod t (Oracle.pack output); the
    plaintext lifts and transforms happen inside the call, as in the reference's first run()."""
    K, b, E = pt_coeff.shape[:3]
    if out is None:
        out = np.zeros((b, 2, o.L, o.N), dtype=np.uint64)
    lib().orc_run_cold(o._h, K, b, E, _p(pt_coeff), _p(mask_coeff), _p(idx), _p(minus), _p(evk_b), _p(evk_a), _p(out), 0, b,
                       nthreads)
    return out


def max_threads():
    return lib().orc_max_threads()
