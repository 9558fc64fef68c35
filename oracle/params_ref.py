"""
ORACLE (test infrastructure, not product code) — exact big-integer restatement of the
BFV-RNS context OpenFHE builds for the reference's BatchedFHEPIE path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product has its own, independent C++ generator
(nested-hashing-psi_b200/host/params.cpp); tests/test_params.py checks the two agree
table by table.

PARITY UNPINNED: the arithmetic of this path lives in OpenFHE (openfheorg/openfhe-development),
which the reference neither vendors nor pins (/root/reference/CMakeLists.txt:10
`FIND_PACKAGE(OpenFHE REQUIRED)`; only hint: commented `-DOPENFHE_VERSION=0.9.2`, :15) and which
is not installed here.  Everything below is restated from the published algorithms
(Halevi-Polyakov-Shoup 2018 "An Improved RNS Variant of the BFV HE Scheme"; Kim-Polyakov-Zucca
2021 "Revisiting HE Schemes for Finite Fields" (HPSPOVERQ); Longa-Naehrig NTT) and from the
OpenFHE 1.0.x source layout as recalled; each recalled detail is marked (recalled).

Reference call sites that fix the parameters:
  /root/reference/src/Client/FHE/BatchedFHEPSIClient.cpp:22-38   plaintext modulus by bit size
  /root/reference/src/Client/FHE/BatchedFHEPSIClient.cpp:46-57   multiplicative depth by E
  /root/reference/src/Client/FHE/BatchedFHEPSIClient.cpp:72-78   ring dim 16384, 128-bit classic
  /root/reference/tests/TestBatchedFHEPIE.cpp:14-26              test context (depth 2, ring dim by library)
Library defaults used (recalled): scalingModSize 60, HPSPOVERQ, BV with digit size 0,
uniform ternary secret, sigma 3.19, assurance alpha 36.
"""
import ctypes
import math
from functools import reduce

MAX_LIMBS = 8

PLAINTEXT_MODULUS = {16: 65537, 32: 4296540161, 40: 1099579260929, 48: 281474981953537}


def depth_for_E(E):
    """BatchedFHEPSIClient.cpp:46-57."""
    if E < 500:
        return 3
    if E < 5000:
        return 5
    return 10


# ----------------------------------------------------------------------------- number theory
_MR_BASES = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)


def is_prime(n):
    if n < 2:
        return False
    for b in _MR_BASES:
        if n % b == 0:
            return n == b
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in _MR_BASES:
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def first_prime(nbits, m):
    """OpenFHE FirstPrime (recalled): first prime >= 2^nbits congruent to 1 mod m."""
    x = 1 << nbits
    r = x % m
    q = x + (m - r) % m + 1 if r else x + 1
    while not is_prime(q):
        q += m
    return q


def previous_prime(q, m):
    """OpenFHE PreviousPrime (recalled): next smaller prime congruent to 1 mod m."""
    q -= m
    while not is_prime(q):
        q -= m
    return q


def root_of_unity_min(m, q):
    """OpenFHE RootOfUnity (recalled): the MINIMUM primitive m-th root of unity mod q
    (the library cycles over all primitive roots so that contexts are reproducible)."""
    assert (q - 1) % m == 0 and m & (m - 1) == 0
    x = 2
    while True:
        r = pow(x, (q - 1) // m, q)
        if pow(r, m // 2, q) == q - 1:
            break
        x += 1
    r2 = r * r % q
    best, cur = r, r
    for _ in range(m // 2 - 1):
        cur = cur * r2 % q
        if cur < best:
            best = cur
    return best


# ----------------------------------------------------------------------------- sizeQ
_STD_TERNARY_128_CLASSIC = {1024: 27, 2048: 54, 4096: 109, 8192: 218, 16384: 438, 32768: 881}


def _find_ring_dim(logq_bits):
    for n in sorted(_STD_TERNARY_128_CLASSIC):
        if _STD_TERNARY_128_CLASSIC[n] >= logq_bits:
            return n
    return 65536


def bfv_logq(n, t, depth, dcrt_bits=60):
    """Natural-log size of Q from OpenFHE's BFVrns noise estimate, EvalMult-only branch
    (evalAddCount = keySwitchCount = 0), BV key switching with digit size 0 (recalled from
    ParameterGenerationBFVRNS::ParamsGenBFVRNS)."""
    sigma, alpha = 3.19, 36.0
    p = float(t)
    Berr = sigma * math.sqrt(alpha)
    Bkey = 1.0
    delta = 2.0 * math.sqrt(n)
    Vnorm = Berr * (1.0 + 2.0 * delta * Bkey)
    w = 2.0 ** dcrt_bits
    C1 = delta * delta * p * Bkey

    def noise_ks(logq_prev):
        return delta * (math.floor(logq_prev / (math.log(2) * dcrt_bits)) + 1) * w * Berr

    def logq_bfv(logq_prev):
        C2 = delta * delta * Bkey * Bkey / 2.0 + noise_ks(logq_prev)
        return math.log(4 * p) + (depth - 1) * math.log(C1) + math.log(C1 * Vnorm + depth * C2)

    logq_prev = 6.0 * math.log(10)
    logq = logq_bfv(logq_prev)
    logq = logq_bfv(logq)
    return logq


def size_q(n, t, depth, dcrt_bits=60):
    logq = bfv_logq(n, t, depth, dcrt_bits)
    return int(math.ceil((math.ceil(logq / math.log(2)) + 1.0) / dcrt_bits))


def choose_ring_dim(t, depth, dcrt_bits=60):
    """Ring dimension the library picks when the caller does not fix it
    (TestBatchedFHEPIE.cpp:14-26): smallest n whose 128-bit-classic bound admits k*dcrtBits."""
    n = 1024
    while True:
        k = size_q(n, t, depth, dcrt_bits)
        if _find_ring_dim(k * dcrt_bits) <= n:
            return n
        n *= 2


# ----------------------------------------------------------------------------- tables
def _prod(xs):
    return reduce(lambda a, b: a * b, xs, 1)


class RefParams:
    """Plain-Python container of every table psi_params carries, computed straight from the
    definitions with unbounded integers."""

    def __init__(self, N, t, depth=None, L=None, dcrt_bits=60, Lp=None, mult_technique=1, ks_technique=0, fp_contract=0):
        assert N & (N - 1) == 0
        assert (t - 1) % (2 * N) == 0, "packed encoding needs t = 1 mod 2N"
        if L is None:
            L = size_q(N, t, depth, dcrt_bits)
        self.N, self.t, self.L = N, t, L
        # HPSPOVERQ: sizeP = sizeQ; HPS: sizeP = sizeQ + 1 (the tensor of two centred lifts needs QP > N Q^2 / 2) (recalled)
        self.Lp = (L + 1 if mult_technique == 0 else L) if Lp is None else Lp
        m = 2 * N
        q = [previous_prime(first_prime(dcrt_bits, m), m)]
        for _ in range(1, L):
            q.append(previous_prime(q[-1], m))
        p = [previous_prime(q[-1], m)]             # aux basis continues below Q (recalled)
        for _ in range(1, self.Lp):
            p.append(previous_prime(p[-1], m))
        self.q, self.p = q, p
        self.psi_q = [root_of_unity_min(m, x) for x in q]
        self.psi_p = [root_of_unity_min(m, x) for x in p]
        self.psi_t = root_of_unity_min(m, t)
        Q, P = _prod(q), _prod(p)
        S = Q * P
        self.Q, self.P = Q, P
        Lq, Lp_ = L, self.Lp
        self.QHatInvModq = [pow(Q // q[i], -1, q[i]) for i in range(Lq)]
        self.QHatModp = [[(Q // q[i]) % p[j] for i in range(Lq)] for j in range(Lp_)]
        self.alphaQModp = [[(a * Q) % p[j] for j in range(Lp_)] for a in range(Lq + 1)]
        self.qInv = [1.0 / float(q[i]) for i in range(Lq)]
        self.negPQHatInvModq = [(-P * pow(Q // q[i], -1, q[i])) % q[i] for i in range(Lq)]
        self.qInvModp = [[pow(q[i], -1, p[j]) for j in range(Lp_)] for i in range(Lq)]
        self.PHatInvModp = [pow(P // p[j], -1, p[j]) for j in range(Lp_)]
        self.PHatModq = [[(P // p[j]) % q[i] for j in range(Lp_)] for i in range(Lq)]
        self.alphaPModq = [[(a * P) % q[i] for i in range(Lq)] for a in range(Lp_ + 1)]
        self.pInv = [1.0 / float(p[j]) for j in range(Lp_)]
        # ScaleAndRound by t/P, output Q:  S = Q*P, inputs = P limbs then the own Q limb
        self.tQSHatInvModsDivsModq = [[0] * (Lp_ + 1) for _ in range(Lq)]
        self.tQSHatInvModsDivsFrac = []
        for i in range(Lp_):
            c = t * Q * pow(S // p[i], -1, p[i])
            quo, rem = divmod(c, p[i])
            for j in range(Lq):
                self.tQSHatInvModsDivsModq[j][i] = quo % q[j]
            self.tQSHatInvModsDivsFrac.append(float(rem) / float(p[i]))
        for j in range(Lq):
            self.tQSHatInvModsDivsModq[j][Lp_] = (t * (Q // q[j]) * pow(S // q[j], -1, q[j])) % q[j]
        # HPS: ScaleAndRound by t/Q, output P: inputs = Q limbs then the own P limb
        self.tPSHatInvModsDivsModp = [[0] * (Lq + 1) for _ in range(Lp_)]
        self.tPSHatInvModsDivsFrac = []
        for i in range(Lq):
            c = t * P * pow(S // q[i], -1, q[i])
            quo, rem = divmod(c, q[i])
            for j in range(Lp_):
                self.tPSHatInvModsDivsModp[j][i] = quo % p[j]
            self.tPSHatInvModsDivsFrac.append(float(rem) / float(q[i]))
        for j in range(Lp_):
            self.tPSHatInvModsDivsModp[j][Lq] = (t * (P // p[j]) * pow(S // p[j], -1, p[j])) % p[j]
        self.mult_technique, self.ks_technique, self.fp_contract = mult_technique, ks_technique, fp_contract
        # HYBRID key switching (recalled): numPartQ = 3 for depth > 3, 2 for depth > 0, else 1 (ComputeNumLargeDigits),
        # never more than sizeQ; sizeP of the key-switching basis = limbs of the largest digit; special primes = the
        # next primes below the other bases
        self.ks_num_parts, self.Lk, self.pk, self.psi_pk = 0, 0, [], []
        if ks_technique == 1:
            d = depth if depth is not None else 2
            self.ks_num_parts = min(L, 3 if d > 3 else (2 if d > 0 else 1))
            alpha = -(-L // self.ks_num_parts)
            self.ks_num_parts = -(-L // alpha)          # digits actually populated
            self.Lk = alpha
            prev = p[-1]
            for _ in range(self.Lk):
                prev = previous_prime(prev, m)
                self.pk.append(prev)
            self.psi_pk = [root_of_unity_min(m, x) for x in self.pk]

    def to_struct(self):
        s = PsiParams()
        s.N, s.L, s.Lp = self.N, self.L, self.Lp
        s.mult_technique, s.ks_technique, s.reserved = self.mult_technique, self.ks_technique, 0
        s.fp_contract, s.ks_num_parts, s.Lk, s.reserved2 = self.fp_contract, self.ks_num_parts, self.Lk, 0
        for i in range(self.Lk):
            s.pk[i], s.psi_pk[i] = self.pk[i], self.psi_pk[i]
        for j in range(self.Lp):
            for i in range(self.L + 1):
                s.tPSHatInvModsDivsModp[j][i] = self.tPSHatInvModsDivsModp[j][i]
        for i in range(self.L):
            s.tPSHatInvModsDivsFrac[i] = self.tPSHatInvModsDivsFrac[i]
        s.t, s.psi_t = self.t, self.psi_t
        for i in range(self.L):
            s.q[i], s.psi_q[i] = self.q[i], self.psi_q[i]
            s.QHatInvModq[i] = self.QHatInvModq[i]
            s.qInv[i] = self.qInv[i]
            s.negPQHatInvModq[i] = self.negPQHatInvModq[i]
            for j in range(self.Lp):
                s.qInvModp[i][j] = self.qInvModp[i][j]
                s.PHatModq[i][j] = self.PHatModq[i][j]
            for i2 in range(self.Lp + 1):
                s.tQSHatInvModsDivsModq[i][i2] = self.tQSHatInvModsDivsModq[i][i2]
        for j in range(self.Lp):
            s.p[j], s.psi_p[j] = self.p[j], self.psi_p[j]
            s.PHatInvModp[j] = self.PHatInvModp[j]
            s.pInv[j] = self.pInv[j]
            s.tQSHatInvModsDivsFrac[j] = self.tQSHatInvModsDivsFrac[j]
            for i in range(self.L):
                s.QHatModp[j][i] = self.QHatModp[j][i]
        for a in range(self.L + 1):
            for j in range(self.Lp):
                s.alphaQModp[a][j] = self.alphaQModp[a][j]
        for a in range(self.Lp + 1):
            for i in range(self.L):
                s.alphaPModq[a][i] = self.alphaPModq[a][i]
        return s


_U64x8 = ctypes.c_uint64 * MAX_LIMBS
_U64x9 = ctypes.c_uint64 * (MAX_LIMBS + 1)
_F64x8 = ctypes.c_double * MAX_LIMBS


class PsiParams(ctypes.Structure):
    """ctypes mirror of `struct psi_params` (include/psi_b200.h)."""
    _fields_ = [
        ("N", ctypes.c_uint32), ("L", ctypes.c_uint32), ("Lp", ctypes.c_uint32),
        ("mult_technique", ctypes.c_uint32), ("ks_technique", ctypes.c_uint32),
        ("reserved", ctypes.c_uint32),
        ("t", ctypes.c_uint64),
        ("q", _U64x8), ("p", _U64x8), ("psi_q", _U64x8), ("psi_p", _U64x8),
        ("psi_t", ctypes.c_uint64),
        ("QHatInvModq", _U64x8),
        ("QHatModp", _U64x8 * MAX_LIMBS),
        ("alphaQModp", _U64x8 * (MAX_LIMBS + 1)),
        ("qInv", _F64x8),
        ("negPQHatInvModq", _U64x8),
        ("qInvModp", _U64x8 * MAX_LIMBS),
        ("PHatInvModp", _U64x8),
        ("PHatModq", _U64x8 * MAX_LIMBS),
        ("alphaPModq", _U64x8 * (MAX_LIMBS + 1)),
        ("pInv", _F64x8),
        ("tQSHatInvModsDivsModq", _U64x9 * MAX_LIMBS),
        ("tQSHatInvModsDivsFrac", _F64x8),
        ("fp_contract", ctypes.c_uint32), ("ks_num_parts", ctypes.c_uint32), ("Lk", ctypes.c_uint32),
        ("reserved2", ctypes.c_uint32), ("pk", _U64x8), ("psi_pk", _U64x8),
        ("tPSHatInvModsDivsModp", _U64x9 * MAX_LIMBS), ("tPSHatInvModsDivsFrac", _F64x8),
    ]


def struct_to_dict(s):
    """Flatten a PsiParams into comparable python values (used by tests)."""
    L, Lp = s.L, s.Lp
    return {
        "N": s.N, "L": L, "Lp": Lp, "t": s.t, "psi_t": s.psi_t,
        "mult_technique": s.mult_technique, "ks_technique": s.ks_technique,
        "q": list(s.q[:L]), "p": list(s.p[:Lp]),
        "psi_q": list(s.psi_q[:L]), "psi_p": list(s.psi_p[:Lp]),
        "QHatInvModq": list(s.QHatInvModq[:L]),
        "QHatModp": [list(s.QHatModp[j][:L]) for j in range(Lp)],
        "alphaQModp": [list(s.alphaQModp[a][:Lp]) for a in range(L + 1)],
        "qInv": list(s.qInv[:L]),
        "negPQHatInvModq": list(s.negPQHatInvModq[:L]),
        "qInvModp": [list(s.qInvModp[i][:Lp]) for i in range(L)],
        "PHatInvModp": list(s.PHatInvModp[:Lp]),
        "PHatModq": [list(s.PHatModq[i][:Lp]) for i in range(L)],
        "alphaPModq": [list(s.alphaPModq[a][:L]) for a in range(Lp + 1)],
        "pInv": list(s.pInv[:Lp]),
        "tQSHatInvModsDivsModq": [list(s.tQSHatInvModsDivsModq[j][:Lp + 1]) for j in range(L)],
        "tQSHatInvModsDivsFrac": list(s.tQSHatInvModsDivsFrac[:Lp]),
        "fp_contract": s.fp_contract, "ks_num_parts": s.ks_num_parts, "Lk": s.Lk,
        "pk": list(s.pk[:s.Lk]), "psi_pk": list(s.psi_pk[:s.Lk]),
        "tPSHatInvModsDivsModp": [list(s.tPSHatInvModsDivsModp[j][:L + 1]) for j in range(Lp)],
        "tPSHatInvModsDivsFrac": list(s.tPSHatInvModsDivsFrac[:L]),
    }
