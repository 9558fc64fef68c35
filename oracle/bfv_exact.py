"""
ORACLE (test infrastructure, not product code) — exact big-integer model of textbook BFV.

Only tests/ may import this.  It pins the RNS restatement in oracle/psi_oracle.c to the
mathematical definition of the scheme (Fan-Vercauteren 2012): ciphertext multiplication is
    round(t/Q * (c0 c0', c0 c1' + c1 c0', c1 c1'))  in  Z[x]/(x^N + 1),  reduced mod Q,
with unbounded Python integers and exact rational rounding, no RNS, no floating point.  The
HPS / HPSPOVERQ RNS procedures (oracle/psi_oracle.c: orc_mul_core) approximate exactly this value;
tests/test_oracle.py bounds the difference and checks that both decrypt to the slot-wise product.

Reference call site whose semantics this pins: EvalMult(ct,ct), BatchedFHEHIPPIE.cpp:123.
"""
import numpy as np


def crt_reconstruct(limbs, moduli):
    """limbs: [L][N] residues -> list of N python ints in [0, Q)."""
    Q = 1
    for q in moduli:
        Q *= q
    out = [0] * len(limbs[0])
    for l, q in enumerate(moduli):
        Qh = Q // q
        c = Qh * pow(Qh, -1, q)
        row = limbs[l]
        for j in range(len(out)):
            out[j] = (out[j] + int(row[j]) * c) % Q
    return out, Q


def centre(v, Q):
    return [x - Q if x > Q // 2 else x for x in v]


def negacyclic_mul(a, b):
    """Exact product in Z[x]/(x^N + 1); O(N^2), meant for N <= 256."""
    n = len(a)
    res = [0] * n
    for i, ai in enumerate(a):
        if ai == 0:
            continue
        for j, bj in enumerate(b):
            k = i + j
            if k < n:
                res[k] += ai * bj
            else:
                res[k - n] -= ai * bj
    return res


def round_div(num, den):
    """round(num/den), ties away from zero (den > 0)."""
    if num >= 0:
        return (2 * num + den) // (2 * den)
    return -((-2 * num + den) // (2 * den))


class ExactBFV:
    def __init__(self, oracle):
        self.o = oracle
        self.N, self.L, self.t = oracle.N, oracle.L, int(oracle.t)
        self.q = [int(oracle.params.q[i]) for i in range(self.L)]
        self.Q = 1
        for q in self.q:
            self.Q *= q

    def poly_to_int(self, poly_eval):
        """[L][N] EVALUATION limbs -> centred integer coefficients."""
        coeff = [self.o.ntt(poly_eval[l], l, inverse=True) for l in range(self.L)]
        v, Q = crt_reconstruct(coeff, self.q)
        return centre(v, Q)

    def int_to_limbs(self, v):
        """centred integer coefficients -> [L][N] COEFFICIENT limbs."""
        out = np.empty((self.L, self.N), dtype=np.uint64)
        for l, q in enumerate(self.q):
            out[l] = np.array([x % q for x in v], dtype=np.uint64)
        return out

    def mul(self, ct1, ct2):
        """Exact BFV tensor + scale: three centred integer polynomials (mod Q)."""
        a0, a1 = self.poly_to_int(ct1[0]), self.poly_to_int(ct1[1])
        b0, b1 = self.poly_to_int(ct2[0]), self.poly_to_int(ct2[1])
        t0 = negacyclic_mul(a0, b0)
        t1 = [x + y for x, y in zip(negacyclic_mul(a0, b1), negacyclic_mul(a1, b0))]
        t2 = negacyclic_mul(a1, b1)
        Q, t = self.Q, self.t
        return [centre([round_div(t * x, Q) % Q for x in tt], Q) for tt in (t0, t1, t2)]

    def decrypt_int(self, comps, sk_eval):
        """comps: list of centred integer polys (c0, c1[, c2]); returns t/Q-scaled message coeffs mod t
        and the largest rounding distance (noise indicator in [0, 0.5))."""
        s = self.poly_to_int(sk_eval)
        acc = list(comps[0])
        sp = s
        for c in comps[1:]:
            prod = negacyclic_mul(c, sp)
            acc = [x + y for x, y in zip(acc, prod)]
            sp = negacyclic_mul(sp, s)
        Q, t = self.Q, self.t
        acc = centre([x % Q for x in acc], Q)
        m, worst = [], 0.0
        for x in acc:
            r = round_div(t * x, Q)
            d = abs(t * x - r * Q) / Q
            worst = max(worst, d)
            m.append(r % t)
        return m, worst
