"""
Derives the polynomial used by mkb_exp (myokit_b200/kernelgen.py prelude):
exp(r) on |r| <= ln(2)/2 as a degree-11 polynomial, coefficients from Chebyshev
interpolation in 60-digit arithmetic, rounded to double. Also checks the whole
algorithm (range reduction + Horner with fused multiply-adds + scaling) against
mpmath on random arguments, emulating double rounding after every operation.

    python scripts/gen_exp_coeffs.py
"""
import random
import mpmath as mp

mp.mp.dps = 60
DEG = 11
A = mp.log(2) / 2 * mp.mpf('1.0001')


def cheb_coeffs(f, a, deg):
    n = deg + 1
    nodes = [a * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    # Solve Vandermonde in high precision
    M = mp.matrix(n, n)
    b = mp.matrix(n, 1)
    for i, x in enumerate(nodes):
        for j in range(n):
            M[i, j] = x ** j
        b[i] = f(x)
    c = mp.lu_solve(M, b)
    return [c[j] for j in range(n)]


def to_double(x):
    return float(mp.nstr(x, 25))


def rd(x):
    """Round an mpf to the nearest double (as mpf)."""
    return mp.mpf(float(x))


def fma(a, b, c):
    return rd(a * b + c)


L2E = rd(1 / mp.log(2))
LN2_HI = mp.mpf(float.fromhex('0x1.62e42fefa39efp-1'))
LN2_LO = rd(mp.log(2) - LN2_HI)


def mkb_exp(x, C):
    x = rd(x)
    t = rd(rd(x * L2E))
    n = mp.nint(t)
    r = fma(n, -LN2_HI, x)
    r = fma(n, -LN2_LO, r)
    p = C[DEG]
    for k in range(DEG - 1, -1, -1):
        p = fma(p, r, C[k])
    return rd(p * mp.mpf(2) ** int(n))


def mkb_exp_estrin(x, C):
    """The Estrin-scheme variant (option fast_exp='estrin'): same reduction and
    coefficients, the polynomial evaluated as a tree of depth 6 instead of a
    chain of 11 dependent fused multiply-adds (3 more FP64 instructions)."""
    x = rd(x)
    t = rd(rd(x * L2E))
    n = mp.nint(t)
    r = fma(n, -LN2_HI, x)
    r = fma(n, -LN2_LO, r)
    r2 = rd(r * r)
    a1 = fma(C[3], r, C[2])
    a2 = fma(C[5], r, C[4])
    a3 = fma(C[7], r, C[6])
    a4 = fma(C[9], r, C[8])
    a5 = fma(C[11], r, C[10])
    r4 = rd(r2 * r2)
    b0 = fma(a2, r2, a1)
    b1 = fma(a4, r2, a3)
    r8 = rd(r4 * r4)
    d = fma(b1, r4, b0)
    q = fma(a5, r8, d)
    p = fma(r2, q, r)
    p = rd(p + 1)
    return rd(p * mp.mpf(2) ** int(n))


def main():
    c = cheb_coeffs(mp.exp, A, DEG)
    C = [rd(x) for x in c]
    C[0] = mp.mpf(1)
    C[1] = mp.mpf(1)
    print('static const double coefficients c0..c%d:' % DEG)
    for k, x in enumerate(C):
        print('    %s,  // c%d' % (float(x).hex(), k))
        print('    //   = %r' % float(x))
    print('LN2_LO =', float(LN2_LO).hex(), repr(float(LN2_LO)))
    print('L2E    =', float(L2E).hex(), repr(float(L2E)))
    random.seed(1)
    worst = 0
    for i in range(20000):
        x = random.uniform(-700, 700) if i % 2 else random.uniform(-5, 5)
        got = mkb_exp(x, C)
        want = mp.exp(rd(x))
        ulp = mp.mpf(2) ** (mp.floor(mp.log(want, 2)) - 52)
        err = abs(got - want) / ulp
        worst = max(worst, err)
    print('max error over 20000 random arguments: %.3f ulp' % float(worst))
    random.seed(1)
    worst = 0
    for i in range(20000):
        x = random.uniform(-700, 700) if i % 2 else random.uniform(-5, 5)
        got = mkb_exp_estrin(x, C)
        want = mp.exp(rd(x))
        ulp = mp.mpf(2) ** (mp.floor(mp.log(want, 2)) - 52)
        worst = max(worst, abs(got - want) / ulp)
    print('Estrin variant, max error over 20000 random arguments: %.3f ulp'
          % float(worst))


if __name__ == '__main__':
    main()


# ---------------------------------------------------------------------------
# Table variant (mkb_exp in the shipped prelude): exp(x) = 2^m * T[j] * e^r,
# n = rint(x * 64 / ln2) = 64 m + j, r = x - n ln2 / 64, |r| <= ln2 / 128,
# e^r - 1 by a degree-5 polynomial: 10 FP64-pipe instructions instead of 17.
# ---------------------------------------------------------------------------
def table_variant():
    N = 64
    T = [rd(mp.mpf(2) ** (mp.mpf(j) / N)) for j in range(N)]
    K = rd(N / mp.log(2))
    C_HI = rd(mp.log(2) / N)
    C_LO = rd(mp.log(2) / N - C_HI)
    # q(r) ~ (e^r - 1 - r) / r^2 on |r| <= ln2/128, degree 3
    a = mp.log(2) / (2 * N) * mp.mpf('1.0001')

    def f(x):
        if x == 0:
            return mp.mpf(1) / 2
        return (mp.exp(x) - 1 - x) / (x * x)
    q = [rd(c) for c in cheb_coeffs(f, a, 3)]

    def mkb_exp_t(x):
        x = rd(x)
        n = mp.nint(rd(x * K))
        r = fma(n, -C_HI, x)
        r = fma(n, -C_LO, r)
        p = fma(q[3], r, q[2])
        p = fma(p, r, q[1])
        p = fma(p, r, q[0])
        r2 = rd(r * r)
        p = fma(r2, p, r)
        j = int(n) % N
        m = (int(n) - j) // N
        y = fma(T[j], p, T[j])
        return rd(y * mp.mpf(2) ** m)

    print('table variant: K =', float(K).hex(), 'C_HI =', float(C_HI).hex(),
          'C_LO =', float(C_LO).hex())
    print('q0..q3 =', [float(c).hex() for c in q])
    random.seed(2)
    worst = 0
    for i in range(20000):
        x = random.uniform(-700, 700) if i % 2 else random.uniform(-5, 5)
        got = mkb_exp_t(x)
        want = mp.exp(rd(x))
        ulp = mp.mpf(2) ** (mp.floor(mp.log(want, 2)) - 52)
        worst = max(worst, abs(got - want) / ulp)
    print('table variant, max error over 20000 random arguments: %.3f ulp'
          % float(worst))
    return T, K, C_HI, C_LO, q


if __name__ == '__main__':
    table_variant()
