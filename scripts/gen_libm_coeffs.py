"""
Derives the polynomials of the branch-free double-precision sqrt / log / cos /
acos routines in the kernel prelude (myokit_b200/kernelgen.py): Chebyshev
interpolation in 60-digit arithmetic of the *kernel* function on its reduced
interval, coefficients rounded to double, and the approximation error that
is left (before rounding errors of the evaluation).

    python scripts/gen_libm_coeffs.py
"""
import mpmath as mp

mp.mp.dps = 60


def cheb_fit(f, a, b, deg):
    """Interpolates f on [a, b] at deg + 1 Chebyshev nodes; power-basis coefficients."""
    n = deg + 1
    nodes = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n))
             for k in range(n)]
    M = mp.matrix(n, n)
    y = mp.matrix(n, 1)
    for i, x in enumerate(nodes):
        for j in range(n):
            M[i, j] = x ** j
        y[i] = f(x)
    c = mp.lu_solve(M, y)
    return [mp.mpf(float(c[j])) for j in range(n)]      # rounded to double


def horner(c, x):
    p = mp.mpf(0)
    for k in reversed(c):
        p = p * x + k
    return p


def max_err(approx, exact, a, b, n=4000):
    worst = 0
    for i in range(1, n):
        x = a + (b - a) * mp.mpf(i) / n
        e = exact(x)
        worst = max(worst, abs((approx(x) - e) / e))
    return worst


def show(name, c):
    print('%s = {' % name)
    for k in c:
        print("    '%s'," % float(k).hex())
    print('}')


def main():
    # asin(s) = s + s z R(z), z = s^2 in [0, 1/4]
    def R(z):
        if z == 0:
            return mp.mpf(1) / 6
        s = mp.sqrt(z)
        return (mp.asin(s) / s - 1) / z
    for deg in (10, 11, 12):
        c = cheb_fit(R, mp.mpf(0), mp.mpf('0.2501'), deg)
        err = max_err(lambda s: s + s * s * s * horner(c, s * s), mp.asin,
                      mp.mpf('1e-4'), mp.mpf('0.5'))
        print('asin kernel degree %d: relative error %.3g (2^%.1f)'
              % (deg, float(err), float(mp.log(err, 2))))
    c = cheb_fit(R, mp.mpf(0), mp.mpf('0.2501'), 12)
    show('ASIN_R', c)

    # log(m) = 2 s + 2 s z A(z), s = (m - 1) / (m + 1), m in [sqrt(1/2), sqrt(2)]
    smax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)

    def A(z):
        if z == 0:
            return mp.mpf(1) / 3
        s = mp.sqrt(z)
        return (mp.atanh(s) / s - 1) / z
    for deg in (5, 6, 7):
        c = cheb_fit(A, mp.mpf(0), smax * smax * mp.mpf('1.0001'), deg)
        err = max_err(lambda s: 2 * s + 2 * s * s * s * horner(c, s * s),
                      lambda s: 2 * mp.atanh(s), mp.mpf('1e-5'), smax)
        print('log kernel degree %d: relative error %.3g (2^%.1f)'
              % (deg, float(err), float(mp.log(err, 2))))
    c = cheb_fit(A, mp.mpf(0), smax * smax * mp.mpf('1.0001'), 6)
    show('LOG_A', c)

    # sin(r) = r + r z S(z), cos(r) = 1 - z / 2 + z^2 C(z), z = r^2, |r| <= pi / 4
    zmax = (mp.pi / 4) ** 2 * mp.mpf('1.001')

    def S(z):
        if z == 0:
            return -mp.mpf(1) / 6
        r = mp.sqrt(z)
        return (mp.sin(r) / r - 1) / z

    def C(z):
        if z == 0:
            return mp.mpf(1) / 24
        r = mp.sqrt(z)
        return (mp.cos(r) - 1 + z / 2) / (z * z)
    for deg in (5, 6):
        cs = cheb_fit(S, mp.mpf(0), zmax, deg)
        cc = cheb_fit(C, mp.mpf(0), zmax, deg)
        es = max_err(lambda r: r + r ** 3 * horner(cs, r * r), mp.sin,
                     mp.mpf('1e-4'), mp.pi / 4)
        ec = max_err(lambda r: 1 - r * r / 2 + r ** 4 * horner(cc, r * r), mp.cos,
                     mp.mpf('1e-4'), mp.pi / 4)
        print('sin / cos kernels degree %d: relative error %.3g / %.3g (2^%.1f / 2^%.1f)'
              % (deg, float(es), float(ec), float(mp.log(es, 2)), float(mp.log(ec, 2))))
    show('SIN_S', cheb_fit(S, mp.mpf(0), zmax, 5))
    show('COS_C', cheb_fit(C, mp.mpf(0), zmax, 5))
    # constants: pi / 2 and pi as sums of doubles, ln 2 with a 32-bit head
    pio2 = mp.pi / 2
    hi = mp.mpf(float(pio2))
    mid = mp.mpf(float(pio2 - hi))
    lo = mp.mpf(float(pio2 - hi - mid))
    print('PIO2 hi / mid / lo:', float(hi).hex(), float(mid).hex(), float(lo).hex())
    phi = mp.mpf(float(mp.pi))
    print('PI hi / lo:', float(phi).hex(), float(mp.pi - phi).hex())
    print('2 / pi:', float(2 / mp.pi).hex())
    ln2_hi = mp.mpf(float.fromhex('0x1.62e42fee00000p-1'))  # e * hi is exact
    print('LN2 hi / lo:', float(ln2_hi).hex(), float(mp.log(2) - ln2_hi).hex())

if __name__ == '__main__':
    main()
