"""
The in-line arithmetic of the generated kernels (division, exp) on the host:
accuracy on random operands and IEEE behaviour for special operands. The
reciprocal seed is the shim's model of rcp.approx.ftz.f64 (20 mantissa bits,
low word zero); tests/test_prelude_math_gpu.py repeats this on the device.
"""
import numpy as np
import pytest

from prelude_math import (DIV_VARIANTS, EXP_VARIANTS, call_host, host_library,
                          ulp_error)


def random_pairs(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal(n) * np.exp2(rng.integers(-200, 200, n))
    b = rng.standard_normal(n) * np.exp2(rng.integers(-200, 200, n))
    b[b == 0] = 1.0
    return a, b


@pytest.mark.parametrize('div', sorted(DIV_VARIANTS))
def test_division_within_one_ulp(div):
    lib = host_library(div)
    a, b = random_pairs(2_000_000, 11)
    q = call_host(lib, 'div', a, b)
    exact = a.astype(np.longdouble) / b.astype(np.longdouble)
    err = ulp_error(q, exact)
    # (cubic: two roundings, reciprocal and product)
    assert err.max() <= (1.5 if div == 'cubic' else 1.0), err.max()
    # the refined forms are correctly rounded almost always
    wrong = np.mean(q != a / b)
    assert wrong < (0.35 if div == 'cubic' else 1e-4), wrong
    # reciprocals (the most common form in gating equations)
    one = np.ones_like(b)
    err = ulp_error(call_host(lib, 'div', one, b), 1 / b.astype(np.longdouble))
    assert err.max() <= 1.0


@pytest.mark.parametrize('div', ['parallel', 'cubic'])
def test_division_special_operands_follow_ieee(div):
    lib = host_library(div)
    inf, nan = np.inf, np.nan
    a = np.array([1.0, -2.0, 1.0, -3.0, 0.0, inf, -inf, 5.0, nan, inf, 0.0, 7.0])
    b = np.array([inf, inf, 0.0, 0.0, 0.0, 2.0, 4.0, nan, 1.0, inf, 3.0, -inf])
    q = call_host(lib, 'div', a, b)
    with np.errstate(all='ignore'):
        want = a / b
    assert np.array_equal(np.isnan(q), np.isnan(want)), (q, want)
    ok = ~np.isnan(want)
    assert np.array_equal(q[ok], want[ok]), (q, want)
    assert np.array_equal(np.signbit(q[ok]), np.signbit(want[ok]))


@pytest.mark.parametrize('name', EXP_VARIANTS)
def test_exp_accuracy(name):
    import mpmath as mp
    mp.mp.dps = 40
    lib = host_library()
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-700, 700, 3000), rng.uniform(-40, 40, 3000),
                        rng.uniform(-1, 1, 2000), [0.0, 1e-300, -1e-300, 708.0, -708.0]])
    y = call_host(lib, name, x)
    worst = 0.0
    for xi, yi in zip(x, y):
        e = mp.exp(mp.mpf(float(xi)))
        ulp = mp.mpf(float(np.spacing(float(e))))
        worst = max(worst, float(abs(mp.mpf(float(yi)) - e) / ulp))
    assert worst <= 1.1, worst


@pytest.mark.parametrize('name', ['mkb_exp_stab', 'mkb_exp_tab'])
def test_exp_table_forms_saturate(name):
    lib = host_library()
    x = np.array([800.0, 710.0, 1e4, -800.0, -745.2, -1e4, 709.0, -708.0, 0.0])
    y = call_host(lib, name, x)
    # overflow: just below DBL_MAX, never inf or NaN; underflow: exactly 0
    assert np.all(np.isfinite(y[:3])) and np.all(y[:3] > 1.7e308), y
    assert np.all(y[3:6] == 0.0), y
    assert np.isfinite(y[6]) and abs(y[6] / np.exp(709.0) - 1) < 1e-15
    assert abs(y[7] / np.exp(-708.0) - 1) < 1e-15 and y[8] == 1.0
    # and the gating-variable forms come out as IEEE would give them:
    # 1 / (1 + exp(big)) = 0, 1 / (1 + exp(-big)) = 1
    lib_c = host_library('cubic')
    big = call_host(lib_c, name, np.array([900.0, -900.0]))
    q = call_host(lib_c, 'div', np.ones(2), 1.0 + big)
    assert q[0] == 0.0 and q[1] == 1.0


@pytest.mark.parametrize('name', ['mkb_exp_poly', 'mkb_exp_estrin'])
def test_exp_polynomial_forms_overflow_and_underflow_like_ieee(name):
    # the power of two is applied with a multiplication: +inf / 0 outside the
    # double range (the advisor's case: 1 / (1 + exp(big)) must be 0)
    lib = host_library()
    x = np.array([800.0, 710.0, 2000.0, 1e9, -800.0, -2000.0, -1e9, 709.0, -708.0, 0.0])
    y = call_host(lib, name, x)
    assert np.all(np.isposinf(y[:4])), y
    assert np.all(y[4:7] == 0.0), y
    assert abs(y[7] / np.exp(709.0) - 1) < 1e-15 and abs(y[8] / np.exp(-708.0) - 1) < 1e-15
    assert y[9] == 1.0
    assert np.isnan(call_host(lib, name, np.array([np.nan])))[0]
    lib_c = host_library('cubic')
    q = call_host(lib_c, 'div', np.ones(7), 1.0 + y[:7])
    assert np.all(q[:4] == 0.0) and np.all(q[4:] == 1.0)


def _mp_ulp_error(lib, name, x, f):
    import mpmath as mp
    mp.mp.dps = 40
    y = call_host(lib, name, x)
    worst = 0.0
    for xi, yi in zip(x, y):
        want = f(mp.mpf(float(xi)))
        ulp = mp.mpf(float(np.spacing(abs(float(want)))))
        worst = max(worst, float(abs(mp.mpf(float(yi)) - want) / ulp))
    return worst


def test_branch_free_libm_accuracy():
    import mpmath as mp
    lib = host_library('cubic')
    rng = np.random.default_rng(8)
    n = 4000
    wide = np.exp(rng.uniform(-700, 700, n))
    assert _mp_ulp_error(lib, 'mkb_sqrt', wide, mp.sqrt) <= 0.501
    assert _mp_ulp_error(lib, 'mkb_log', wide, mp.log) <= 1.0
    assert _mp_ulp_error(lib, 'mkb_log', rng.uniform(0.4, 2.5, n), mp.log) <= 1.0
    assert _mp_ulp_error(lib, 'mkb_log', 1 + rng.uniform(-1e-3, 1e-3, n), mp.log) <= 1.0
    assert _mp_ulp_error(lib, 'mkb_cos', rng.uniform(-10, 10, n), mp.cos) <= 1.6
    assert _mp_ulp_error(lib, 'mkb_cos', rng.uniform(-1e5, 1e5, n), mp.cos) <= 1.6
    assert _mp_ulp_error(lib, 'mkb_cos', rng.uniform(-1e9, 1e9, n), mp.cos) <= 1.6
    assert _mp_ulp_error(lib, 'mkb_acos', rng.uniform(-1, 1, n), mp.acos) <= 1.5
    near = np.concatenate([1 - np.exp(rng.uniform(-30, 0, n)), np.exp(rng.uniform(-30, 0, n)) - 1])
    assert _mp_ulp_error(lib, 'mkb_acos', near, mp.acos) <= 1.5


def test_branch_free_libm_special_operands():
    lib = host_library('cubic')
    inf, nan = np.inf, np.nan
    x = np.array([0.0, -0.0, inf, -inf, nan, -1.0, 1.0, 4.0])
    y = call_host(lib, 'mkb_sqrt', x)
    assert y[0] == 0 and y[1] == 0 and np.signbit(y[1]) and y[2] == inf
    assert np.all(np.isnan(y[3:6])) and y[6] == 1.0 and y[7] == 2.0
    y = call_host(lib, 'mkb_log', x)
    assert y[0] == -inf and y[1] == -inf and y[2] == inf
    assert np.all(np.isnan(y[3:6])) and y[6] == 0.0
    y = call_host(lib, 'mkb_cos', x)
    assert y[0] == 1.0 and y[1] == 1.0 and np.all(np.isnan(y[2:5]))
    y = call_host(lib, 'mkb_acos', np.array([1.0, -1.0, 0.0, 2.0, -2.0, nan, 0.5, -0.5]))
    assert y[0] == 0.0 and y[1] == np.pi and y[2] == np.pi / 2
    assert np.all(np.isnan(y[3:6]))
    assert abs(y[6] - np.pi / 3) < 3e-16 and abs(y[7] - 2 * np.pi / 3) < 5e-16


def test_pow_through_exp_and_log():
    import mpmath as mp
    mp.mp.dps = 40
    lib = host_library('cubic')
    rng = np.random.default_rng(9)
    x = np.exp(rng.uniform(-8, 8, 3000))
    y = rng.uniform(-6, 6, 3000)
    got = call_host(lib, 'pow', x, y)
    worst = 0.0
    for xi, yi, gi in zip(x, y, got):
        want = mp.power(mp.mpf(float(xi)), mp.mpf(float(yi)))
        ulp = mp.mpf(float(np.spacing(float(want))))
        err = float(abs(mp.mpf(float(gi)) - want) / ulp)
        worst = max(worst, err / (2.0 + 1.5 * abs(yi * np.log(xi))))
    assert worst <= 1.0, worst
    # 0^y, overflow and underflow, NaN
    got = call_host(lib, 'pow', np.array([0.0, 0.0, 10.0, 10.0, np.nan, 2.0]),
                    np.array([1.4, -1.4, 400.0, -400.0, 1.0, 0.5]))
    assert got[0] == 0.0 and got[1] == np.inf and got[2] == np.inf and got[3] == 0.0
    assert np.isnan(got[4]) and abs(got[5] - np.sqrt(2.0)) < 5e-16
