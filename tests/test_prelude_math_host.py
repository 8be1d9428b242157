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
def test_exp_polynomial_forms_never_produce_nan(name):
    # saturating forms: huge / tiny finite values outside the double range
    lib = host_library()
    x = np.array([800.0, 710.0, 2000.0, -800.0, -2000.0])
    y = call_host(lib, name, x)
    assert np.all(np.isfinite(y)) and np.all(y > 0), y
    assert np.all(y[:3] > 6e307) and np.all(y[3:] < 1e-300)
    lib_c = host_library('cubic')
    q = call_host(lib_c, 'div', np.ones(5), 1.0 + y)
    assert np.all(q[:3] < 2e-308) and np.all(q[3:] == 1.0)
