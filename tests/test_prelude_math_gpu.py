"""
The in-line arithmetic of the generated kernels on the device: what
tests/test_prelude_math_host.py checks with a model of the reciprocal seed,
checked here with the real ``rcp.approx.ftz.f64`` (MUFU.RCP64H).
"""
import numpy as np
import pytest

from prelude_math import DIV_VARIANTS, EXP_VARIANTS, DeviceFunctions, ulp_error

pytestmark = pytest.mark.gpu


def random_pairs(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal(n) * np.exp2(rng.integers(-200, 200, n))
    b = rng.standard_normal(n) * np.exp2(rng.integers(-200, 200, n))
    b[b == 0] = 1.0
    return a, b


def test_reciprocal_seed_is_as_the_host_model_assumes():
    dev = DeviceFunctions()
    _, b = random_pairs(4_000_000, 3)
    r = dev.call('rcp_seed', b)
    rel = np.abs(r * b - 1.0)
    # 20 good bits (measured on a B200: max relative error 2^-19.95), low
    # word zero; 0 and inf map to inf and 0
    assert rel.max() < 2.0 ** -19.9, rel.max()
    assert np.all(r.view(np.uint64) & np.uint64(0xFFFFFFFF) == 0)
    s = dev.call('rcp_seed', np.array([0.0, -0.0, np.inf, -np.inf, 1e-310, 1.7e308]))
    assert np.isposinf(s[0]) and np.isneginf(s[1]) and s[2] == 0 and s[3] == 0
    assert np.isposinf(s[4]) and s[5] == 0


@pytest.mark.parametrize('div', sorted(DIV_VARIANTS))
def test_division_on_device(div):
    dev = DeviceFunctions(div)
    a, b = random_pairs(8_000_000, 17)
    q = dev.call('div', a, b)
    err = ulp_error(q, a.astype(np.longdouble) / b.astype(np.longdouble))
    assert err.max() <= (1.5 if div == 'cubic' else 1.0), err.max()
    if div != 'newton':
        inf, nan = np.inf, np.nan
        a = np.array([1.0, -2.0, 1.0, -3.0, 0.0, inf, -inf, 5.0, nan, inf, 0.0, 7.0])
        b = np.array([inf, inf, 0.0, 0.0, 0.0, 2.0, 4.0, nan, 1.0, inf, 3.0, -inf])
        q = dev.call('div', a, b)
        with np.errstate(all='ignore'):
            want = a / b
        assert np.array_equal(np.isnan(q), np.isnan(want)), (q, want)
        ok = ~np.isnan(want)
        assert np.array_equal(q[ok], want[ok]), (q, want)


@pytest.mark.parametrize('name', EXP_VARIANTS)
def test_exp_on_device(name):
    dev = DeviceFunctions()
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-700, 700, 1_000_000), rng.uniform(-40, 40, 1_000_000)])
    y = dev.call(name, x)
    # (glibc's exp is within 1 ulp: 2 ulp here allows for both)
    err = ulp_error(y, np.exp(x.astype(np.longdouble)))
    assert err.max() <= 1.6, err.max()
    y = dev.call(name, np.array([800.0, -800.0, 0.0]))
    if name in ('mkb_exp_poly', 'mkb_exp_estrin'):
        assert np.isposinf(y[0]) and y[1] == 0 and y[2] == 1.0
    else:
        assert np.isfinite(y[0]) and y[0] > 6e307 and 0 <= y[1] < 1e-300 and y[2] == 1.0


def test_branch_free_libm_on_device():
    # against numpy's long double routines (x87: 64-bit mantissa)
    dev = DeviceFunctions('cubic')
    rng = np.random.default_rng(21)
    n = 2_000_000
    L = np.longdouble
    wide = np.exp(rng.uniform(-700, 700, n))
    seed = dev.call('rsqrt_seed', wide)
    assert np.abs(seed * np.sqrt(wide) - 1).max() < 2.0 ** -19, 'rsqrt.approx seed'
    assert ulp_error(dev.call('mkb_sqrt', wide), np.sqrt(wide.astype(L))).max() <= 0.51
    assert ulp_error(dev.call('mkb_log', wide), np.log(wide.astype(L))).max() <= 1.0
    x = rng.uniform(0.4, 2.5, n)
    assert ulp_error(dev.call('mkb_log', x), np.log(x.astype(L))).max() <= 1.0
    x = rng.uniform(-1, 1, n)
    assert ulp_error(dev.call('mkb_acos', x), np.arccos(x.astype(L))).max() <= 1.6
    # (long double cos reduces with a 64-bit pi: compare where that is exact enough)
    x = rng.uniform(-20, 20, n)
    assert ulp_error(dev.call('mkb_cos', x), np.cos(x.astype(L))).max() <= 2.0
    a = np.exp(rng.uniform(-8, 8, n))
    b = rng.uniform(-6, 6, n)
    err = ulp_error(dev.call('pow', a, b), np.power(a.astype(L), b.astype(L)))
    # (the rounding of log and of the product both scale with the exponent argument)
    assert (err / (2.0 + 1.5 * np.abs(b * np.log(a)))).max() <= 1.0
    inf, nan = np.inf, np.nan
    y = dev.call('mkb_sqrt', np.array([0.0, inf, -1.0, nan, 4.0]))
    assert y[0] == 0 and y[1] == inf and np.isnan(y[2]) and np.isnan(y[3]) and y[4] == 2.0
    y = dev.call('mkb_log', np.array([0.0, inf, -1.0, nan, 1.0]))
    assert y[0] == -inf and y[1] == inf and np.isnan(y[2]) and np.isnan(y[3]) and y[4] == 0.0
    y = dev.call('mkb_acos', np.array([1.0, -1.0, 2.0]))
    assert y[0] == 0.0 and y[1] == np.pi and np.isnan(y[2])
