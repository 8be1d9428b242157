"""
The arithmetic routines of the kernel prelude (``mkb_div``, ``mkb_exp_*``) as
stand-alone functions, for tests: built for the host behind the CUDA shim
(g++) and, on a GPU box, as a small cubin through the product's own JIT.
TEST INFRASTRUCTURE ONLY.
"""
import ctypes
import hashlib
import os
import subprocess

import numpy as np

from myokit_b200 import kernelgen

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, 'cuda_shim')
_BUILD = os.path.join(_SHIM, '_build')

DIV_VARIANTS = {
    'newton': dict(MKB_DIV_PARALLEL=0, MKB_DIV_CUBIC=0),
    'parallel': dict(MKB_DIV_PARALLEL=1, MKB_DIV_CUBIC=0),
    'cubic': dict(MKB_DIV_PARALLEL=0, MKB_DIV_CUBIC=1),
}
EXP_VARIANTS = ['mkb_exp_poly', 'mkb_exp_estrin', 'mkb_exp_tab', 'mkb_exp_stab']
UNARY = EXP_VARIANTS + [name for name in getattr(kernelgen, 'PRELUDE_UNARY', [])]


def _defines(div):
    d = dict(MKB_DIV_INT_CHECK=0)
    d.update(DIV_VARIANTS[div])
    return ''.join('#define %s %d\n' % kv for kv in sorted(d.items()))


def host_library(div='newton'):
    """ctypes library with f_div(a, b, out, n) and f_<unary>(x, out, n)."""
    code = ['#include "mkb_cuda_shim.h"', 'typedef double Real;', _defines(div),
            kernelgen._PRELUDE, 'extern "C" {',
            'void f_init() { MKB_EXP_TABLE_INIT(0u, 1u); }',
            'void f_div(const double* a, const double* b, double* out, long n) {'
            ' for (long i = 0; i < n; i++) out[i] = mkb_div(a[i], b[i]); }',
            'void f_pow(const double* a, const double* b, double* out, long n) {'
            ' for (long i = 0; i < n; i++) out[i] = mkb_pow(a[i], b[i]); }']
    for name in UNARY:
        code.append('void f_%s(const double* x, double* out, long n) {'
                    ' for (long i = 0; i < n; i++) out[i] = %s(x[i]); }' % (name, name))
    code.append('}')
    code = '\n'.join(code)
    with open(os.path.join(_SHIM, 'mkb_cuda_shim.h'), 'rb') as f:
        shim = f.read()
    key = hashlib.sha1(code.encode() + shim).hexdigest()[:20]
    os.makedirs(_BUILD, exist_ok=True)
    so = os.path.join(_BUILD, 'm_%s.so' % key)
    if not os.path.isfile(so):
        src = os.path.join(_BUILD, 'm_%s.cpp' % key)
        with open(src, 'w') as f:
            f.write(code)
        tmp = so + '.tmp%d' % os.getpid()
        r = subprocess.run(
            ['g++', '-O1', '-std=c++17', '-fPIC', '-shared', '-mfma',
             '-ffp-contract=off', '-Wno-unknown-pragmas', '-I' + _SHIM, src,
             '-o', tmp], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stderr[-3000:])
        os.replace(tmp, so)
    lib = ctypes.CDLL(so)
    lib.f_init()
    return lib


def call_host(lib, name, *arrays):
    arrays = [np.ascontiguousarray(a, dtype=np.float64) for a in arrays]
    out = np.empty_like(arrays[0])
    args = [a.ctypes.data_as(ctypes.c_void_p) for a in arrays]
    getattr(lib, 'f_' + name)(*args, out.ctypes.data_as(ctypes.c_void_p),
                              ctypes.c_long(out.size))
    return out


def device_source(div='newton'):
    code = ['typedef double Real;', '#define MKB_BX 64', '#define MKB_BY 4',
            _defines(div), kernelgen._PRELUDE,
            'extern "C" __global__ void f_div(const double* a, const double* b, double* out, long n) {',
            '    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;',
            '    if (i < n) out[i] = mkb_div(a[i], b[i]);',
            '}',
            'extern "C" __global__ void f_pow(const double* a, const double* b, double* out, long n) {',
            '    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;',
            '    if (i < n) out[i] = mkb_pow(a[i], b[i]);',
            '}',
            'extern "C" __global__ void f_rsqrt_seed(const double* b, double* out, long n) {',
            '    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;',
            '    if (i < n) { double r; MKB_ASM_RSQRT64(r, b[i]); out[i] = r; }',
            '}',
            # the raw seed, to pin the host model of rcp.approx.ftz.f64
            'extern "C" __global__ void f_rcp_seed(const double* b, double* out, long n) {',
            '    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;',
            '    if (i < n) { double r; MKB_ASM_RCP64(r, b[i]); out[i] = r; }',
            '}']
    for name in UNARY:
        code += ['extern "C" __global__ void f_%s(const double* x, double* out, long n) {' % name,
                 '    MKB_EXP_TABLE_INIT(threadIdx.x, blockDim.x);',
                 '    __syncthreads();',
                 '    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;',
                 '    if (i < n) out[i] = %s(x[i]);' % name,
                 '}']
    return '\n'.join(code)


class DeviceFunctions:
    """Launches the f_* kernels of :func:`device_source` (cuda-python driver API)."""
    def __init__(self, div='newton'):
        import torch
        from cuda.bindings import driver
        from myokit_b200 import capi
        self.torch, self.driver = torch, driver
        torch.zeros(1, device='cuda')      # primary context
        cubin, log = capi.jit_compile(device_source(div), ('--fmad=false',))
        err, self.module = driver.cuModuleLoadData(cubin)
        assert err == driver.CUresult.CUDA_SUCCESS, err

    def call(self, name, *arrays):
        torch, driver = self.torch, self.driver
        dev = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
               for a in arrays]
        out = torch.empty_like(dev[0])
        n = out.numel()
        err, fn = driver.cuModuleGetFunction(self.module, ('f_' + name).encode())
        assert err == driver.CUresult.CUDA_SUCCESS, (name, err)
        ptrs = [np.array([t.data_ptr()], dtype=np.uint64) for t in dev + [out]]
        ptrs.append(np.array([n], dtype=np.int64))
        args = np.array([p.ctypes.data for p in ptrs], dtype=np.uint64)
        err, = driver.cuLaunchKernel(fn, (n + 255) // 256, 1, 1, 256, 1, 1, 0,
                                     0, args.ctypes.data, 0)
        assert err == driver.CUresult.CUDA_SUCCESS, err
        torch.cuda.synchronize()
        return out.cpu().numpy()


def ulp_error(got, exact_hi):
    """|got - exact| in units of the last place of ``exact``; exact_hi: longdouble."""
    exact = np.asarray(exact_hi, dtype=np.longdouble)
    ref = exact.astype(np.float64)
    ulp = np.spacing(np.abs(ref)).astype(np.longdouble)
    return np.abs((np.asarray(got, dtype=np.longdouble) - exact) / ulp).astype(np.float64)
