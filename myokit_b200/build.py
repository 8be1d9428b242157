"""
Builds ``libmyokit_b200.so`` (host runtime + fixed helper kernels) in-tree with
nvcc for sm_100a. The model kernels themselves are generated per model and
JIT-compiled through the library (NVRTC), like the reference builds its OpenCL
program inside ``sim_init`` (``myokit/_sim/openclsim.c:815-836``).
"""
import hashlib
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
LIB_PATH = os.path.join(_HERE, 'libmyokit_b200.so')
_STAMP = LIB_PATH + '.stamp'

SOURCES = ['mkb_runtime.cu']
HEADERS = ['mkb_device_abi.h', 'mkb_pacing.hpp', 'mkb_schedule.hpp',
           os.path.join('..', '..', 'include', 'myokit_b200.h')]


def _nvcc():
    nvcc = shutil.which('nvcc')
    if nvcc is None and os.path.isfile('/usr/local/cuda/bin/nvcc'):
        nvcc = '/usr/local/cuda/bin/nvcc'
    if nvcc is None:
        raise RuntimeError('nvcc not found; cannot build libmyokit_b200.so')
    return nvcc


def _cuda_lib_dir(nvcc):
    root = os.path.dirname(os.path.dirname(os.path.realpath(nvcc)))
    for d in ('lib64', os.path.join('targets', 'x86_64-linux', 'lib')):
        p = os.path.join(root, d)
        if os.path.isfile(os.path.join(p, 'libnvrtc.so')):
            return p
    return os.path.join(root, 'lib64')


def _source_hash():
    h = hashlib.sha1()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def write_abi_include():
    """Embeds mkb_device_abi.h as a C++ raw string for NVRTC."""
    with open(os.path.join(CSRC, 'mkb_device_abi.h'), 'r') as f:
        text = f.read()
    out = 'R"MKBABI(' + text + ')MKBABI"\n'
    path = os.path.join(CSRC, 'mkb_device_abi_text.inc')
    old = None
    if os.path.isfile(path):
        with open(path, 'r') as f:
            old = f.read()
    if old != out:
        with open(path, 'w') as f:
            f.write(out)


def build_library(force=False, verbose=False):
    """Compiles the library if sources changed; returns its path."""
    stamp = _source_hash()
    if not force and os.path.isfile(LIB_PATH) and os.path.isfile(_STAMP):
        with open(_STAMP, 'r') as f:
            if f.read().strip() == stamp:
                return LIB_PATH
    nvcc = _nvcc()
    libdir = _cuda_lib_dir(nvcc)
    write_abi_include()
    tmp = LIB_PATH + '.tmp%d' % os.getpid()
    cmd = [
        nvcc, '-O3', '-std=c++17',
        '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
        '-Xcompiler', '-fPIC,-Wall', '-shared', '-cudart', 'static',
        '-o', tmp,
    ] + [os.path.join(CSRC, s) for s in SOURCES] + [
        '-L' + libdir, '-lnvrtc', '-ldl', '-Xlinker', '-rpath=' + libdir,
    ]
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError(
            'Building libmyokit_b200.so failed:\n' + ' '.join(cmd) + '\n'
            + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    os.replace(tmp, LIB_PATH)
    with open(_STAMP, 'w') as f:
        f.write(stamp)
    return LIB_PATH


if __name__ == '__main__':
    print(build_library(force=True, verbose=True))
