"""
Kernel generator: ``myokit.Model`` expression trees -> one fused CUDA kernel
for sm_100a.

Takes the place of the reference's pype template ``myokit/_sim/openclsim.cl``.
What is kept from the reference, because it defines the arithmetic: equation
order (``model.solvable_order()``, ``openclsim.cl:32-39``), which equations are
evaluated (``:235-243``: Rush-Larsen states' derivatives and bound variables
are skipped), variable naming (``:99-113``), the expression text (myokit's own
``CudaExpressionWriter``), the literal forms of the zero-flux stencil
(``:403-434``, ``:474-483``) and the update lines (``:358-364``).

What is different, because the target is a B200 and not an arbitrary OpenCL
device:

* ONE ``__global__`` function per model: the diffusion stencil is fused into
  the cell update through a shared-memory tile of V with a one-cell halo, and V
  is double-buffered in HBM, so there is no ``idiff`` round trip and no second
  launch per step;
* state, fields and logged intermediaries are structure-of-arrays planes
  (``state[k * stride + cid]``), so every warp load/store is one coalesced
  128/256-byte request (the reference's array-of-structs makes each a strided
  gather);
* everything is inlined into one scope: no ``calc_<component>`` functions with
  pointer outputs, constants are ``const Real`` locals that the compiler folds
  (fields turn the dependent ones into per-cell values automatically);
* the paced rectangle and the per-step scalars are runtime arguments (the
  rectangle in ``MkbGridArgs``, time/dt/pace in a device schedule ring), so a
  compiled kernel is reused across runs, protocols and pacing areas; an
  explicit paced-cell list is a byte mask (O(1) per cell) instead of one ``if``
  per paced cell;
* ``set_connections`` graphs are a CSR gather inside the same kernel
  (deterministic summation order) instead of an edge-parallel atomic scatter.
"""
import hashlib

import myokit
from myokit.formats.cuda import CudaExpressionWriter

KERNEL_NAME = 'mkb_cell_step'

DIFF_NONE, DIFF_HOMOGENEOUS, DIFF_FIELD, DIFF_CONNECTIONS = range(4)


class _NativeCudaExpressionWriter(CudaExpressionWriter):
    """
    Single-precision writer using the hardware approximations (``__expf``
    etc.), the CUDA counterpart of the reference's ``native_maths=True``
    (``native_exp`` ..., ``myokit/formats/opencl/_ewriter.py:57-95``).
    """
    def _ex_exp(self, e):
        return self._ex_function(e, '__expf')

    def _ex_log(self, e):
        if len(e) == 1:
            return self._ex_function(e, '__logf')
        return '(__logf(' + self.ex(e[0]) + ') / __logf(' + self.ex(e[1]) + '))'

    def _ex_log10(self, e):
        return self._ex_function(e, '__log10f')

    def _ex_power(self, e):
        return '__powf(' + self.ex(e[0]) + ', ' + self.ex(e[1]) + ')'

    def _ex_sin(self, e):
        return self._ex_function(e, '__sinf')

    def _ex_cos(self, e):
        return self._ex_function(e, '__cosf')

    def _ex_tan(self, e):
        return self._ex_function(e, '__tanf')

    def _ex_divide(self, e):
        return '__fdividef(' + self.ex(e[0]) + ', ' + self.ex(e[1]) + ')'


def default_block(nx, ny, precision):
    """Thread-block tile ``(bx, by)`` for a grid of ``nx * ny`` cells."""
    if ny <= 1:
        return (128, 1)
    return (32, 4)


class KernelSource:
    """Generated source plus what the runtime must know about it."""
    def __init__(self, code, block, n_state, i_vm, n_inter, n_field,
                 diffusion_mode, options):
        self.code = code
        self.block = block
        self.n_state = n_state
        self.i_vm = i_vm
        self.n_inter = n_inter
        self.n_field = n_field
        self.diffusion_mode = diffusion_mode
        self.options = tuple(options)
        self.kernel_name = KERNEL_NAME

    def key(self):
        h = hashlib.sha256()
        h.update(self.code.encode('utf-8'))
        h.update('\0'.join(self.options).encode('utf-8'))
        return h.hexdigest()


def generate(model, precision, bound_variables, inter_log, fields, rl_states,
             diffusion_mode, paced_list, block, native_maths=False, fmad=True,
             max_registers=None):
    """
    Generates the fused cell-step kernel for a prepared ``model`` (bindings
    processed and unique names created, ``openclsim.py:284-290``).

    ``inter_log`` and ``fields`` are lists of variables in storage order;
    ``rl_states`` maps state -> (inf, tau); ``paced_list`` is True when paced
    cells are an explicit list (byte mask) instead of a rectangle.
    """
    sp = (precision == myokit.SINGLE_PRECISION)
    if native_maths and sp:
        w = _NativeCudaExpressionWriter(precision)
    else:
        w = CudaExpressionWriter(precision)
    fields = list(fields)
    inter_log = list(inter_log)
    bx, by = block
    diffusion = diffusion_mode != DIFF_NONE

    def v(var):
        # openclsim.cl:99-113
        if isinstance(var, myokit.Derivative):
            return 'D_' + var.var().uname()
        if isinstance(var, myokit.Name):
            var = var.var()
        if var in bound_variables:
            return bound_variables[var]
        return 'V_' + var.uname()
    w.set_lhs_function(v)

    equations = model.solvable_order()
    del equations['*remaining*']

    n_state = model.count_states()
    vm = model.label('membrane_potential') if diffusion else None
    i_vm = vm.index() if vm is not None else -1
    real = 'float' if sp else 'double'
    exp = 'expf' if sp else 'exp'
    if native_maths and sp:
        exp = '__expf'

    out = []
    p = out.append
    p('// Generated by myokit_b200.kernelgen for sm_100a — do not edit.')
    p('// Model: %s' % model.name())
    p('#include "mkb_device_abi.h"')
    p('typedef %s Real;' % real)
    p('#define MKB_BX %d' % bx)
    p('#define MKB_BY %d' % by)
    p('')
    p('extern "C" __global__ void __launch_bounds__(MKB_BX * MKB_BY)')
    p('%s(const MkbGridArgs g, const MkbStepParams* __restrict__ sp,' % KERNEL_NAME)
    p('    const Real* __restrict__ v_in, Real* __restrict__ v_out)')
    p('{')
    p('    const unsigned int tx = threadIdx.x, ty = threadIdx.y;')
    p('    const unsigned long long nx = g.nx, ny = g.ny, stride = g.stride;')
    p('    const unsigned long long nbx = (nx + MKB_BX - 1) / MKB_BX;')
    p('    const unsigned long long bid = blockIdx.x;')
    p('    const unsigned long long ix = (bid % nbx) * MKB_BX + tx;')
    p('    const unsigned long long iy = (bid / nbx) * MKB_BY + ty;')
    p('    const bool active = (ix < nx) && (iy < ny);')
    p('    const unsigned long long cid = ix + iy * nx;')
    p('    Real* const state = (Real*)g.state;')
    p('    // Per-step scalars, cast like openclsim.c:1063,1148,1155')
    p('    const Real time = (Real)sp->time;')
    p('    const Real dt = (Real)sp->dt;')
    p('    const Real pace_in = (Real)sp->pace;')
    p('    const bool store_aux = (sp->flags & MKB_FLAG_STORE_AUX) != 0;')
    p('    (void)time; (void)pace_in; (void)store_aux; (void)v_in; (void)v_out;')
    p('')

    if diffusion_mode in (DIFF_HOMOGENEOUS, DIFF_FIELD):
        # --- V tile with halo in shared memory --------------------------
        p('    // V(t) tile + one-cell halo in shared memory. Out-of-grid halo')
        p('    // entries hold the cell\'s own V and are never used: the edge')
        p('    // formulas below drop those terms exactly as openclsim.cl does.')
        p('    const unsigned long long iyg = iy + g.iy_offset;  // global row')
        p('    const unsigned long long nyg = g.ny_global;')
        p('    __shared__ Real tile[MKB_BY + 2][MKB_BX + 2];')
        p('    const Real vc = active ? v_in[cid] : (Real)0;')
        p('    tile[ty + 1][tx + 1] = vc;')
        p('    if (active) {')
        p('        if (tx == 0) tile[ty + 1][0] = (ix > 0) ? v_in[cid - 1] : vc;')
        p('        if (tx == MKB_BX - 1 || ix == nx - 1)')
        p('            tile[ty + 1][tx + 2] = (ix < nx - 1) ? v_in[cid + 1] : vc;')
        p('        if (ty == 0) {')
        p('            Real vn = vc;')
        p('            if (iy > 0) vn = v_in[cid - nx];')
        p('            else if (iyg > 0 && g.halo_lo) vn = ((const Real*)g.halo_lo)[ix];')
        p('            tile[0][tx + 1] = vn;')
        p('        }')
        p('        if (ty == MKB_BY - 1 || iy == ny - 1) {')
        p('            Real vn = vc;')
        p('            if (iy < ny - 1) vn = v_in[cid + nx];')
        p('            else if (iyg < nyg - 1 && g.halo_hi) vn = ((const Real*)g.halo_hi)[ix];')
        p('            tile[ty + 2][tx + 1] = vn;')
        p('        }')
        p('    }')
        p('    __syncthreads();')
        p('    if (!active) return;')
        p('    const Real vxm = tile[ty + 1][tx], vxp = tile[ty + 1][tx + 2];')
        p('    const Real vym = tile[ty][tx + 1], vyp = tile[ty + 2][tx + 1];')
        p('    Real idiff;')
        if diffusion_mode == DIFF_HOMOGENEOUS:
            p('    // openclsim.cl:401-434 (diff_step)')
            p('    const Real gx = (Real)g.gx, gy = (Real)g.gy;')
            p('    if (nx > 1) {')
            p('        if (ix == 0) idiff = gx * (vc - vxp);')
            p('        else if (ix == nx - 1) idiff = gx * (vc - vxm);')
            p('        else idiff = gx * (2 * vc - vxm - vxp);')
            p('    } else {')
            p('        idiff = 0;')
            p('    }')
            p('    if (nyg > 1) {')
            p('        if (iyg == 0) idiff += gy * (vc - vyp);')
            p('        else if (iyg == nyg - 1) idiff += gy * (vc - vym);')
            p('        else idiff += gy * (2 * vc - vym - vyp);')
            p('    }')
        else:
            p('    // openclsim.cl:469-486 (diff_hetero); gx[(ny, nx-1)], gy[(ny-1, nx)].')
            p('    // Both pointers are slab-relative: gyf[-nx .. -1] is the gy row that')
            p('    // couples this slab\'s first row to the slab above it.')
            p('    const Real* const gxf = (const Real*)g.gx_field;')
            p('    const Real* const gyf = (const Real*)g.gy_field;')
            p('    idiff = 0.0;')
            p('    if (nx > 1) {')
            p('        if (ix > 0) { idiff += gxf[cid - iy - 1] * (vc - vxm); }')
            p('        if (ix < nx - 1) { idiff += gxf[cid - iy] * (vc - vxp); }')
            p('    }')
            p('    if (nyg > 1) {')
            p('        if (iyg > 0) idiff += gyf[(long long)cid - (long long)nx] * (vc - vym);')
            p('        if (iyg < nyg - 1) idiff += gyf[cid] * (vc - vyp);')
            p('    }')
    elif diffusion_mode == DIFF_CONNECTIONS:
        p('    if (!active) return;')
        p('    // openclsim.cl:537-556 as a per-cell CSR gather: same terms')
        p('    // g * (V_i - V_j), summed in edge-list order, no atomics.')
        p('    const Real vc = v_in[cid];')
        p('    Real idiff = 0;')
        p('    {')
        p('        const Real* const cg = (const Real*)g.csr_g;')
        p('        const unsigned long long e1 = g.csr_row[cid + 1];')
        p('        for (unsigned long long e = g.csr_row[cid]; e < e1; e++) {')
        p('            idiff += cg[e] * (vc - v_in[g.csr_col[e]]);')
        p('        }')
        p('    }')
    else:
        p('    if (!active) return;')
    p('')

    # --- pacing ---------------------------------------------------------
    if diffusion:
        p('    // openclsim.cl:249-280, 322-329')
        if paced_list:
            p('    const Real pace = g.paced_mask[cid] ? pace_in : (Real)0;')
        else:
            if diffusion_mode == DIFF_CONNECTIONS:
                p('    const long long pix = (long long)ix, piy = 0;')
            else:
                p('    const long long pix = (long long)ix, piy = (long long)iyg;')
            p('    const Real pace = (pix >= g.pace_x0 && pix < g.pace_x1 &&')
            p('                       piy >= g.pace_y0 && piy < g.pace_y1) ? pace_in : (Real)0;')
        p('    if (store_aux) ((Real*)g.idiff)[cid] = idiff;')
    else:
        p('    const Real pace = pace_in;')
    p('    (void)pace;')
    p('')

    # --- fields, constants, states --------------------------------------
    p('    // Scalar fields (set_field): one plane each')
    for k, var in enumerate(fields):
        p('    const Real %s = ((const Real*)g.field)[%dull * stride + cid];'
          % (v(var), k))
    p('    // Literal constants (openclsim.cl:173-178)')
    for group in equations.values():
        for eq in group.equations(const=True):
            if isinstance(eq.rhs, myokit.Number):
                if eq.lhs.var() not in fields:
                    p('    const Real %s = %s;' % (v(eq.lhs), w.ex(eq.rhs)))
    p('    // Calculated constants (openclsim.cl:181-186); folded at compile')
    p('    // time unless they depend on a field')
    for group in equations.values():
        for eq in group.equations(const=True):
            if not isinstance(eq.rhs, myokit.Number):
                if eq.lhs.var() not in fields:
                    p('    const Real %s = %s;' % (v(eq.lhs), w.ex(eq.rhs)))
    p('    // States at time t')
    for var in model.states():
        k = var.index()
        if k == i_vm:
            p('    const Real %s = vc;' % v(var))
        else:
            p('    const Real %s = state[%dull * stride + cid];' % (v(var), k))
    p('')

    # --- equations --------------------------------------------------------
    inter_index = dict((var, k) for k, var in enumerate(inter_log))
    for name, group in equations.items():
        eqs = []
        for eq in group.equations(const=False):
            var = eq.lhs.var()
            if var in rl_states or var in bound_variables:
                continue
            eqs.append(eq)
        if not eqs:
            continue
        p('    // Component: %s' % name)
        for eq in eqs:
            var = eq.lhs.var()
            p('    const Real %s = %s;' % (v(eq.lhs), w.ex(eq.rhs)))
            if var in inter_index and not eq.lhs.is_derivative():
                p('    if (store_aux) ((Real*)g.inter)[%dull * stride + cid] = %s;'
                  % (inter_index[var], v(eq.lhs)))
    p('')

    # --- update -----------------------------------------------------------
    p('    // Update (openclsim.cl:358-364)')
    for var in model.states():
        k = var.index()
        if var in rl_states:
            inf, tau = rl_states[var]
            inf, tau, x = v(inf), v(tau), v(var)
            rhs = '%s - (%s - %s) * %s(-dt / %s)' % (inf, inf, x, exp, tau)
        else:
            rhs = '%s + dt * %s' % (v(var), v(var.lhs()))
        if k == i_vm:
            p('    v_out[cid] = %s;' % rhs)
        else:
            p('    state[%dull * stride + cid] = %s;' % (k, rhs))
    p('}')
    p('')

    options = ['--fmad=true' if fmad else '--fmad=false']
    if max_registers:
        options.append('--maxrregcount=%d' % int(max_registers))
    return KernelSource('\n'.join(out), block, n_state, i_vm, len(inter_log),
                        len(fields), diffusion_mode, options)
