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


def _integer_exponent(e):
    """Returns k if ``e`` is a literal with a small integer value, else None."""
    try:
        if not e.is_literal():
            return None
        x = e.eval()
    except Exception:
        return None
    if x != x or x in (float('inf'), float('-inf')):
        return None
    k = int(x)
    if k != x or abs(k) > 32:
        return None
    return k


class _PowMixin:
    """
    ``pow(x, k)`` with a small integer literal ``k`` becomes a fixed
    square-and-multiply chain (``mkb_powi<k>``): 2-7 multiplications instead
    of the ~100-instruction general ``pow``. Differs from ``pow`` by at most
    a few ulp; covered by the parity tests against the oracle (which keeps
    ``pow``). Exponent 0.5 becomes ``sqrt``. Anything else stays ``pow``.
    """
    _pow_multiply = True

    def _ex_power(self, e):
        if self._pow_multiply:
            k = _integer_exponent(e[1])
            if k is not None:
                return 'mkb_powi<%d>(%s)' % (k, self.ex(e[0]))
            # half-integer exponents: x^(k + 1/2) = x^k * sqrt(x)
            try:
                x2 = 2 * e[1].eval() if e[1].is_literal() else None
            except Exception:
                x2 = None
            if x2 is not None and x2 == int(x2) and 0 < x2 <= 15:
                k = (int(x2) - 1) // 2
                sq = 'sqrtf' if self._sp else 'sqrt'
                if getattr(self, '_fast_libm', False) and not self._sp:
                    sq = 'mkb_sqrt'
                if k == 0:
                    return '%s(%s)' % (sq, self.ex(e[0]))
                return 'mkb_powh<%d>(%s)' % (k, self.ex(e[0]))
        return super()._ex_power(e)


class _DivMixin:
    """Optional branch-free division (``mkb_div``), see the kernel prelude."""
    _fast_div = False
    # Set by generate(): expression -> float for compile-time constants
    # (literals and folded model constants), else None.
    const_value = None

    def _ex_divide(self, e):
        if self._fast_div:
            if self.const_value is not None:
                # x / c with a compile-time c: one multiplication by 1 / c
                # (<= 1 ulp from the quotient, like mkb_div itself) instead
                # of a reciprocal seed, two Newton steps and a correction.
                c = self.const_value(e[1])
                if c is not None and c != 0:
                    r = 1.0 / c
                    if r == r and 1e-290 < abs(r) < 1e290:
                        return ('((' + self.ex(e[0]) + ') * '
                                + self.ex(myokit.Number(r)) + ')')
            return 'mkb_div(' + self.ex(e[0]) + ', ' + self.ex(e[1]) + ')'
        return super()._ex_divide(e)


class _ConstPoolMixin:
    """
    Double-precision literals that do not fit an instruction immediate go to
    one ``__constant__`` table (``mkb_k``), in order of first use. sm_100a has
    no 64-bit immediates: ptxas otherwise builds every such literal from two
    32-bit moves (UMOV / IMAD.MOV), which was ~30% of all issued instructions
    of a large fp64 model; from the table two adjacent constants arrive with
    one uniform 128-bit load (LDCU.128) or are used as c-bank operands.
    Literals whose low 32 mantissa bits are zero (1.0, 0.5, -80.0, ...) stay
    in the instruction stream, where they are free.
    """
    _pool = None        # list of floats, or None when pooling is off

    @staticmethod
    def _is_cheap(x):
        import struct
        bits = struct.unpack('<Q', struct.pack('<d', x))[0]
        return (bits & 0xffffffff) == 0

    def pool_ref(self, x):
        x = float(x)
        if self._pool is None or x != x or self._is_cheap(x):
            return None
        key = x.hex()
        i = self._pool_index.get(key)
        if i is None:
            i = len(self._pool)
            self._pool.append(x)
            self._pool_index[key] = i
        return 'mkb_k[%d]' % i

    def enable_pool(self):
        self._pool = []
        self._pool_index = {}

    def _ex_number(self, e):
        if self._pool is not None:
            ref = self.pool_ref(e.eval())
            if ref is not None:
                return ref
        return super()._ex_number(e)


class _ExpMixin:
    """Optional in-line double-precision exp (``mkb_exp``), see the prelude."""
    _fast_exp = False

    def _ex_exp(self, e):
        if self._fast_exp and not self._sp:
            return self._ex_function(e, self._fast_exp)
        if self._fast_exp == 'mkb_expf_ex2' and self._sp:
            return self._ex_function(e, 'mkb_expf_ex2')
        return super()._ex_exp(e)


class _LibmMixin:
    """
    Optional branch-free double-precision ``sqrt`` / ``log`` / ``cos`` /
    ``acos`` / general ``pow`` (``mkb_sqrt`` ..., see the prelude), and
    conditional expressions as selects: both arms are evaluated and one is
    chosen (``mkb_sel``), where the ternary operator the host framework's
    writer emits becomes a branch whenever an arm holds a division or an exp.
    Either keeps the kernel body one straight line for the instruction
    scheduler.
    """
    _fast_libm = False
    _select = False

    def _libm(self, e, name, fallback):
        if self._fast_libm and not self._sp:
            return self._ex_function(e, name)
        return fallback(e)

    def _ex_sqrt(self, e):
        return self._libm(e, 'mkb_sqrt', super()._ex_sqrt)

    def _ex_log(self, e):
        if len(e) != 1:
            return super()._ex_log(e)
        return self._libm(e, 'mkb_log', super()._ex_log)

    def _ex_cos(self, e):
        return self._libm(e, 'mkb_cos', super()._ex_cos)

    def _ex_acos(self, e):
        return self._libm(e, 'mkb_acos', super()._ex_acos)

    def _ex_power(self, e):
        text = super()._ex_power(e)
        if self._fast_libm and not self._sp and text.startswith('pow('):
            return 'mkb_pow(' + self.ex(e[0]) + ', ' + self.ex(e[1]) + ')'
        return text

    _COSTLY = (myokit.Exp, myokit.Log, myokit.Log10, myokit.Power, myokit.Sin,
               myokit.Cos, myokit.Tan, myokit.ASin, myokit.ACos, myokit.ATan)

    def _selectable(self, arms):
        """Whether evaluating all ``arms`` is cheap enough to drop the branch."""
        if self._select == 'cheap':
            for arm in arms:
                if isinstance(arm, self._COSTLY):
                    return False
                for x in arm.walk(self._COSTLY):
                    return False
        return bool(self._select)

    def _ex_if(self, e):
        if not self._selectable([e._t, e._e]):
            return super()._ex_if(e)
        return 'mkb_sel(%s, %s, %s)' % (
            self.ex(e._i), self.ex(e._t), self.ex(e._e))

    def _ex_piecewise(self, e):
        if not self._selectable(list(e._e)):
            return super()._ex_piecewise(e)
        ifs = [self.ex(x) for x in e._i]
        thens = [self.ex(x) for x in e._e]
        text = thens[-1]
        for c, t in zip(reversed(ifs), reversed(thens[:-1])):
            text = 'mkb_sel(%s, %s, %s)' % (c, t, text)
        return text


class _Writer(_ConstPoolMixin, _LibmMixin, _PowMixin, _DivMixin, _ExpMixin,
              CudaExpressionWriter):
    """
    The writer used for all but ``native_maths`` kernels. ``fold`` (set by
    :func:`generate`) maps a constant sub-expression to its text, or returns
    None: the in-line division / exp contain ``asm`` and cannot be folded by
    the compiler, so constant sub-trees are folded here.
    """
    fold = None

    def ex(self, e):
        if self.fold is not None and not isinstance(
                e, (myokit.Number, myokit.Name)):
            text = self.fold(e)
            if text is not None:
                return text
        return super().ex(e)


class _NativeWriter(_PowMixin, _NativeCudaExpressionWriter):
    pass


_PRELUDE = r"""
// The two pieces of inline PTX go through macros so that a host build of this
// source (tests/cuda_shim, test infrastructure) can supply its own.
#ifndef MKB_ASM_RCP64
#define MKB_ASM_RCP64(r, b) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b))
#define MKB_ASM_RSQRT64(r, b) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b))
#define MKB_ASM_SREG(v, name) asm volatile("mov.u32 %0, %%" name ";" : "=r"(v))
#define MKB_ASM_EX2F(y, t) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(t))
#define MKB_PREFETCH_L1(p) asm volatile("prefetch.global.L1 [%0];" :: "l"(p))
#define MKB_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" :: "l"(p))
// Streaming kernels: mbarrier + TMA (cp.async.bulk.tensor) and warp shuffles
#define MKB_SMEM_ADDR(p) ((unsigned int)__cvta_generic_to_shared(p))
#define MKB_MBAR_INIT(bar, count) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(MKB_SMEM_ADDR(bar)), "r"(count))
#define MKB_MBAR_FENCE_INIT() asm volatile("fence.mbarrier_init.release.cluster;\n\tfence.proxy.async.shared::cta;" ::: "memory")
// one arrival that announces `bytes`, then the 2-d box (bw x bh elements) whose
// first element is (cx, cy) of the tensor `tmap` describes; cells outside the
// tensor arrive as zeros
#define MKB_TMA_LOAD_2D(dst, tmap, cx, cy, bar, bw, bh, bytes) do { \
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" \
                 :: "r"(MKB_SMEM_ADDR(bar)), "r"(bytes) : "memory"); \
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" \
                 :: "r"(MKB_SMEM_ADDR(dst)), "l"((unsigned long long)(tmap)), "r"(cx), "r"(cy), \
                    "r"(MKB_SMEM_ADDR(bar)) : "memory"); \
} while (0)
#define MKB_MBAR_WAIT(bar, parity) do { \
    unsigned int done_; \
    do { \
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" \
                     : "=r"(done_) : "r"(MKB_SMEM_ADDR(bar)), "r"(parity) : "memory"); \
    } while (!done_); \
} while (0)
#define MKB_SHFL_UP(v, d) __shfl_up_sync(0xffffffffu, (v), (d))
#define MKB_SHFL_DOWN(v, d) __shfl_down_sync(0xffffffffu, (v), (d))
// Staged kernels (option stage): the thread block's tile of every state plane
// travels HBM <-> shared memory by TMA, one box of (MKB_BX, MKB_BY, 1 plane)
// per instruction, through the 3-d descriptor MkbGridArgs::tmap_state
// ([plane][row][column]); cells outside the grid arrive as zeros and are not
// written back.
#define MKB_STAGE_DECL(bytes) extern __shared__ __align__(128) unsigned char mkb_stage_mem[]
#define MKB_SYNCWARP() __syncwarp()
#define MKB_MBAR_EXPECT_TX(bar, bytes) \
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" \
                 :: "r"(MKB_SMEM_ADDR(bar)), "r"(bytes) : "memory")
#define MKB_TMA_LOAD_3D(dst, tmap, cx, cy, cz, bar) \
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" \
                 :: "r"(MKB_SMEM_ADDR(dst)), "l"((unsigned long long)(tmap)), "r"(cx), "r"(cy), "r"(cz), \
                    "r"(MKB_SMEM_ADDR(bar)) : "memory")
#define MKB_TMA_STORE_3D(tmap, cx, cy, cz, src) \
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" \
                 :: "l"((unsigned long long)(tmap)), "r"(cx), "r"(cy), "r"(cz), "r"(MKB_SMEM_ADDR(src)) : "memory")
// commit the stores issued so far and wait until they have been performed
#define MKB_TMA_STORE_FINISH() \
    asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group 0;" ::: "memory")
// the same box only as far as L2 (no destination): a hint
#define MKB_TMA_PREFETCH_3D(tmap, cx, cy, cz) \
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" \
                 :: "l"((unsigned long long)(tmap)), "r"(cx), "r"(cy), "r"(cz) : "memory")
#define MKB_NSM(v) asm("mov.u32 %0, %%nsmid;" : "=r"(v))
// tile-loop kernels: commit / wait separately; per-thread asynchronous copies
// (LDGSTS) of the next tile's membrane potentials
#define MKB_TMA_STORE_COMMIT() asm volatile("cp.async.bulk.commit_group;" ::: "memory")
#define MKB_TMA_STORE_READ_WAIT() asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory")
#define MKB_CP_ASYNC(dst, src) \
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" \
                 :: "r"(MKB_SMEM_ADDR(dst)), "l"(src), "n"(sizeof(Real)) : "memory")
#define MKB_CP_ASYNC_WAIT() asm volatile("cp.async.wait_all;" ::: "memory")
// commit, and wait only until the shared memory has been read
#define MKB_TMA_STORE_READ_DONE() \
    asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group.read 0;" ::: "memory")
// shared-memory writes of this thread become visible to the TMA unit
#define MKB_FENCE_ASYNC_SMEM() asm volatile("fence.proxy.async.shared::cta;" ::: "memory")
// global writes of the TMA unit observed through a flag become visible to later TMA reads
#define MKB_FENCE_ASYNC_GLOBAL() asm volatile("fence.proxy.async.global;" ::: "memory")
// max(min(a, b), 0) in one instruction (VIMNMX.RELU)
#define MKB_MIN_RELU(d, a, b) asm("min.relu.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))
// Kernels generated for one plane stride (option plane_stride) refuse any other
#define MKB_STRIDE_MISMATCH() __trap()
#endif

// Overlapping consecutive steps (option overlap). A step kernel normally starts
// when the previous one has drained: every SM then idles through the tail of
// the last wave — about half a thread's latency (27 us for the 48-state
// model) per step, 2 % of a 2048^2 step but 12 % of a 256-row slab's. Kernels
// generated with MKB_OVERLAP_STEPS are launched with programmatic stream
// serialization: each block lets the next kernel start launching as soon as
// it runs itself (griddepcontrol.launch_dependents), and orders itself
// against the previous step by data instead of by kernel boundary: it waits
// until its own tile and its four neighbours have published step - 1 in
// MkbGridArgs::tile_done (their V(t) rows are then written, and they have
// finished reading the V plane this block is about to overwrite), and
// publishes `step` behind a device-scope release when its stores are done.
// All earlier blocks are resident or finished whenever a block waits (the
// next grid only launches once every block of this one has started), so the
// waits cannot deadlock. Loads of data another step wrote bypass L1
// (ld.global.cg): the previous kernel may still be running on this SM.
#ifndef MKB_OVERLAP_STEPS
#define MKB_OVERLAP_STEPS 0
#endif
#if MKB_OVERLAP_STEPS
#define MKB_LD(p) __ldcg(p)
#else
#define MKB_LD(p) (*(p))
#endif
#ifndef MKB_PDL_TRIGGER
#define MKB_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
__device__ __forceinline__ void mkb_wait_tile(const unsigned int* flag, unsigned int want, unsigned int* error) {
    unsigned int seen, spins = 0;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
        if (seen >= want) break;
        __nanosleep(spins < 64 ? 20 : 100);
        if (++spins > 20000000u) {      // ~2 s: surface as an error, never hang the GPU
            if (error) atomicExch(error, 2u);
            break;
        }
    }
}
__device__ __forceinline__ void mkb_publish_tile(unsigned int* flag, unsigned int step) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(flag), "r"(step) : "memory");
}
#endif

// Plane k of a cell, from the cell's address in plane 0. With the plane
// stride as a run-time value every access costs three integer instructions
// (IMAD.WIDE.U32 + IMAD + IADD for the 64 x 64-bit product; a 32-bit stride
// and an inline mad.wide.u32 were tried: ptxas knows the stride to be uniform
// and turns both into a uniform multiply, a copy to ordinary registers and a
// 64-bit addition in two halves, no fewer). Kernels are compiled per
// simulation anyway, so the main kernel takes the stride as a compile-time
// constant (option plane_stride): the access is then base + constant, two
// instructions, and the stride, its multiples and their copies leave the
// register file: 640 of the 4200 instructions of the 48-state kernel.
#define MKB_AT(base, k) ((base)[(unsigned long long)(k) * stride])

// Conditional expression with both arms evaluated (function arguments are),
// then one select: no branch.
__device__ __forceinline__ Real mkb_sel(bool c, Real a, Real b) { return c ? a : b; }

// x^k for a compile-time integer k: square-and-multiply, fixed order.
template <int N>
__device__ __forceinline__ Real mkb_powi(Real x) {
    if constexpr (N < 0) {
        return (Real)1 / mkb_powi<-N>(x);
    } else if constexpr (N == 0) {
        return (Real)1;
    } else if constexpr (N == 1) {
        return x;
    } else if constexpr (N % 2 == 0) {
        const Real h = mkb_powi<N / 2>(x);
        return h * h;
    } else {
        return x * mkb_powi<N - 1>(x);
    }
}

// Thread / block coordinates read again where they are needed (volatile: the
// compiler may not reuse an earlier read), so that they do not occupy
// registers across the whole cell model between the tile load at the top of a
// slab kernel and its flag publication at the bottom (option slab_lean).
struct MkbSlabPos {
    unsigned int tx, ty, bxb, byb, nby;
};
template <int BY>
__device__ __forceinline__ MkbSlabPos mkb_slab_pos(unsigned int ny) {
    unsigned int by, bz, gy;
    MkbSlabPos p;
    MKB_ASM_SREG(p.tx, "tid.x");
    MKB_ASM_SREG(p.ty, "tid.y");
    MKB_ASM_SREG(p.bxb, "ctaid.x");
    MKB_ASM_SREG(by, "ctaid.y");
    MKB_ASM_SREG(bz, "ctaid.z");
    MKB_ASM_SREG(gy, "nctaid.y");
    p.nby = (ny + BY - 1) / BY;
    const unsigned int byr = by + bz * gy;
    p.byb = (byr == 0) ? 0 : ((byr == 1) ? p.nby - 1 : byr - 1);
    return p;
}

// Ghost-row arrival: spin (with back-off) until the neighbouring GPU has
// delivered the row for `step`; gives up after ~10 s and raises halo_error so
// a stalled neighbour surfaces as an error instead of a hung GPU.
__device__ __forceinline__ void mkb_wait_flag(
    const unsigned int* flag, unsigned int step, unsigned int* error) {
    const volatile unsigned int* f = flag;
    unsigned int spins = 0;
    while (*f < step) {
        __nanosleep(spins < 64 ? 20 : 200);
        if (++spins > 50000000u) {
            if (error) atomicExch(error, 1u);
            break;
        }
    }
}

// Branch-free division (option fast_div): hardware reciprocal seed (~2^-23),
// one Newton step on the reciprocal (~2^-46), then one residual correction of
// the quotient (error ~2^-92 before the final rounding): within 1 ulp, not
// guaranteed correctly rounded (20 M random operand pairs on the host, with a
// 20-bit seed: all correctly rounded). Operands and quotient must be in the
// normal range (|x| in [2^-1000, 2^1000]); denormal divisors are not
// supported. A NaN operand or an infinite dividend give the IEEE result; a
// divisor of exactly 0 or inf gives NaN (IEEE: inf / 0) unless the kernel is
// built with div_parallel, whose unrefined product a * rcp(b) is at hand for
// those cases.
__device__ __forceinline__ double mkb_div(double a, double b) {
    double r;
    MKB_ASM_RCP64(r, b);
    const double e = fma(-b, r, 1.0);
#if MKB_DIV_CUBIC
    // Option div_cubic: one third-order step on the reciprocal, r (1 + e + e^2)
    // (relative error e^3 < 2^-60 from the >= 20-bit seed), then the product:
    // 4 FP64 instructions instead of 6, no residual correction and therefore
    // no NaN from inf - inf for an infinite dividend. A divisor of 0 or inf
    // gives a seed of inf / 0 and e = NaN (either sign) or inf. The correction
    // factor e + e^2 is then replaced by a finite number with ONE instruction
    // on the otherwise idle FP32 pipe: its high word, read as a float, is a
    // float NaN exactly when the double is NaN or inf (exponent field all
    // ones), and fminf(x, 1.0f) returns 1.0f for a NaN x and x itself — same
    // bits — for every other value a correction can take (|e + e^2| < 2^-19:
    // the float reading is below 1.0f; flush-to-zero only concerns doubles
    // below 2^-1015, and an exact zero stays zero). So r stays inf / 0 and
    // a * r is the IEEE quotient (a / 0 = inf, a / inf = 0, 0 / 0 = NaN).
    // (Option div_int_check: the first form of this test, on the seed's
    // exponent field with the integer pipe — add, and, compare, select: 4
    // issue slots per division; kept for comparison.)
    // Error <= 1.5 ulp of the quotient (two roundings: reciprocal and product;
    // reciprocals 1 / b: <= 1 ulp); denormal divisors count as 0
    // (rcp.approx.ftz).
    double e2 = fma(e, e, e);
#if MKB_DIV_INT_CHECK
    const unsigned int rh = (unsigned int)__double2hiint(r);
    const bool special = ((rh + 0x00100000u) & 0x7fe00000u) == 0u;
    e2 = __hiloint2double(special ? 0x3ff00000 : __double2hiint(e2), __double2loint(e2));
#else
    e2 = __hiloint2double(__float_as_int(fminf(__int_as_float(__double2hiint(e2)), 1.0f)),
                          __double2loint(e2));
#endif
    r = fma(r, e2, r);
    return a * r;
#endif
#if MKB_DIV_PARALLEL
    // Quotient and reciprocal are refined side by side (dependent chain of 4
    // instead of 5 after the seed, same instruction count).
    const double q0 = a * r;
    const double q1 = fma(q0, e, q0);
    r = fma(r, e, r);
    const double dp = fma(-b, q1, a);
    const double qp = fma(dp, r, q1);
    // NaN residual: special operands; a * rcp(b) is the IEEE answer then
    return (dp == dp) ? qp : q0;
#endif
    r = fma(r, e, r);
    double q = a * r;
    const double d = fma(-b, q, a);
    // A NaN residual (special operands): keep the uncorrected product.
#if MKB_DIV_INT_CHECK
    // the same test on the integer pipe (exponent field all zeros / ones)
    const unsigned int eb = ((unsigned int)__double2hiint(b) << 1) + 0x00200000u;
    const unsigned int ea = ((unsigned int)__double2hiint(a) << 1) + 0x00200000u;
    if (!(eb < 0x00400000u || ea < 0x00200000u)) q = fma(d, r, q);
#else
    if (d == d) q = fma(d, r, q);   // (ptxas turns this into DSETP + 2 FSEL)
#endif
    return q;
}
__device__ __forceinline__ float mkb_div(float a, float b) {
    return __fdividef(a, b);
}

// exp(x) in double precision (option fast_exp): Cody-Waite reduction with a
// fused multiply-add, degree-11 polynomial (scripts/gen_exp_coeffs.py, max
// error 0.85 ulp against mpmath for |x| < 708), no branches. The power of two
// is built on the integer pipe from n clamped to [-1023, 1024] and applied
// with one multiplication, so the ends of the range follow IEEE: the scale is
// exactly 0 for n <= -1023 (exp underflows to 0; results below 2^-1022 are
// flushed instead of going through the denormals) and +inf for n >= 1024 (exp
// overflows to +inf, from x > 709.44 on: up to 0.05 % below the true
// threshold 709.78). NaN gives NaN. Valid for |x| < 2^31 ln 2 = 1.4e9; beyond
// that, and for infinite x, n wraps and the result is unspecified (mkb_pow
// clamps its exponent argument for that reason). The coefficients live in
// constant memory so they arrive as c-bank operands / paired uniform loads
// instead of two 32-bit moves each.
__constant__ double mkb_exp_c[14] = {
@EXP_TABLE@
};
__device__ __forceinline__ double mkb_exp_scale(int n) {
    const int nc = min(max(n, -1023), 1024);
    return __hiloint2double((nc + 1023) << 20, 0);
}
// p 2^n for p in [1/2, 2): by multiplication (above), or — option
// exp_scale='add' — by adding n to the exponent field, one integer
// instruction instead of one on the FP64 pipe; the result then saturates
// (x > 709.4: above 6e307 but finite; x < -708: below 4.5e-308 but not 0).
#ifndef MKB_EXP_SCALE_ADD
#define MKB_EXP_SCALE_ADD 0
#endif
__device__ __forceinline__ double mkb_exp_apply(double p, int n) {
#if MKB_EXP_SCALE_ADD
    const int nc = min(max(n, -1021), 1023);
    return __hiloint2double(__double2hiint(p) + (nc << 20), __double2loint(p));
#else
    return p * mkb_exp_scale(n);
#endif
}
__device__ __forceinline__ double mkb_exp_poly(double x) {
#if MKB_EXP_SCALE_ADD
    double t = fma(x, mkb_exp_c[0], 6755399441055744.0);
    const int n = __double2loint(t);
    t -= 6755399441055744.0;
#else
    // The rounding shift carries the exponent bias: the low word of t is
    // n + 1023, which ONE min-with-relu clamps to [0, 2047] = the exponent
    // fields of 0 and +inf (see mkb_exp_scale; its clamp was two instructions
    // and the bias a third).
    double t = fma(x, mkb_exp_c[0], 6755399441056767.0);     // 1.5 * 2^52 + 1023
    int nb;
    MKB_MIN_RELU(nb, __double2loint(t), 2047);
    t -= 6755399441056767.0;
#endif
    double r = fma(t, mkb_exp_c[2], x);
    r = fma(t, mkb_exp_c[3], r);
    double p = mkb_exp_c[4];
#pragma unroll
    for (int k = 5; k < 14; k++) p = fma(p, r, mkb_exp_c[k]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
#if MKB_EXP_SCALE_ADD
    return mkb_exp_apply(p, n);
#else
    return p * __hiloint2double(nb << 20, 0);
#endif
}

// Single precision (option fast_exp = 'ex2'): expf(x) = 2^t with t = x log2(e)
// carried as a rounded product plus its exact residual, so that the argument
// error does not grow with |x| as it does in __expf: FMUL, 2 FFMA, MUFU.EX2,
// FMUL, FFMA = 6 instructions against the 8-10 of libdevice's expf (which
// splits off 2^n by hand; ex2.approx saturates to 0 / inf by itself).
// Error: the 2^-22.5 of ex2.approx plus up to 2 ulp. Overflow gives inf and
// underflow 0 as they should (the correction is a factor next to 1, so it
// never meets inf - inf); infinite x and |x| > 2e38 are not supported.
__device__ __forceinline__ float mkb_expf_ex2(float x) {
    const float t = x * 1.44269502e+0f;
    float r = fmaf(x, 1.44269502e+0f, -t);      // exact: x * L2E_hi - t
    r = fmaf(x, 1.92596303e-8f, r);             // + x * L2E_lo
    float y;
    MKB_ASM_EX2F(y, t);
    return y * fmaf(r, 6.93147182e-1f, 1.0f);   // 2^(t + r) = 2^t (1 + r ln 2)
}

// Estrin variant (option fast_exp = 'estrin'): the same reduction and
// coefficients, the polynomial evaluated as a tree — dependent chain of 6
// instead of 11 fused multiply-adds, for 3 more FP64 instructions. Max error
// 0.97 ulp (scripts/gen_exp_coeffs.py). For kernels that wait on dependent
// FP64 results (stall_wait) more than on the FP64 pipe itself.
__device__ __forceinline__ double mkb_exp_estrin(double x) {
    double t = fma(x, mkb_exp_c[0], mkb_exp_c[1]);
    const int n = __double2loint(t);
    t -= mkb_exp_c[1];
    double r = fma(t, mkb_exp_c[2], x);
    r = fma(t, mkb_exp_c[3], r);
    const double r2 = r * r;
    const double a1 = fma(mkb_exp_c[12], r, mkb_exp_c[13]);    // c2 + c3 r
    const double a2 = fma(mkb_exp_c[10], r, mkb_exp_c[11]);    // c4 + c5 r
    const double a3 = fma(mkb_exp_c[8], r, mkb_exp_c[9]);      // c6 + c7 r
    const double a4 = fma(mkb_exp_c[6], r, mkb_exp_c[7]);      // c8 + c9 r
    const double a5 = fma(mkb_exp_c[4], r, mkb_exp_c[5]);      // c10 + c11 r
    const double r4 = r2 * r2;
    const double b0 = fma(a2, r2, a1);
    const double b1 = fma(a4, r2, a3);
    const double r8 = r4 * r4;
    const double d = fma(b1, r4, b0);
    const double q = fma(a5, r8, d);
    double p = fma(r2, q, r);
    p += 1.0;
    return mkb_exp_apply(p, n);
}

// Table variant (option fast_exp = 'table'): exp(x) = 2^m T[j] e^r
// with n = rint(64 x / ln2) = 64 m + j and |r| <= ln2 / 128, so e^r - 1 needs
// only a degree-5 polynomial: 10 FP64-pipe instructions instead of 17, plus
// one cached 8-byte table load. Max error 1.0 ulp (scripts/gen_exp_coeffs.py).
// Measured on C3 it is no faster than the polynomial form (the kernel is as
// much issue- as FP64-bound and this variant trades 7 FP64 for 6 integer /
// load instructions), so the polynomial form is the default.
// Saturation as in mkb_exp_stab below.
__device__ const double mkb_exp_t[64] = {
@EXP_POW2@
};
__constant__ double mkb_expt_c[8] = {
@EXP_TCOEF@
};
// Shared-memory table variant (option fast_exp = 'stab'): the same algorithm
// with the 64-entry table copied to shared memory by every thread block
// (short, fixed latency instead of a global load) and the power of two
// applied to the table entry before the last multiply-add. n is clamped to
// [-1023 * 64, 1023 * 64 + 63]: at the lower end the scaled entry is exactly
// 0, so exp underflows to 0 (results below 2^-1022 are flushed, and between
// -709.8 and -709.1 some tiny denormal comes out); at the upper end it is
// 1.98 * 2^1023, so exp saturates just below DBL_MAX instead of overflowing
// to +inf (inf * p + inf would be NaN for a negative p) — never NaN or inf
// for a finite argument, and 1 / (1 + exp(big)) is 0. 10 FP64-pipe
// instructions.
__shared__ double mkb_exp_ts[64];
#define MKB_EXP_TABLE_INIT(tid, nthreads) do { for (unsigned int i_ = (tid); i_ < 64u; i_ += (nthreads)) mkb_exp_ts[i_] = mkb_exp_t[i_]; } while (0)
__device__ __forceinline__ double mkb_exp_stab(double x) {
    double t = fma(x, mkb_expt_c[0], mkb_expt_c[1]);
    const int n = __double2loint(t);
    t -= mkb_expt_c[1];
    double r = fma(t, mkb_expt_c[2], x);
    r = fma(t, mkb_expt_c[3], r);
    const int nc = min(max(n, -1023 * 64), 1023 * 64 + 63);
    const double tj = mkb_exp_ts[nc & 63];
    double p = fma(mkb_expt_c[4], r, mkb_expt_c[5]);
    p = fma(p, r, mkb_expt_c[6]);
    p = fma(p, r, mkb_expt_c[7]);
    p = fma(r * r, p, r);
    const double ts = __hiloint2double(__double2hiint(tj) + ((nc >> 6) << 20), __double2loint(tj));
    return fma(ts, p, ts);
}
__device__ __forceinline__ double mkb_exp_tab(double x) {
    double t = fma(x, mkb_expt_c[0], mkb_expt_c[1]);
    const int n = __double2loint(t);
    t -= mkb_expt_c[1];
    double r = fma(t, mkb_expt_c[2], x);
    r = fma(t, mkb_expt_c[3], r);
    double p = fma(mkb_expt_c[4], r, mkb_expt_c[5]);
    p = fma(p, r, mkb_expt_c[6]);
    p = fma(p, r, mkb_expt_c[7]);
    p = fma(r * r, p, r);
    const int nc = min(max(n, -1023 * 64), 1023 * 64 + 63);
    const double tj = __ldg(&mkb_exp_t[nc & 63]);
    const double ts = __hiloint2double(__double2hiint(tj) + ((nc >> 6) << 20), __double2loint(tj));
    return fma(ts, p, ts);
}

// ---------------------------------------------------------------------------
// Branch-free double-precision sqrt / log / cos / acos / pow (option
// fast_libm). libdevice's versions are accurate but each contains branches
// (special operands, slow paths), and a branch ends the region in which ptxas
// can interleave independent dependency chains: a large model kernel built on
// them is a string of short blocks, each waiting on its own chain of FP64
// latencies. These have no branches, handle special operands with selects,
// and use fewer integer instructions. Polynomials: scripts/gen_libm_coeffs.py.
// Accuracy (tests/test_prelude_math_host.py, _gpu.py): sqrt <= 0.5 ulp + 2^-20,
// log <= 1 ulp, cos, acos <= 1.5 ulp, pow <= 2 + 1.5 |y ln x| ulp.
// ---------------------------------------------------------------------------

// sqrt: reciprocal-square-root seed (>= 20 bits), one coupled Newton step on
// (g, h) ~ (sqrt x, 1 / (2 sqrt x)), one residual correction. x = 0, +inf and
// denormals (seed inf / 0 / inf) return x itself; x < 0 gives NaN.
__device__ __forceinline__ double mkb_sqrt(double x) {
    double y;
    MKB_ASM_RSQRT64(y, x);
    double g = x * y;
    double h = 0.5 * y;
    const double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    const double d = fma(-g, g, x);
    g = fma(d, h, g);
    const unsigned int u = (unsigned int)__double2hiint(y) << 1;
    return (u == 0xffe00000u || u == 0u) ? x : g;
}

__constant__ double mkb_libm_c[] = {
    // [0..6] log: R(z) / z = 2 (A0 + A1 z + ...), log(1 + f) = f - f^2 / 2 + s (f^2 / 2 + R)
    0x1.5555555555558p-1, 0x1.99999999952d7p-2, 0x1.2492492df281ap-2, 0x1.c71c62e3f11e6p-3,
    0x1.7462b51cb66b1p-3, 0x1.39fe51a7c18f9p-3, 0x1.2b5900de53b32p-3,
    // [7] ln2 head (32 bits), [8] ln2 tail
    0x1.62e42fee00000p-1, 0x1.a39ef35793c76p-33,
    // [9] spare
    0.0,
    // [10..15] cos kernel (even quadrants): cos r = 1 - z / 2 + z^2 C(z)
    0x1.5555555555555p-5, -0x1.6c16c16c16963p-10, 0x1.a01a019f4ddfdp-16,
    -0x1.27e4fa16cc2d1p-22, 0x1.1eeb67d026ec6p-29, -0x1.907cd8ff5ede7p-37,
    // [16..21] sin kernel (odd quadrants): sin r = r + r z S(z)
    -0x1.5555555555555p-3, 0x1.1111111110babp-7, -0x1.a01a019e820aep-13,
    0x1.71de37946ed6dp-19, -0x1.ae6008d114485p-26, 0x1.5e0a4fc16be59p-33,
    // [22] 2 / pi, [23..25] pi / 2 as a sum of three doubles
    0x1.45f306dc9c883p-1, 0x1.921fb54442d18p+0, 0x1.1a62633145c07p-54, -0x1.f1976b7ed8fbcp-110,
    // [26..38] asin kernel: asin s = s + s z R(z), z = s^2 <= 1 / 4
    0x1.5555555555556p-3, 0x1.3333333332ec4p-4, 0x1.6db6db6e3292ep-5, 0x1.f1c71c1d4aa0dp-6,
    0x1.6e8bb1d89492bp-6, 0x1.1c4d343fa57f6p-6, 0x1.c9cf3759fd8ffp-7, 0x1.78246f32d075cp-7,
    0x1.524ea8b3aad31p-7, 0x1.653a8f8cfd43cp-8, 0x1.1d661531b222ep-6, -0x1.e7a24f548c99fp-7,
    0x1.d78189767524ep-6,
    // [39] spare
    0.0,
    // acos = (c0_hi + (c1 p + c0_lo)), index = 2 [|x| > 1/2] + [x < 0]:
    // [40..43] c1, [44..47] c0_hi, [48..51] c0_lo
    -1.0, 1.0, 2.0, -2.0,
    0x1.921fb54442d18p+0, 0x1.921fb54442d18p+0, 0.0, 0x1.921fb54442d18p+1,
    0x1.1a62633145c07p-54, 0x1.1a62633145c07p-54, 0.0, 0x1.1a62633145c07p-53,
};

// log: x = m 2^e with m in [sqrt(1/2), sqrt(2)), s = f / (2 + f), f = m - 1;
// the scheme of fdlibm's e_log.c with this file's division and coefficients.
// x < 0 and NaN give NaN, +inf gives +inf, 0 and denormals give -inf.
__device__ __forceinline__ double mkb_log(double x) {
    const int hx = __double2hiint(x);
    const int ha = hx + (0x3ff00000 - 0x3fe6a09e);
    const int e = (ha >> 20) - 1023;
    const double m = __hiloint2double((ha & 0x000fffff) + 0x3fe6a09e, __double2loint(x));
    const double f = m - 1.0;
    const double s = mkb_div(f, 2.0 + f);
    const double z = s * s;
    double p = mkb_libm_c[6];
#pragma unroll
    for (int k = 5; k >= 0; k--) p = fma(p, z, mkb_libm_c[k]);
    const double hfsq = (0.5 * f) * f;
    // (e as a double without a conversion instruction: 2^52 + 2^31 + e)
    const double dk = __hiloint2double(0x43300000, e ^ 0x80000000) - 4503601774854144.0;
    double t = fma(z, p, hfsq);                 // hfsq + R
    t = fma(s, t, dk * mkb_libm_c[8]);
    t = f - (hfsq - t);
    double res = fma(dk, mkb_libm_c[7], t);
    const unsigned int ux = (unsigned int)hx;
    // not a positive normal number
    const double odd = fma(x, __hiloint2double(0x7ff00000, 0), __hiloint2double(0x7ff00000, 0));
    if (ux - 0x00100000u >= 0x7fe00000u) res = odd;         // +inf -> +inf, negative / NaN -> NaN
    if ((ux << 1) < 0x00200000u) res = __hiloint2double(0xfff00000, 0);     // +-0, denormals -> -inf
    return res;
}

// cos: n = rint(x 2 / pi), r = x - n pi / 2 with pi / 2 in three parts
// (three fused multiply-adds), then the sin or cos kernel on |r| <= pi / 4
// chosen by the parity of n with an indexed constant load, sign from n mod 4.
// Valid for |x| < 2^31 (libdevice switches to Payne-Hanek above 1e5: the
// three-part reduction stays below 1 ulp far beyond that); inf / NaN -> NaN.
__device__ __forceinline__ double mkb_cos(double x) {
    double t = fma(x, mkb_libm_c[22], 6755399441055744.0);
    const int j = __double2loint(t);
    t -= 6755399441055744.0;
    double r = fma(-t, mkb_libm_c[23], x);
    r = fma(-t, mkb_libm_c[24], r);
    r = fma(-t, mkb_libm_c[25], r);
    const double z = r * r;
    const bool odd = (j & 1) != 0;
    const double* c = mkb_libm_c + (odd ? 16 : 10);
    double p = c[5];
#pragma unroll
    for (int k = 4; k >= 0; k--) p = fma(p, z, c[k]);
    const double a = odd ? r : z;
    const double b = odd ? r : fma(z, -0.5, 1.0);
    const double res = fma(a * z, p, b);
    // cos x = cos r, -sin r, -cos r, sin r for n mod 4 = 0, 1, 2, 3
    return __hiloint2double(__double2hiint(res) ^ (((j + 1) & 2) << 30), __double2loint(res));
}

// acos: for |x| <= 1/2, pi / 2 - asin x; beyond, 2 asin(sqrt((1 - |x|) / 2))
// (from pi for negative x), with one polynomial for asin on [0, 1/2]. Both
// reductions are computed and selected; |x| > 1 gives NaN through the sqrt.
__device__ __forceinline__ double mkb_acos(double x) {
    const int hx = __double2hiint(x);
    const double a = fabs(x);
    const bool big = (unsigned int)(hx & 0x7fffffff) >= 0x3fe00000u;    // |x| >= 1/2
    const double zb = fma(a, -0.5, 0.5);
    const double sb = mkb_sqrt(zb);
    const double z = big ? zb : a * a;
    const double s = big ? sb : a;
    double p = mkb_libm_c[38];
#pragma unroll
    for (int k = 37; k >= 26; k--) p = fma(p, z, mkb_libm_c[k]);
    p = fma(s * z, p, s);                       // asin(s)
    const int idx = (big ? 2 : 0) + (hx < 0 ? 1 : 0);
    return mkb_libm_c[44 + idx] + fma(mkb_libm_c[40 + idx], p, mkb_libm_c[48 + idx]);
}

// pow for exponents that are not small integer literals: exp(y log x), the
// product clamped to +-1024 so that 0^y and overflow come out as 0 / inf.
// Error <= 2 + 1.5 |y ln x| ulp (OpenCL allows its pow 16 ulp). A negative base
// gives NaN also for integer-valued y (integer literals never get here: they
// are multiplication chains), and 0^0 is NaN.
#ifndef MKB_EXP_FN
#define MKB_EXP_FN mkb_exp_poly
#endif
#ifndef MKB_FAST_LIBM
#define MKB_FAST_LIBM 0
#endif
__device__ __forceinline__ double mkb_pow(double x, double y) {
    double a = y * mkb_log(x);
    const int ha = __double2hiint(a);
    if ((unsigned int)((ha & 0x7fffffff) - 0x40900000) <= 0x3f600000u)  // 1024 <= |a| <= inf
        a = __hiloint2double((ha & 0x80000000) | 0x40900000, 0);
    return MKB_EXP_FN(a);
}

// x^(K + 1/2) = x^K * sqrt(x)
__device__ __forceinline__ float mkb_sqrt(float x) { return sqrtf(x); }
template <int K>
__device__ __forceinline__ Real mkb_powh(Real x) {
#if MKB_FAST_LIBM
    return mkb_powi<K>(x) * mkb_sqrt(x);
#else
    return mkb_powi<K>(x) * sqrt(x);
#endif
}
"""


# Unary double-precision routines of the prelude (tests/prelude_math.py)
PRELUDE_UNARY = ['mkb_sqrt', 'mkb_log', 'mkb_cos', 'mkb_acos']

_VECTOR_PRELUDE = r"""
// N consecutive Reals as one 8/16-byte access (two for 32 bytes). The address
// is aligned: planes start 128-byte aligned, nx and the thread's first cell
// are multiples of N.
template <int N>
__device__ __forceinline__ void mkb_vload(Real (&d)[N], const Real* p, bool ok) {
    if (!ok) {
#pragma unroll
        for (int c = 0; c < N; c++) d[c] = (Real)0;
        return;
    }
    if constexpr (sizeof(Real) * N == 8) {
        const float2 t = *reinterpret_cast<const float2*>(p);
        d[0] = t.x; d[1] = t.y;
    } else if constexpr (sizeof(Real) == 4) {
#pragma unroll
        for (int c = 0; c < N; c += 4) {
            const float4 t = *reinterpret_cast<const float4*>(p + c);
            d[c] = t.x; d[c + 1] = t.y; d[c + 2] = t.z; d[c + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int c = 0; c < N; c += 2) {
            const double2 t = *reinterpret_cast<const double2*>(p + c);
            d[c] = t.x; d[c + 1] = t.y;
        }
    }
}
template <int N>
__device__ __forceinline__ void mkb_vstore(Real* p, const Real (&d)[N]) {
    if constexpr (sizeof(Real) * N == 8) {
        *reinterpret_cast<float2*>(p) = make_float2(d[0], d[1]);
    } else if constexpr (sizeof(Real) == 4) {
#pragma unroll
        for (int c = 0; c < N; c += 4) {
            *reinterpret_cast<float4*>(p + c) = make_float4(d[c], d[c + 1], d[c + 2], d[c + 3]);
        }
    } else {
#pragma unroll
        for (int c = 0; c < N; c += 2) {
            *reinterpret_cast<double2*>(p + c) = make_double2(d[c], d[c + 1]);
        }
    }
}
"""

# Output of scripts/gen_exp_coeffs.py (degree 11, c11 .. c2)
_EXP_COEFFS = [
    '0x1.af632a0f7e2cep-26', '0x1.28b4101c77212p-22', '0x1.71ddf56d8deb5p-19',
    '0x1.a01991a10d9aep-16', '0x1.a01a01b1461c5p-13', '0x1.6c16c1880029fp-10',
    '0x1.111111110f21ep-7', '0x1.555555554f0bap-5', '0x1.555555555555ap-3',
    '0x1.0000000000011p-1',
]
# 2^(j/64), j = 0..63, correctly rounded (scripts/gen_exp_coeffs.py)
_EXP_POW2 = [
    '0x1.0000000000000p+0', '0x1.02c9a3e778061p+0', '0x1.059b0d3158574p+0', '0x1.0874518759bc8p+0',
    '0x1.0b5586cf9890fp+0', '0x1.0e3ec32d3d1a2p+0', '0x1.11301d0125b51p+0', '0x1.1429aaea92de0p+0',
    '0x1.172b83c7d517bp+0', '0x1.1a35beb6fcb75p+0', '0x1.1d4873168b9aap+0', '0x1.2063b88628cd6p+0',
    '0x1.2387a6e756238p+0', '0x1.26b4565e27cddp+0', '0x1.29e9df51fdee1p+0', '0x1.2d285a6e4030bp+0',
    '0x1.306fe0a31b715p+0', '0x1.33c08b26416ffp+0', '0x1.371a7373aa9cbp+0', '0x1.3a7db34e59ff7p+0',
    '0x1.3dea64c123422p+0', '0x1.4160a21f72e2ap+0', '0x1.44e086061892dp+0', '0x1.486a2b5c13cd0p+0',
    '0x1.4bfdad5362a27p+0', '0x1.4f9b2769d2ca7p+0', '0x1.5342b569d4f82p+0', '0x1.56f4736b527dap+0',
    '0x1.5ab07dd485429p+0', '0x1.5e76f15ad2148p+0', '0x1.6247eb03a5585p+0', '0x1.6623882552225p+0',
    '0x1.6a09e667f3bcdp+0', '0x1.6dfb23c651a2fp+0', '0x1.71f75e8ec5f74p+0', '0x1.75feb564267c9p+0',
    '0x1.7a11473eb0187p+0', '0x1.7e2f336cf4e62p+0', '0x1.82589994cce13p+0', '0x1.868d99b4492edp+0',
    '0x1.8ace5422aa0dbp+0', '0x1.8f1ae99157736p+0', '0x1.93737b0cdc5e5p+0', '0x1.97d829fde4e50p+0',
    '0x1.9c49182a3f090p+0', '0x1.a0c667b5de565p+0', '0x1.a5503b23e255dp+0', '0x1.a9e6b5579fdbfp+0',
    '0x1.ae89f995ad3adp+0', '0x1.b33a2b84f15fbp+0', '0x1.b7f76f2fb5e47p+0', '0x1.bcc1e904bc1d2p+0',
    '0x1.c199bdd85529cp+0', '0x1.c67f12e57d14bp+0', '0x1.cb720dcef9069p+0', '0x1.d072d4a07897cp+0',
    '0x1.d5818dcfba487p+0', '0x1.da9e603db3285p+0', '0x1.dfc97337b9b5fp+0', '0x1.e502ee78b3ff6p+0',
    '0x1.ea4afa2a490dap+0', '0x1.efa1bee615a27p+0', '0x1.f50765b6e4540p+0', '0x1.fa7c1819e90d8p+0',
]
# Table variant: 64 / ln2, -ln2/64 (hi, lo), q3..q0 of (e^r - 1 - r) / r^2
_EXP_TCOEF = [
    ('0x1.71547652b82fep+6', '64 / ln2'),
    ('0x1.8p+52', '1.5 * 2^52'),
    ('-0x1.62e42fefa39efp-7', '-ln2 / 64, high part'),
    ('-0x1.abc9e3b39803fp-62', '-ln2 / 64, low part'),
    ('0x1.11111d90623e3p-7', 'q3'),
    ('0x1.55556b342389bp-5', 'q2'),
    ('0x1.5555555555255p-3', 'q1'),
    ('0x1.ffffffffff57dp-2', 'q0'),
]
_EXP_TABLE = [
    ('0x1.71547652b82fep+0', 'log2(e)'),
    ('0x1.8p+52', '1.5 * 2^52: round-to-nearest-integer shift'),
    ('-0x1.62e42fefa39efp-1', '-ln2, high part'),
    ('-0x1.abc9e3b39803fp-56', '-ln2, low part'),
] + [(c, 'c%d' % (11 - i)) for i, c in enumerate(_EXP_COEFFS)]
_PRELUDE = _PRELUDE.replace('@EXP_TABLE@', '\n'.join(
    '    %r,  // %s' % (float.fromhex(h), name) for h, name in _EXP_TABLE))
_PRELUDE = _PRELUDE.replace('@EXP_TCOEF@', '\n'.join(
    '    %r,  // %s' % (float.fromhex(h), name) for h, name in _EXP_TCOEF))
_PRELUDE = _PRELUDE.replace('@EXP_POW2@', '\n'.join(
    '    %r,' % float.fromhex(h) for h in _EXP_POW2))


def default_block(nx, ny, precision, n_state=0):
    """Thread-block tile ``(bx, by)`` for a grid of ``nx * ny`` cells."""
    if ny <= 1:
        return (128, 1) if nx <= 4096 else (256, 1)
    if nx >= 256 and ny >= 2 and n_state > 16 \
            and precision == myokit.DOUBLE_PRECISION:
        # (C3 sweep: 128 x 2 is 1.5 % faster than 64 x 4 for the large model)
        return (128, 2)
    if nx >= 64:
        return (64, 4)
    return (32, 8) if ny >= 8 else (32, 4)


def default_options(precision, n_state):
    """
    Generator defaults, from sweeps on a B200 (profiles/): double precision
    uses the in-line division, exp and branch-free sqrt / log / cos / acos /
    pow, and for large models asks for two resident 256-thread CTAs per SM
    (128 registers per thread), prefetches every state plane into L1 at the
    top of the kernel and loads each state 24 equations ahead of its use.
    """
    if precision == myokit.SINGLE_PRECISION:
        # __fdividef: 2 ulp, inside the 2.5 ulp OpenCL allows its single-
        # precision division (the reference's arithmetic contract for fp32)
        return dict(min_blocks=None, fast_div=True, fast_exp=False,
                    load_ahead=32, stream=n_state <= 4,
                    cells_per_thread=4 if n_state <= 4 else 1,
                    rows_per_thread=4 if n_state <= 4 else 1)
    big = n_state > 16
    return dict(min_blocks=2 if big else None, fast_div=True,
                fast_exp='poly', div_cubic=True, fast_libm=True,
                # (measured on C3, profiles/r02_sweeps.md: conditionals as
                # selects make the whole body one scheduling region, which
                # ptxas fills until it spills 430 bytes per thread at 128
                # registers — slower than leaving the model's own branches)
                select=False,
                load_ahead=24 if big else 32,
                prefetch='l1' if big else None,
                slab_lean=True,
                # large models: every state plane's tile through shared memory
                # by TMA where the grid allows (profiles/r03_sweeps.md: 0.707
                # -> 0.652 ms on C3; SimulationCUDA.kernel_source then also
                # loads 4 equations ahead and takes select='cheap')
                stage=big,
                stage_group=(8,),
                # ... and without a V tile: neighbours from the adjacent lanes
                # and from memory (0.647 -> 0.638 ms)
                v_direct=big,
                # small models: vector path, TMA-fed where the grid allows
                # (stencil-only on 8192 x 4096: 92-97 % of the measured copy
                # bandwidth in double precision against 69 % without)
                stream=n_state <= 4,
                cells_per_thread=2 if n_state <= 4 else 1,
                rows_per_thread=4 if n_state <= 4 else 1)


# Shared memory two resident thread blocks of a staged kernel may take together
# (227 KiB per SM less the V tile, the exp table and the per-block reserve)
STAGE_SMEM_LIMIT = 106 * 1024


def stage_applies(precision, n_state, block, diffusion_mode, nx, cells_per_thread=1,
                  persistent=False, split_gates=False, junction=None,
                  debug_mem=None, fast_exp=None, lazy_state=True, **ignored):
    """
    True if :func:`generate` can stage the states of this configuration in
    shared memory (option ``stage``): scalar path on a regular grid or
    uncoupled cells, rows and tile rows that are 16-byte multiples (the TMA
    descriptor's strides and box), and a tile of all state planes small enough
    for two resident thread blocks per SM.
    """
    rs = 4 if precision == myokit.SINGLE_PRECISION else 8
    bx, by = block
    return bool(
        (cells_per_thread or 1) == 1 and lazy_state
        and not (persistent or split_gates or junction or debug_mem)
        and fast_exp != 'stab'
        and diffusion_mode != DIFF_CONNECTIONS
        and (nx * rs) % 16 == 0 and (bx * rs) % 16 == 0
        and (bx * by * rs) % 128 == 0 and bx <= 256 and by <= 256
        and n_state * bx * by * rs + 128 <= STAGE_SMEM_LIMIT)


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
        self.persistent = False
        self.gate_kernel = False        # a second kernel, mkb_gate_step
        self.kernel_flags = 0           # MKB_KERNEL_* of include/myokit_b200.h
        self.stream_box = (0, 0)        # TMA box (cells, rows) of a streaming kernel
        self.gate_states = []
        self.cells_per_thread = 1
        self.rows_per_thread = 1
        self.plane_stride = 0           # elements; 0: any (read from MkbGridArgs)
        self.smem_bytes = 0             # dynamic shared memory per block (staged kernels)

    def key(self):
        h = hashlib.sha256()
        h.update(self.code.encode('utf-8'))
        h.update('\0'.join(self.options).encode('utf-8'))
        return h.hexdigest()


def _emit_tile_loop(p, L):
    """
    The staged kernel as a loop over tiles (option ``tile_loop``): a fixed grid
    of ``SMs x blocks per SM`` thread blocks, block b taking tiles b, b + grid,
    ... in launch order. What a one-tile block waits for at its start — its
    first loads from HBM, about a third of a warp's residence (ncu,
    profiles/r03_summary.md) — is requested one tile ahead instead:

    * while tile i is computed, the membrane potentials of tile i + grid (own
      cell and rim, per thread, LDGSTS) and the ``stage_early`` state planes
      its equations need first (TMA) travel into the second of two small
      buffers;
    * the other state planes of tile i + grid are requested as soon as the
      TMA stores of tile i have read the main buffer, and arrive while the
      block works on the early planes;
    * the diffusion current is formed where the equations first use it (the
      update of V, at the end), so its conductances are loaded late and never
      waited for at the top.

    Same arithmetic in the same order as the one-tile kernel: same bits.
    ``L`` is the local namespace of :func:`generate`.
    """
    bx, by = L['bx'], L['by']
    stage_slot, stage_sizes = L['stage_slot'], L['stage_sizes']
    diffusion_mode, diffusion = L['diffusion_mode'], L['diffusion']
    fields, consts, body = L['fields'], L['consts'], L['body']
    plane_stride, min_blocks = L['plane_stride'], L['min_blocks']
    paced_list, i_vm, n_state = L['paced_list'], L['i_vm'], L['n_state']
    load_ahead = max(int(L['load_ahead']), 0)
    n_slots = len(stage_slot)
    ne = min(stage_sizes[0], n_slots)
    nm = n_slots - ne
    grid = diffusion_mode in (DIFF_HOMOGENEOUS, DIFF_FIELD)
    p('#define MKB_STAGE_EARLY %d' % ne)
    p('#define MKB_STAGE_AT(j) (*((j) < MKB_STAGE_EARLY ? early_c + (j) * MKB_STAGE_TILE'
      ' : main_c + ((j) - MKB_STAGE_EARLY) * MKB_STAGE_TILE))')
    p('#define MKB_STAGE_WAIT(k) do { if ((k) == 0) MKB_MBAR_WAIT(&stage_bar[1 + pb_], pe_);'
      ' else MKB_MBAR_WAIT(&stage_bar[0], pm_); } while (0)')
    p('')
    p('extern "C" __global__ void __launch_bounds__(MKB_BX * MKB_BY, %d)' % int(min_blocks or 2))
    p('%s(const __grid_constant__ MkbGridArgs g, const MkbStepParams* __restrict__ sp,' % KERNEL_NAME)
    p('    const Real* __restrict__ v_in, Real* __restrict__ v_out)')
    p('{')
    p('    const unsigned int tx = threadIdx.x, ty = threadIdx.y;')
    p('    const unsigned int nx = (unsigned int)g.nx, ny = (unsigned int)g.ny;')
    if plane_stride:
        p('    constexpr unsigned long long stride = %dull;' % int(plane_stride))
        p('    if (g.stride != stride) { MKB_STRIDE_MISMATCH(); return; }')
    else:
        p('    const unsigned long long stride = g.stride;')
    p('    const unsigned int nbx = (nx + MKB_BX - 1) / MKB_BX;')
    p('    const unsigned int ntiles = nbx * ((ny + MKB_BY - 1) / MKB_BY);')
    p('    if (blockIdx.x >= ntiles) return;')
    p('    const unsigned int t_ = ty * MKB_BX + tx;')
    p('    const bool lane0_ = (t_ & 31u) == 0u;')
    p('    Real* const state = (Real*)g.state;')
    p('    // Per-step scalars, cast like openclsim.c:1063,1148,1155')
    p('    const Real time = (Real)sp->time;')
    p('    const Real dt = (Real)sp->dt;')
    p('    const Real pace_in = (Real)sp->pace;')
    p('    const bool store_aux_ = (sp->flags & MKB_FLAG_STORE_AUX) != 0;')
    p('    (void)time; (void)pace_in; (void)store_aux_; (void)v_in; (void)v_out; (void)stride; (void)state;')
    p('    // Shared memory: arrival barriers (main, early[2]); the main buffer (%d' % nm)
    p('    // state planes of this tile); two early buffers (%d planes each: this' % ne)
    p('    // tile\'s and the next one\'s); two V tiles with their rims.')
    p('    MKB_STAGE_DECL(MKB_LOOP_BYTES);')
    p('    unsigned long long* const stage_bar = (unsigned long long*)mkb_stage_mem;')
    p('    Real* const main_base = (Real*)(mkb_stage_mem + 128);')
    p('    Real* const early_base = main_base + %d * MKB_STAGE_TILE;' % nm)
    if grid:
        p('    __shared__ Real tile2[2][MKB_BY + 2][MKB_BX + 2];')
    p('    if (t_ == 0) {')
    p('        for (int k_ = 0; k_ < 3; k_++) MKB_MBAR_INIT(&stage_bar[k_], MKB_STAGE_WARPS);')
    p('        MKB_MBAR_FENCE_INIT();')
    p('    }')
    p('    __syncthreads();')

    def request_early(tile, buf):
        # V of the tile (own cell and rim) by per-thread asynchronous copies,
        # its early planes by TMA, into buffer `buf`
        p('        {')
        p('            const unsigned int byr_ = %s / nbx, bxb_ = %s - byr_ * nbx;' % (tile, tile))
        if grid:
            p('            const unsigned int ix_ = bxb_ * MKB_BX + tx, iy_ = byr_ * MKB_BY + ty;')
            p('            if (ix_ < nx && iy_ < ny) {')
            p('                const Real* const vs_ = v_in + ((unsigned long long)iy_ * nx + ix_);')
            p('                Real (*const tl_)[MKB_BX + 2] = tile2[%s];' % buf)
            p('                MKB_CP_ASYNC(&tl_[ty + 1][tx + 1], vs_);')
            p('                // (a rim slot outside the grid takes the cell\'s own V and is never used)')
            p('                if (tx == 0) MKB_CP_ASYNC(&tl_[ty + 1][0], (ix_ > 0) ? vs_ - 1 : vs_);')
            p('                if (tx == MKB_BX - 1 || ix_ == nx - 1)')
            p('                    MKB_CP_ASYNC(&tl_[ty + 1][tx + 2], (ix_ < nx - 1) ? vs_ + 1 : vs_);')
            p('                if (ty == 0) {')
            p('                    const Real* src_ = vs_;')
            p('                    if (iy_ > 0) src_ = vs_ - nx;')
            p('                    else if (g.iy_offset > 0 && g.halo_lo) src_ = (const Real*)g.halo_lo + ix_;')
            p('                    MKB_CP_ASYNC(&tl_[0][tx + 1], src_);')
            p('                }')
            p('                if (ty == MKB_BY - 1 || iy_ == ny - 1) {')
            p('                    const Real* src_ = vs_;')
            p('                    if (iy_ < ny - 1) src_ = vs_ + nx;')
            p('                    else if (g.iy_offset + iy_ < g.ny_global - 1 && g.halo_hi) src_ = (const Real*)g.halo_hi + ix_;')
            p('                    MKB_CP_ASYNC(&tl_[ty + 2][tx + 1], src_);')
            p('                }')
            p('            }')
        p('            if (lane0_) {')
        p('                unsigned int n_ = 0;')
        p('                for (unsigned int j_ = t_ >> 5; j_ < %du; j_ += MKB_STAGE_WARPS) n_++;' % ne)
        p('                MKB_MBAR_EXPECT_TX(&stage_bar[1 + %s], n_ * MKB_STAGE_TILE * (unsigned int)sizeof(Real));' % buf)
        p('                for (unsigned int j_ = t_ >> 5; j_ < %du; j_ += MKB_STAGE_WARPS)' % ne)
        p('                    MKB_TMA_LOAD_3D(early_base + (%s * %du + j_) * MKB_STAGE_TILE, g.tmap_state,' % (buf, ne))
        p('                                    (int)(bxb_ * MKB_BX), (int)(byr_ * MKB_BY), (int)mkb_stage_plane[j_],')
        p('                                    &stage_bar[1 + %s]);' % buf)
        p('            }')
        p('        }')

    p('    // the first tile\'s early data')
    p('    {')
    request_early('blockIdx.x', '0u')
    p('    }')
    p('    unsigned int it_ = 0;')
    p('    for (unsigned int tile_ = blockIdx.x; tile_ < ntiles; tile_ += gridDim.x, it_++) {')
    p('        // buffer and barrier phases of this tile')
    p('        const unsigned int pb_ = it_ & 1u, pm_ = it_ & 1u, pe_ = (it_ >> 1) & 1u;')
    p('        const unsigned int byr = tile_ / nbx, bxb = tile_ - byr * nbx;')
    p('        const unsigned int ix = bxb * MKB_BX + tx, iy = byr * MKB_BY + ty;')
    p('        const bool active = (ix < nx) && (iy < ny);')
    p('        const unsigned long long cid = (unsigned long long)iy * nx + ix;')
    p('        Real* const state_c = state + cid;')
    p('        const Real* const field_c = (const Real*)g.field + cid;')
    p('        Real* const inter_c = (Real*)g.inter + cid;')
    p('        (void)state_c; (void)field_c; (void)inter_c;')
    p('        Real* const main_c = main_base + t_;')
    p('        Real* const early_c = early_base + pb_ * %du * MKB_STAGE_TILE + t_;' % ne)
    p('        (void)main_c; (void)early_c;')
    p('        // This thread\'s copies of V have landed; the stores of the previous tile')
    p('        // have read the buffers; then the block meets.')
    p('        MKB_CP_ASYNC_WAIT();')
    p('        if (lane0_) MKB_TMA_STORE_READ_WAIT();')
    p('        __syncthreads();')
    p('        if (lane0_) {')
    p('            unsigned int n_ = 0;')
    p('            for (unsigned int j_ = %du + (t_ >> 5); j_ < %du; j_ += MKB_STAGE_WARPS) n_++;' % (ne, n_slots))
    p('            MKB_MBAR_EXPECT_TX(&stage_bar[0], n_ * MKB_STAGE_TILE * (unsigned int)sizeof(Real));')
    p('            for (unsigned int j_ = %du + (t_ >> 5); j_ < %du; j_ += MKB_STAGE_WARPS)' % (ne, n_slots))
    p('                MKB_TMA_LOAD_3D(main_base + (j_ - %du) * MKB_STAGE_TILE, g.tmap_state,' % ne)
    p('                                (int)(bxb * MKB_BX), (int)(byr * MKB_BY), (int)mkb_stage_plane[j_],')
    p('                                &stage_bar[0]);')
    p('        }')
    p('        // the next tile\'s early data, into the other buffers')
    p('        if (tile_ + gridDim.x < ntiles) {')
    request_early('(tile_ + gridDim.x)', '(pb_ ^ 1u)')
    p('        }')
    p('        // Every thread takes the step, also those of a rim tile that lie outside')
    p('        // the grid (their states arrive as zeros, their results are clipped, their')
    p('        // accesses to global memory are predicated): the body stays uniform')
    p('        // control flow, in which ptxas keeps the constants in uniform registers —')
    p('        // inside `if (active)` it copies them to ordinary ones, some twenty')
    p('        // registers that the schedule of the FP64 chains then lacks.')
    p('        {')
    p('        const bool store_aux = store_aux_ && active;')
    p('        (void)store_aux;')
    kernel_p = p
    active_lines = []
    p = active_lines.append
    if grid:
        p('        Real (*const tile)[MKB_BX + 2] = tile2[pb_];')
        p('        const Real vc = tile[ty + 1][tx + 1];')
        p('        const unsigned int iyg = iy + (unsigned int)g.iy_offset;  // global row')
        p('        const unsigned int nyg = (unsigned int)g.ny_global;')
        p('        (void)nyg;')
    for k, var in enumerate(fields):
        pass
    for line in L['early']:
        p('    ' + line)
    if diffusion:
        p('        // openclsim.cl:249-280, 322-329')
        if paced_list:
            p('        const Real pace = (active && g.paced_mask[cid]) ? pace_in : (Real)0;')
        else:
            p('        const int pix = (int)ix, piy = (int)iyg;')
            p('        const Real pace = (pix >= (int)g.pace_x0 && pix < (int)g.pace_x1 &&')
            p('                           piy >= (int)g.pace_y0 && piy < (int)g.pace_y1) ? pace_in : (Real)0;')
    else:
        p('        const Real pace = pace_in;')
    p('        (void)pace;')
    for line in consts:
        p('    ' + line)

    # the diffusion current, where the equations first need it; its
    # conductances are requested `cond_ahead` statements earlier
    cond_ahead = 40
    dloads = []
    dblock = []
    d = dblock.append
    if grid:
        d('        // Diffusion current of this cell (formed here, where it is first used)')
        d('        Real idiff;')
        d('        {')
        d('            const Real vxm = tile[ty + 1][tx], vxp = tile[ty + 1][tx + 2];')
        d('            const Real vym = tile[ty][tx + 1], vyp = tile[ty + 2][tx + 1];')
        if diffusion_mode == DIFF_HOMOGENEOUS:
            d('            // openclsim.cl:401-434 (diff_step)')
            d('            const Real gx = (Real)g.gx, gy = (Real)g.gy;')
            d('            if (nx > 1) {')
            d('                if (ix == 0) idiff = gx * (vc - vxp);')
            d('                else if (ix == nx - 1) idiff = gx * (vc - vxm);')
            d('                else idiff = gx * (2 * vc - vxm - vxp);')
            d('            } else {')
            d('                idiff = 0;')
            d('            }')
            d('            if (nyg > 1) {')
            d('                if (iyg == 0) idiff += gy * (vc - vyp);')
            d('                else if (iyg == nyg - 1) idiff += gy * (vc - vym);')
            d('                else idiff += gy * (2 * vc - vym - vyp);')
            d('            }')
        else:
            d('            // openclsim.cl:469-486 (diff_hetero)')
            dl = dloads.append
            dl('        // Edge conductances gx[(ny, nx-1)], gy[(ny-1, nx)], ahead of the diffusion current')
            dl('        const Real* const gxf = (const Real*)g.gx_field;')
            dl('        const Real* const gyf = (const Real*)g.gy_field;')
            dl('        const bool has_xm = active && nx > 1 && ix > 0, has_xp = active && nx > 1 && ix < nx - 1;')
            dl('        const bool has_ym = active && nyg > 1 && iyg > 0, has_yp = active && nyg > 1 && iyg < nyg - 1;')
            dl('        const Real gxm = has_xm ? gxf[cid - iy - 1] : (Real)0;')
            dl('        const Real gxp = has_xp ? gxf[cid - iy] : (Real)0;')
            dl('        const Real gym = has_ym ? gyf[(long long)cid - (long long)nx] : (Real)0;')
            dl('        const Real gyp = has_yp ? gyf[cid] : (Real)0;')
            d('            idiff = 0.0;')
            d('            if (has_xm) { idiff += gxm * (vc - vxm); }')
            d('            if (has_xp) { idiff += gxp * (vc - vxp); }')
            d('            if (has_ym) idiff += gym * (vc - vym);')
            d('            if (has_yp) idiff += gyp * (vc - vyp);')
        d('        }')
        d('        if (store_aux) ((Real*)g.idiff)[cid] = idiff;')
    first = None
    if dblock:
        import re as _re
        for i, line in enumerate(body):
            if _re.search(r'\bidiff\b', line):
                first = i
                break
        if first is None:
            first = len(body)
    for i, line in enumerate(body):
        if dloads and i == max(first - cond_ahead, 0):
            for x in dloads:
                p(x)
        if dblock and i == first:
            for x in dblock:
                p(x)
        if line == '@PREFETCH_NEXT@':
            continue
        for x in line.split('\n'):
            if x.startswith('    v_out[cid] = '):
                x = '    if (active) ' + x.strip()
            p('    ' + x)
    if dblock and first == len(body):
        for x in dloads + dblock:
            p(x)
    p = kernel_p
    if L.get('tile_call'):
        p('            mkb_tile_body(g, tx, ty, ix, iy, active, pb_, pe_, pm_, time, dt, pace_in, store_aux,')
        p('                          v_out, %s, main_c, early_c);' % ('&tile2[0][0][0]' if grid else '(Real*)0'))
    else:
        for x in active_lines:
            p(x)
    p('        }   // active')
    p('        // The tile goes back: stores visible to the TMA unit, the block meets,')
    p('        // the first lane of every warp issues its share of the boxes.')
    p('        MKB_FENCE_ASYNC_SMEM();')
    p('        __syncthreads();')
    p('        if (lane0_) {')
    p('            for (unsigned int j_ = t_ >> 5; j_ < %du; j_ += MKB_STAGE_WARPS)' % n_slots)
    p('                MKB_TMA_STORE_3D(g.tmap_state, (int)(bxb * MKB_BX), (int)(byr * MKB_BY), (int)mkb_stage_plane[j_],')
    p('                    (j_ < %du) ? early_base + (pb_ * %du + j_) * MKB_STAGE_TILE' % (ne, ne))
    p('                               : main_base + (j_ - %du) * MKB_STAGE_TILE);' % ne)
    p('            MKB_TMA_STORE_COMMIT();')
    p('        }')
    p('    }')
    p('    if (lane0_) MKB_TMA_STORE_READ_WAIT();')
    p('}')
    p('')
    if not L.get('tile_call'):
        return []
    # One cell's step as a function of its own: ptxas schedules a loop body
    # with far fewer independent chains in flight than straight-line code
    # (r03_tile_loop.md: 36 % of the FP64 instructions two or fewer
    # instructions behind their producer, against 20 %).
    F = []
    q = F.append
    q('static __device__ __noinline__ void mkb_tile_body(')
    q('    const MkbGridArgs& g, const unsigned int tx, const unsigned int ty,')
    q('    const unsigned int ix, const unsigned int iy, const bool active, const unsigned int pb_,')
    q('    const unsigned int pe_, const unsigned int pm_, const Real time, const Real dt,')
    q('    const Real pace_in, const bool store_aux, Real* __restrict__ v_out, Real* const tile2_,')
    q('    Real* __restrict__ const main_c, Real* __restrict__ const early_c)')
    q('{')
    q('    const unsigned int nx = (unsigned int)g.nx, ny = (unsigned int)g.ny;')
    if plane_stride:
        q('    constexpr unsigned long long stride = %dull;' % int(plane_stride))
    else:
        q('    const unsigned long long stride = g.stride;')
    q('    (void)time; (void)pace_in; (void)store_aux; (void)stride; (void)ny; (void)tile2_;')
    q('    const unsigned int t_ = ty * MKB_BX + tx;')
    q('    const unsigned long long cid = (unsigned long long)iy * nx + ix;')
    q('    Real* const state = (Real*)g.state;')
    q('    Real* const state_c = state + cid;')
    q('    const Real* const field_c = (const Real*)g.field + cid;')
    q('    Real* const inter_c = (Real*)g.inter + cid;')
    q('    (void)state_c; (void)field_c; (void)inter_c;')
    q('    MKB_STAGE_DECL(MKB_LOOP_BYTES);')
    q('    unsigned long long* const stage_bar = (unsigned long long*)mkb_stage_mem;')
    q('    // (the two buffers never overlap: told to the compiler, which otherwise')
    q('    // keeps every shared-memory load behind every earlier store)')
    q('    (void)main_c; (void)early_c; (void)stage_bar; (void)t_;')
    if grid:
        q('    Real (*const tile2)[MKB_BY + 2][MKB_BX + 2] = (Real (*)[MKB_BY + 2][MKB_BX + 2])tile2_;')
    for x in active_lines:
        q(x)
    q('}')
    q('')
    return F


def generate(model, precision, bound_variables, inter_log, fields, rl_states,
             diffusion_mode, paced_list, block, native_maths=False, fmad=True,
             max_registers=None, pow_multiply=True, fast_div=False,
             lazy_state=True, min_blocks=None, fast_exp=False,
             const_pool=True, load_ahead=8, slab=False, cells_per_thread=1,
             rows_per_thread=1, div_int_check=False, partitioned=False,
             const_div=True, slab_lean=False, div_parallel=False,
             junction=None, persistent=False, split_gates=False,
             div_cubic=False, prefetch=None, debug_mem=None,
             fast_libm=False, select=False, exp_scale='mul', stream=False,
             overlap=False, plane_stride=None, stage=False,
             stage_group=8, stage_store=True, prefetch_next=None,
             tile_loop=False, stage_early=4, tile_call=False, v_direct=False):
    """
    Generates the fused cell-step kernel for a prepared ``model`` (bindings
    processed and unique names created, ``openclsim.py:284-290``).

    ``inter_log`` and ``fields`` are lists of variables in storage order;
    ``rl_states`` maps state -> (inf, tau); ``paced_list`` is True when paced
    cells are an explicit list (byte mask) instead of a rectangle.

    Code-shape options (none changes which operations are evaluated or their
    order, except ``pow_multiply`` / ``fast_div`` / ``fast_exp``, which swap a
    library routine for a cheaper one of equal or near-equal accuracy):

    ``lazy_state`` / ``load_ahead``
        Load each state ``load_ahead`` equations before its first use and
        store its new value as soon as it exists (short live ranges, loads
        still ahead of use), instead of all loads first and all stores last.
    ``min_blocks``
        Second argument of ``__launch_bounds__``: resident CTAs per SM the
        register allocation must allow.
    ``const_pool``
        Double-precision literals in a ``__constant__`` table.
    ``cells_per_thread``
        2 or 4: each thread owns that many x-adjacent cells, so every plane
        access is one 8/16/32-byte vector load or store and the index / halo
        bookkeeping is paid once per thread. For small models, where that
        bookkeeping is a large share of the instructions (the stencil-only
        kernel is otherwise instruction-, not HBM-bound). Needs
        ``nx % cells_per_thread == 0``; not combined with ``slab`` or
        connections.
    ``rows_per_thread``
        With ``cells_per_thread`` > 1: each thread also walks that many rows
        (more bytes in flight per thread, fewer and fatter thread blocks).
    ``const_div``
        With ``fast_div``: ``x / c`` for a compile-time constant ``c`` becomes
        ``x * (1 / c)`` (within 1 ulp of the quotient). A third of the
        divisions of a large model have constant divisors. Likewise the
        Rush-Larsen exponent ``-dt / tau`` with ``tau = c / X`` becomes
        ``-dt * X / c``.
    ``partitioned``
        Connection graphs cut over several GPUs: CSR columns beyond the local
        cells are ghost cells whose V is read from the ghost buffer the
        owning GPUs push into.
    ``split_gates``
        Two kernels per step instead of one: states whose update needs only V
        and the state itself (gating variables with voltage-dependent rates)
        get a kernel of their own, ``mkb_gate_step``, with few live values and
        therefore high occupancy; ``mkb_cell_step`` reads those states but
        neither computes their rate equations nor updates them. Same
        expressions, same results; the gates are read twice. For large models
        held at low occupancy by their register count. One cell per thread.
    ``persistent``
        For grids that fit one thread block (``nx <= bx`` and ``ny <= by``;
        the runtime checks): the kernel ``mkb_cell_step_persistent`` keeps
        every state in registers and takes all steps up to the next logged
        one inside a single launch — V travels through shared memory, two
        barriers per step, nothing goes to global memory in between. For
        small cables, where a step is otherwise bound by launch latency.
        One cell per thread; not for connection graphs, slabs or junctions.
    ``junction``
        ``'fiber'`` or ``'tissue'``: the kernel of one of two grids stepped in
        lockstep (``FiberTissueSimulationCUDA``); cells on the junction add
        the current to / from the other grid's cell to ``idiff`` after their
        own stencil, as ``diff_step_fiber_tissue`` does
        (``openclsim.cl:601-628``). Homogeneous 2-d grids, one cell per thread.
    ``prefetch``
        ``'l1'`` or ``'l2'``: every state that is loaded later in the body
        (``load_ahead``) is prefetched into that cache level at the top of the
        kernel, so the load proper finds it close by: the memory latency is
        covered without holding a register for the value.
    ``debug_mem``
        Diagnostic builds that give WRONG results, for measuring where the
        time of a step goes: ``'l1'`` reads every state but V from one of
        256 cells (always a cache hit: the kernel without its load latency),
        ``'l1ns'`` also predicates every state store off.
    ``fast_libm``
        In-line branch-free double-precision ``sqrt`` / ``log`` / ``cos`` /
        ``acos`` / general ``pow`` (``mkb_sqrt`` ... in the prelude) instead
        of libdevice's, whose internal branches cut the kernel body into
        short scheduling regions.
    ``select``
        Conditional expressions evaluate both arms and select (``True``), or
        only where no arm holds an exp / log / pow / trig call (``'cheap'``).
    ``exp_scale``
        ``'mul'`` (default): ``mkb_exp_*`` apply 2^n with a multiplication,
        +inf / 0 outside the double range; ``'add'``: exponent-field add,
        saturating.
    ``stream``
        Vector path fed by TMA: persistent blocks walk tiles, the V tile and
        its halo arrive through ``cp.async.bulk.tensor.2d`` into a two-stage
        mbarrier ring (2-d grids on one GPU whose rows are 16-byte multiples,
        ``block[0] == 32``).
    ``overlap``
        Consecutive steps overlap: programmatic dependent launch plus
        per-tile step counters (``MkbGridArgs::tile_done``) instead of the
        kernel boundary; scalar path on regular grids.
    ``plane_stride``
        The plane stride (elements) as a compile-time constant of the main
        kernel: plane addresses become base + constant. The kernel traps on
        any other ``MkbGridArgs::stride``; ``KernelSource.plane_stride``
        tells the runtime (``mkb_sim_config::kernel_stride``).
    ``stage`` / ``stage_group`` / ``stage_store``
        The thread block's tile of every state plane but V staged in shared
        memory by TMA (see the main kernel's prologue; :func:`stage_applies`):
        ``stage_group`` planes per arrival barrier, or explicit group sizes
        ``(first, second, ...)`` with the last group taking the rest;
        ``stage_store=False`` writes results with plain stores instead of TMA
        stores (better where steps overlap).
    ``v_direct``
        Staged kernels: no V tile in shared memory; a thread takes the
        potentials left and right of its cell from the adjacent lanes of its
        warp and those above and below from memory.
    ``tile_loop`` / ``stage_early`` / ``tile_call``
        The staged kernel as a loop over tiles (:func:`_emit_tile_loop`),
        ``stage_early`` state planes requested a tile ahead; ``tile_call``
        puts the per-cell step into a function of its own. Measured slower
        than the one-tile kernel; not a default.
    ``prefetch_next``
        Staged kernels: L2 hints for the tiles of the next wave of thread
        blocks, issued at this fraction of the equations. Measured slower.
    ``div_cubic``
        ``mkb_div`` with one third-order refinement of the reciprocal and no
        residual correction: 4 FP64 instructions instead of 6, IEEE results
        for divisors 0 and inf, at most 1.5 ulp from the quotient.
    ``div_parallel``
        ``mkb_div`` refines quotient and reciprocal side by side: a shorter
        dependent chain, and IEEE results for divisors 0 and inf. Same
        instruction count; off by default until measured.
    ``slab_lean``
        With ``slab``: threads outside the grid leave before the cell model
        (as in the single-GPU kernel) instead of skipping it inside a
        conditional region, and the flag publication at the end re-reads its
        block coordinates and the step number. ptxas then allocates the slab
        kernel like the plain one (decker-2009 fp64: 96 bytes of spills per
        thread instead of 348). The closing ``__syncthreads()`` is reached by
        all non-exited threads only, which is what the barrier waits for on
        sm_70 and later. Off by default until measured on a multi-GPU box.
    ``slab``
        Row-slab variant for multi-GPU grids: boundary row blocks run first,
        wait for the neighbouring GPU's ghost row (arrival flags), and push
        their new V row into the neighbour's ghost buffer with peer stores
        followed by a system-scope fence and a flag write.
    """
    sp = (precision == myokit.SINGLE_PRECISION)
    if native_maths and sp:
        w = _NativeWriter(precision)
    else:
        w = _Writer(precision)
        w._fast_div = bool(fast_div)
        w._fast_exp = ({'table': 'mkb_exp_tab', 'poly': 'mkb_exp_poly',
                        'stab': 'mkb_exp_stab',
                        'estrin': 'mkb_exp_estrin', 'ex2': 'mkb_expf_ex2'}.get(
            fast_exp, 'mkb_exp_poly') if fast_exp else False)
        if (fast_exp == 'ex2') != bool(sp) and fast_exp == 'ex2':
            raise ValueError("fast_exp='ex2' is a single-precision variant.")
        if const_pool and not sp:
            w.enable_pool()
        w._fast_libm = bool(fast_libm) and not sp
        w._select = select if select == 'cheap' else bool(select)
    w._pow_multiply = bool(pow_multiply)
    math_defines = '#define MKB_OVERLAP_STEPS @OVERLAP@\n#define MKB_EXP_SCALE_ADD %d\n#define MKB_FAST_LIBM %d\n#define MKB_EXP_FN %s' % (
        1 if exp_scale == 'add' else 0,
        1 if (fast_libm and not sp and not native_maths) else 0,
        getattr(w, '_fast_exp', None) if (getattr(w, '_fast_exp', None)
                                          and not sp) else 'exp')
    # the shared-memory exp table must be filled by every thread block
    stab = (fast_exp == 'stab') and not sp and not native_maths
    pooled = getattr(w, '_pool', None) is not None
    fields = list(fields)
    inter_log = list(inter_log)
    bx, by = block
    diffusion = diffusion_mode != DIFF_NONE
    slab = bool(slab) and diffusion_mode in (DIFF_HOMOGENEOUS, DIFF_FIELD)
    cpt = int(cells_per_thread or 1)
    if slab or diffusion_mode == DIFF_CONNECTIONS or cpt not in (2, 4, 8) \
            or (sp and cpt == 2) or junction or persistent or split_gates:
        cpt = 1
    if cpt > 1 and diffusion_mode in (DIFF_HOMOGENEOUS, DIFF_FIELD):
        # rim-exchange arrays of the register-patch path (static shared memory)
        rpt_ = max(int(rows_per_thread or 1), 1)
        smem = (2 * by * rpt_ * (bx + 2) + 2 * (by + 2) * bx * cpt) * (4 if sp else 8)
        if smem > 48 * 1024:
            raise ValueError(
                'Tile too large: block %dx%d with cells_per_thread=%d,'
                ' rows_per_thread=%d needs %d bytes of shared memory for its'
                ' rim exchange (limit 49152).' % (bx, by, cpt, rpt_, smem))

    # Streaming form of the vector path (see the emitter below): regular 2-d
    # grids whose rows are 16-byte multiples, one warp across a tile row
    stream = bool(stream) and cpt > 1 and diffusion_mode in (
        DIFF_HOMOGENEOUS, DIFF_FIELD) and bx == 32 and not stab

    # Consecutive steps overlap (see MKB_OVERLAP_STEPS in the prelude): scalar
    # path on regular grids only
    overlap = bool(overlap) and cpt == 1 and diffusion_mode in (
        DIFF_HOMOGENEOUS, DIFF_FIELD) and not (junction or persistent
                                               or split_gates or debug_mem)

    # States staged in shared memory by TMA (see the main kernel's prologue):
    # scalar path, regular grids and uncoupled cells; the caller checks that
    # rows are 16-byte multiples (the descriptor's strides)
    rs_ = 4 if sp else 8
    stage = (bool(stage) and cpt == 1 and lazy_state and not (
        persistent or split_gates or debug_mem or junction or stab)
        and diffusion_mode != DIFF_CONNECTIONS
        and (bx * rs_) % 16 == 0 and (bx * by * rs_) % 128 == 0
        and bx <= 256 and by <= 256)
    # Staged kernel as a loop over tiles (see the emitter below): plain grids
    # and uncoupled cells on one GPU, steps that do not overlap
    tile_loop = bool(tile_loop) and stage and stage_store and not (
        slab or overlap or partitioned or junction)
    if tile_loop:
        stage_group = (max(int(stage_early or 4), 1),)
    if isinstance(stage_group, (tuple, list)):
        # explicit group sizes; the last group takes what is left
        stage_sizes = [max(int(x), 1) for x in stage_group]
    else:
        stage_sizes = None
        stage_group = max(int(stage_group or 8), 1)

    def stage_group_of(j):
        if stage_sizes is None:
            return j // stage_group
        at = 0
        for gi, n in enumerate(stage_sizes):
            at += n
            if j < at:
                return gi
        return len(stage_sizes)

    if junction not in (None, 'fiber', 'tissue'):
        raise ValueError('junction must be None, "fiber" or "tissue".')
    if junction and (diffusion_mode != DIFF_HOMOGENEOUS or slab or cpt > 1):
        raise ValueError('A junction needs a homogeneous grid kernel with one'
                         ' cell per thread.')
    if split_gates and (persistent or junction):
        raise ValueError('split_gates cannot be combined with persistent or'
                         ' junction kernels.')
    if persistent and (diffusion_mode == DIFF_CONNECTIONS or slab or junction
                       or partitioned):
        raise ValueError('The persistent kernel is for unsharded grids and'
                         ' uncoupled cells.')

    equations = model.solvable_order()
    del equations['*remaining*']

    # Double precision: constants that do not depend on a field are evaluated
    # here (host double arithmetic, the same values a compiler's constant
    # folding produces up to the last ulp of log/exp/pow) and enter the
    # kernel as literals / constant-table entries at their point of use.
    folded = {}
    if pooled:
        field_dep = set(fields)
        for group in equations.values():
            for eq in group.equations(const=True):
                var = eq.lhs.var()
                if var in field_dep:
                    continue
                dep = False
                for ref in eq.rhs.references():
                    if ref.var() in field_dep:
                        dep = True
                        break
                if dep:
                    field_dep.add(var)
                    continue
                try:
                    x = float(eq.rhs.eval())
                except Exception:
                    x = float('nan')
                if x == x and x not in (float('inf'), float('-inf')):
                    folded[var] = x

    def number(x):
        ref = w.pool_ref(x)
        if ref is not None:
            return ref
        t = repr(float(x))
        return '(' + t + ')' if t[0] == '-' else t

    def v(var):
        # openclsim.cl:99-113
        if isinstance(var, myokit.Derivative):
            return 'D_' + var.var().uname()
        if isinstance(var, myokit.Name):
            var = var.var()
        if var in bound_variables:
            return bound_variables[var]
        if var in folded:
            return number(folded[var])
        return 'V_' + var.uname()
    w.set_lhs_function(v)

    def fold(e):
        # Constant sub-expression (only folded constants and literals)?
        if isinstance(e, myokit.Condition):
            return None
        for ref in e.references():
            if not isinstance(ref, myokit.Name) or ref.var() not in folded:
                return None
        try:
            x = float(e.eval())
        except Exception:
            return None
        if x != x or x in (float('inf'), float('-inf')):
            return None
        return number(x)
    if pooled:
        w.fold = fold

    def const_value(e):
        if isinstance(e, myokit.Condition):
            return None
        for ref in e.references():
            if not isinstance(ref, myokit.Name) or ref.var() not in folded:
                return None
        try:
            x = float(e.eval())
        except Exception:
            return None
        return x if x == x else None
    if fast_div and const_div and not native_maths:
        w.const_value = const_value

    n_state = model.count_states()
    vm = model.label('membrane_potential') if diffusion else None
    i_vm = vm.index() if vm is not None else -1
    real = 'float' if sp else 'double'
    exp = 'expf' if sp else (w._fast_exp if fast_exp else 'exp')
    if sp and fast_exp == 'ex2' and not native_maths:
        exp = 'mkb_expf_ex2'
    if native_maths and sp:
        exp = '__expf'
    states = list(model.states())
    state_set = set(states)
    inter_index = dict((var, k) for k, var in enumerate(inter_log))

    stage_slot = {}         # state -> slot in the staged tile (option stage)
    stage_waited = set()    # arrival groups already waited for

    def state_load(var, guarded=False):
        k = var.index()
        if k == i_vm:
            return '    const Real %s = vc;' % v(var)
        if var in stage_slot:
            j = stage_slot[var]
            lines = []
            if stage_group_of(j) not in stage_waited:
                stage_waited.add(stage_group_of(j))
                lines.append('    MKB_STAGE_WAIT(%d);' % stage_group_of(j))
            lines.append('    const Real %s = MKB_STAGE_AT(%d);' % (v(var), j))
            return '\n'.join(lines)
        src = 'MKB_LD(&MKB_AT(state_c, %d))' % k
        if debug_mem:
            src = 'MKB_AT(state + (cid & 255ull), %d)' % k
        if guarded:
            src = 'active ? %s : (Real)0' % src
        return '    const Real %s = %s;' % (v(var), src)

    def state_rhs(var):
        # openclsim.cl:358-364
        if var in rl_states:
            inf, tau = rl_states[var]
            inf, tau, x = v(inf), v(tau), v(var)
            arg = '-dt / %s' % tau
            if fast_div and not sp and not native_maths:
                arg = 'mkb_div(-dt, %s)' % tau
                trhs = rl_states[var][1].rhs()
                if (w.const_value is not None
                        and isinstance(trhs, myokit.Divide)):
                    # tau = c / X: -dt / tau = -dt * X * (1 / c), no division
                    # (X is the expression tau was computed from: the
                    # compiler reuses its value)
                    c = w.const_value(trhs[0])
                    if c is not None and c != 0 and 1e-290 < abs(c) < 1e290:
                        arg = '(-dt * (%s))' % w.ex(trhs[1])
                        if c != 1.0:
                            arg = '(%s * %s)' % (arg, w.ex(myokit.Number(1.0 / c)))
            return '%s - (%s - %s) * %s(%s)' % (inf, inf, x, exp, arg)
        return '%s + dt * %s' % (v(var), v(var.lhs()))

    def state_update(var):
        k = var.index()
        rhs = state_rhs(var)
        if k == i_vm and slab:
            return (
                '    {\n'
                '        const Real vnew = %s;\n'
                '        v_out[cid] = vnew;\n'
                '        // boundary rows go straight into the neighbouring GPU\'s ghost row\n'
                '        if (iy == 0 && g.peer_lo_halo_hi)\n'
                '            ((Real*)g.peer_lo_halo_hi)[((step + 1u) %% 3u) * nx + ix] = vnew;\n'
                '        if (iy == ny - 1 && g.peer_hi_halo_lo)\n'
                '            ((Real*)g.peer_hi_halo_lo)[((step + 1u) %% 3u) * nx + ix] = vnew;\n'
                '    }' % rhs)
        if k == i_vm:
            return '    v_out[cid] = %s;' % rhs
        if var in stage_slot and stage_store:
            return '    MKB_STAGE_AT(%d) = %s;' % (stage_slot[var], rhs)
        if debug_mem == 'l1ns':
            return '    { const Real vnew = %s; if (dt < (Real)0) MKB_AT(state_c, %d) = vnew; }' % (rhs, k)
        return '    MKB_AT(state_c, %d) = %s;' % (k, rhs)

    # Equations to emit, in the reference's order (openclsim.cl:235-243)
    todo = []       # (component name or None, equation)
    for name, group in equations.items():
        first = True
        for eq in group.equations(const=False):
            var = eq.lhs.var()
            if var in rl_states or var in bound_variables:
                continue
            todo.append((name if first else None, eq))
            first = False

    def refs(expr):
        out = []
        for ref in expr.references():
            if isinstance(ref, myokit.Name) and ref.var() in state_set:
                out.append(ref.var())
        return sorted(set(out), key=lambda x: x.index())

    # ------------------------------------------------------------------
    # split_gates: which states can leave the big kernel
    # ------------------------------------------------------------------
    gate_states = []        # states updated by mkb_gate_step
    gate_todo = []          # its equations, in the reference's order
    # (only with diffusion: the second kernel must still find V(t) after the
    # first has produced V(t + dt), which takes the two V planes)
    if split_gates and diffusion:
        def eq_key(eq):
            return ('D', eq.lhs.var()) if eq.lhs.is_derivative() else eq.lhs.var()
        by_key = dict((eq_key(eq), eq) for name, eq in todo)
        bound_set = set(bound_variables)

        def closure(roots):
            # equations (keys of todo) and states / bound variables reached
            keys, reads = set(), set()
            stack = list(roots)
            while stack:
                key = stack.pop()
                if key in keys or key not in by_key:
                    continue
                keys.add(key)
                for ref in by_key[key].rhs.references():
                    var = ref.var()
                    if isinstance(ref, myokit.Derivative):
                        stack.append(('D', var))
                    elif var in state_set or var in bound_set:
                        reads.add(var)
                    else:
                        stack.append(var)
            return keys, reads

        def roots_of(var):
            if var in rl_states:
                return list(rl_states[var])
            return [('D', var)]
        vtime = model.time()
        gate_keys = set()
        for var in states:
            if var is vm:
                continue
            keys, reads = closure(roots_of(var))
            allowed = set([var, vtime])
            if vm is not None:
                allowed.add(vm)
            if reads <= allowed:
                gate_states.append(var)
                gate_keys |= keys
        rest_keys = set()
        for var in states:
            if var not in gate_states:
                rest_keys |= closure(roots_of(var))[0]
        only_gates = gate_keys - rest_keys
        if gate_states:
            gate_todo = [(name, eq) for name, eq in todo if eq_key(eq) in gate_keys]
            todo = [(name, eq) for name, eq in todo if eq_key(eq) not in only_gates]
    gate_set = set(gate_states)

    # ------------------------------------------------------------------
    # Section: model body (equations, loads, stores)
    # ------------------------------------------------------------------
    early = []      # state loads hoisted above the stencil (guarded)
    early_states = set()
    gate_set_unused = set()
    body = []
    if not lazy_state:
        used = set()
        for name, eq in todo:
            used.update(refs(eq.rhs))
        for var in states:
            if var in gate_set and var not in used:
                continue
            early.append(state_load(var, guarded=True))
        for name, eq in todo:
            if name:
                body.append('    // Component: %s' % name)
            var = eq.lhs.var()
            body.append('    const Real %s = %s;' % (v(eq.lhs), w.ex(eq.rhs)))
            if var in inter_index and not eq.lhs.is_derivative():
                body.append(
                    '    if (store_aux) MKB_AT(inter_c, %d) = %s;'
                    % (inter_index[var], v(eq.lhs)))
        body.append('    // Update (openclsim.cl:358-364)')
        for var in states:
            if var not in gate_set:
                body.append(state_update(var))
    else:
        # Same equations, same order, same arithmetic; only the position of
        # the loads and stores differs. No thread reads another thread's
        # states except V, which is double-buffered, and a state is loaded
        # exactly once, so a store can never be observed by a load of the
        # same step.
        first_use = {}
        for i, (name, eq) in enumerate(todo):
            for r in refs(eq.rhs):
                first_use.setdefault(r, i)
        for var in states:      # never referenced: needed by its own update
            if var not in gate_set:
                first_use.setdefault(var, len(todo))
        order = sorted([x for x in states if x in first_use],
                       key=lambda x: (first_use[x], x.index()))
        if stage:
            for var in order:
                if var.index() != i_vm:
                    stage_slot[var] = len(stage_slot)
        loaded = set()
        have = set()
        done = set(gate_set)    # updated by mkb_gate_step, if any
        ahead = max(int(load_ahead), 0)

        def emit_loads(limit, dest, guarded):
            for var in order:
                if var in loaded:
                    continue
                if first_use[var] <= limit:
                    dest.append(state_load(var, guarded))
                    loaded.add(var)

        def flush_updates():
            for var in states:
                if var in done:
                    continue
                if var in rl_states:
                    ready = all(x in have for x in rl_states[var])
                else:
                    ready = ('D', var) in have
                if ready:
                    if var not in loaded:
                        body.append(state_load(var))
                        loaded.add(var)
                    body.append(state_update(var))
                    done.add(var)

        # (staged states are read from shared memory: after the tile barrier)
        emit_loads(ahead, body if stage else early, True)
        early_states = set(loaded)
        gate_set_unused = set(x for x in gate_set if x not in first_use)
        next_at = None
        if stage and prefetch_next:
            next_at = min(max(int(float(prefetch_next) * len(todo)), 0), len(todo) - 1)
        for i, (name, eq) in enumerate(todo):
            if name:
                body.append('    // Component: %s' % name)
            if i == next_at:
                body.append('@PREFETCH_NEXT@')
            emit_loads(i + ahead, body, False)
            var = eq.lhs.var()
            body.append('    const Real %s = %s;' % (v(eq.lhs), w.ex(eq.rhs)))
            if eq.lhs.is_derivative():
                have.add(('D', var))
            else:
                have.add(var)
                if var in inter_index:
                    body.append(
                        '    if (store_aux) MKB_AT(inter_c, %d) = %s;'
                        % (inter_index[var], v(eq.lhs)))
            flush_updates()
        emit_loads(len(todo), body, False)
        flush_updates()
        missing = [x.qname() for x in states if x not in done]
        if missing:     # pragma: no cover
            raise RuntimeError('No update generated for: ' + ', '.join(missing))

    # Constants that stay in the kernel (fields' dependants, single precision)
    consts = []
    for k, var in enumerate(fields):
        early.append(
            '    const Real %s = active ? MKB_AT(field_c, %d) : (Real)0;'
            % (v(var), k))
    consts.append('    // Literal constants (openclsim.cl:173-178)')
    for group in equations.values():
        for eq in group.equations(const=True):
            if isinstance(eq.rhs, myokit.Number):
                var = eq.lhs.var()
                if var not in fields and var not in folded:
                    consts.append('    const Real %s = %s;' % (v(eq.lhs), w.ex(eq.rhs)))
    consts.append('    // Calculated constants (openclsim.cl:181-186): folded (by the')
    consts.append('    // compiler, or on the host for double precision) unless they')
    consts.append('    // depend on a field')
    for group in equations.values():
        for eq in group.equations(const=True):
            if not isinstance(eq.rhs, myokit.Number):
                var = eq.lhs.var()
                if var not in fields and var not in folded:
                    consts.append('    const Real %s = %s;' % (v(eq.lhs), w.ex(eq.rhs)))

    # The body of mkb_gate_step, rendered here: its literals must be in the
    # constant table before the table is printed
    gate_lines = []
    if gate_states:
        if vm is not None:
            if diffusion:
                gate_lines.append('    const Real %s = v_in[cid];' % v(vm))
            else:
                gate_lines.append('    const Real %s = MKB_AT(state_c, %d);'
                                  % (v(vm), vm.index()))
        for k, var in enumerate(fields):
            gate_lines.append(
                '    const Real %s = MKB_AT(field_c, %d);'
                % (v(var), k))
        for var in gate_states:
            gate_lines.append('    const Real %s = MKB_AT(state_c, %d);'
                              % (v(var), var.index()))
        gate_lines.extend(consts)
        for name, eq in gate_todo:
            var = eq.lhs.var()
            gate_lines.append('    const Real %s = %s;' % (v(eq.lhs), w.ex(eq.rhs)))
            if var in inter_index and not eq.lhs.is_derivative():
                gate_lines.append(
                    '    if (store_aux) MKB_AT(inter_c, %d) = %s;'
                    % (inter_index[var], v(eq.lhs)))
        for var in gate_states:
            gate_lines.append('    MKB_AT(state_c, %d) = %s;'
                              % (var.index(), state_rhs(var)))

    # ------------------------------------------------------------------
    # Persistent path: the whole grid in one thread block, many steps per launch
    # ------------------------------------------------------------------
    if persistent:
        o = []
        q = o.append
        q('// Generated by myokit_b200.kernelgen for sm_100a — do not edit.')
        q('// Model: %s (persistent: one block, all steps up to the next logged one)' % model.name())
        q('#include "mkb_device_abi.h"')
        q('typedef %s Real;' % real)
        q('#define MKB_BX %d' % bx)
        q('#define MKB_BY %d' % by)
        q('#define MKB_DIV_INT_CHECK %d' % (1 if div_int_check else 0))
        q('#define MKB_DIV_PARALLEL %d' % (1 if div_parallel else 0))
        q('#define MKB_DIV_CUBIC %d' % (1 if div_cubic else 0))
        q(math_defines.replace('@OVERLAP@', '0'))
        q(_PRELUDE)
        if pooled and w._pool:
            q('__constant__ double mkb_k[%d] = {' % len(w._pool))
            for x in w._pool:
                q('    %r,' % x)
            q('};')
        q('extern "C" __global__ void __launch_bounds__(MKB_BX * MKB_BY)')
        q('%s_persistent(const MkbGridArgs g, const MkbStepParams* __restrict__ sp_first,' % KERNEL_NAME)
        q('    const Real* __restrict__ v_in, Real* __restrict__ v_out)')
        q('{')
        q('    // The block is the grid: thread (tx, ty) owns cell (tx, ty).')
        q('    const unsigned int tx = threadIdx.x, ty = threadIdx.y;')
        q('    const unsigned int nx = (unsigned int)g.nx, ny = (unsigned int)g.ny;')
        q('    const unsigned long long stride = g.stride;')
        q('    const unsigned int ix = tx, iy = ty;')
        q('    const bool active = (ix < nx) && (iy < ny);')
        q('    const unsigned long long cid = (unsigned long long)iy * nx + ix;')
        q('    Real* const state = (Real*)g.state;')
        q('    (void)v_in; (void)v_out; (void)stride;')
        q('    // Every state stays in a register from here to the end of the launch')
        for var in states:
            k = var.index()
            src = 'v_in[cid]' if k == i_vm else 'state[%dull * stride + cid]' % k
            q('    Real S%d = active ? %s : (Real)0;' % (k, src))
        for k, var in enumerate(fields):
            q('    const Real %s = active ? ((const Real*)g.field)[%dull * stride + cid] : (Real)0;'
              % (v(var), k))
        if diffusion_mode == DIFF_FIELD:
            q('    // Edge conductances (openclsim.cl:475-482), constant over the steps')
            q('    const Real* const gxf = (const Real*)g.gx_field;')
            q('    const Real* const gyf = (const Real*)g.gy_field;')
            q('    const bool has_xm = active && nx > 1 && ix > 0;')
            q('    const bool has_xp = active && nx > 1 && ix < nx - 1;')
            q('    const bool has_ym = active && ny > 1 && iy > 0;')
            q('    const bool has_yp = active && ny > 1 && iy < ny - 1;')
            q('    const Real gxm = has_xm ? gxf[cid - iy - 1] : (Real)0;')
            q('    const Real gxp = has_xp ? gxf[cid - iy] : (Real)0;')
            q('    const Real gym = has_ym ? gyf[(long long)cid - (long long)nx] : (Real)0;')
            q('    const Real gyp = has_yp ? gyf[cid] : (Real)0;')
        if diffusion:
            q('    __shared__ Real tile[MKB_BY + 2][MKB_BX + 2];')
        if stab:
            q('    MKB_EXP_TABLE_INIT(ty * MKB_BX + tx, MKB_BX * MKB_BY);')
            q('    __syncthreads();')
        q('    // number of steps of this launch: upper bits of the first record\'s flags')
        q('    const unsigned int count = sp_first->flags >> 8;')
        q('    for (unsigned int it = 0; it < count; it++) {')
        q('    const MkbStepParams* const sp = sp_first + it;')
        q('    // Per-step scalars, cast like openclsim.c:1063,1148,1155')
        q('    const Real time = (Real)sp->time;')
        q('    const Real dt = (Real)sp->dt;')
        q('    const Real pace_in = (Real)sp->pace;')
        q('    const bool store_aux = (sp->flags & MKB_FLAG_STORE_AUX) != 0;')
        q('    (void)time; (void)pace_in; (void)store_aux;')
        if diffusion:
            q('    const Real vc = S%d;' % i_vm)
            q('    // V(t) of the whole grid through shared memory; cells at the rim see')
            q('    // their own V beyond it (zero flux, openclsim.cl:401-434)')
            q('    if (active) {')
            q('        tile[ty + 1][tx + 1] = vc;')
            q('        if (ix == 0) tile[ty + 1][0] = vc;')
            q('        if (ix == nx - 1) tile[ty + 1][tx + 2] = vc;')
            q('        if (iy == 0) tile[0][tx + 1] = vc;')
            q('        if (iy == ny - 1) tile[ty + 2][tx + 1] = vc;')
            q('    }')
            q('    __syncthreads();')
        q('    if (active) {')
        if diffusion:
            q('    const Real vxm = tile[ty + 1][tx], vxp = tile[ty + 1][tx + 2];')
            q('    const Real vym = tile[ty][tx + 1], vyp = tile[ty + 2][tx + 1];')
            q('    Real idiff;')
            if diffusion_mode == DIFF_HOMOGENEOUS:
                q('    // openclsim.cl:401-434 (diff_step)')
                q('    const Real gx = (Real)g.gx, gy = (Real)g.gy;')
                q('    if (nx > 1) {')
                q('        if (ix == 0) idiff = gx * (vc - vxp);')
                q('        else if (ix == nx - 1) idiff = gx * (vc - vxm);')
                q('        else idiff = gx * (2 * vc - vxm - vxp);')
                q('    } else {')
                q('        idiff = 0;')
                q('    }')
                q('    if (ny > 1) {')
                q('        if (iy == 0) idiff += gy * (vc - vyp);')
                q('        else if (iy == ny - 1) idiff += gy * (vc - vym);')
                q('        else idiff += gy * (2 * vc - vym - vyp);')
                q('    }')
            else:
                q('    // openclsim.cl:469-486 (diff_hetero)')
                q('    idiff = 0.0;')
                q('    if (has_xm) { idiff += gxm * (vc - vxm); }')
                q('    if (has_xp) { idiff += gxp * (vc - vxp); }')
                q('    if (has_ym) idiff += gym * (vc - vym);')
                q('    if (has_yp) idiff += gyp * (vc - vyp);')
            q('    // openclsim.cl:249-280, 322-329')
            if paced_list:
                q('    const Real pace = g.paced_mask[cid] ? pace_in : (Real)0;')
            else:
                q('    const int pix = (int)ix, piy = (int)iy;')
                q('    const Real pace = (pix >= (int)g.pace_x0 && pix < (int)g.pace_x1 &&')
                q('                       piy >= (int)g.pace_y0 && piy < (int)g.pace_y1) ? pace_in : (Real)0;')
            q('    if (store_aux) ((Real*)g.idiff)[cid] = idiff;')
        else:
            q('    const Real pace = pace_in;')
        q('    (void)pace;')
        for line in consts:
            q(line)
        q('    // the states as this step sees them')
        for var in states:
            q('    const Real %s = S%d;' % (v(var), var.index()))
        for name, eq in todo:
            if name:
                q('    // Component: %s' % name)
            var = eq.lhs.var()
            q('    const Real %s = %s;' % (v(eq.lhs), w.ex(eq.rhs)))
            if var in inter_index and not eq.lhs.is_derivative():
                q('    if (store_aux) ((Real*)g.inter)[%dull * stride + cid] = %s;'
                  % (inter_index[var], v(eq.lhs)))
        q('    // Update (openclsim.cl:358-364): all from the values of time t')
        for var in states:
            q('    const Real N%d = %s;' % (var.index(), state_rhs(var)))
        for var in states:
            q('    S%d = N%d;' % (var.index(), var.index()))
        q('    }   // active')
        if diffusion:
            q('    __syncthreads();    // everybody has read the tile')
        q('    }   // steps')
        q('    if (active) {')
        for var in states:
            k = var.index()
            if k == i_vm:
                q('        v_out[cid] = S%d;' % k)
            else:
                q('        state[%dull * stride + cid] = S%d;' % (k, k))
        q('    }')
        q('}')
        q('')
        options = ['--fmad=true' if fmad else '--fmad=false']
        if max_registers:
            options.append('--maxrregcount=%d' % int(max_registers))
        ks = KernelSource('\n'.join(o), block, n_state, i_vm, len(inter_log),
                          len(fields), diffusion_mode, options)
        ks.kernel_name = KERNEL_NAME + '_persistent'
        ks.kernel_flags = 1             # MKB_KERNEL_PERSISTENT
        ks.persistent = True
        return ks

    # ------------------------------------------------------------------
    # Vector path: several x-adjacent cells (and rows) per thread
    # ------------------------------------------------------------------
    if cpt > 1:
        rpt = max(int(rows_per_thread or 1), 1)
        o = []
        q = o.append
        q('// Generated by myokit_b200.kernelgen for sm_100a — do not edit.')
        q('// Model: %s (%d x %d cells per thread)' % (model.name(), cpt, rpt))
        q('#include "mkb_device_abi.h"')
        q('typedef %s Real;' % real)
        q('#define MKB_BX %d' % bx)
        q('#define MKB_BY %d' % by)
        q('#define MKB_CPT %d' % cpt)
        q('#define MKB_RPT %d' % rpt)
        q('#define MKB_DIV_INT_CHECK %d' % (1 if div_int_check else 0))
        q('#define MKB_DIV_PARALLEL %d' % (1 if div_parallel else 0))
        q('#define MKB_DIV_CUBIC %d' % (1 if div_cubic else 0))
        q(math_defines.replace('@OVERLAP@', '0'))
        q(_PRELUDE)
        q(_VECTOR_PRELUDE)
        if pooled and w._pool:
            q('__constant__ double mkb_k[%d] = {' % len(w._pool))
            for x in w._pool:
                q('    %r,' % x)
            q('};')
        if stream:
            # ----------------------------------------------------------
            # Streaming form: persistent thread blocks walk the tiles of
            # the grid; the V tile with its halo arrives in shared memory
            # by TMA (cp.async.bulk.tensor.2d, out-of-grid cells zero
            # filled) into a two-stage ring guarded by mbarriers, so the
            # next tile's load is in flight while this one is computed and
            # no thread holds a register or issues an instruction for it.
            # One warp spans a tile row: left / right neighbours of a
            # thread's patch come by shuffle, rows above / below from the
            # tile. Everything after the loads is the vector path's code.
            # ----------------------------------------------------------
            q('#define MKB_TW (MKB_BX * MKB_CPT)')
            q('#define MKB_TH (MKB_BY * MKB_RPT)')
            q('#define MKB_PAD (16 / (int)sizeof(Real))')
            q('#define MKB_BW (MKB_TW + 2 * MKB_PAD)')
            q('#define MKB_BH (MKB_TH + 2)')
            q('#define MKB_STAGES 2')
            q('#define MKB_I_VM %d' % i_vm)
            q('extern "C" __global__ void __launch_bounds__(MKB_BX * MKB_BY, %d)' % int(min_blocks or 2))
            q('%s(const __grid_constant__ MkbGridArgs g, const MkbStepParams* __restrict__ sp,' % KERNEL_NAME)
            q('    const Real* __restrict__ v_in, Real* __restrict__ v_out)')
            q('{')
            q('    // (each stage 128-byte aligned, as the TMA destination must be)')
            q('    struct alignas(128) MkbStage { Real v[MKB_BH][MKB_BW]; };')
            q('    __shared__ MkbStage tile[MKB_STAGES];')
            q('    __shared__ unsigned long long full[MKB_STAGES];')
            q('    const unsigned int tx = threadIdx.x, ty = threadIdx.y;')
            q('    const unsigned int tid = ty * MKB_BX + tx;')
            q('    const unsigned int nx = (unsigned int)g.nx, ny = (unsigned int)g.ny;')
            q('    const unsigned long long stride = g.stride;')
            q('    const unsigned int ntx = (nx + MKB_TW - 1) / MKB_TW;')
            q('    const unsigned int n_tiles = ntx * ((ny + MKB_TH - 1) / MKB_TH);')
            q('    Real* const state = (Real*)g.state;')
            q('    // descriptor of the V plane this step reads')
            q('    const void* const tmap = (v_in == state + (unsigned long long)MKB_I_VM * stride)')
            q('        ? (const void*)g.tmap[0] : (const void*)g.tmap[1];')
            q('    const Real time = (Real)sp->time;')
            q('    const Real dt = (Real)sp->dt;')
            q('    const Real pace_in = (Real)sp->pace;')
            q('    const bool store_aux = (sp->flags & MKB_FLAG_STORE_AUX) != 0;')
            q('    (void)time; (void)pace_in; (void)store_aux; (void)v_out;')
            q('    const unsigned int nyg = (unsigned int)g.ny_global;')
            q('    if (tid == 0) {')
            q('        for (int s = 0; s < MKB_STAGES; s++) MKB_MBAR_INIT(&full[s], 1);')
            q('        MKB_MBAR_FENCE_INIT();')
            q('        for (unsigned int s = 0; s < MKB_STAGES; s++) {')
            q('            const unsigned int t = blockIdx.x + s * gridDim.x;')
            q('            if (t < n_tiles) {')
            q('                const unsigned int tyb = t / ntx, txb = t - tyb * ntx;')
            q('                MKB_TMA_LOAD_2D(&tile[s].v[0][0], tmap, (int)(txb * MKB_TW) - MKB_PAD, (int)(tyb * MKB_TH) - 1,')
            q('                                &full[s], MKB_BW, MKB_BH, (unsigned int)(MKB_BW * MKB_BH * sizeof(Real)));')
            q('            }')
            q('        }')
            q('    }')
            q('    __syncthreads();')
            q('    unsigned int k_iter = 0;')
            q('    for (unsigned int t = blockIdx.x; t < n_tiles; t += gridDim.x, k_iter++) {')
            q('    const unsigned int stg = k_iter % MKB_STAGES, phase = (k_iter / MKB_STAGES) & 1u;')
            q('    const unsigned int tyb = t / ntx, txb = t - tyb * ntx;')
            q('    const unsigned int ix0 = txb * MKB_TW + tx * MKB_CPT;')
            q('    const unsigned int iy0 = tyb * MKB_TH + ty * MKB_RPT;')
            q('    const bool in_x = ix0 < nx;')
            q('    // Vector loads of the other planes first: in flight while the tile is awaited')
            q('    Real vc[MKB_RPT][MKB_CPT];')
            for var in states:
                if var.index() != i_vm:
                    q('    Real S%d[MKB_RPT][MKB_CPT];' % var.index())
            for k, var in enumerate(fields):
                q('    Real F%d[MKB_RPT][MKB_CPT];' % k)
            if len(states) > 1 or fields:
                q('    #pragma unroll')
                q('    for (int r = 0; r < MKB_RPT; r++) {')
                q('        const unsigned int iy = iy0 + r;')
                q('        const bool active = in_x && (iy < ny);')
                q('        const unsigned long long cid0 = (unsigned long long)iy * nx + ix0;')
                for var in states:
                    if var.index() != i_vm:
                        q('        mkb_vload<MKB_CPT>(S%d[r], state + %dull * stride + cid0, active);'
                          % (var.index(), var.index()))
                for k, var in enumerate(fields):
                    q('        mkb_vload<MKB_CPT>(F%d[r], (const Real*)g.field + %dull * stride + cid0, active);'
                      % (k, k))
                q('    }')
            q('    MKB_MBAR_WAIT(&full[stg], phase);')
            q('    // this thread\'s patch, the rows above and below it, the cells left and right')
            q('    Real v_up[MKB_CPT], v_dn[MKB_CPT], v_lf[MKB_RPT], v_rt[MKB_RPT];')
            q('    {')
            q('        const Real (*tl)[MKB_BW] = tile[stg].v;')
            q('        const unsigned int c0 = MKB_PAD + tx * MKB_CPT, r0 = ty * MKB_RPT + 1;')
            q('        mkb_vload<MKB_CPT>(v_up, &tl[r0 - 1][c0], true);')
            q('        mkb_vload<MKB_CPT>(v_dn, &tl[r0 + MKB_RPT][c0], true);')
            q('        #pragma unroll')
            q('        for (int r = 0; r < MKB_RPT; r++) {')
            q('            mkb_vload<MKB_CPT>(vc[r], &tl[r0 + r][c0], true);')
            q('            const Real lf = MKB_SHFL_UP(vc[r][MKB_CPT - 1], 1);')
            q('            const Real rt = MKB_SHFL_DOWN(vc[r][0], 1);')
            q('            v_lf[r] = (tx == 0) ? tl[r0 + r][MKB_PAD - 1] : lf;')
            q('            v_rt[r] = (tx == MKB_BX - 1) ? tl[r0 + r][MKB_PAD + MKB_TW] : rt;')
            q('        }')
            q('    }')
            q('    __syncthreads();        // every thread has read the stage: refill it')
            q('    if (tid == 0) {')
            q('        const unsigned int tn = t + MKB_STAGES * gridDim.x;')
            q('        if (tn < n_tiles) {')
            q('            const unsigned int nyb = tn / ntx, nxb = tn - nyb * ntx;')
            q('            MKB_TMA_LOAD_2D(&tile[stg].v[0][0], tmap, (int)(nxb * MKB_TW) - MKB_PAD, (int)(nyb * MKB_TH) - 1,')
            q('                            &full[stg], MKB_BW, MKB_BH, (unsigned int)(MKB_BW * MKB_BH * sizeof(Real)));')
            q('        }')
            q('    }')
            q('    if (!in_x || iy0 >= ny) continue;')
            q('    const bool at_x0 = (ix0 == 0), at_x1 = (ix0 + MKB_CPT == nx);')
        else:
            q('extern "C" __global__ void __launch_bounds__(MKB_BX * MKB_BY)')
            q('%s(const MkbGridArgs g, const MkbStepParams* __restrict__ sp,' % KERNEL_NAME)
            q('    const Real* __restrict__ v_in, Real* __restrict__ v_out)')
            q('{')
            q('    const unsigned int tx = threadIdx.x, ty = threadIdx.y;')
            q('    const unsigned int nx = (unsigned int)g.nx, ny = (unsigned int)g.ny;')
            q('    const unsigned long long stride = g.stride;')
            q('    const unsigned int nby = (ny + MKB_BY * MKB_RPT - 1) / (MKB_BY * MKB_RPT);')
            q('    const unsigned int byr = blockIdx.y + blockIdx.z * gridDim.y;')
            q('    if (byr >= nby) return;')
            q('    // This thread owns the patch of cells ix0 .. ix0 + MKB_CPT - 1 (nx is a')
            q('    // multiple of MKB_CPT: all inside or all outside) by rows iy0 .. iy0 +')
            q('    // MKB_RPT - 1. Neighbours inside the patch are in registers; only the')
            q('    // rim of the patch is exchanged through shared memory.')
            q('    const unsigned int ix0 = (blockIdx.x * MKB_BX + tx) * MKB_CPT;')
            q('    const unsigned int iy0 = (byr * MKB_BY + ty) * MKB_RPT;')
            q('    const bool in_x = ix0 < nx;')
            q('    Real* const state = (Real*)g.state;')
            q('    const Real time = (Real)sp->time;')
            q('    const Real dt = (Real)sp->dt;')
            q('    const Real pace_in = (Real)sp->pace;')
            q('    const bool store_aux = (sp->flags & MKB_FLAG_STORE_AUX) != 0;')
            q('    (void)time; (void)pace_in; (void)store_aux; (void)v_in; (void)v_out;')
            q('')
            q('    // All vector loads of this thread first: they are in flight together')
            if diffusion:
                q('    Real vc[MKB_RPT][MKB_CPT];')
            for var in states:
                if var.index() != i_vm:
                    q('    Real S%d[MKB_RPT][MKB_CPT];' % var.index())
            for k, var in enumerate(fields):
                q('    Real F%d[MKB_RPT][MKB_CPT];' % k)
            q('    #pragma unroll')
            q('    for (int r = 0; r < MKB_RPT; r++) {')
            q('        const unsigned int iy = iy0 + r;')
            q('        const bool active = in_x && (iy < ny);')
            q('        const unsigned long long cid0 = (unsigned long long)iy * nx + ix0;')
            if diffusion:
                q('        mkb_vload<MKB_CPT>(vc[r], v_in + cid0, active);')
            for var in states:
                if var.index() != i_vm:
                    q('        mkb_vload<MKB_CPT>(S%d[r], state + %dull * stride + cid0, active);'
                      % (var.index(), var.index()))
            for k, var in enumerate(fields):
                q('        mkb_vload<MKB_CPT>(F%d[r], (const Real*)g.field + %dull * stride + cid0, active);'
                  % (k, k))
            q('    }')
            if diffusion_mode in (DIFF_HOMOGENEOUS, DIFF_FIELD):
                q('    const unsigned int nyg = (unsigned int)g.ny_global;')
                q('    // Rim exchange: first / last column of every row of the patch, first')
                q('    // and last row of the patch; block-edge threads fetch the true halo.')
                q('    __shared__ Real col_l[MKB_BY * MKB_RPT][MKB_BX + 2];   // [row][tx + 1]: first cell')
                q('    __shared__ Real col_r[MKB_BY * MKB_RPT][MKB_BX + 2];   // [row][tx + 1]: last cell')
                q('    __shared__ Real row_t[MKB_BY + 2][MKB_BX * MKB_CPT];   // [ty + 1]: first row')
                q('    __shared__ Real row_b[MKB_BY + 2][MKB_BX * MKB_CPT];   // [ty + 1]: last row')
                q('    #pragma unroll')
                q('    for (int r = 0; r < MKB_RPT; r++) {')
                q('        const unsigned int iy = iy0 + r;')
                q('        const unsigned int tr = ty * MKB_RPT + r;')
                q('        // (own slots only for threads of the grid: the slot of a thread past the')
                q('        // last column is the halo slot of its left neighbour, written below)')
                q('        if (in_x && iy < ny) {')
                q('            col_l[tr][tx + 1] = vc[r][0];')
                q('            col_r[tr][tx + 1] = vc[r][MKB_CPT - 1];')
                q('            const unsigned long long cid0 = (unsigned long long)iy * nx + ix0;')
                q('            // halo cells left and right of the block')
                q('            if (tx == 0) col_r[tr][0] = (ix0 > 0) ? v_in[cid0 - 1] : vc[r][0];')
                q('            if (tx == MKB_BX - 1 || ix0 + MKB_CPT >= nx)')
                q('                col_l[tr][tx + 2] = (ix0 + MKB_CPT < nx) ? v_in[cid0 + MKB_CPT] : vc[r][MKB_CPT - 1];')
                q('        }')
                q('    }')
                q('    #pragma unroll')
                q('    for (int c = 0; c < MKB_CPT; c++) {')
                q('        row_t[ty + 1][tx * MKB_CPT + c] = vc[0][c];')
                q('        row_b[ty + 1][tx * MKB_CPT + c] = vc[MKB_RPT - 1][c];')
                q('    }')
                q('    if (in_x && iy0 < ny) {')
                q('        if (ty == 0) {')
                q('            // row above the block')
                q('            Real vn[MKB_CPT];')
                q('            mkb_vload<MKB_CPT>(vn, v_in + (unsigned long long)(iy0 - 1) * nx + ix0, iy0 > 0);')
                q('            #pragma unroll')
                q('            for (int c = 0; c < MKB_CPT; c++) row_b[0][tx * MKB_CPT + c] = vn[c];')
                q('        }')
                q('        if (ty == MKB_BY - 1) {')
                q('            // row below the block')
                q('            Real vn[MKB_CPT];')
                q('            mkb_vload<MKB_CPT>(vn, v_in + (unsigned long long)(iy0 + MKB_RPT) * nx + ix0, iy0 + MKB_RPT < ny);')
                q('            #pragma unroll')
                q('            for (int c = 0; c < MKB_CPT; c++) row_t[MKB_BY + 1][tx * MKB_CPT + c] = vn[c];')
                q('        }')
                q('    }')
                if stab:
                    q('    MKB_EXP_TABLE_INIT(ty * MKB_BX + tx, MKB_BX * MKB_BY);')
                q('    __syncthreads();')
                q('    if (!in_x || iy0 >= ny) return;')
                q('    const bool at_x0 = (ix0 == 0), at_x1 = (ix0 + MKB_CPT == nx);')
            else:
                if stab:
                    q('    MKB_EXP_TABLE_INIT(ty * MKB_BX + tx, MKB_BX * MKB_BY);')
                    q('    __syncthreads();')
                q('    if (!in_x || iy0 >= ny) return;')
        if diffusion_mode == DIFF_HOMOGENEOUS:
            q('    const Real gx = (Real)g.gx, gy = (Real)g.gy;')
        if diffusion_mode == DIFF_FIELD:
            q('    const Real* const gxf = (const Real*)g.gx_field;')
            q('    const Real* const gyf = (const Real*)g.gy_field;')
        if diffusion and not paced_list:
            q('    // paced rectangle, relative to this patch (openclsim.cl:260-274)')
            q('    const bool pacing_on = (pace_in != (Real)0);')
            q('    const int pc0 = (int)g.pace_x0 - (int)ix0, pc1 = (int)g.pace_x1 - (int)ix0;')
        q('')
        q('    #pragma unroll')
        q('    for (int r = 0; r < MKB_RPT; r++) {')
        q('    const unsigned int iy = iy0 + r;')
        q('    if (iy >= ny) break;')
        q('    const unsigned long long cid0 = (unsigned long long)iy * nx + ix0;')
        for var in states:
            q('    Real N%d[MKB_CPT];' % var.index())
        if diffusion_mode in (DIFF_HOMOGENEOUS, DIFF_FIELD):
            q('    const unsigned int tr = ty * MKB_RPT + r;')
            q('    const unsigned int iyg = iy + (unsigned int)g.iy_offset;')
            q('    const bool at_y0 = (iyg == 0), at_y1 = (iyg == nyg - 1);')
            if not paced_list:
                q('    const bool row_paced = pacing_on && (int)iyg >= (int)g.pace_y0 && (int)iyg < (int)g.pace_y1;')
        q('    #pragma unroll')
        q('    for (int c = 0; c < MKB_CPT; c++) {')
        q('    const unsigned int ix = ix0 + c;')
        q('    const unsigned long long cid = cid0 + c;')
        q('    (void)ix; (void)cid;')
        if diffusion_mode in (DIFF_HOMOGENEOUS, DIFF_FIELD):
            q('    const Real vcc = vc[r][c];')
            if stream:
                q('    // neighbours: registers inside the patch, its rim from the tile / the warp')
                q('    const Real vxm = (c > 0) ? vc[r][c > 0 ? c - 1 : 0] : v_lf[r];')
                q('    const Real vxp = (c < MKB_CPT - 1) ? vc[r][c < MKB_CPT - 1 ? c + 1 : 0] : v_rt[r];')
                q('    const Real vym = (r > 0) ? vc[r > 0 ? r - 1 : 0][c] : v_up[c];')
                q('    const Real vyp = (r < MKB_RPT - 1) ? vc[r < MKB_RPT - 1 ? r + 1 : 0][c] : v_dn[c];')
            else:
                q('    // neighbours: registers inside the patch, shared memory on its rim')
                q('    const Real vxm = (c > 0) ? vc[r][c > 0 ? c - 1 : 0] : col_r[tr][tx];')
                q('    const Real vxp = (c < MKB_CPT - 1) ? vc[r][c < MKB_CPT - 1 ? c + 1 : 0] : col_l[tr][tx + 2];')
                q('    const Real vym = (r > 0) ? vc[r > 0 ? r - 1 : 0][c] : row_b[ty][tx * MKB_CPT + c];')
                q('    const Real vyp = (r < MKB_RPT - 1) ? vc[r < MKB_RPT - 1 ? r + 1 : 0][c] : row_t[ty + 2][tx * MKB_CPT + c];')
            q('    Real idiff;')
            if diffusion_mode == DIFF_HOMOGENEOUS:
                q('    // openclsim.cl:401-434 (diff_step); nx >= MKB_CPT > 1 here. Only the')
                q('    // first / last cell of a patch can sit on the grid edge.')
                q('    if (c == 0 && at_x0) idiff = gx * (vcc - vxp);')
                q('    else if (c == MKB_CPT - 1 && at_x1) idiff = gx * (vcc - vxm);')
                q('    else idiff = gx * (2 * vcc - vxm - vxp);')
                q('    if (nyg > 1) {')
                q('        if (at_y0) idiff += gy * (vcc - vyp);')
                q('        else if (at_y1) idiff += gy * (vcc - vym);')
                q('        else idiff += gy * (2 * vcc - vym - vyp);')
                q('    }')
            else:
                q('    // openclsim.cl:469-486 (diff_hetero)')
                q('    idiff = 0.0;')
                q('    if (!(c == 0 && at_x0)) { idiff += gxf[cid - iy - 1] * (vcc - vxm); }')
                q('    if (!(c == MKB_CPT - 1 && at_x1)) { idiff += gxf[cid - iy] * (vcc - vxp); }')
                q('    if (nyg > 1) {')
                q('        if (!at_y0) idiff += gyf[(long long)cid - (long long)nx] * (vcc - vym);')
                q('        if (!at_y1) idiff += gyf[cid] * (vcc - vyp);')
                q('    }')
            if paced_list:
                q('    const Real pace = g.paced_mask[cid] ? pace_in : (Real)0;')
            else:
                q('    const Real pace = (row_paced && c >= pc0 && c < pc1) ? pace_in : (Real)0;')
            q('    if (store_aux) ((Real*)g.idiff)[cid] = idiff;')
        else:
            q('    const Real pace = pace_in;')
        q('    (void)pace;')
        for k, var in enumerate(fields):
            q('    const Real %s = F%d[r][c];' % (v(var), k))
        for line in consts:
            q(line)
        for var in states:
            if var.index() == i_vm:
                q('    const Real %s = vcc;' % v(var))
            else:
                q('    const Real %s = S%d[r][c];' % (v(var), var.index()))
        for name, eq in todo:
            if name:
                q('    // Component: %s' % name)
            var = eq.lhs.var()
            q('    const Real %s = %s;' % (v(eq.lhs), w.ex(eq.rhs)))
            if var in inter_index and not eq.lhs.is_derivative():
                q('    if (store_aux) ((Real*)g.inter)[%dull * stride + cid] = %s;'
                  % (inter_index[var], v(eq.lhs)))
        q('    // Update (openclsim.cl:358-364)')
        for var in states:
            if var in rl_states:
                inf, tau = rl_states[var]
                inf, tau, x = v(inf), v(tau), v(var)
                arg = '-dt / %s' % tau
                if fast_div and not sp and not native_maths:
                    arg = 'mkb_div(-dt, %s)' % tau
                rhs = '%s - (%s - %s) * %s(%s)' % (inf, inf, x, exp, arg)
            else:
                rhs = '%s + dt * %s' % (v(var), v(var.lhs()))
            q('    N%d[c] = %s;' % (var.index(), rhs))
        q('    }   // cells of this row')
        for var in states:
            if var.index() == i_vm:
                q('    mkb_vstore<MKB_CPT>(v_out + cid0, N%d);' % var.index())
            else:
                q('    mkb_vstore<MKB_CPT>(state + %dull * stride + cid0, N%d);'
                  % (var.index(), var.index()))
        q('    }   // rows of this thread')
        if stream:
            q('    }   // tiles of this thread block')
        q('}')
        q('')
        options = ['--fmad=true' if fmad else '--fmad=false']
        if max_registers:
            options.append('--maxrregcount=%d' % int(max_registers))
        ks = KernelSource('\n'.join(o), block, n_state, i_vm, len(inter_log),
                          len(fields), diffusion_mode, options)
        ks.cells_per_thread = cpt
        ks.rows_per_thread = rpt
        if stream:
            # MKB_KERNEL_STREAM | thread blocks per SM << 8; the TMA box
            blocks_per_sm = int(min_blocks or 2)
            ks.kernel_flags = 2 | (blocks_per_sm << 8)
            pad = 4 if sp else 2
            ks.stream_box = (bx * cpt + 2 * pad, by * rpt + 2)
        return ks

    # ------------------------------------------------------------------
    # Assembly
    # ------------------------------------------------------------------
    out = []
    p = out.append
    p('// Generated by myokit_b200.kernelgen for sm_100a — do not edit.')
    p('// Model: %s' % model.name())
    p('#include "mkb_device_abi.h"')
    p('typedef %s Real;' % real)
    p('#define MKB_BX %d' % bx)
    p('#define MKB_BY %d' % by)
    p('#define MKB_DIV_INT_CHECK %d' % (1 if div_int_check else 0))
    p('#define MKB_DIV_PARALLEL %d' % (1 if div_parallel else 0))
    p('#define MKB_DIV_CUBIC %d' % (1 if div_cubic else 0))
    p(math_defines.replace('@OVERLAP@', '1' if overlap else '0'))
    p(_PRELUDE)
    if pooled and w._pool:
        p('// Model constants (double precision), in order of first use')
        p('__constant__ double mkb_k[%d] = {' % len(w._pool))
        for x in w._pool:
            p('    %r,' % x)
        p('};')
        p('')
    stage_head = ((stage_group_of(max(len(stage_slot) - 1, 0)) + 1) * 8 + 127) // 128 * 128
    if stage and stage_slot:
        slot_plane_ = [None] * len(stage_slot)
        for var, j in stage_slot.items():
            slot_plane_[j] = var.index()
        p('// Staged states: slot -> state plane, in order of first use')
        p('#define MKB_STAGE_TILE (MKB_BX * MKB_BY)')
        p('#define MKB_STAGE_WARPS %d' % ((bx * by + 31) // 32))
        p('#define MKB_STAGE_HEAD %d    // the arrival barriers' % stage_head)
        p('#define MKB_STAGE_BYTES %d' % (stage_head + len(stage_slot) * bx * by * rs_))
        p('__constant__ unsigned short mkb_stage_plane[%d] = {%s};'
          % (len(slot_plane_), ', '.join(str(x) for x in slot_plane_)))
        p('__constant__ unsigned char mkb_stage_group[%d] = {%s};'
          % (len(slot_plane_), ', '.join(str(stage_group_of(j)) for j in range(len(slot_plane_)))))
        if tile_loop:
            p('#define MKB_LOOP_BYTES %d' % (
                128 + (len(stage_slot) + min(stage_sizes[0], len(stage_slot))) * bx * by * rs_))
        if not tile_loop:
            p('#define MKB_STAGE_AT(j) stage_c[(j) * MKB_STAGE_TILE]')
            p('#define MKB_STAGE_WAIT(k) MKB_MBAR_WAIT(&stage_bar[k], 0u)')
        p('')
    if tile_loop and stage_slot:
        kernel_lines = []
        fn_lines = _emit_tile_loop(kernel_lines.append, locals())
        for x in kernel_lines[:3]:      # (the macros of this form)
            p(x)
        for x in fn_lines:
            p(x)
        for x in kernel_lines[3:]:
            p(x)
        code = '\n'.join(out)
        options = ['--fmad=true' if fmad else '--fmad=false']
        ks = KernelSource(code, block, n_state, i_vm, len(inter_log),
                          len(fields), diffusion_mode, options)
        ks.plane_stride = int(plane_stride or 0)
        ne_ = stage_sizes[0]
        ks.smem_bytes = 128 + (len(stage_slot) + min(ne_, len(stage_slot))) * bx * by * rs_
        # MKB_KERNEL_STAGE | MKB_KERNEL_TILE_LOOP, thread blocks per SM
        ks.kernel_flags = 8 | 16 | (int(min_blocks or 2) << 8)
        return ks
    if min_blocks:
        p('extern "C" __global__ void __launch_bounds__(MKB_BX * MKB_BY, %d)'
          % int(min_blocks))
    elif max_registers:
        # (ptxas ignores --maxrregcount for kernels with launch bounds, and
        # derives only 128 / 96 / 80 ... registers from a block count)
        p('extern "C" __global__ void __maxnreg__(%d)' % int(max_registers))
    else:
        p('extern "C" __global__ void __launch_bounds__(MKB_BX * MKB_BY)')
    p('%s(const %sMkbGridArgs g, const MkbStepParams* __restrict__ sp,'
      % (KERNEL_NAME, '__grid_constant__ ' if stage else ''))
    p('    const Real* __restrict__ v_in, Real* __restrict__ v_out)')
    p('{')
    p('    // 32-bit indices (64-bit integer arithmetic is emulated on the GPU);')
    p('    // only the flat cell id, which can pass 2^32, is 64-bit. The launch')
    p('    // grid is (column blocks, row blocks mod 32768, row blocks / 32768).')
    p('    const unsigned int tx = threadIdx.x, ty = threadIdx.y;')
    p('    const unsigned int nx = (unsigned int)g.nx, ny = (unsigned int)g.ny;')
    if plane_stride:
        p('    // compiled for this plane stride (see MKB_AT)')
        p('    constexpr unsigned long long stride = %dull;' % int(plane_stride))
        p('    if (g.stride != stride) { MKB_STRIDE_MISMATCH(); return; }')
    else:
        p('    const unsigned long long stride = g.stride;')
    p('    const unsigned int nby = (ny + MKB_BY - 1) / MKB_BY;')
    p('    const unsigned int byr = blockIdx.y + blockIdx.z * gridDim.y;')
    p('    if (byr >= nby) return;')
    p('    const unsigned int bxb = blockIdx.x;')
    p('    const unsigned int ix = bxb * MKB_BX + tx;')
    if slab:
        p('    // Boundary row blocks first: their rows reach the neighbouring GPUs')
        p('    // while the interior is still being computed.')
        p('    const unsigned int byb = (byr == 0) ? 0 : ((byr == 1) ? nby - 1 : byr - 1);')
        p('    const unsigned int iy = byb * MKB_BY + ty;')
        p('    const unsigned int step = sp->step;')
        p('    const bool wait_lo = (byb == 0) && g.halo_lo;')
        p('    const bool wait_hi = (byb == nby - 1) && g.halo_hi;')
        p('    if (wait_lo || wait_hi) {')
        p('        if (tx == 0 && ty == 0) {')
        p('            if (wait_lo) mkb_wait_flag(g.flag_lo + bxb, step, g.halo_error);')
        p('            if (wait_hi) mkb_wait_flag(g.flag_hi + bxb, step, g.halo_error);')
        p('            __threadfence();')
        p('        }')
        p('        __syncthreads();')
        p('    }')
    else:
        p('    const unsigned int iy = byr * MKB_BY + ty;')
    if overlap:
        p('    // Steps overlap: the next kernel may start launching now; this block')
        p('    // waits for step - 1 of its own tile and its four neighbours.')
        p('    MKB_PDL_TRIGGER();')
        p('    const unsigned int tile_row = %s;' % ('byb' if slab else 'byr'))
        p('    {')
        p('        const unsigned int t5 = ty * MKB_BX + tx;')
        p('        if (t5 < 5u) {')
        p('            const int nbx_ = (int)bxb + (t5 == 1u ? -1 : (t5 == 2u ? 1 : 0));')
        p('            const int nby_ = (int)tile_row + (t5 == 3u ? -1 : (t5 == 4u ? 1 : 0));')
        p('            if (nbx_ >= 0 && nbx_ < (int)gridDim.x && nby_ >= 0 && nby_ < (int)nby)')
        p('                mkb_wait_tile(g.tile_done + (unsigned int)nby_ * gridDim.x + (unsigned int)nbx_,')
        p('                              sp->step - 1u, g.halo_error);')
        p('        }')
        p('        __syncthreads();')
        p('    }')
    if partitioned:
        p('    // Ghost cells: V(t) of this step must have arrived from every')
        p('    // exporting GPU before any thread of this block gathers it.')
        p('    if (g.n_ghost_import) {')
        p('        const unsigned int nimp = (unsigned int)g.n_ghost_import;')
        p('        for (unsigned int k = ty * MKB_BX + tx; k < nimp; k += MKB_BX * MKB_BY) {')
        p('            mkb_wait_flag(g.ghost_flags + g.ghost_import[k], sp->step, g.halo_error);')
        p('            __threadfence();')
        p('        }')
        p('        __syncthreads();')
        p('    }')
    n_slots = len(stage_slot)
    group_of_slot = [stage_group_of(j) for j in range(n_slots)]
    n_groups = (max(group_of_slot) + 1) if n_slots else 0
    def stage_prologue():
        if not (stage and n_slots):
            return
        p('    // Staged states: the tile of every state plane but V arrives in shared')
        p('    // memory by TMA, one box per plane, in the order the equations first')
        p('    // need them, announced on %d arrival barrier(s). One thread sets the' % n_groups)
        p('    // barriers up, then the first lane of every warp issues its share of')
        p('    // the boxes; a thread reads and updates only its own element of each')
        p('    // plane, and the tiles go back by TMA at the end.')
        p('    MKB_STAGE_DECL(MKB_STAGE_BYTES);')
        p('    unsigned long long* const stage_bar = (unsigned long long*)mkb_stage_mem;')
        p('    Real* const stage_base = (Real*)(mkb_stage_mem + MKB_STAGE_HEAD);')
        p('    Real* const stage_c = stage_base + ty * MKB_BX + tx;')
        p('    {')
        p('        const unsigned int t_ = ty * MKB_BX + tx;')
        p('        if (t_ == 0) {')
        p('            for (int k_ = 0; k_ < %d; k_++) MKB_MBAR_INIT(&stage_bar[k_], 1);' % n_groups)
        p('            MKB_MBAR_FENCE_INIT();')
        for gi in range(n_groups):
            p('            MKB_MBAR_EXPECT_TX(&stage_bar[%d], %du * MKB_STAGE_TILE * (unsigned int)sizeof(Real));'
              % (gi, group_of_slot.count(gi)))
        p('        }')
        p('        __syncthreads();')
        p('        if ((t_ & 31u) == 0u) {')
        if overlap:
            p('            // (the tiles were written by the TMA unit of the previous step,')
            p('            // which this block has just observed through tile_done)')
            p('            MKB_FENCE_ASYNC_GLOBAL();')
        p('            for (unsigned int j_ = t_ >> 5; j_ < %du; j_ += MKB_STAGE_WARPS)' % n_slots)
        p('                MKB_TMA_LOAD_3D(stage_base + j_ * MKB_STAGE_TILE, g.tmap_state,')
        p('                                (int)(ix - tx), (int)(iy - ty), (int)mkb_stage_plane[j_],')
        p('                                &stage_bar[mkb_stage_group[j_]]);')
        p('        }')
        p('    }')
    p('    const bool active = (ix < nx) && (iy < ny);')
    p('    const unsigned long long cid = (unsigned long long)iy * nx + ix;')
    p('    Real* const state = (Real*)g.state;')
    p('    Real* const state_c = state + cid;')
    p('    const Real* const field_c = (const Real*)g.field + cid;')
    p('    Real* const inter_c = (Real*)g.inter + cid;')
    p('    (void)state_c; (void)field_c; (void)inter_c;')
    p('    // Per-step scalars, cast like openclsim.c:1063,1148,1155')
    p('    const Real time = (Real)sp->time;')
    p('    const Real dt = (Real)sp->dt;')
    p('    const Real pace_in = (Real)sp->pace;')
    p('    const bool store_aux = (sp->flags & MKB_FLAG_STORE_AUX) != 0;')
    p('    (void)time; (void)pace_in; (void)store_aux; (void)v_in; (void)v_out;')
    p('')
    if diffusion:
        p('    const Real vc = active ? MKB_LD(v_in + cid) : (Real)0;')
    if diffusion_mode in (DIFF_HOMOGENEOUS, DIFF_FIELD):
        p('    const unsigned int iyg = iy + (unsigned int)g.iy_offset;  // global row')
        p('    const unsigned int nyg = (unsigned int)g.ny_global;')
    if diffusion_mode == DIFF_FIELD:
        p('    // Edge conductances, gx[(ny, nx-1)], gy[(ny-1, nx)] (openclsim.cl:')
        p('    // 475-482); both pointers are slab-relative: gyf[-nx .. -1] is the')
        p('    // gy row that couples this slab\'s first row to the slab above it.')
        p('    // Loaded here so they are in flight together with everything else.')
        p('    const Real* const gxf = (const Real*)g.gx_field;')
        p('    const Real* const gyf = (const Real*)g.gy_field;')
        p('    const bool has_xm = active && nx > 1 && ix > 0;')
        p('    const bool has_xp = active && nx > 1 && ix < nx - 1;')
        p('    const bool has_ym = active && nyg > 1 && iyg > 0;')
        p('    const bool has_yp = active && nyg > 1 && iyg < nyg - 1;')
        p('    const Real gxm = has_xm ? gxf[cid - iy - 1] : (Real)0;')
        p('    const Real gxp = has_xp ? gxf[cid - iy] : (Real)0;')
        p('    const Real gym = has_ym ? gyf[(long long)cid - (long long)nx] : (Real)0;')
        p('    const Real gyp = has_yp ? gyf[cid] : (Real)0;')
    p('    // Loads issued ahead of use: fields and the first states')
    for line in early:
        p(line)
    if stage and prefetch_next and diffusion_mode in (DIFF_HOMOGENEOUS, DIFF_FIELD) and (
            fields or diffusion_mode == DIFF_FIELD):
        p('    // Hint (see the tile prefetch further down): conductances and fields of')
        p('    // the cells whose blocks start a wave from now, as far as L2')
        p('    {')
        p('        unsigned int nsm_;')
        p('        MKB_NSM(nsm_);')
        p('        const unsigned int r_ = ((nsm_ * %du + gridDim.x - 1u) / gridDim.x) * MKB_BY;'
          % int(min_blocks or 2))
        p('        if (active && iy + r_ < ny) {')
        if diffusion_mode == DIFF_FIELD:
            p('            if (ix < nx - 1) MKB_PREFETCH_L2(gxf + (cid - iy) + (unsigned long long)r_ * (nx - 1));')
            p('            MKB_PREFETCH_L2(gyf + cid + (unsigned long long)r_ * nx);')
        for k in range(len(fields)):
            p('            MKB_PREFETCH_L2(&MKB_AT(field_c, %d) + (unsigned long long)r_ * nx);' % k)
        p('        }')
        p('    }')
    if prefetch and lazy_state and not stage:
        p('    // The other states: only prefetched here, loaded where they are used')
        p('    if (active) {')
        for var in states:
            if var.index() != i_vm and var not in early_states and var not in gate_set_unused:
                p('        MKB_PREFETCH_%s(&MKB_AT(state_c, %d));'
                  % ('L1' if (prefetch == 'l1' and not overlap) else 'L2', var.index()))
        p('    }')
    p('')

    v_direct = bool(v_direct) and stage and n_slots and diffusion_mode in (
        DIFF_HOMOGENEOUS, DIFF_FIELD) and not stab
    if v_direct:
        # No V tile: a thread takes the potentials left and right of its cell
        # from the neighbouring lanes of its warp (the first and last lane
        # from memory) and those above and below from memory (rows another
        # warp of the block loads at the same moment: one request to L2), so
        # nothing is written to shared memory, read back, or waited for at a
        # block barrier before the first equation.
        lo_src = ('__ldcg((const Real*)g.halo_lo + (step % 3u) * nx + ix)' if slab
                  else '((const Real*)g.halo_lo)[ix]')
        hi_src = ('__ldcg((const Real*)g.halo_hi + (step % 3u) * nx + ix)' if slab
                  else '((const Real*)g.halo_hi)[ix]')
        p('    // V(t) of the four neighbours: lanes of the warp left and right, memory')
        p('    // above and below. A neighbour outside the grid leaves the cell\'s own V,')
        p('    // which the edge formulas below never use (as openclsim.cl).')
        p('    Real vxm = vc, vxp = vc, vym = vc, vyp = vc;')
        p('    {')
        p('        const unsigned int lane_ = (ty * MKB_BX + tx) & 31u;')
        p('        const Real left_ = MKB_SHFL_UP(vc, 1), right_ = MKB_SHFL_DOWN(vc, 1);')
        p('        if (active) {')
        p('            if (ix > 0) vxm = (lane_ > 0u && tx > 0u) ? left_ : MKB_LD(v_in + cid - 1);')
        p('            if (ix < nx - 1) vxp = (lane_ < 31u && tx < MKB_BX - 1) ? right_ : MKB_LD(v_in + cid + 1);')
        p('            if (iy > 0) vym = MKB_LD(v_in + cid - nx);')
        p('            else if (iyg > 0 && g.halo_lo) vym = %s;' % lo_src)
        p('            if (iy < ny - 1) vyp = MKB_LD(v_in + cid + nx);')
        p('            else if (iyg < nyg - 1 && g.halo_hi) vyp = %s;' % hi_src)
        p('        }')
        p('    }')
        stage_prologue()
        if slab and not slab_lean:
            p('    if (active) {')
        else:
            p('    if (!active) return;')
    if diffusion_mode in (DIFF_HOMOGENEOUS, DIFF_FIELD) and not v_direct:
        p('    // V(t) tile + one-cell halo in shared memory. Out-of-grid halo')
        p('    // entries hold the cell\'s own V and are never used: the edge')
        p('    // formulas below drop those terms exactly as openclsim.cl does.')
        p('    __shared__ Real tile[MKB_BY + 2][MKB_BX + 2];')
        p('    // (only cells of the grid: the slot of a thread beyond the last row or')
        p('    // column is the halo slot of its neighbour, written below)')
        lo_src = ('__ldcg((const Real*)g.halo_lo + (step % 3u) * nx + ix)' if slab
                  else '((const Real*)g.halo_lo)[ix]')
        hi_src = ('__ldcg((const Real*)g.halo_hi + (step % 3u) * nx + ix)' if slab
                  else '((const Real*)g.halo_hi)[ix]')
        if stage and n_slots:
            # every load from HBM is in flight before the block meets to set
            # up the arrival barriers; the tile is written after
            p('    const bool rim_l = active && tx == 0, rim_r = active && (tx == MKB_BX - 1 || ix == nx - 1);')
            p('    const bool rim_u = active && ty == 0, rim_d = active && (ty == MKB_BY - 1 || iy == ny - 1);')
            p('    Real vrl = vc, vrr = vc, vru = vc, vrd = vc;')
            p('    if (rim_l && ix > 0) vrl = MKB_LD(v_in + cid - 1);')
            p('    if (rim_r && ix < nx - 1) vrr = MKB_LD(v_in + cid + 1);')
            p('    if (rim_u) {')
            p('        if (iy > 0) vru = MKB_LD(v_in + cid - nx);')
            p('        else if (iyg > 0 && g.halo_lo) vru = %s;' % lo_src)
            p('    }')
            p('    if (rim_d) {')
            p('        if (iy < ny - 1) vrd = MKB_LD(v_in + cid + nx);')
            p('        else if (iyg < nyg - 1 && g.halo_hi) vrd = %s;' % hi_src)
            p('    }')
            stage_prologue()
            p('    if (active) tile[ty + 1][tx + 1] = vc;')
            p('    if (rim_l) tile[ty + 1][0] = vrl;')
            p('    if (rim_r) tile[ty + 1][tx + 2] = vrr;')
            p('    if (rim_u) tile[0][tx + 1] = vru;')
            p('    if (rim_d) tile[ty + 2][tx + 1] = vrd;')
        else:
            p('    if (active) tile[ty + 1][tx + 1] = vc;')
            p('    if (active) {')
            p('        if (tx == 0) tile[ty + 1][0] = (ix > 0) ? MKB_LD(v_in + cid - 1) : vc;')
            p('        if (tx == MKB_BX - 1 || ix == nx - 1)')
            p('            tile[ty + 1][tx + 2] = (ix < nx - 1) ? MKB_LD(v_in + cid + 1) : vc;')
            p('        if (ty == 0) {')
            p('            Real vn = vc;')
            p('            if (iy > 0) vn = MKB_LD(v_in + cid - nx);')
            p('            else if (iyg > 0 && g.halo_lo) vn = %s;' % lo_src)
            p('            tile[0][tx + 1] = vn;')
            p('        }')
            p('        if (ty == MKB_BY - 1 || iy == ny - 1) {')
            p('            Real vn = vc;')
            p('            if (iy < ny - 1) vn = MKB_LD(v_in + cid + nx);')
            p('            else if (iyg < nyg - 1 && g.halo_hi) vn = %s;' % hi_src)
            p('            tile[ty + 2][tx + 1] = vn;')
            p('        }')
            p('    }')
        if stab:
            p('    MKB_EXP_TABLE_INIT(ty * MKB_BX + tx, MKB_BX * MKB_BY);')
        p('    __syncthreads();')
        if slab and not slab_lean:
            p('    if (active) {')
        else:
            p('    if (!active) return;')
        p('    const Real vxm = tile[ty + 1][tx], vxp = tile[ty + 1][tx + 2];')
        p('    const Real vym = tile[ty][tx + 1], vyp = tile[ty + 2][tx + 1];')
    if diffusion_mode in (DIFF_HOMOGENEOUS, DIFF_FIELD):
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
            if junction:
                p('    // openclsim.cl:617-627 (diff_step_fiber_tissue), this grid\'s side')
                p('    if (g.junction_v0 && ix == (unsigned int)g.jx && iy >= (unsigned int)g.jy0')
                p('            && iy < (unsigned int)(g.jy0 + g.jn)) {')
                p('        const Real* const other = (const Real*)((v_in == (const Real*)g.state + %dull * stride)' % i_vm)
                p('            ? g.junction_v0 : g.junction_v1);')
                p('        const Real vo = other[g.joff + (unsigned long long)(iy - (unsigned int)g.jy0) * g.jstride];')
                if junction == 'fiber':
                    p('        idiff += (Real)g.jg * (vc - vo);')
                else:
                    p('        idiff -= (Real)g.jg * (vo - vc);')
                p('    }')
        else:
            p('    // openclsim.cl:469-486 (diff_hetero)')
            p('    idiff = 0.0;')
            p('    if (has_xm) { idiff += gxm * (vc - vxm); }')
            p('    if (has_xp) { idiff += gxp * (vc - vxp); }')
            p('    if (has_ym) idiff += gym * (vc - vym);')
            p('    if (has_yp) idiff += gyp * (vc - vyp);')
    elif diffusion_mode == DIFF_CONNECTIONS:
        if stab:
            p('    MKB_EXP_TABLE_INIT(ty * MKB_BX + tx, MKB_BX * MKB_BY);')
            p('    __syncthreads();')
        p('    if (!active) return;')
        p('    // openclsim.cl:537-556 as a per-cell CSR gather: same terms')
        p('    // g * (V_i - V_j), summed in edge-list order, no atomics.')
        p('    Real idiff = 0;')
        p('    {')
        p('        const Real* const cg = (const Real*)g.csr_g;')
        p('        const unsigned long long e1 = g.csr_row[cid + 1];')
        if partitioned:
            p('        // columns >= nx are ghost cells: V(t) pushed here by their owners')
            p('        const Real* const ghost = (const Real*)g.ghost + (unsigned long long)(sp->step % 3u) * g.n_ghost;')
            p('        for (unsigned long long e = g.csr_row[cid]; e < e1; e++) {')
            p('            const unsigned int col = g.csr_col[e];')
            p('            const Real vn = (col < nx) ? v_in[col] : __ldcg(ghost + (col - nx));')
            p('            idiff += cg[e] * (vc - vn);')
            p('        }')
        else:
            p('        for (unsigned long long e = g.csr_row[cid]; e < e1; e++) {')
            p('            idiff += cg[e] * (vc - v_in[g.csr_col[e]]);')
            p('        }')
        p('    }')
    else:
        if stab:
            p('    MKB_EXP_TABLE_INIT(ty * MKB_BX + tx, MKB_BX * MKB_BY);')
            p('    __syncthreads();')
        elif stage and n_slots:
            stage_prologue()
        p('    if (!active) return;')
    p('')

    if diffusion:
        p('    // openclsim.cl:249-280, 322-329')
        if paced_list:
            p('    const Real pace = g.paced_mask[cid] ? pace_in : (Real)0;')
        else:
            if diffusion_mode == DIFF_CONNECTIONS:
                p('    const int pix = (int)ix, piy = 0;')
            else:
                p('    const int pix = (int)ix, piy = (int)iyg;')
            p('    const Real pace = (pix >= (int)g.pace_x0 && pix < (int)g.pace_x1 &&')
            p('                       piy >= (int)g.pace_y0 && piy < (int)g.pace_y1) ? pace_in : (Real)0;')
        p('    if (store_aux) ((Real*)g.idiff)[cid] = idiff;')
    else:
        p('    const Real pace = pace_in;')
    p('    (void)pace;')
    p('')
    for line in consts:
        p(line)
    p('')
    for line in body:
        if line == '@PREFETCH_NEXT@':
            # The thread blocks that start when this wave retires find their
            # tiles in L2: about a quarter of a warp's residence was the wait
            # for its first loads from HBM (ncu, profiles/r03_summary.md). The
            # grid is handed out in launch order, x fastest: the tile a wave
            # ahead is `resident blocks / column blocks` block rows down.
            p('    // Hint: bring the tiles of the blocks that start a wave from now as far as L2')
            p('    {')
            p('        unsigned int tx_, ty_;')
            p('        MKB_ASM_SREG(tx_, "tid.x"); MKB_ASM_SREG(ty_, "tid.y");')
            p('        const unsigned int t_ = ty_ * MKB_BX + tx_;')
            p('        if ((t_ & 31u) == 0u) {')
            p('            unsigned int bx_, by_, bz_, gy_, gx_, nsm_;')
            p('            MKB_ASM_SREG(bx_, "ctaid.x"); MKB_ASM_SREG(by_, "ctaid.y"); MKB_ASM_SREG(bz_, "ctaid.z");')
            p('            MKB_ASM_SREG(gy_, "nctaid.y"); MKB_ASM_SREG(gx_, "nctaid.x");')
            p('            MKB_NSM(nsm_);')
            p('            const unsigned int row_ = by_ + bz_ * gy_ + (nsm_ * %du + gx_ - 1u) / gx_;'
              % int(min_blocks or 2))
            p('            const int x_ = (int)(bx_ * MKB_BX), y_ = (int)(row_ * MKB_BY);')
            p('            for (unsigned int j_ = t_ >> 5; j_ < %du; j_ += MKB_STAGE_WARPS)' % n_slots)
            p('                MKB_TMA_PREFETCH_3D(g.tmap_state, x_, y_, (int)mkb_stage_plane[j_]);')
            if diffusion and i_vm >= 0:
                p('            // (V(t): the plane inside `state`, or the second V plane behind the states)')
                p('            if (t_ == 0u) MKB_TMA_PREFETCH_3D(g.tmap_state, x_, y_,')
                p('                (v_in == (const Real*)g.state + %dull * stride) ? %d : %d);'
                  % (i_vm, i_vm, n_state))
            p('        }')
            p('    }')
            continue
        p(line)

    def stage_epilogue():
        if not (stage and n_slots and stage_store):
            return
        p('    // Staged states go back: every writer makes its shared-memory stores')
        p('    // visible to the TMA unit, the threads still here meet, and one thread')
        p('    // issues a box per plane and waits until they have been written')
        p('    // (cells outside the grid are clipped). Coordinates are read again.')
        p('    MKB_FENCE_ASYNC_SMEM();')
        p('    __syncthreads();')
        p('    {')
        p('        unsigned int tx_, ty_;')
        p('        MKB_ASM_SREG(tx_, "tid.x"); MKB_ASM_SREG(ty_, "tid.y");')
        p('        const unsigned int t_ = ty_ * MKB_BX + tx_;')
        p('        if ((t_ & 31u) == 0u) {')
        p('            unsigned int bx_, by_, bz_, gy_;')
        p('            MKB_ASM_SREG(bx_, "ctaid.x"); MKB_ASM_SREG(by_, "ctaid.y"); MKB_ASM_SREG(bz_, "ctaid.z");')
        p('            MKB_ASM_SREG(gy_, "nctaid.y");')
        p('            unsigned int row_ = by_ + bz_ * gy_;')
        if slab:
            p('            const unsigned int nby_ = ((unsigned int)g.ny + MKB_BY - 1) / MKB_BY;')
            p('            row_ = (row_ == 0) ? 0 : ((row_ == 1) ? nby_ - 1 : row_ - 1);')
        p('            // a tile inside the grid: every warp is still here and takes its')
        p('            // share; a tile across the rim: thread (0, 0), which always is')
        p('            const bool full_ = (bx_ + 1u) * MKB_BX <= (unsigned int)g.nx')
        p('                && (row_ + 1u) * MKB_BY <= (unsigned int)g.ny;')
        p('            const unsigned int j0_ = full_ ? (t_ >> 5) : 0u, dj_ = full_ ? MKB_STAGE_WARPS : 1u;')
        p('            if (full_ || t_ == 0u) {')
        p('                Real* const base_ = (Real*)(mkb_stage_mem + MKB_STAGE_HEAD);')
        p('                for (unsigned int j_ = j0_; j_ < %du; j_ += dj_)' % n_slots)
        p('                    MKB_TMA_STORE_3D(g.tmap_state, (int)(bx_ * MKB_BX), (int)(row_ * MKB_BY),')
        p('                                     (int)mkb_stage_plane[j_], base_ + j_ * MKB_STAGE_TILE);')
        if overlap:
            p('                // (written, and visible to the next step\'s TMA loads, before')
            p('                // the tile is published)')
            p('                MKB_TMA_STORE_FINISH();')
            p('                MKB_FENCE_ASYNC_GLOBAL();')
        else:
            p('                // (the shared memory has been read; the kernel boundary')
            p('                // completes the writes)')
            p('                MKB_TMA_STORE_READ_DONE();')
        p('            }')
        p('        }')
        p('    }')
    if not (slab and not slab_lean):
        stage_epilogue()
    if slab and slab_lean:
        p('    // Publish the boundary rows: data first, then (after a system-scope')
        p('    // fence by every writer and a barrier of the threads still here)')
        p('    // the arrival flag. Coordinates and step are read again.')
        p('    const MkbSlabPos q = mkb_slab_pos<MKB_BY>((unsigned int)g.ny);')
        p('    const bool send_lo = (q.byb == 0) && g.peer_lo_halo_hi;')
        p('    const bool send_hi = (q.byb == q.nby - 1) && g.peer_hi_halo_lo;')
        p('    if (send_lo || send_hi) {')
        p('        __threadfence_system();')
        p('        __syncthreads();')
        p('        // (thread (0, 0) owns the block\'s lowest cell: it is in the grid)')
        p('        if (q.tx == 0 && q.ty == 0) {')
        p('            const unsigned int next = *(const volatile unsigned int*)&sp->step + 1u;')
        p('            if (send_lo) *((volatile unsigned int*)g.peer_lo_flag_hi + q.bxb) = next;')
        p('            if (send_hi) *((volatile unsigned int*)g.peer_hi_flag_lo + q.bxb) = next;')
        p('        }')
        p('    }')
    elif slab:
        p('    }   // active')
        stage_epilogue()
        p('    // Publish the boundary rows: data first, then (after a system-scope')
        p('    // fence by every writer and a CTA barrier) the arrival flag.')
        p('    const bool send_lo = (byb == 0) && g.peer_lo_halo_hi;')
        p('    const bool send_hi = (byb == nby - 1) && g.peer_hi_halo_lo;')
        p('    if (send_lo || send_hi) {')
        p('        __threadfence_system();')
        p('        __syncthreads();')
        p('        if (tx == 0 && ty == 0) {')
        p('            if (send_lo) *((volatile unsigned int*)g.peer_lo_flag_hi + bxb) = step + 1u;')
        p('            if (send_hi) *((volatile unsigned int*)g.peer_hi_flag_lo + bxb) = step + 1u;')
        p('        }')
        p('    }')
    if overlap:
        p('    // This tile has taken its step: stores first (barrier of the threads')
        p('    // still here, device-scope release), then the flag. Coordinates are')
        p('    // read again rather than kept in registers across the model.')
        p('    __syncthreads();')
        p('    {')
        p('        unsigned int tx_, ty_;')
        p('        MKB_ASM_SREG(tx_, "tid.x"); MKB_ASM_SREG(ty_, "tid.y");')
        p('        if (tx_ == 0 && ty_ == 0) {')
        p('            unsigned int bx_, by_, bz_, gy_, gx_;')
        p('            MKB_ASM_SREG(bx_, "ctaid.x"); MKB_ASM_SREG(by_, "ctaid.y"); MKB_ASM_SREG(bz_, "ctaid.z");')
        p('            MKB_ASM_SREG(gy_, "nctaid.y"); MKB_ASM_SREG(gx_, "nctaid.x");')
        p('            unsigned int row_ = by_ + bz_ * gy_;')
        if slab:
            p('            const unsigned int nby_ = ((unsigned int)g.ny + MKB_BY - 1) / MKB_BY;')
            p('            row_ = (row_ == 0) ? 0 : ((row_ == 1) ? nby_ - 1 : row_ - 1);')
        p('            mkb_publish_tile(g.tile_done + row_ * gx_ + bx_, *(const volatile unsigned int*)&sp->step);')
        p('        }')
        p('    }')
    p('}')
    p('')
    if gate_states:
        # The second kernel of a split step: runs after mkb_cell_step on the
        # same stream, reads V(t) from the plane that kernel only read, and
        # the gates it alone updates.
        p('// Gating variables whose rates depend on V alone (%d of %d states):'
          % (len(gate_states), n_state))
        p('// ' + ', '.join(x.qname() for x in gate_states))
        p('extern "C" __global__ void __launch_bounds__(MKB_BX * MKB_BY)')
        p('mkb_gate_step(const MkbGridArgs g, const MkbStepParams* __restrict__ sp,')
        p('    const Real* __restrict__ v_in, Real* __restrict__ v_out)')
        p('{')
        p('    const unsigned int nx = (unsigned int)g.nx, ny = (unsigned int)g.ny;')
        p('    const unsigned long long stride = g.stride;')
        p('    const unsigned int nby = (ny + MKB_BY - 1) / MKB_BY;')
        p('    const unsigned int byr = blockIdx.y + blockIdx.z * gridDim.y;')
        p('    if (byr >= nby) return;')
        p('    const unsigned int ix = blockIdx.x * MKB_BX + threadIdx.x;')
        p('    const unsigned int iy = byr * MKB_BY + threadIdx.y;')
        if stab:
            p('    MKB_EXP_TABLE_INIT(threadIdx.y * MKB_BX + threadIdx.x, MKB_BX * MKB_BY);')
            p('    __syncthreads();')
        p('    if (ix >= nx || iy >= ny) return;')
        p('    const unsigned long long cid = (unsigned long long)iy * nx + ix;')
        p('    Real* const state = (Real*)g.state;')
        p('    Real* const state_c = state + cid;')
        p('    const Real* const field_c = (const Real*)g.field + cid;')
        p('    Real* const inter_c = (Real*)g.inter + cid;')
        p('    (void)state_c; (void)field_c; (void)inter_c; (void)state;')
        p('    const Real time = (Real)sp->time;')
        p('    const Real dt = (Real)sp->dt;')
        p('    const bool store_aux = (sp->flags & MKB_FLAG_STORE_AUX) != 0;')
        p('    (void)time; (void)store_aux; (void)v_in; (void)v_out;')
        for line in gate_lines:
            p(line)
        p('}')
        p('')
    code = '\n'.join(out)

    options = ['--fmad=true' if fmad else '--fmad=false']
    if max_registers:
        options.append('--maxrregcount=%d' % int(max_registers))
    ks = KernelSource(code, block, n_state, i_vm, len(inter_log),
                      len(fields), diffusion_mode, options)
    ks.gate_kernel = bool(gate_states)
    ks.gate_states = [x.qname() for x in gate_states]
    if overlap:
        ks.kernel_flags |= 4            # MKB_KERNEL_OVERLAP
    ks.plane_stride = int(plane_stride or 0)
    if stage and stage_slot:
        ks.kernel_flags |= 8            # MKB_KERNEL_STAGE
        ks.smem_bytes = stage_head + len(stage_slot) * bx * by * rs_
    return ks
