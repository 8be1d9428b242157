/*
 * runner.cpp — TEST INFRASTRUCTURE ONLY (see mkb_cuda_shim.h).
 *
 * Compiled once per generated kernel: -DMKB_KERNEL_FILE="\"kernel.cu\"". Lays
 * the data out exactly as mkb_runtime.cu does on the device (state planes of
 * `stride` Reals, second V plane, idiff / inter / field planes, conductance
 * fields with the extra leading gy row, CSR in edge-list order), then runs the
 * given steps, every thread block in turn, every thread of a block as a fiber.
 */
#include <ucontext.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mkb_cuda_shim.h"
#include MKB_KERNEL_FILE

namespace {

struct Fiber {
    ucontext_t ctx;
    bool done = false, waiting = false;
    unsigned int tx = 0, ty = 0;
};

ucontext_t g_sched;
Fiber* g_cur = nullptr;
int g_or_acc = 0, g_or_res = 0;
int g_reverse = 0;
int g_persistent = 0;
int g_stream_blocks = 0;     // > 0: persistent-grid streaming kernel, this many blocks

struct Launch {
    MkbGridArgs g;
    const MkbStepParams* sp;
    const Real* v_in;
    Real* v_out;
} g_launch;

#ifndef MKB_KERNEL_FN
#define MKB_KERNEL_FN mkb_cell_step
#endif

// a split step has a second kernel, launched after the first over the same grid
int g_second = 0;

void fiber_main() {
#ifdef MKB_KERNEL_FN2
    if (g_second) {
        MKB_KERNEL_FN2(g_launch.g, g_launch.sp, g_launch.v_in, g_launch.v_out);
        g_cur->done = true;
        swapcontext(&g_cur->ctx, &g_sched);
    }
#endif
    MKB_KERNEL_FN(g_launch.g, g_launch.sp, g_launch.v_in, g_launch.v_out);
    g_cur->done = true;
    swapcontext(&g_cur->ctx, &g_sched);
}

}   // namespace

void shim_barrier() {
    g_cur->waiting = true;
    swapcontext(&g_cur->ctx, &g_sched);
}

// a thread that polls (an arrival barrier another thread has yet to serve)
// lets the others run and comes back
void shim_yield() {
    swapcontext(&g_cur->ctx, &g_sched);
}

int shim_barrier_or(int pred) {
    g_or_acc |= (pred != 0);
    shim_barrier();
    return g_or_res;
}

namespace {

const size_t kStack = 1 << 20;

void run_block(std::vector<Fiber>& fibers, std::vector<char>& stacks, unsigned int bx_, unsigned int by_) {
    const unsigned int nthreads = bx_ * by_;
    for (unsigned int t = 0; t < nthreads; t++) {
        Fiber& f = fibers[t];
        f.done = f.waiting = false;
        f.tx = t % bx_;
        f.ty = t / bx_;
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = stacks.data() + (size_t)t * kStack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = &g_sched;
        makecontext(&f.ctx, fiber_main, 0);
    }
    g_or_acc = g_or_res = 0;
    for (;;) {
        bool alive = false;
        for (unsigned int k = 0; k < nthreads; k++) {
            // thread order within a barrier phase is not defined on the GPU:
            // tests run both directions and require identical results
            const unsigned int t = g_reverse ? nthreads - 1 - k : k;
            Fiber& f = fibers[t];
            if (f.done || f.waiting) continue;
            threadIdx.x = f.tx;
            threadIdx.y = f.ty;
            threadIdx.z = 0;
            g_cur = &f;
            swapcontext(&g_sched, &f.ctx);
        }
        bool polling = false;
        for (unsigned int t = 0; t < nthreads; t++) {
            alive = alive || !fibers[t].done;
            polling = polling || (!fibers[t].done && !fibers[t].waiting);
        }
        if (!alive) break;
        if (polling) continue;      // somebody yielded: another round before any barrier opens
        // every thread that has not exited is waiting: release the barrier
        g_or_res = g_or_acc;
        g_or_acc = 0;
        for (unsigned int t = 0; t < nthreads; t++) fibers[t].waiting = false;
    }
}

}   // namespace

extern "C" int shim_real_size(void) { return (int)sizeof(Real); }
extern "C" void shim_set_thread_order(int reverse) { g_reverse = reverse; }
// persistent kernels: consecutive unlogged steps share one launch (as
// sim_step_typed groups them); the run length rides in flags >> 8
extern "C" void shim_set_persistent(int on) { g_persistent = on; }
extern "C" void shim_set_stream_blocks(int n) { g_stream_blocks = n; }
double shim_shfl_buf[1024];

extern "C" int shim_run(
    int nx, int ny, int n_state, int i_vm, int n_inter, int n_field, int diffusion_mode,
    double gx, double gy, const double* gx_field, const double* gy_field,
    long long px0, long long px1, long long py0, long long py1, const unsigned char* paced_mask,
    unsigned long long n_conn, const unsigned long long* conn_i, const unsigned long long* conn_j,
    const double* conn_g,
    int n_steps, const double* st_time, const double* st_dt, const double* st_pace,
    const unsigned char* st_log,
    double* state_aos, const double* field_aos, double* log_v, double* log_idiff, double* log_inter,
    int block_x, int block_y, int cpt, int rpt)
{
    const size_t n = (size_t)nx * ny;
    const size_t stride = (n + 31) / 32 * 32;
    std::vector<Real> state((size_t)n_state * stride, (Real)0), v_alt(stride, (Real)0);
    std::vector<Real> idiff(stride, (Real)0), inter((size_t)(n_inter > 0 ? n_inter : 1) * stride, (Real)0);
    std::vector<Real> field((size_t)(n_field > 0 ? n_field : 1) * stride, (Real)0);
    for (size_t c = 0; c < n; c++) {
        for (int k = 0; k < n_state; k++) state[(size_t)k * stride + c] = (Real)state_aos[c * n_state + k];
        for (int k = 0; k < n_field; k++) field[(size_t)k * stride + c] = (Real)field_aos[c * n_field + k];
    }
    std::vector<Real> gxf, gyf;
    if (gx_field) {
        const size_t ngx = (size_t)ny * (nx > 1 ? nx - 1 : 0);
        gxf.assign(ngx + 1, (Real)0);
        for (size_t k = 0; k < ngx; k++) gxf[k] = (Real)gx_field[k];
    }
    if (gy_field) {
        // one extra leading row (the row shared with a slab above: none here)
        gyf.assign((size_t)(ny + 1) * nx, (Real)0);
        for (size_t k = 0; k < (size_t)(ny - 1) * nx; k++) gyf[nx + k] = (Real)gy_field[k];
    }
    std::vector<unsigned long long> row;
    std::vector<unsigned int> col;
    std::vector<Real> cg;
    if (diffusion_mode == 3) {
        row.assign(n + 1, 0);
        for (unsigned long long e = 0; e < n_conn; e++) {
            row[conn_i[e] + 1]++;
            row[conn_j[e] + 1]++;
        }
        for (size_t i = 0; i < n; i++) row[i + 1] += row[i];
        col.resize(2 * n_conn + 1);
        cg.resize(2 * n_conn + 1);
        std::vector<unsigned long long> fill(row.begin(), row.end() - 1);
        for (unsigned long long e = 0; e < n_conn; e++) {
            const unsigned long long i = conn_i[e], j = conn_j[e];
            col[fill[i]] = (unsigned int)j;
            cg[fill[i]++] = (Real)conn_g[e];
            col[fill[j]] = (unsigned int)i;
            cg[fill[j]++] = (Real)conn_g[e];
        }
    }

    MkbGridArgs g;
    memset(&g, 0, sizeof(g));
    g.state = state.data();
    g.idiff = idiff.data();
    g.inter = inter.data();
    g.field = field.data();
    g.gx_field = gx_field ? gxf.data() : nullptr;
    g.gy_field = gy_field ? gyf.data() + nx : nullptr;
    g.paced_mask = paced_mask;
    if (diffusion_mode == 3) {
        g.csr_row = row.data();
        g.csr_col = col.data();
        g.csr_g = cg.data();
    }
    g.nx = nx;
    g.ny = ny;
    g.stride = stride;
    // stand-in descriptor of the state planes (staged kernels)
    g.tmap_state[0] = (unsigned long long)(uintptr_t)state.data();
    g.tmap_state[1] = nx; g.tmap_state[2] = ny; g.tmap_state[3] = sizeof(Real);
    g.tmap_state[4] = stride; g.tmap_state[5] = (unsigned long long)n_state;
    g.tmap_state[6] = (unsigned long long)block_x; g.tmap_state[7] = (unsigned long long)block_y;
    g.iy_offset = 0;
    g.ny_global = ny;
    g.gx = gx;
    g.gy = gy;
    g.pace_x0 = px0; g.pace_x1 = px1; g.pace_y0 = py0; g.pace_y1 = py1;

    const unsigned long long cells_x = (unsigned long long)block_x * cpt, cells_y = (unsigned long long)block_y * rpt;
    unsigned int gbx = (unsigned int)((nx + cells_x - 1) / cells_x);
    const unsigned long long by_blocks = (ny + cells_y - 1) / cells_y;
    unsigned int gby = (unsigned int)(by_blocks < 32768 ? by_blocks : 32768);
    unsigned int gbz = (unsigned int)((by_blocks + gby - 1) / gby);
    if (g_stream_blocks > 0) {
        // a fixed number of blocks walks all tiles; stand-in tensor maps of the two V planes
        gbx = (unsigned int)g_stream_blocks; gby = gbz = 1;
        const unsigned long long planes[2] = {(unsigned long long)(uintptr_t)(state.data() + (size_t)i_vm * stride),
                                              (unsigned long long)(uintptr_t)v_alt.data()};
        for (int k = 0; k < 2; k++) {
            g.tmap[k][0] = planes[k];
            g.tmap[k][1] = nx;
            g.tmap[k][2] = ny;
            g.tmap[k][3] = sizeof(Real);
        }
    }
    gridDim.x = gbx; gridDim.y = gby; gridDim.z = gbz;
    // kernels whose steps overlap on the GPU: per-block step counters (here every
    // wait finds its flag set, which the shim checks)
    std::vector<unsigned int> tile_done((size_t)gbx * gby * gbz, 0u);
    g.tile_done = tile_done.data();
    blockDim.x = block_x; blockDim.y = block_y; blockDim.z = 1;

    std::vector<Fiber> fibers((size_t)block_x * block_y);
    std::vector<char> stacks((size_t)block_x * block_y * kStack);
    Real* v_main = (i_vm >= 0) ? state.data() + (size_t)i_vm * stride : nullptr;
    int parity = 0;
    size_t row_out = 0;
    std::vector<MkbStepParams> run;
    for (int s = 0; s < n_steps;) {
        // one launch: a single step, or (persistent) a run of unlogged steps
        int len = 1;
        if (g_persistent && !st_log[s]) {
            while (s + len < n_steps && !st_log[s + len]) len++;
        }
        run.resize(len);
        for (int j = 0; j < len; j++) {
            MkbStepParams& sp = run[j];
            sp.time = st_time[s + j];
            sp.dt = st_dt[s + j];
            sp.pace = st_pace[s + j];
            sp.flags = st_log[s + j] ? MKB_FLAG_STORE_AUX : 0u;
            sp.step = (unsigned int)(s + j + 1);
        }
        if (g_persistent) run[0].flags |= (unsigned int)len << 8;
        const Real* v_in = v_main ? (parity ? v_alt.data() : v_main) : nullptr;
        Real* v_out = v_main ? (parity ? v_main : v_alt.data()) : nullptr;
        if (st_log[s] && log_v) {
            // states are logged before the update (openclsim.c:1079-1084)
            const Real* vsrc = v_in ? v_in : state.data();
            for (size_t c = 0; c < n; c++) log_v[row_out * n + c] = (double)vsrc[c];
        }
        g_launch.g = g;
        g_launch.sp = run.data();
        g_launch.v_in = v_in;
        g_launch.v_out = v_out;
#ifdef MKB_KERNEL_FN2
        for (int pass = 0; pass < 2; pass++) {
            g_second = pass;
#else
        {
#endif
            for (unsigned int bz = 0; bz < gbz; bz++) {
                for (unsigned int by = 0; by < gby; by++) {
                    for (unsigned int bx = 0; bx < gbx; bx++) {
                        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
                        run_block(fibers, stacks, (unsigned int)block_x, (unsigned int)block_y);
                    }
                }
            }
        }
        g_second = 0;
        if (st_log[s]) {
            if (log_idiff) for (size_t c = 0; c < n; c++) log_idiff[row_out * n + c] = (double)idiff[c];
            if (log_inter) {
                for (int k = 0; k < n_inter; k++) {
                    for (size_t c = 0; c < n; c++) {
                        log_inter[(row_out * n_inter + k) * n + c] = (double)inter[(size_t)k * stride + c];
                    }
                }
            }
            row_out++;
        }
        if (v_main) parity ^= 1;
        s += len;
    }
    if (v_main && parity) memcpy(v_main, v_alt.data(), n * sizeof(Real));
    for (size_t c = 0; c < n; c++) {
        for (int k = 0; k < n_state; k++) state_aos[c * n_state + k] = (double)state[(size_t)k * stride + c];
    }
    return 0;
}


/*
 * Row slabs: the grid cut into `n_slabs` row ranges, every slab with its own
 * planes and exchange block ([halo_lo 3 x nx][halo_hi 3 x nx][flags lo][flags
 * hi][error], as mkb_runtime.cu lays it out), neighbours connected by plain
 * pointers. Slabs take each step one after the other: a slab's wait for its
 * neighbour's row of step n is satisfied by the neighbour's step n - 1, which
 * has already run — the same protocol as on the GPUs, minus the concurrency.
 */
namespace {

struct Slab {
    size_t iy0, ny, n, stride;
    std::vector<Real> state, v_alt, idiff, inter, field, gxf, gyf;
    std::vector<unsigned char> mask;
    std::vector<char> xchg;
    std::vector<unsigned int> tile_done;
    MkbGridArgs g;
};

}   // namespace

extern "C" int shim_run_slabs(
    int nx, int ny, int n_slabs, const int* row0, const int* row1,
    int n_state, int i_vm, int n_inter, int n_field, int diffusion_mode,
    double gx, double gy, const double* gx_field, const double* gy_field,
    long long px0, long long px1, long long py0, long long py1, const unsigned char* paced_mask,
    int n_steps, const double* st_time, const double* st_dt, const double* st_pace,
    const unsigned char* st_log,
    double* state_aos, const double* field_aos, double* log_v, double* log_idiff,
    int block_x, int block_y, unsigned int* halo_error_out)
{
    const size_t ntot = (size_t)nx * ny;
    const size_t nbx = ((size_t)nx + block_x - 1) / block_x;
    const size_t halo = 3 * (size_t)nx * sizeof(Real);
    const size_t flags = nbx * sizeof(unsigned int);
    std::vector<Slab> slabs(n_slabs);
    for (int r = 0; r < n_slabs; r++) {
        Slab& s = slabs[r];
        s.iy0 = row0[r];
        s.ny = row1[r] - row0[r];
        s.n = s.ny * nx;
        s.stride = (s.n + 31) / 32 * 32;
        s.state.assign((size_t)n_state * s.stride, (Real)0);
        s.v_alt.assign(s.stride, (Real)0);
        s.idiff.assign(s.stride, (Real)0);
        s.inter.assign((size_t)(n_inter > 0 ? n_inter : 1) * s.stride, (Real)0);
        s.field.assign((size_t)(n_field > 0 ? n_field : 1) * s.stride, (Real)0);
        const size_t c0 = s.iy0 * nx;
        for (size_t c = 0; c < s.n; c++) {
            for (int k = 0; k < n_state; k++) s.state[(size_t)k * s.stride + c] = (Real)state_aos[(c0 + c) * n_state + k];
            for (int k = 0; k < n_field; k++) s.field[(size_t)k * s.stride + c] = (Real)field_aos[(c0 + c) * n_field + k];
        }
        if (gx_field) {
            const size_t ngx = s.ny * (nx > 1 ? nx - 1 : 0);
            s.gxf.assign(ngx + 1, (Real)0);
            for (size_t k = 0; k < ngx; k++) s.gxf[k] = (Real)gx_field[s.iy0 * (nx - 1) + k];
        }
        if (gy_field) {
            // local row q <- global gy row iy0 - 1 + q; rows that do not exist stay zero
            s.gyf.assign((s.ny + 1) * nx, (Real)0);
            for (size_t q = 0; q <= s.ny; q++) {
                const long long j = (long long)s.iy0 - 1 + (long long)q;
                if (j < 0 || j > (long long)ny - 2) continue;
                for (int x = 0; x < nx; x++) s.gyf[q * nx + x] = (Real)gy_field[(size_t)j * nx + x];
            }
        }
        if (paced_mask) s.mask.assign(paced_mask + c0, paced_mask + c0 + s.n);
        s.xchg.assign(2 * halo + 2 * flags + 64, 0);
    }
    for (int r = 0; r < n_slabs; r++) {
        Slab& s = slabs[r];
        MkbGridArgs& g = s.g;
        memset(&g, 0, sizeof(g));
        g.state = s.state.data();
        g.idiff = s.idiff.data();
        g.inter = s.inter.data();
        g.field = s.field.data();
        g.gx_field = gx_field ? s.gxf.data() : nullptr;
        g.gy_field = gy_field ? s.gyf.data() + nx : nullptr;
        g.paced_mask = paced_mask ? s.mask.data() : nullptr;
        g.nx = nx;
        g.ny = s.ny;
        g.stride = s.stride;
        g.tmap_state[0] = (unsigned long long)(uintptr_t)s.state.data();
        g.tmap_state[1] = nx; g.tmap_state[2] = s.ny; g.tmap_state[3] = sizeof(Real);
        g.tmap_state[4] = s.stride; g.tmap_state[5] = (unsigned long long)n_state;
        g.tmap_state[6] = (unsigned long long)block_x; g.tmap_state[7] = (unsigned long long)block_y;
        g.iy_offset = s.iy0;
        g.ny_global = ny;
        g.gx = gx;
        g.gy = gy;
        g.pace_x0 = px0; g.pace_x1 = px1; g.pace_y0 = py0; g.pace_y1 = py1;
        char* b = s.xchg.data();
        const bool has_lo = r > 0, has_hi = r + 1 < n_slabs;
        g.halo_lo = has_lo ? b : nullptr;
        g.halo_hi = has_hi ? b + halo : nullptr;
        g.flag_lo = (const unsigned int*)(b + 2 * halo);
        g.flag_hi = (const unsigned int*)(b + 2 * halo + flags);
        g.halo_error = (unsigned int*)(b + 2 * halo + 2 * flags);
        if (has_lo) {
            char* pb = slabs[r - 1].xchg.data();
            g.peer_lo_halo_hi = pb + halo;
            g.peer_lo_flag_hi = (unsigned int*)(pb + 2 * halo + flags);
        }
        if (has_hi) {
            char* pb = slabs[r + 1].xchg.data();
            g.peer_hi_halo_lo = pb;
            g.peer_hi_flag_lo = (unsigned int*)(pb + 2 * halo);
        }
    }
    // seed: V(t0) of the boundary rows into the neighbours' slot of step 1
    for (int r = 0; r < n_slabs; r++) {
        Slab& s = slabs[r];
        const Real* v = s.state.data() + (size_t)i_vm * s.stride;
        if (s.g.peer_lo_halo_hi) {
            memcpy((Real*)s.g.peer_lo_halo_hi + 1 * (size_t)nx, v, nx * sizeof(Real));
            for (size_t k = 0; k < nbx; k++) s.g.peer_lo_flag_hi[k] = 1u;
        }
        if (s.g.peer_hi_halo_lo) {
            memcpy((Real*)s.g.peer_hi_halo_lo + 1 * (size_t)nx, v + (s.ny - 1) * nx, nx * sizeof(Real));
            for (size_t k = 0; k < nbx; k++) s.g.peer_hi_flag_lo[k] = 1u;
        }
    }

    if (getenv("SHIM_DEBUG")) {
        for (int r = 0; r < n_slabs; r++) {
            Slab& s = slabs[r];
            fprintf(stderr, "slab %d iy0 %zu ny %zu halo_lo %p halo_hi %p peer_lo_halo_hi %p peer_hi_halo_lo %p\n", r, s.iy0, s.ny,
                    s.g.halo_lo, s.g.halo_hi, s.g.peer_lo_halo_hi, s.g.peer_hi_halo_lo);
            if (s.g.halo_hi) for (int x = 0; x < nx; x++) fprintf(stderr, " %g", (double)((const Real*)s.g.halo_hi)[nx + x]);
            fprintf(stderr, "\n");
        }
    }
    blockDim.x = block_x; blockDim.y = block_y; blockDim.z = 1;
    std::vector<Fiber> fibers((size_t)block_x * block_y);
    std::vector<char> stacks((size_t)block_x * block_y * kStack);
    int parity = 0;
    size_t row_out = 0;
    for (int st = 0; st < n_steps; st++) {
        MkbStepParams sp;
        sp.time = st_time[st];
        sp.dt = st_dt[st];
        sp.pace = st_pace[st];
        sp.flags = st_log[st] ? MKB_FLAG_STORE_AUX : 0u;
        sp.step = (unsigned int)(st + 1);
        for (int r = 0; r < n_slabs; r++) {
            Slab& s = slabs[r];
            Real* v_main = s.state.data() + (size_t)i_vm * s.stride;
            const Real* v_in = parity ? s.v_alt.data() : v_main;
            Real* v_out = parity ? v_main : s.v_alt.data();
            if (st_log[st] && log_v) {
                for (size_t c = 0; c < s.n; c++) log_v[row_out * ntot + s.iy0 * nx + c] = (double)v_in[c];
            }
            const unsigned long long by_blocks = (s.ny + block_y - 1) / block_y;
            gridDim.x = (unsigned int)nbx;
            gridDim.y = (unsigned int)(by_blocks < 32768 ? by_blocks : 32768);
            gridDim.z = (unsigned int)((by_blocks + gridDim.y - 1) / gridDim.y);
            if (s.tile_done.empty()) s.tile_done.assign((size_t)gridDim.x * gridDim.y * gridDim.z, 0u);
            s.g.tile_done = s.tile_done.data();
            g_launch.g = s.g;
            g_launch.sp = &sp;
            g_launch.v_in = v_in;
            g_launch.v_out = v_out;
#ifdef MKB_KERNEL_FN2
            for (int pass = 0; pass < 2; pass++) {
                g_second = pass;
#else
            {
#endif
                for (unsigned int bz = 0; bz < gridDim.z; bz++) {
                    for (unsigned int by = 0; by < gridDim.y; by++) {
                        for (unsigned int bx = 0; bx < gridDim.x; bx++) {
                            blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
                            run_block(fibers, stacks, (unsigned int)block_x, (unsigned int)block_y);
                        }
                    }
                }
            }
            g_second = 0;
            if (st_log[st] && log_idiff) {
                for (size_t c = 0; c < s.n; c++) log_idiff[row_out * ntot + s.iy0 * nx + c] = (double)s.idiff[c];
            }
        }
        if (st_log[st]) row_out++;
        parity ^= 1;
    }
    unsigned int err = 0;
    for (int r = 0; r < n_slabs; r++) {
        Slab& s = slabs[r];
        if (parity) memcpy(s.state.data() + (size_t)i_vm * s.stride, s.v_alt.data(), s.n * sizeof(Real));
        const size_t c0 = s.iy0 * nx;
        for (size_t c = 0; c < s.n; c++) {
            for (int k = 0; k < n_state; k++) state_aos[(c0 + c) * n_state + k] = (double)s.state[(size_t)k * s.stride + c];
        }
        err |= *s.g.halo_error;
    }
    if (halo_error_out) *halo_error_out = err;
    return 0;
}


/*
 * Partitioned connection graphs: one handle per partition, stepped from Python
 * in lockstep (the push of the exported V into the other partitions' ghost
 * planes and the flag updates are done by the caller between steps, as
 * k_push_ghosts does on the GPU).
 */
namespace {

struct Part {
    size_t n, n_ghost, stride;
    int n_state, i_vm, block_x;
    std::vector<Real> state, v_alt, idiff, inter, field, ghost, cg;
    std::vector<unsigned long long> row;
    std::vector<unsigned int> col, flags, import;
    std::vector<unsigned char> mask;
    unsigned int error = 0;
    int parity = 0;
    MkbGridArgs g;
};

}   // namespace

extern "C" void* shim_part_create(
    unsigned long long n, int n_state, int i_vm, int n_inter, int n_field,
    unsigned long long n_conn, const unsigned long long* ci, const unsigned long long* cj, const double* cgd,
    unsigned long long n_ghost, int n_ranks, int n_import, const unsigned int* import_ranks,
    long long px0, long long px1, const unsigned char* paced_mask,
    const double* state_aos, const double* field_aos, int block_x)
{
    Part* p = new Part();
    p->n = n;
    p->n_ghost = n_ghost;
    p->stride = (n + 31) / 32 * 32;
    p->n_state = n_state;
    p->i_vm = i_vm;
    p->block_x = block_x;
    p->state.assign((size_t)n_state * p->stride, (Real)0);
    p->v_alt.assign(p->stride, (Real)0);
    p->idiff.assign(p->stride, (Real)0);
    p->inter.assign((size_t)(n_inter > 0 ? n_inter : 1) * p->stride, (Real)0);
    p->field.assign((size_t)(n_field > 0 ? n_field : 1) * p->stride, (Real)0);
    for (size_t c = 0; c < n; c++) {
        for (int k = 0; k < n_state; k++) p->state[(size_t)k * p->stride + c] = (Real)state_aos[c * n_state + k];
        for (int k = 0; k < n_field; k++) p->field[(size_t)k * p->stride + c] = (Real)field_aos[c * n_field + k];
    }
    // CSR as mkb_runtime.cu builds it: per-cell order = edge-list order; an
    // endpoint >= n is a ghost cell and only contributes to its local end
    p->row.assign(n + 1, 0);
    for (unsigned long long e = 0; e < n_conn; e++) {
        p->row[ci[e] + 1]++;
        if (cj[e] < n) p->row[cj[e] + 1]++;
    }
    for (size_t i = 0; i < n; i++) p->row[i + 1] += p->row[i];
    p->col.assign(2 * n_conn + 1, 0);
    p->cg.assign(2 * n_conn + 1, (Real)0);
    std::vector<unsigned long long> fill(p->row.begin(), p->row.end() - 1);
    for (unsigned long long e = 0; e < n_conn; e++) {
        const unsigned long long i = ci[e], j = cj[e];
        p->col[fill[i]] = (unsigned int)j;
        p->cg[fill[i]++] = (Real)cgd[e];
        if (j < n) {
            p->col[fill[j]] = (unsigned int)i;
            p->cg[fill[j]++] = (Real)cgd[e];
        }
    }
    p->ghost.assign(3 * (n_ghost ? n_ghost : 1), (Real)0);
    p->flags.assign(n_ranks > 0 ? n_ranks : 1, 0u);
    p->import.assign(import_ranks, import_ranks + n_import);
    if (paced_mask) p->mask.assign(paced_mask, paced_mask + n);
    MkbGridArgs& g = p->g;
    memset(&g, 0, sizeof(g));
    g.state = p->state.data();
    g.idiff = p->idiff.data();
    g.inter = p->inter.data();
    g.field = p->field.data();
    g.paced_mask = paced_mask ? p->mask.data() : nullptr;
    g.csr_row = p->row.data();
    g.csr_col = p->col.data();
    g.csr_g = p->cg.data();
    g.ghost = n_ghost ? p->ghost.data() : nullptr;
    g.n_ghost = n_ghost;
    g.ghost_flags = p->flags.data();
    g.ghost_import = p->import.data();
    g.n_ghost_import = (unsigned long long)n_import;
    g.halo_error = &p->error;
    g.nx = n;
    g.ny = 1;
    g.stride = p->stride;
    g.iy_offset = 0;
    g.ny_global = 1;
    g.pace_x0 = px0; g.pace_x1 = px1; g.pace_y0 = 0; g.pace_y1 = 1;
    return p;
}

extern "C" void* shim_part_ghost(void* h) { return ((Part*)h)->ghost.data(); }
extern "C" void* shim_part_flags(void* h) { return ((Part*)h)->flags.data(); }
extern "C" unsigned int shim_part_error(void* h) { return ((Part*)h)->error; }

// V(t) of the partition's cells: what the next step will read
extern "C" void shim_part_v(void* h, double* out) {
    Part* p = (Part*)h;
    const Real* v = p->parity ? p->v_alt.data() : p->state.data() + (size_t)p->i_vm * p->stride;
    for (size_t c = 0; c < p->n; c++) out[c] = (double)v[c];
}
extern "C" void shim_part_idiff(void* h, double* out) {
    Part* p = (Part*)h;
    for (size_t c = 0; c < p->n; c++) out[c] = (double)p->idiff[c];
}

extern "C" void shim_part_step(void* h, double time, double dt, double pace, int logging, unsigned int step) {
    Part* p = (Part*)h;
    MkbStepParams sp;
    sp.time = time;
    sp.dt = dt;
    sp.pace = pace;
    sp.flags = logging ? MKB_FLAG_STORE_AUX : 0u;
    sp.step = step;
    Real* v_main = p->state.data() + (size_t)p->i_vm * p->stride;
    g_launch.g = p->g;
    g_launch.sp = &sp;
    g_launch.v_in = p->parity ? p->v_alt.data() : v_main;
    g_launch.v_out = p->parity ? v_main : p->v_alt.data();
    blockDim.x = p->block_x; blockDim.y = 1; blockDim.z = 1;
    gridDim.x = (unsigned int)((p->n + p->block_x - 1) / p->block_x);
    gridDim.y = gridDim.z = 1;
    static std::vector<Fiber> fibers;
    static std::vector<char> stacks;
    if (fibers.size() < (size_t)p->block_x) {
        fibers.resize(p->block_x);
        stacks.resize((size_t)p->block_x * kStack);
    }
    for (unsigned int bx = 0; bx < gridDim.x; bx++) {
        blockIdx.x = bx; blockIdx.y = 0; blockIdx.z = 0;
        run_block(fibers, stacks, (unsigned int)p->block_x, 1);
    }
    p->parity ^= 1;
}

extern "C" void shim_part_finish(void* h, double* state_aos) {
    Part* p = (Part*)h;
    if (p->parity) memcpy(p->state.data() + (size_t)p->i_vm * p->stride, p->v_alt.data(), p->n * sizeof(Real));
    for (size_t c = 0; c < p->n; c++) {
        for (int k = 0; k < p->n_state; k++) state_aos[c * p->n_state + k] = (double)p->state[(size_t)k * p->stride + c];
    }
    delete p;
}


/*
 * One homogeneous grid as a handle, stepped from Python: two of them (two
 * kernels, two libraries) coupled through MkbGridArgs::junction_* make the
 * fibre-tissue pair.
 */
namespace {

struct Grid {
    size_t nx, ny, n, stride;
    int n_state, i_vm, n_inter, block_x, block_y;
    std::vector<Real> state, v_alt, idiff, inter, field;
    std::vector<unsigned char> mask;
    int parity = 0;
    MkbGridArgs g;
};

}   // namespace

extern "C" void* shim_grid_create(
    int nx, int ny, int n_state, int i_vm, int n_inter, int n_field, double gx, double gy,
    long long px0, long long px1, long long py0, long long py1, const unsigned char* paced_mask,
    const double* state_aos, const double* field_aos, int block_x, int block_y)
{
    Grid* p = new Grid();
    p->nx = nx; p->ny = ny; p->n = (size_t)nx * ny;
    p->stride = (p->n + 31) / 32 * 32;
    p->n_state = n_state; p->i_vm = i_vm; p->n_inter = n_inter;
    p->block_x = block_x; p->block_y = block_y;
    p->state.assign((size_t)n_state * p->stride, (Real)0);
    p->v_alt.assign(p->stride, (Real)0);
    p->idiff.assign(p->stride, (Real)0);
    p->inter.assign((size_t)(n_inter > 0 ? n_inter : 1) * p->stride, (Real)0);
    p->field.assign((size_t)(n_field > 0 ? n_field : 1) * p->stride, (Real)0);
    for (size_t c = 0; c < p->n; c++) {
        for (int k = 0; k < n_state; k++) p->state[(size_t)k * p->stride + c] = (Real)state_aos[c * n_state + k];
        for (int k = 0; k < n_field; k++) p->field[(size_t)k * p->stride + c] = (Real)field_aos[c * n_field + k];
    }
    if (paced_mask) p->mask.assign(paced_mask, paced_mask + p->n);
    MkbGridArgs& g = p->g;
    memset(&g, 0, sizeof(g));
    g.state = p->state.data();
    g.idiff = p->idiff.data();
    g.inter = p->inter.data();
    g.field = p->field.data();
    g.paced_mask = paced_mask ? p->mask.data() : nullptr;
    g.nx = nx; g.ny = ny; g.stride = p->stride;
    g.iy_offset = 0; g.ny_global = ny;
    g.gx = gx; g.gy = gy;
    g.pace_x0 = px0; g.pace_x1 = px1; g.pace_y0 = py0; g.pace_y1 = py1;
    return p;
}

extern "C" void* shim_grid_vplane(void* h, int which) {
    Grid* p = (Grid*)h;
    return which ? p->v_alt.data() : p->state.data() + (size_t)p->i_vm * p->stride;
}

extern "C" void shim_grid_set_junction(void* h, const void* v0, const void* v1, double jg,
                                       unsigned long long jx, unsigned long long jy0, unsigned long long jn,
                                       unsigned long long joff, unsigned long long jstride) {
    Grid* p = (Grid*)h;
    p->g.junction_v0 = v0;
    p->g.junction_v1 = v1;
    p->g.jg = jg;
    p->g.jx = jx; p->g.jy0 = jy0; p->g.jn = jn; p->g.joff = joff; p->g.jstride = jstride;
}

extern "C" void shim_grid_step(void* h, double time, double dt, double pace, int logging, unsigned int step) {
    Grid* p = (Grid*)h;
    MkbStepParams sp;
    sp.time = time; sp.dt = dt; sp.pace = pace;
    sp.flags = logging ? MKB_FLAG_STORE_AUX : 0u;
    sp.step = step;
    Real* v_main = p->state.data() + (size_t)p->i_vm * p->stride;
    g_launch.g = p->g;
    g_launch.sp = &sp;
    g_launch.v_in = p->parity ? p->v_alt.data() : v_main;
    g_launch.v_out = p->parity ? v_main : p->v_alt.data();
    blockDim.x = p->block_x; blockDim.y = p->block_y; blockDim.z = 1;
    gridDim.x = (unsigned int)((p->nx + p->block_x - 1) / p->block_x);
    gridDim.y = (unsigned int)((p->ny + p->block_y - 1) / p->block_y);
    gridDim.z = 1;
    static std::vector<Fiber> fibers;
    static std::vector<char> stacks;
    const size_t nthreads = (size_t)p->block_x * p->block_y;
    if (fibers.size() < nthreads) {
        fibers.resize(nthreads);
        stacks.resize(nthreads * kStack);
    }
    for (unsigned int by = 0; by < gridDim.y; by++) {
        for (unsigned int bx = 0; bx < gridDim.x; bx++) {
            blockIdx.x = bx; blockIdx.y = by; blockIdx.z = 0;
            run_block(fibers, stacks, (unsigned int)p->block_x, (unsigned int)p->block_y);
        }
    }
    p->parity ^= 1;
}

// V(t) as the next step will read it / idiff and logged intermediaries of the last logged step
extern "C" void shim_grid_v(void* h, double* out) {
    Grid* p = (Grid*)h;
    const Real* v = p->parity ? p->v_alt.data() : p->state.data() + (size_t)p->i_vm * p->stride;
    for (size_t c = 0; c < p->n; c++) out[c] = (double)v[c];
}
extern "C" void shim_grid_idiff(void* h, double* out) {
    Grid* p = (Grid*)h;
    for (size_t c = 0; c < p->n; c++) out[c] = (double)p->idiff[c];
}
extern "C" void shim_grid_inter(void* h, int k, double* out) {
    Grid* p = (Grid*)h;
    for (size_t c = 0; c < p->n; c++) out[c] = (double)p->inter[(size_t)k * p->stride + c];
}
extern "C" void shim_grid_finish(void* h, double* state_aos) {
    Grid* p = (Grid*)h;
    if (p->parity) memcpy(p->state.data() + (size_t)p->i_vm * p->stride, p->v_alt.data(), p->n * sizeof(Real));
    for (size_t c = 0; c < p->n; c++) {
        for (int k = 0; k < p->n_state; k++) state_aos[c * p->n_state + k] = (double)p->state[(size_t)k * p->stride + c];
    }
    delete p;
}
