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

struct Launch {
    MkbGridArgs g;
    const MkbStepParams* sp;
    const Real* v_in;
    Real* v_out;
} g_launch;

void fiber_main() {
    mkb_cell_step(g_launch.g, g_launch.sp, g_launch.v_in, g_launch.v_out);
    g_cur->done = true;
    swapcontext(&g_cur->ctx, &g_sched);
}

}   // namespace

void shim_barrier() {
    g_cur->waiting = true;
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
        for (unsigned int t = 0; t < nthreads; t++) {
            Fiber& f = fibers[t];
            if (f.done || f.waiting) continue;
            threadIdx.x = f.tx;
            threadIdx.y = f.ty;
            threadIdx.z = 0;
            g_cur = &f;
            swapcontext(&g_sched, &f.ctx);
        }
        for (unsigned int t = 0; t < nthreads; t++) alive = alive || !fibers[t].done;
        if (!alive) break;
        // every thread that has not exited is waiting: release the barrier
        g_or_res = g_or_acc;
        g_or_acc = 0;
        for (unsigned int t = 0; t < nthreads; t++) fibers[t].waiting = false;
    }
}

}   // namespace

extern "C" int shim_real_size(void) { return (int)sizeof(Real); }

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
    g.iy_offset = 0;
    g.ny_global = ny;
    g.gx = gx;
    g.gy = gy;
    g.pace_x0 = px0; g.pace_x1 = px1; g.pace_y0 = py0; g.pace_y1 = py1;

    const unsigned long long cells_x = (unsigned long long)block_x * cpt, cells_y = (unsigned long long)block_y * rpt;
    const unsigned int gbx = (unsigned int)((nx + cells_x - 1) / cells_x);
    const unsigned long long by_blocks = (ny + cells_y - 1) / cells_y;
    const unsigned int gby = (unsigned int)(by_blocks < 32768 ? by_blocks : 32768);
    const unsigned int gbz = (unsigned int)((by_blocks + gby - 1) / gby);
    gridDim.x = gbx; gridDim.y = gby; gridDim.z = gbz;
    blockDim.x = block_x; blockDim.y = block_y; blockDim.z = 1;

    std::vector<Fiber> fibers((size_t)block_x * block_y);
    std::vector<char> stacks((size_t)block_x * block_y * kStack);
    Real* v_main = (i_vm >= 0) ? state.data() + (size_t)i_vm * stride : nullptr;
    int parity = 0;
    size_t row_out = 0;
    for (int s = 0; s < n_steps; s++) {
        MkbStepParams sp;
        sp.time = st_time[s];
        sp.dt = st_dt[s];
        sp.pace = st_pace[s];
        sp.flags = st_log[s] ? MKB_FLAG_STORE_AUX : 0u;
        sp.step = (unsigned int)(s + 1);
        const Real* v_in = v_main ? (parity ? v_alt.data() : v_main) : nullptr;
        Real* v_out = v_main ? (parity ? v_main : v_alt.data()) : nullptr;
        if (st_log[s] && log_v) {
            // states are logged before the update (openclsim.c:1079-1084)
            const Real* vsrc = v_in ? v_in : state.data();
            for (size_t c = 0; c < n; c++) log_v[row_out * n + c] = (double)vsrc[c];
        }
        g_launch.g = g;
        g_launch.sp = &sp;
        g_launch.v_in = v_in;
        g_launch.v_out = v_out;
        for (unsigned int bz = 0; bz < gbz; bz++) {
            for (unsigned int by = 0; by < gby; by++) {
                for (unsigned int bx = 0; bx < gbx; bx++) {
                    blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
                    run_block(fibers, stacks, (unsigned int)block_x, (unsigned int)block_y);
                }
            }
        }
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
    }
    if (v_main && parity) memcpy(v_main, v_alt.data(), n * sizeof(Real));
    for (size_t c = 0; c < n; c++) {
        for (int k = 0; k < n_state; k++) state_aos[c * n_state + k] = (double)state[(size_t)k * stride + c];
    }
    return 0;
}
