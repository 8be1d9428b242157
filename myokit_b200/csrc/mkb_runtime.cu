// mkb_runtime.cu — host runtime + fixed helper kernels of libmyokit_b200.so.
//
// Replaces the reference's generated host driver for SimulationOpenCL
// (myokit/_sim/openclsim.c): device setup and buffers (sim_init, :309-1031),
// the time loop (sim_step, :1036-1211) and teardown (sim_clean, :216-304),
// re-designed for one B200:
//   * the model kernel arrives as an sm_100a cubin (NVRTC, mkb_jit_compile)
//     and is ONE fused launch per time step (stencil + cell update) instead of
//     the reference's diffusion kernel + cell kernel + 3 clSetKernelArg;
//   * per-step scalars (time, dt, pace, flags) are written ahead into a device
//     ring, so launches carry no per-step host state and batches of steps can
//     be replayed as CUDA graphs;
//   * state is structure-of-arrays with a double-buffered V plane;
//   * logging is a strided device gather into a device row ring that a side
//     stream drains into pinned host memory — never a full-state copy per log
//     point (the reference does one, openclsim.c:1083);
//   * the host never synchronises inside a batch; NaN detection
//     (openclsim.c:1087) rides along as a hidden log column.
//
// No CPU fallback exists: every compute entry point fails with MKB_ERR_CUDA
// when there is no device.
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvrtc.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/myokit_b200.h"
#include "mkb_device_abi.h"
#include "mkb_pacing.hpp"
#include "mkb_schedule.hpp"

typedef unsigned long long u64;

// ---------------------------------------------------------------------------
// Errors
// ---------------------------------------------------------------------------
static thread_local std::string g_error;

static int fail(int code, const char* fmt, ...) {
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                        \
    do {                                                                      \
        cudaError_t e_ = (expr);                                              \
        if (e_ != cudaSuccess) {                                              \
            return fail(MKB_ERR_CUDA, "CUDA error %d (%s) at %s:%d: %s",      \
                        (int)e_, cudaGetErrorName(e_), __FILE__, __LINE__,    \
                        cudaGetErrorString(e_));                              \
        }                                                                     \
    } while (0)

static const char* const kDeviceAbiText =
#include "mkb_device_abi_text.inc"
    ;

// ---------------------------------------------------------------------------
// Fixed helper kernels (data movement only; the model kernel is generated)
// ---------------------------------------------------------------------------

// Reference layout aos[(c0 + c) * nvar + k]  ->  planes[k * stride + c0 + c].
// One thread per (cell, k), k fastest: coalesced reads of the staging chunk.
template <typename TH, typename TR>
__global__ void k_aos_to_soa(const TH* __restrict__ aos, TR* __restrict__ planes,
                             u64 c0, u64 ncells, int nvar, u64 stride) {
    u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    u64 total = ncells * (u64)nvar;
    for (; i < total; i += (u64)gridDim.x * blockDim.x) {
        u64 c = i / nvar;
        int k = (int)(i - c * nvar);
        planes[(u64)k * stride + c0 + c] = (TR)aos[i];
    }
}

// planes -> aos chunk; plane `i_vm` is read from `v_cur` (ping-pong buffer).
template <typename TR, typename TH>
__global__ void k_soa_to_aos(const TR* __restrict__ planes, const TR* __restrict__ v_cur,
                             int i_vm, TH* __restrict__ aos, u64 c0, u64 ncells,
                             int nvar, u64 stride) {
    u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    u64 total = ncells * (u64)nvar;
    for (; i < total; i += (u64)gridDim.x * blockDim.x) {
        u64 c = i / nvar;
        int k = (int)(i - c * nvar);
        TR val = (k == i_vm) ? v_cur[c0 + c] : planes[(u64)k * stride + c0 + c];
        aos[i] = (TH)val;
    }
}

// One cell's values for every cell: planes[k * stride + c] = cell[k].
template <typename TH, typename TR>
__global__ void k_fill_planes(const TH* __restrict__ cell, TR* __restrict__ planes,
                              u64 ncells, int nvar, u64 stride) {
    const int k = blockIdx.y;
    if (k >= nvar) return;
    const TR v = (TR)cell[k];
    TR* const plane = planes + (u64)k * stride;
    for (u64 c = blockIdx.x * (u64)blockDim.x + threadIdx.x; c < ncells; c += (u64)gridDim.x * blockDim.x)
        plane[c] = v;
}

template <typename TH, typename TR>
__global__ void k_convert(const TH* __restrict__ in, TR* __restrict__ out, u64 n) {
    u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    for (; i < n; i += (u64)gridDim.x * blockDim.x) out[i] = (TR)in[i];
}

// Strided log gather: row[col[j]] = src[off[j]]. Offsets are element offsets
// from `base` (all planes live in one allocation); bit 63 marks an entry of
// the ping-pong V plane, resolved against `v_cur`.
struct GatherEntry {
    u64 off;
    unsigned int col;
    unsigned int pad;
};
#define MKB_GATHER_V (1ull << 63)

template <typename TR>
__global__ void k_log_gather(const TR* __restrict__ base, const TR* __restrict__ v_cur,
                             const GatherEntry* __restrict__ tab, u64 n,
                             TR* __restrict__ row) {
    u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    for (; i < n; i += (u64)gridDim.x * blockDim.x) {
        GatherEntry e = tab[i];
        TR val = (e.off & MKB_GATHER_V) ? v_cur[e.off & ~MKB_GATHER_V] : base[e.off];
        row[e.col] = val;
    }
}

__global__ void k_fill_u32(unsigned int* p, u64 n, unsigned int value) {
    u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (i < n) p[i] = value;
}

// --- partitioned connection graphs ---------------------------------------
// One launch per step for all peers: entry e copies v[src[e]] into ghost slot
// slot[e] of peer peer[e] (a store over NVLink), in that peer's plane for
// `step`. The block that finishes last — every other block's stores are
// behind a system-scope fence by then — raises this rank's flag on every
// peer, so no second launch is needed to publish the step.
template <typename TR>
__global__ void k_push_ghosts(const TR* __restrict__ v, const u64* __restrict__ src,
                              const u64* __restrict__ slot, const unsigned int* __restrict__ peer,
                              u64 n, void* const* peer_base, const u64* __restrict__ peer_n_ghost,
                              unsigned int* const* flags, unsigned int n_peers, unsigned int step_arg,
                              const MkbStepParams* __restrict__ sp, unsigned int* done) {
    __shared__ bool last;
    // the step this push delivers for: the one after the step record `sp` (so
    // that the launch can sit in a replayed CUDA graph), or `step_arg`
    const unsigned int step = sp ? sp->step + 1u : step_arg;
    const u64 plane = step % 3u;
    u64 e = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    for (; e < n; e += (u64)gridDim.x * blockDim.x) {
        const unsigned int p = peer[e];
        TR* dst = (TR*)peer_base[p] + plane * peer_n_ghost[p];
        dst[slot[e]] = v[src[e]];
    }
    // one system-scope fence per block, after the block has met: it is
    // cumulative over the stores of every thread that reached the barrier
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        last = (atomicAdd(done, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence_system();
    for (unsigned int k = threadIdx.x; k < n_peers; k += blockDim.x) {
        *((volatile unsigned int*)flags[k]) = step;
    }
    if (threadIdx.x == 0) *done = 0;
}

// ---------------------------------------------------------------------------
// Library-level API
// ---------------------------------------------------------------------------
extern "C" int mkb_abi_version(void) { return MKB_ABI_VERSION; }

extern "C" int mkb_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(MKB_ERR_INVALID, "null argument");
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(MKB_ERR_CUDA, "cudaHostAlloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    return MKB_OK;
}

extern "C" void mkb_host_free(void* p) {
    if (p) cudaFreeHost(p);
}
extern "C" const char* mkb_last_error(void) { return g_error.c_str(); }
extern "C" void mkb_free(void* p) { free(p); }
extern "C" const char* mkb_device_abi_header(void) { return kDeviceAbiText; }

extern "C" int mkb_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        return fail(MKB_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return n;
}

extern "C" int mkb_device_info(int device, mkb_device_info_t* out) {
    if (!out) return fail(MKB_ERR_INVALID, "out is null");
    cudaDeviceProp p;
    CUDA_TRY(cudaGetDeviceProperties(&p, device));
    memset(out, 0, sizeof(*out));
    snprintf(out->name, sizeof(out->name), "%s", p.name);
    out->cc_major = p.major;
    out->cc_minor = p.minor;
    out->sm_count = p.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    out->clock_khz = khz;
    out->total_mem = p.totalGlobalMem;
    out->l2_bytes = (size_t)p.l2CacheSize;
    out->smem_per_block_optin = p.sharedMemPerBlockOptin;
    return MKB_OK;
}

extern "C" int mkb_jit_compile(const char* source, const char* options,
                               void** cubin, size_t* cubin_size, char** log) {
    if (!source || !cubin || !cubin_size) return fail(MKB_ERR_INVALID, "null argument");
    *cubin = nullptr;
    *cubin_size = 0;
    if (log) *log = nullptr;
    nvrtcProgram prog;
    const char* hdr_src[] = {kDeviceAbiText};
    const char* hdr_name[] = {"mkb_device_abi.h"};
    nvrtcResult r = nvrtcCreateProgram(&prog, source, "mkb_model.cu", 1, hdr_src, hdr_name);
    if (r != NVRTC_SUCCESS) return fail(MKB_ERR_JIT, "nvrtcCreateProgram: %s", nvrtcGetErrorString(r));
    std::vector<const char*> opts;
    opts.push_back("--gpu-architecture=sm_100a");
    opts.push_back("--std=c++17");
    opts.push_back("-lineinfo");
    opts.push_back("-default-device");
    if (options) {
        for (const char* o = options; *o; o += strlen(o) + 1) opts.push_back(o);
    }
    r = nvrtcCompileProgram(prog, (int)opts.size(), opts.data());
    size_t log_size = 0;
    nvrtcGetProgramLogSize(prog, &log_size);
    std::string logtext(log_size ? log_size : 1, '\0');
    if (log_size) nvrtcGetProgramLog(prog, &logtext[0]);
    if (log) {
        *log = (char*)malloc(logtext.size() + 1);
        memcpy(*log, logtext.c_str(), strlen(logtext.c_str()) + 1);
    }
    if (r != NVRTC_SUCCESS) {
        nvrtcDestroyProgram(&prog);
        return fail(MKB_ERR_JIT, "NVRTC compilation failed (%s):\n%s", nvrtcGetErrorString(r),
                    logtext.c_str());
    }
    size_t n = 0;
    r = nvrtcGetCUBINSize(prog, &n);
    if (r != NVRTC_SUCCESS || n == 0) {
        nvrtcDestroyProgram(&prog);
        return fail(MKB_ERR_JIT, "nvrtcGetCUBINSize: %s", nvrtcGetErrorString(r));
    }
    void* img = malloc(n);
    r = nvrtcGetCUBIN(prog, (char*)img);
    nvrtcDestroyProgram(&prog);
    if (r != NVRTC_SUCCESS) {
        free(img);
        return fail(MKB_ERR_JIT, "nvrtcGetCUBIN: %s", nvrtcGetErrorString(r));
    }
    *cubin = img;
    *cubin_size = n;
    return MKB_OK;
}

extern "C" int mkb_schedule_probe(double tmin, double tmax, double dt, double log_interval,
                                  int n_events, const double* events, uint64_t max_steps,
                                  double* times, double* dts, double* paces, unsigned char* logging,
                                  uint64_t* n_steps) {
    if (!n_steps) return fail(MKB_ERR_INVALID, "null argument");
    mkb::StepScheduler sched;
    int rc = sched.init(tmin, tmax, dt, log_interval, n_events, events);
    u64 n = 0;
    while (rc == 0 && !sched.finished() && n < max_steps) {
        mkb::StepInfo info;
        rc = sched.next(&info);
        if (rc) break;
        if (times) times[n] = info.time;
        if (dts) dts[n] = info.dt;
        if (paces) paces[n] = info.pace;
        if (logging) logging[n] = info.logging ? 1 : 0;
        n++;
    }
    *n_steps = n;
    if (rc == mkb::PACING_SIMULTANEOUS_EVENT) {
        return fail(MKB_ERR_SIMULTANEOUS,
                    "E-Pacing error: Event scheduled or re-occuring at the same time as another event.");
    }
    if (rc) return fail(MKB_ERR_PACING, "E-Pacing error %d", rc);
    return sched.finished() ? MKB_OK : 1;
}

extern "C" int mkb_pacing_probe(double t0, int n_events, const double* events, int n_times,
                                const double* times, double* levels, double* next_times) {
    mkb::EventPacing p;
    int rc = p.init(t0, n_events, events);
    for (int i = 0; rc == 0 && i < n_times; i++) {
        rc = p.advance(times[i]);
        if (rc) break;
        levels[i] = p.level();
        next_times[i] = p.next_time();
    }
    if (rc == mkb::PACING_SIMULTANEOUS_EVENT) {
        return fail(MKB_ERR_SIMULTANEOUS,
                    "E-Pacing error: Event scheduled or re-occuring at the same time as another event.");
    }
    if (rc) return fail(MKB_ERR_PACING, "E-Pacing error %d", rc);
    return MKB_OK;
}

// ---------------------------------------------------------------------------
// Simulation object
// ---------------------------------------------------------------------------
static const int kRingHalf = 1024;      // schedule entries per ring half
// At most 2 * kThrottle step kernels are in flight per simulation. The driver's
// launch queue is finite and a full queue blocks the launching thread inside
// the runtime; with slabs of several GPUs driven from one process that could
// stop the very thread whose kernels the queued ones are waiting for.
static const int kThrottle = 128;
// Steps replayed by one CUDA graph launch (even, so the V ping-pong parity is
// the same before and after): one driver call and one device-side launch
// chain instead of kGraphSteps separate launches. Matters when a step is only
// a few microseconds of work (small and medium grids).
static const int kGraphSteps = 64;

struct StepRec {
    MkbStepParams p;
    bool logging;
    double log_time, log_pace;          // already rounded to Real
};

struct mkb_sim {
    // configuration
    int device = 0, precision = 64, host_precision = 64;
    size_t rs = 8, hs = 8;              // sizeof(Real), sizeof(host element)
    int n_state = 0, i_vm = 0, n_inter = 0, n_field = 0, diff_mode = 0;
    u64 nx = 0, ny = 0, n = 0, stride = 0;
    int block_x = 32, block_y = 1, cpt = 1, rpt = 1;
    u64 steps_per_call = 1000;

    // CUDA objects
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kern = nullptr;
    // optional second kernel of a split step ("mkb_gate_step", kernelgen's
    // split_gates): launched after kern with the same arguments
    cudaKernel_t kern2 = nullptr;
    cudaStream_t stream = nullptr, side = nullptr;
    cudaEvent_t ev_ring[2] = {nullptr, nullptr};
    cudaEvent_t ev_rows = nullptr, ev_copied[2] = {nullptr, nullptr};
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    cudaEvent_t ev_throttle[2] = {nullptr, nullptr};
    u64 throttle_count = 0;             // groups of kThrottle launches issued
    u64 issued = 0;                     // step kernels launched in this run
    bool copied_pending[2] = {false, false};

    // device memory
    char* d_planes = nullptr;           // all Real planes, one allocation
    u64 plane_alt_v = 0, plane_idiff = 0, plane_inter = 0, plane_field = 0, n_planes = 0;
    char* d_gx = nullptr;
    char* d_gy = nullptr;
    unsigned char* d_mask = nullptr;
    u64* d_csr_row = nullptr;
    unsigned int* d_csr_col = nullptr;
    char* d_csr_g = nullptr;
    MkbStepParams* d_ring = nullptr;
    MkbStepParams* h_ring = nullptr;    // pinned
    u64 ring_chunk = 0;                 // chunks issued so far
    int parity = 0;                     // 0: V(t) in plane i_vm, 1: in plane alt_v
    MkbGridArgs grid{};
    dim3 launch_grid, launch_block;
    unsigned int smem_bytes = 0;        // dynamic shared memory of the step kernel (staged kernels)

    // schedule (openclsim.c globals :84-211)
    double tmin = 0, tmax = 0, default_dt = 0, log_interval = 1;
    double engine_time = 0;
    mkb::StepScheduler sched;
    bool finished = false, halted = false;

    // logging
    u64 n_log = 0, row_stride = 0;      // columns; row_stride = n_log + 1 (hidden NaN probe)
    struct FieldCopy {
        u64 plane;                      // plane index, or ~0ull for the current V plane
        u64 col;                        // first column in the row
    };
    std::vector<FieldCopy> pre_fields, post_fields;
    u64 h_log_bytes = 0;                // capacity of h_log in bytes
    std::vector<char*> h_log_retired;   // outgrown pinned matrices, freed at clean
    u64 h_log_retired_bytes = 0;
    std::vector<int> time_cols, pace_cols;
    GatherEntry* d_tab_pre = nullptr;   // states at t (before the step kernel)
    GatherEntry* d_tab_post = nullptr;  // idiff / intermediaries at t (after it)
    u64 n_pre = 0, n_post = 0;
    bool logging_states = false, store_aux = false;
    char* d_log = nullptr;              // device row ring (null: no device rows this run)
    // Allocations behind d_log / d_tab_*: kept across re-armed runs, because
    // cudaFree / cudaMalloc synchronise the device and were measured at up to
    // several hundred ms between two runs of 20-200 ms.
    char* d_log_store = nullptr;
    u64 d_log_bytes = 0;
    GatherEntry* d_tab_store[2] = {nullptr, nullptr};
    u64 d_tab_cap[2] = {0, 0};
    u64 log_cap = 0, log_half = 0;      // rows in ring / per half
    u64 rows_written = 0, rows_flushed = 0, rows_final = 0;
    char* h_log = nullptr;              // pinned host matrix
    u64 h_log_cap = 0;                  // rows
    std::vector<double> row_time, row_pace;     // for rows >= rows_final

    // CUDA graphs of kGraphSteps plain (non-logging) steps
    struct GraphSlot {
        cudaGraphExec_t exec[2] = {nullptr, nullptr};   // by V parity at entry
        MkbStepParams* h_stage = nullptr;               // pinned, kGraphSteps entries
        MkbStepParams* d_params = nullptr;              // device, kGraphSteps entries
        cudaEvent_t done = nullptr;
        bool pending = false;
    };
    GraphSlot gslot[2];
    bool use_graphs = false;
    u64 graph_seq = 0, graph_launches = 0;

    // row-slab halo exchange (multi-GPU)
    char* d_xchg = nullptr;             // [halo_lo 3*nx][halo_hi 3*nx][flag_lo nbx][flag_hi nbx][error]
    size_t xchg_bytes = 0;
    u64 nbx = 0;                        // column blocks = flags per side
    void* peer_lo_base = nullptr;       // neighbour exchange blocks as mapped here
    void* peer_hi_base = nullptr;
    bool peer_lo_ipc = false, peer_hi_ipc = false;
    bool has_lo = false, has_hi = false;
    u64 step_index = 0;                 // steps taken in this run (1-based in the kernel)
    bool halo_live = false;             // exchange block consistent with step_index (see arm_run)
    bool halo_connected = false;        // halo_connect / ghost_connect + first seed done
    bool state_replaced = false;        // mkb_sim_set_state since the last run
    void* d_stage[2] = {nullptr, nullptr};  // state up / download staging (kept: cudaMalloc and
    size_t stage_bytes = 0;                 // cudaFree per transfer cost ms and synchronise the device)
    bool overlap = false;               // MKB_KERNEL_OVERLAP: step kernels launched with PDL
    unsigned int* d_tile_done = nullptr;
    size_t n_tiles = 0;

    // partitioned connection graphs: ghost cells (multi-GPU)
    u64 n_ghost = 0;
    unsigned int n_flags = 0;           // flags in this partition's block (one per rank)
    struct GhostPeerRt {
        void* base = nullptr;           // peer's exchange block as mapped here
        bool ipc = false;
        u64 peer_n_ghost = 0;
        u64 n_export = 0;
        unsigned int* flag = nullptr;   // the flag this rank raises on the peer
    };
    std::vector<GhostPeerRt> gpeers;
    // all peers' export lists, concatenated (one push launch per step)
    u64 n_export = 0;
    u64* d_exp_src = nullptr;           // local cells to send
    u64* d_exp_slot = nullptr;          // ghost slots on the receiving peer
    unsigned int* d_exp_peer = nullptr; // which peer (index into gpeers)
    void** d_peer_base = nullptr;       // device array: the peers' exchange blocks
    u64* d_peer_n_ghost = nullptr;      // device array: ghost count of each peer
    unsigned int* d_push_done = nullptr;// block counter of the push kernel
    unsigned int** d_peer_flags = nullptr;  // device array of the peers' flag pointers
    unsigned int* d_import = nullptr;       // indices of own flags to wait for
    unsigned int n_import = 0;
    bool ghosts_connected = false;

    // kernel "mkb_cell_step_persistent": the block is the grid, consecutive
    // unlogged steps share one launch (run length in flags >> 8)
    bool persistent = false;
    std::vector<u64> runlen;

    // fibre-tissue pair (mkb_sim_junction_connect): the other grid, and the
    // events this one records after its step kernels (odd / even steps)
    mkb_sim* partner = nullptr;
    cudaEvent_t ev_step[2] = {nullptr, nullptr};

    // counters
    u64 launches = 0, steps = 0;
    double device_ms = 0;

    std::vector<StepRec> recs;
};

template <typename T>
static T* plane_ptr(mkb_sim* s, u64 plane) {
    return (T*)(s->d_planes + plane * s->stride * s->rs);
}

static void sim_destroy(mkb_sim* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->side) cudaStreamSynchronize(s->side);
    if (s->partner) {
        // the other grid must not wait for, or read from, this one any more
        if (s->partner->stream) cudaStreamSynchronize(s->partner->stream);
        s->partner->grid.junction_v0 = s->partner->grid.junction_v1 = nullptr;
        s->partner->partner = nullptr;
        s->partner = nullptr;
    }
    for (int k = 0; k < 2; k++) {
        if (s->ev_step[k]) cudaEventDestroy(s->ev_step[k]);
    }
    cudaFree(s->d_planes);
    cudaFree(s->d_gx);
    cudaFree(s->d_gy);
    cudaFree(s->d_mask);
    cudaFree(s->d_csr_row);
    cudaFree(s->d_csr_col);
    cudaFree(s->d_csr_g);
    cudaFree(s->d_ring);
    cudaFree(s->d_tab_store[0]);
    cudaFree(s->d_tab_store[1]);
    cudaFree(s->d_log_store);
    for (int i = 0; i < 2; i++) {
        mkb_sim::GraphSlot& gs = s->gslot[i];
        for (int p = 0; p < 2; p++) {
            if (gs.exec[p]) cudaGraphExecDestroy(gs.exec[p]);
        }
        if (gs.h_stage) cudaFreeHost(gs.h_stage);
        cudaFree(gs.d_params);
        if (gs.done) cudaEventDestroy(gs.done);
    }
    for (auto& gp : s->gpeers) {
        if (gp.base && gp.ipc) cudaIpcCloseMemHandle(gp.base);
    }
    cudaFree(s->d_exp_src);
    cudaFree(s->d_exp_slot);
    cudaFree(s->d_exp_peer);
    cudaFree(s->d_peer_base);
    cudaFree(s->d_peer_n_ghost);
    cudaFree(s->d_push_done);
    cudaFree(s->d_peer_flags);
    cudaFree(s->d_import);
    if (s->peer_lo_base && s->peer_lo_ipc) cudaIpcCloseMemHandle(s->peer_lo_base);
    if (s->peer_hi_base && s->peer_hi_ipc) cudaIpcCloseMemHandle(s->peer_hi_base);
    cudaFree(s->d_xchg);
    if (s->h_ring) cudaFreeHost(s->h_ring);
    if (s->h_log) cudaFreeHost(s->h_log);
    for (char* q : s->h_log_retired) cudaFreeHost(q);
    for (int i = 0; i < 2; i++) {
        if (s->ev_ring[i]) cudaEventDestroy(s->ev_ring[i]);
        if (s->ev_copied[i]) cudaEventDestroy(s->ev_copied[i]);
        if (s->ev_throttle[i]) cudaEventDestroy(s->ev_throttle[i]);
    }
    if (s->ev_rows) cudaEventDestroy(s->ev_rows);
    if (s->ev_t0) cudaEventDestroy(s->ev_t0);
    if (s->ev_t1) cudaEventDestroy(s->ev_t1);
    if (s->d_tile_done) cudaFree(s->d_tile_done);
    if (s->d_stage[0]) cudaFree(s->d_stage[0]);
    if (s->d_stage[1]) cudaFree(s->d_stage[1]);
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->side) cudaStreamDestroy(s->side);
    if (s->lib) cudaLibraryUnload(s->lib);
    delete s;
}

// NVTX range for the lifetime of a scope (shows up in Nsight Systems / ncu
// --nvtx timelines: init, arm, step calls, log flush, state transfers).
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

static int grid_for(u64 n) {
    u64 b = (n + 255) / 256;
    return (int)std::min<u64>(std::max<u64>(b, 1), 148 * 16);
}

// Host -> device on the simulation's own stream, complete on return. (Not
// cudaMemcpy: from pageable memory that call may return while the DMA of the
// staged copy is still in flight, and only the legacy default stream waits for
// it — the simulation's streams are non-blocking, so a kernel launched on
// them next could read the destination too early. Found as garbage
// conductances in a 264 x 77 single-precision run; copies of 64 KiB or less
// are inline, which is why small grids never showed it.)
// Two device staging buffers of at least `bytes` each, kept with the simulation.
static cudaError_t stage_buffers(mkb_sim* s, size_t bytes, int count) {
    if (s->stage_bytes < bytes) {
        for (int i = 0; i < 2; i++) {
            if (s->d_stage[i]) cudaFree(s->d_stage[i]);
            s->d_stage[i] = nullptr;
        }
        s->stage_bytes = 0;
    }
    for (int i = 0; i < count; i++) {
        if (!s->d_stage[i]) {
            cudaError_t e = cudaMalloc(&s->d_stage[i], bytes);
            if (e != cudaSuccess) return e;
        }
    }
    if (s->stage_bytes < bytes) s->stage_bytes = bytes;
    return cudaSuccess;
}

static cudaError_t h2d_sync(mkb_sim* s, void* dst, const void* src, size_t bytes) {
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    return e;
}

// Uploads a host array in the reference's cell-major layout into SoA planes:
// 64 MiB chunks through two device staging buffers on the launching stream,
// one synchronisation at the end. From pinned host memory (mkb_host_alloc) the
// copies are asynchronous DMA at PCIe speed and overlap the re-layout kernel
// of the previous chunk's partner; from pageable memory the driver stages
// them itself (measured: ~6 GB/s).
template <typename TH, typename TR>
static int upload_aos(mkb_sim* s, const void* host, TR* planes, int nvar) {
    if (nvar == 0) return MKB_OK;
    const u64 chunk_cells = std::max<u64>(1, (64ull << 20) / ((u64)nvar * sizeof(TH)));
    const u64 cap = std::min<u64>(chunk_cells, s->n);
    const int nbuf = (cap < s->n) ? 2 : 1;
    cudaError_t e = stage_buffers(s, cap * nvar * sizeof(TH), nbuf);
    if (e != cudaSuccess) return fail(MKB_ERR_CUDA, "state upload failed: %s", cudaGetErrorString(e));
    TH* d_stage[2] = {(TH*)s->d_stage[0], (TH*)s->d_stage[1]};
    int b = 0;
    for (u64 c0 = 0; c0 < s->n && e == cudaSuccess; c0 += cap, b = (b + 1) % nbuf) {
        const u64 nc = std::min<u64>(cap, s->n - c0);
        e = cudaMemcpyAsync(d_stage[b], (const TH*)host + c0 * nvar, nc * nvar * sizeof(TH),
                            cudaMemcpyHostToDevice, s->stream);
        if (e == cudaSuccess) {
            k_aos_to_soa<TH, TR><<<grid_for(nc * nvar), 256, 0, s->stream>>>(
                d_stage[b], planes, c0, nc, nvar, s->stride);
            s->launches++;
            e = cudaGetLastError();
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) return fail(MKB_ERR_CUDA, "state upload failed: %s", cudaGetErrorString(e));
    return MKB_OK;
}

// The same value set for every cell (mkb_sim_config::state_uniform): nvar
// host values instead of n_cells * nvar.
template <typename TH, typename TR>
static int upload_uniform(mkb_sim* s, const void* host, TR* planes, int nvar) {
    if (nvar == 0) return MKB_OK;
    TH* d_cell = nullptr;
    CUDA_TRY(cudaMalloc(&d_cell, nvar * sizeof(TH)));
    cudaError_t e = cudaMemcpyAsync(d_cell, host, nvar * sizeof(TH), cudaMemcpyHostToDevice, s->stream);
    if (e == cudaSuccess) {
        dim3 grid((unsigned)std::min<u64>((s->n + 1023) / 1024, 148 * 4), (unsigned)nvar);
        k_fill_planes<TH, TR><<<grid, 256, 0, s->stream>>>(d_cell, planes, s->n, nvar, s->stride);
        s->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(d_cell);
    if (e != cudaSuccess) return fail(MKB_ERR_CUDA, "state upload failed: %s", cudaGetErrorString(e));
    return MKB_OK;
}

template <typename TR, typename TH>
static int download_aos(mkb_sim* s, void* host) {
    const int nvar = s->n_state;
    const u64 chunk_cells = std::max<u64>(1, (64ull << 20) / ((u64)nvar * sizeof(TH)));
    const u64 cap = std::min<u64>(chunk_cells, s->n);
    const int nbuf = (cap < s->n) ? 2 : 1;
    cudaError_t e0 = stage_buffers(s, cap * nvar * sizeof(TH), nbuf);
    if (e0 != cudaSuccess) return fail(MKB_ERR_CUDA, "state download failed: %s", cudaGetErrorString(e0));
    TH* d_stage[2] = {(TH*)s->d_stage[0], (TH*)s->d_stage[1]};
    const TR* planes = plane_ptr<TR>(s, 0);
    const TR* v_cur = plane_ptr<TR>(s, s->parity ? s->plane_alt_v : (u64)std::max(s->i_vm, 0));
    cudaError_t e = cudaSuccess;
    int b = 0;
    for (u64 c0 = 0; c0 < s->n && e == cudaSuccess; c0 += cap, b = (b + 1) % nbuf) {
        const u64 nc = std::min<u64>(cap, s->n - c0);
        k_soa_to_aos<TR, TH><<<grid_for(nc * nvar), 256, 0, s->stream>>>(
            planes, v_cur, s->i_vm, d_stage[b], c0, nc, nvar, s->stride);
        s->launches++;
        e = cudaGetLastError();
        if (e == cudaSuccess) {
            e = cudaMemcpyAsync((TH*)host + c0 * nvar, d_stage[b], nc * nvar * sizeof(TH),
                                cudaMemcpyDeviceToHost, s->stream);
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) return fail(MKB_ERR_CUDA, "state download failed: %s", cudaGetErrorString(e));
    return MKB_OK;
}

// Host array (host precision) -> existing device array of Real.
template <typename TH, typename TR>
static int upload_convert(mkb_sim* s, const void* host, u64 count, TR* d) {
    if (count == 0) return MKB_OK;
    if (sizeof(TH) == sizeof(TR)) {
        CUDA_TRY(h2d_sync(s, d, host, count * sizeof(TR)));
        return MKB_OK;
    }
    TH* stage = nullptr;
    CUDA_TRY(cudaMalloc(&stage, count * sizeof(TH)));
    cudaError_t e = cudaMemcpyAsync(stage, host, count * sizeof(TH), cudaMemcpyHostToDevice, s->stream);
    if (e == cudaSuccess) {
        k_convert<TH, TR><<<grid_for(count), 256, 0, s->stream>>>(stage, d, count);
        s->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(stage);
    if (e != cudaSuccess) return fail(MKB_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e));
    return MKB_OK;
}

template <typename TR>
static int sim_init_typed(mkb_sim* s, const mkb_sim_config* c) {
    const bool host_double = (c->host_precision == MKB_DOUBLE);
    int rc;

    // --- planes ---
    s->plane_alt_v = (u64)s->n_state;
    s->plane_idiff = s->plane_alt_v + 1;
    s->plane_inter = s->plane_idiff + 1;
    s->plane_field = s->plane_inter + (u64)s->n_inter;
    s->n_planes = s->plane_field + (u64)s->n_field;
    CUDA_TRY(cudaMalloc(&s->d_planes, s->n_planes * s->stride * s->rs));
    CUDA_TRY(cudaMemsetAsync(s->d_planes, 0, s->n_planes * s->stride * s->rs, s->stream));

    // --- state + fields ---
    if (!c->state_in) return fail(MKB_ERR_INVALID, "state_in is null");
    if (c->state_uniform) {
        rc = host_double ? upload_uniform<double, TR>(s, c->state_in, plane_ptr<TR>(s, 0), s->n_state)
                         : upload_uniform<TR, TR>(s, c->state_in, plane_ptr<TR>(s, 0), s->n_state);
    } else {
        rc = host_double ? upload_aos<double, TR>(s, c->state_in, plane_ptr<TR>(s, 0), s->n_state)
                         : upload_aos<TR, TR>(s, c->state_in, plane_ptr<TR>(s, 0), s->n_state);
    }
    if (rc) return rc;
    if (s->n_field > 0) {
        if (!c->field_data) return fail(MKB_ERR_INVALID, "field_data is null");
        rc = host_double
                 ? upload_aos<double, TR>(s, c->field_data, plane_ptr<TR>(s, s->plane_field), s->n_field)
                 : upload_aos<TR, TR>(s, c->field_data, plane_ptr<TR>(s, s->plane_field), s->n_field);
        if (rc) return rc;
    }

    // --- conductance fields (config arrays are GLOBAL, reference layout) ---
    if (s->diff_mode == MKB_DIFF_FIELD) {
        const u64 iy0 = c->iy_offset;
        const u64 nyg = c->ny_global ? c->ny_global : s->ny;
        const size_t hs = host_double ? 8 : sizeof(TR);
        if (s->nx > 1) {
            if (!c->gx_field) return fail(MKB_ERR_INVALID, "gx_field is null");
            // rows [iy0, iy0 + ny) of gx[(ny_global, nx - 1)]
            const u64 ngx = (s->nx - 1) * s->ny;
            const char* src = (const char*)c->gx_field + iy0 * (s->nx - 1) * hs;
            CUDA_TRY(cudaMalloc(&s->d_gx, ngx * sizeof(TR)));
            rc = host_double ? upload_convert<double, TR>(s, src, ngx, (TR*)s->d_gx)
                             : upload_convert<TR, TR>(s, src, ngx, (TR*)s->d_gx);
            if (rc) return rc;
        }
        if (nyg > 1) {
            if (!c->gy_field) return fail(MKB_ERR_INVALID, "gy_field is null");
            // Slab copy with ny + 1 rows: local row r <- global gy row iy0 - 1 + r
            // (gy row j couples grid rows j and j + 1); missing rows stay zero.
            CUDA_TRY(cudaMalloc(&s->d_gy, (s->ny + 1) * s->nx * sizeof(TR)));
            CUDA_TRY(cudaMemsetAsync(s->d_gy, 0, (s->ny + 1) * s->nx * sizeof(TR), s->stream));
            CUDA_TRY(cudaStreamSynchronize(s->stream));
            const u64 first = iy0 > 0 ? iy0 - 1 : 0;               // first global gy row
            const u64 last = std::min<u64>(iy0 + s->ny - 1, nyg - 2);   // last global gy row
            if (last >= first && nyg >= 2) {
                const u64 rows = last - first + 1;
                const char* src = (const char*)c->gy_field + first * s->nx * hs;
                TR* dst = (TR*)s->d_gy + (first + 1 - iy0) * s->nx;
                rc = host_double ? upload_convert<double, TR>(s, src, rows * s->nx, dst)
                                 : upload_convert<TR, TR>(s, src, rows * s->nx, dst);
                if (rc) return rc;
            }
        }
    }

    // --- connections -> CSR (stable: per-cell order = edge-list order) ---
    if (s->diff_mode == MKB_DIFF_CONNECTIONS) {
        const u64 ne = c->n_connections;
        std::vector<u64> row(s->n + 1, 0);
        // An endpoint j >= n is a ghost cell (owned by another partition): the
        // edge then only contributes to the row of its local endpoint i.
        const u64 ntot = s->n + c->n_ghost;
        if (ntot > 0xffffffffull) return fail(MKB_ERR_INVALID, "too many cells for 32-bit CSR columns");
        for (u64 e = 0; e < ne; e++) {
            u64 i = c->conn_i[e], j = c->conn_j[e];
            if (i >= s->n || j >= ntot || i == j) {
                return fail(MKB_ERR_INVALID, "invalid connection %llu: (%llu, %llu)", e, i, j);
            }
            row[i + 1]++;
            if (j < s->n) row[j + 1]++;
        }
        for (u64 i = 0; i < s->n; i++) row[i + 1] += row[i];
        std::vector<unsigned int> col(2 * ne + 1);
        std::vector<TR> g(2 * ne + 1);
        std::vector<u64> fill(row.begin(), row.end() - 1);
        for (u64 e = 0; e < ne; e++) {
            u64 i = c->conn_i[e], j = c->conn_j[e];
            TR ge = host_double ? (TR)((const double*)c->conn_g)[e] : ((const TR*)c->conn_g)[e];
            col[fill[i]] = (unsigned int)j;
            g[fill[i]++] = ge;
            if (j < s->n) {
                col[fill[j]] = (unsigned int)i;
                g[fill[j]++] = ge;
            }
        }
        CUDA_TRY(cudaMalloc(&s->d_csr_row, (s->n + 1) * sizeof(u64)));
        CUDA_TRY(cudaMalloc(&s->d_csr_col, (2 * ne + 1) * sizeof(unsigned int)));
        CUDA_TRY(cudaMalloc(&s->d_csr_g, (2 * ne + 1) * sizeof(TR)));
        CUDA_TRY(h2d_sync(s, s->d_csr_row, row.data(), (s->n + 1) * sizeof(u64)));
        CUDA_TRY(h2d_sync(s, s->d_csr_col, col.data(), (2 * ne + 1) * sizeof(unsigned int)));
        CUDA_TRY(h2d_sync(s, s->d_csr_g, g.data(), (2 * ne + 1) * sizeof(TR)));
    }
    return MKB_OK;
}

// CUDA loads kernels lazily: the first launch of a kernel may need a
// context-wide synchronisation. A step kernel that is spinning on a
// neighbour's arrival flag would then wait for a kernel that cannot be loaded
// while it runs. Everything that can be launched during stepping is therefore
// loaded up front.
template <typename TR>
static int preload_kernels(mkb_sim* s) {
    cudaFuncAttributes a;
    CUDA_TRY(cudaFuncGetAttributes(&a, (const void*)s->kern));
    if (s->kern2) CUDA_TRY(cudaFuncGetAttributes(&a, (const void*)s->kern2));
    CUDA_TRY(cudaFuncGetAttributes(&a, (const void*)k_log_gather<TR>));
    CUDA_TRY(cudaFuncGetAttributes(&a, (const void*)k_fill_u32));
    CUDA_TRY(cudaFuncGetAttributes(&a, (const void*)k_push_ghosts<TR>));
    CUDA_TRY(cudaFuncGetAttributes(&a, (const void*)k_soa_to_aos<TR, TR>));
    CUDA_TRY(cudaFuncGetAttributes(&a, (const void*)k_soa_to_aos<TR, double>));
    return MKB_OK;
}

// MKB_DEBUG_TIMING=1: host wall-clock of the set-up phases on stderr.
static bool debug_timing() {
    static const bool on = getenv("MKB_DEBUG_TIMING") != nullptr;
    return on;
}
static double wall_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#define MKB_PHASE(label)                                                     \
    do {                                                                     \
        if (debug_timing()) {                                                \
            const double t_ = wall_s();                                      \
            fprintf(stderr, "[mkb] %-28s %9.3f ms\n", label, (t_ - t_phase) * 1e3); \
            t_phase = t_;                                                    \
        }                                                                    \
    } while (0)

// Prepares a run on the state that is resident on the device: log tables and
// rings, pacing (openclsim.c:488-496), schedule (:501, :1018-1022). Used by
// mkb_sim_init and mkb_sim_rearm.
static int arm_run(mkb_sim* s, const mkb_run_config* r) {
    NvtxRange nvtx_("mkb: arm run");
    if (!(r->dt > 0)) return fail(MKB_ERR_INVALID, "Step size must be greater than zero.");
    if (r->tmax < r->tmin) return fail(MKB_ERR_INVALID, "Simulation time can't be negative.");
    double t_phase = wall_s();
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->side));
    MKB_PHASE("arm: stream sync");

    // Drop the previous run's log (its device allocations are reused)
    s->d_tab_pre = s->d_tab_post = nullptr;
    s->d_log = nullptr;
    // (the pinned host matrix is kept and reused when it is large enough)
    s->h_log_cap = 0;
    s->pre_fields.clear();
    s->post_fields.clear();
    s->time_cols.clear();
    s->pace_cols.clear();
    s->row_time.clear();
    s->row_pace.clear();
    s->rows_written = s->rows_flushed = s->rows_final = 0;
    s->copied_pending[0] = s->copied_pending[1] = false;
    s->logging_states = false;
    s->store_aux = false;
    s->log_cap = s->log_half = 0;

    s->tmin = r->tmin;
    s->tmax = r->tmax;
    s->default_dt = r->dt;
    s->log_interval = r->log_interval;
    if (r->steps_per_call) {
        s->steps_per_call = r->steps_per_call;
    } else {
        s->steps_per_call = std::max<u64>(1000, 500 + 200000 / s->n);   // openclsim.c:1046-1047
    }

    // Logging tables
    std::vector<GatherEntry> pre, post;
    u64 col = 0;
    for (u64 j = 0; j < r->n_log; j++) {
        GatherEntry e;
        e.col = (unsigned int)col;
        e.pad = 0;
        e.off = 0;
        const u64 idx = r->log_index[j];
        bool bad = false;
        u64 width = 1;
        if (col > 0xfffffff0ull) return fail(MKB_ERR_INVALID, "too many log columns");
        switch (r->log_kind[j]) {
        case MKB_LOG_TIME:
            s->time_cols.push_back((int)col);
            break;
        case MKB_LOG_PACE:
            s->pace_cols.push_back((int)col);
            break;
        case MKB_LOG_STATE_FIELD: {
            if (idx >= (u64)s->n_state) { bad = true; break; }
            mkb_sim::FieldCopy f;
            f.plane = ((int)idx == s->i_vm) ? ~0ull : idx;
            f.col = col;
            s->pre_fields.push_back(f);
            s->logging_states = true;
            width = s->n;
            break;
        }
        case MKB_LOG_INTER_FIELD: {
            if (idx >= (u64)s->n_inter) { bad = true; break; }
            mkb_sim::FieldCopy f;
            f.plane = s->plane_inter + idx;
            f.col = col;
            s->post_fields.push_back(f);
            width = s->n;
            break;
        }
        case MKB_LOG_IDIFF_FIELD: {
            if (s->diff_mode == MKB_DIFF_NONE) { bad = true; break; }
            mkb_sim::FieldCopy f;
            f.plane = s->plane_idiff;
            f.col = col;
            s->post_fields.push_back(f);
            width = s->n;
            break;
        }
        case MKB_LOG_IDIFF:
            if (idx >= s->n || s->diff_mode == MKB_DIFF_NONE) { bad = true; break; }
            e.off = s->plane_idiff * s->stride + idx;
            post.push_back(e);
            break;
        case MKB_LOG_STATE: {
            const u64 cid = idx / (u64)s->n_state;
            const u64 k = idx % (u64)s->n_state;
            if (cid >= s->n) { bad = true; break; }
            e.off = ((int)k == s->i_vm) ? (MKB_GATHER_V | cid) : (k * s->stride + cid);
            pre.push_back(e);
            s->logging_states = true;
            break;
        }
        case MKB_LOG_INTER: {
            if (s->n_inter == 0) { bad = true; break; }
            const u64 cid = idx / (u64)s->n_inter;
            const u64 k = idx % (u64)s->n_inter;
            if (cid >= s->n) { bad = true; break; }
            e.off = (s->plane_inter + k) * s->stride + cid;
            post.push_back(e);
            break;
        }
        default:
            bad = true;
            break;
        }
        if (bad) return fail(MKB_ERR_INVALID, "Unknown variables found in logging dictionary.");
        col += width;
    }
    s->n_log = col;
    s->row_stride = col + 1;
    s->store_aux = !post.empty() || !s->post_fields.empty();
    if (s->logging_states) {
        // Hidden NaN probe: first state of cell 0 (openclsim.c:1087)
        GatherEntry e;
        e.col = (unsigned int)s->n_log;
        e.pad = 0;
        e.off = (s->i_vm == 0) ? (MKB_GATHER_V | 0ull) : 0ull;
        pre.push_back(e);
    }
    s->n_pre = pre.size();
    s->n_post = post.size();
    const std::vector<GatherEntry>* tabs[2] = {&pre, &post};
    for (int k = 0; k < 2; k++) {
        const u64 cnt = tabs[k]->size();
        if (!cnt) continue;
        if (cnt > s->d_tab_cap[k]) {
            cudaFree(s->d_tab_store[k]);
            s->d_tab_store[k] = nullptr;
            s->d_tab_cap[k] = 0;
            CUDA_TRY(cudaMalloc(&s->d_tab_store[k], cnt * sizeof(GatherEntry)));
            s->d_tab_cap[k] = cnt;
        }
        CUDA_TRY(cudaMemcpyAsync(s->d_tab_store[k], tabs[k]->data(), cnt * sizeof(GatherEntry),
                                 cudaMemcpyHostToDevice, s->stream));
        (k == 0 ? s->d_tab_pre : s->d_tab_post) = s->d_tab_store[k];
    }
    if (s->n_pre + s->n_post + s->pre_fields.size() + s->post_fields.size() > 0) {
        // Device row ring: two halves, ~64 MiB each at most
        const u64 row_bytes = s->row_stride * s->rs;
        const u64 half = std::max<u64>(1, std::min<u64>(4096, (64ull << 20) / row_bytes));
        s->log_half = half;
        s->log_cap = 2 * half;
        if (s->log_cap * row_bytes > s->d_log_bytes) {
            cudaFree(s->d_log_store);
            s->d_log_store = nullptr;
            s->d_log_bytes = 0;
            CUDA_TRY(cudaMalloc(&s->d_log_store, s->log_cap * row_bytes));
            s->d_log_bytes = s->log_cap * row_bytes;
        }
        s->d_log = s->d_log_store;
        CUDA_TRY(cudaMemsetAsync(s->d_log, 0, s->log_cap * row_bytes, s->stream));
    }
    MKB_PHASE("arm: tables + ring alloc");
    CUDA_TRY(cudaStreamSynchronize(s->stream));     // tables copied from stack vectors
    MKB_PHASE("arm: sync after memset");

    // Pacing and schedule: openclsim.c:488-502, 1018-1022
    int rc = s->sched.init(r->tmin, r->tmax, r->dt, r->log_interval, r->n_events, r->events);
    if (rc == mkb::PACING_SIMULTANEOUS_EVENT) {
        return fail(MKB_ERR_SIMULTANEOUS,
                    "E-Pacing error: Event scheduled or re-occuring at the same time as another event.");
    }
    if (rc) return fail(MKB_ERR_PACING, "E-Pacing error %d", rc);
    s->engine_time = r->tmin;
    s->finished = s->sched.finished();
    s->halted = false;
    // Sharded simulations whose last run ended normally keep counting: the
    // ghost rows / ghost cells for step_index + 1 are already in their slots
    // on every neighbour (the last step kernel delivered them), so a re-armed
    // run continues the exchange protocol where it stopped — no flag reset,
    // no barrier, no re-seed.
    if (!s->halo_live) {
        s->step_index = 0;
        // (stream order: after everything of the previous run, before the first step)
        if (s->d_tile_done) {
            CUDA_TRY(cudaMemsetAsync(s->d_tile_done, 0, s->n_tiles * sizeof(unsigned int), s->stream));
        }
    }
    s->issued = 0;
    s->throttle_count = 0;
    s->ring_chunk = 0;
    MKB_PHASE("arm: schedule init");
    return MKB_OK;
}

extern "C" int mkb_sim_init(const mkb_sim_config* c, mkb_sim** out) {
    NvtxRange nvtx_("mkb_sim_init");
    if (!c || !out) return fail(MKB_ERR_INVALID, "null argument");
    *out = nullptr;
    double t_phase = wall_s();
    if (c->abi_version != MKB_ABI_VERSION) {
        return fail(MKB_ERR_INVALID, "ABI version mismatch: library %d, caller %d", MKB_ABI_VERSION,
                    c->abi_version);
    }
    if (c->precision != MKB_SINGLE && c->precision != MKB_DOUBLE) {
        return fail(MKB_ERR_INVALID, "Only single and double precision are supported.");
    }
    if (c->host_precision != MKB_DOUBLE && c->host_precision != c->precision) {
        return fail(MKB_ERR_INVALID, "host_precision must be 64 or equal to precision");
    }
    if (c->nx < 1 || c->ny < 1) {
        return fail(MKB_ERR_INVALID, "The number of cells in any direction must be at least 1.");
    }
    if (c->n_state < 1 || c->n_inter < 0 || c->n_field < 0) return fail(MKB_ERR_INVALID, "bad model shape");
    if (c->diffusion_mode < 0 || c->diffusion_mode > 3) return fail(MKB_ERR_INVALID, "bad diffusion_mode");
    if (c->diffusion_mode != MKB_DIFF_NONE && (c->i_vm < 0 || c->i_vm >= c->n_state)) {
        return fail(MKB_ERR_INVALID, "i_vm out of range");
    }
    if (c->diffusion_mode == MKB_DIFF_CONNECTIONS && c->ny != 1) {
        return fail(MKB_ERR_INVALID, "Connections can only be specified in 1d mode.");
    }
    if (!c->cubin || !c->cubin_size || !c->kernel_name) return fail(MKB_ERR_INVALID, "no kernel image");
    if (!(c->dt > 0)) return fail(MKB_ERR_INVALID, "Step size must be greater than zero.");
    if (c->tmax < c->tmin) return fail(MKB_ERR_INVALID, "Simulation time can't be negative.");
    if (c->block_x < 1 || c->block_y < 1 || c->block_x * c->block_y > 1024) {
        return fail(MKB_ERR_INVALID, "bad block shape");
    }

    int ndev = 0;
    {
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) {
            return fail(MKB_ERR_CUDA, "No CUDA device available (%s); this back-end has no CPU fallback.",
                        cudaGetErrorString(e));
        }
    }
    if (c->device < 0 || c->device >= ndev) return fail(MKB_ERR_INVALID, "device ordinal out of range");
    CUDA_TRY(cudaSetDevice(c->device));

    mkb_sim* s = new mkb_sim();
    s->device = c->device;
    s->precision = c->precision;
    s->host_precision = c->host_precision;
    s->rs = c->precision == MKB_DOUBLE ? 8 : 4;
    s->hs = c->host_precision == MKB_DOUBLE ? 8 : 4;
    s->n_state = c->n_state;
    // Without diffusion nothing reads a neighbour's V: every state updates in place
    s->i_vm = c->diffusion_mode == MKB_DIFF_NONE ? -1 : c->i_vm;
    s->n_inter = c->n_inter;
    s->n_field = c->n_field;
    s->diff_mode = c->diffusion_mode;
    s->nx = c->nx;
    s->ny = c->ny;
    s->n = c->nx * c->ny;
    s->stride = (s->n + 31) / 32 * 32;
    if (c->kernel_stride && c->kernel_stride != s->stride) {
        const u64 want = s->stride;
        delete s;
        return fail(MKB_ERR_INVALID, "The kernel was compiled for a plane stride of %llu elements; this grid needs %llu.",
                    (unsigned long long)c->kernel_stride, (unsigned long long)want);
    }
    s->block_x = c->block_x;
    s->block_y = c->block_y;
    s->cpt = c->cells_per_thread > 1 ? c->cells_per_thread : 1;
    s->rpt = c->rows_per_thread > 1 ? c->rows_per_thread : 1;
    if (s->cpt > 1 && (c->nx % (u64)s->cpt) != 0) {
        delete s;
        return fail(MKB_ERR_INVALID, "nx must be a multiple of cells_per_thread");
    }
    s->use_graphs = c->use_graphs != 0;
    s->tmin = c->tmin;
    s->tmax = c->tmax;
    s->default_dt = c->dt;
    s->log_interval = c->log_interval;
    if (c->steps_per_call) {
        s->steps_per_call = c->steps_per_call;
    } else {
        // openclsim.c:1046-1047
        s->steps_per_call = std::max<u64>(1000, 500 + 200000 / s->n);
    }

#define INIT_TRY(expr)          \
    do {                        \
        int rc_ = (expr);       \
        if (rc_) {              \
            sim_destroy(s);     \
            return rc_;         \
        }                       \
    } while (0)
#define INIT_CUDA(expr)                                                                      \
    do {                                                                                     \
        cudaError_t e_ = (expr);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            sim_destroy(s);                                                                  \
            return fail(MKB_ERR_CUDA, "CUDA error %d (%s) at %s:%d: %s", (int)e_,            \
                        cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_));   \
        }                                                                                    \
    } while (0)

    INIT_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    INIT_CUDA(cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        INIT_CUDA(cudaEventCreateWithFlags(&s->ev_ring[i], cudaEventDisableTiming));
        INIT_CUDA(cudaEventCreateWithFlags(&s->ev_copied[i], cudaEventDisableTiming));
        INIT_CUDA(cudaEventCreateWithFlags(&s->ev_throttle[i], cudaEventDisableTiming));
    }
    INIT_CUDA(cudaEventCreateWithFlags(&s->ev_rows, cudaEventDisableTiming));
    INIT_CUDA(cudaEventCreate(&s->ev_t0));
    INIT_CUDA(cudaEventCreate(&s->ev_t1));

    MKB_PHASE("init: streams + events");
    // Model kernel
    INIT_CUDA(cudaLibraryLoadData(&s->lib, c->cubin, nullptr, nullptr, 0, nullptr, nullptr, 0));
    INIT_CUDA(cudaLibraryGetKernel(&s->kern, s->lib, c->kernel_name));
    // The caller says what else the image holds; a missing symbol is an error
    if (c->second_kernel_name && c->second_kernel_name[0]) {
        cudaError_t e2 = cudaLibraryGetKernel(&s->kern2, s->lib, c->second_kernel_name);
        if (e2 != cudaSuccess) {
            cudaGetLastError();
            sim_destroy(s);
            return fail(MKB_ERR_INVALID, "The kernel image does not contain the second kernel '%s' (%s).",
                        c->second_kernel_name, cudaGetErrorName(e2));
        }
    }
    s->persistent = (c->kernel_flags & MKB_KERNEL_PERSISTENT) != 0;
    if (s->persistent) {
        const u64 cpt = c->cells_per_thread > 0 ? (u64)c->cells_per_thread : 1;
        const u64 rpt = c->rows_per_thread > 0 ? (u64)c->rows_per_thread : 1;
        if (c->nx > (uint64_t)c->block_x || c->ny > (uint64_t)c->block_y || cpt != 1 || rpt != 1 ||
            c->diffusion_mode == MKB_DIFF_CONNECTIONS || c->n_ghost > 0 ||
            (c->ny_global && c->ny_global != c->ny)) {
            sim_destroy(s);
            return fail(MKB_ERR_INVALID, "The persistent kernel needs an unsharded grid that fits one thread block "
                                         "(%llu x %llu cells, block %d x %d).",
                        (unsigned long long)c->nx, (unsigned long long)c->ny, c->block_x, c->block_y);
        }
    }
    // (Measured on C3: asking for the maximum-L1 carve-out makes the step kernel
    // 50 % slower — 1.58 vs 1.05 ms — because the shared-memory tile then limits
    // residency; the driver's default split is kept.)

    MKB_PHASE("init: kernel image");
    // Buffers
    if (c->precision == MKB_DOUBLE) {
        INIT_TRY(preload_kernels<double>(s));
        MKB_PHASE("init: preload kernels");
        INIT_TRY(sim_init_typed<double>(s, c));
    } else {
        INIT_TRY(preload_kernels<float>(s));
        MKB_PHASE("init: preload kernels");
        INIT_TRY(sim_init_typed<float>(s, c));
    }
    MKB_PHASE("init: planes, state, fields");

    // Paced cells: rectangle stays symbolic, a list becomes a byte mask
    s->grid.pace_x0 = s->grid.pace_x1 = s->grid.pace_y0 = s->grid.pace_y1 = 0;
    if (c->diffusion_mode != MKB_DIFF_NONE) {
        if (c->pace_rect) {
            // openclsim.cl:260-274
            // clamped to what a 32-bit compare in the kernel can hold
            auto clamp32 = [](int64_t v) {
                return (long long)std::max<int64_t>(-(1ll << 30), std::min<int64_t>(v, 1ll << 30));
            };
            s->grid.pace_x0 = clamp32(c->pace_x);
            s->grid.pace_x1 = clamp32(c->pace_x + c->pace_nx);
            s->grid.pace_y0 = clamp32(c->pace_y);
            s->grid.pace_y1 = clamp32(c->pace_y + c->pace_ny);
        } else {
            std::vector<unsigned char> mask(s->n, 0);
            const u64 cell0 = c->iy_offset * c->nx;
            for (u64 i = 0; i < c->n_paced; i++) {
                u64 cid = c->paced_cells[i];
                if (cid >= cell0 && cid - cell0 < s->n) mask[cid - cell0] = 1;
            }
            INIT_CUDA(cudaMalloc(&s->d_mask, s->n));
            INIT_CUDA(h2d_sync(s, s->d_mask, mask.data(), s->n));
        }
    }

    // Schedule ring
    INIT_CUDA(cudaMalloc(&s->d_ring, 2 * kRingHalf * sizeof(MkbStepParams)));
    INIT_CUDA(cudaHostAlloc(&s->h_ring, 2 * kRingHalf * sizeof(MkbStepParams), cudaHostAllocDefault));

    // Grid arguments
    MkbGridArgs& g = s->grid;
    g.state = s->d_planes;
    g.idiff = plane_ptr<char>(s, s->plane_idiff);
    g.inter = plane_ptr<char>(s, s->plane_inter);
    g.field = plane_ptr<char>(s, s->plane_field);
    g.gx_field = s->d_gx;
    // slab-relative: element [-nx .. -1] is the row shared with the slab above
    g.gy_field = s->d_gy ? s->d_gy + s->nx * s->rs : nullptr;
    g.paced_mask = s->d_mask;
    g.csr_row = s->d_csr_row;
    g.csr_col = s->d_csr_col;
    g.csr_g = s->d_csr_g;
    g.halo_lo = nullptr;
    g.halo_hi = nullptr;
    g.flag_lo = nullptr;
    g.flag_hi = nullptr;
    g.peer_lo_halo_hi = nullptr;
    g.peer_hi_halo_lo = nullptr;
    g.peer_lo_flag_hi = nullptr;
    g.peer_hi_flag_lo = nullptr;
    g.halo_error = nullptr;
    {
        const u64 nyg = c->ny_global ? c->ny_global : s->ny;
        s->has_lo = c->iy_offset > 0;
        s->has_hi = c->iy_offset + s->ny < nyg;
        s->nbx = (s->nx + (u64)s->block_x * s->cpt - 1) / ((u64)s->block_x * s->cpt);
        if ((s->has_lo || s->has_hi) &&
            (s->diff_mode == MKB_DIFF_HOMOGENEOUS || s->diff_mode == MKB_DIFF_FIELD)) {
            // Exchange block: this GPU's ghost rows and arrival flags
            const size_t halo = 3 * s->nx * s->rs;
            const size_t flags = s->nbx * sizeof(unsigned int);
            s->xchg_bytes = 2 * halo + 2 * flags + 256;
            INIT_CUDA(cudaMalloc(&s->d_xchg, s->xchg_bytes));
            INIT_CUDA(cudaMemsetAsync(s->d_xchg, 0, s->xchg_bytes, s->stream));
            g.halo_error = (unsigned int*)(s->d_xchg + 2 * halo + 2 * flags);
        }
        g.ghost = nullptr;
        g.n_ghost = 0;
        if (s->diff_mode == MKB_DIFF_CONNECTIONS && c->n_ghost > 0) {
            // Exchange block: [ghost V: 3 x n_ghost][error][flags, sized at connect]
            s->n_ghost = c->n_ghost;
            const size_t ghost_bytes = (3 * s->n_ghost * s->rs + 255) / 256 * 256;
            s->xchg_bytes = ghost_bytes + 256 + 4096 * sizeof(unsigned int);
            INIT_CUDA(cudaMalloc(&s->d_xchg, s->xchg_bytes));
            INIT_CUDA(cudaMemsetAsync(s->d_xchg, 0, s->xchg_bytes, s->stream));
            g.ghost = s->d_xchg;
            g.n_ghost = s->n_ghost;
            g.halo_error = (unsigned int*)(s->d_xchg + ghost_bytes);
        }
    }
    g.nx = s->nx;
    g.ny = s->ny;
    g.stride = s->stride;
    g.iy_offset = c->iy_offset;
    g.ny_global = c->ny_global ? c->ny_global : s->ny;
    g.gx = c->gx;
    g.gy = c->gy;
    {
        const u64 cells_x = (u64)s->block_x * (u64)s->cpt;
        u64 bx = (s->nx + cells_x - 1) / cells_x;
        const u64 cells_y = (u64)s->block_y * (u64)s->rpt;
        u64 by = (s->ny + cells_y - 1) / cells_y;
        // (column blocks, row blocks mod 32768, row blocks / 32768): the kernel
        // reads its block coordinates without an integer division
        const u64 gy = std::min<u64>(by, 32768);
        const u64 gz = (by + gy - 1) / gy;
        if (bx > 0x7fffffffull || gz > 65535) {
            sim_destroy(s);
            return fail(MKB_ERR_INVALID, "grid too large for one launch");
        }
        s->launch_grid = dim3((unsigned int)bx, (unsigned int)gy, (unsigned int)gz);
        s->launch_block = dim3((unsigned int)s->block_x, (unsigned int)s->block_y, 1);
        if (c->kernel_flags & MKB_KERNEL_STREAM) {
            // Streaming kernel: a persistent grid walks the tiles; the two V
            // planes are described to the TMA unit as [ny][nx] tensors with
            // the kernel's box (halo included; cells outside arrive as zeros).
            if (s->i_vm < 0 || c->stream_box_w <= 0 || c->stream_box_h <= 0 || c->stream_box_w > 256 ||
                c->stream_box_h > 256 || (s->nx * s->rs) % 16 != 0 || (c->stream_box_w * s->rs) % 16 != 0) {
                sim_destroy(s);
                return fail(MKB_ERR_INVALID, "Streaming kernel: rows of %llu cells / a box of %d x %d cannot be "
                                             "described to the TMA unit (16-byte multiples, at most 256 per side).",
                            (unsigned long long)s->nx, c->stream_box_w, c->stream_box_h);
            }
            typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                          const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                          CUtensorMapFloatOOBfill);
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            INIT_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
            if (!fn || qres != cudaDriverEntryPointSuccess) {
                sim_destroy(s);
                return fail(MKB_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
            }
            static_assert(sizeof(CUtensorMap) == sizeof(g.tmap[0]), "tensor map size");
            const u64 planes[2] = {(u64)std::max(s->i_vm, 0), s->plane_alt_v};
            for (int k = 0; k < 2; k++) {
                const cuuint64_t dims[2] = {s->nx, s->ny};
                const cuuint64_t strides[1] = {s->nx * s->rs};
                const cuuint32_t box[2] = {(cuuint32_t)c->stream_box_w, (cuuint32_t)c->stream_box_h};
                const cuuint32_t estr[2] = {1, 1};
                CUresult r = ((encode_fn)fn)(
                    (CUtensorMap*)g.tmap[k], s->rs == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64,
                    2, s->d_planes + planes[k] * s->stride * s->rs, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) {
                    sim_destroy(s);
                    return fail(MKB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
                }
            }
            int sms = 0;
            INIT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
            const u64 per_sm = std::max<u64>(1, ((u64)c->kernel_flags >> MKB_KERNEL_FLAG_SHIFT_BLOCKS) & 0xff);
            const u64 tiles = bx * by;
            s->launch_grid = dim3((unsigned int)std::min<u64>(tiles, (u64)sms * per_sm), 1, 1);
        }
    }

    if (c->kernel_flags & MKB_KERNEL_STAGE) {
        // Staged kernel: the state planes as one 3-d tensor [plane][row][column]
        // with a box of one thread-block tile of one plane; the kernel asks for
        // kernel_smem_bytes of dynamic shared memory (above the 48 KiB default:
        // opt in).
        if (s->persistent || (c->kernel_flags & MKB_KERNEL_STREAM) || s->cpt != 1 || s->rpt != 1 ||
            c->kernel_smem_bytes <= 0 || (s->nx * s->rs) % 16 != 0 || (s->block_x * s->rs) % 16 != 0 ||
            s->block_x > 256 || s->block_y > 256) {
            sim_destroy(s);
            return fail(MKB_ERR_INVALID, "Staged kernel: rows of %llu cells / a tile of %d x %d cannot be described "
                                         "to the TMA unit (16-byte multiples, at most 256 per side), or the kernel "
                                         "form does not stage its states.",
                        (unsigned long long)s->nx, s->block_x, s->block_y);
        }
        typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                      CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                      CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        INIT_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) {
            sim_destroy(s);
            return fail(MKB_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        }
        static_assert(sizeof(CUtensorMap) == sizeof(g.tmap_state), "tensor map size");
        // (n_state planes and the second V plane right behind them, for L2 prefetches)
        const cuuint64_t dims[3] = {s->nx, s->ny, (cuuint64_t)s->n_state + (s->i_vm >= 0 ? 1u : 0u)};
        const cuuint64_t strides[2] = {s->nx * s->rs, s->stride * s->rs};
        const cuuint32_t box[3] = {(cuuint32_t)s->block_x, (cuuint32_t)s->block_y, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = ((encode_fn)fn)(
            (CUtensorMap*)g.tmap_state, s->rs == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64,
            3, s->d_planes, dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            sim_destroy(s);
            return fail(MKB_ERR_CUDA, "cuTensorMapEncodeTiled (state planes) failed (%d)", (int)r);
        }
        s->smem_bytes = (unsigned int)c->kernel_smem_bytes;
        cudaError_t ea = cudaFuncSetAttribute((const void*)s->kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)s->smem_bytes);
        if (ea != cudaSuccess) {
            sim_destroy(s);
            return fail(MKB_ERR_CUDA, "Staged kernel: %u bytes of shared memory per block refused (%s).",
                        s->smem_bytes, cudaGetErrorName(ea));
        }
        if (c->kernel_flags & MKB_KERNEL_TILE_LOOP) {
            int sms = 0;
            INIT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
            const u64 per_sm = std::max<u64>(1, ((u64)c->kernel_flags >> MKB_KERNEL_FLAG_SHIFT_BLOCKS) & 0xff);
            const u64 tiles = ((s->nx + s->block_x - 1) / s->block_x) * ((s->ny + s->block_y - 1) / s->block_y);
            s->launch_grid = dim3((unsigned int)std::min<u64>(tiles, (u64)sms * per_sm), 1, 1);
        }
        // (all of the SM's shared memory for the two resident blocks)
        cudaFuncSetAttribute((const void*)s->kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    }

    s->overlap = (c->kernel_flags & MKB_KERNEL_OVERLAP) != 0;
    if (s->overlap) {
        if (s->kern2 || s->persistent || (c->kernel_flags & MKB_KERNEL_STREAM)) {
            sim_destroy(s);
            return fail(MKB_ERR_INVALID, "Overlapping steps need a single ordinary step kernel.");
        }
        s->n_tiles = (size_t)s->launch_grid.x * s->launch_grid.y * s->launch_grid.z;
        INIT_CUDA(cudaMalloc(&s->d_tile_done, s->n_tiles * sizeof(unsigned int)));
        INIT_CUDA(cudaMemsetAsync(s->d_tile_done, 0, s->n_tiles * sizeof(unsigned int), s->stream));
        g.tile_done = s->d_tile_done;
    }

    // Logging, pacing and schedule of the first run
    {
        mkb_run_config r;
        memset(&r, 0, sizeof(r));
        r.tmin = c->tmin;
        r.tmax = c->tmax;
        r.dt = c->dt;
        r.log_interval = c->log_interval;
        r.n_events = c->n_events;
        r.events = c->events;
        r.n_log = c->n_log;
        r.log_kind = c->log_kind;
        r.log_index = c->log_index;
        r.steps_per_call = c->steps_per_call;
        INIT_TRY(arm_run(s, &r));
    }

    INIT_CUDA(cudaStreamSynchronize(s->stream));
#undef INIT_TRY
#undef INIT_CUDA
    *out = s;
    return MKB_OK;
}

static int ensure_host_rows(mkb_sim* s, u64 rows) {
    if (rows <= s->h_log_cap) return MKB_OK;
    const u64 row_bytes = s->row_stride * s->rs;
    // Estimate the whole run on first use, then grow geometrically
    u64 want = rows;
    if (s->h_log_cap == 0) {
        double est = (s->tmax - s->tmin) / s->log_interval;
        double est2 = 2.0 * (s->tmax - s->tmin) / s->default_dt;
        double m = std::min(est, est2) + 2;
        if (m > 0 && m < 4.0e9) want = std::max<u64>(rows, (u64)m);
        // Do not pin more than 2 GiB speculatively
        u64 lim = std::max<u64>(rows, (2ull << 30) / row_bytes);
        want = std::min(want, std::max<u64>(lim, rows));
        if (s->h_log && want * row_bytes <= s->h_log_bytes) {
            // the previous run's pinned matrix is large enough: reuse it
            s->h_log_cap = s->h_log_bytes / row_bytes;
            return MKB_OK;
        }
    } else {
        want = std::max(rows, s->h_log_cap * 2);
    }
    double t_phase = wall_s();
    CUDA_TRY(cudaStreamSynchronize(s->side));
    char* p = nullptr;
    CUDA_TRY(cudaHostAlloc(&p, want * row_bytes, cudaHostAllocDefault));
    MKB_PHASE("log: pinned alloc");
    if (s->h_log) {
        if (s->h_log_cap) memcpy(p, s->h_log, s->rows_flushed * row_bytes);
        // Unpinning is slow (measured: ~0.5 s for 100 MB) and synchronises the
        // device: outgrown matrices are retired and freed with the simulation,
        // unless they add up to more than 2 GiB.
        s->h_log_retired.push_back(s->h_log);
        s->h_log_retired_bytes += s->h_log_bytes;
        if (s->h_log_retired_bytes > (2ull << 30)) {
            for (char* q : s->h_log_retired) cudaFreeHost(q);
            s->h_log_retired.clear();
            s->h_log_retired_bytes = 0;
        }
    }
    MKB_PHASE("log: pinned copy + free");
    s->h_log = p;
    s->h_log_cap = want;
    s->h_log_bytes = want * row_bytes;
    return MKB_OK;
}

// Issues the D2H copy of rows [rows_flushed, rows_written) on the side stream.
static int flush_rows(mkb_sim* s) {
    NvtxRange nvtx_("mkb: flush log rows");
    if (s->rows_flushed == s->rows_written) return MKB_OK;
    int rc = ensure_host_rows(s, s->rows_written);
    if (rc) return rc;
    if (s->d_log) {
        const u64 row_bytes = s->row_stride * s->rs;
        const u64 slot = s->rows_flushed % s->log_cap;
        const int half = (int)(slot / s->log_half);
        const u64 nrows = s->rows_written - s->rows_flushed;
        CUDA_TRY(cudaEventRecord(s->ev_rows, s->stream));
        CUDA_TRY(cudaStreamWaitEvent(s->side, s->ev_rows, 0));
        CUDA_TRY(cudaMemcpyAsync(s->h_log + s->rows_flushed * row_bytes, s->d_log + slot * row_bytes,
                                 nrows * row_bytes, cudaMemcpyDeviceToHost, s->side));
        CUDA_TRY(cudaEventRecord(s->ev_copied[half], s->side));
        s->copied_pending[half] = true;
    }
    s->rows_flushed = s->rows_written;
    return MKB_OK;
}

template <typename TR>
static int finalize_rows(mkb_sim* s) {
    // Host-known columns and the NaN probe, for rows that have landed
    CUDA_TRY(cudaStreamSynchronize(s->side));
    for (u64 r = s->rows_final; r < s->rows_flushed; r++) {
        TR* row = (TR*)(s->h_log) + r * s->row_stride;
        const u64 k = r - s->rows_final;
        for (int col : s->time_cols) row[col] = (TR)s->row_time[k];
        for (int col : s->pace_cols) row[col] = (TR)s->row_pace[k];
        if (s->logging_states && !s->halted) {
            TR probe = row[s->n_log];
            if (probe != probe) {
                // openclsim.c:1087-1089: the row with the NaN is kept, then the run stops
                s->halted = true;
                s->finished = true;
                s->rows_flushed = s->rows_written = r + 1;
                s->rows_final = r + 1;
                s->row_time.clear();
                s->row_pace.clear();
                return MKB_OK;
            }
        }
    }
    s->rows_final = s->rows_flushed;
    s->row_time.clear();
    s->row_pace.clear();
    return MKB_OK;
}

static size_t ghost_flags_offset(u64 n_ghost, size_t rs) {
    return (3 * n_ghost * rs + 255) / 256 * 256 + 256;
}

// Pushes V(t) of the exported cells into the peers' slot for `step` and raises
// the flags to `step` (seeding uses step = 1 with the current V plane).
template <typename TR>
static int ghost_push(mkb_sim* s, const TR* v, unsigned int step, const MkbStepParams* sp = nullptr) {
    if (s->gpeers.empty()) return MKB_OK;
    u64 blocks = (s->n_export + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 4) blocks = 148 * 4;
    k_push_ghosts<TR><<<(unsigned int)blocks, 256, 0, s->stream>>>(
        v, s->d_exp_src, s->d_exp_slot, s->d_exp_peer, s->n_export, s->d_peer_base,
        s->d_peer_n_ghost, s->d_peer_flags, (unsigned int)s->gpeers.size(), step, sp, s->d_push_done);
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    return MKB_OK;
}

// One step-kernel launch. Overlapping kernels carry the programmatic-stream-
// serialization attribute: the launch may begin while the previous kernel in
// the stream is still running (once all its blocks have started); the blocks
// order themselves through MkbGridArgs::tile_done. Captured into a graph the
// attribute becomes a programmatic dependency edge.
static cudaError_t launch_step(mkb_sim* s, cudaKernel_t kern, void** args) {
    const unsigned int smem = (kern == s->kern) ? s->smem_bytes : 0u;
    if (!s->overlap) {
        return cudaLaunchKernel((const void*)kern, s->launch_grid, s->launch_block, args, smem, s->stream);
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = s->launch_grid;
    cfg.blockDim = s->launch_block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelExC(&cfg, (const void*)kern, args);
}

// Builds (once per slot and entry parity) the graph
//   [H2D: staged step parameters -> device] -> step kernel x kGraphSteps
// by stream capture. Each node has its parameter pointer and its V planes
// baked in; the per-step scalars travel through the staged copy.
template <typename TR>
static int graph_get(mkb_sim* s, int slot, int parity, cudaGraphExec_t* out) {
    mkb_sim::GraphSlot& gs = s->gslot[slot];
    if (!gs.h_stage) {
        CUDA_TRY(cudaHostAlloc(&gs.h_stage, kGraphSteps * sizeof(MkbStepParams), cudaHostAllocDefault));
        CUDA_TRY(cudaMalloc(&gs.d_params, kGraphSteps * sizeof(MkbStepParams)));
        CUDA_TRY(cudaEventCreateWithFlags(&gs.done, cudaEventDisableTiming));
    }
    if (!gs.exec[parity]) {
        const u64 vm_plane = (u64)std::max(s->i_vm, 0);
        cudaGraph_t graph = nullptr;
        CUDA_TRY(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        cudaError_t e = cudaMemcpyAsync(gs.d_params, gs.h_stage, kGraphSteps * sizeof(MkbStepParams),
                                        cudaMemcpyHostToDevice, s->stream);
        int par = parity;
        for (int j = 0; e == cudaSuccess && j < kGraphSteps; j++) {
            TR* v_in = plane_ptr<TR>(s, par ? s->plane_alt_v : vm_plane);
            TR* v_out = plane_ptr<TR>(s, par ? vm_plane : s->plane_alt_v);
            const MkbStepParams* sp = gs.d_params + j;
            void* args[] = {(void*)&s->grid, (void*)&sp, (void*)&v_in, (void*)&v_out};
            e = launch_step(s, s->kern, args);
            if (e == cudaSuccess && s->kern2) {
                e = cudaLaunchKernel((const void*)s->kern2, s->launch_grid, s->launch_block, args, 0, s->stream);
            }
            if (e == cudaSuccess && !s->gpeers.empty()) {
                // partitioned graphs: V(t + dt) of the exported cells -> the peers'
                // slot for the next step (the push reads its step number from sp)
                if (ghost_push<TR>(s, v_out, 0u, sp)) e = cudaErrorUnknown;
                s->launches--;      // counted when the graph is launched
            }
            par ^= 1;
        }
        cudaError_t e2 = cudaStreamEndCapture(s->stream, &graph);
        if (e == cudaSuccess) e = e2;
        if (e == cudaSuccess) e = cudaGraphInstantiate(&gs.exec[parity], graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
            gs.exec[parity] = nullptr;
            return fail(MKB_ERR_CUDA, "CUDA graph construction failed: %s", cudaGetErrorString(e));
        }
    }
    *out = gs.exec[parity];
    return MKB_OK;
}

template <typename TR>
static int sim_step_typed(mkb_sim* s, bool drain = true) {
    NvtxRange nvtx_("mkb: step batch");
    u64 steps_left = s->steps_per_call;
    bool timing_started = false;

    while (steps_left > 0 && !s->finished) {
        // ---- Phase A: host schedule for up to kRingHalf steps ----
        s->recs.clear();
        while (s->recs.size() < (size_t)kRingHalf && steps_left > 0) {
            StepRec rec;
            mkb::StepInfo info;
            int prc = s->sched.next(&info);
            if (prc == mkb::PACING_SIMULTANEOUS_EVENT) {
                return fail(MKB_ERR_SIMULTANEOUS,
                            "E-Pacing error: Event scheduled or re-occuring at the same time as another "
                            "event.");
            }
            if (prc) return fail(MKB_ERR_PACING, "E-Pacing error %d", prc);
            rec.logging = info.logging;
            rec.p.time = info.time;
            rec.p.dt = info.dt;
            rec.p.pace = info.pace;
            rec.p.flags = (rec.logging && s->store_aux) ? MKB_FLAG_STORE_AUX : 0u;
            rec.p.step = (unsigned int)(++s->step_index);
            rec.log_time = (double)(TR)info.time;       // openclsim.c:936-943: logged as Real
            rec.log_pace = (double)(TR)info.pace;
            s->engine_time = s->sched.time();
            s->recs.push_back(rec);
            steps_left--;
            if (s->sched.finished()) {
                s->finished = true;
                break;
            }
        }

        // ---- Phase B: upload the schedule chunk, then launch in order ----
        const int half = (int)(s->ring_chunk & 1);
        MkbStepParams* h = s->h_ring + half * kRingHalf;
        MkbStepParams* dring = s->d_ring + half * kRingHalf;
        if (s->ring_chunk >= 2) {
            // The H2D copy that last used this pinned half must have executed
            CUDA_TRY(cudaEventSynchronize(s->ev_ring[half]));
        }
        for (size_t i = 0; i < s->recs.size(); i++) h[i] = s->recs[i].p;
        if (s->persistent) {
            // consecutive unlogged steps share one launch; a logged step is a run of one
            s->runlen.assign(s->recs.size(), 1);
            for (size_t i = 0; i < s->recs.size();) {
                u64 len = 1;
                if (!s->recs[i].logging) {
                    while (i + len < s->recs.size() && !s->recs[i + len].logging && len < (1u << 20)) len++;
                }
                s->runlen[i] = len;
                h[i].flags |= (unsigned int)len << 8;
                i += len;
            }
        }
        if (!timing_started) {
            CUDA_TRY(cudaEventRecord(s->ev_t0, s->stream));
            timing_started = true;
        }
        CUDA_TRY(cudaMemcpyAsync(dring, h, s->recs.size() * sizeof(MkbStepParams),
                                 cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaEventRecord(s->ev_ring[half], s->stream));
        s->ring_chunk++;

        for (size_t i = 0; i < s->recs.size(); i++) {
            // (partitioned graphs: step + push pairs CAN be replayed as graphs — the push reads its
            // step from the device-side record — but measured slower at 8 GPUs: 0.042 vs 0.033 ms per step)
            if (s->use_graphs && !s->ghosts_connected && !s->partner && !s->persistent &&
                i + kGraphSteps <= s->recs.size()) {
                bool plain = true;
                for (int j = 0; j < kGraphSteps && plain; j++) plain = !s->recs[i + j].logging;
                if (plain) {
                    const int slot = (int)(s->graph_seq & 1);
                    mkb_sim::GraphSlot& gs = s->gslot[slot];
                    cudaGraphExec_t exec = nullptr;
                    int grc = graph_get<TR>(s, slot, s->parity, &exec);
                    if (grc) return grc;
                    if (gs.pending) {
                        // the launch that last used this slot's staging area
                        CUDA_TRY(cudaEventSynchronize(gs.done));
                        gs.pending = false;
                    }
                    for (int j = 0; j < kGraphSteps; j++) gs.h_stage[j] = s->recs[i + j].p;
                    CUDA_TRY(cudaGraphLaunch(exec, s->stream));
                    CUDA_TRY(cudaEventRecord(gs.done, s->stream));
                    gs.pending = true;
                    s->graph_seq++;
                    s->graph_launches++;
                    s->launches += ((s->kern2 ? 2 : 1) + (s->gpeers.empty() ? 0 : 1)) * kGraphSteps;
                    s->steps += kGraphSteps;
                    s->issued += kGraphSteps;
                    i += kGraphSteps - 1;       // parity unchanged: kGraphSteps is even
                    continue;
                }
            }
            const StepRec& rec = s->recs[i];
            if (s->issued++ % kThrottle == 0) {
                const int slot = (int)(s->throttle_count & 1);
                if (s->throttle_count >= 2) CUDA_TRY(cudaEventSynchronize(s->ev_throttle[slot]));
                CUDA_TRY(cudaEventRecord(s->ev_throttle[slot], s->stream));
                s->throttle_count++;
            }
            const u64 vm_plane = (u64)std::max(s->i_vm, 0);
            TR* v_in = plane_ptr<TR>(s, s->parity ? s->plane_alt_v : vm_plane);
            TR* v_out = plane_ptr<TR>(s, s->parity ? vm_plane : s->plane_alt_v);
            TR* row = nullptr;
            const bool dev_row = rec.logging && s->d_log;
            if (dev_row) {
                const u64 slot = s->rows_written % s->log_cap;
                if (slot % s->log_half == 0) {
                    const int lh = (int)(slot / s->log_half);
                    if (s->copied_pending[lh]) {
                        // rows previously in this half must have left the device
                        CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_copied[lh], 0));
                        s->copied_pending[lh] = false;
                    }
                }
                row = (TR*)s->d_log + slot * s->row_stride;
                if (s->n_pre) {
                    // states at time t (openclsim.c:1079-1084)
                    k_log_gather<TR><<<grid_for(s->n_pre), 256, 0, s->stream>>>(
                        plane_ptr<TR>(s, 0), v_in, s->d_tab_pre, s->n_pre, row);
                    s->launches++;
                }
                for (const mkb_sim::FieldCopy& f : s->pre_fields) {
                    const TR* src = (f.plane == ~0ull) ? v_in : plane_ptr<TR>(s, f.plane);
                    CUDA_TRY(cudaMemcpyAsync(row + f.col, src, s->n * sizeof(TR),
                                             cudaMemcpyDeviceToDevice, s->stream));
                }
            }
            if (s->partner) {
                // The other grid's PREVIOUS step wrote the V(t) plane this
                // kernel reads and was the last reader of the plane this
                // kernel overwrites; its kernel of the same step may overlap.
                // (An event that was never recorded does not block.)
                CUDA_TRY(cudaStreamWaitEvent(s->stream, s->partner->ev_step[(rec.p.step - 1u) & 1u], 0));
            }
            // fused diffusion + cell step: states -> t + dt (openclsim.c:1066-1096)
            const MkbStepParams* sp = dring + i;
            void* args[] = {(void*)&s->grid, (void*)&sp, (void*)&v_in, (void*)&v_out};
            CUDA_TRY(launch_step(s, s->kern, args));
            if (s->kern2) {
                // gates: reads V(t) (v_in) and the states only it updates
                CUDA_TRY(cudaLaunchKernel((const void*)s->kern2, s->launch_grid, s->launch_block, args, 0,
                                          s->stream));
                s->launches++;
            }
            if (s->partner) CUDA_TRY(cudaEventRecord(s->ev_step[rec.p.step & 1u], s->stream));
            s->launches++;
            if (s->persistent) {
                // the launch took runlen[i] steps and left V in the other plane once
                s->steps += s->runlen[i];
                i += (size_t)s->runlen[i] - 1;
            } else {
                s->steps++;
            }
            s->parity ^= 1;
            if (!s->gpeers.empty()) {
                // V(t + dt) of the exported cells -> the peers' slot for the next step
                int prc = ghost_push<TR>(s, v_out, rec.p.step + 1u);
                if (prc) return prc;
            }
            if (rec.logging) {
                if (dev_row && s->n_post) {
                    // idiff(t), intermediaries(t) (openclsim.c:1110-1119)
                    k_log_gather<TR><<<grid_for(s->n_post), 256, 0, s->stream>>>(
                        plane_ptr<TR>(s, 0), v_in, s->d_tab_post, s->n_post, row);
                    s->launches++;
                }
                if (dev_row) {
                    for (const mkb_sim::FieldCopy& f : s->post_fields) {
                        CUDA_TRY(cudaMemcpyAsync(row + f.col, plane_ptr<TR>(s, f.plane),
                                                 s->n * sizeof(TR), cudaMemcpyDeviceToDevice,
                                                 s->stream));
                    }
                }
                if (s->n_log > 0) {
                    s->row_time.push_back(rec.log_time);
                    s->row_pace.push_back(rec.log_pace);
                    s->rows_written++;
                    if (s->d_log && s->rows_written % s->log_half == 0) {
                        int rc = flush_rows(s);
                        if (rc) return rc;
                    }
                }
            }
        }
    }

    if (!drain) return MKB_OK;      // mkb_sim_step_pair drains both grids itself

    // End of call: drain, like the clFinish at openclsim.c:1172-1176
    if (timing_started) CUDA_TRY(cudaEventRecord(s->ev_t1, s->stream));
    int rc = flush_rows(s);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (timing_started) {
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, s->ev_t0, s->ev_t1));
        s->device_ms += ms;
    }
    if (s->grid.halo_error) {
        unsigned int err = 0;
        CUDA_TRY(cudaMemcpy(&err, s->grid.halo_error, sizeof(err), cudaMemcpyDeviceToHost));
        if (err) {
            return fail(MKB_ERR_CUDA, "Timed out waiting for a neighbouring GPU's boundary row "
                                      "(a rank stopped stepping or was never connected).");
        }
    }
    return finalize_rows<TR>(s);
}

extern "C" int mkb_sim_step(mkb_sim* s, double* engine_time, int* halted) {
    NvtxRange nvtx_("mkb_sim_step");
    if (!s) return fail(MKB_ERR_STATE, "Simulation not initialized.");
    if (s->n_ghost > 0 && !s->ghosts_connected) {
        return fail(MKB_ERR_STATE, "Graph partition not connected to its peers (mkb_sim_ghost_connect).");
    }
    if (s->d_xchg && !s->n_ghost && ((s->has_lo && !s->peer_lo_base) || (s->has_hi && !s->peer_hi_base))) {
        return fail(MKB_ERR_STATE, "Row slab not connected to its neighbours (mkb_sim_halo_connect).");
    }
    CUDA_TRY(cudaSetDevice(s->device));
    int rc = MKB_OK;
    if (!s->finished) {
        rc = (s->precision == MKB_DOUBLE) ? sim_step_typed<double>(s) : sim_step_typed<float>(s);
    }
    if (engine_time) *engine_time = s->engine_time;
    if (halted) *halted = s->halted ? 1 : 0;
    if (rc) return rc;
    return s->finished ? 0 : 1;
}

// ---------------------------------------------------------------------------
// Fibre-tissue pair: two grids, two kernels, every step together
// (myokit/_sim/fiber_tissue.c:1001-1155)
// ---------------------------------------------------------------------------
extern "C" int mkb_sim_junction_connect(mkb_sim* f, mkb_sim* t, double g, uint64_t cty) {
    if (!f || !t) return fail(MKB_ERR_STATE, "Simulation not initialized.");
    if (f == t) return fail(MKB_ERR_INVALID, "a junction needs two simulations");
    if (f->partner || t->partner) return fail(MKB_ERR_STATE, "already part of a pair");
    if (f->diff_mode != MKB_DIFF_HOMOGENEOUS || t->diff_mode != MKB_DIFF_HOMOGENEOUS) {
        return fail(MKB_ERR_INVALID, "a junction needs two homogeneous grids");
    }
    if (f->precision != t->precision || f->device != t->device) {
        return fail(MKB_ERR_INVALID, "both grids must use the same precision and device");
    }
    if (f->d_xchg || t->d_xchg || f->cpt != 1 || t->cpt != 1) {
        return fail(MKB_ERR_INVALID, "a junction needs unsharded grids with one cell per thread");
    }
    if (f->step_index != 0 || t->step_index != 0) {
        return fail(MKB_ERR_STATE, "junction_connect must precede the first step");
    }
    // fiber_tissue.py:168-170, 219-222
    if (f->ny > t->ny || cty + f->ny > t->ny) {
        return fail(MKB_ERR_INVALID, "The fiber y-dimension cannot exceed that of the tissue.");
    }
    CUDA_TRY(cudaSetDevice(f->device));
    mkb_sim* both[2] = {f, t};
    for (mkb_sim* s : both) {
        for (int k = 0; k < 2; k++) {
            if (!s->ev_step[k]) CUDA_TRY(cudaEventCreateWithFlags(&s->ev_step[k], cudaEventDisableTiming));
        }
    }
    auto plane = [](mkb_sim* s, bool alt) -> const void* {
        const u64 k = alt ? s->plane_alt_v : (u64)s->i_vm;
        return s->d_planes + k * s->stride * s->rs;
    };
    // fibre cell (nfx - 1, k) <-> tissue cell (0, cty + k)   (openclsim.cl:617-621)
    MkbGridArgs& gf = f->grid;
    gf.junction_v0 = plane(t, false);
    gf.junction_v1 = plane(t, true);
    gf.jg = g;
    gf.jx = f->nx - 1; gf.jy0 = 0; gf.jn = f->ny;
    gf.joff = 0 + cty * t->nx; gf.jstride = t->nx;
    MkbGridArgs& gt = t->grid;
    gt.junction_v0 = plane(f, false);
    gt.junction_v1 = plane(f, true);
    gt.jg = g;
    gt.jx = 0; gt.jy0 = cty; gt.jn = f->ny;
    gt.joff = f->nx - 1; gt.jstride = f->nx;
    f->partner = t;
    t->partner = f;
    return MKB_OK;
}

template <typename TR>
static int sim_drain_typed(mkb_sim* s) {
    NvtxRange nvtx_("mkb: drain log");
    int rc = flush_rows(s);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return finalize_rows<TR>(s);
}

extern "C" int mkb_sim_step_pair(mkb_sim* f, mkb_sim* t, uint64_t steps, double* engine_time, int* halted) {
    if (!f || !t) return fail(MKB_ERR_STATE, "Simulation not initialized.");
    if (f->partner != t || t->partner != f) return fail(MKB_ERR_STATE, "not a connected pair");
    if (f->parity != t->parity || f->finished != t->finished) {
        return fail(MKB_ERR_STATE, "the two grids of a pair must have taken the same steps");
    }
    CUDA_TRY(cudaSetDevice(f->device));
    const bool dp = f->precision == MKB_DOUBLE;
    const u64 keep_f = f->steps_per_call, keep_t = t->steps_per_call;
    f->steps_per_call = t->steps_per_call = 1;
    int rc = MKB_OK;
    for (uint64_t k = 0; k < steps && !f->finished && !t->finished && !rc; k++) {
        // one step each, strictly alternating: each kernel waits for the event
        // the other grid recorded after its previous step
        rc = dp ? sim_step_typed<double>(f, false) : sim_step_typed<float>(f, false);
        if (!rc) rc = dp ? sim_step_typed<double>(t, false) : sim_step_typed<float>(t, false);
    }
    f->steps_per_call = keep_f;
    t->steps_per_call = keep_t;
    if (!rc) rc = dp ? sim_drain_typed<double>(f) : sim_drain_typed<float>(f);
    if (!rc) rc = dp ? sim_drain_typed<double>(t) : sim_drain_typed<float>(t);
    if (engine_time) *engine_time = f->engine_time;
    if (halted) *halted = (f->halted || t->halted) ? 1 : 0;
    if (rc) return rc;
    if (f->halted || t->halted) f->finished = t->finished = true;
    return (f->finished && t->finished) ? 0 : 1;
}

extern "C" int mkb_sim_log_view(mkb_sim* s, const void** data, uint64_t* rows, uint64_t* cols,
                                uint64_t* row_stride) {
    if (!s) return fail(MKB_ERR_STATE, "Simulation not initialized.");
    if (data) *data = s->h_log;
    if (rows) *rows = s->rows_final;
    if (cols) *cols = s->n_log;
    if (row_stride) *row_stride = s->row_stride;
    return MKB_OK;
}

extern "C" int mkb_sim_get_state(mkb_sim* s, void* state_out) {
    NvtxRange nvtx_("mkb_sim_get_state");
    if (!s) return fail(MKB_ERR_STATE, "Simulation not initialized.");
    if (!state_out) return fail(MKB_ERR_INVALID, "state_out is null");
    CUDA_TRY(cudaSetDevice(s->device));
    if (s->precision == MKB_DOUBLE) return download_aos<double, double>(s, state_out);
    if (s->host_precision == MKB_DOUBLE) return download_aos<float, double>(s, state_out);
    return download_aos<float, float>(s, state_out);
}

template <typename TR>
static int set_state_typed(mkb_sim* s, const void* state_in, int uniform) {
    const bool host_double = (s->host_precision == MKB_DOUBLE);
    int rc;
    if (uniform) {
        rc = host_double ? upload_uniform<double, TR>(s, state_in, plane_ptr<TR>(s, 0), s->n_state)
                         : upload_uniform<TR, TR>(s, state_in, plane_ptr<TR>(s, 0), s->n_state);
    } else {
        rc = host_double ? upload_aos<double, TR>(s, state_in, plane_ptr<TR>(s, 0), s->n_state)
                         : upload_aos<TR, TR>(s, state_in, plane_ptr<TR>(s, 0), s->n_state);
    }
    return rc;
}

extern "C" int mkb_sim_set_state(mkb_sim* s, const void* state_in, int uniform) {
    NvtxRange nvtx_("mkb_sim_set_state");
    if (!s) return fail(MKB_ERR_STATE, "Simulation not initialized.");
    if (!state_in) return fail(MKB_ERR_INVALID, "state_in is null");
    CUDA_TRY(cudaSetDevice(s->device));
    // nothing of the previous run may still be in flight
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->side));
    int rc = (s->precision == MKB_DOUBLE) ? set_state_typed<double>(s, state_in, uniform)
                                          : set_state_typed<float>(s, state_in, uniform);
    if (rc) return rc;
    s->parity = 0;              // V(t) is in its own plane again
    s->state_replaced = true;   // ghost rows / cells on the neighbours are stale
    return MKB_OK;
}

extern "C" int mkb_sim_counters(mkb_sim* s, uint64_t* kernel_launches, uint64_t* steps) {
    if (!s) return fail(MKB_ERR_STATE, "Simulation not initialized.");
    if (kernel_launches) *kernel_launches = s->launches;
    if (steps) *steps = s->steps;
    return MKB_OK;
}

extern "C" int mkb_sim_device_ms(mkb_sim* s, double* ms) {
    if (!s) return fail(MKB_ERR_STATE, "Simulation not initialized.");
    if (ms) *ms = s->device_ms;
    return MKB_OK;
}

extern "C" int mkb_sim_set_steps_per_call(mkb_sim* s, uint64_t steps) {
    if (!s) return fail(MKB_ERR_STATE, "Simulation not initialized.");
    if (steps < 1) return fail(MKB_ERR_INVALID, "steps must be at least 1");
    s->steps_per_call = steps;
    return MKB_OK;
}

extern "C" int mkb_sim_reset_counters(mkb_sim* s) {
    if (!s) return fail(MKB_ERR_STATE, "Simulation not initialized.");
    s->launches = 0;
    s->steps = 0;
    s->device_ms = 0;
    return MKB_OK;
}

// ---------------------------------------------------------------------------
// Row-slab halo exchange
// ---------------------------------------------------------------------------
extern "C" int mkb_sim_halo_info(mkb_sim* s, int* has_lower, int* has_upper, uint64_t* bytes) {
    if (!s) return fail(MKB_ERR_STATE, "Simulation not initialized.");
    if (has_lower) *has_lower = (s->d_xchg && s->has_lo) ? 1 : 0;
    if (has_upper) *has_upper = (s->d_xchg && s->has_hi) ? 1 : 0;
    if (bytes) *bytes = s->xchg_bytes;
    return MKB_OK;
}

extern "C" int mkb_sim_halo_export(mkb_sim* s, void* ipc_handle_64, void** device_pointer) {
    if (!s) return fail(MKB_ERR_STATE, "Simulation not initialized.");
    if (!s->d_xchg) return fail(MKB_ERR_STATE, "This simulation has no neighbouring slabs.");
    CUDA_TRY(cudaSetDevice(s->device));
    if (ipc_handle_64) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        cudaIpcMemHandle_t h;
        CUDA_TRY(cudaIpcGetMemHandle(&h, s->d_xchg));
        memcpy(ipc_handle_64, &h, 64);
    }
    if (device_pointer) *device_pointer = s->d_xchg;
    return MKB_OK;
}

template <typename TR>
static int halo_seed_typed(mkb_sim* s);

// --- partitioned connection graphs (kernels are defined with the helpers above) ---
extern "C" int mkb_sim_ghost_connect(mkb_sim* s, uint32_t n_flags, uint32_t n_peers,
                                     const mkb_ghost_peer* peers, int direct, uint32_t n_import,
                                     const uint32_t* import_flags) {
    if (!s) return fail(MKB_ERR_STATE, "Simulation not initialized.");
    if (s->diff_mode != MKB_DIFF_CONNECTIONS) return fail(MKB_ERR_STATE, "not a connection graph");
    if (s->step_index != 0 || s->ghosts_connected) return fail(MKB_ERR_STATE, "ghost_connect must be called once, before the first step");
    if (n_flags > 4096) return fail(MKB_ERR_INVALID, "too many ranks");
    if ((n_peers && !peers) || (n_import && !import_flags)) return fail(MKB_ERR_INVALID, "null argument");
    if ((n_import > 0 || s->n_ghost > 0) && !s->d_xchg) return fail(MKB_ERR_STATE, "no exchange block");
    CUDA_TRY(cudaSetDevice(s->device));
    s->n_flags = n_flags;
    std::vector<unsigned int*> flag_ptrs;
    std::vector<u64> exp_src, exp_slot;
    std::vector<unsigned int> exp_peer;
    for (uint32_t k = 0; k < n_peers; k++) {
        const mkb_ghost_peer& in = peers[k];
        mkb_sim::GhostPeerRt gp;
        if (!in.handle) return fail(MKB_ERR_INVALID, "missing peer handle");
        if (direct) {
            void* ptr = *(void* const*)in.handle;
            cudaPointerAttributes attr;
            CUDA_TRY(cudaPointerGetAttributes(&attr, ptr));
            if (attr.device != s->device) {
                int can = 0;
                CUDA_TRY(cudaDeviceCanAccessPeer(&can, s->device, attr.device));
                if (!can) return fail(MKB_ERR_CUDA, "GPU %d cannot access GPU %d", s->device, attr.device);
                cudaError_t e = cudaDeviceEnablePeerAccess(attr.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_TRY(e);
                cudaGetLastError();
            }
            gp.base = ptr;
        } else {
            cudaIpcMemHandle_t h;
            memcpy(&h, in.handle, 64);
            CUDA_TRY(cudaIpcOpenMemHandle(&gp.base, h, cudaIpcMemLazyEnablePeerAccess));
            gp.ipc = true;
        }
        gp.peer_n_ghost = in.peer_n_ghost;
        gp.n_export = in.n_export;
        if (in.flag_index >= in.peer_n_flags) return fail(MKB_ERR_INVALID, "flag index out of range");
        gp.flag = (unsigned int*)((char*)gp.base + ghost_flags_offset(in.peer_n_ghost, s->rs)) + in.flag_index;
        for (u64 e = 0; e < gp.n_export; e++) {
            if (in.src_cell[e] >= s->n || in.dst_slot[e] >= in.peer_n_ghost) {
                return fail(MKB_ERR_INVALID, "export entry out of range");
            }
            exp_src.push_back(in.src_cell[e]);
            exp_slot.push_back(in.dst_slot[e]);
            exp_peer.push_back(k);
        }
        s->gpeers.push_back(gp);
        flag_ptrs.push_back(gp.flag);
    }
    if (!flag_ptrs.empty()) {
        std::vector<void*> bases;
        std::vector<u64> counts;
        for (auto& gp : s->gpeers) {
            bases.push_back(gp.base);
            counts.push_back(gp.peer_n_ghost);
        }
        const size_t np = flag_ptrs.size();
        CUDA_TRY(cudaMalloc(&s->d_peer_flags, np * sizeof(unsigned int*)));
        CUDA_TRY(h2d_sync(s, s->d_peer_flags, flag_ptrs.data(), np * sizeof(unsigned int*)));
        CUDA_TRY(cudaMalloc(&s->d_peer_base, np * sizeof(void*)));
        CUDA_TRY(h2d_sync(s, s->d_peer_base, bases.data(), np * sizeof(void*)));
        CUDA_TRY(cudaMalloc(&s->d_peer_n_ghost, np * sizeof(u64)));
        CUDA_TRY(h2d_sync(s, s->d_peer_n_ghost, counts.data(), np * sizeof(u64)));
        CUDA_TRY(cudaMalloc(&s->d_push_done, sizeof(unsigned int)));
        CUDA_TRY(cudaMemset(s->d_push_done, 0, sizeof(unsigned int)));
        s->n_export = exp_src.size();
        if (s->n_export) {
            CUDA_TRY(cudaMalloc(&s->d_exp_src, s->n_export * sizeof(u64)));
            CUDA_TRY(cudaMalloc(&s->d_exp_slot, s->n_export * sizeof(u64)));
            CUDA_TRY(cudaMalloc(&s->d_exp_peer, s->n_export * sizeof(unsigned int)));
            CUDA_TRY(h2d_sync(s, s->d_exp_src, exp_src.data(), s->n_export * sizeof(u64)));
            CUDA_TRY(h2d_sync(s, s->d_exp_slot, exp_slot.data(), s->n_export * sizeof(u64)));
            CUDA_TRY(h2d_sync(s, s->d_exp_peer, exp_peer.data(), s->n_export * sizeof(unsigned int)));
        }
    }
    s->n_import = n_import;
    if (n_import) {
        for (uint32_t k = 0; k < n_import; k++) {
            if (import_flags[k] >= n_flags) return fail(MKB_ERR_INVALID, "import flag out of range");
        }
        CUDA_TRY(cudaMalloc(&s->d_import, n_import * sizeof(unsigned int)));
        CUDA_TRY(h2d_sync(s, s->d_import, import_flags, n_import * sizeof(unsigned int)));
        // the step kernel itself waits for these flags (no separate launch)
        s->grid.ghost_flags = (const unsigned int*)(s->d_xchg + ghost_flags_offset(s->n_ghost, s->rs));
        s->grid.ghost_import = s->d_import;
        s->grid.n_ghost_import = n_import;
    }
    s->ghosts_connected = true;
    s->halo_connected = true;
    return mkb_sim_halo_seed(s);
}

extern "C" int mkb_sim_rearm(mkb_sim* s, const mkb_run_config* r) {
    NvtxRange nvtx_("mkb_sim_rearm");
    if (!s) return fail(MKB_ERR_STATE, "Simulation not initialized.");
    if (!r) return fail(MKB_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(s->device));
    // (a run that halted on a NaN or was abandoned half way leaves ranks at
    // different steps: start the exchange protocol afresh then)
    s->halo_live = s->d_xchg && s->halo_connected && s->finished && !s->halted && !s->state_replaced;
    s->state_replaced = false;
    int rc = arm_run(s, r);
    if (rc) return rc;
    // counters describe one run
    s->launches = 0;
    s->steps = 0;
    s->device_ms = 0;
    if (s->d_xchg && !s->halo_live) {
        // Arrival flags restart from zero; the caller barriers, then reseeds
        const size_t keep = s->n_ghost ? (3 * s->n_ghost * s->rs + 255) / 256 * 256
                                       : 2 * 3 * s->nx * s->rs;
        CUDA_TRY(cudaMemsetAsync(s->d_xchg + keep, 0, s->xchg_bytes - keep, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
    }
    return MKB_OK;
}

extern "C" int mkb_sim_halo_live(mkb_sim* s, int* live) {
    if (!s || !live) return fail(MKB_ERR_STATE, "Simulation not initialized.");
    *live = s->halo_live ? 1 : 0;
    return MKB_OK;
}

extern "C" int mkb_sim_halo_seed(mkb_sim* s) {
    if (!s) return fail(MKB_ERR_STATE, "Simulation not initialized.");
    if (s->halo_live) return MKB_OK;    // nothing to deliver: see arm_run
    if (s->ghosts_connected) {
        if (s->step_index != 0) return fail(MKB_ERR_STATE, "halo_seed must precede the first step");
        CUDA_TRY(cudaSetDevice(s->device));
        const u64 vm = (u64)std::max(s->i_vm, 0);
        int rc;
        if (s->precision == MKB_DOUBLE) {
            rc = ghost_push<double>(s, plane_ptr<double>(s, s->parity ? s->plane_alt_v : vm), 1u);
        } else {
            rc = ghost_push<float>(s, plane_ptr<float>(s, s->parity ? s->plane_alt_v : vm), 1u);
        }
        if (rc) return rc;
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        return MKB_OK;
    }
    if (!s->d_xchg) return MKB_OK;
    if (s->step_index != 0) return fail(MKB_ERR_STATE, "halo_seed must precede the first step");
    if ((s->has_lo && !s->peer_lo_base) || (s->has_hi && !s->peer_hi_base)) {
        return fail(MKB_ERR_STATE, "Row slab not connected to its neighbours (mkb_sim_halo_connect).");
    }
    CUDA_TRY(cudaSetDevice(s->device));
    return s->precision == MKB_DOUBLE ? halo_seed_typed<double>(s) : halo_seed_typed<float>(s);
}

template <typename TR>
static int halo_connect_typed(mkb_sim* s) {
    MkbGridArgs& g = s->grid;
    const size_t halo = 3 * s->nx * s->rs;
    const size_t flags = s->nbx * sizeof(unsigned int);
    // Own block
    g.halo_lo = s->has_lo ? s->d_xchg : nullptr;
    g.halo_hi = s->has_hi ? s->d_xchg + halo : nullptr;
    g.flag_lo = (const unsigned int*)(s->d_xchg + 2 * halo);
    g.flag_hi = (const unsigned int*)(s->d_xchg + 2 * halo + flags);
    // Neighbours' blocks have the same layout (same nx, same column blocks)
    if (s->peer_lo_base) {
        char* b = (char*)s->peer_lo_base;
        g.peer_lo_halo_hi = b + halo;
        g.peer_lo_flag_hi = (unsigned int*)(b + 2 * halo + flags);
    }
    if (s->peer_hi_base) {
        char* b = (char*)s->peer_hi_base;
        g.peer_hi_halo_lo = b;
        g.peer_hi_flag_lo = (unsigned int*)(b + 2 * halo);
    }
    s->halo_connected = true;
    return halo_seed_typed<TR>(s);
}

// Delivers V(tmin) of the boundary rows into the neighbours' slot for step 1.
template <typename TR>
static int halo_seed_typed(mkb_sim* s) {
    MkbGridArgs& g = s->grid;
    const TR* v = plane_ptr<TR>(s, s->parity ? s->plane_alt_v : (u64)std::max(s->i_vm, 0));
    const size_t row = s->nx * s->rs;
    const int grid = (int)((s->nbx + 255) / 256);
    if (s->peer_lo_base) {
        CUDA_TRY(cudaMemcpyAsync((char*)g.peer_lo_halo_hi + 1 * row, v, row, cudaMemcpyDefault, s->stream));
        k_fill_u32<<<grid, 256, 0, s->stream>>>(g.peer_lo_flag_hi, s->nbx, 1u);
        s->launches++;
    }
    if (s->peer_hi_base) {
        CUDA_TRY(cudaMemcpyAsync((char*)g.peer_hi_halo_lo + 1 * row, v + (s->ny - 1) * s->nx, row,
                                 cudaMemcpyDefault, s->stream));
        k_fill_u32<<<grid, 256, 0, s->stream>>>(g.peer_hi_flag_lo, s->nbx, 1u);
        s->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return MKB_OK;
}

extern "C" int mkb_sim_halo_connect(mkb_sim* s, const void* lower, const void* upper, int direct) {
    if (!s) return fail(MKB_ERR_STATE, "Simulation not initialized.");
    if (!s->d_xchg) return fail(MKB_ERR_STATE, "This simulation has no neighbouring slabs.");
    if (s->step_index != 0) return fail(MKB_ERR_STATE, "halo_connect must precede the first step");
    if ((s->has_lo && !lower) || (s->has_hi && !upper)) {
        return fail(MKB_ERR_INVALID, "missing neighbour handle");
    }
    CUDA_TRY(cudaSetDevice(s->device));
    const void* in[2] = {s->has_lo ? lower : nullptr, s->has_hi ? upper : nullptr};
    void** out[2] = {&s->peer_lo_base, &s->peer_hi_base};
    bool* ipc[2] = {&s->peer_lo_ipc, &s->peer_hi_ipc};
    for (int k = 0; k < 2; k++) {
        if (!in[k]) continue;
        if (direct) {
            // Same process: `in` is the neighbour's device pointer
            void* ptr = *(void* const*)in[k];
            cudaPointerAttributes attr;
            CUDA_TRY(cudaPointerGetAttributes(&attr, ptr));
            if (attr.device != s->device) {
                int can = 0;
                CUDA_TRY(cudaDeviceCanAccessPeer(&can, s->device, attr.device));
                if (!can) return fail(MKB_ERR_CUDA, "GPU %d cannot access GPU %d", s->device, attr.device);
                cudaError_t e = cudaDeviceEnablePeerAccess(attr.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_TRY(e);
                cudaGetLastError();
            }
            *out[k] = ptr;
            *ipc[k] = false;
        } else {
            cudaIpcMemHandle_t h;
            memcpy(&h, in[k], 64);
            void* ptr = nullptr;
            CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
            *out[k] = ptr;
            *ipc[k] = true;
        }
    }
    return s->precision == MKB_DOUBLE ? halo_connect_typed<double>(s) : halo_connect_typed<float>(s);
}

extern "C" void mkb_sim_clean(mkb_sim* s) { sim_destroy(s); }

// ---------------------------------------------------------------------------
// Pipe micro-benchmarks: the denominators of the compute rooflines
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_peak_fma(T* out, int iters, T a, T b) {
    // 8 independent chains per thread: enough ILP to saturate the pipe at
    // full occupancy
    T x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    T x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
            x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void __launch_bounds__(256) k_peak_mufu(float* out, int iters, float a) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 0.1f, x2 = x0 + 0.2f, x3 = x0 + 0.3f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x2));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x3));
        }
        x0 *= a; x1 *= a; x2 *= a; x3 *= a;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}

__global__ void __launch_bounds__(256) k_peak_copy(const float4* __restrict__ in,
                                                   float4* __restrict__ out, u64 n) {
    u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    const u64 step = (u64)gridDim.x * blockDim.x;
    for (; i + 3 * step < n; i += 4 * step) {
        float4 a = in[i], b = in[i + step], c = in[i + 2 * step], d = in[i + 3 * step];
        out[i] = a; out[i + step] = b; out[i + 2 * step] = c; out[i + 3 * step] = d;
    }
    for (; i < n; i += step) out[i] = in[i];
}

// Fills out[0..5]: fp64 FMA Ginstr/s (thread-level), fp32 FMA Ginstr/s, MUFU.EX2
// Gop/s, copy GB/s (read + write), SM clock MHz seen by the driver, SM count.
extern "C" int mkb_measure_peaks(int device, double* out6) {
    if (!out6) return fail(MKB_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int sms = prop.multiProcessorCount;
    const int blocks = sms * 8;
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    void* buf = nullptr;
    CUDA_TRY(cudaMalloc(&buf, (size_t)blocks * 256 * 8));
    float ms = 0;
    auto timed = [&](auto launch) -> double {
        double best = 1e30;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        return best;
    };
    const int iters = 4096;
    const double fma_per_thread = (double)iters * 64;
    double t = timed([&] { k_peak_fma<double><<<blocks, 256>>>((double*)buf, iters, 1.0000001, 1e-9); });
    out6[0] = fma_per_thread * blocks * 256 / (t * 1e-3) / 1e9;
    t = timed([&] { k_peak_fma<float><<<blocks, 256>>>((float*)buf, iters, 1.0000001f, 1e-9f); });
    out6[1] = fma_per_thread * blocks * 256 / (t * 1e-3) / 1e9;
    t = timed([&] { k_peak_mufu<<<blocks, 256>>>((float*)buf, iters, 0.5f); });
    out6[2] = (double)iters * 32 * blocks * 256 / (t * 1e-3) / 1e9;
    cudaFree(buf);
    const u64 nbytes = 2ull << 30;
    void *a = nullptr, *b = nullptr;
    CUDA_TRY(cudaMalloc(&a, nbytes));
    CUDA_TRY(cudaMalloc(&b, nbytes));
    CUDA_TRY(cudaMemset(a, 1, nbytes));
    t = timed([&] { k_peak_copy<<<sms * 16, 256>>>((const float4*)a, (float4*)b, nbytes / 16); });
    out6[3] = 2.0 * nbytes / (t * 1e-3) / 1e9;
    cudaFree(a);
    cudaFree(b);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    out6[4] = khz / 1000.0;
    out6[5] = sms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CUDA_TRY(cudaGetLastError());
    return MKB_OK;
}
