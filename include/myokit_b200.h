/*
 * myokit_b200.h — C ABI of libmyokit_b200.so: the B200 (sm_100a) back-end for
 * the multi-cell time-stepping path of myokit.SimulationOpenCL.
 *
 * The reference's native boundary for this path is a generated CPython
 * extension exporting exactly sim_init / sim_step / sim_clean
 * (myokit/_sim/openclsim.c:1216-1221). The entry points below are what a
 * binding for that path would call instead; they take plain pointers and
 * sizes, no Python and no torch types:
 *
 *   reference (file:line)                            this library
 *   -----------------------------------------------  ---------------------------
 *   clBuildProgram of the rendered kernel            mkb_jit_compile
 *     (openclsim.c:815-836)
 *   sim_init, 21 arguments (openclsim.c:309-1031,    mkb_sim_init(mkb_sim_config)
 *     format string :385-407)
 *   sim_step (openclsim.c:1036-1211)                 mkb_sim_step
 *   log lists appended per sample (:1122-1131)       mkb_sim_log_view
 *   state_out filled in place (:1185-1190)           mkb_sim_get_state
 *   sim_clean (openclsim.c:216-304)                  mkb_sim_clean
 *   ESys_* pacing (pacing.h:240-600)                 events passed as doubles in
 *                                                    mkb_sim_config; mkb_pacing_probe
 *   mcl_select_device / mcl_info (mcl.h:265,823)     mkb_device_count / mkb_device_info
 *
 * All functions return 0 on success and a negative code on failure unless
 * stated otherwise; mkb_last_error() returns a message for the calling thread.
 * There is no CPU fallback: without a CUDA device every compute entry point
 * fails with MKB_ERR_CUDA.
 */
#ifndef MYOKIT_B200_H
#define MYOKIT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MKB_ABI_VERSION 5

/* Error codes */
#define MKB_OK               0
#define MKB_ERR_INVALID     -1   /* bad argument */
#define MKB_ERR_CUDA        -2   /* CUDA runtime / driver error (no device, OOM, launch) */
#define MKB_ERR_JIT         -3   /* NVRTC compilation failed (see log) */
#define MKB_ERR_PACING      -4   /* pacing error other than simultaneous events */
#define MKB_ERR_SIMULTANEOUS -5  /* myokit.SimultaneousProtocolEventError (pacing.h:112) */
#define MKB_ERR_STATE       -6   /* call out of order (not initialised / finished) */

/* Precision constants: same values as myokit.SINGLE_PRECISION / DOUBLE_PRECISION */
#define MKB_SINGLE 32
#define MKB_DOUBLE 64

/* Diffusion modes */
#define MKB_DIFF_NONE        0   /* diffusion=False: every cell paced (openclsim.cl:329) */
#define MKB_DIFF_HOMOGENEOUS 1   /* set_conductance      (openclsim.cl:384-435) */
#define MKB_DIFF_FIELD       2   /* set_conductance_field (openclsim.cl:452-487) */
#define MKB_DIFF_CONNECTIONS 3   /* set_connections      (openclsim.cl:537-573) */

/* Log column kinds; `index` uses the REFERENCE's array-of-structs numbering
 * (openclsim.c:936-1001) so a binding can reuse the reference's key table. */
#define MKB_LOG_TIME  0          /* global, (Real)engine_time */
#define MKB_LOG_PACE  1          /* global, (Real)engine_pace */
#define MKB_LOG_IDIFF 2          /* index = cid */
#define MKB_LOG_STATE 3          /* index = cid * n_state + k */
#define MKB_LOG_INTER 4          /* index = cid * n_inter + k */
/* Whole-field entries (no reference equivalent; the reference needs one dict
 * key per cell): one entry expands to n_cells consecutive columns, cell-id
 * order, copied plane-to-row on the device without an index table. */
#define MKB_LOG_STATE_FIELD 5    /* index = k: state k of every cell */
#define MKB_LOG_INTER_FIELD 6    /* index = k: logged intermediary k of every cell */
#define MKB_LOG_IDIFF_FIELD 7    /* diffusion current of every cell */

typedef struct mkb_sim mkb_sim;

typedef struct mkb_device_info_t {
    char name[256];
    int cc_major, cc_minor;
    int sm_count;
    int clock_khz;
    size_t total_mem;
    size_t l2_bytes;
    size_t smem_per_block_optin;
} mkb_device_info_t;

typedef struct mkb_sim_config {
    int abi_version;            /* MKB_ABI_VERSION */
    int device;                 /* CUDA device ordinal */
    int precision;              /* MKB_SINGLE | MKB_DOUBLE: arithmetic type Real */
    int host_precision;         /* dtype of state_in/field_data/gx_field/gy_field/conn_g
                                   and of mkb_sim_get_state: MKB_DOUBLE (the reference's
                                   Python floats) or equal to `precision` */

    /* Compiled model kernel (from mkb_jit_compile) */
    const void* cubin;
    size_t cubin_size;
    const char* kernel_name;    /* "mkb_cell_step" */
    int block_x, block_y;       /* thread block the kernel was generated for */
    int cells_per_thread;       /* x-adjacent cells per thread (0 or 1: one) */
    int rows_per_thread;        /* rows per thread (0 or 1: one) */

    /* Model shape */
    int n_state;
    int i_vm;                   /* index of membrane potential in the state (diffusion on) */
    int n_inter;                /* logged intermediary variables */
    int n_field;                /* set_field variables */

    /* Geometry: cid = ix + iy * nx (openclsim.cl:316) */
    uint64_t nx, ny;
    int diffusion_mode;         /* MKB_DIFF_* */
    double gx, gy;
    const void* gx_field;       /* [ny][nx-1] */
    const void* gy_field;       /* [ny-1][nx], may be null when ny == 1 */
    uint64_t n_ghost;           /* partitioned graphs: cells owned by other GPUs that
                                   this partition's edges reach; an edge endpoint
                                   conn_j >= nx names ghost cell conn_j - nx */
    uint64_t n_connections;     /* edges (i < j, g), openclsim.py:1405-1449 */
    const uint64_t* conn_i;
    const uint64_t* conn_j;
    const void* conn_g;

    /* Pacing */
    int pace_rect;              /* 1: rectangle below; 0: explicit list */
    int64_t pace_nx, pace_ny, pace_x, pace_y;   /* openclsim.py:1569-1593 */
    uint64_t n_paced;           /* explicit list (openclsim.py:1612-1629) */
    const uint64_t* paced_cells;
    int n_events;               /* protocol events, 5 doubles each: */
    const double* events;       /* level, start, duration, period, multiplier */

    /* Time */
    double tmin, tmax;
    double dt;                  /* default step size */
    double log_interval;

    /* Initial data, reference layouts */
    const void* state_in;       /* [n_cells * n_state], state_in[cid * n_state + k] */
    const void* field_data;     /* [n_cells * n_field], field_data[cid * n_field + k] */

    /* Logging */
    uint64_t n_log;
    const int32_t* log_kind;    /* MKB_LOG_* */
    const uint64_t* log_index;

    /* Row-slab sharding (multi-GPU); single GPU: iy_offset = 0, ny_global = ny */
    uint64_t iy_offset;
    uint64_t ny_global;

    /* Tuning (0 = default) */
    uint64_t steps_per_call;    /* steps before mkb_sim_step returns (openclsim.c:1046-1047) */
    int use_graphs;             /* 1: replay batches of steps as CUDA graphs */

    /* What the image is expected to hold beyond `kernel_name`; a missing
     * symbol is an error (never a silent single-kernel run). */
    int state_uniform;          /* 1: state_in holds ONE cell's n_state values, which
                                   every cell starts from (the reference's default:
                                   openclsim.py:236 tiles the model's initial state) */
    const char* second_kernel_name; /* NULL, or the kernel launched after every
                                       `kernel_name` launch ("mkb_gate_step") */
    int kernel_flags;           /* MKB_KERNEL_* */
    int stream_box_w, stream_box_h; /* MKB_KERNEL_STREAM: the TMA box (cells, rows) the
                                       kernel loads per tile, halo included */
    int kernel_smem_bytes;      /* MKB_KERNEL_STAGE: dynamic shared memory per thread block */
    uint64_t kernel_stride;     /* 0, or the plane stride (elements) the kernel was
                                   compiled for: nx * ny rounded up to 32. Anything
                                   else is refused (the kernel would trap) */
} mkb_sim_config;

/* kernel_flags */
#define MKB_KERNEL_PERSISTENT 1 /* kernel_name takes `flags >> 8` steps per launch; the
                                   grid fits one thread block */
#define MKB_KERNEL_STREAM     2 /* kernel_name is a streaming kernel: a persistent grid of
                                   min(tiles, sm_count * blocks_per_sm) thread blocks walks
                                   the tiles of the grid; the V tile + halo arrives by TMA
                                   through the descriptors in MkbGridArgs::tmap */
#define MKB_KERNEL_OVERLAP    4 /* consecutive launches of kernel_name overlap: launched with
                                   programmatic stream serialization, ordered by the
                                   per-block step counters in MkbGridArgs::tile_done */
#define MKB_KERNEL_STAGE      8 /* kernel_name stages its thread block's tile of every state
                                   plane in `kernel_smem_bytes` of dynamic shared memory by
                                   TMA, through the 3-d descriptor MkbGridArgs::tmap_state
                                   ([plane][row][column], box block_x x block_y x 1) */
#define MKB_KERNEL_TILE_LOOP 16 /* with MKB_KERNEL_STAGE: a fixed grid of
                                   min(tiles, sm_count * blocks_per_sm) thread blocks, block b
                                   taking tiles b, b + grid, ... (x fastest) */
#define MKB_KERNEL_FLAG_SHIFT_BLOCKS 8   /* bits 8..15: thread blocks per SM (stream, tile loop) */

/* A run on the state that is already resident on the device (mkb_sim_rearm):
 * the time span, step size, protocol and log selection of mkb_sim_config. */
typedef struct mkb_run_config {
    double tmin, tmax;
    double dt;
    double log_interval;
    int n_events;
    const double* events;
    uint64_t n_log;
    const int32_t* log_kind;
    const uint64_t* log_index;
    uint64_t steps_per_call;
} mkb_run_config;

/* ---- library ---- */
int mkb_abi_version(void);
const char* mkb_last_error(void);
void mkb_free(void* p);

/* Page-locked host memory for states and logs the caller wants moved at PCIe
 * speed (mkb_sim_config::state_in, mkb_sim_get_state: any host pointer works,
 * pinned ones are copied by asynchronous DMA). */
int mkb_host_alloc(size_t bytes, void** out);
void mkb_host_free(void* p);

/* ---- devices (replaces mcl.h device selection / info) ---- */
int mkb_device_count(void);                                 /* < 0 on error */
int mkb_device_info(int device, mkb_device_info_t* out);

/* ---- JIT ---- */
/* Text of the device ABI header generated kernels #include as "mkb_device_abi.h". */
const char* mkb_device_abi_header(void);
/* Compiles CUDA C++ `source` to an sm_100a cubin with NVRTC. `options` is a
 * NUL-separated, double-NUL-terminated list of extra NVRTC options or NULL.
 * On return *cubin (malloc'ed, free with mkb_free) holds the image and *log
 * (malloc'ed, may be empty) the compiler log, also on failure. */
int mkb_jit_compile(const char* source, const char* options,
                    void** cubin, size_t* cubin_size, char** log);

/* ---- simulation (replaces sim_init / sim_step / sim_clean) ---- */
int mkb_sim_init(const mkb_sim_config* cfg, mkb_sim** out);
/* Starts another run from the state the device holds (the reference keeps
 * its state in a Python list between runs, openclsim.py:1104,1149; here it
 * stays in HBM). Row slabs: call on every rank, then either step straight
 * away (mkb_sim_halo_live says 1) or barrier, mkb_sim_halo_seed, barrier. */
int mkb_sim_rearm(mkb_sim* sim, const mkb_run_config* run);
/* Replaces the resident state between runs (before mkb_sim_rearm) without
 * rebuilding the simulation: `state_in` as in mkb_sim_config, `uniform` as
 * mkb_sim_config::state_uniform. The reference re-creates its buffers and
 * uploads the state on every run (openclsim.c:512-562). */
int mkb_sim_set_state(mkb_sim* sim, const void* state_in, int uniform);
/* Runs up to steps_per_call time steps. Returns 1 while t < tmax, 0 when the
 * run has finished (final state available), < 0 on error. *engine_time gets
 * the current time. *halted (may be null) is set when a NaN was found in the
 * first state of cell 0 at a logged step (openclsim.c:1087). */
int mkb_sim_step(mkb_sim* sim, double* engine_time, int* halted);
/* Logged rows so far: pinned host matrix of Real; element (r, c) of the log is
 * data[r * row_stride + c] for c < cols. Columns follow the order of the log
 * entries; a *_FIELD entry takes n_cells columns. Valid until mkb_sim_clean,
 * mkb_sim_rearm or the next mkb_sim_step. */
int mkb_sim_log_view(mkb_sim* sim, const void** data, uint64_t* rows, uint64_t* cols,
                     uint64_t* row_stride);
/* Copies the state, reference layout [cid * n_state + k], host_precision. */
int mkb_sim_get_state(mkb_sim* sim, void* state_out);
/* Counters: kernels launched by this library for this simulation, steps taken. */
int mkb_sim_counters(mkb_sim* sim, uint64_t* kernel_launches, uint64_t* steps);
/* Device time (ms) spent between the first and last step kernel of the calls
 * made so far, measured with CUDA events on the launching stream. */
int mkb_sim_device_ms(mkb_sim* sim, double* ms);
/* Changes the number of steps the next mkb_sim_step calls take (>= 1). */
int mkb_sim_set_steps_per_call(mkb_sim* sim, uint64_t steps);
/* Zeroes the launch / step counters and the accumulated device time. */
int mkb_sim_reset_counters(mkb_sim* sim);
void mkb_sim_clean(mkb_sim* sim);

/* ---- row-slab sharding over several GPUs (no reference equivalent) ----
 * A slab (iy_offset / ny_global in mkb_sim_config) with neighbours owns an
 * "exchange block" in its own HBM: ghost rows of V (3 slots) and arrival flags.
 * Neighbouring slabs write their boundary rows straight into it from the step
 * kernel (peer stores over NVLink) and bump the flags; the step kernel of the
 * owner waits on the flags. Sequence per rank: mkb_sim_init; halo_export;
 * exchange handles (any transport, e.g. torch.distributed); halo_connect;
 * barrier; mkb_sim_step ... */
int mkb_sim_halo_info(mkb_sim* sim, int* has_lower, int* has_upper, uint64_t* bytes);
/* ipc_handle_64: 64 bytes (cudaIpcMemHandle_t) for another process;
 * device_pointer: the raw pointer, for a neighbour in the same process. */
int mkb_sim_halo_export(mkb_sim* sim, void* ipc_handle_64, void** device_pointer);
/* lower / upper: the neighbours' exports (rows below iy_offset / above the
 * slab); direct = 0: pointers to 64-byte IPC handles, direct = 1: pointers to
 * device pointers of sims living in this process. Null where no neighbour. */
int mkb_sim_halo_connect(mkb_sim* sim, const void* lower, const void* upper, int direct);
/* After mkb_sim_rearm (and a barrier): delivers the boundary rows of the
 * current state to the neighbours again. mkb_sim_halo_connect includes it. */
int mkb_sim_halo_seed(mkb_sim* sim);
/* After mkb_sim_rearm: *live = 1 when the exchange protocol simply continues
 * (the previous run ended normally on every rank, so each neighbour already
 * holds this slab's boundary row for the next step): no barrier and no
 * mkb_sim_halo_seed are needed before stepping. Every rank gets the same
 * answer, because they all took the same steps. */
int mkb_sim_halo_live(mkb_sim* sim, int* live);

/* Partitioned connection graphs: a partition (contiguous cell ids, n_ghost > 0)
 * owns an exchange block [ghost V: 3 x n_ghost][flags: one per rank]. Every
 * step, after the step kernel, the owner of a cell pushes its new V into the
 * ghost slots of the partitions that reference it and then raises its flag
 * there; a partition's next step first waits for the flags of the ranks it
 * imports from. Same handles and sequence as the row-slab calls above
 * (mkb_sim_halo_export / this call instead of mkb_sim_halo_connect /
 * mkb_sim_halo_seed after a re-arm). */
typedef struct mkb_ghost_peer {
    const void* handle;         /* peer's export: IPC handle, or pointer to its device pointer */
    uint64_t peer_n_ghost;      /* the peer's n_ghost (layout of its block) */
    uint32_t peer_n_flags;      /* the peer's flag count */
    uint32_t flag_index;        /* which of the peer's flags this rank raises */
    uint64_t n_export;          /* cells of this partition the peer needs */
    const uint64_t* src_cell;   /* their local ids here */
    const uint64_t* dst_slot;   /* their ghost slots there */
} mkb_ghost_peer;
int mkb_sim_ghost_connect(mkb_sim* sim, uint32_t n_flags, uint32_t n_peers,
                          const mkb_ghost_peer* peers, int direct,
                          uint32_t n_import, const uint32_t* import_flags);

/* ---- fibre-tissue pair (myokit/_sim/fiber_tissue.py:17, fiber_tissue.c:1001-1155) ----
 * Two homogeneous 2-d simulations (two models, kernels generated with
 * junction='fiber' / 'tissue') that take every step together: fibre cell
 * (nfx - 1, k) is tied to tissue cell (0, cty + k), k < nfy, with conductance
 * g (diff_step_fiber_tissue, myokit/_sim/openclsim.cl:601-628). Both must be
 * initialised with the same time span, step size, log interval and protocol.
 * mkb_sim_step_pair replaces mkb_sim_step for the pair: up to `steps` steps of
 * both, strictly alternating; returns 1 while there is more to do, 0 when both
 * have finished, < 0 on error; *halted as for mkb_sim_step (either grid). */
int mkb_sim_junction_connect(mkb_sim* fiber, mkb_sim* tissue, double g, uint64_t cty);
int mkb_sim_step_pair(mkb_sim* fiber, mkb_sim* tissue, uint64_t steps,
                      double* engine_time, int* halted);

/* ---- roofline denominators ----
 * Micro-benchmarks of the pipes the cell step is bound by. out6: fp64 FMA
 * Ginstr/s (thread-level), fp32 FMA Ginstr/s, MUFU.EX2 Gop/s, device copy GB/s
 * (read + write), SM clock MHz (driver attribute), SM count. */
int mkb_measure_peaks(int device, double* out6);

/* ---- the host schedule alone (unit tests, no GPU needed) ----
 * Runs the step selection of the time loop (openclsim.c:1051-1178) without
 * any device work: for each step its start time, size, pacing level and
 * whether a log row is written. Arrays may be null. Returns 0 when the run
 * ended within max_steps, 1 if it was cut off, < 0 on a pacing error. */
int mkb_schedule_probe(double tmin, double tmax, double dt, double log_interval,
                       int n_events, const double* events, uint64_t max_steps,
                       double* times, double* dts, double* paces,
                       unsigned char* logging, uint64_t* n_steps);

/* ---- pacing alone (unit tests; mirrors tests/ansic_event_based_pacing.c) ---- */
int mkb_pacing_probe(double t0, int n_events, const double* events,
                     int n_times, const double* times,
                     double* levels, double* next_times);

#ifdef __cplusplus
}
#endif
#endif
