/*
 * oracle/driver.c — CPU ORACLE (TEST INFRASTRUCTURE ONLY; never on the product path).
 *
 * Restates the host loop of the reference's SimulationOpenCL back-end:
 *   - time-step schedule, logging order, pacing advance:
 *       myokit/_sim/openclsim.c:1036-1211 (sim_step), :488-502, :1018-1022
 *   - event pacing system (ESys_*): myokit/_sim/pacing.h:74-80, 192-218, 483-548
 *   - diffusion kernels: myokit/_sim/openclsim.cl:384-435 (diff_step),
 *       :452-487 (diff_hetero), :537-573 (diff_arb_step / diff_arb_reset; the
 *       atomics are replaced by a serial, fixed-order edge loop)
 *
 * Compiled together with a model header that defines N_STATE, N_INTER,
 * N_FIELD, I_VM, Real and cell_step():
 *   - ORACLE_REF_KERNEL undefined: header from oracle/cgen.py (the port)
 *   - ORACLE_REF_KERNEL defined:   the reference's own rendered openclsim.cl
 *       behind oracle/cl_shim.h (built into oracle/_ref/ by oracle/build_ref.py)
 *
 * Plain C, doubles at the ABI; values are cast to Real exactly where the
 * reference casts (openclsim.c:1063, 1148, 1155).
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ORACLE_REF_KERNEL
#include "cl_shim.h"
#endif
#include ORACLE_MODEL_HEADER

/* ------------------------------------------------------------------------ */
/* Event pacing: restatement of ESys (pacing.h)                              */
/* ------------------------------------------------------------------------ */
typedef struct {
    double level, start, duration, period, multiplier;
    int next;                   /* index of next event in queue, -1 = none */
} OEvent;

typedef struct {
    int n;
    OEvent* ev;
    int head, fire;             /* indices, -1 = none */
    double time, tnext, tdown, level;
} OPacing;

static double o_scale(double a, double b)
{
    return fabs(a) > fabs(b) ? fabs(a) : fabs(b);
}
/* pacing.h:74 */
static int o_eq(double a, double b)
{
    return (a == b) || (fabs(a - b) / o_scale(a, b) < DBL_EPSILON);
}
/* pacing.h:80 */
static int o_geq(double a, double b)
{
    return (a >= b) || o_eq(a, b);
}

/* pacing.h:192-218. Returns new head; *flag = 1 on simultaneous events. */
static int o_schedule(OEvent* ev, int head, int add, int* flag)
{
    int e;
    *flag = 0;
    ev[add].next = -1;
    if (head < 0) return add;
    if (ev[add].start < ev[head].start) {
        ev[add].next = head;
        return add;
    }
    e = head;
    while (ev[e].next >= 0 && ev[add].start >= ev[ev[e].next].start) {
        e = ev[e].next;
    }
    if (ev[add].start == ev[e].start) *flag = 1;
    ev[add].next = ev[e].next;
    ev[e].next = add;
    return head;
}

/* pacing.h:240-262 + 289-326 (create + populate + reset) */
static int o_pacing_init(OPacing* p, double t0, int n, const double* events)
{
    int i, flag, head;
    p->n = n;
    p->ev = NULL;
    p->head = p->fire = -1;
    p->time = t0;
    p->tnext = t0;
    p->tdown = t0;
    p->level = 0;
    if (n <= 0) return 0;
    p->ev = (OEvent*)malloc(sizeof(OEvent) * (size_t)n);
    for (i = 0; i < n; i++) {
        p->ev[i].level = events[5 * i + 0];
        p->ev[i].start = events[5 * i + 1];
        p->ev[i].duration = events[5 * i + 2];
        p->ev[i].period = events[5 * i + 3];
        p->ev[i].multiplier = events[5 * i + 4];
        p->ev[i].next = -1;
        if (p->ev[i].period == 0 && p->ev[i].multiplier != 0) return -24;
        if (p->ev[i].period < 0) return -23;
        if (p->ev[i].multiplier < 0) return -25;
    }
    head = 0;
    for (i = 1; i < n; i++) {
        head = o_schedule(p->ev, head, i, &flag);
        if (flag) return -50;
    }
    p->head = head;
    return 0;
}

/* pacing.h:483-548 */
static int o_pacing_advance(OPacing* p, double new_time)
{
    int flag;
    if (new_time < p->time) return -40;
    p->time = new_time;
    while (o_geq(p->time, p->tnext)) {
        if (p->fire >= 0 && o_geq(p->tnext, p->tdown)) {
            p->fire = -1;
            p->level = 0;
        }
        if (p->head >= 0 && o_geq(p->tnext, p->ev[p->head].start)) {
            OEvent* f;
            p->fire = p->head;
            f = &p->ev[p->fire];
            p->head = f->next;
            p->tdown = f->start + f->duration;
            p->level = f->level;
            if (f->period > 0) {
                if (f->multiplier != 1) {
                    if (f->multiplier > 1) f->multiplier--;
                    f->start += f->period;
                    p->head = o_schedule(p->ev, p->head, p->fire, &flag);
                    if (flag) return -50;
                } else {
                    f->period = 0;
                }
            }
            if (p->head >= 0 && o_eq(p->ev[p->head].start, p->tdown)) {
                p->tdown = p->ev[p->head].start;
            }
        }
        p->tnext = HUGE_VAL;
        if (p->fire >= 0 && p->tnext > p->tdown) p->tnext = p->tdown;
        if (p->head >= 0 && p->tnext > p->ev[p->head].start)
            p->tnext = p->ev[p->head].start;
    }
    return 0;
}

/*
 * Stand-alone pacing probe (for tests against myokit.PacingSystem and the
 * known answers of myokit/tests/test_pacing_system_c.py): advances to each of
 * times[i] and reports level and next-event time after each advance.
 */
int oracle_pacing_probe(
    double t0, int n_events, const double* events,
    int n_times, const double* times, double* levels, double* tnexts)
{
    OPacing p;
    int i, flag;
    flag = o_pacing_init(&p, t0, n_events, events);
    if (flag) { free(p.ev); return flag; }
    for (i = 0; i < n_times; i++) {
        flag = o_pacing_advance(&p, times[i]);
        if (flag) { free(p.ev); return flag; }
        levels[i] = p.level;
        tnexts[i] = p.tnext;
    }
    free(p.ev);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Diffusion (restated), used when the reference's own kernels are absent    */
/* ------------------------------------------------------------------------ */
#ifndef ORACLE_REF_KERNEL

/* openclsim.cl:384-435 */
static void diff_step_cell(
    size_t ix, size_t iy, size_t nx, size_t ny, Real gx, Real gy,
    const Real* state, Real* idiff)
{
    const size_t cid = ix + iy * nx;
    const size_t of1 = cid * N_STATE + I_VM;
    size_t ofp, ofm;
    if (nx > 1) {
        ofp = of1 + N_STATE;
        ofm = of1 - N_STATE;
        if (ix == 0) {
            idiff[cid] = gx * (state[of1] - state[ofp]);
        } else if (ix == nx - 1) {
            idiff[cid] = gx * (state[of1] - state[ofm]);
        } else {
            idiff[cid] = gx * (2 * state[of1] - state[ofm] - state[ofp]);
        }
    } else {
        idiff[cid] = 0;
    }
    if (ny > 1) {
        ofp = of1 + N_STATE * nx;
        ofm = of1 - N_STATE * nx;
        if (iy == 0) {
            idiff[cid] += gy * (state[of1] - state[ofp]);
        } else if (iy == ny - 1) {
            idiff[cid] += gy * (state[of1] - state[ofm]);
        } else {
            idiff[cid] += gy * (2 * state[of1] - state[ofm] - state[ofp]);
        }
    }
}

/* openclsim.cl:452-487 */
static void diff_hetero_cell(
    size_t ix, size_t iy, size_t nx, size_t ny, const Real* gx, const Real* gy,
    const Real* state, Real* idiff)
{
    const size_t cid = ix + iy * nx;
    const size_t off = cid * N_STATE + I_VM;
    Real i = 0.0;
    Real v = state[off];
    if (nx > 1) {
        if (ix > 0) { i += gx[cid - iy - 1] * (v - state[off - N_STATE]); }
        if (ix < nx - 1) { i += gx[cid - iy] * (v - state[off + N_STATE]); }
    }
    if (ny > 1) {
        if (iy > 0) i += gy[cid - nx] * (v - state[off - N_STATE * nx]);
        if (iy < ny - 1) i += gy[cid] * (v - state[off + N_STATE * nx]);
    }
    idiff[cid] = i;
}

#endif /* !ORACLE_REF_KERNEL */

/* ------------------------------------------------------------------------ */
/* The run                                                                   */
/* ------------------------------------------------------------------------ */

/* Log column kinds */
#define LOG_TIME  0
#define LOG_PACE  1
#define LOG_IDIFF 2     /* index = cid */
#define LOG_STATE 3     /* index = cid * N_STATE + k */
#define LOG_INTER 4     /* index = cid * N_INTER + k */

int oracle_n_state(void) { return N_STATE; }
int oracle_n_inter(void) { return N_INTER; }
int oracle_n_field(void) { return N_FIELD; }
int oracle_real_size(void) { return (int)sizeof(Real); }

/*
 * diffusion_mode: 0 none, 1 homogeneous grid, 2 heterogeneous grid,
 *                 3 connections (1-d ids).
 * state: [nx*ny*N_STATE] doubles, in/out (values are rounded to Real).
 * paced: [nx*ny] bytes (only read when diffusion_mode != 0).
 * events: n_events * (level, start, duration, period, multiplier).
 * log_out: [max_rows * n_log] doubles.
 * Returns 0, or a negative pacing error flag, or 1 if log_out overflowed.
 */
/* Wall-clock seconds the last oracle_run spent in its time-step loop
 * (sim_step's loop, openclsim.c:1051-1178), without set-up and state copies:
 * what bench.py's CPU arms report, so that a 20-step call and a 900-step call
 * measure the same thing. */
static double g_loop_seconds = 0;
double oracle_loop_seconds(void) { return g_loop_seconds; }
static double o_now(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int oracle_run(
    size_t nx, size_t ny, int diffusion_mode,
    double gx_in, double gy_in,
    const double* gx_field_in, const double* gy_field_in,
    size_t n_conn, const uint64_t* conn1, const uint64_t* conn2,
    const double* conn_g_in,
    double tmin, double tmax, double default_dt, double log_interval,
    double* state_io, const double* field_in,
    const unsigned char* paced,
    int n_events, const double* events,
    size_t n_log, const int* log_kind, const uint64_t* log_index,
    double* log_out, size_t max_rows,
    uint64_t* n_rows_out, uint64_t* n_steps_out, int* halted_out,
    double* final_time_out, int nthreads)
{
    const size_t n = nx * ny;
    size_t i;
    int rc = 0;
    OPacing pacing;
    Real *state, *idiff, *inter_log, *field_data;
    Real *gxf = NULL, *gyf = NULL, *cg = NULL;
    Real *snap_state = NULL;
    const Real gx = (Real)gx_in, gy = (Real)gy_in;
    double engine_time = tmin, engine_pace, tnext_pace, tnext_log, dt, d;
    const double dt_min = 0;                            /* openclsim.c:413 */
    unsigned long istep, inext_log;
    uint64_t n_rows = 0, n_steps = 0;
    int halt = 0, logging_states = 0;
    Real arg_time, arg_dt, arg_pace;

    (void)nthreads;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif

    state = (Real*)malloc(sizeof(Real) * (n * N_STATE + 1));
    idiff = (Real*)calloc(n + 1, sizeof(Real));
    inter_log = (Real*)calloc(n * N_INTER + 1, sizeof(Real));
    field_data = (Real*)malloc(sizeof(Real) * (n * N_FIELD + 1));
    for (i = 0; i < n * N_STATE; i++) state[i] = (Real)state_io[i];
    for (i = 0; i < n * N_FIELD; i++) field_data[i] = (Real)field_in[i];
    if (diffusion_mode == 2) {
        size_t ngx = (nx - 1) * ny, ngy = nx * (ny - 1);
        gxf = (Real*)malloc(sizeof(Real) * (ngx + 1));
        gyf = (Real*)malloc(sizeof(Real) * (ngy + 1));
        for (i = 0; i < ngx; i++) gxf[i] = (Real)gx_field_in[i];
        for (i = 0; i < ngy; i++) gyf[i] = (Real)gy_field_in[i];
    }
    if (diffusion_mode == 3) {
        cg = (Real*)malloc(sizeof(Real) * (n_conn + 1));
        for (i = 0; i < n_conn; i++) cg[i] = (Real)conn_g_in[i];
    }
    for (i = 0; i < n_log; i++) {
        if (log_kind[i] == LOG_STATE) logging_states = 1;
    }
    if (logging_states) {
        snap_state = (Real*)malloc(sizeof(Real) * (n * N_STATE + 1));
    }

    /* Pacing: openclsim.c:488-496 */
    rc = o_pacing_init(&pacing, tmin, n_events, events);
    if (!rc) rc = o_pacing_advance(&pacing, tmin);
    if (rc) goto done;
    tnext_pace = pacing.tnext;
    engine_pace = pacing.level;
    arg_pace = (Real)engine_pace;

    /* openclsim.c:501-502, 1018-1022 */
    engine_time = tmin;
    arg_time = (Real)engine_time;
    istep = 1;
    inext_log = 0;
    tnext_log = tmin;

    const double t_loop0 = o_now();
    while (1) {
        int logging_condition, intermediary_step = 0;
        long ic;

        /* openclsim.c:1054 */
        logging_condition = (engine_time >= tnext_log);

        /* openclsim.c:1057-1063 */
        dt = tmin + (double)istep * default_dt - engine_time;
        d = tmax - engine_time;
        if (d > dt_min && d < dt) { dt = d; intermediary_step = 1; }
        d = tnext_pace - engine_time;
        if (d > dt_min && d < dt) { dt = d; intermediary_step = 1; }
        d = tnext_log - engine_time;
        if (d > dt_min && d < dt) { dt = d; intermediary_step = 1; }
        if (!intermediary_step) istep++;
        arg_dt = (Real)dt;

        /* Diffusion current at time t: openclsim.c:1066-1076 */
        if (diffusion_mode == 3) {
#ifdef ORACLE_REF_KERNEL
            for (ic = 0; ic < (long)n; ic++) {
                cl_set_gid((size_t)ic, 0);
                diff_arb_reset(n, idiff);
            }
            for (ic = 0; ic < (long)n_conn; ic++) {
                cl_set_gid((size_t)ic, 0);
                diff_arb_step(n_conn, (const unsigned long*)conn1,
                              (const unsigned long*)conn2, cg, state, idiff);
            }
#else
            /* openclsim.cl:566-573, then :545-555 in edge order */
            for (i = 0; i < n; i++) idiff[i] = 0;
            for (i = 0; i < n_conn; i++) {
                const size_t i1 = conn1[i], i2 = conn2[i];
                const Real i12 = cg[i] * (state[i1 * N_STATE + I_VM]
                                          - state[i2 * N_STATE + I_VM]);
                idiff[i1] = idiff[i1] + i12;
                idiff[i2] = idiff[i2] + (-i12);
            }
#endif
        } else if (diffusion_mode == 2) {
            #pragma omp parallel for schedule(static)
            for (ic = 0; ic < (long)n; ic++) {
                const size_t ix = (size_t)ic % nx, iy = (size_t)ic / nx;
#ifdef ORACLE_REF_KERNEL
                cl_set_gid(ix, iy);
                diff_hetero(nx, ny, gxf, gyf, state, idiff);
#else
                diff_hetero_cell(ix, iy, nx, ny, gxf, gyf, state, idiff);
#endif
            }
        } else if (diffusion_mode == 1) {
            #pragma omp parallel for schedule(static)
            for (ic = 0; ic < (long)n; ic++) {
                const size_t ix = (size_t)ic % nx, iy = (size_t)ic / nx;
#ifdef ORACLE_REF_KERNEL
                cl_set_gid(ix, iy);
                diff_step(nx, ny, gx, gy, state, idiff);
#else
                diff_step_cell(ix, iy, nx, ny, gx, gy, state, idiff);
#endif
            }
        }

        /* State snapshot at t: openclsim.c:1079-1090 */
        if (logging_condition && logging_states) {
            memcpy(snap_state, state, sizeof(Real) * n * N_STATE);
            if (isnan(snap_state[0])) halt = 1;
        }

        /* Cell kernel: states -> t + dt, intermediates at t: :1093-1096 */
        #pragma omp parallel for schedule(static)
        for (ic = 0; ic < (long)n; ic++) {
#ifdef ORACLE_REF_KERNEL
            cl_set_gid((size_t)ic % nx, (size_t)ic / nx);
            cell_step(nx, ny, arg_time, arg_dt, arg_pace, state, idiff,
                      inter_log, field_data);
            (void)paced;
#else
            Real pace = arg_pace;
            /* openclsim.cl:322-329 with :249-280 folded into a mask */
            if (diffusion_mode != 0) pace = paced[ic] ? arg_pace : 0;
            cell_step((size_t)ic, arg_time, arg_dt, pace, state, idiff,
                      inter_log, field_data);
#endif
        }
        n_steps++;

        /* Log row for time t: openclsim.c:1108-1145 */
        if (logging_condition) {
            if (n_rows >= max_rows) { rc = 1; goto done; }
            for (i = 0; i < n_log; i++) {
                double val = 0;
                switch (log_kind[i]) {
                case LOG_TIME: val = (double)arg_time; break;
                case LOG_PACE: val = (double)arg_pace; break;
                case LOG_IDIFF: val = (double)idiff[log_index[i]]; break;
                case LOG_STATE: val = (double)snap_state[log_index[i]]; break;
                case LOG_INTER: val = (double)inter_log[log_index[i]]; break;
                }
                log_out[n_rows * n_log + i] = val;
            }
            n_rows++;
            inext_log++;
            tnext_log = tmin + (double)inext_log * log_interval;
        }

        /* openclsim.c:1147-1155 */
        engine_time += dt;
        arg_time = (Real)engine_time;
        rc = o_pacing_advance(&pacing, engine_time);
        if (rc) goto done;
        tnext_pace = pacing.tnext;
        engine_pace = pacing.level;
        arg_pace = (Real)engine_pace;

        /* openclsim.c:1162 */
        if (engine_time >= tmax || halt) break;
    }
    g_loop_seconds = o_now() - t_loop0;

    /* Final state: openclsim.c:1185-1190 */
    for (i = 0; i < n * N_STATE; i++) state_io[i] = (double)state[i];

done:
    *n_rows_out = n_rows;
    *n_steps_out = n_steps;
    *halted_out = halt;
    *final_time_out = engine_time;
    free(pacing.ev);
    free(state); free(idiff); free(inter_log); free(field_data);
    free(gxf); free(gyf); free(cg); free(snap_state);
    return rc;
}

/* -------------------------------------------------------------------------
 * Step-wise entry points for coupled runs (oracle/fiber_tissue.py): the
 * fibre-tissue simulation (myokit/_sim/fiber_tissue.c:1001-1155) steps two
 * grids with two models and one pacing system, so its loop lives in Python
 * and calls the pieces of one model's library. Arrays are `Real`.
 * ------------------------------------------------------------------------- */

/* diff_step for every cell: fiber_tissue.c:1031-1032 (openclsim.cl:384-435) */
void oracle_k_diff(size_t nx, size_t ny, double gx_in, double gy_in,
                   const Real* state, Real* idiff)
{
    const Real gx = (Real)gx_in, gy = (Real)gy_in;
    long ic;
    for (ic = 0; ic < (long)(nx * ny); ic++) {
        const size_t ix = (size_t)ic % nx, iy = (size_t)ic / nx;
#ifdef ORACLE_REF_KERNEL
        cl_set_gid(ix, iy);
        diff_step(nx, ny, gx, gy, state, idiff);
#else
        diff_step_cell(ix, iy, nx, ny, gx, gy, (Real*)state, idiff);
#endif
    }
}

/* diff_step_fiber_tissue: fiber_tissue.c:1033 (openclsim.cl:601-628). Called
 * on the FIBRE model's library; nst / ivt describe the tissue's state vector. */
void oracle_k_junction(size_t nfx, size_t nfy, size_t ntx, size_t ctx, size_t cty,
                       size_t nst, int ivt, double gft_in,
                       const Real* state_f, const Real* state_t,
                       Real* idiff_f, Real* idiff_t)
{
    const Real gft = (Real)gft_in;
    size_t cid;
#if defined(ORACLE_REF_KERNEL) && defined(ORACLE_REF_FT)
    for (cid = 0; cid < nfy; cid++) {
        cl_set_gid(cid, 0);
        diff_step_fiber_tissue(nfx, nfy, ntx, ctx, cty, N_STATE, nst, I_VM, ivt, gft,
                               state_f, state_t, idiff_f, idiff_t);
    }
#else
    for (cid = 0; cid < nfy; cid++) {
        const size_t iff = (nfx - 1) + cid * nfx;
        const size_t ift = ctx + (cty + cid) * ntx;
        const Real i = gft * (state_f[iff * N_STATE + I_VM] - state_t[ift * nst + ivt]);
        idiff_f[iff] += i;
        idiff_t[ift] -= i;
    }
#endif
}

/* cell_step for every cell: fiber_tissue.c:1055-1062 */
void oracle_k_cells(size_t nx, size_t ny, double time_in, double dt_in, double pace_in,
                    const unsigned char* paced, Real* state, const Real* idiff,
                    Real* inter_log, const Real* field_data)
{
    const Real arg_time = (Real)time_in, arg_dt = (Real)dt_in, arg_pace = (Real)pace_in;
    long ic;
    for (ic = 0; ic < (long)(nx * ny); ic++) {
#ifdef ORACLE_REF_KERNEL
        cl_set_gid((size_t)ic % nx, (size_t)ic / nx);
        cell_step(nx, ny, arg_time, arg_dt, arg_pace, state, idiff, inter_log, field_data);
        (void)paced;
#else
        cell_step((size_t)ic, arg_time, arg_dt, paced[ic] ? arg_pace : 0, state,
                  (Real*)idiff, inter_log, field_data);
#endif
    }
}

/* The pacing system as an object: fiber_tissue.c:473-481, 1135-1139 */
void* oracle_pacing_new(double t0, int n_events, const double* events, int* rc_out)
{
    OPacing* p = (OPacing*)calloc(1, sizeof(OPacing));
    int rc = o_pacing_init(p, t0, n_events, events);
    if (!rc) rc = o_pacing_advance(p, t0);
    *rc_out = rc;
    return p;
}

int oracle_pacing_advance(void* h, double t, double* level, double* tnext)
{
    OPacing* p = (OPacing*)h;
    int rc = o_pacing_advance(p, t);
    *level = p->level;
    *tnext = p->tnext;
    return rc;
}

void oracle_pacing_state(void* h, double* level, double* tnext)
{
    OPacing* p = (OPacing*)h;
    *level = p->level;
    *tnext = p->tnext;
}

void oracle_pacing_free(void* h)
{
    OPacing* p = (OPacing*)h;
    if (p) free(p->ev);
    free(p);
}
