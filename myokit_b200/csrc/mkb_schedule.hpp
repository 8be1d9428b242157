// mkb_schedule.hpp — the host-side time-step schedule of the CUDA back-end.
//
// Restates the step selection of the reference's sim_step loop
// (myokit/_sim/openclsim.c:1051-1178): the step that ends at the next
// multiple of the default step size, shortened to hit tmax, the next pacing
// event or the next log point when one of them comes first (dt_min = 0,
// openclsim.c:413, so sub-ulp "intermediary" steps do occur); logging decided
// before the step (:1054) with the next log point at tmin + k * log_interval
// (:1134-1136); pacing advanced after the step (:1147-1155); the run ends when
// engine_time >= tmax (:1162), so the final time point is never logged.
//
// Pure host code, no CUDA: the runtime feeds its output into the device
// schedule ring, and mkb_schedule_probe exposes it to CPU-only tests.
#pragma once
#include "mkb_pacing.hpp"

namespace mkb {

struct StepInfo {
    double time;        // engine_time at the start of the step
    double dt;          // step size actually taken
    double pace;        // pacing level during the step
    bool logging;       // a log row is written for `time`
};

class StepScheduler {
public:
    // Returns a PacingStatus (openclsim.c:488-496: create, populate, advance to tmin)
    int init(double tmin, double tmax, double default_dt, double log_interval,
             int n_events, const double* events) {
        tmin_ = tmin;
        tmax_ = tmax;
        default_dt_ = default_dt;
        log_interval_ = log_interval;
        int rc = pacing_.init(tmin, n_events, events);
        if (rc == 0) rc = pacing_.advance(tmin);
        if (rc) return rc;
        tnext_pace_ = pacing_.next_time();
        engine_pace_ = pacing_.level();
        engine_time_ = tmin;            // openclsim.c:501
        istep_ = 1;                     // openclsim.c:1018
        inext_log_ = 0;
        tnext_log_ = tmin;              // openclsim.c:1021-1022
        finished_ = !(tmax > tmin);
        return PACING_OK;
    }

    // Computes the next step and advances time and pacing past it.
    int next(StepInfo* out) {
        const double dt_min = 0;        // openclsim.c:413
        out->logging = (engine_time_ >= tnext_log_);                    // :1054
        bool intermediary = false;                                      // :1057-1063
        double dt = tmin_ + (double)istep_ * default_dt_ - engine_time_;
        double d = tmax_ - engine_time_;
        if (d > dt_min && d < dt) { dt = d; intermediary = true; }
        d = tnext_pace_ - engine_time_;
        if (d > dt_min && d < dt) { dt = d; intermediary = true; }
        d = tnext_log_ - engine_time_;
        if (d > dt_min && d < dt) { dt = d; intermediary = true; }
        if (!intermediary) istep_++;
        out->time = engine_time_;
        out->dt = dt;
        out->pace = engine_pace_;
        if (out->logging) {                                             // :1134-1136
            inext_log_++;
            tnext_log_ = tmin_ + (double)inext_log_ * log_interval_;
        }
        engine_time_ += dt;                                             // :1147-1155
        int rc = pacing_.advance(engine_time_);
        if (rc) return rc;
        tnext_pace_ = pacing_.next_time();
        engine_pace_ = pacing_.level();
        if (engine_time_ >= tmax_) finished_ = true;                    // :1162
        return PACING_OK;
    }

    bool finished() const { return finished_; }
    void finish() { finished_ = true; }
    double time() const { return engine_time_; }

private:
    double tmin_ = 0, tmax_ = 0, default_dt_ = 0, log_interval_ = 1;
    double engine_time_ = 0, engine_pace_ = 0, tnext_pace_ = 0, tnext_log_ = 0;
    unsigned long long istep_ = 1, inext_log_ = 0;
    bool finished_ = true;
    EventPacing pacing_;
};

}  // namespace mkb
