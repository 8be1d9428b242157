// mkb_pacing.hpp — host-side event pacing for the CUDA back-end.
//
// Same observable behaviour as the reference's ESys_* (myokit/_sim/pacing.h):
// tolerant comparisons (:74-80), queue insertion incl. the simultaneous-event
// error (:192-218), advance (:483-548), level / next time (:560-600). Events
// arrive as plain doubles (level, start, duration, period, multiplier) instead
// of Python objects (ESys_Populate, :340-470), and the queue is an index-linked
// vector instead of malloc'ed pointers.
#pragma once
#include <cfloat>
#include <cmath>
#include <vector>

namespace mkb {

enum PacingStatus {
    PACING_OK = 0,
    PACING_NEGATIVE_PERIOD = -23,        // ESys_POPULATE_NEGATIVE_PERIOD
    PACING_NON_ZERO_MULTIPLIER = -24,    // ESys_POPULATE_NON_ZERO_MULTIPLIER
    PACING_NEGATIVE_MULTIPLIER = -25,    // ESys_POPULATE_NEGATIVE_MULTIPLIER
    PACING_NEGATIVE_TIME_INCREMENT = -40,
    PACING_SIMULTANEOUS_EVENT = -50,
};

class EventPacing {
public:
    int init(double t0, int n, const double* events) {
        ev_.clear();
        head_ = fire_ = -1;
        time_ = tnext_ = tdown_ = t0;
        level_ = 0;
        for (int i = 0; i < n; i++) {
            Event e;
            e.level = events[5 * i];
            e.start = events[5 * i + 1];
            e.duration = events[5 * i + 2];
            e.period = events[5 * i + 3];
            e.multiplier = events[5 * i + 4];
            e.next = -1;
            if (e.period == 0 && e.multiplier != 0) return PACING_NON_ZERO_MULTIPLIER;
            if (e.period < 0) return PACING_NEGATIVE_PERIOD;
            if (e.multiplier < 0) return PACING_NEGATIVE_MULTIPLIER;
            ev_.push_back(e);
        }
        if (n > 0) {
            head_ = 0;
            for (int i = 1; i < n; i++) {
                bool clash = false;
                head_ = schedule(head_, i, &clash);
                if (clash) return PACING_SIMULTANEOUS_EVENT;
            }
        }
        return PACING_OK;
    }

    int advance(double new_time) {
        if (new_time < time_) return PACING_NEGATIVE_TIME_INCREMENT;
        time_ = new_time;
        while (geq(time_, tnext_)) {
            // Active event finished
            if (fire_ >= 0 && geq(tnext_, tdown_)) {
                fire_ = -1;
                level_ = 0;
            }
            // New event starting
            if (head_ >= 0 && geq(tnext_, ev_[head_].start)) {
                fire_ = head_;
                Event& f = ev_[fire_];
                head_ = f.next;
                tdown_ = f.start + f.duration;
                level_ = f.level;
                if (f.period > 0) {
                    if (f.multiplier != 1) {
                        if (f.multiplier > 1) f.multiplier--;
                        f.start += f.period;
                        bool clash = false;
                        head_ = schedule(head_, fire_, &clash);
                        if (clash) return PACING_SIMULTANEOUS_EVENT;
                    } else {
                        f.period = 0;
                    }
                }
                // Snap a computed end onto an indistinguishable event start
                if (head_ >= 0 && eq(ev_[head_].start, tdown_)) {
                    tdown_ = ev_[head_].start;
                }
            }
            tnext_ = HUGE_VAL;
            if (fire_ >= 0 && tnext_ > tdown_) tnext_ = tdown_;
            if (head_ >= 0 && tnext_ > ev_[head_].start) tnext_ = ev_[head_].start;
        }
        return PACING_OK;
    }

    double level() const { return level_; }
    double next_time() const { return tnext_; }

private:
    struct Event {
        double level, start, duration, period, multiplier;
        int next;
    };

    static bool eq(double a, double b) {
        if (a == b) return true;
        double s = std::fabs(a) > std::fabs(b) ? std::fabs(a) : std::fabs(b);
        return std::fabs(a - b) / s < DBL_EPSILON;
    }
    static bool geq(double a, double b) { return a >= b || eq(a, b); }

    int schedule(int head, int add, bool* clash) {
        ev_[add].next = -1;
        if (head < 0) return add;
        if (ev_[add].start < ev_[head].start) {
            ev_[add].next = head;
            return add;
        }
        int e = head;
        while (ev_[e].next >= 0 && ev_[add].start >= ev_[ev_[e].next].start) {
            e = ev_[e].next;
        }
        if (ev_[add].start == ev_[e].start) *clash = true;
        ev_[add].next = ev_[e].next;
        ev_[e].next = add;
        return head;
    }

    std::vector<Event> ev_;
    int head_ = -1, fire_ = -1;
    double time_ = 0, tnext_ = 0, tdown_ = 0, level_ = 0;
};

}  // namespace mkb
