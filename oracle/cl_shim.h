/*
 * oracle/cl_shim.h — lets the OpenCL C text that the REFERENCE renders from
 * myokit/_sim/openclsim.cl compile as ISO C (oracle/_ref builds only; test
 * infrastructure). Work-item ids become thread-local loop indices set by the
 * driver; address-space qualifiers vanish; the compare-exchange used by the
 * reference's AtomicAdd (openclsim.cl:499-524) is serial.
 */
#ifndef ORACLE_CL_SHIM_H
#define ORACLE_CL_SHIM_H
#include <stddef.h>
#include <tgmath.h>     /* OpenCL's exp/log/pow/... are overloaded on float/double */

#define __kernel
#define __global
#define __private
#define __constant const

static _Thread_local size_t cl_gid_[2];
#define get_global_id(i) (cl_gid_[(i)])
static inline void cl_set_gid(size_t x, size_t y)
{
    cl_gid_[0] = x;
    cl_gid_[1] = y;
}

static inline unsigned long atom_cmpxchg(
    volatile unsigned long* p, unsigned long cmp, unsigned long val)
{
    unsigned long old = *p;
    if (old == cmp) *p = val;
    return old;
}
static inline unsigned int atomic_cmpxchg(
    volatile unsigned int* p, unsigned int cmp, unsigned int val)
{
    unsigned int old = *p;
    if (old == cmp) *p = val;
    return old;
}

/* `inline Real f(...)` in OpenCL C has internal linkage semantics here */
#define inline static inline
#endif
