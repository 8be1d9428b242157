"""
Host execution of generated kernels — TEST INFRASTRUCTURE ONLY.

``run_on_host(sim, duration, ...)`` takes a configured ``SimulationCUDA``
(never run, no GPU needed), compiles the CUDA source its next run would JIT
(``sim.kernel_source()``) as host C++ behind ``mkb_cuda_shim.h`` and steps it
with the library's own schedule (``mkb_schedule_probe``). The product never
imports this package.
"""
import ctypes
import hashlib
import os
import subprocess

import numpy as np

from myokit_b200 import capi, kernelgen

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_BUILD = os.path.join(_HERE, '_build')
_CSRC = os.path.join(_ROOT, 'myokit_b200', 'csrc')


def _compile(code, contract, kernel_name='mkb_cell_step', second=None):
    os.makedirs(_BUILD, exist_ok=True)
    with open(os.path.join(_HERE, 'runner.cpp'), 'rb') as f:
        runner = f.read()
    with open(os.path.join(_HERE, 'mkb_cuda_shim.h'), 'rb') as f:
        shim = f.read()
    key = hashlib.sha1(code.encode('utf-8') + runner + shim
                       + (b'c' if contract else b'n')
                       + kernel_name.encode('ascii')
                       + (second or '').encode('ascii')).hexdigest()[:20]
    so = os.path.join(_BUILD, 'k_%s.so' % key)
    if not os.path.isfile(so):
        cu = os.path.join(_BUILD, 'k_%s.cu.h' % key)
        with open(cu, 'w') as f:
            f.write(code)
        tmp = so + '.tmp%d' % os.getpid()
        cmd = ['g++', '-O1', '-std=c++17', '-fPIC', '-shared', '-mfma',
               '-ffp-contract=' + ('fast' if contract else 'off'),
               '-Wno-unknown-pragmas', '-Wno-unused-variable',
               '-DMKB_KERNEL_FILE="%s"' % cu, '-DMKB_KERNEL_FN=' + kernel_name,
               '-I' + _CSRC, '-I' + _HERE,
               os.path.join(_HERE, 'runner.cpp'), '-o', tmp]
        if second:
            cmd.insert(-3, '-DMKB_KERNEL_FN2=' + second)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('host build of the generated kernel failed:\n'
                               + r.stderr[-4000:])
        os.replace(tmp, so)
    return ctypes.CDLL(so)


def schedule(sim, duration, log_interval):
    """The library's step list for a run: (times, dts, paces, logging)."""
    lib = capi.library()
    events, n_events = sim._events()
    tmin = sim.time()
    cap = int(duration / sim.step_size()) * 2 + 64
    times = np.zeros(cap)
    dts = np.zeros(cap)
    paces = np.zeros(cap)
    logging = np.zeros(cap, dtype=np.uint8)
    n = ctypes.c_uint64(0)
    rc = lib.mkb_schedule_probe(
        ctypes.c_double(tmin), ctypes.c_double(tmin + duration),
        ctypes.c_double(sim.step_size()), ctypes.c_double(log_interval),
        ctypes.c_int(n_events), events.ctypes.data_as(ctypes.c_void_p),
        ctypes.c_uint64(cap), times.ctypes.data_as(ctypes.c_void_p),
        dts.ctypes.data_as(ctypes.c_void_p), paces.ctypes.data_as(ctypes.c_void_p),
        logging.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n))
    assert rc == 0, rc
    k = int(n.value)
    return times[:k], dts[:k], paces[:k], logging[:k]


def run_on_host(sim, duration, log_interval=1.0, inter_log=(), contract=None,
                reverse=False, stream_blocks=3):
    """
    Runs ``duration`` on the host. Returns a dict: ``time`` (nt,), ``V``
    (nt, ncells) — V(t) at the logged steps —, ``idiff`` (nt, ncells),
    ``inter`` (nt, n_inter, ncells), ``state`` (ncells, n_state), ``steps``.
    ``contract``: let g++ contract a*b+c like nvcc's --fmad (default: what
    the kernel source was generated for).
    """
    inter_vars = [sim._model.get(q) for q in inter_log]
    src = sim.kernel_source(inter_vars)
    if contract is None:
        contract = '--fmad=true' in src.options
    lib = _compile(src.code, contract, src.kernel_name,
                   'mkb_gate_step' if getattr(src, 'gate_kernel', False) else None)
    lib.shim_set_thread_order(1 if reverse else 0)
    lib.shim_set_persistent(1 if getattr(src, 'persistent', False) else 0)
    # streaming kernels: a persistent grid of a few blocks walks all tiles
    lib.shim_set_stream_blocks(stream_blocks if (src.kernel_flags & (2 | 16)) else 0)
    nx, ny = sim._nx, sim._ny
    if getattr(src, 'persistent', False):
        assert nx <= src.block[0] and ny <= src.block[1]
    n = nx * ny
    times, dts, paces, logging = schedule(sim, duration, log_interval)
    rows = int(logging.sum())

    state = np.ascontiguousarray(sim._state, dtype=np.float64).copy()
    assert state.size == n * src.n_state
    fields = [np.asarray(f, dtype=np.float64).ravel() for f in sim._fields.values()]
    field_aos = (np.ascontiguousarray(np.vstack(fields).T).ravel()
                 if fields else np.zeros(1))
    mode = src.diffusion_mode
    gxf = gyf = None
    if mode == kernelgen.DIFF_FIELD:
        gxf = np.ascontiguousarray(sim._gx_field, dtype=np.float64).ravel()
        if sim._gy_field is not None:
            gyf = np.ascontiguousarray(sim._gy_field, dtype=np.float64).ravel()
    ci = cj = cg = None
    n_conn = 0
    if mode == kernelgen.DIFF_CONNECTIONS:
        a, b, g = sim._connections
        ci = np.ascontiguousarray(a, dtype=np.uint64)
        cj = np.ascontiguousarray(b, dtype=np.uint64)
        cg = np.ascontiguousarray(g, dtype=np.float64)
        n_conn = len(ci)
    px0 = px1 = py0 = py1 = 0
    mask = None
    if sim._diffusion_enabled:
        if type(sim._paced_cells) == tuple:
            pnx, pny, px, py = sim._paced_cells
            px0, px1, py0, py1 = px, px + pnx, py, py + pny
        else:
            mask = np.zeros(n, dtype=np.uint8)
            mask[np.array(sim._paced_cells, dtype=np.int64)] = 1

    def ptr(a):
        return None if a is None else a.ctypes.data_as(ctypes.c_void_p)

    log_v = np.zeros((rows, n))
    log_idiff = np.zeros((rows, n))
    log_inter = np.zeros((rows, max(src.n_inter, 1), n))
    lib.shim_run.restype = ctypes.c_int
    rc = lib.shim_run(
        ctypes.c_int(nx), ctypes.c_int(ny), ctypes.c_int(src.n_state),
        ctypes.c_int(src.i_vm), ctypes.c_int(src.n_inter),
        ctypes.c_int(src.n_field), ctypes.c_int(mode),
        ctypes.c_double(sim._gx or 0), ctypes.c_double(sim._gy or 0),
        ptr(gxf), ptr(gyf),
        ctypes.c_longlong(px0), ctypes.c_longlong(px1),
        ctypes.c_longlong(py0), ctypes.c_longlong(py1), ptr(mask),
        ctypes.c_ulonglong(n_conn), ptr(ci), ptr(cj), ptr(cg),
        ctypes.c_int(len(times)), ptr(times), ptr(dts), ptr(paces),
        ptr(logging), ptr(state), ptr(field_aos), ptr(log_v), ptr(log_idiff),
        ptr(log_inter),
        ctypes.c_int(src.block[0]), ctypes.c_int(src.block[1]),
        ctypes.c_int(src.cells_per_thread), ctypes.c_int(src.rows_per_thread))
    assert rc == 0
    return {
        'time': times[logging.astype(bool)], 'V': log_v, 'idiff': log_idiff,
        'inter': log_inter[:, :src.n_inter], 'steps': len(times),
        'state': state.reshape(n, src.n_state),
        'real_size': lib.shim_real_size(),
    }


def run_slabs_on_host(make, n_slabs, duration, log_interval=1.0, options=None,
                      reverse=False):
    """
    Row-slab kernels on the host: ``make(comm=None)`` builds the simulation;
    the slab variant of its kernel is what a rank of an ``n_slabs``-GPU run
    compiles. All slabs run in this process, step by step (see runner.cpp).
    Returns the same dict as :func:`run_on_host` for the whole grid plus
    ``halo_error``.
    """
    from myokit_b200 import multigpu
    # one host library serves every slab: the kernel reads its plane stride at
    # run time (on a GPU each rank compiles for its own)
    options = dict(options or {}, plane_stride=False)
    whole = make(None)
    whole.set_kernel_options(**options)
    box = {}

    def work(comm):
        s = make(comm)
        s.set_kernel_options(**options)
        if comm.rank == 0:
            box['src'] = s.kernel_source()
    multigpu.run_threads(n_slabs, work)
    src = box['src']
    assert 'peer_lo_halo_hi' in src.code      # the slab variant
    contract = '--fmad=true' in src.options
    lib = _compile(src.code, contract, src.kernel_name,
                   'mkb_gate_step' if getattr(src, 'gate_kernel', False) else None)
    lib.shim_set_thread_order(1 if reverse else 0)
    nx, ny = whole._nx, whole._ny
    n = nx * ny
    rows = multigpu.slab_rows(ny, n_slabs)
    row0 = np.array([r[0] for r in rows], dtype=np.int32)
    row1 = np.array([r[1] for r in rows], dtype=np.int32)
    times, dts, paces, logging = schedule(whole, duration, log_interval)
    nrows = int(logging.sum())
    state = np.ascontiguousarray(whole._state, dtype=np.float64).copy()
    fields = [np.asarray(f, dtype=np.float64).ravel() for f in whole._fields.values()]
    field_aos = (np.ascontiguousarray(np.vstack(fields).T).ravel()
                 if fields else np.zeros(1))
    gxf = gyf = None
    if src.diffusion_mode == kernelgen.DIFF_FIELD:
        gxf = np.ascontiguousarray(whole._gx_field, dtype=np.float64).ravel()
        gyf = np.ascontiguousarray(whole._gy_field, dtype=np.float64).ravel()
    px0 = px1 = py0 = py1 = 0
    mask = None
    if type(whole._paced_cells) == tuple:
        pnx, pny, px, py = whole._paced_cells
        px0, px1, py0, py1 = px, px + pnx, py, py + pny
    else:
        mask = np.zeros(n, dtype=np.uint8)
        mask[np.array(whole._paced_cells, dtype=np.int64)] = 1

    def ptr(a):
        return None if a is None else a.ctypes.data_as(ctypes.c_void_p)

    log_v = np.zeros((nrows, n))
    log_idiff = np.zeros((nrows, n))
    err = ctypes.c_uint(0)
    lib.shim_run_slabs.restype = ctypes.c_int
    rc = lib.shim_run_slabs(
        ctypes.c_int(nx), ctypes.c_int(ny), ctypes.c_int(n_slabs),
        ptr(row0), ptr(row1),
        ctypes.c_int(src.n_state), ctypes.c_int(src.i_vm),
        ctypes.c_int(src.n_inter), ctypes.c_int(src.n_field),
        ctypes.c_int(src.diffusion_mode),
        ctypes.c_double(whole._gx or 0), ctypes.c_double(whole._gy or 0),
        ptr(gxf), ptr(gyf),
        ctypes.c_longlong(px0), ctypes.c_longlong(px1),
        ctypes.c_longlong(py0), ctypes.c_longlong(py1), ptr(mask),
        ctypes.c_int(len(times)), ptr(times), ptr(dts), ptr(paces),
        ptr(logging), ptr(state), ptr(field_aos), ptr(log_v), ptr(log_idiff),
        ctypes.c_int(src.block[0]), ctypes.c_int(src.block[1]),
        ctypes.byref(err))
    assert rc == 0
    return {
        'time': times[logging.astype(bool)], 'V': log_v, 'idiff': log_idiff,
        'steps': len(times), 'state': state.reshape(n, src.n_state),
        'halo_error': int(err.value),
    }


def run_parts_on_host(make, n_parts, duration, log_interval=1.0, options=None,
                      reverse=False):
    """
    Partitioned ``set_connections`` graphs on the host: ``make(comm)`` builds
    the simulation on every rank of an ``n_parts`` job; the partitions (the
    product's own ``_partition_graph``) are stepped in lockstep in this
    process, the exported V pushed into the other partitions' ghost planes
    between steps like ``k_push_ghosts`` does. Returns ``time``, ``V``
    (nt, ncells, global cell order), ``idiff``, ``state``, ``halo_error``.
    """
    from myokit_b200 import multigpu
    options = dict(options or {}, plane_stride=False)     # (as in run_slabs_on_host)
    ranks = [None] * n_parts

    def work(comm):
        s = make(comm)
        s.set_kernel_options(**options)
        src = s.kernel_source()
        ci, cj, cg, ghost_ids = s._partition_graph()
        ranks[comm.rank] = dict(
            sim=s, src=src, x0=s._sx0, n=s._snx, ci=ci, cj=cj, cg=cg,
            ghost_ids=ghost_ids, state=np.array(s._state, dtype=np.float64),
            fields=[np.asarray(s._local_slice(f), dtype=np.float64).ravel()
                    for f in s._fields.values()])
    multigpu.run_threads(n_parts, work)
    src = ranks[0]['src']
    assert 'g.ghost' in src.code            # the partitioned variant
    lib = _compile(src.code, '--fmad=true' in src.options)
    lib.shim_set_thread_order(1 if reverse else 0)
    lib.shim_part_create.restype = ctypes.c_void_p
    lib.shim_part_ghost.restype = ctypes.c_void_p
    lib.shim_part_flags.restype = ctypes.c_void_p
    lib.shim_part_error.restype = ctypes.c_uint
    real = np.float32 if lib.shim_real_size() == 4 else np.float64
    whole = ranks[0]['sim']
    ntot = whole._nx
    times, dts, paces, logging = schedule(whole, duration, log_interval)

    def ptr(a):
        return None if a is None else a.ctypes.data_as(ctypes.c_void_p)

    keep = []
    for r, d in enumerate(ranks):
        s = d['sim']
        # ranks that own ghost cells of this partition raise a flag here
        imp = [q for q, e in enumerate(ranks) if q != r and len(d['ghost_ids'])
               and np.any((d['ghost_ids'] >= e['x0']) & (d['ghost_ids'] < e['x0'] + e['n']))]
        imp = np.array(imp, dtype=np.uint32)
        px0 = px1 = 0
        mask = None
        if type(s._paced_cells) == tuple:
            pnx, pny, px, py = s._paced_cells
            px0, px1 = px - d['x0'], px - d['x0'] + pnx
        else:
            mask = np.zeros(d['n'], dtype=np.uint8)
            pc = np.array(s._paced_cells, dtype=np.int64)
            pc = pc[(pc >= d['x0']) & (pc < d['x0'] + d['n'])] - d['x0']
            mask[pc] = 1
        field_aos = (np.ascontiguousarray(np.vstack(d['fields']).T).ravel()
                     if d['fields'] else np.zeros(1))
        ci = np.ascontiguousarray(d['ci'], dtype=np.uint64)
        cj = np.ascontiguousarray(d['cj'], dtype=np.uint64)
        cg = np.ascontiguousarray(d['cg'], dtype=np.float64)
        keep += [imp, mask, field_aos, ci, cj, cg]
        h = lib.shim_part_create(
            ctypes.c_ulonglong(d['n']), ctypes.c_int(src.n_state),
            ctypes.c_int(src.i_vm), ctypes.c_int(src.n_inter),
            ctypes.c_int(src.n_field), ctypes.c_ulonglong(len(ci)),
            ptr(ci), ptr(cj), ptr(cg),
            ctypes.c_ulonglong(len(d['ghost_ids'])), ctypes.c_int(n_parts),
            ctypes.c_int(len(imp)), ptr(imp) if len(imp) else None,
            ctypes.c_longlong(px0), ctypes.c_longlong(px1), ptr(mask),
            ptr(d['state']), ptr(field_aos), ctypes.c_int(src.block[0]))
        d['h'] = ctypes.c_void_p(h)
        ng = max(len(d['ghost_ids']), 1)
        d['ghost'] = np.ctypeslib.as_array(
            ctypes.cast(lib.shim_part_ghost(d['h']), ctypes.POINTER(
                ctypes.c_float if real is np.float32 else ctypes.c_double)),
            shape=(3, ng))
        d['flags'] = np.ctypeslib.as_array(
            ctypes.cast(lib.shim_part_flags(d['h']), ctypes.POINTER(ctypes.c_uint)),
            shape=(n_parts,))
        d['v'] = np.zeros(d['n'])
    # export lists: cells of partition r that are ghost cells of partition q
    for r, d in enumerate(ranks):
        d['exports'] = []
        for q, e in enumerate(ranks):
            if q == r:
                continue
            ids = e['ghost_ids']
            sel = np.nonzero((ids >= d['x0']) & (ids < d['x0'] + d['n']))[0]
            if len(sel):
                d['exports'].append((q, ids[sel] - d['x0'], sel))

    def push(r, step):
        d = ranks[r]
        lib.shim_part_v(d['h'], ptr(d['v']))
        for q, src_cells, slots in d['exports']:
            ranks[q]['ghost'][step % 3, slots] = d['v'][src_cells]
        for q, e in enumerate(ranks):
            if q != r and any(x[0] == q for x in d['exports']):
                e['flags'][r] = step

    for r in range(n_parts):
        push(r, 1)                          # seed: V(t0) for step 1
    nrows = int(logging.sum())
    log_v = np.zeros((nrows, ntot))
    log_idiff = np.zeros((nrows, ntot))
    row = 0
    tmp = None
    for k in range(len(times)):
        step = k + 1
        for r, d in enumerate(ranks):
            if logging[k]:
                lib.shim_part_v(d['h'], ptr(d['v']))
                log_v[row, d['x0']:d['x0'] + d['n']] = d['v']
            lib.shim_part_step(d['h'], ctypes.c_double(times[k]),
                               ctypes.c_double(dts[k]), ctypes.c_double(paces[k]),
                               ctypes.c_int(int(logging[k])), ctypes.c_uint(step))
            if logging[k]:
                tmp = np.zeros(d['n'])
                lib.shim_part_idiff(d['h'], ptr(tmp))
                log_idiff[row, d['x0']:d['x0'] + d['n']] = tmp
            push(r, step + 1)
        if logging[k]:
            row += 1
    err = 0
    state = np.zeros((ntot, src.n_state))
    for d in ranks:
        err |= lib.shim_part_error(d['h'])
        out = np.zeros(d['n'] * src.n_state)
        d.pop('ghost')
        d.pop('flags')
        lib.shim_part_finish(d['h'], ptr(out))
        state[d['x0']:d['x0'] + d['n']] = out.reshape(d['n'], src.n_state)
    del keep
    return {'time': times[logging.astype(bool)], 'V': log_v, 'idiff': log_idiff,
            'state': state, 'halo_error': int(err), 'steps': len(times)}


def run_pair_on_host(fiber, tissue, g_fiber_tissue, cty, duration,
                     log_interval=1.0, reverse=False):
    """
    Two homogeneous 2-d grids stepped in lockstep on the host, coupled through
    ``MkbGridArgs::junction_*``: ``fiber`` and ``tissue`` are configured
    ``SimulationCUDA`` objects whose kernels were generated with
    ``junction='fiber'`` / ``'tissue'``. The last fibre column is tied to
    tissue cells ``(0, cty + y)``. Returns ``time`` and, per part, ``V``,
    ``idiff`` (nt, ny, nx) and ``state``.
    """
    parts = []
    for sim in (fiber, tissue):
        src = sim.kernel_source()
        assert 'g.junction_v0' in src.code
        lib = _compile(src.code, '--fmad=true' in src.options)
        lib.shim_set_thread_order(1 if reverse else 0)
        lib.shim_grid_create.restype = ctypes.c_void_p
        lib.shim_grid_vplane.restype = ctypes.c_void_p
        nx, ny = sim._nx, sim._ny
        pnx, pny, px, py = sim._paced_cells
        state = np.ascontiguousarray(sim._state, dtype=np.float64)
        h = ctypes.c_void_p(lib.shim_grid_create(
            ctypes.c_int(nx), ctypes.c_int(ny), ctypes.c_int(src.n_state),
            ctypes.c_int(src.i_vm), ctypes.c_int(src.n_inter), ctypes.c_int(0),
            ctypes.c_double(sim._gx), ctypes.c_double(sim._gy),
            ctypes.c_longlong(px), ctypes.c_longlong(px + pnx),
            ctypes.c_longlong(py), ctypes.c_longlong(py + pny), None,
            state.ctypes.data_as(ctypes.c_void_p),
            np.zeros(1).ctypes.data_as(ctypes.c_void_p),
            ctypes.c_int(src.block[0]), ctypes.c_int(src.block[1])))
        parts.append(dict(sim=sim, src=src, lib=lib, h=h, nx=nx, ny=ny,
                          n=nx * ny))
    f, t = parts
    for me, other, jx, jy0, joff, jstride in (
            (f, t, f['nx'] - 1, 0, 0 + cty * t['nx'], t['nx']),
            (t, f, 0, cty, f['nx'] - 1, f['nx'])):
        v0 = other['lib'].shim_grid_vplane(other['h'], 0)
        v1 = other['lib'].shim_grid_vplane(other['h'], 1)
        me['lib'].shim_grid_set_junction(
            me['h'], ctypes.c_void_p(v0), ctypes.c_void_p(v1),
            ctypes.c_double(g_fiber_tissue), ctypes.c_ulonglong(jx),
            ctypes.c_ulonglong(jy0), ctypes.c_ulonglong(f['ny']),
            ctypes.c_ulonglong(joff), ctypes.c_ulonglong(jstride))
    times, dts, paces, logging = schedule(fiber, duration, log_interval)
    rows = {0: [], 1: []}
    for k in range(len(times)):
        for i, p in enumerate(parts):
            if logging[k]:
                v = np.zeros(p['n'])
                p['lib'].shim_grid_v(p['h'], v.ctypes.data_as(ctypes.c_void_p))
                rows[i].append([v])
        # both kernels of a step read V(t) of both grids and write V(t + dt)
        for i, p in enumerate(parts):
            p['lib'].shim_grid_step(
                p['h'], ctypes.c_double(times[k]), ctypes.c_double(dts[k]),
                ctypes.c_double(paces[k]), ctypes.c_int(int(logging[k])),
                ctypes.c_uint(k + 1))
            if logging[k]:
                d = np.zeros(p['n'])
                p['lib'].shim_grid_idiff(p['h'], d.ctypes.data_as(ctypes.c_void_p))
                rows[i][-1].append(d)
    out = {'time': times[logging.astype(bool)], 'steps': len(times)}
    for i, name in enumerate(('fiber', 'tissue')):
        p = parts[i]
        nt = len(rows[i])
        state = np.zeros(p['n'] * p['src'].n_state)
        p['lib'].shim_grid_finish(p['h'], state.ctypes.data_as(ctypes.c_void_p))
        out[name] = {
            'V': np.array([r[0] for r in rows[i]]).reshape(nt, p['ny'], p['nx']),
            'idiff': np.array([r[1] for r in rows[i]]).reshape(nt, p['ny'], p['nx']),
            'state': state,
        }
    return out
