"""
The C ABI driven directly (ctypes, no SimulationCUDA in the loop), the way a
binding for the reference would use it: JIT, init from plain
arrays — here with single-precision host buffers — step, read the log matrix
and the state, re-arm, clean. Checked against the oracle.
"""
import ctypes

import numpy as np
import pytest

import myokit_b200
from myokit_b200 import capi
import myokit

from oracle.oracle import OracleSimulation

pytestmark = pytest.mark.gpu


def test_c_abi_end_to_end_fp32_host_buffers():
    lib = capi.library()
    m, _, _ = myokit.load('example')
    p = myokit.pacing.blocktrain(duration=2, offset=1, period=1000)
    nx, ny = 24, 10
    n = nx * ny

    # The generator is the only Python piece: source text in, cubin out
    gen = myokit_b200.SimulationCUDA(m, p, ncells=(nx, ny),
                                     precision=myokit.SINGLE_PRECISION)
    gen.set_conductance(8, 6)
    gen.set_paced_cells(3, ny, 0, 0)
    src = gen.kernel_source()
    cubin, log = capi.jit_compile(src.code, src.options)
    n_state = src.n_state

    state = np.tile(np.array(m.initial_values(True), dtype=np.float32), n)
    events = np.array([1.0, 1.0, 2.0, 1000.0, 0.0])
    # log: time, pace, V of every cell (field entry), m-gate of cell (5, 3)
    kinds = np.array([capi.LOG_TIME, capi.LOG_PACE, capi.LOG_STATE_FIELD,
                      capi.LOG_STATE], dtype=np.int32)
    index = np.array([0, 0, 0, (5 + 3 * nx) * n_state + 1], dtype=np.uint64)

    cfg = capi.SimConfig()
    cfg.abi_version = capi.MKB_ABI_VERSION
    cfg.device = 0
    cfg.precision = 32
    cfg.host_precision = 32
    buf = ctypes.create_string_buffer(cubin, len(cubin))
    cfg.cubin = ctypes.cast(buf, ctypes.c_void_p)
    cfg.cubin_size = len(cubin)
    cfg.kernel_name = src.kernel_name.encode()
    cfg.kernel_flags = src.kernel_flags      # (what kind of kernel the generator made)
    cfg.stream_box_w, cfg.stream_box_h = src.stream_box
    cfg.kernel_stride = src.plane_stride     # (0, or the stride it was compiled for)
    cfg.kernel_smem_bytes = src.smem_bytes   # (staged kernels)
    cfg.block_x, cfg.block_y = src.block
    cfg.cells_per_thread = src.cells_per_thread
    cfg.rows_per_thread = src.rows_per_thread
    cfg.n_state, cfg.i_vm = n_state, src.i_vm
    cfg.n_inter, cfg.n_field = 0, 0
    cfg.nx, cfg.ny = nx, ny
    cfg.diffusion_mode = src.diffusion_mode
    cfg.gx, cfg.gy = 8.0, 6.0
    cfg.pace_rect = 1
    cfg.pace_nx, cfg.pace_ny, cfg.pace_x, cfg.pace_y = 3, ny, 0, 0
    cfg.n_events = 1
    cfg.events = events.ctypes.data
    cfg.tmin, cfg.tmax, cfg.dt, cfg.log_interval = 0.0, 6.0, 0.005, 0.5
    cfg.state_in = state.ctypes.data
    cfg.n_log = len(kinds)
    cfg.log_kind = kinds.ctypes.data
    cfg.log_index = index.ctypes.data
    cfg.ny_global = ny
    cfg.use_graphs = 1
    cfg.steps_per_call = 1000

    sim = ctypes.c_void_p()
    capi.check(lib.mkb_sim_init(ctypes.byref(cfg), ctypes.byref(sim)))
    try:
        t = ctypes.c_double(0)
        halted = ctypes.c_int(0)
        calls = 0
        while True:
            rc = capi.check(lib.mkb_sim_step(sim, ctypes.byref(t),
                                             ctypes.byref(halted)))
            calls += 1
            if rc == 0:
                break
        assert t.value >= 6.0 and not halted.value and calls == 2

        def rows_matrix():
            data, rows = ctypes.c_void_p(), ctypes.c_uint64()
            cols, stride = ctypes.c_uint64(), ctypes.c_uint64()
            capi.check(lib.mkb_sim_log_view(
                sim, ctypes.byref(data), ctypes.byref(rows),
                ctypes.byref(cols), ctypes.byref(stride)))
            assert cols.value == 2 + n + 1
            raw = (ctypes.c_float * (rows.value * stride.value)).from_address(
                data.value)
            return np.array(raw, dtype=np.float32).reshape(
                rows.value, stride.value)[:, :cols.value]
        mat = rows_matrix()
        assert mat.shape[0] == 12

        out = np.empty(n * n_state, dtype=np.float32)
        capi.check(lib.mkb_sim_get_state(sim, out.ctypes.data))
        launches, steps = ctypes.c_uint64(), ctypes.c_uint64()
        capi.check(lib.mkb_sim_counters(sim, ctypes.byref(launches),
                                        ctypes.byref(steps)))
        assert steps.value == 1200 and launches.value >= 1200

        # oracle on the same inputs
        o = OracleSimulation(m, p, ncells=(nx, ny),
                             precision=myokit.SINGLE_PRECISION)
        o.set_conductance(8, 6)
        o.set_paced_cells(3, ny, 0, 0)
        ol, ostate = o.run(6, log=['engine.time', 'engine.pace', 'membrane.V',
                                   '5.3.ina.m'], log_interval=0.5)
        assert np.array_equal(mat[:, 0].astype(np.float64), ol['engine.time'])
        assert np.array_equal(mat[:, 1].astype(np.float64), ol['engine.pace'])
        V = mat[:, 2:2 + n].reshape(12, ny, nx)
        ref = np.array([[ol['%d.%d.membrane.V' % (x, y)] for x in range(nx)]
                        for y in range(ny)]).transpose(2, 0, 1)
        assert ref[-1].max() > 0
        # fp32 with 2-ulp division vs IEEE division: small, not zero
        assert np.max(np.abs(V - ref)) < 0.05
        assert np.max(np.abs(mat[:, -1] - ol['5.3.ina.m'])) < 1e-3
        scale = np.abs(ostate).reshape(-1, n_state).max(axis=0)
        assert np.max(np.abs(out - ostate).reshape(-1, n_state) / scale) < 1e-3

        # second run on the resident state
        rc2 = capi.RunConfig()
        rc2.tmin, rc2.tmax, rc2.dt, rc2.log_interval = 6.0, 8.0, 0.005, 1.0
        rc2.n_events, rc2.events = 1, events.ctypes.data
        rc2.n_log, rc2.log_kind = 1, kinds.ctypes.data
        rc2.log_index = index.ctypes.data
        capi.check(lib.mkb_sim_rearm(sim, ctypes.byref(rc2)))
        while capi.check(lib.mkb_sim_step(sim, ctypes.byref(t),
                                          ctypes.byref(halted))):
            pass
        data, rows = ctypes.c_void_p(), ctypes.c_uint64()
        cols, stride = ctypes.c_uint64(), ctypes.c_uint64()
        capi.check(lib.mkb_sim_log_view(sim, ctypes.byref(data),
                                        ctypes.byref(rows), ctypes.byref(cols),
                                        ctypes.byref(stride)))
        assert (rows.value, cols.value) == (2, 1)
        times = (ctypes.c_float * (2 * stride.value)).from_address(data.value)
        assert [times[0], times[stride.value]] == [6.0, 7.0]
    finally:
        lib.mkb_sim_clean(sim)


def test_c_abi_rejects_bad_input():
    lib = capi.library()
    cfg = capi.SimConfig()
    sim = ctypes.c_void_p()
    assert lib.mkb_sim_init(ctypes.byref(cfg), ctypes.byref(sim)) == \
        capi.MKB_ERR_INVALID
    assert b'ABI version' in lib.mkb_last_error()
    cfg.abi_version = capi.MKB_ABI_VERSION
    cfg.precision = 16
    assert lib.mkb_sim_init(ctypes.byref(cfg), ctypes.byref(sim)) == \
        capi.MKB_ERR_INVALID
    assert b'single and double' in lib.mkb_last_error()
    assert lib.mkb_sim_step(None, None, None) == capi.MKB_ERR_STATE
    peaks = capi.measure_peaks(0)
    assert peaks['fp64_fma_ginstr_s'] > 1000 and peaks['copy_gb_s'] > 1000
