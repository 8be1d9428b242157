"""Shared helpers for the parity tests (CUDA path vs the CPU oracle)."""
import os

import numpy as np

import myokit_b200  # noqa: F401  (puts myokit on sys.path)
import myokit

from oracle.oracle import OracleSimulation


def data_model(name):
    """Loads a model shipped with the host framework's own test data."""
    path = os.path.join(os.path.dirname(myokit.__file__), 'tests', 'data', name)
    return myokit.load_model(path)


def example():
    m, p, _ = myokit.load('example')
    return m, p


def configure(sim, cfg):
    """Applies the same setter calls to a SimulationCUDA or OracleSimulation."""
    if 'dt' in cfg:
        sim.set_step_size(cfg['dt'])
    if 'conductance' in cfg:
        sim.set_conductance(*cfg['conductance'])
    if 'conductance_field' in cfg:
        sim.set_conductance_field(*cfg['conductance_field'])
    if 'connections' in cfg:
        sim.set_connections(cfg['connections'])
    if 'paced_cells' in cfg:
        sim.set_paced_cells(*cfg['paced_cells'])
    if 'paced_cell_list' in cfg:
        sim.set_paced_cell_list(cfg['paced_cell_list'])
    for var, values in cfg.get('fields', {}).items():
        sim.set_field(var, values)
    for var, value in cfg.get('constants', {}).items():
        sim.set_constant(var, value)
    if 'state' in cfg:
        sim.set_state(cfg['state'])
    if 'time' in cfg:
        sim.set_time(cfg['time'])


def run_pair(model, protocol, ncells, duration, log, log_interval=1.0,
             precision=myokit.DOUBLE_PRECISION, diffusion=True, rl=False,
             cfg=None, kernel='port', fmad=True, block=None):
    """
    Runs the CUDA path and the oracle on identical inputs.
    Returns (cuda_log, cuda_state, oracle_log, oracle_state).
    """
    cfg = cfg or {}
    s = myokit_b200.SimulationCUDA(
        model, protocol, ncells=ncells, diffusion=diffusion,
        precision=precision, rl=rl)
    s.set_kernel_options(fmad=fmad, block=block)
    configure(s, cfg)
    d = s.run(duration, log=log, log_interval=log_interval)
    cl = dict((k, np.array(v, dtype=np.float64)) for k, v in d.items())

    o = OracleSimulation(
        model, protocol, ncells=ncells, diffusion=diffusion,
        precision=precision, rl=rl, kernel=kernel)
    configure(o, cfg)
    ol, ostate = o.run(duration, log=log, log_interval=log_interval)
    return cl, s.state_array(), ol, ostate


def max_abs_diff(a, b, keys=None, suffix=None):
    """Max |a[k] - b[k]| over keys (optionally only keys ending in suffix)."""
    worst = 0.0
    for k in (keys or a.keys()):
        if suffix is not None and not k.endswith(suffix):
            continue
        assert len(a[k]) == len(b[k]), (k, len(a[k]), len(b[k]))
        if len(a[k]):
            worst = max(worst, float(np.max(np.abs(a[k] - b[k]))))
    return worst
