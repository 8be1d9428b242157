"""
Generates tests/golden/*.npz by running the REFERENCE itself
(myokit.Simulation1d, the CPU sibling of SimulationOpenCL: myokit/_sim/cable.py
+ cable.c) in the build container. Run from the repo root:

    PYTHONPATH=baseline/_ref python tests/golden/make_golden.py

The fixtures pin the oracle (tests/test_oracle.py) and, on the GPU box, the
CUDA path (tests/test_parity_gpu.py); nothing at test time needs the
reference's compiled simulation.

Cases
  sim1d_lr91_c1     BASELINE configs[0] shape: LR1991 ('example'), 128 cells,
                    dt 0.005, 1 Hz protocol, g = 10, 5 paced cells; first
                    120 ms (stimulus at 50 ms, upstroke and propagation),
                    V of every cell logged each 1 ms, plus the final state.
  sim1d_br77        the reference's own cross-check configuration
                    (myokit/tests/test_simulation_opencl_vs_sim1d.py:28-136):
                    Beeler-Reuter 1977, 10 cells, dt 0.005, 15 ms,
                    log_interval 0.5; time, pace, V, i_diff, Isi.
  sim1d_lr91_rl     LR1991 with Rush-Larsen updates, 32 cells, 80 ms.
  sim1d_decker_rl   the bench model (decker-2009, 48 states) with Rush-Larsen
                    updates, 12 cells, dt 0.005, 12 ms with a 2 ms stimulus
                    from t = 1; V, a gate, a concentration and i_diff.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle._locate import import_myokit  # noqa: E402

myokit = import_myokit()


def save(name, log, state, meta):
    out = {'state': np.array(state, dtype=np.float64)}
    for k, v in log.items():
        out['log:' + k] = np.array(v, dtype=np.float64)
    for k, v in meta.items():
        out['meta:' + k] = np.array(v)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, len(out), 'arrays')


def main():
    m, p, _ = myokit.load('example')
    s = myokit.Simulation1d(m, p, ncells=128)
    s.set_conductance(10)
    s.set_paced_cells(5)
    s.set_step_size(0.005)
    d = s.run(120, log=['engine.time', 'engine.pace', 'membrane.V'],
              log_interval=1)
    save('sim1d_lr91_c1', d, s.state(),
         dict(ncells=128, dt=0.005, duration=120, log_interval=1, g=10,
              paced=5))

    mb = myokit.load_model(os.path.join(
        os.path.dirname(myokit.__file__), 'tests', 'data',
        'beeler-1977-model.mmt'))
    pb = myokit.pacing.blocktrain(duration=2, offset=1, period=1000)
    s = myokit.Simulation1d(mb, pb, ncells=10)
    s.set_conductance(10)
    s.set_paced_cells(3)
    s.set_step_size(0.005)
    logvars = ['engine.time', 'engine.pace', 'membrane.V', 'membrane.i_diff',
               'isi.Isi']
    d = s.run(15, log=logvars, log_interval=0.5)
    save('sim1d_br77', d, s.state(),
         dict(ncells=10, dt=0.005, duration=15, log_interval=0.5, g=10,
              paced=3))

    s = myokit.Simulation1d(m, p, ncells=32, rl=True)
    s.set_conductance(10)
    s.set_paced_cells(5)
    s.set_step_size(0.01)
    d = s.run(80, log=['engine.time', 'membrane.V', 'ina.m'], log_interval=1)
    save('sim1d_lr91_rl', d, s.state(),
         dict(ncells=32, dt=0.01, duration=80, log_interval=1, g=10, paced=5))

    md = myokit.load_model(os.path.join(
        os.path.dirname(myokit.__file__), 'tests', 'data', 'decker-2009.mmt'))
    s = myokit.Simulation1d(md, pb, ncells=12, rl=True)
    s.set_conductance(10)
    s.set_paced_cells(3)
    s.set_step_size(0.005)
    d = s.run(12, log=['engine.time', 'engine.pace', 'membrane.V',
                       'membrane.i_diff', 'ina.m', 'calcium.uCa_i'],
              log_interval=0.5)
    save('sim1d_decker_rl', d, s.state(),
         dict(ncells=12, dt=0.005, duration=12, log_interval=0.5, g=10,
              paced=3))


if __name__ == '__main__':
    main()
