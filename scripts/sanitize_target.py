"""
Small runs of every kernel family for compute-sanitizer (scripts/round2_sanitize.sh):
the fused step kernel with logging and graph replay, the register-patch and
streaming (TMA) forms, row slabs (2 slabs sharing one device: peer stores,
arrival flags), a partitioned connection graph (ghost cells + push kernel),
the persistent one-block kernel and the fibre-tissue pair.
"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import myokit_b200, myokit
from myokit_b200 import workloads, multigpu

S = myokit_b200.SimulationCUDA
which = sys.argv[1:] or ['plain', 'vector', 'stream', 'slab', 'graph', 'persistent']
DP, SP = myokit.DOUBLE_PRECISION, myokit.SINGLE_PRECISION

if 'plain' in which:
    # 70 plain steps in a row: one 64-step CUDA graph + single launches, logged rows
    s = workloads.c3_hetero(S, nx=72, ny=20)
    t, f = s.run_fields(0.7, ['membrane.V', 'membrane.i_diff', 'ikr.IKr'], log_interval=0.25)
    print('plain', f['membrane.V'].shape, s.last_run_info()['kernel_launches'])
if 'loop' in which:
    # the staged kernel as a loop over tiles: 396 tiles of 128 x 2 for 296 blocks, ragged rims
    s = workloads.c3_hetero(S, nx=328, ny=66)
    s.set_kernel_options(tile_loop=True, overlap=False)
    assert s.kernel_source().kernel_flags & 16
    t, f = s.run_fields(0.4, ['membrane.V', 'membrane.i_diff'], log_interval=0.2)
    print('loop', f['membrane.V'].shape, s.last_run_info()['kernel_launches'])
if 'vector' in which:
    s = workloads.stencil_only(S, 136, 37, precision=SP, hetero=True)
    s.set_kernel_options(stream=False)
    t, f = s.run_fields(0.5, ['membrane.V'], log_interval=0.25)
    print('vector', f['membrane.V'].shape)
if 'stream' in which:
    for prec in (SP, DP):
        s = workloads.stencil_only(S, 264, 70, precision=prec)
        s.set_kernel_options(stream=True)
        assert s.kernel_source().kernel_flags & 2
        t, f = s.run_fields(0.5, ['membrane.V'], log_interval=0.25)
        print('stream', f['membrane.V'].shape)
if 'slab' in which:
    def target(comm):
        s = workloads.c3_hetero(S, nx=64, ny=24, device=0, comm=comm)
        t, f = s.run_fields(0.4, ['membrane.V'], log_interval=0.2)
        return f['membrane.V'].shape
    print('slab', multigpu.run_threads(2, target))
if 'graph' in which:
    n, edges = workloads.fibre_mesh(16, 8, 8)

    def target(comm):
        m, _, _ = myokit.load('example')
        p = myokit.pacing.blocktrain(period=1000, duration=2, offset=1)
        s = S(m, p, ncells=n, precision=SP, device=0, comm=comm)
        s.set_connections(edges)
        s.set_paced_cells(16)
        t, f = s.run_fields(0.4, ['membrane.V'], log_interval=0.2)
        return f['membrane.V'].shape
    print('graph', multigpu.run_threads(2, target))
if 'persistent' in which:
    s = workloads.c1_cable(S, 128)
    s.set_kernel_options(persistent=True)
    d = s.run(1.0, log=['engine.time', 'membrane.V'], log_interval=0.5)
    print('persistent', len(d['engine.time']))
