"""Kernel-option sweep on the C3 workload: compile stats here, timings on a GPU.

    SWEEP_SET=r2a SWEEP_STEPS=30 python scripts/sweep_c3.py
"""
import os, sys, time, re
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import myokit_b200, myokit
from myokit_b200 import workloads, capi

grid = int(os.environ.get('SWEEP_GRID', '2048'))
steps = int(os.environ.get('SWEEP_STEPS', '20'))
B3 = dict(div_cubic=True, fast_exp='stab', fast_libm=False, select=False)
OLD = dict(div_cubic=False, fast_libm=False, select=False)
LM = dict(select=False)
ND = dict(block=(128, 2), min_blocks=2, load_ahead=24, prefetch='l1', select=False)
ST = dict(ND, stage=True, stage_group=64)
SB = dict(ND, stage=True, stage_group=(8,), load_ahead=4, select='cheap')


def also(base, **kw):
    d = dict(base)
    d.update(kw)
    return d


SETS = {
    'r1': [
        ('default', dict(OLD)),
        ('div_parallel', dict(div_parallel=True)),
        ('exp estrin', dict(fast_exp='estrin')),
        ('exp table', dict(fast_exp='table')),
        ('split gates', dict(split_gates=True)),
        ('const_div off', dict(const_div=False)),
    ],
    # round 2, first pass: cheaper division / exp, thread-block shapes, prefetch
    'r2a': [
        ('default', dict()),
        ('div_cubic', dict(div_cubic=True)),
        ('exp stab', dict(fast_exp='stab')),
        ('cubic+stab', B3),
        ('cubic+stab 64x2 mb4', also(B3, block=(64, 2), min_blocks=4)),
        ('cubic+stab 32x4 mb4', also(B3, block=(32, 4), min_blocks=4)),
        ('cubic+stab 128x1 mb4', also(B3, block=(128, 1), min_blocks=4)),
        ('cubic+stab 32x2 mb8', also(B3, block=(32, 2), min_blocks=8)),
        ('cubic+stab 64x1 mb8', also(B3, block=(64, 1), min_blocks=8)),
        ('cubic+stab 64x2 mb5 (96 regs)', also(B3, block=(64, 2), min_blocks=5)),
        ('cubic+stab 64x2 mb6 (80 regs)', also(B3, block=(64, 2), min_blocks=6)),
        ('cubic+stab 64x6 mb1 (168 regs)', also(B3, block=(64, 6), min_blocks=1, max_registers=168)),
        ('cubic+stab 64x3 mb2 (168 regs)', also(B3, block=(64, 3), min_blocks=2, max_registers=168)),
        ('cubic+stab la4', also(B3, load_ahead=4)),
        ('cubic+stab la8', also(B3, load_ahead=8)),
        ('cubic+stab la16', also(B3, load_ahead=16)),
        ('cubic+stab la64', also(B3, load_ahead=64)),
        ('cubic+stab pf-l1 la4', also(B3, prefetch='l1', load_ahead=4)),
        ('cubic+stab pf-l1 la8', also(B3, prefetch='l1', load_ahead=8)),
        ('cubic+stab pf-l1 la16', also(B3, prefetch='l1', load_ahead=16)),
        ('cubic+stab pf-l2 la8', also(B3, prefetch='l2', load_ahead=8)),
        ('cubic+stab pf-l1 la8 64x2 mb4', also(B3, prefetch='l1', load_ahead=8, block=(64, 2), min_blocks=4)),
        ('cubic+stab split gates', also(B3, split_gates=True)),
        ('cubic+stab nofmad', also(B3, fmad=False)),
    ],
    # round 2, second pass: where does the time go (diagnostic builds, wrong results)
    'r2b': [
        ('default', dict()),
        ('default, loads hit L1 (diagnostic)', dict(debug_mem='l1')),
        ('default, L1 loads + no stores (diag.)', dict(debug_mem='l1ns')),
        ('cubic pf-l1 la16', dict(div_cubic=True, prefetch='l1', load_ahead=16)),
        ('cubic pf-l1 la32', dict(div_cubic=True, prefetch='l1', load_ahead=32)),
        ('cubic pf-l1 la24', dict(div_cubic=True, prefetch='l1', load_ahead=24)),
        ('cubic+stab pf-l1 la16', also(B3, prefetch='l1', load_ahead=16)),
        ('cubic+stab pf-l1 la16 (diag. l1)', also(B3, prefetch='l1', load_ahead=16, debug_mem='l1')),
        ('cubic+stab pf-l1 la16 (diag. l1ns)', also(B3, prefetch='l1', load_ahead=16, debug_mem='l1ns')),
        ('cubic+stab la4 (diag. l1)', also(B3, load_ahead=4, debug_mem='l1')),
        ('cubic+stab la4 (diag. l1ns)', also(B3, load_ahead=4, debug_mem='l1ns')),
        ('cubic+stab la4 168 regs 64x3 (diag. l1ns)', also(B3, load_ahead=4, debug_mem='l1ns', block=(64, 3), min_blocks=2, max_registers=168)),
        ('cubic+stab la4 96 regs 64x2 (diag. l1ns)', also(B3, load_ahead=4, debug_mem='l1ns', block=(64, 2), min_blocks=5)),
        ('cubic+stab la4 80 regs 64x2 (diag. l1ns)', also(B3, load_ahead=4, debug_mem='l1ns', block=(64, 2), min_blocks=6)),
    ],
    # round 2, third pass: branch-free body (selects, in-line sqrt/log/cos/acos/pow)
    'r2c': [
        ('old default (libdevice libm, ternaries)', dict(OLD)),
        ('old + cubic', also(OLD, div_cubic=True)),
        ('select only', also(OLD, div_cubic=True, select=True)),
        ('libm only', also(OLD, div_cubic=True, fast_libm=True)),
        ('NEW default (cubic, libm, select)', dict()),
        ('new la16', dict(load_ahead=16)),
        ('new la8', dict(load_ahead=8)),
        ('new la4', dict(load_ahead=4)),
        ('new la64', dict(load_ahead=64)),
        ('new 64x3 mb2 (168 regs)', dict(block=(64, 3), min_blocks=2, max_registers=168)),
        ('new 64x2 mb3 (168 regs)', dict(block=(64, 2), min_blocks=3, max_registers=168)),
        ('new 64x2 mb2 (255 regs)', dict(block=(64, 2), min_blocks=2, max_registers=255)),
        ('new 64x4 mb1 (255 regs)', dict(block=(64, 4), min_blocks=1, max_registers=255)),
        ('new 64x2 mb5 (96 regs)', dict(block=(64, 2), min_blocks=5)),
        ('new 128x1 mb4', dict(block=(128, 1), min_blocks=4)),
        ('new 32x4 mb4', dict(block=(32, 4), min_blocks=4)),
        ('new pf-l1 la16', dict(prefetch='l1', load_ahead=16)),
        ('new estrin', dict(fast_exp='estrin')),
        ('new stab', dict(fast_exp='stab')),
        ('new nofmad', dict(fmad=False)),
        ('new (diag. l1)', dict(debug_mem='l1')),
    ],
    'r2d': [
        ('libm', dict(LM)),
        ('libm exp-add', also(LM, exp_scale='add')),
        ('libm la16', also(LM, load_ahead=16)),
        ('libm la48', also(LM, load_ahead=48)),
        ('libm pf-l1 la16', also(LM, prefetch='l1', load_ahead=16)),
        ('libm 64x2 mb3 (168 regs)', also(LM, block=(64, 2), min_blocks=3, max_registers=168)),
        ('libm 64x3 mb2 (168 regs)', also(LM, block=(64, 3), min_blocks=2, max_registers=168)),
        ('libm 32x3 mb5 (136 regs)', also(LM, block=(32, 3), min_blocks=5, max_registers=136)),
        ('libm 64x1 mb7 (144 regs)', also(LM, block=(64, 1), min_blocks=7, max_registers=144)),
        ('libm cheap-select', also(LM, select='cheap')),
        ('libm cheap-select exp-add', also(LM, select='cheap', exp_scale='add')),
        ('libm cheap-select 64x2 mb3 (168 regs)', also(LM, select='cheap', block=(64, 2), min_blocks=3, max_registers=168)),
        ('libm cheap-select la16', also(LM, select='cheap', load_ahead=16)),
        ('libm select 64x2 mb3 (168) la16', dict(block=(64, 2), min_blocks=3, max_registers=168, load_ahead=16)),
        ('libm select 64x2 mb3 (168) la8', dict(block=(64, 2), min_blocks=3, max_registers=168, load_ahead=8)),
        ('libm select 64x2 mb3 (168) exp-add', dict(block=(64, 2), min_blocks=3, max_registers=168, exp_scale='add')),
        ('libm select 32x4 mb3 (168)', dict(block=(32, 4), min_blocks=3, max_registers=168)),
        ('libm select 128x1 mb3 (168)', dict(block=(128, 1), min_blocks=3, max_registers=168)),
        ('libm (diag. l1)', also(LM, debug_mem='l1')),
        ('libm cheap-select (diag. l1)', also(LM, select='cheap', debug_mem='l1')),
    ],
    'r2e': [
        ('libm pf-l1 la16', also(LM, prefetch='l1', load_ahead=16)),
        ('libm pf-l1 la8', also(LM, prefetch='l1', load_ahead=8)),
        ('libm pf-l1 la4', also(LM, prefetch='l1', load_ahead=4)),
        ('libm pf-l1 la24', also(LM, prefetch='l1', load_ahead=24)),
        ('libm pf-l1 la32', also(LM, prefetch='l1', load_ahead=32)),
        ('libm pf-l2 la16', also(LM, prefetch='l2', load_ahead=16)),
        ('libm pf-l1 la16 128x1 mb4', also(LM, prefetch='l1', load_ahead=16, block=(128, 1), min_blocks=4)),
        ('libm pf-l1 la16 128x2 mb2', also(LM, prefetch='l1', load_ahead=16, block=(128, 2), min_blocks=2)),
        ('libm pf-l1 la16 64x2 mb4', also(LM, prefetch='l1', load_ahead=16, block=(64, 2), min_blocks=4)),
        ('libm pf-l1 la16 32x8 mb2', also(LM, prefetch='l1', load_ahead=16, block=(32, 8), min_blocks=2)),
        ('libm pf-l1 la16 cheap-select', also(LM, prefetch='l1', load_ahead=16, select='cheap')),
        ('libm pf-l1 la8 cheap-select', also(LM, prefetch='l1', load_ahead=8, select='cheap')),
        ('libm pf-l1 la16 select 128x1 mb3 (168)', dict(prefetch='l1', load_ahead=16, block=(128, 1), min_blocks=3, max_registers=168)),
        ('libm pf-l1 la8 select 128x1 mb3 (168)', dict(prefetch='l1', load_ahead=8, block=(128, 1), min_blocks=3, max_registers=168)),
        ('libm pf-l1 la16 64x1 mb7', also(LM, prefetch='l1', load_ahead=16, block=(64, 1), min_blocks=7, max_registers=144)),
        ('libm pf-l1 la16 nofmad', also(LM, prefetch='l1', load_ahead=16, fmad=False)),
        ('libm pf-l1 la16 div newton', also(LM, prefetch='l1', load_ahead=16, div_cubic=False)),
        ('libm pf-l1 la16 estrin', also(LM, prefetch='l1', load_ahead=16, fast_exp='estrin')),
        ('libm pf-l1 la16 (diag. l1)', also(LM, prefetch='l1', load_ahead=16, debug_mem='l1')),
    ],
    'r2f': [
        ('default', dict(block=None, min_blocks=2, load_ahead=24, prefetch='l1', select=False)),
        ('64x4 la24', also(LM, prefetch='l1', load_ahead=24)),
        ('128x2 la24', also(LM, prefetch='l1', load_ahead=24, block=(128, 2))),
        ('128x2 la20', also(LM, prefetch='l1', load_ahead=20, block=(128, 2))),
        ('128x2 la28', also(LM, prefetch='l1', load_ahead=28, block=(128, 2))),
        ('256x1 la24', also(LM, prefetch='l1', load_ahead=24, block=(256, 1))),
        ('128x2 la24 pf-l2', also(LM, prefetch='l2', load_ahead=24, block=(128, 2))),
        ('128x4 mb1 la24', also(LM, prefetch='l1', load_ahead=24, block=(128, 4), min_blocks=1, max_registers=128)),
        ('128x2 la24 (diag. l1)', also(LM, prefetch='l1', load_ahead=24, block=(128, 2), debug_mem='l1')),
    ],
    # round 2, last pass: the plane stride as a compile-time constant and the
    # one-instruction special-divisor test (4217 -> 3578 SASS instructions)
    'r3a': [
        ('r2 default (run-time stride, int check)', also(ND, plane_stride=False, div_int_check=True)),
        ('fminf test only', also(ND, plane_stride=False)),
        ('const stride only', also(ND, div_int_check=True)),
        ('NEW default', dict(ND)),
        ('new 64x4', also(ND, block=(64, 4))),
        ('new 256x1', also(ND, block=(256, 1))),
        ('new 128x1 mb4', also(ND, block=(128, 1), min_blocks=4)),
        ('new 64x3 (112 regs)', also(ND, block=(64, 3), min_blocks=None, max_registers=112)),
        ('new 64x1 (112 regs)', also(ND, block=(64, 1), min_blocks=None, max_registers=112)),
        ('new select 64x3 (112 regs)', also(ND, select=True, block=(64, 3), min_blocks=None, max_registers=112)),
        ('new 128x1 mb5 (96 regs)', also(ND, block=(128, 1), min_blocks=5)),
        ('new la16', also(ND, load_ahead=16)),
        ('new la20', also(ND, load_ahead=20)),
        ('new la28', also(ND, load_ahead=28)),
        ('new la32', also(ND, load_ahead=32)),
        ('new la40', also(ND, load_ahead=40)),
        ('new no prefetch', also(ND, prefetch=None)),
        ('new pf-l2', also(ND, prefetch='l2')),
        ('new select', also(ND, select=True)),
        ('new select la8', also(ND, select=True, load_ahead=8)),
        ('new select 128x1 mb3 (168)', also(ND, select=True, block=(128, 1), min_blocks=3, max_registers=168)),
        ('new select 64x3 mb2 (168)', also(ND, select=True, block=(64, 3), min_blocks=2, max_registers=168)),
        ('new 128x1 mb3 (168)', also(ND, block=(128, 1), min_blocks=3, max_registers=168)),
        ('new cheap-select', also(ND, select='cheap')),
        ('new exp-add', also(ND, exp_scale='add')),
        ('new estrin', also(ND, fast_exp='estrin')),
        ('new (diag. l1)', also(ND, debug_mem='l1')),
    ],
    # staged states: tiles of every state plane through shared memory by TMA
    'r3b': [
        ('NEW default', dict(ND)),
        ('stage la24', also(ND, stage=True)),
        ('stage la8', also(ND, stage=True, load_ahead=8)),
        ('stage la4', also(ND, stage=True, load_ahead=4)),
        ('stage la2', also(ND, stage=True, load_ahead=2)),
        ('stage la1', also(ND, stage=True, load_ahead=1)),
        ('stage la4 groups of 4', also(ND, stage=True, load_ahead=4, stage_group=4)),
        ('stage la4 groups of 16', also(ND, stage=True, load_ahead=4, stage_group=16)),
        ('stage la4 one group', also(ND, stage=True, load_ahead=4, stage_group=64)),
        ('stage la4 64x4', also(ND, stage=True, load_ahead=4, block=(64, 4))),
        ('stage la4 256x1', also(ND, stage=True, load_ahead=4, block=(256, 1))),
        ('stage la4 128x1 mb4', also(ND, stage=True, load_ahead=4, block=(128, 1), min_blocks=4)),
        ('stage la4 64x2 mb4', also(ND, stage=True, load_ahead=4, block=(64, 2), min_blocks=4)),
        ('stage la4 select', also(ND, stage=True, load_ahead=4, select=True)),
        ('stage la8 select', also(ND, stage=True, load_ahead=8, select=True)),
        ('stage la4 cheap-select', also(ND, stage=True, load_ahead=4, select='cheap')),
        ('stage la4 128x1 mb3 (168)', also(ND, stage=True, load_ahead=4, block=(128, 1), min_blocks=3, max_registers=168)),
        ('stage la4 select 128x1 mb3 (168)', also(ND, stage=True, load_ahead=4, select=True, block=(128, 1), min_blocks=3, max_registers=168)),
        ('stage la4 128x1 mb5 (96)', also(ND, stage=True, load_ahead=4, block=(128, 1), min_blocks=5)),
        ('stage la4 estrin', also(ND, stage=True, load_ahead=4, fast_exp='estrin')),
        ('stage la4 exp-add', also(ND, stage=True, load_ahead=4, exp_scale='add')),
        ('stage la4 overlap', also(ND, stage=True, load_ahead=4, overlap=True)),
        ('overlap (no stage)', also(ND, overlap=True)),
    ],
    'r3c': [
        ('NEW default', dict(ND)),
        ('NEW default, 300 steps', also(ND, _steps=300)),
        ('stage one group la4', also(ST, load_ahead=4)),
        ('stage one group la4, 300 steps', also(ST, load_ahead=4, _steps=300)),
        ('stage one group la2', also(ST, load_ahead=2)),
        ('stage one group la8', also(ST, load_ahead=8)),
        ('stage one group la24', also(ST, load_ahead=24)),
        ('stage 8 + rest la4', also(ST, load_ahead=4, stage_group=(8,))),
        ('stage 8 + rest la8', also(ST, load_ahead=8, stage_group=(8,))),
        ('stage 16 + rest la4', also(ST, load_ahead=4, stage_group=(16,))),
        ('stage 4 + 12 + rest la4', also(ST, load_ahead=4, stage_group=(4, 12))),
        ('stage one group la4 cheap-select', also(ST, load_ahead=4, select='cheap')),
        ('stage one group la8 cheap-select', also(ST, load_ahead=8, select='cheap')),
        ('stage 8 + rest la4 cheap-select', also(ST, load_ahead=4, stage_group=(8,), select='cheap')),
        ('stage one group la8 select', also(ST, load_ahead=8, select=True)),
        ('stage one group la4 64x4', also(ST, load_ahead=4, block=(64, 4))),
        ('stage one group la4 256x1', also(ST, load_ahead=4, block=(256, 1))),
        ('stage one group la4 64x2 mb4', also(ST, load_ahead=4, block=(64, 2), min_blocks=4)),
        ('stage one group la4 overlap', also(ST, load_ahead=4, overlap=True)),
        ('stage one group la4 exp-add', also(ST, load_ahead=4, exp_scale='add')),
    ],
    'r3d': [
        ('NEW default', dict(ND)),
        ('stage 8+ la4 cheap', dict(SB)),
        ('stage 8+ la4 cheap next .3', also(SB, prefetch_next=0.3)),
        ('stage 8+ la4 cheap next .5', also(SB, prefetch_next=0.5)),
        ('stage 8+ la4 cheap next .7', also(SB, prefetch_next=0.7)),
        ('stage 8+ la4 cheap next .9', also(SB, prefetch_next=0.9)),
        ('stage 8+ la4 cheap next .01', also(SB, prefetch_next=0.01)),
        ('stage 8+ la4 cheap, direct stores', also(SB, stage_store=False)),
        ('stage 8+ la4 cheap, direct stores, next .5', also(SB, stage_store=False, prefetch_next=0.5)),
        ('stage 8+ la8 cheap next .5', also(SB, load_ahead=8, prefetch_next=0.5)),
        ('stage 8+ la4 next .5', also(SB, select=False, prefetch_next=0.5)),
        ('stage 4+12+ la4 cheap next .5', also(SB, stage_group=(4, 12), prefetch_next=0.5)),
        ('stage one group la8 next .5', also(ST, load_ahead=8, prefetch_next=0.5)),
        ('stage 8+ la4 cheap next .5, 300 steps', also(SB, prefetch_next=0.5, _steps=300)),
        ('stage 8+ la4 cheap next .5 overlap', also(SB, prefetch_next=0.5, overlap=True)),
    ],
    # small grids (SWEEP_GRID=768: the cells of a 2048 x 256 slab): overlapping steps x staging
    'r3e': [
        ('plain', also(ND, _steps=200)),
        ('plain overlap', also(ND, overlap=True, _steps=200)),
        ('stage', also(SB, _steps=200)),
        ('stage overlap', also(SB, overlap=True, _steps=200)),
        ('stage overlap direct stores', also(SB, overlap=True, stage_store=False, _steps=200)),
        ('stage one group la8 overlap', also(ST, load_ahead=8, overlap=True, _steps=200)),
    ],
    # V and conductance loads issued before the staging barrier
    'r3f': [
        ('stage 8+ la4 cheap', dict(SB)),
        ('stage 8+ la4 cheap (again)', dict(SB)),
        ('stage 4+ la4 cheap', also(SB, stage_group=(4,))),
        ('stage 16+ la4 cheap', also(SB, stage_group=(16,))),
        ('stage 4+12+ la4 cheap', also(SB, stage_group=(4, 12))),
        ('stage one group la4 cheap', also(SB, stage_group=64)),
        ('stage 8+ la2 cheap', also(SB, load_ahead=2)),
        ('stage 8+ la8 cheap', also(SB, load_ahead=8)),
        ('stage 8+ la4', also(SB, select=False)),
        ('stage 8+ la4 cheap 64x4', also(SB, block=(64, 4))),
        ('stage 8+ la4 cheap 256x1', also(SB, block=(256, 1))),
        ('stage 8+ la4 cheap direct stores', also(SB, stage_store=False)),
    ],
    # the staged kernel as a loop over tiles (next tile's V and first planes requested a tile ahead)
    'r3g': [
        ('stage 8+ la4 cheap', dict(SB)),
        ('tile loop early 4 la4 cheap', also(SB, tile_loop=True)),
        ('tile loop early 4 la2 cheap', also(SB, tile_loop=True, load_ahead=2)),
        ('tile loop early 4 la8 cheap', also(SB, tile_loop=True, load_ahead=8)),
        ('tile loop early 2 la4 cheap', also(SB, tile_loop=True, stage_early=2)),
        ('tile loop early 3 la4 cheap', also(SB, tile_loop=True, stage_early=3)),
        ('tile loop early 5 la4 cheap', also(SB, tile_loop=True, stage_early=5)),
        ('tile loop early 4 la4', also(SB, tile_loop=True, select=False)),
        ('tile loop early 4 la4 cheap 256x1', also(SB, tile_loop=True, block=(256, 1))),
        ('tile loop early 4 la4 cheap 64x4', also(SB, tile_loop=True, block=(64, 4))),
        ('tile loop early 4 la4 cheap, 300 steps', also(SB, tile_loop=True, _steps=300)),
    ],
    # loop over tiles, body in uniform control flow
    'r3i': [
        ('stage 8+ la4 cheap', dict(SB)),
        ('tile loop early 4 la4 cheap', also(SB, tile_loop=True)),
        ('tile loop early 4 la2 cheap', also(SB, tile_loop=True, load_ahead=2)),
        ('tile loop early 4 la8 cheap', also(SB, tile_loop=True, load_ahead=8)),
        ('tile loop early 3 la4 cheap', also(SB, tile_loop=True, stage_early=3)),
        ('tile loop early 5 la4 cheap', also(SB, tile_loop=True, stage_early=5)),
        ('tile loop early 4 la4', also(SB, tile_loop=True, select=False)),
        ('tile loop early 4 la4 select', also(SB, tile_loop=True, select=True)),
        ('tile loop early 4 la4 cheap 64x4', also(SB, tile_loop=True, block=(64, 4))),
        ('tile loop early 4 la4 cheap call', also(SB, tile_loop=True, tile_call=True)),
        ('tile loop early 4 la4 cheap, 300 steps', also(SB, tile_loop=True, _steps=300)),
    ],
    # no V tile: left / right neighbours by warp shuffle, above / below from memory
    'r3j': [
        ('stage 8+ la4 cheap', dict(SB)),
        ('v_direct', also(SB, v_direct=True)),
        ('v_direct (again)', also(SB, v_direct=True)),
        ('v_direct 256x1', also(SB, v_direct=True, block=(256, 1))),
        ('v_direct 64x4', also(SB, v_direct=True, block=(64, 4))),
        ('v_direct 4+', also(SB, v_direct=True, stage_group=(4,))),
        ('v_direct one group', also(SB, v_direct=True, stage_group=64)),
        ('v_direct la2', also(SB, v_direct=True, load_ahead=2)),
        ('v_direct la8', also(SB, v_direct=True, load_ahead=8)),
        ('v_direct, 300 steps', also(SB, v_direct=True, _steps=300)),
        ('stage 8+ la4 cheap, 300 steps', also(SB, _steps=300)),
    ],
}
variants = SETS[os.environ.get('SWEEP_SET', 'r2a')]
only = os.environ.get('SWEEP_ONLY')
gpu = capi.device_count() > 0
s = None
for name, opts in variants:
    if only and only not in name:
        continue
    opts = dict(opts)
    nsteps = opts.pop('_steps', steps)
    if s is None or not gpu:
        s = workloads.c3_hetero(myokit_b200.SimulationCUDA, nx=grid if gpu else 16)
    # one simulation, re-optioned: the state stays in HBM across variants
    s.set_kernel_options(**dict(dict(
        block=(64, 4), min_blocks=2, max_registers=0, load_ahead=32, prefetch=None,
        div_cubic=True, fast_libm=True, select=True,
        fast_exp='poly', split_gates=False, div_parallel=False,
        const_div=True, fmad=True, debug_mem=None, exp_scale='mul',
        plane_stride=True, div_int_check=False, stage=False, stage_group=8,
        overlap=False, stage_store=True, prefetch_next=None, tile_loop=False,
        stage_early=4, tile_call=False, v_direct=False), **opts))
    src = s.kernel_source()
    t0 = time.time()
    try:
        cubin, log = capi.jit_compile(src.code, src.options + ('--ptxas-options=-v',))
    except Exception as e:
        print('%-36s compile failed: %s' % (name, str(e)[:200]))
        continue
    m = re.search(r"Compiling entry function 'mkb_cell_step'.*?Used (\d+) registers", log, re.S)
    sp = re.search(r"mkb_cell_step\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores", log, re.S)
    line = '%-36s regs %s stack/spill %s' % (
        name, m.group(1) if m else '?', '/'.join(sp.groups()) if sp else '?')
    if gpu:
        try:
            info = s.benchmark_steps(nsteps, warmup=3)
            ms = info['device_ms'] / info['steps']
            line += '  %.4f ms/step  %.3e cell-steps/s' % (ms, grid * grid / ms * 1e3)
        except Exception as e:
            line += '  run failed: %s' % str(e)[:200]
            s = None
    print(line, flush=True)
