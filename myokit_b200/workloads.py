"""
The BASELINE.json configurations as code: each builder applies the same
setter calls to any simulation class with the ``SimulationOpenCL`` surface
(``SimulationCUDA`` for the product, ``oracle.OracleSimulation`` for checks
and CPU baselines), with fixed seeds.

Model note: O'Hara-Rudy CiPA and ten Tusscher 2006 are not shipped with the
reference. The "O'Hara-Rudy-class" configurations run on the reference's own
``myokit/tests/data/decker-2009.mmt`` (Decker 2009, 48 states, 341 variables),
the closest shipped model in size and structure; this is stated in every
result they produce.
"""
import os

import numpy as np

import myokit


def data_model(name):
    """A model from the host framework's own test data directory."""
    path = os.path.join(os.path.dirname(myokit.__file__), 'tests', 'data', name)
    return myokit.load_model(path)


def c1_cable(sim_class, n=128, **kw):
    """C1: LR1991 1-D cable, fp64 forward Euler, 1 Hz pacing of 5 cells."""
    m, p, _ = myokit.load('example')
    s = sim_class(m, p, ncells=n, precision=myokit.DOUBLE_PRECISION, **kw)
    s.set_conductance(10)
    s.set_paced_cells(5)
    s.set_step_size(0.005)
    return s


def c2_planar(sim_class, n=512, **kw):
    """C2: LR1991 2-D planar wave, fp32 forward Euler, paced left edge."""
    m, _, _ = myokit.load('example')
    p = myokit.pacing.blocktrain(period=1000, duration=2, offset=1)
    s = sim_class(m, p, ncells=(n, n), precision=myokit.SINGLE_PRECISION, **kw)
    s.set_conductance(10, 10)
    s.set_paced_cells(nx=5, ny=n, x=0, y=0)
    s.set_step_size(0.005)
    return s


def c3_fields(nx, ny):
    """Conductance fields and the ikr.Gbar field of C3 (seeds 1234 / 1235)."""
    rng = np.random.default_rng(1234)
    gx = 10.0 * (1.0 + 0.2 * rng.uniform(-1, 1, size=(ny, nx - 1)))
    gy = 10.0 * (1.0 + 0.2 * rng.uniform(-1, 1, size=(ny - 1, nx)))
    # Non-conducting scar: an (nx/8 x ny/8) block at the centre
    sx, sy = max(nx // 8, 1), max(ny // 8, 1)
    x0, y0 = nx // 2 - sx // 2, ny // 2 - sy // 2
    gx[y0:y0 + sy, max(x0 - 1, 0):x0 + sx] = 0
    gy[max(y0 - 1, 0):y0 + sy, x0:x0 + sx] = 0
    rng = np.random.default_rng(1235)
    base = 0.0138542    # decker-2009 ikr.Gbar
    scale = np.clip(0.3 * rng.standard_normal(size=(ny, nx)), -0.6, 0.6)
    gkr = base * (1.0 + scale)
    return gx, gy, gkr


def c3_hetero(sim_class, nx=2048, ny=None, model=None, **kw):
    """
    C3: ORd-class (decker-2009 proxy) 2-D, fp64 Rush-Larsen, heterogeneous
    conduction (set_conductance_field) + a per-cell ikr.Gbar (set_field),
    paced left edge (2 ms pulse from t = 1 ms).
    """
    ny = nx if ny is None else ny
    if model is None:
        model = data_model('decker-2009.mmt')
    p = myokit.pacing.blocktrain(period=1000, duration=2, offset=1)
    s = sim_class(model, p, ncells=(nx, ny),
                  precision=myokit.DOUBLE_PRECISION, rl=True, **kw)
    gx, gy, gkr = c3_fields(nx, ny)
    s.set_conductance_field(gx, gy)
    s.set_field('ikr.Gbar', gkr)
    s.set_paced_cells(nx=5, ny=ny, x=0, y=0)
    s.set_step_size(0.005)
    return s


STENCIL_ONLY_MMT = """
[[model]]
name: stencil-only
desc: V' = V - dt * i_diff; the fused kernel with the cell model removed
membrane.V = -80

[engine]
time = 0 bind time
pace = 0 bind pace

[membrane]
dot(V) = -(i_diff + engine.pace * stim)
    label membrane_potential
stim = -10
i_diff = 0 bind diffusion_current
"""


def stencil_only(sim_class, nx, ny=None, precision=None, hetero=False, **kw):
    """
    The stencil-only variant SURVEY.md §8(d) asks for: same tile / halo code,
    cell model replaced by ``V' = V - dt * i_diff``. Algorithmic bytes per
    cell-step: 2 * sizeof(Real) (4 * sizeof(Real) with gx / gy fields).
    """
    ny = nx if ny is None else ny
    if precision is None:
        precision = myokit.SINGLE_PRECISION
    m = myokit.parse_model(STENCIL_ONLY_MMT)
    p = myokit.pacing.blocktrain(period=1000, duration=2, offset=1)
    s = sim_class(m, p, ncells=(nx, ny), precision=precision, **kw)
    if hetero:
        rng = np.random.default_rng(99)
        s.set_conductance_field(rng.uniform(0.5, 1.5, size=(ny, nx - 1)),
                                rng.uniform(0.5, 1.5, size=(ny - 1, nx)))
    else:
        s.set_conductance(1.0, 0.5)
    s.set_paced_cells(nx=5, ny=ny, x=0, y=0)
    s.set_step_size(0.005)
    return s


def fibre_mesh(nx=256, ny=128, nz=128, g_long=10.0, g_trans=2.0,
               extra=0.01, seed=7):
    """
    Edge list of C5-ii: an ``nx * ny * nz`` lattice flattened
    to 1-d cell ids (x fastest) with 6-neighbour edges, ``g_long`` along x and
    ``g_trans`` across, plus ``extra`` (fraction of cells) random long-range
    edges. Returns ``(n_cells, (i, j, g))`` with ``i < j``, no duplicates.
    """
    n = nx * ny * nz
    ids = np.arange(n, dtype=np.int64).reshape(nz, ny, nx)
    parts = []
    for a, b, g in ((ids[:, :, :-1], ids[:, :, 1:], g_long),
                    (ids[:, :-1, :], ids[:, 1:, :], g_trans),
                    (ids[:-1, :, :], ids[1:, :, :], g_trans)):
        parts.append((a.ravel(), b.ravel(), np.full(a.size, float(g))))
    rng = np.random.default_rng(seed)
    k = int(extra * n)
    if k:
        a = rng.integers(0, n, size=k)
        b = rng.integers(0, n, size=k)
        keep = np.abs(a - b) > nx * ny      # never a lattice neighbour
        a, b = a[keep], b[keep]
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        _, first = np.unique(lo * n + hi, return_index=True)
        parts.append((lo[first], hi[first], np.full(len(first), 0.5)))
    i = np.concatenate([p[0] for p in parts])
    j = np.concatenate([p[1] for p in parts])
    g = np.concatenate([p[2] for p in parts])
    return n, (i, j, g)


def c5_mesh(sim_class, nx=256, ny=128, nz=128, precision=None, **kw):
    """C5-ii: LR1991 on the fibre mesh through set_connections (one GPU)."""
    if precision is None:
        precision = myokit.SINGLE_PRECISION
    m, _, _ = myokit.load('example')
    p = myokit.pacing.blocktrain(period=1000, duration=2, offset=1)
    n, edges = fibre_mesh(nx, ny, nz)
    s = sim_class(m, p, ncells=n, precision=precision, **kw)
    s.set_connections(edges)
    s.set_paced_cells(nx)       # the first row of the first plane
    s.set_step_size(0.005)
    return s


# Algorithmic bytes per cell-step, SURVEY.md §8(d) / BASELINE.md §3:
# B = (2 * n_state + n_field + n_gfield + 1) * sizeof(Real)
def algorithmic_bytes(n_state, n_field, n_gfield, real_size):
    return (2 * n_state + n_field + n_gfield + 1) * real_size
