"""
CPU oracle for the SimulationOpenCL hot path — TEST INFRASTRUCTURE ONLY.

Nothing in here is part of the product. Only ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` may
import, call, link or execute anything under ``oracle/``; ``myokit_b200``
never does and has no CPU fallback.

``cgen.py``      restatement of the reference's kernel generator (plain C out)
``driver.c``     restatement of the reference's host loop, pacing, diffusion
``cl_shim.h``    lets the reference's own rendered OpenCL kernel compile as C
``oracle.py``    ``OracleSimulation`` (SimulationOpenCL surface) over both
``_ref/``        (git-ignored) reference-rendered kernels, built on demand
``_build/``      (git-ignored) compiled restatements

Parity status: PINNED (see ``oracle.py`` and ``tests/test_oracle.py``).
"""
