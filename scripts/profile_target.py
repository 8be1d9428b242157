"""One short run of a named workload, for ncu to capture (scripts/collect_profiles.sh)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import myokit_b200, myokit
from myokit_b200 import workloads

S = myokit_b200.SimulationCUDA
name = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
if name == 'c3':
    s = workloads.c3_hetero(S, nx=2048)
elif name == 'stencil32':
    s = workloads.stencil_only(S, 8192, 4096, precision=myokit.SINGLE_PRECISION)
elif name == 'stencil64':
    s = workloads.stencil_only(S, 8192, 4096, precision=myokit.DOUBLE_PRECISION)
elif name == 'lr91_fp32':
    s = workloads.c2_planar(S, 4096)
else:
    raise SystemExit('unknown workload ' + name)
if os.environ.get('MKB_PROFILE_OPTS'):
    # kernel options of the variant to capture, e.g. "dict(div_cubic=True)"
    s.set_kernel_options(**eval(os.environ['MKB_PROFILE_OPTS']))
if os.environ.get('MKB_PROFILE_KEYFILE'):
    # which kernel the capture is of: bench.py trusts the profile only for this key
    with open(os.environ['MKB_PROFILE_KEYFILE'], 'w') as f:
        f.write(s.kernel_source().key())
info = s.benchmark_steps(steps, warmup=3)
print(name, info)
