"""
Runs the reference's OWN behavioural tests of SimulationOpenCL against
SimulationCUDA: the unmodified test modules of the
installed host framework (baseline/_ref/myokit/tests/test_simulation_opencl*.py)
with `myokit.SimulationOpenCL` bound to `myokit_b200.SimulationCUDA` and the
"OpenCL found" switches of `myokit.tests` forced on. Needs a B200.

    python scripts/run_reference_suite.py [out.md]

Writes one line per test (pass / FAIL / ERROR / skip, with the first line of
the reason) and exits 0: the list is the deliverable, not an all-green run —
tests that probe OpenCL itself (device selection, kernel text) cannot pass.
"""
import importlib
import os
import sys
import time
import unittest

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import myokit_b200                      # noqa: E402 (puts myokit on sys.path)
import myokit                           # noqa: E402
import myokit.tests                     # noqa: E402

myokit.SimulationOpenCL = myokit_b200.SimulationCUDA
myokit.FiberTissueSimulation = myokit_b200.FiberTissueSimulationCUDA
for name in ('OpenCL_FOUND', 'OpenCL_DOUBLE_PRECISION',
             'OpenCL_DOUBLE_PRECISION_CONNECTIONS'):
    setattr(myokit.tests, name, True)

MODULES = ['test_simulation_opencl', 'test_simulation_opencl_log_interval',
           'test_simulation_opencl_vs_sim1d', 'test_simulation_opencl_vs_cvode']


class Recorder(unittest.TestResult):
    def __init__(self):
        super().__init__()
        self.rows = []

    def _first(self, text):
        lines = [x for x in str(text).strip().splitlines() if x.strip()]
        if not lines:
            return ''
        # (compilation errors of the host framework's own CVODE simulation end
        # with an argument dump: name the exception instead)
        for x in lines:
            if 'Error' in x and not x.startswith(' '):
                return x[:160]
        return lines[-1][:160]

    def addSuccess(self, test):
        self.rows.append((test.id(), 'pass', ''))

    def addFailure(self, test, err):
        super().addFailure(test, err)
        self.rows.append((test.id(), 'FAIL', self._first(self.failures[-1][1])))

    def addError(self, test, err):
        super().addError(test, err)
        self.rows.append((test.id(), 'ERROR', self._first(self.errors[-1][1])))

    def addSkip(self, test, reason):
        super().addSkip(test, reason)
        self.rows.append((test.id(), 'skip', reason[:160]))


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else None
    rec = Recorder()
    t0 = time.time()
    for name in MODULES:
        try:
            mod = importlib.import_module('myokit.tests.' + name)
        except Exception as e:      # e.g. sundials missing for the CVODE module
            rec.rows.append(('myokit.tests.' + name, 'ERROR',
                             'import failed: %s' % str(e)[:140]))
            continue
        unittest.defaultTestLoader.loadTestsFromModule(mod).run(rec)
    counts = {}
    for _, status, _ in rec.rows:
        counts[status] = counts.get(status, 0) + 1
    lines = ['# The reference\'s own SimulationOpenCL tests, run against SimulationCUDA',
             '',
             '`python scripts/run_reference_suite.py` on a B200 (%.0f s): '
             % (time.time() - t0)
             + ', '.join('%d %s' % (v, k) for k, v in sorted(counts.items())) + '.',
             '',
             '| test | result | note |', '|---|---|---|']
    for tid, status, note in rec.rows:
        tid = tid.replace('myokit.tests.', '')
        lines.append('| `%s` | %s | %s |' % (tid, status, note.replace('|', '/')))
    text = '\n'.join(lines) + '\n'
    if out:
        with open(out, 'w') as f:
            f.write(text)
    print(text)


if __name__ == '__main__':
    main()
