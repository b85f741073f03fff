import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz')))
    return load


def pytest_terminal_summary(terminalreporter):
    """Largest decode error seen per quantity in this session (tests/test_gpu_parity.py::close), relative and absolute;
    also written to gpurun_out/parity_errors.json when that directory exists."""
    mod = sys.modules.get('test_gpu_parity')
    seen = getattr(mod, 'OBSERVED', None)
    if not seen:
        return
    terminalreporter.write_line('max observed decode error per quantity (relative | absolute):')
    for what in sorted(seen):
        rel, ab = seen[what]
        terminalreporter.write_line(f'  {what:28s} {rel:.3e} | {ab:.3e}')
    out_dir = os.path.join(ROOT, 'gpurun_out')
    if os.path.isdir(out_dir):
        import json
        with open(os.path.join(out_dir, 'parity_errors.json'), 'w') as f:
            json.dump({k: {'max_rel': v[0], 'max_abs': v[1]} for k, v in seen.items()}, f, indent=1, sort_keys=True)
