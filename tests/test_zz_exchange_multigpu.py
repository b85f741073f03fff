"""The exchange protocol across GPUs: scripts/exchange_2gpu.py under torchrun, with and without the NVLS multicast
mapping.  Needs >= 2 GPUs on the box (skipped on the driver's 1-GPU test box; run with `gpurun --gpus 2`)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('flags', [[], ['--no-multicast']])
def test_exchange_protocol_across_gpus(flags):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs')
    port = 29500 + (os.getpid() % 400) + (1 if flags else 0)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(min(n, 8)),
           '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'scripts', 'exchange_2gpu.py')] + flags
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    line = json.loads([l for l in res.stdout.splitlines() if l.startswith('{')][-1])
    print(line)
    assert line['all_ranks_ok'] and line['timeouts'] == 0
