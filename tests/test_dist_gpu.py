"""pytest -m gpu (needs >= 2 GPUs, skipped otherwise): 2-rank data-parallel parity through torch.distributed.run."""
import json
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_ddp_matches_oracle():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29651', os.path.join(ROOT, 'tests', 'dist_gpu_worker.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith('DIST_RESULT ')]
    assert line, (r.stdout[-2000:], r.stderr[-2000:])
    res = json.loads(line[0][len('DIST_RESULT '):])
    check_dist_result(res)


def check_dist_result(res):
    """pass/fail bars of the 2-rank worker report (shared with __graft_entry__.smoke)."""
    for rr in res:
        for k, tol in (('loss_ct', 1.5e-2), ('loss_rc', 1e-2)):   # loss_ct: SyncBN over only 32 rows (see model_checks)
            a, b = rr[k]
            assert abs(a - b) <= tol * abs(b), rr
        assert rr['grad_cos_all'] > 0.97, rr          # bf16 vs fp32 over ALL parameters (see tests/model_checks.py)
        assert rr['grads_equal_across_ranks'], rr
        assert rr['bn_rm_err'] < 5e-2, rr
    assert res[0]['loss_ct'][1] != res[1]['loss_ct'][1]
    for rr in res:
        mo = rr['moco']
        assert mo['ptr'] == 128 and mo['queues_equal'] and mo['loss'] > 0, rr
        assert mo['own_keys_err'] < 2e-2 and mo['rows_match_queue'] < 1e-2, rr   # bf16 encoder / bf16 row copy
