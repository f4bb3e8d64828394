"""The one-line swap of INTEGRATION.md §2 really selects the drop-in: with the UNMODIFIED reference package loaded
(oracle/ref_loader.import_cmae, mmengine/mmcv stand-ins), `try_register_mmengine()` force-registers the drop-in classes
into the reference's OWN child registry `cmae.registry.MODELS` (cmae/registry.py:83-84) and
`MODELS.build(cfg.model)` (cmae/models/builder.py:12-14) from the unmodified `configs/cmunet_config.py:5-42` returns
the drop-in `CM_UNet` with the golden parameter keys.  Needs /root/reference (absent on the GPU box -> skipped)."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = textwrap.dedent('''
    import json, sys
    sys.path.insert(0, %r)
    from oracle import ref_loader as RL
    models, MODELS, cfg = RL.import_cmae()
    before = MODELS.get('CM_UNet').__module__
    import contrastive_masked_unet_b200 as C
    done = C.try_register_mmengine()
    import torch
    torch.manual_seed(60)
    m = MODELS.build(cfg)                          # unmodified config (projector.in_channels = 224 * 224)
    from cmae.models.builder import build_algorithm
    m2 = build_algorithm(cfg)
    from mmengine.registry import MODELS as ROOT_MODELS
    print('RESULT ' + json.dumps({
        'before': before, 'done': done, 'module': type(m).__module__, 'module2': type(m2).__module__,
        'children': {n: type(c).__module__ for n, c in m.named_children()},
        'predictor': type(m.head.predictor).__module__,
        'root': ROOT_MODELS.get('CM_UNet').__module__,
        'keys': [k for k, _ in m.named_parameters()],
        'fc0': list(m.projector.fc0.weight.shape)}))
''')


def test_reference_registry_builds_the_drop_in():
    from oracle import ref_loader as RL
    if not RL.reference_available():
        pytest.skip('the reference tree is not present on this box')
    r = subprocess.run([sys.executable, '-c', _SCRIPT % ROOT], capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith('RESULT ')]
    assert line, (r.stdout[-2000:], r.stderr[-3000:])
    res = json.loads(line[0][len('RESULT '):])
    assert res['before'].startswith('cmae.models'), res['before']          # the reference class was registered first
    assert 'cmae.registry.MODELS' in res['done'] and 'mmengine.registry.MODELS' in res['done']
    drop_in = 'contrastive_masked_unet_b200.modules'
    assert res['module'] == drop_in and res['module2'] == drop_in and res['root'] == drop_in
    assert all(v == drop_in for v in res['children'].values()), res['children']
    assert res['predictor'] == drop_in
    gold = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'pretrain.json')))['cases'][0]
    assert res['keys'] == gold['param_keys']
    assert res['fc0'] == [1536, 224 * 224]
