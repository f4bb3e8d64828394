"""Pins oracle/mask_oracle.py (CPU restatement of UNet_encoder.py:106-139) against numpy itself and against
golden vectors minted from the unmodified reference (tests/golden/masks.json)."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle.mask_oracle import MT19937, num_masked_patches, patch_mask

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'masks.json')))


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.mark.parametrize('seed,P', [(60, 196), (60, 1024), (61, 256), (42, 4096), (0, 2), (5, 1)])
def test_shuffle_matches_numpy(seed, P):
    np.random.seed(seed)
    r = MT19937(seed)
    for _ in range(4):
        a = np.arange(P)
        np.random.shuffle(a)
        assert a.tolist() == r.shuffle_arange(P)


def test_state_import_from_numpy():
    np.random.seed(123)
    np.random.rand(1000)  # advance
    _, key, pos, _, _ = np.random.get_state()
    r = MT19937()
    r.set_state(key, pos)
    a = np.arange(300)
    np.random.shuffle(a)
    assert a.tolist() == r.shuffle_arange(300)


def test_survey_kats():
    """SURVEY.md Appendix D table."""
    r = MT19937(60)
    assert r.shuffle_arange(196)[:8] == [29, 46, 0, 83, 52, 27, 1, 73]
    assert r.shuffle_arange(196)[:8] == [74, 164, 111, 62, 70, 47, 23, 124]
    r = MT19937(61)
    assert r.shuffle_arange(1024)[:8] == [994, 629, 774, 258, 702, 535, 658, 1002]
    assert num_masked_patches(224, 16, 0.65) == 127
    assert num_masked_patches(512, 16, 0.65) == 665
    assert num_masked_patches(1024, 16, 0.65) == 2662
    assert num_masked_patches(512, 16, 0.0) == 0


@pytest.mark.parametrize('case', GOLD['cases'], ids=lambda c: f"s{c['seed']}_B{c['B']}_S{c['S']}_p{c['patch_size']}_r{c['mask_ratio']}")
def test_golden_masks(case):
    r = MT19937(case['seed'])
    for st in case['steps']:
        m, _ = patch_mask(r, case['B'], case['S'], case['patch_size'], case['mask_ratio'])
        assert sha16(m) == st['online_sha16']
        assert int(m[0].sum()) == st['online_sum_per_image']
        if 'target_sum' in st:
            mt, _ = patch_mask(r, case['B'], case['S'], case['patch_size'], 0.0)   # Q2: target still shuffles
            assert int(mt.sum()) == st['target_sum'] == 0
    if 'next_u32' in case:   # stream position after the two steps
        assert [r.interval(2 ** 32 - 1) for _ in range(4)] == case['next_u32']
