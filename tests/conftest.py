import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box via gpurun)')


def load_golden(name):
    """A fixture minted by oracle/gen_golden.py from the unmodified reference; inputs and
    weights that were not stored are regenerated from the recorded seeds."""
    import torch
    from openglue_b200.synthetic import synthetic_pairs, synthetic_state_dict
    fx = torch.load(os.path.join(GOLDEN_DIR, f'{name}.pt'), weights_only=False)
    cfg = fx['config']
    if 'state_dict' not in fx:
        fx['state_dict'] = synthetic_state_dict(cfg, seed=fx['weights_seed'])
    if 'data' not in fx:
        fx['data'] = synthetic_pairs(fx['batch'], fx['n'], fx['m'], cfg['descriptor_dim'],
                                     cfg['positional_encoding']['side_info_size'],
                                     family=fx['family'], seed=fx['inputs_seed'])
    return fx


GOLDEN_FULL = ['tiny_flat', 'tiny_planted', 'tiny_offset_s6', 'small_planted']
GOLDEN_SAMPLED = ['C1_planted', 'C1_flat']
# BASELINE.json configs[1], [2] (headline), [4] at full depth and at the batch bench.py times (oracle/gen_golden.py)
GOLDEN_BIG = ['C2_planted', 'C3_planted', 'C3_flat', 'C5_planted']


@pytest.fixture(scope='session')
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]
    return get
