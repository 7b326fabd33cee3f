"""CPU-only checks of the host side: state_dict contract, config translation, the C-ABI
library (loads, exports every declared symbol, argument validation without touching a GPU),
weight folding/packing, and that the product never routes through the oracle."""
import os
import re

import pytest
import torch

from conftest import ROOT, GOLDEN_FULL
from openglue_b200 import _cabi
from openglue_b200.packing import pack_weights
from openglue_b200.superglue import SuperGlue
from openglue_b200.synthetic import default_config, synthetic_state_dict
from oracle import superglue_oracle as O
from packed_emulation import forward_packed


def test_library_exports_every_header_symbol():
    header = open(os.path.join(ROOT, 'include', 'openglue_b200.h')).read()
    declared = set(re.findall(r'\b(og_[a-z0-9_]+)\s*\(', header))
    declared -= {'og_status', 'og_config'}
    lib = _cabi.lib()
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    assert lib.og_version() == 100


def test_abi_struct_sizes_match_header():
    # og_config: 5 + 8 ints, int, float, float, int, int = 18 * 4 bytes
    import ctypes as C
    assert C.sizeof(_cabi.OgConfig) == 18 * 4
    assert C.sizeof(_cabi.OgLinearArgs) == 192


def test_argument_validation_without_gpu():
    lib = _cabi.lib()
    cfg = _cabi.make_config(default_config())
    assert lib.og_workspace_bytes(cfg, 0, 10, 10) < 0
    assert b'positive' in lib.og_last_error()
    bad = _cabi.make_config(default_config(descriptor_dim=30, num_heads=4))
    assert lib.og_packed_weight_floats(bad) < 0
    assert lib.og_superglue_forward(cfg, None, None, None, 1, 8, 8, None, None, None, None, None, None, None, None, None, None,
                                    None, None, None, None, None, 0, None) == -1     # OG_EINVAL, no CUDA call made
    assert lib.og_workspace_bytes(cfg, 2, 100, 3000) > 0                             # wide rows: several warps share a row
    assert lib.og_workspace_bytes(cfg, 2, 100, 9000) < 0 and b'8192' in lib.og_last_error()   # the documented limit


def test_state_dict_contract_matches_reference_layout(golden):
    for name in ['tiny_flat', 'tiny_offset_s6']:
        fx = golden(name)
        model = SuperGlue(dict(fx['config']))
        ours = model.state_dict()
        assert list(ours.keys()) == list(fx['state_dict'].keys())
        for k, v in fx['state_dict'].items():
            assert tuple(ours[k].shape) == tuple(v.shape), k
        model.load_state_dict(fx['state_dict'], strict=True)
    full = SuperGlue(default_config())
    assert len(full.state_dict()) == 333
    assert sum(p.numel() for p in full.parameters()) == 11_957_249


def test_unsupported_options_raise():
    cfg = default_config()
    cfg['attention_gnn']['attention'] = 'linear'
    with pytest.raises(ValueError):
        SuperGlue(cfg)
    cfg = default_config()
    cfg['positional_encoding']['encoder_name'] = 'Nope'
    with pytest.raises(NameError):
        SuperGlue(cfg)
    model = SuperGlue(default_config(descriptor_dim=32, num_stages=1)).eval()
    data = {'keypoints0': torch.zeros(1, 4, 2), 'keypoints1': torch.zeros(1, 4, 2)}
    with pytest.raises(RuntimeError, match='no CPU path'):
        model(data)


@pytest.mark.parametrize('name', GOLDEN_FULL)
def test_weight_folding_and_packing(golden, name):
    """packed weights + the kernel schedule (torch emulation, fp64) == oracle fp64."""
    fx = golden(name)
    cfg = _cabi.make_config(fx['config'])
    packed64 = pack_weights(fx['state_dict'], fx['config'], cfg, dtype=torch.float64)
    got = forward_packed(packed64, fx['config'], cfg, fx['data'])
    assert (got['scores'] - fx['scores_f64']).abs().max() < 5e-6     # the reference keeps log_a/log_b/norm in fp32
    ref = O.run(fx['state_dict'], fx['config'], fx['data'], dtype=torch.float64)
    assert (got['context_descriptors0'] - ref['context_descriptors0']).abs().max() < 1e-10
    # and the fp32 packing is the rounding of the fp64 one
    packed = pack_weights(fx['state_dict'], fx['config'], cfg)
    assert packed.dtype == torch.float32
    assert (packed.double() - packed64).abs().max() <= packed64.abs().max() * 2 ** -23


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'openglue_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', text, re.M), f
                assert '/root/reference' not in text, f


def test_tensor_core_kernels_are_blackwell_native_and_address_shared_memory_directly():
    """Static check of the built library (cuobjdump, no GPU): every kernel that issues tcgen05 MMAs (UTCHMMA in SASS) also uses TMEM
    loads / stores and TMA, and none of them contains a generic LD.E / ST.E - the dynamic shared-memory block is aligned as an offset
    from the __shared__ symbol (tc::align_smem_1024); aligning it through uintptr_t hid the address space and cost 7 % of the attention
    kernel (profiles/README.md, finding 7)."""
    import collections
    import re
    import shutil
    import subprocess
    from openglue_b200 import _cabi
    tool = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(tool):
        pytest.skip('cuobjdump not available')
    out = subprocess.run([tool, '-sass', _cabi.LIB_PATH], capture_output=True, text=True).stdout
    hist, cur = {}, None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = hist.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m and cur is not None:
            cur[m.group(1).split('.')[0] + ('.E' if m.group(1).startswith(('LD.E', 'ST.E')) else '')] += 1
    tc = {k: c for k, c in hist.items() if c['UTCHMMA'] > 0}
    assert len(tc) >= 20                                   # GEMM and attention kernels in their fp16 / tf32, paired / single-CTA forms
    for name, c in tc.items():
        assert c['LD.E'] == 0 and c['ST.E'] == 0, (name, c['LD.E'], c['ST.E'])
        assert c['LDTM'] > 0 and c['UTMALDG'] > 0, name    # accumulators read back from TMEM, operands staged by TMA
    fp16_attn = [c for k, c in tc.items() if 'attention_f16t_kernel' in k]
    assert len(fp16_attn) == 4 and all(c['STTM'] > 0 and c['MUFU'] >= 32 for c in fp16_attn)


def test_bench_work_accounting_reproduces_the_survey_table():
    """bench.py's algorithmic FLOP / byte formulas against the values SURVEY.md (section 8d, Appendix B) states for the BASELINE
    configurations - the numerators of `roofline.achieved` and of the judge's own check."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('og_bench', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    rel = lambda a, b: abs(a - b) / b
    # (n, m, d, stages, S, T) -> (F_total, F_attn, Q_sink GB)
    table = {'C1': ((512, 512, 256, 9, 1, 20), (3.418e10, 9.66e9, 0.0221)),
             'C2': ((1024, 1024, 256, 9, 1, 100), (8.796e10, 3.865e10, 0.4245)),
             'C3': ((2048, 2048, 256, 9, 1, 100), (2.543e11, 1.546e11, 1.696)),
             'C5': ((4096, 1024, 128, 18, 6, 50), (3.035e11, 2.416e11, 0.857))}
    for name, ((n, m, d, stages, s, t), (ftot, fattn, qs)) in table.items():
        fl = bench.flops_per_pair(n, m, d, stages, s)
        assert rel(fl['total'], ftot) < 2e-3 and rel(fl['attn'], fattn) < 2e-3, (name, fl)
        assert rel(bench.sinkhorn_bytes_per_pair(n, m, t) / 1e9, qs) < 3e-3, name
        wl = bench.BASELINE_CONFIGS[name]
        assert (wl['n'], wl['m'], wl['cfg']['descriptor_dim'], wl['cfg']['num_stages'], wl['cfg']['num_iters']) == (n, m, d, stages, t)
    assert bench.BASELINE_CONFIGS['C4']['n'] == 2048 and bench.BASELINE_CONFIGS['C4']['cfg']['num_iters'] == 100
