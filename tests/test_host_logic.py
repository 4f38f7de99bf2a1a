"""CPU suite: host-side logic of the drop-in module and the C-ABI library
(load + exported symbols only — no compute without a GPU)."""
import ctypes
import math
import os
import re

import numpy as np
import pytest
import torch

import tpnet_b200
from tpnet_b200 import RandomProjectionModule, _lib
from tpnet_b200.build import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make(**over):
    kw = dict(node_num=37, edge_num=500, dim_factor=3, num_layer=2, time_decay_weight=1e-3, device='cpu',
              use_matrix=False, beginning_time=np.float64(12.5), not_scale=False, enforce_dim=-1)
    kw.update(over)
    return RandomProjectionModule(**kw)


def test_library_builds_loads_and_exports_every_declared_symbol():
    path = build()
    assert os.path.exists(path)
    lib = _lib.load()
    header = open(os.path.join(ROOT, 'include', 'tpnet_b200.h')).read()
    declared = set(re.findall(r'\b(tpn_[a-z_]+)\s*\(', header))
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    raw = ctypes.CDLL(path)
    for sym in declared:
        assert hasattr(raw, sym), sym
    assert lib.tpn_version() == _lib.ABI_VERSION
    assert lib.tpn_error_string(0) == b'ok'
    st = _lib.TpnState(num_layer=3, row_stride=144)
    assert lib.tpn_update_workspace_bytes(ctypes.byref(st), 200) > 0
    assert lib.tpn_update_workspace_bytes(ctypes.byref(st), 100000) > lib.tpn_update_workspace_bytes(ctypes.byref(st), 200)


def test_struct_layout_matches_header(tmp_path):
    """The ctypes mirrors against the header itself: a C program (gcc) prints sizeof / offsetof of every field."""
    import subprocess
    fields = {'tpn_state_t': [f[0] for f in _lib.TpnState._fields_], 'tpn_shard_t': [f[0] for f in _lib.TpnShard._fields_]}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "tpnet_b200.h"', 'int main(void) {']
    for name, fs in fields.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        lines += [f'  printf("{name}.{f} %zu\\n", offsetof({name}, {f}));' for f in fs]
    lines += ['  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, cls in (('tpn_state_t', _lib.TpnState), ('tpn_shard_t', _lib.TpnShard)):
        assert int(out[name]) == ctypes.sizeof(cls), name
        for f in fields[name]:
            assert int(out[f'{name}.{f}']) == getattr(cls, f).offset, (name, f)


def test_argument_validation_without_gpu():
    lib = _lib.load()
    st = _lib.TpnState()           # null data pointer
    assert lib.tpn_pairwise(ctypes.byref(st), None, None, 4, None, 1, None, None) < 0
    assert lib.tpn_gather(ctypes.byref(st), None, 4, None, None) < 0
    assert lib.tpn_materialize(ctypes.byref(st), None) < 0
    assert lib.tpn_pairwise_neighbors(ctypes.byref(st), None, None, None, 4, 5, 1, None, None) < 0
    # the fused head exists for the default 64 -> 256 -> 64 shape only; other shapes are the caller's GEMMs
    assert lib.tpn_head_forward(None, 10, None, 36, 144, None, None, None, None, None, None) == _lib.TPN_ERR_UNSUPPORTED
    assert lib.tpn_head_forward(None, 10, None, 64, 256, None, None, None, None, None, None) < 0      # null pointers
    assert lib.tpn_head_forward(None, 0, None, 64, 256, None, None, None, None, None, None) == 0      # nothing to do


def test_constructor_matches_reference_contract():
    torch.manual_seed(0)
    m = make()
    assert isinstance(m, torch.nn.Module)
    assert m.dim == int(math.log(1000)) * 3 == 18
    assert m.pair_wise_feature_dim == 36 and m.num_layer == 2
    assert len(m.random_projections) == 3
    for p in m.random_projections:
        assert p.shape == (37, 18) and p.dtype == torch.float32 and not p.requires_grad
    assert m.now_time.dtype == torch.float64 and float(m.now_time) == 12.5
    keys = list(m.state_dict().keys())
    assert keys == ['begging_time', 'now_time', 'random_projections.0', 'random_projections.1',
                    'random_projections.2', 'mlp.0.weight', 'mlp.0.bias', 'mlp.2.weight', 'mlp.2.bias']
    assert m.mlp[0].in_features == 36 and m.mlp[0].out_features == 144
    # same RNG consumption as the reference constructor: P_0 = first torch.normal draw
    torch.manual_seed(0)
    expect = torch.normal(0, 1 / math.sqrt(18), (37, 18))
    assert torch.equal(m.random_projections[0].data, expect)
    assert not m.random_projections[1].any() and not m.random_projections[2].any()
    # trainable parameters are exactly the head's
    assert [n for n, p in m.named_parameters() if p.requires_grad] == \
        ['mlp.0.weight', 'mlp.0.bias', 'mlp.2.weight', 'mlp.2.bias']


def test_dim_rules_and_use_matrix():
    assert make(enforce_dim=64).dim == 64
    assert make(node_num=10, dim_factor=10).dim == 10          # capped by node_num
    m = make(node_num=9, use_matrix=True)
    assert m.dim == 9 and torch.equal(m.random_projections[0].data, torch.eye(9))


def test_state_is_packed_node_major_and_survives_to():
    m = make()
    assert m._is_packed()
    assert m.row_stride % 8 == 0 and m.row_stride >= m.dim
    assert m._state.shape == (37, 3, m.row_stride)
    assert not m._state[:, :, m.dim:].any()                    # pad columns are zero
    m2 = m.to('cpu').double().float() if False else m.to('cpu')
    assert m2 is m and m._is_packed()
    # the three parents of the reference share one instance: _apply is called repeatedly
    seq = torch.nn.Sequential(torch.nn.ModuleDict({'a': m}), torch.nn.ModuleDict({'b': m}))
    seq.to('cpu')
    assert m._is_packed()
    m.random_projections[1].data[3, 2] = 5.0                   # writes through the view
    assert m._state[3, 1, 2] == 5.0


def test_state_dict_roundtrip_repacks():
    torch.manual_seed(1)
    a, b = make(), make()
    a.random_projections[1].data.normal_()
    a.now_time.data.fill_(99.0)
    sd = {k: v.clone() for k, v in a.state_dict().items()}
    for k in ('random_projections.0', 'random_projections.1'):
        assert sd[k].shape == (37, 18)
    b.load_state_dict(sd)
    assert b._is_packed() and b._now_host == 99.0
    for i in range(3):
        assert torch.equal(b.random_projections[i].data, a.random_projections[i].data)


def test_compute_without_cuda_fails_loudly():
    m = make()
    ids = np.array([1, 2, 3])
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m.update(ids, ids, np.array([1.0, 2.0, 3.0]))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m.get_pair_wise_feature(ids, ids)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m.get_random_projections(ids)


def test_head_is_the_pytorch_module_whenever_autograd_or_shape_require_it():
    """`_head` = `self.mlp` (TPNet.py:125/:129).  With autograd on (training) and for tensors / shapes the
    fused kernel does not serve it must be the PyTorch module, gradients included."""
    torch.manual_seed(0)
    m = make()
    x = torch.randn(7, m.pair_wise_feature_dim)
    y = m._head(x)
    assert y.requires_grad and torch.equal(y, m.mlp(x))
    y.sum().backward()
    assert m.mlp[0].weight.grad is not None and m.mlp[2].bias.grad is not None
    with torch.no_grad():                       # CPU tensors, F = 36: not the kernel's case -> PyTorch head
        assert torch.equal(m._head(x), m.mlp(x))
        assert m._head(x[:0]).shape == (0, m.pair_wise_feature_dim)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m.get_neighbor_pair_wise_feature(np.ones((3, 4), dtype=np.int64), np.ones(3, dtype=np.int64),
                                         np.ones(3, dtype=np.int64))


def test_product_never_imports_the_oracle():
    pkg = os.path.dirname(tpnet_b200.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(root, f)).read()
                assert 'oracle' not in text.lower().replace('test infrastructure', ''), f
