"""CPU: the C-ABI library loads and exports every symbol include/mpb.h declares; the host layer
refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'mpb.h')).read()
    return sorted(set(re.findall(r'\b(mpb_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from motion_planning_baselines_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared()
    assert len(declared) >= 8
    for name in declared:
        assert hasattr(handle, name), f'{name} declared in include/mpb.h but not exported'
    assert sorted(_lib.exported_symbols()) == declared, 'ctypes signature table out of sync with include/mpb.h'
    assert _lib.lib().mpb_version() >= 100


def test_no_cpu_fallback():
    from motion_planning_baselines_b200 import _lib
    from motion_planning_baselines_b200.models import point_mass_model
    from motion_planning_baselines_b200.robots import Robot
    with pytest.raises(_lib.MpbError):
        Robot(point_mass_model(2), tensor_args=dict(device='cpu', dtype=torch.float32))
    with pytest.raises(_lib.MpbError):
        _lib.ptr(torch.zeros(4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'motion_planning_baselines_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, re.M), f'{f} imports the oracle'


def test_models_and_configs():
    from motion_planning_baselines_b200 import configs
    from motion_planning_baselines_b200.models import panda_model
    m = panda_model()
    assert m.n_spheres == 50 and m.q_dim == 7 and m.fixed_tf.shape == (7, 3, 4)
    assert (m.sphere_link[1:] >= m.sphere_link[:-1]).all()
    for name, (P, S) in dict(C1=(1, 64), C2=(1024, 1), C3=(256, 128), C4=(512, 64)).items():
        c = configs.config(name)
        assert (c['P'], c['S'], c['H']) == (P, S, 64)
