"""K1 dof-major sampler alone at the C4 shape, with stages disabled (MPB_KRON_DM_DBG: 1 no Philox, 2 no MMAs, 4 no stores).
usage: k1_dm_bench.py"""
import os, sys, ctypes as C, subprocess
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for dbg in (0, 64, 127, 63):
        subprocess.run([sys.executable, __file__, str(dbg)], env=dict(os.environ, MPB_KRON_DM_DBG=str(dbg)))
    sys.exit(0)
import torch
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
from motion_planning_baselines_b200 import _lib
from test_gpu_sample_gen import make_prior
dev = dict(device=torch.device('cuda:0'), dtype=torch.float32)
P, S, H, d = int(os.environ.get('K1P', 512)), 64, 64, 7
prior, means = make_prior(P, dev)
lib, st = _lib.lib(), _lib.stream_ptr()
x = torch.empty(P * S, H * 2 * d, **dev)
y = torch.empty(P, H * 2 * d, **dev)
desc = _lib.NoiseDesc(seed=1, offset=0, s_offset=0, p_offset=0, P_global=P)
def run_dm():
    _lib.check(lib.mpb_sample_gp_kron_gen_dm(_lib.ptr(prior.scale_tril_kron_gen), _lib.ptr(prior.means), C.byref(desc), _lib.ptr(x),
                                             P, S, H, d, _lib.ptr(prior.Sigma_inv), _lib.ptr(y), None, st))
def run_nat():
    _lib.check(lib.mpb_sample_gp_kron_gen_mv(_lib.ptr(prior.scale_tril_kron_gen), _lib.ptr(prior.means), C.byref(desc), _lib.ptr(x),
                                             P, S, H, d, _lib.ptr(prior.Sigma_inv), _lib.ptr(y), None, st))
for name, fn in (('dm', run_dm), ('natural', run_nat)):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(f'dbg {sys.argv[1]} {name:8s}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch')
