"""Per-role clock stamps of CTA 0 of the dof-major sampler (MPB_KRON_DM_TRACE).  usage: k1_dm_trace.py [dbg]"""
import os, sys, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
dev = dict(device=torch.device('cuda:0'), dtype=torch.float32)
trace = torch.zeros(512, device='cuda', dtype=torch.int64)
os.environ['MPB_KRON_DM_TRACE'] = hex(trace.data_ptr())
if len(sys.argv) > 1: os.environ['MPB_KRON_DM_DBG'] = sys.argv[1]
from motion_planning_baselines_b200 import _lib
from test_gpu_sample_gen import make_prior
P, S, H, d = 512, 64, 64, 7
prior, means = make_prior(P, dev)
lib, st = _lib.lib(), _lib.stream_ptr()
x = torch.empty(P * S, H * 2 * d, **dev); y = torch.empty(P, H * 2 * d, **dev)
desc = _lib.NoiseDesc(seed=1, offset=0, s_offset=0, p_offset=0, P_global=P)
for _ in range(3):
    _lib.check(lib.mpb_sample_gp_kron_gen_dm(_lib.ptr(prior.scale_tril_kron_gen), _lib.ptr(prior.means), C.byref(desc), _lib.ptr(x),
                                             P, S, H, d, _lib.ptr(prior.Sigma_inv), _lib.ptr(y), None, st))
torch.cuda.synchronize()
t = trace.cpu().view(64, 8)
t0 = int(t[t > 0].min())
print('unit: prod_gen_done prod_got_stage prod_arrived | mma_start mma_committed | epi_start epi_end   (cycles from first stamp)')
for u in range(24):
    r = [int(v) - t0 if v > 0 else -1 for v in t[u]]
    f2 = int(t[u + 32][5]) - t0 if t[u + 32][5] > 0 else -1
    print(f'{u:2d}: {r[0]:7d} {r[6]:7d} {r[1]:7d} | {r[2]:7d} {r[3]:7d} | {r[4]:7d} {r[5]:7d} | before fence {r[7]:7d} after fence {f2:7d}')
