"""K1 experiment: factor tiles staged in tensor memory (MPB_KRON_GEN_DBG=512) vs read from shared memory by every MMA.
Compares the samples bit for bit and times both.  usage: k1_tmem_a.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from motion_planning_baselines_b200 import _lib
dev = dict(device=torch.device('cuda:0'), dtype=torch.float32)
dof, H, P, S = 7, 64, 512, 64
N, M = 2 * H, 2 * H * dof
gen = torch.Generator(device='cuda').manual_seed(0)
LkT = torch.zeros(dof, N, N, **dev)
for j in range(dof):
    LkT[j] = (torch.tril(torch.randn(N, N, generator=gen, **dev)) * 0.03).t()
lib = _lib.lib()
mu = torch.randn(P, M, generator=gen, **dev)
Limg = torch.empty(lib.mpb_sample_gp_kron_gen_bytes(H, dof), device=dev['device'], dtype=torch.uint8)
_lib.check(lib.mpb_sample_gp_kron_gen_prepare(_lib.ptr(LkT), _lib.ptr(Limg), H, dof, _lib.stream_ptr()))
out = {}
for dbg in (0, 512, 512 + 1024):
    os.environ['MPB_KRON_GEN_DBG'] = str(dbg)
    x = torch.empty(P, S, M, **dev)
    def run(i):
        nd = _lib.NoiseDesc(seed=1, offset=i, s_offset=0, p_offset=0, P_global=P)
        _lib.check(lib.mpb_sample_gp_kron_gen(_lib.ptr(Limg), _lib.ptr(mu), C.byref(nd), _lib.ptr(x), P, S, H, dof, _lib.stream_ptr()))
    run(0); torch.cuda.synchronize()
    out[dbg] = x.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30): run(0)
    e1.record(); torch.cuda.synchronize()
    print(f'dbg={dbg}: {e0.elapsed_time(e1) / 30:.4f} ms/launch')
for k in (512, 1536):
    print(k, 'max |difference|', (out[0] - out[k]).abs().max().item(), 'bit-identical', torch.equal(out[0], out[k]))
