"""Times mpb_sample_gp_kron_gen alone at the C4 shape, with the stage-disable hooks (MPB_KRON_GEN_DBG bit mask:
1 no Philox, 2 no MMAs, 4 no output stores, 8 no factor loads).  Usage: python profiles/tools/k1_gen_bench.py [reps]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from motion_planning_baselines_b200 import _lib
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = dict(device=torch.device('cuda:0'), dtype=torch.float32)
dof, H, P, S = 7, 64, 512, 64
N, M = 2 * H, 2 * H * dof
gen = torch.Generator(device='cuda').manual_seed(0)
LkT = torch.zeros(dof, N, N, **dev)
for j in range(dof):
    LkT[j] = (torch.tril(torch.randn(N, N, generator=gen, **dev)) * 0.03).t()
lib = _lib.lib()
mu = torch.randn(P, M, generator=gen, **dev)
x = torch.empty(P, S, M, **dev)
Limg = torch.empty(lib.mpb_sample_gp_kron_gen_bytes(H, dof), device=dev['device'], dtype=torch.uint8)
_lib.check(lib.mpb_sample_gp_kron_gen_prepare(_lib.ptr(LkT), _lib.ptr(Limg), H, dof, _lib.stream_ptr()))
for dbg in [0, 1, 4 + 128 + 256]:
    os.environ['MPB_KRON_GEN_DBG'] = str(dbg)
    def run(i):
        nd = _lib.NoiseDesc(seed=1, offset=i, s_offset=0, p_offset=0, P_global=P)
        _lib.check(lib.mpb_sample_gp_kron_gen(_lib.ptr(Limg), _lib.ptr(mu), C.byref(nd), _lib.ptr(x), P, S, H, dof, _lib.stream_ptr()))
    for i in range(3): run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): run(i)
    e1.record(); torch.cuda.synchronize()
    off = [n for b, n in ((1, 'Philox'), (2, 'MMA'), (4, 'stores'), (8, 'factor loads'), (16, 'epilogue fence'), (32, 'producer fence'), (64, 'epilogue'), (128, 'epilogue TMEM loads'), (256, 'epilogue smem writes')) if dbg & b]
    print(f'dbg={dbg:2d} without [{", ".join(off) or "-"}]: {e0.elapsed_time(e1) / reps:.4f} ms/launch')

# timeline of CTA 0 (clock64 stamps through the MPB_KRON_GEN_TRACE debug hook)
names = ['MMA chunk0 ready', 'MMA tile committed', 'epilogue start', 'epilogue end', 'producer chunk0 written', 'producer chunk7 written', 'MMA got acc_empty']
for dbg in [0, 1, 4 + 128 + 256]:
    os.environ['MPB_KRON_GEN_DBG'] = str(dbg)
    tr = torch.zeros(64, dtype=torch.int64, device=dev['device'])
    os.environ['MPB_KRON_GEN_TRACE'] = str(tr.data_ptr())
    nd = _lib.NoiseDesc(seed=1, offset=0, s_offset=0, p_offset=0, P_global=P)
    _lib.check(lib.mpb_sample_gp_kron_gen(_lib.ptr(Limg), _lib.ptr(mu), C.byref(nd), _lib.ptr(x), P, S, H, dof, _lib.stream_ptr()))
    torch.cuda.synchronize()
    del os.environ['MPB_KRON_GEN_TRACE']
    t = tr.cpu().view(8, 8)
    t0 = int(t[t > 0].min())
    nt = 4 if os.environ.get('MPB_KRON_GEN_TS') == '64' else 8
    print(f'dbg={dbg}: per tile (cycles): ' + ' | '.join(
        f'k-loop {int(t[o, 1] - t[o, 0])}, production c0->c7 {int(t[o, 5] - t[o, 4])}, epilogue {int(t[o, 3] - t[o, 2])}, tile {int(t[o, 3] - t[o, 6])}' for o in range(nt)))
    print('   absolute (cycles since the first stamp): ' + ' | '.join(
        f'mma {int(t[o, 0] - t0)}-{int(t[o, 1] - t0)} epi {int(t[o, 2] - t0)}-{int(t[o, 3] - t0)} prod {int(t[o, 4] - t0)}-{int(t[o, 5] - t0)}' for o in range(nt)))
    continue
    f = tr.cpu()[40:56].view(2, 8)
    for i in range(2):
        v = [int(x) for x in f[i]]
        print('   epilogue batch %d of tile 1 (cycles): wait_read %d, bar %d, tmem ld+wait %d, ffma+sts %d, fence %d, bar %d, store+commit %d' % (4 + i, v[1]-v[0], v[2]-v[1], v[3]-v[2], v[4]-v[3], v[5]-v[4], v[6]-v[5], v[7]-v[6]))
    print('   batch 4 start -> batch 5 start', int(f[1][0] - f[0][0]))
