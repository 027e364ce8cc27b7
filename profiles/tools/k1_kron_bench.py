"""Times mpb_sample_gp_kron alone at the C4 shape (and the dense samplers beside it). Usage: python profiles/tools/k1_kron_bench.py [reps]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from motion_planning_baselines_b200 import _lib
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = dict(device=torch.device('cuda:0'), dtype=torch.float32)
dof, H, P, S = 7, 64, 512, 64
N, M = 2 * H, 2 * H * dof
gen = torch.Generator(device='cuda').manual_seed(0)
L6 = torch.zeros(H, 2, dof, H, 2, dof, **dev)
for j in range(dof):
    L6[:, :, j, :, :, j] = (torch.tril(torch.randn(N, N, generator=gen, **dev)) * 0.03).view(H, 2, H, 2)
L = L6.reshape(M, M).contiguous()
LkT = torch.empty(dof, N, N, **dev)
ok = C.c_int(0)
lib = _lib.lib()
_lib.check(lib.mpb_sample_gp_kron_pack(_lib.ptr(L), _lib.ptr(LkT), H, dof, C.byref(ok), _lib.stream_ptr()))
assert ok.value == 1
mu = torch.randn(P, M, generator=gen, **dev)
eps = [torch.randn(S, P, M, generator=gen, **dev) for _ in range(4)]
x = torch.empty(P, S, M, **dev)
LkF = torch.empty(lib.mpb_sample_gp_kron_tc_bytes(H, dof), device=dev['device'], dtype=torch.uint8)
_lib.check(lib.mpb_sample_gp_kron_tc_prepare(_lib.ptr(LkT), _lib.ptr(LkF), H, dof, _lib.stream_ptr()))
mode = os.environ.get('KRON', 'tc')
Lp = torch.empty(lib.mpb_sample_gp_kron_umma_floats(H, dof), **dev)
_lib.check(lib.mpb_sample_gp_kron_umma_prepare(_lib.ptr(LkT), _lib.ptr(Lp), H, dof, _lib.stream_ptr()))
fn = {'tc': lib.mpb_sample_gp_kron_tc, 'umma': lib.mpb_sample_gp_kron_umma}.get(mode, lib.mpb_sample_gp_kron)
Lop = {'tc': LkF, 'umma': Lp}.get(mode, LkT)
def run(i):
    _lib.check(fn(_lib.ptr(Lop), _lib.ptr(mu), _lib.ptr(eps[i % 4]), _lib.ptr(x), P, S, H, dof, _lib.stream_ptr()))
for i in range(3): run(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps): run(i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
ref = mu[:2].double().unsqueeze(1) + torch.einsum('ik,spk->psi', L.double(), eps[(reps - 1) % 4][:, :2].double())
print('max err vs fp64', float((x[:2].double() - ref).abs().max()), 'noise amp', float((ref - mu[:2].double().unsqueeze(1)).abs().max()))
print(f'kron[{os.environ.get("KRON", "tc")}]: {ms:.4f} ms/launch  ({2 * M * 4 * P * S / ms / 1e6:.0f} GB/s, {dof * N * (N + 1) * P * S / ms / 1e9:.1f} TFLOP/s structured)')
