import sys, time, torch
sys.path.insert(0, '.')
from motion_planning_baselines_b200 import _lib
lib = _lib.lib()
dev = dict(device='cuda:0', dtype=torch.float32)
def run(P, S, M, check_rows=256):
    g = torch.Generator(device='cuda').manual_seed(P * 1000 + M)
    L = torch.tril(torch.randn(M, M, generator=g, **dev)) / 30
    mu = torch.randn(P, M, generator=g, **dev)
    eps = torch.randn(S, P, M, generator=g, **dev)
    split = torch.empty(2, M, M, **dev)
    _lib.check(lib.mpb_split_tf32(_lib.ptr(L), _lib.ptr(split[0]), _lib.ptr(split[1]), M * M, _lib.stream_ptr()))
    x_tc = torch.full((P, S, M), float('nan'), **dev)
    x_simt = torch.empty(P, S, M, **dev)
    _lib.check(lib.mpb_sample_gp_tc(_lib.ptr(split[0]), _lib.ptr(split[1]), _lib.ptr(mu), _lib.ptr(eps), _lib.ptr(x_tc), P, S, M, _lib.stream_ptr()))
    _lib.check(lib.mpb_sample_gp(_lib.ptr(L), _lib.ptr(mu), _lib.ptr(eps), _lib.ptr(x_simt), P, S, M, _lib.stream_ptr()))
    torch.cuda.synchronize()
    npart = max(1, min(P, check_rows // S))
    ref = mu[:npart].double().unsqueeze(1) + torch.einsum('ik,spk->psi', L.double(), eps[:, :npart].double())
    noise_scale = float((ref - mu[:npart].double().unsqueeze(1)).abs().max())
    e_tc = float((x_tc[:npart].double() - ref).abs().max())
    e_simt = float((x_simt[:npart].double() - ref).abs().max())
    d = float((x_tc - x_simt).abs().max())
    nan = int(torch.isnan(x_tc).sum())
    print(f'P={P} S={S} M={M}: max|tc-fp64|={e_tc:.3e} max|simt-fp64|={e_simt:.3e} max|tc-simt|={d:.3e} (noise scale {noise_scale:.2f}) nan={nan}')
    return e_tc, nan
for shape in [(2, 64, 256), (4, 32, 896), (3, 7, 96), (512, 64, 896), (5, 33, 336), (1, 1, 32), (7, 19, 224)]:
    run(*shape)
# timing at C4
P, S, M = 512, 64, 896
L = torch.tril(torch.randn(M, M, **dev)) / 30; mu = torch.randn(P, M, **dev); eps = torch.randn(S, P, M, **dev)
split = torch.empty(2, M, M, **dev); x = torch.empty(P, S, M, **dev)
_lib.check(lib.mpb_split_tf32(_lib.ptr(L), _lib.ptr(split[0]), _lib.ptr(split[1]), M * M, _lib.stream_ptr()))
for name, fn in [('tc', lambda: lib.mpb_sample_gp_tc(_lib.ptr(split[0]), _lib.ptr(split[1]), _lib.ptr(mu), _lib.ptr(eps), _lib.ptr(x), P, S, M, _lib.stream_ptr())),
                 ('simt', lambda: lib.mpb_sample_gp(_lib.ptr(L), _lib.ptr(mu), _lib.ptr(eps), _lib.ptr(x), P, S, M, _lib.stream_ptr()))]:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, 'ms/launch', e0.elapsed_time(e1) / 20)
