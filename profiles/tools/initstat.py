"""Debug helper: initial-particle statistics and best-sample cost gaps of the C4 'moderate' regime under both K1 samplers."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_gpu_bench_shape import build, MODERATE
dev = dict(device=torch.device('cuda:0'), dtype=torch.float32)
for mode in ('kron', 'auto'):
    os.environ['MPB_SAMPLE_GP'] = mode
    cfg, sig, pl = build('C4', dev, MODERATE)
    m = pl._particle_means
    print(mode, 'gen' if pl._sample_dist.scale_tril_kron_gen is not None else 'mma', 'init means: std over particles (pos, mid waypoint)',
          m[:, 32, :7].std(0).cpu().numpy().round(3), 'vel', m[:, 32, 7:].std(0).cpu().numpy().round(3), 'abs max', float(m.abs().max()))
    print('   start/goal rows std', m[:, 0, :7].std(0).cpu().numpy().round(5), m[:, -1, :7].std(0).cpu().numpy().round(5))
    print('   mean over particles mid', m[:, 32, :7].mean(0).cpu().numpy().round(3), 'roughness', float((m[:, 1:, :7] - m[:, :-1, :7]).abs().mean()))
    gen = torch.Generator(device='cuda').manual_seed(7)
    eps = torch.randn(64, 512, 896, generator=gen, **dev)
    pl.optimize(opt_iters=1, eps=[eps])
    c = pl.costs.sort(1).values
    print('   costs: median best', float(c[:, 0].median()), 'gap best->2nd: median', float((c[:, 1] - c[:, 0]).median()), 'min', float((c[:, 1] - c[:, 0]).min()),
          'max weight < 0.999 in', int((pl._weights.view(512, 64).max(1).values < 0.999).sum()), 'of 512 particles')
