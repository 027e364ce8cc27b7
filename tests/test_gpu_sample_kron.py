"""GPU tests (through the C ABI) of the structured GP-prior sampler mpb_sample_gp_kron (csrc/sample_gp_kron.cu).

The sampler drops the entries of scale_tril that couple different dofs after mpb_sample_gp_kron_pack has verified
bit-exactly that they are 0.0f.  Dropping exact zeros from an fp32 fma chain over ascending k changes nothing, so the
result must be BIT-IDENTICAL to the dense FP32 sampler mpb_sample_gp (which is itself checked against the golden
vectors of the unmodified reference's MultiMPPrior.sample, mp_priors_multi.py:253-256) -- and within 1e-5 of fp64.
"""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

from motion_planning_baselines_b200 import configs  # noqa: E402


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    return dict(device=torch.device('cuda:0'), dtype=torch.float32)


def structured_factor(H, dof, gen, dev, scale=0.05):
    """A lower-triangular [M,M] factor in the state order (t,[pos|vel],j) that couples only equal dofs."""
    N = 2 * H
    L6 = torch.zeros(H, 2, dof, H, 2, dof, **dev)
    for j in range(dof):
        L6[:, :, j, :, :, j] = (torch.tril(torch.randn(N, N, generator=gen, **dev)) * scale).view(H, 2, H, 2)
    return L6.reshape(N * dof, N * dof).contiguous()


def pack(L, H, dof, dev):
    from motion_planning_baselines_b200 import _lib
    out = torch.full((dof, 2 * H, 2 * H), float('nan'), **dev)
    ok = C.c_int(-1)
    _lib.check(_lib.lib().mpb_sample_gp_kron_pack(_lib.ptr(L), _lib.ptr(out), H, dof, C.byref(ok), _lib.stream_ptr()))
    return out, ok.value


def sample_both(L, LkT, mu, eps, H, dof, dev):
    from motion_planning_baselines_b200 import _lib
    S, P, M = eps.shape
    lib = _lib.lib()
    xk = torch.full((P, S, M), float('nan'), **dev)
    xd = torch.full((P, S, M), float('nan'), **dev)
    _lib.check(lib.mpb_sample_gp_kron(_lib.ptr(LkT), _lib.ptr(mu), _lib.ptr(eps), _lib.ptr(xk), P, S, H, dof, _lib.stream_ptr()))
    _lib.check(lib.mpb_sample_gp(_lib.ptr(L), _lib.ptr(mu), _lib.ptr(eps), _lib.ptr(xd), P, S, M, _lib.stream_ptr()))
    torch.cuda.synchronize()
    return xk, xd


@pytest.mark.parametrize('dof,H,P,S', [(2, 32, 3, 7), (2, 64, 1, 64), (2, 128, 5, 13), (3, 32, 4, 8), (3, 64, 7, 33),
                                       (3, 128, 2, 31), (7, 32, 3, 11), (7, 64, 5, 13), (7, 64, 16, 64), (7, 64, 1, 1),
                                       (2, 16, 4, 9), (3, 48, 2, 5), (4, 32, 3, 6), (4, 64, 2, 9), (5, 64, 3, 5), (6, 32, 2, 8),
                                       (6, 64, 2, 5), (7, 16, 3, 7), (7, 48, 2, 6), (8, 32, 2, 7), (8, 64, 2, 5)])
def test_kron_bit_identical_to_dense_fp32(dof, H, P, S, dev):
    from motion_planning_baselines_b200 import _lib
    assert _lib.lib().mpb_sample_gp_kron_supported(H, dof)
    gen = torch.Generator(device='cuda').manual_seed(100 * dof + H + P)
    M = 2 * H * dof
    L = structured_factor(H, dof, gen, dev)
    LkT, ok = pack(L, H, dof, dev)
    assert ok == 1
    # the packed blocks are the per-dof blocks, k-major
    L6 = L.view(H, 2, dof, H, 2, dof)
    for j in range(dof):
        assert torch.equal(LkT[j], L6[:, :, j, :, :, j].reshape(2 * H, 2 * H).t())
    mu = torch.randn(P, M, generator=gen, **dev)
    eps = torch.randn(S, P, M, generator=gen, **dev)
    xk, xd = sample_both(L, LkT, mu, eps, H, dof, dev)
    assert not torch.isnan(xk).any()
    assert torch.equal(xk, xd), f'max |diff| {float((xk - xd).abs().max()):.3e}'
    ref = mu.double().unsqueeze(1) + torch.einsum('ik,spk->psi', L.double(), eps.double())
    err = (xk.double() - ref).abs().max()
    assert float(err) <= 1e-5 * float(ref.abs().max()), float(err)
    # zero noise returns the means exactly
    xz, _ = sample_both(L, LkT, mu, torch.zeros_like(eps), H, dof, dev)
    assert torch.equal(xz, mu.unsqueeze(1).expand(P, S, M))


def test_pack_rejects_coupled_or_upper_factors(dev):
    gen = torch.Generator(device='cuda').manual_seed(5)
    H, dof = 32, 3
    L = structured_factor(H, dof, gen, dev)
    assert pack(L, H, dof, dev)[1] == 1
    Lc = L.clone()
    Lc[40, 3] = 1e-30            # row (t=6, pos, j=4%3...) couples different dofs: (40 % 6) % 3 = 1 vs 3 % 3 = 0
    assert (40 % 6) % 3 != (3 % 6) % 3
    assert pack(Lc, H, dof, dev)[1] == 0
    Lu = L.clone()
    Lu[0, 6] = 0.5               # same dof (0), but above the diagonal
    assert pack(Lu, H, dof, dev)[1] == 0
    dense = torch.tril(torch.randn(2 * H * dof, 2 * H * dof, generator=gen, **dev))
    assert pack(dense, H, dof, dev)[1] == 0


@pytest.mark.parametrize('name', ['C3', 'C4'])
def test_reference_prior_is_structured_and_sampler_matches(name, dev):
    """The factor torch builds from the reference's precision (factors.MultiMPPrior, the reference's own routine)
    has the exact zero pattern; MultiMPPrior.sample then runs the structured kernel and equals the dense FP32 one."""
    from motion_planning_baselines_b200 import _lib
    from motion_planning_baselines_b200.factors import GPFactor, MultiMPPrior, UnaryFactor
    cfg = configs.config(name)
    prm, H, dof = cfg['params'], cfg['H'], cfg['robot'].q_dim
    D = 2 * dof
    start = torch.cat((torch.tensor(cfg['start']), torch.zeros(dof))).to(**dev)
    goal = torch.cat((torch.tensor(cfg['goal']), torch.zeros(dof))).to(**dev).unsqueeze(0)
    prior = MultiMPPrior(H - 1, cfg['dt'], D, dof,
                         UnaryFactor(D, prm['sigma_start_sample'], tensor_args=dev).K,
                         GPFactor(dof, prm['sigma_gp_sample'], cfg['dt'], H - 1, tensor_args=dev).Q_inv[0],
                         start, K_g_inv=UnaryFactor(D, prm['sigma_goal_sample'], tensor_args=dev).K, goal_states=goal,
                         tensor_args=dev)
    assert prior.scale_tril_kron is not None, 'the reference prior factor must decouple over the dofs'
    P, S, M = 6, 40, H * D
    gen = torch.Generator(device='cuda').manual_seed(11)
    prior.means = torch.randn(P, M, generator=gen, **dev)
    prior.num_modes = P
    eps = torch.randn(S, P, M, generator=gen, **dev)
    x_tc = prior.sample(S, eps=eps).reshape(P, S, M).clone()       # default: tensor-core variant
    assert prior.scale_tril_kron_tc is not None
    kind, prior.kron_tc_kind = prior.kron_tc_kind, 0
    assert kind == 1, 'the default sampler of a structured prior is the warp-MMA tensor-core variant'
    x = prior.sample(S, eps=eps).reshape(P, S, M)                    # exact FP32 variant
    prior.kron_tc_kind = kind
    xd = torch.empty(P, S, M, **dev)
    _lib.check(_lib.lib().mpb_sample_gp(_lib.ptr(prior.scale_tril), _lib.ptr(prior.means), _lib.ptr(eps), _lib.ptr(xd),
                                        P, S, M, _lib.stream_ptr()))
    assert torch.equal(x, xd)
    ref = prior.means.double().unsqueeze(1) + torch.einsum('ik,spk->psi', prior.scale_tril.double(), eps.double())
    noise = (ref - prior.means.double().unsqueeze(1)).abs().max()
    assert float((x.double() - ref).abs().max()) <= 2e-6 * float(noise) + 1e-6
    assert float((x_tc.double() - ref).abs().max()) <= 5e-6 * float(noise) + 1e-6


def test_kron_full_size_properties(dev):
    """C4 shape (512 x 64 x 896): linearity in eps and agreement with the dense sampler on the whole batch."""
    from motion_planning_baselines_b200 import _lib
    dof, H, P, S = 7, 64, 512, 64
    M = 2 * H * dof
    gen = torch.Generator(device='cuda').manual_seed(3)
    L = structured_factor(H, dof, gen, dev, scale=0.03)
    LkT, ok = pack(L, H, dof, dev)
    assert ok == 1
    mu = torch.randn(P, M, generator=gen, **dev)
    eps = torch.randn(S, P, M, generator=gen, **dev)
    xk, xd = sample_both(L, LkT, mu, eps, H, dof, dev)
    assert torch.equal(xk, xd)
    x2 = torch.empty_like(xk)
    _lib.check(_lib.lib().mpb_sample_gp_kron(_lib.ptr(LkT), _lib.ptr(torch.zeros_like(mu)), _lib.ptr(2 * eps), _lib.ptr(x2),
                                             P, S, H, dof, _lib.stream_ptr()))
    x0 = torch.empty_like(xk)
    _lib.check(_lib.lib().mpb_sample_gp_kron(_lib.ptr(LkT), _lib.ptr(torch.zeros_like(mu)), _lib.ptr(eps), _lib.ptr(x0),
                                             P, S, H, dof, _lib.stream_ptr()))
    assert torch.equal(x2, 2 * x0), 'scaling the noise by 2 is exact in fp32'


@pytest.mark.parametrize('dof,H,P,S', [(2, 32, 3, 7), (2, 64, 1, 64), (2, 128, 5, 13), (3, 32, 4, 8), (3, 64, 7, 33),
                                       (3, 128, 2, 31), (7, 32, 3, 11), (7, 64, 5, 13), (7, 64, 16, 64), (7, 64, 1, 1),
                                       (2, 48, 4, 9), (3, 16, 2, 5), (4, 64, 3, 6), (5, 32, 2, 9), (6, 64, 3, 5), (7, 16, 3, 7),
                                       (7, 48, 2, 6), (8, 32, 2, 7), (8, 64, 2, 33)])
def test_kron_tc_matches_fp64(dof, H, P, S, dev):
    """Tensor-core variant (warp MMA, two-term fp16 split): within 5e-6 of the noise amplitude of an fp64 product, zero noise
    returns the means exactly, and rows past the ragged end are untouched."""
    from motion_planning_baselines_b200 import _lib
    gen = torch.Generator(device='cuda').manual_seed(7 * dof + H + S)
    M = 2 * H * dof
    L = structured_factor(H, dof, gen, dev)
    LkT, ok = pack(L, H, dof, dev)
    assert ok == 1
    mu = torch.randn(P, M, generator=gen, **dev)
    eps = torch.randn(S, P, M, generator=gen, **dev)
    lib = _lib.lib()
    LkF = torch.empty(lib.mpb_sample_gp_kron_tc_bytes(H, dof), device=dev['device'], dtype=torch.uint8)
    _lib.check(lib.mpb_sample_gp_kron_tc_prepare(_lib.ptr(LkT), _lib.ptr(LkF), H, dof, _lib.stream_ptr()))
    x = torch.full((P * S + 3, M), float('nan'), **dev)
    _lib.check(lib.mpb_sample_gp_kron_tc(_lib.ptr(LkF), _lib.ptr(mu), _lib.ptr(eps), _lib.ptr(x), P, S, H, dof, _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.isnan(x[P * S:]).all(), 'wrote past the last row'
    x = x[:P * S].view(P, S, M)
    assert not torch.isnan(x).any()
    noise = torch.einsum('ik,spk->psi', L.double(), eps.double())
    err = (x.double() - (mu.double().unsqueeze(1) + noise)).abs().max()
    assert float(err) <= 5e-6 * float(noise.abs().max()), (float(err), float(noise.abs().max()))
    xz = torch.empty(P, S, M, **dev)
    _lib.check(lib.mpb_sample_gp_kron_tc(_lib.ptr(LkF), _lib.ptr(mu), _lib.ptr(torch.zeros_like(eps)), _lib.ptr(xz), P, S, H, dof, _lib.stream_ptr()))
    assert torch.equal(xz, mu.unsqueeze(1).expand(P, S, M))


@pytest.mark.parametrize('name,P', [('C3', 37), ('C4', 512), ('C4', 3)])
def test_structured_prior_matvec_bit_identical(name, P, dev):
    """mpb_prior_matvec_dof (7 non-zeros per row) == mpb_prior_matvec (55-entry band walk), bit for bit, on the
    reference's precision matrix; a precision that couples dofs is rejected by mpb_prior_dof_structured."""
    from motion_planning_baselines_b200 import _lib
    from motion_planning_baselines_b200.factors import GPFactor, MultiMPPrior, UnaryFactor
    cfg = configs.config(name)
    prm, H, dof = cfg['params'], cfg['H'], cfg['robot'].q_dim
    D, M = 2 * dof, 2 * dof * H
    start = torch.cat((torch.tensor(cfg['start']), torch.zeros(dof))).to(**dev)
    goal = torch.cat((torch.tensor(cfg['goal']), torch.zeros(dof))).to(**dev).unsqueeze(0)
    prior = MultiMPPrior(H - 1, cfg['dt'], D, dof,
                         UnaryFactor(D, prm['sigma_start_sample'], tensor_args=dev).K,
                         GPFactor(dof, prm['sigma_gp_sample'], cfg['dt'], H - 1, tensor_args=dev).Q_inv[0],
                         start, K_g_inv=UnaryFactor(D, prm['sigma_goal_sample'], tensor_args=dev).K, goal_states=goal,
                         tensor_args=dev)
    lib = _lib.lib()
    ok = C.c_int(-1)
    _lib.check(lib.mpb_prior_dof_structured(_lib.ptr(prior.Sigma_inv), H, dof, C.byref(ok), _lib.stream_ptr()))
    assert ok.value == 1
    gen = torch.Generator(device='cuda').manual_seed(P)
    mu = torch.randn(P, M, generator=gen, **dev)
    y_dof = torch.full((P, M), float('nan'), **dev)
    y_band = torch.full((P, M), float('nan'), **dev)
    _lib.check(lib.mpb_prior_matvec_dof(_lib.ptr(prior.Sigma_inv), _lib.ptr(mu), _lib.ptr(y_dof), P, H, dof, _lib.stream_ptr()))
    _lib.check(lib.mpb_prior_matvec(_lib.ptr(prior.Sigma_inv), _lib.ptr(mu), _lib.ptr(y_band), P, M, 2 * D - 1, _lib.stream_ptr()))
    assert torch.equal(y_dof, y_band)
    ref = mu.double() @ prior.Sigma_inv.double().t()
    assert float((y_dof.double() - ref).abs().max()) <= 1e-6 * float(ref.abs().max())
    bad = prior.Sigma_inv.clone()
    bad[D + 1, 0] = 1.0          # (t=1, pos, j=1) x (t=0, pos, j=0): couples two dofs
    _lib.check(lib.mpb_prior_dof_structured(_lib.ptr(bad), H, dof, C.byref(ok), _lib.stream_ptr()))
    assert ok.value == 0


@pytest.mark.parametrize('dof,H,P,S', [(2, 32, 3, 7), (2, 64, 1, 64), (2, 128, 5, 13), (3, 32, 4, 8), (3, 64, 7, 33), (3, 16, 9, 5),
                                       (3, 128, 2, 31), (7, 32, 3, 11), (7, 64, 5, 13), (7, 64, 16, 64), (7, 64, 1, 1), (7, 64, 37, 23)])
def test_kron_umma_matches_fp64(dof, H, P, S, dev):
    """tcgen05 variant (3xTF32, per-dof M128xN32 MMA chains): within 5e-6 of the noise amplitude of an fp64 product, zero
    noise returns the means exactly, rows past the ragged end are untouched."""
    from motion_planning_baselines_b200 import _lib
    lib = _lib.lib()
    assert lib.mpb_sample_gp_kron_umma_supported(H, dof)
    gen = torch.Generator(device='cuda').manual_seed(13 * dof + H + S)
    M = 2 * H * dof
    L = structured_factor(H, dof, gen, dev)
    LkT, ok = pack(L, H, dof, dev)              # packing has no shape restriction
    assert ok == 1
    Lp = torch.empty(lib.mpb_sample_gp_kron_umma_floats(H, dof), **dev)
    _lib.check(lib.mpb_sample_gp_kron_umma_prepare(_lib.ptr(LkT), _lib.ptr(Lp), H, dof, _lib.stream_ptr()))
    mu = torch.randn(P, M, generator=gen, **dev)
    eps = torch.randn(S, P, M, generator=gen, **dev)
    x = torch.full((P * S + 3, M), float('nan'), **dev)
    _lib.check(lib.mpb_sample_gp_kron_umma(_lib.ptr(Lp), _lib.ptr(mu), _lib.ptr(eps), _lib.ptr(x), P, S, H, dof, _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.isnan(x[P * S:]).all(), 'wrote past the last row'
    x = x[:P * S].view(P, S, M)
    assert not torch.isnan(x).any()
    noise = torch.einsum('ik,spk->psi', L.double(), eps.double())
    err = (x.double() - (mu.double().unsqueeze(1) + noise)).abs().max()
    assert float(err) <= 5e-6 * float(noise.abs().max()), (float(err), float(noise.abs().max()))
    xz = torch.empty(P, S, M, **dev)
    _lib.check(lib.mpb_sample_gp_kron_umma(_lib.ptr(Lp), _lib.ptr(mu), _lib.ptr(torch.zeros_like(eps)), _lib.ptr(xz), P, S, H, dof, _lib.stream_ptr()))
    assert torch.equal(xz, mu.unsqueeze(1).expand(P, S, M))


def test_structured_entry_points_reject_bad_arguments(dev):
    """Error behaviour at the C boundary: unsupported shapes, null / misaligned pointers -> MPB_EINVAL with a message;
    empty batches are a no-op."""
    from motion_planning_baselines_b200 import _lib
    lib = _lib.lib()
    assert lib.mpb_sample_gp_kron_supported(64, 7) == 1 and lib.mpb_sample_gp_kron_supported(48, 7) == 1
    assert lib.mpb_sample_gp_kron_supported(64, 5) == 1 and lib.mpb_sample_gp_kron_supported(64, 8) == 1
    assert lib.mpb_sample_gp_kron_supported(40, 7) == 0 and lib.mpb_sample_gp_kron_supported(128, 7) == 0      # H % 16, dof * H <= 512
    assert lib.mpb_sample_gp_kron_supported(64, 9) == 0 and lib.mpb_sample_gp_kron_umma_supported(24, 7) == 0
    H, dof, P, S = 64, 7, 2, 4
    M = 2 * H * dof
    buf = torch.zeros(dof * 4 * H * H + 64, **dev)
    mu, eps, x = torch.zeros(P, M, **dev), torch.zeros(S, P, M, **dev), torch.full((P, S, M), 7.0, **dev)
    st = _lib.stream_ptr()
    for fn in (lib.mpb_sample_gp_kron, lib.mpb_sample_gp_kron_tc, lib.mpb_sample_gp_kron_umma):
        assert fn(_lib.ptr(buf), _lib.ptr(mu), _lib.ptr(eps), _lib.ptr(x), P, S, 40, dof, st) != 0        # unsupported H
        assert b'H=40' in lib.mpb_last_error()
        assert fn(None, _lib.ptr(mu), _lib.ptr(eps), _lib.ptr(x), P, S, H, dof, st) != 0                  # null factor
        assert fn(_lib.ptr(buf), _lib.ptr(mu), _lib.ptr(eps), _lib.ptr(x), -1, S, H, dof, st) != 0        # negative size
        assert fn(_lib.ptr(buf), _lib.ptr(mu), eps.data_ptr() + 4, _lib.ptr(x), P, S, H, dof, st) != 0    # misaligned eps
        assert fn(_lib.ptr(buf), _lib.ptr(mu), _lib.ptr(eps), _lib.ptr(x), 0, S, H, dof, st) == 0         # empty batch
    torch.cuda.synchronize()
    assert bool((x == 7.0).all()), 'a rejected or empty call must not write'
    ok = C.c_int(5)
    assert lib.mpb_sample_gp_kron_pack(None, _lib.ptr(buf), H, dof, C.byref(ok), st) != 0
    assert lib.mpb_prior_dof_structured(None, H, dof, C.byref(ok), st) != 0
    assert lib.mpb_prior_matvec_dof(_lib.ptr(buf), _lib.ptr(mu), None, P, H, dof, st) != 0
    assert lib.mpb_sample_gp_kron_tc_prepare(_lib.ptr(buf), _lib.ptr(buf), 40, dof, st) != 0
    assert lib.mpb_sample_gp_kron_umma_prepare(_lib.ptr(buf), _lib.ptr(buf), 24, dof, st) != 0


def test_planner_samplers_agree(dev, monkeypatch):
    """StochGPMP at a C4-like shape: one iteration with the default (warp-MMA), the exact-FP32, the tcgen05 and the dense
    samplers on identical injected noise gives the same costs (1e-5) and the same argmin sample per particle."""
    from motion_planning_baselines_b200.fields import CollisionField
    from motion_planning_baselines_b200.planners import StochGPMP
    from motion_planning_baselines_b200.robots import Robot
    cfg = configs.config('C4')
    P, S, H = 8, 32, 64
    gen = torch.Generator(device='cuda').manual_seed(77)
    eps = torch.randn(S, P, H * 14, generator=gen, **dev)
    out = {}
    for mode in ('simt', 'kron', 'kron_fp32', 'kron_umma'):
        monkeypatch.setenv('MPB_SAMPLE_GP', mode)
        torch.manual_seed(5)
        robot = Robot(cfg['robot'], dt=cfg['dt'], tensor_args=dev)
        field = CollisionField(cfg['obstacles'], tensor_args=dev)
        planner = StochGPMP(robot=robot, n_dof=7, n_support_points=H, num_particles_per_goal=P, opt_iters=1, dt=cfg['dt'],
                            start_state=torch.tensor(cfg['start']).to(**dev),
                            multi_goal_states=torch.tensor(cfg['goal']).to(**dev).unsqueeze(0),
                            collision_fields=[field], tensor_args=dev, num_samples=S, **cfg['params'])
        kinds = dict(kron=1, kron_fp32=0, kron_umma=2, simt=0)
        assert planner._sample_dist.kron_tc_kind == kinds[mode]
        assert (planner._sample_dist.scale_tril_kron is not None) == (mode != 'simt')
        if mode == 'simt':
            means0 = planner._particle_means.clone()
        else:
            planner._particle_means.copy_(means0)
        traj = planner.optimize(opt_iters=1, eps=[eps])
        out[mode] = (planner.costs.clone(), traj.clone(), planner.state_samples.clone())
    ref_c, ref_t, ref_x = out['simt']
    for mode in ('kron', 'kron_fp32', 'kron_umma'):
        c, t, xs = out[mode]
        assert float((xs - ref_x).abs().max()) <= 2e-6, mode
        assert float(((c - ref_c).abs() / ref_c.abs().clamp_min(1e-6)).max()) <= 1e-5, mode
        assert torch.equal(c.argmin(1), ref_c.argmin(1)), mode
        assert float((t - ref_t).abs().max()) <= 1e-5 * float(ref_t.abs().max()), mode
    assert torch.equal(out['kron_fp32'][2], ref_x), 'the exact-FP32 structured sampler is bit-identical to the dense FP32 one'
