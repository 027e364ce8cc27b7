"""K1, Blackwell path (csrc/sample_gp_kron_gen.cu: tcgen05 kind::f16 with the factor as the M operand, noise drawn by
warp-specialised Philox producers, bulk-async factor loads and row stores).  Replaces MultiMPPrior.sample
(mp_priors_multi.py:253-256) including the noise draw.  Checked against x = mu + L @ eps in fp64 on the noise the
kernel drew (mpb_philox_normal dump, layout NOISE_SPMD): 1e-5 of the noise amplitude is the north-star bar, the
measured error is ~1e-6; sharding the particles or samples over "ranks" must not change a single bit."""
import numpy as np
import pytest
import torch

from motion_planning_baselines_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    return dict(device=torch.device('cuda:0'), dtype=torch.float32)


def nd(seed, offset, p_off=0, P_glob=1, s_off=0):
    return _lib.NoiseDesc(seed=seed, offset=offset, s_offset=s_off, p_offset=p_off, P_global=P_glob)


def make_prior(P, dev, means=None, seed=0, d=7):
    from motion_planning_baselines_b200.factors import GPFactor, MultiMPPrior, UnaryFactor
    H, dt = 64, 5 / 64
    K = UnaryFactor(2 * d, 1e-3, None, dev).K
    Q = GPFactor(d, 1e-1, dt, H - 1, dev).Q_inv[0]
    if means is None:
        means = torch.randn(P, H, 2 * d, generator=torch.Generator().manual_seed(seed)).to(**dev)
    return MultiMPPrior(H - 1, dt, 2 * d, d, K, Q, torch.zeros(2 * d, **dev), means=means, K_g_inv=K,
                        goal_states=torch.zeros(1, 2 * d, **dev), tensor_args=dev), means


@pytest.mark.parametrize('P,S', [(8, 64), (2, 128), (5, 24), (3, 7), (1, 1), (300, 64)])
def test_gen_sampler_matches_fp64_on_its_own_noise(P, S, dev):
    prior, means = make_prior(P, dev)
    assert prior.scale_tril_kron_gen is not None, 'the tcgen05 sampler must be the default at (H, dof) = (64, 7)'
    assert prior.noise_layout == _lib.NOISE_SPMD
    desc = nd(11, 4, p_off=3, P_glob=P + 5)
    x = prior.sample(S, noise_desc=desc).clone()
    eps = prior.replay_noise(desc, S)                                   # [S,P,M]
    L = prior.scale_tril.double()
    ref = means.view(P, 1, -1).double() + torch.einsum('ik,spk->psi', L, eps.double())
    amp = float((ref - means.view(P, 1, -1).double()).abs().max())
    err = float((x.view(P, S, -1).double() - ref).abs().max()) / amp
    assert err < 5e-6, f'max error {err:.2e} of the noise amplitude'
    # the injected-noise samplers see the same noise and agree with it
    x_inj = prior.sample(S, eps=eps)
    assert float((x_inj.double() - x.double()).abs().max()) / amp < 5e-6


def test_gen_noise_is_standard_normal_and_stream_advances(dev):
    P, S = 16, 64
    prior, means = make_prior(P, dev)
    n = prior.replay_noise(nd(2024, 0, P_glob=P), S).double().flatten()
    N = n.numel()
    assert abs(float(n.mean())) < 5 / np.sqrt(N) and abs(float(n.var()) - 1) < 5 * np.sqrt(2 / N)
    assert abs(float((n ** 3).mean())) < 5 * np.sqrt(15 / N) and abs(float((n ** 4).mean()) - 3) < 5 * np.sqrt(96 / N)
    a = n.view(S * P, -1)
    assert abs(float((a[:, :-1] * a[:, 1:]).mean())) < 5 / np.sqrt(N)        # neighbours inside a row
    assert abs(float((a[:-1] * a[1:]).mean())) < 5 / np.sqrt(N)              # neighbouring rows
    assert abs(float((a[:, :-7] * a[:, 7:]).mean())) < 5 / np.sqrt(N)        # consecutive k of one dof (one Philox call)
    x1, x2 = prior.sample(S).clone(), prior.sample(S).clone()
    assert not torch.equal(x1, x2)
    prior.noise.offset = 0
    assert torch.equal(prior.sample(S), x1)


def test_gen_sampler_is_independent_of_the_sharding(dev):
    P, S = 8, 64
    prior, means = make_prior(P, dev)
    whole = prior.sample(S, noise_desc=nd(7, 2, P_glob=P)).clone()
    halves = []
    for r in range(2):                                           # particles over two "ranks"
        pr, _ = make_prior(P // 2, dev, means=means[r * P // 2:(r + 1) * P // 2])
        halves.append(pr.sample(S, noise_desc=nd(7, 2, p_off=r * P // 2, P_glob=P)).clone())
    assert torch.equal(torch.cat(halves, dim=0), whole)
    parts = [prior.sample(S // 2, noise_desc=nd(7, 2, s_off=r * S // 2, P_glob=P)).clone() for r in range(2)]   # samples over two
    assert torch.equal(torch.cat(parts, dim=1), whole)
    # the dump obeys the same rule
    e = prior.replay_noise(nd(7, 2, P_glob=P), S)
    e2 = torch.cat([_lib.philox_normal(nd(7, 2, s_off=r * S // 2, P_glob=P), _lib.NOISE_SPMD, (S // 2, P, prior.M), dev['device'], dof=7)
                    for r in range(2)], dim=0)
    assert torch.equal(e, e2)


def test_gen_sampler_rejects_bad_arguments(dev):
    lib = _lib.lib()
    assert lib.mpb_sample_gp_kron_gen_supported(64, 7) == 1 and lib.mpb_sample_gp_kron_gen_supported(32, 7) == 0
    assert lib.mpb_sample_gp_kron_gen_supported(64, 2) == 1 and lib.mpb_sample_gp_kron_gen_supported(64, 8) == 0
    prior, means = make_prior(2, dev)
    x = torch.empty(2, 4, prior.M, **dev)
    import ctypes as C
    d = nd(1, 0, P_glob=2)
    assert lib.mpb_sample_gp_kron_gen(None, _lib.ptr(means), C.byref(d), _lib.ptr(x), 2, 4, 64, 7, _lib.stream_ptr()) != 0
    assert lib.mpb_sample_gp_kron_gen(_lib.ptr(prior.scale_tril_kron_gen), _lib.ptr(means), C.byref(d), _lib.ptr(x), 2, 4, 32, 7,
                                      _lib.stream_ptr()) != 0
    assert lib.mpb_sample_gp_kron_gen(_lib.ptr(prior.scale_tril_kron_gen), _lib.ptr(means), C.byref(d), _lib.ptr(x), 0, 4, 64, 7,
                                      _lib.stream_ptr()) == 0


@pytest.mark.parametrize('P,S', [(512, 64), (5, 24), (1, 1), (300, 7)])
def test_gen_sampler_with_prior_matvec_warp(P, S, dev):
    """mpb_sample_gp_kron_gen_mv: the extra warp's y = Sigma^-1 mu (IS term, stoch_gpmp.py:239-241) is bit-identical to
    mpb_prior_matvec_dof, and the samples are bit-identical to the plain launch."""
    import ctypes as C
    prior, means = make_prior(P, dev)
    lib = _lib.lib()
    Sinv = prior.Sigma_inv.contiguous()
    ok = C.c_int(0)
    _lib.check(lib.mpb_prior_dof_structured(_lib.ptr(Sinv), 64, 7, C.byref(ok), _lib.stream_ptr()))
    assert ok.value == 1
    desc = nd(5, 9, P_glob=P)
    x0 = prior.sample(S, noise_desc=desc).clone()
    y_ref = torch.empty(P, 64 * 14, **dev)
    _lib.check(lib.mpb_prior_matvec_dof(_lib.ptr(Sinv), _lib.ptr(means), _lib.ptr(y_ref), P, 64, 7, _lib.stream_ptr()))
    x1 = torch.empty_like(x0)
    y = torch.full((P, 64 * 14), float('nan'), **dev)
    mu_c = torch.full((P, 64 * 14), float('nan'), **dev)
    _lib.check(lib.mpb_sample_gp_kron_gen_mv(_lib.ptr(prior.scale_tril_kron_gen), _lib.ptr(means.contiguous()), C.byref(desc),
                                             _lib.ptr(x1), P, S, 64, 7, _lib.ptr(Sinv), _lib.ptr(y), _lib.ptr(mu_c), _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(x1, x0)
    assert torch.equal(y, y_ref)
    assert torch.equal(mu_c, means.view(P, -1)), 'the copy of the means written by the mat-vec warp'
    with pytest.raises(_lib.MpbError):
        _lib.check(lib.mpb_sample_gp_kron_gen_mv(_lib.ptr(prior.scale_tril_kron_gen), _lib.ptr(means), C.byref(desc), _lib.ptr(x1),
                                                 P, S, 64, 7, _lib.ptr(Sinv), None, None, _lib.stream_ptr()))


@pytest.mark.parametrize('d,P,S', [(2, 9, 64), (3, 256, 128), (4, 5, 24), (5, 3, 7), (6, 40, 64)])
def test_gen_sampler_other_robots(d, P, S, dev):
    """The tcgen05 sampler for 2..6 dofs at H = 64 (point robots, 6-dof arms): samples against fp64 on its own dumped noise,
    the mat-vec warp's y and copy of the means bit-identical to mpb_prior_matvec_dof / the means, sharding-independent."""
    import ctypes as C
    prior, means = make_prior(P, dev, d=d)
    assert prior.scale_tril_kron_gen is not None, f'the tcgen05 sampler must be the default at (H, dof) = (64, {d})'
    lib = _lib.lib()
    M = 64 * 2 * d
    desc = nd(21, 3, p_off=1, P_glob=P + 2)
    x = prior.sample(S, noise_desc=desc).clone()
    eps = prior.replay_noise(desc, S)
    ref = means.view(P, 1, -1).double() + torch.einsum('ik,spk->psi', prior.scale_tril.double(), eps.double())
    amp = float((ref - means.view(P, 1, -1).double()).abs().max())
    assert float((x.view(P, S, -1).double() - ref).abs().max()) / amp < 5e-6
    Sinv = prior.Sigma_inv.contiguous()
    y_ref = torch.empty(P, M, **dev)
    _lib.check(lib.mpb_prior_matvec_dof(_lib.ptr(Sinv), _lib.ptr(means), _lib.ptr(y_ref), P, 64, d, _lib.stream_ptr()))
    x1, y, mu_c = torch.empty_like(x), torch.full((P, M), float('nan'), **dev), torch.full((P, M), float('nan'), **dev)
    _lib.check(lib.mpb_sample_gp_kron_gen_mv(_lib.ptr(prior.scale_tril_kron_gen), _lib.ptr(means.contiguous()), C.byref(desc),
                                             _lib.ptr(x1), P, S, 64, d, _lib.ptr(Sinv), _lib.ptr(y), _lib.ptr(mu_c), _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(x1, x) and torch.equal(y, y_ref) and torch.equal(mu_c, means.view(P, -1))
    # the second half of the particles as its own "rank": same bits
    h = P // 2
    if h:
        lo = make_prior(h, dev, means=means[:h].contiguous(), d=d)[0].sample(S, noise_desc=nd(21, 3, p_off=1, P_glob=P + 2))
        hi = make_prior(P - h, dev, means=means[h:].contiguous(), d=d)[0].sample(S, noise_desc=nd(21, 3, p_off=1 + h, P_glob=P + 2))
        assert torch.equal(torch.cat((lo, hi)), x)
