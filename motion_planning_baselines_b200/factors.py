"""GP / unary factors and the GP trajectory prior with the reference's names and signatures
(mp_baselines/planners/costs/factors/{gp_factor,unary_factor,mp_priors_multi}.py).

Construction is one-off host-side setup: the precision is assembled in fp64
on the CPU exactly like the reference does, and the sampling factor ``scale_tril`` comes from the
very routine the reference uses (torch.distributions.MultivariateNormal(precision_matrix=...) on the
CPU in the working dtype), so that the fused sampler multiplies by bit-identical numbers.  The hot
part -- drawing samples every iteration -- runs in the mpb_sample_gp kernel, and ``set_mean`` only
swaps the mean: the factor of an unchanged precision is never recomputed (reference quirk B4).
"""
import ctypes as C
import os

import torch
import torch.distributions as dist

from . import _lib

_CPU32 = dict(device='cpu', dtype=torch.float32)


class UnaryFactor:
    """K = I / sigma^2, error = mean - x (unary_factor.py:6-32)."""

    def __init__(self, dim, sigma, mean=None, tensor_args=None):
        self.sigma, self.dim, self.tensor_args = sigma, dim, tensor_args
        self.mean = torch.zeros(dim, **tensor_args) if mean is None else mean
        self.K = (torch.eye(dim, **_CPU32) / sigma ** 2).to(**tensor_args)

    def set_mean(self, x):
        self.mean = x.clone().detach()


class GPFactor:
    """Constant-velocity GP factor: Phi and Q^-1 (gp_factor.py:4-50)."""

    def __init__(self, dim, sigma, d_t, num_factors, tensor_args=None, Q_c_inv=None):
        self.dim, self.d_t, self.tensor_args = dim, d_t, tensor_args
        self.state_dim = 2 * dim
        self.num_factors = num_factors
        eye = torch.eye(dim, **_CPU32)
        Qc = eye / sigma ** 2 if Q_c_inv is None else Q_c_inv.to(**_CPU32)
        a, b, c = 12. * (d_t ** -3.) * Qc, -6. * (d_t ** -2.) * Qc, 4. * (d_t ** -1.) * Qc
        Q = torch.cat((torch.cat((a, b), dim=-1), torch.cat((b, c), dim=-1)), dim=-2)
        self.Q_c_inv = Qc.to(**tensor_args)
        self.Q_inv = Q.unsqueeze(0).to(**tensor_args)           # [1,D,D]; the reference repeats it num_factors times
        Phi = torch.eye(2 * dim, **_CPU32)
        Phi[:dim, dim:] = eye * d_t
        self.phi = Phi.to(**tensor_args)


def _precision(num_steps, dt, state_dim, dof, K_s_inv, K_gp_inv, K_g_inv):
    """Sigma^-1 = A^T Q^-1 A in fp64 on the CPU (mp_priors_multi.py:213-251)."""
    H, D, M = num_steps + 1, state_dim, state_dim * (num_steps + 1)
    f64 = dict(device='cpu', dtype=torch.float64)
    Phi = torch.eye(D, **f64)
    Phi[:dof, dof:] = torch.eye(dof, **f64) * dt
    A = torch.eye(M, **f64)
    A[D:, :-D] -= torch.kron(torch.eye(H - 1, **f64), Phi)
    blocks = [K_s_inv.to(**f64)] + [K_gp_inv.to(**f64)] * (H - 1)
    if K_g_inv is not None:
        tail = torch.zeros(D, M, **f64)
        tail[:, -D:] = torch.eye(D, **f64)
        A = torch.cat((A, tail))
        blocks.append(K_g_inv.to(**f64))
    return A.t() @ torch.block_diag(*blocks) @ A


class MultiMPPrior:
    """Gaussian trajectory prior N(means, Sigma) with Sigma^-1 block-tridiagonal
    (mp_priors_multi.py:15-259).  ``sample`` runs on the GPU (csrc/sample_gp.cu)."""

    def __init__(self, num_steps, dt, state_dim, dof, K_s_inv, K_gp_inv, start_state, means=None, K_g_inv=None,
                 goal_states=None, use_numpy=False, tensor_args=None, factor_dtype=torch.float32, noise=None):
        self.state_dim, self.dof, self.num_steps = state_dim, dof, num_steps
        self.M = state_dim * (num_steps + 1)
        self.tensor_args = tensor_args
        self.goal_directed = goal_states is not None
        if means is None:
            self.num_modes = goal_states.shape[0] if self.goal_directed else 1
            means = self.get_const_vel_mean(start_state, goal_states, dt, num_steps, dof)
        else:
            self.num_modes = means.shape[0]
        self.means = means.reshape(self.num_modes, -1).to(**tensor_args).contiguous()
        # in-kernel noise (csrc/philox.cuh): seed from torch's global seed unless the owner hands in its own stream
        self.noise = noise if noise is not None else _lib.NoiseStream(P_global=self.num_modes)

        Sinv64 = _precision(num_steps, dt, state_dim, dof, K_s_inv, K_gp_inv, K_g_inv if self.goal_directed else None)
        Sinv_cpu = Sinv64.to(factor_dtype)      # float64: the reference's init path (base.py:155-158, quirk B8)
        self.Sigma_inv = Sinv_cpu.to(**tensor_args).contiguous()
        # the reference's own factorisation routine, once, on the CPU, in the working dtype.  NOTE (measured):
        # at H = 64 this fp32 factor is accurate to ~5e-3 only (cond(Sigma^-1) ~ 2e6) and differs by ~1e-2 between LAPACK
        # thread counts -- in the reference too -- so a bit-level replay of reference samples needs the reference's
        # factor (set_scale_tril), not just its noise.
        self.set_scale_tril(dist.MultivariateNormal(torch.zeros(self.M), precision_matrix=Sinv_cpu).scale_tril)

    def set_scale_tril(self, scale_tril):
        """Install a lower-triangular factor [M,M] and prepare the sampler operands for it (one-off, setup time)."""
        tensor_args, state_dim, dof, num_steps = self.tensor_args, self.state_dim, self.dof, self.num_steps
        assert scale_tril.shape == (self.M, self.M)
        self.scale_tril = scale_tril.to(**tensor_args).contiguous()
        # Sampler selection (MPB_SAMPLE_GP = kron | kron_umma | kron_fp32 | tc | simt forces one; default: the first that applies).
        #  kron: the factor decouples over the dofs (verified bit-exactly on the device) -> per-dof [2H,2H] blocks
        #  tc  : dense tcgen05 3xTF32 sampler; L pre-split into two TF32-representable parts
        #  simt: dense FP32 sampler
        mode = os.environ.get('MPB_SAMPLE_GP', 'auto')
        H = num_steps + 1
        self.scale_tril_kron = None
        self.scale_tril_kron_tc, self.kron_tc_kind = None, 0
        self.scale_tril_kron_gen = None
        if mode in ('auto', 'kron', 'kron_umma', 'kron_fp32') and state_dim == 2 * dof and _lib.lib().mpb_sample_gp_kron_supported(H, dof):
            packed = torch.empty(dof, 2 * H, 2 * H, **tensor_args)
            ok = C.c_int(0)
            _lib.check(_lib.lib().mpb_sample_gp_kron_pack(_lib.ptr(self.scale_tril), _lib.ptr(packed), H, dof,
                                                          C.byref(ok), _lib.stream_ptr()))
            if ok.value:
                self.scale_tril_kron = packed
                # tensor-core operand: warp-MMA fp16 fragments ('kron', default: fastest measured, 0.107 ms at C4) or
                # tcgen05 tiles ('kron_umma': 0.173 ms, bound by the MMA-issue / TMEM-store handshakes of its many
                # small N=32 MMAs); 'kron_fp32' keeps the exact FP32 kernel
                lib = _lib.lib()
                if mode == 'kron_umma' and lib.mpb_sample_gp_kron_umma_supported(H, dof):
                    self.scale_tril_kron_tc = torch.empty(lib.mpb_sample_gp_kron_umma_floats(H, dof), **tensor_args)
                    _lib.check(lib.mpb_sample_gp_kron_umma_prepare(_lib.ptr(packed), _lib.ptr(self.scale_tril_kron_tc),
                                                                   H, dof, _lib.stream_ptr()))
                    self.kron_tc_kind = 2
                elif mode != 'kron_fp32':
                    self.scale_tril_kron_tc = torch.empty(lib.mpb_sample_gp_kron_tc_bytes(H, dof), device=tensor_args['device'],
                                                          dtype=torch.uint8)
                    _lib.check(lib.mpb_sample_gp_kron_tc_prepare(_lib.ptr(packed), _lib.ptr(self.scale_tril_kron_tc),
                                                                 H, dof, _lib.stream_ptr()))
                    self.kron_tc_kind = 1
                    # Blackwell path for the in-kernel noise draw (tcgen05 + bulk-async copies; noise layout NOISE_SPMD):
                    # the default wherever it is built; MPB_SAMPLE_GP=kron keeps the warp-MMA kernel for both draws
                    if mode == 'auto' and lib.mpb_sample_gp_kron_gen_supported(H, dof):
                        self.scale_tril_kron_gen = torch.empty(lib.mpb_sample_gp_kron_gen_bytes(H, dof),
                                                               device=tensor_args['device'], dtype=torch.uint8)
                        _lib.check(lib.mpb_sample_gp_kron_gen_prepare(_lib.ptr(packed), _lib.ptr(self.scale_tril_kron_gen),
                                                                      H, dof, _lib.stream_ptr()))
        self.scale_tril_split = None
        if self.scale_tril_kron is None and mode != 'simt' and _lib.lib().mpb_sample_gp_tc_supported(1, 1, self.M):
            self.scale_tril_split = torch.empty(2, self.M, self.M, **tensor_args)
            _lib.check(_lib.lib().mpb_split_tf32(_lib.ptr(self.scale_tril), _lib.ptr(self.scale_tril_split[0]),
                                                 _lib.ptr(self.scale_tril_split[1]), self.M * self.M, _lib.stream_ptr()))

    @classmethod
    def const_vel_trajectory(cls, start_state, goal_state, dt, num_steps, dof, set_initial_final_vel_to_zero=True,
                             tensor_args=None):
        """Straight line, constant velocity, zero velocity at both ends (mp_priors_multi.py:130-151)."""
        w = torch.arange(num_steps + 1, **tensor_args).unsqueeze(-1)
        traj = torch.zeros(num_steps + 1, 2 * dof, **tensor_args)
        traj[:, :dof] = start_state[:dof] * (num_steps - w) * 1. / num_steps + goal_state[:dof] * w * 1. / num_steps
        vel = (goal_state[:dof] - start_state[:dof]) / (num_steps * dt)
        if set_initial_final_vel_to_zero:
            traj[1:-1, dof:] = vel
        else:
            traj[:, dof:] = vel
        return traj

    def get_const_vel_mean(self, start_state, goal_states, dt, num_steps, dof):
        if self.goal_directed:
            return torch.stack([self.const_vel_trajectory(start_state, g, dt, num_steps, dof, tensor_args=self.tensor_args)
                                for g in goal_states], dim=0)
        return start_state.repeat(num_steps + 1, 1)

    def get_mean(self, reshape=True):
        m = self.means.clone().detach()
        return m.reshape(self.num_modes, self.num_steps + 1, self.state_dim) if reshape else m

    def set_mean(self, means_new):
        assert means_new.shape == self.means.shape
        self.means = means_new.clone().detach().contiguous()

    @property
    def noise_layout(self):
        """Layout of the virtual global noise tensor the in-kernel draw of sample() uses (for _lib.philox_normal replays)."""
        return _lib.NOISE_SPMD if self.scale_tril_kron_gen is not None else _lib.NOISE_SPM

    def replay_noise(self, noise_desc, num_samples):
        """The [S,P,M] noise sample(num_samples, noise_desc=noise_desc) draws -- feed it back as ``eps`` to replay a run."""
        return _lib.philox_normal(noise_desc, self.noise_layout, (num_samples, self.num_modes, self.M),
                                  self.tensor_args['device'], dof=self.dof)

    def sample(self, num_samples, eps=None, out=None, noise_desc=None):
        """-> [num_modes, num_samples, H, state_dim].  ``eps`` ([S,P,M], the layout torch draws) can be injected for
        parity runs; otherwise the noise is drawn on the device by Philox keyed on the global element index
        (``noise_desc``, default: the next draw of ``self.noise``) -- inside the sampler itself for the default
        structured kernel, by mpb_philox_normal in front of the other variants (same numbers either way)."""
        P, S, M = self.num_modes, num_samples, self.M
        x = out if out is not None else torch.empty(P, S, M, **self.tensor_args)
        lib = _lib.lib()
        if eps is None:
            nd = noise_desc if noise_desc is not None else self.noise.next()
            if self.scale_tril_kron_gen is not None:
                _lib.check(lib.mpb_sample_gp_kron_gen(_lib.ptr(self.scale_tril_kron_gen), _lib.ptr(self.means), C.byref(nd),
                                                      _lib.ptr(x), P, S, self.num_steps + 1, self.dof, _lib.stream_ptr()))
                return x.view(P, S, self.num_steps + 1, self.state_dim)
            if self.kron_tc_kind == 1:
                _lib.check(lib.mpb_sample_gp_kron_tc_rng(_lib.ptr(self.scale_tril_kron_tc), _lib.ptr(self.means), C.byref(nd),
                                                         _lib.ptr(x), P, S, self.num_steps + 1, self.dof, _lib.stream_ptr()))
                return x.view(P, S, self.num_steps + 1, self.state_dim)
            eps = _lib.philox_normal(nd, _lib.NOISE_SPM, (S, P, M), self.tensor_args['device'])
        _lib.require_f32(eps)
        assert eps.shape == (S, P, M)
        eps = eps.contiguous()
        if self.kron_tc_kind == 2:
            _lib.check(lib.mpb_sample_gp_kron_umma(_lib.ptr(self.scale_tril_kron_tc), _lib.ptr(self.means), _lib.ptr(eps),
                                                   _lib.ptr(x), P, S, self.num_steps + 1, self.dof, _lib.stream_ptr()))
        elif self.kron_tc_kind == 1:
            _lib.check(lib.mpb_sample_gp_kron_tc(_lib.ptr(self.scale_tril_kron_tc), _lib.ptr(self.means), _lib.ptr(eps),
                                                 _lib.ptr(x), P, S, self.num_steps + 1, self.dof, _lib.stream_ptr()))
        elif self.scale_tril_kron is not None:
            _lib.check(lib.mpb_sample_gp_kron(_lib.ptr(self.scale_tril_kron), _lib.ptr(self.means), _lib.ptr(eps),
                                              _lib.ptr(x), P, S, self.num_steps + 1, self.dof, _lib.stream_ptr()))
        elif self.scale_tril_split is not None:
            _lib.check(lib.mpb_sample_gp_tc(_lib.ptr(self.scale_tril_split[0]), _lib.ptr(self.scale_tril_split[1]),
                                            _lib.ptr(self.means), _lib.ptr(eps), _lib.ptr(x), P, S, M, _lib.stream_ptr()))
        else:
            _lib.check(lib.mpb_sample_gp(_lib.ptr(self.scale_tril), _lib.ptr(self.means), _lib.ptr(eps),
                                         _lib.ptr(x), P, S, M, _lib.stream_ptr()))
        return x.view(P, S, self.num_steps + 1, self.state_dim)
