"""Oracle planner iterations (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Functional restatement of one iteration of each planner's hot loop:
  * Stoch-GPMP   mp_baselines/planners/stoch_gpmp.py:235-279
  * STOMP        mp_baselines/planners/stomp.py:68-108,162-220
  * MPPI         mp_baselines/planners/mppi.py:72-134,164-169,190-210
                 mp_baselines/planners/dynamics/point.py:102-140,154-226
                 mp_baselines/planners/priors/gaussian.py:166-198,271-298
  * CHOMP        mp_baselines/planners/chomp.py:81-101,127-169
  * GPMP2        mp_baselines/planners/gpmp2.py:308-368,432-452,493-495
Noise is always INJECTED (eps tensors shaped as torch.distributions would draw them).
Pinned by tests/golden/*.npz (generated from the unmodified reference planners).
"""
import torch

from .gp_prior import precision_to_scale_tril, sample_prior, sample_prior_faithful


# ------------------------------------------------------------------ Stoch-GPMP
def stoch_gpmp_costs(cost, samples, means, Sigma_inv, temperature):
    """cost.eval + importance-sampling ratio term (stoch_gpmp.py:235-242)."""
    P, S, H, D = samples.shape
    c = cost.eval(samples).reshape(P, S)
    V = samples.reshape(P, S, H * D)
    U = means.reshape(P, 1, H * D)
    return c + temperature * (V @ Sigma_inv @ U.transpose(1, 2)).squeeze(2)


def softmax_update(costs, samples, means, temperature, step_size):
    """w = softmax(-c/T) over samples; g = sum_s w (x_s - mu); mu += step*g
    (stoch_gpmp.py:267-279).  Returns (weights [P,S], grad [P,H,D], new_means)."""
    w = torch.softmax(-costs / temperature, dim=1)
    g = (w.reshape(*w.shape, 1, 1) * (samples - means.unsqueeze(1))).sum(1)
    return w, g, means + step_size * g


def stoch_gpmp_iteration(cost, means, L, Sigma_inv, eps, temperature, step_size,
                         faithful=False):
    """One optimize() iteration.  means [P,H,D]; L [M,M]; eps [S,P,M].
    faithful=True reproduces the reference's cost model for CPU-baseline timing: per-particle
    scale_tril [P,M,M] re-factorised from the (unchanged) precision and a broadcast batched
    mat-vec sampler (mp_priors_multi.py:100-110,120-123) -- identical maths."""
    P, H, D = means.shape
    if faithful:
        Lb = precision_to_scale_tril(Sigma_inv.repeat(P, 1, 1))
        samples = sample_prior_faithful(means.reshape(P, -1), Lb, eps)
    else:
        samples = sample_prior(means.reshape(P, -1), L, eps)
    samples = samples.reshape(P, -1, H, D)
    costs = stoch_gpmp_costs(cost, samples, means, Sigma_inv, temperature)
    w, g, new_means = softmax_update(costs, samples, means, temperature, step_size)
    return dict(samples=samples, costs=costs, weights=w, grad=g, means=new_means)


# ------------------------------------------------------------------ STOMP
def stomp_R(H, dt, sigma_spectral, tensor_args):
    """R = A^T A, A = second differences padded with a zero row on both sides and corner ones,
    scaled by sigma_spectral/dt^2 (stomp.py:68-86; built in torch's default fp32 then cast)."""
    A0 = (torch.diag(torch.ones(H - 1), 1) + torch.diag(torch.ones(H - 1), -1) - 2 * torch.eye(H))
    A = torch.cat((torch.zeros(1, H), A0, torch.zeros(1, H)), dim=0)
    A[0, 0] = 1.
    A[-1, -1] = 1.
    A = A * 1. / dt ** 2 * sigma_spectral
    return (A.t() @ A).to(**tensor_args)


def stomp_sample(means, L_R, eps):
    """eps [S,D,P,H] (sample_shape (S,D) + batch P + event H) -> samples [P,S,H,D];
    noise on the first/last waypoint is zeroed (stomp.py:97-108)."""
    noise = eps @ L_R.transpose(-1, -2)                         # [S,D,P,H]
    noise = noise.permute(2, 0, 3, 1).clone()                   # [P,S,H,D]
    noise[..., -1, :] = 0
    noise[..., 0, :] = 0
    return means.unsqueeze(1) + noise


def stomp_iteration(cost, means, L_R, Sigma_R, eps, temperature, lr):
    """One STOMP iteration: mu += lr * Sigma_R @ sum_s w (x_s - mu) (stomp.py:162-211)."""
    P, H, D = means.shape
    samples = stomp_sample(means, L_R, eps)
    S = samples.shape[1]
    costs = cost(samples.flatten(0, 1)).reshape(P, S)
    w = torch.softmax(-costs / temperature, dim=1)
    g = (w.reshape(P, S, 1, 1) * (samples - means.unsqueeze(1))).sum(1)
    return dict(samples=samples, costs=costs, weights=w, grad=g, means=means + lr * Sigma_R @ g)


# ------------------------------------------------------------------ MPPI
def mppi_cov(T, sigma, kind, tensor_args):
    """Per-control-dim covariance [T,T,C] (gaussian.py:143-198)."""
    sig = torch.as_tensor(sigma, **tensor_args).reshape(-1) if isinstance(sigma, (list, tuple)) else sigma
    if kind == 'const_ctrl':
        Lo = torch.tril(torch.ones(T, T - 1, **tensor_args), diagonal=-1)
        base = Lo @ Lo.t() + torch.ones(T, T, **tensor_args)
    else:
        base = torch.eye(T, **tensor_args)
    C = sig.numel() if torch.is_tensor(sig) else 1
    return base.unsqueeze(-1).repeat(1, 1, C) * sig ** 2


def mppi_rollout(U, state0, dt, ctrl_min, ctrl_max):
    """Velocity-control point dynamics: x_{t+1} = x_t + clamp(u_t)*dt, x_0 = state0
    (mppi.py:190-210, point.py:102-140 with control_type='velocity')."""
    N, T, C = U.shape
    X = torch.empty(N, T, state0.numel(), dtype=U.dtype)
    X[:, 0] = state0
    for i in range(T - 1):
        X[:, i + 1] = X[:, i] + U[:, i].clamp(min=ctrl_min, max=ctrl_max) * dt
    return X


def mppi_iteration(mean, L_ctrl, Cov_inv, eps, state0, goal, dt, ctrl_min, ctrl_max, c_weights,
                   temp, step_size, discount=1.0, ext_cost=None, mean_sample=None):
    """One MPPI iteration.  mean [T,C]; L_ctrl [C,T,T] (cholesky of Cov[:,:,i]); Cov_inv [C,T,T];
    eps [C,N,T].  Returns controls, states, costs [N,1], weights, new mean, argmin, min cost.
    mean_sample: the loc of ctrl_dist when it differs from `mean` -- the reference refreshes it only in
    update_ctrl_dist() (mppi.py:68-70,86), so after pop()/shift() (mppi.py:171-178) controls are still sampled around
    the UNSHIFTED mean while the IS term and the update use the shifted one."""
    T, C = mean.shape
    N = eps.shape[1]
    ms = mean if mean_sample is None else mean_sample
    U = torch.stack([ms[:, i] + eps[i] @ L_ctrl[i].t() for i in range(C)], dim=-1)    # [N,T,C]
    X = mppi_rollout(U, state0, dt, ctrl_min, ctrl_max)
    sd = X.shape[-1]
    disc = torch.cumprod(torch.ones(T, dtype=U.dtype) * discount, dim=0) / discount
    dX = X - goal[..., :sd]
    pos = (torch.square(dX[..., :sd]) * c_weights['pos']).sum(-1) * disc
    vel = (torch.square(dX[..., sd:C]) * c_weights['vel']).sum(-1) * disc
    ctl = (torch.square(U) * c_weights['ctrl']).sum(-1) * disc
    term = (torch.square(dX[:, -1, :]) * c_weights['pos_T']).sum(-1) * disc[-1]
    # quirk B2: the external cost is summed over the BATCH -> a scalar shift (point.py:194-196)
    energy = ext_cost.eval(torch.cat((X, U), dim=-1)).sum(-1) if ext_cost is not None else 0.
    costs = (pos.sum(1) + vel.sum(1) + ctl.sum(1) + term + energy).view(N, 1)
    for i in range(C):
        costs = costs + temp * (U[..., i] @ Cov_inv[i] @ mean[..., i]).reshape(-1, 1)
    best = torch.argmin(costs)
    w = torch.softmax(-costs / temp, dim=0)
    new_mean = mean + step_size * (w.reshape(-1, 1, 1) * (U - mean.unsqueeze(0))).sum(0)
    return dict(controls=U, states=X, costs=costs, weights=w, mean=new_mean,
                argmin=best, best_cost=costs.reshape(-1)[best])


# ------------------------------------------------------------------ CHOMP
def chomp_R(H, dt, tensor_args):
    """R = K^T K, K = backward differences with an extra last row, /dt^2 (chomp.py:81-101)."""
    K = torch.eye(H) - torch.diag(torch.ones(H - 1), -1)
    K = torch.cat((K, torch.zeros(1, H)), dim=0)
    K[-1, -1] = -1.
    K = K * 1. / dt ** 2
    return (K.t() @ K).to(**tensor_args)


def chomp_iteration(cost, x, R, weight_prior_cost, lr, grad_clip):
    """One CHOMP step.  The smoothness term is the GLOBAL sum over particles added to every
    particle's cost, so its gradient is scaled by P (quirk B1, chomp.py:139,165-167)."""
    x = x.detach().clone().requires_grad_(True)
    costs = cost(x)
    smooth = torch.einsum('phd,hk,pkd->pd', x, R, x).sum()
    costs = costs + weight_prior_cost * smooth
    g = torch.autograd.grad(costs.sum(), x)[0]
    g_raw = g.clone()
    g = g.clamp(-grad_clip, grad_clip)
    g[..., 0, :] = 0.
    g[..., -1, :] = 0.
    return dict(costs=costs.detach(), grad_raw=g_raw, grad=g, x=(x - lr * g).detach())


# ------------------------------------------------------------------ GPMP2
def gpmp2_step(cost, means, delta, trust_region, step_size):
    """One Gauss-Newton / LM step on the dense system (gpmp2.py:308-368,451-452)."""
    A, b, K = cost.linear_system(means)
    B, _, N = A.shape
    I = torch.eye(N, dtype=A.dtype)
    AtK = A.transpose(-2, -1) @ K
    AtA = AtK @ A
    if trust_region:
        JtJ = AtA + delta * (AtA.mean(0) * I)          # quirk B10: batch mean of the diagonal
    else:
        JtJ = AtA + delta * I
    g = AtK @ b
    Lc, _ = torch.linalg.cholesky_ex(JtJ)
    d_theta = torch.cholesky_solve(g, Lc).view(means.shape)
    costs = (b.transpose(1, 2) @ K @ b).reshape(B)
    return dict(A=A, b=b, K=K, JtJ=JtJ, g=g, d_theta=d_theta, costs=costs,
                means=means + step_size * d_theta)


# ------------------------------------------------------------------ sample-split update (multi-CTA / multi-GPU)
def partial_record(costs, x, mu, temperature, sample_offset=0):
    """Packed partial record of one block of samples:
    [m, Z, cmin, argmin, v...] with m = max_s(-c_s/T), Z = sum_s exp(-c_s/T - m), v = sum_s exp(-c_s/T - m)(x_s - mu).
    costs [P,S], x [P,S,M], mu [P,M] -> [P, 4+M] (float64; argmin stored as a number)."""
    P, S = costs.shape
    M = mu.shape[-1]
    rec = torch.zeros(P, 4 + M, dtype=torch.float64)
    if S == 0:
        rec[:, 0] = -float('inf')
        rec[:, 2] = float('inf')
        rec[:, 3] = 2 ** 31 - 1
        return rec
    a = -costs.double() / temperature
    m = a.max(dim=1).values
    e = torch.exp(a - m.unsqueeze(1))
    rec[:, 0] = m
    rec[:, 1] = e.sum(1)
    rec[:, 2] = costs.double().min(dim=1).values
    rec[:, 3] = costs.argmin(dim=1).double() + sample_offset
    rec[:, 4:] = (e.unsqueeze(-1) * (x.double() - mu.double().unsqueeze(1))).sum(1)
    return rec


def combine_records(recs, mu, step_size, Sigma_R=None, H=None):
    """Fixed-order log-sum-exp merge of records [R,P,4+M] -> dict(means, grad, lse [P,2], best_cost, best_idx).
    Ties of the minimum resolve to the lowest global sample index (torch.argmin semantics, mppi.py:166)."""
    R, P, W = recs.shape
    m = recs[:, :, 0].max(dim=0).values
    scale = torch.exp(recs[:, :, 0] - m.unsqueeze(0))
    scale[recs[:, :, 0] == -float('inf')] = 0
    Z = (scale * recs[:, :, 1]).sum(0)
    g = (scale.unsqueeze(-1) * recs[:, :, 4:]).sum(0) / Z.unsqueeze(-1)
    cmin = recs[:, :, 2].min(dim=0).values
    cand = torch.where(recs[:, :, 2] == cmin.unsqueeze(0), recs[:, :, 3], torch.full_like(recs[:, :, 3], float('inf')))
    best_idx = cand.min(dim=0).values.long()
    upd = g
    if Sigma_R is not None:
        upd = (Sigma_R.double() @ g.reshape(P, H, -1)).reshape(P, -1)
    return dict(means=mu.double() + step_size * upd, grad=g, lse=torch.stack((m, Z), dim=1), best_cost=cmin, best_idx=best_idx)
