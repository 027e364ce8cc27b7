"""Parity checker shared by tests/, __graft_entry__.smoke() and bench.py's parity_check leg
(TEST INFRASTRUCTURE -- see oracle/__init__.py; never on the product path).

`stoch_gpmp_subset` replays one Stoch-GPMP iteration (mp_baselines/planners/stoch_gpmp.py:235-279) of a SUBSET
of particles on the CPU oracle, in fp32 (the reference's arithmetic) or fp64 (the yardstick for tolerances above
1e-5: "our error against fp64 must not exceed the fp32 reference's own error against fp64").
"""
import numpy as np
import torch

from . import planners as oplanners
from .build import oracle_field, oracle_robot
from .costs import CostSpec


def cost_spec(cfg, H, sig, dtype=torch.float32):
    ta = dict(device='cpu', dtype=dtype)
    start, goal = torch.as_tensor(np.asarray(cfg['start'])).to(**ta), torch.as_tensor(np.asarray(cfg['goal'])).to(**ta)
    return CostSpec(oracle_robot(cfg['robot'], cfg['dt'], ta), H, cfg['dt'], start, goal,
                    [oracle_field(cfg['obstacles'], cfg['robot'], ta)], sigma_start=sig['sigma_start'],
                    sigma_gp=sig['sigma_gp'], sigma_coll=sig['sigma_coll'], sigma_goal_prior=sig['sigma_goal_prior'],
                    tensor_args=ta)


def stoch_gpmp_subset(cfg, H, sig, means0, L, Sigma_inv, eps, particles, dtype=torch.float32):
    """means0 [P,H,D], L [M,M], Sigma_inv [M,M], eps [S,P,M] (any device), particles: index list.
    -> dict(samples [p,S,H,D], costs [p,S], weights [p,S], means [p,H,D], free [p,S] bool) on the CPU in `dtype`."""
    idx = torch.as_tensor(list(particles), dtype=torch.long)
    cpu = lambda t: t.detach().cpu()                    # noqa: E731
    m = cpu(means0)[idx].to(dtype)
    e = cpu(eps)[:, idx].to(dtype)
    spec = cost_spec(cfg, H, sig, dtype)
    with torch.no_grad():
        out = oplanners.stoch_gpmp_iteration(spec, m, cpu(L).to(dtype), cpu(Sigma_inv).to(dtype), e,
                                             sig['temperature'], sig['step_size'])
        p, S = out['costs'].shape
        out['free'] = spec.collision_free(out['samples'].reshape(p * S, H, -1)).reshape(p, S)
    return out


def rel_err(a, b, floor=0.0):
    """max |a-b| / max(|b|, floor) in fp64."""
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float(((a - b).abs() / b.abs().clamp_min(max(floor, 1e-300))).max())


def assert_not_worse_than_fp32(ours, ref32, ref64, what='', factor=2.0, rtol=1e-5):
    """The fp64 companion of a tolerance above 1e-5 ("hard parts"): in the max norm, our error
    against the fp64 oracle must be within `rtol` of the result's scale, or else no larger than `factor` x the error the
    reference's own fp32 arithmetic (`ref32`: a golden vector or the fp32 oracle) makes against fp64."""
    o, r32, r64 = (t.detach().cpu().double() for t in (ours, ref32, ref64))
    scale = float(r64.abs().max())
    e_ours, e_ref = float((o - r64).abs().max()), float((r32 - r64).abs().max())
    bound = max(rtol * scale, factor * e_ref)
    assert e_ours <= bound, (f'{what}: |ours - fp64| = {e_ours:.3e} exceeds max({rtol:g} x scale = {rtol * scale:.3e}, '
                             f'{factor:g} x |fp32 reference - fp64| = {factor * e_ref:.3e})')
    return e_ours, e_ref
