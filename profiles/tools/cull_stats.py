import sys, numpy as np, torch
sys.path.insert(0, '.')
from motion_planning_baselines_b200 import configs
from oracle.build import TA, oracle_robot
from oracle import gp_prior
torch.manual_seed(0)
cfg = configs.config(sys.argv[1] if len(sys.argv) > 1 else 'C4')
model, obst = cfg['robot'], cfg['obstacles']
H, d = 64, 7
sig = cfg['params'] if 'sigma_gp_sample' in cfg['params'] else configs.STOCH_GPMP_SIGMAS
robot = oracle_robot(model, cfg['dt'])
start, goal = torch.tensor(cfg['start']), torch.tensor(cfg['goal'])
K_s = gp_prior.unary_K(14, sig['sigma_start_sample'], TA); K_g = gp_prior.unary_K(14, sig['sigma_goal_sample'], TA)
Q = gp_prior.gp_Q_inv(7, cfg['dt'], sig['sigma_gp_sample'], TA)
Sinv = gp_prior.prior_precision(H, 7, cfg['dt'], K_s, Q, K_g, TA)
L = gp_prior.precision_to_scale_tril(Sinv)
mean = gp_prior.const_vel_mean(torch.cat((start, torch.zeros(7))), torch.cat((goal, torch.zeros(7))), cfg['dt'], H, 7, TA)
# particles: init dist samples; samples around them
P, S = 8, 16
Qi = gp_prior.gp_Q_inv(7, cfg['dt'], sig['sigma_gp_init'], TA)
Li = gp_prior.precision_to_scale_tril(gp_prior.prior_precision(H, 7, cfg['dt'], K_s, Qi, K_g, TA))
means = mean.reshape(1, -1) + torch.randn(P, H * 14) @ Li.T
x = (means.unsqueeze(1) + torch.randn(P, S, H * 14) @ L.T).reshape(P * S, H, 14)
q = x[..., :7]
R, t = robot.link_frames(q)                       # [B,H,7,3,3], [B,H,7,3]
centers = robot.fk_map_collision(q)               # [B,H,50,3]
link = robot.sphere_link
rad = robot.link_radii
margin = obst.cutoff_margin
if obst.n_spheres:
    oc = torch.tensor(obst.sphere_centers); orad = torch.tensor(obst.sphere_radii)
else:   # boxes: bounding spheres
    oc = torch.tensor(obst.box_centers); orad = torch.tensor(obst.box_half).norm(dim=-1)
B = x.shape[0]
dist = (centers.unsqueeze(-2) - oc).norm(dim=-1)          # [B,H,50,No]
pair_cand = dist < (rad.view(1, 1, -1, 1) + margin + orad)
print('pair candidate fraction', pair_cand.float().mean().item(), ' sphere candidate fraction', pair_cand.any(-1).float().mean().item())
# link bounding spheres (in world): centre = mean of sphere centres of link, radius = max dist + r
tot_pairs = 0; active_pairs_warp = 0; active_pairs_lane = 0
for j in range(7):
    idx = (link == j).nonzero().flatten()
    c = centers[:, :, idx]                       # [B,H,n,3]
    bc = c.mean(2)                               # [B,H,3]
    br = ((c - bc.unsqueeze(2)).norm(dim=-1) + rad[idx]).max(-1).values  # [B,H]
    dd = (bc.unsqueeze(-2) - oc).norm(dim=-1)                      # [B,H,No]
    act = dd < (br.unsqueeze(-1) + margin + orad)                  # [B,H,No] per lane
    act_w = act.reshape(B, H // 32, 32, -1).any(2)                 # warp-level
    n = len(idx)
    tot_pairs += n * oc.shape[0]
    active_pairs_warp += n * act_w.float().mean().item() * oc.shape[0]
    active_pairs_lane += n * act.float().mean().item() * oc.shape[0]
    print(f'link {j}: n={n} R~{br.mean():.3f} active obstacles/warp {act_w.float().sum(-1).mean():.2f} of {oc.shape[0]} (per lane {act.float().sum(-1).mean():.2f})')
print('total pairs', tot_pairs, 'after warp-level link culling', active_pairs_warp, 'per-lane', active_pairs_lane)

# --- emulate the kernel's warp AABB broad phase (collision.cuh broad_phase) -----------------------------------------
print('\nAABB broad phase as in the kernel (per warp = 32 consecutive waypoints):')
tot_aabb = 0
for j in range(7):
    idx = (link == j).nonzero().flatten()
    c = centers[:, :, idx]
    bc = c.mean(2)                                           # stand-in for the link bounding-sphere centre
    br = ((c - bc.unsqueeze(2)).norm(dim=-1) + rad[idx]).max(-1).values
    Rm = br.max() + margin
    bcw = bc.reshape(B, H // 32, 32, 3)
    lo, hi = bcw.min(2).values, bcw.max(2).values            # [B,2,3]
    dlo = (lo.unsqueeze(-2) - oc).clamp_min(0)               # [B,2,No,3]
    dhi = (oc - hi.unsqueeze(-2)).clamp_min(0)
    dd = torch.maximum(dlo, dhi).norm(dim=-1)
    near = dd < (Rm + orad)
    n_ls = near.float().sum(-1)
    tot_aabb += len(idx) * n_ls.mean().item()
    print(f'link {j}: n={len(idx)} AABB extent {float((hi - lo).mean()):.3f}  n_ls mean {n_ls.mean():.2f} max {int(n_ls.max())}  (links with empty list: {float((n_ls == 0).float().mean()):.2f})')
print('sphere x obstacle pairs per waypoint after the AABB broad phase:', tot_aabb, ' vs exact per-warp bounding-sphere test', active_pairs_warp)
