"""Point-particle system with the reference's constructor (mp_baselines/planners/dynamics/point.py:5-77).
It is a parameter holder: MPPI's rollout and quadratic trajectory cost (point.py:102-140,154-226) run inside the
fused kernel csrc/mppi.cu, which reads the weights / limits / discount from this object."""
import numpy as np
import torch

from . import _lib


class PointParticleDynamics:
    def __init__(self, rollout_steps=None, control_dim=2, state_dim=2, dt=0.01, discount=1.0, deterministic=True,
                 start_state=None, goal_state=None, ctrl_min=None, ctrl_max=None, control_type='velocity',
                 dyn_std=np.zeros(4, ), c_weights=None, verbose=False, tensor_args=None):
        if tensor_args is None or torch.device(tensor_args['device']).type != 'cuda':
            raise _lib.MpbError("tensor_args['device'] must be a CUDA device: the hot path has no CPU implementation")
        self.tensor_args = dict(device=torch.device(tensor_args['device']), dtype=torch.float32)
        if control_type != 'velocity':
            raise NotImplementedError("only control_type='velocity' (what the reference examples use) is fused")
        if not deterministic:
            raise NotImplementedError('stochastic dynamics are not part of the fused rollout')
        self.control_dim = control_dim
        self.state_dim = state_dim
        self._c_weights = c_weights if c_weights is not None else {'pos': 10., 'vel': 10., 'ctrl': 0., 'pos_T': 10., 'vel_T': 0.}
        assert len(ctrl_min) == control_dim and len(ctrl_max) == control_dim
        self.ctrl_min = torch.tensor(ctrl_min).to(**self.tensor_args).contiguous()
        self.ctrl_max = torch.tensor(ctrl_max).to(**self.tensor_args).contiguous()
        self.discount = discount
        self.discount_seq = torch.cumprod(torch.ones(rollout_steps) * discount, dim=0).div_(discount).to(**self.tensor_args)
        if start_state is not None:
            self.start_state = torch.as_tensor(start_state).to(**self.tensor_args)
        else:
            self.start_state = torch.zeros(state_dim, **self.tensor_args)
        self.state = self.start_state.clone()
        if goal_state is not None:
            self.goal_state = torch.as_tensor(goal_state).to(**self.tensor_args)
        else:
            self.goal_state = torch.zeros(state_dim, **self.tensor_args)
        self.rollout_steps = rollout_steps
        self.dt = dt
        self.control_type = control_type
        self.verbose = verbose
        self.deterministic = deterministic

    def dynamics(self, x, u_s, use_crn=False):
        """One Euler step x + clamp(u) dt (point.py:102-140) -- executing an action between planner calls, not the
        planner's rollout."""
        return x + u_s.clamp(min=self.ctrl_min, max=self.ctrl_max) * self.dt

    def step(self, action):
        self.state = self.dynamics(self.state.reshape(1, 1, -1), action.reshape(1, 1, -1)).reshape(-1)
        return self.state
