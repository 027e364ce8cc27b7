"""Stand-in (imported but unused at mp_baselines/planners/costs/cost_functions.py:12)."""


def link_pos_from_link_tensor(H):
    return H[..., :-1, -1]
