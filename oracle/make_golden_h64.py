"""Golden fixtures at the benchmarked horizon (H = 64, Panda) from the UNMODIFIED reference Stoch-GPMP, so that the
default structured K1 sampler (which exists for H in {32, 64, 128} only) meets reference data, not only the oracle.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the build container only (needs /root/reference):

    python oracle/make_golden_h64.py

P = 2 particles, S = 4 samples keep the files small; the [M,M] factor is stored as its diagonal + the dof-0 block.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import make_golden as mg  # noqa: E402

if __name__ == '__main__':
    torch.set_num_threads(4)
    mg.gen_stoch_gpmp('stochgpmp_panda_h64_moderate', 'C4', P=2, S=4, H=64, iters=2, sig=mg.MODERATE, seed=60, store_factor=False)
    mg.gen_stoch_gpmp('stochgpmp_panda_h64_frozen', 'C4', P=2, S=4, H=64, iters=1, sig=mg.FROZEN, seed=61, store_factor=False)
