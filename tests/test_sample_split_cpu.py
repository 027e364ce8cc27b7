"""CPU (gloo, world_size 2): host logic of the sample-split multi-GPU mode -- the SampleSplit partition, the
rank-major all-gather of packed records, owner lookup -- checked against the unsplit softmax update.  The records
themselves are produced here by the oracle (on the GPU box they come from mpb_softmax_partial; tests/test_gpu_mppi.py
compares the two)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from motion_planning_baselines_b200.update import SampleSplit, default_chunks
from oracle import planners as op


def test_partition_is_contiguous_and_complete():
    for S in (0, 1, 7, 64, 1000, 10 ** 6 + 3):
        for world in (1, 2, 3, 8):
            parts = [SampleSplit(rank=r, world=world).local_slice(S) for r in range(world)]
            assert sum(c for _, c in parts) == S
            pos = 0
            for off, c in parts:
                assert off == pos
                pos += c
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    sp = SampleSplit(rank=0, world=3)
    idx = torch.tensor([0, 3, 4, 6, 7, 9])
    assert sp.counts(10) == [4, 3, 3]
    assert sp.owner_of(idx, 10).tolist() == [0, 0, 1, 1, 2, 2]
    assert default_chunks(1, 10 ** 6) >= 10 ** 6 // 16384 and default_chunks(512, 64) == 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, S, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        gen = torch.Generator().manual_seed(0)          # every rank builds the same global problem ...
        P, H, D, temp, step = 3, 5, 4, 0.7, 0.3
        M = H * D
        mu = torch.randn(P, M, generator=gen)
        x = mu.unsqueeze(1) + 0.2 * torch.randn(P, S, M, generator=gen)
        costs = 5 * torch.rand(P, S, generator=gen)
        costs[0, S // 3] = costs[0, S - 1] = -1.0        # tie of the minimum across ranks: lowest index must win
        split = SampleSplit()
        off, cnt = split.local_slice(S)                   # ... and works on its own block of samples
        n_chunks = 2
        sub = [(off + (cnt * c) // n_chunks, off + (cnt * (c + 1)) // n_chunks) for c in range(n_chunks)]
        rec = torch.stack([op.partial_record(costs[:, a:b], x[:, a:b], mu, temp, sample_offset=a) for a, b in sub])
        rec_all = split.all_gather_cat(rec)               # [world*n_chunks, P, 4+M] in global sample order
        assert rec_all.shape == (world * n_chunks, P, 4 + M)
        got = op.combine_records(rec_all, mu, step)
        w_ref, g_ref, mu_ref = op.softmax_update(costs.double(), x.double().reshape(P, S, H, D), mu.double().reshape(P, H, D), temp, step)
        assert torch.allclose(got['means'], mu_ref.reshape(P, M), rtol=1e-12, atol=1e-12)
        assert torch.allclose(got['grad'], g_ref.reshape(P, M), rtol=1e-11, atol=1e-12)
        assert got['best_idx'].tolist() == costs.argmin(dim=1).tolist(), 'first-occurrence argmin across ranks'
        w_local = torch.exp(-costs[:, off:off + cnt].double() / temp - got['lse'][:, :1]) / got['lse'][:, 1:]
        assert torch.allclose(w_local, w_ref[:, off:off + cnt], rtol=1e-12, atol=1e-300)
        owner = split.owner_of(got['best_idx'], S)
        assert all(sum(split.counts(S)[:o]) <= i < sum(split.counts(S)[:o + 1]) for o, i in zip(owner.tolist(), got['best_idx'].tolist()))
        # every rank must end with bit-identical means
        gathered = split.all_gather_cat(got['means'].unsqueeze(0))
        assert torch.equal(gathered[0], gathered[-1])
        if rank == 0:
            out.put('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('S', [64, 37])
def test_two_rank_gather_and_combine_matches_unsplit_update(S):
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, S, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert out.get(timeout=5) == 'ok'
