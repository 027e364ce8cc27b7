"""Host side of the record-based importance-weight update (csrc/softmax_split.cu) and of the sample-split
multi-GPU mode (one problem's samples sharded over the GPUs of a box).

Single GPU: partial records per chunk of samples -> fixed-order combine.
Several GPUs: every rank produces the records of ITS samples, ONE all-gather concatenates them along the chunk
axis (rank-major, i.e. global sample order), and every rank runs the same fixed-order combine, so all ranks end up
with bit-identical means without a second collective.  The payload is R x P x (4 + H*Dw) floats -- a few kB:
latency-bound on NVLink, which is why it is a single all-gather rather than max-allreduce + sum-allreduce.
"""
import torch
import torch.distributed as dist

from . import _lib


class SampleSplit:
    """Contiguous partition of S samples over the ranks of a process group (rank r owns a contiguous block, blocks in
    rank order, sizes differing by at most one)."""

    def __init__(self, group=None, rank=None, world=None):
        self.group = group
        if world is None:
            world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
        assert 0 <= rank < world
        self.rank, self.world = rank, world

    def counts(self, S):
        q, r = divmod(S, self.world)
        return [q + (1 if i < r else 0) for i in range(self.world)]

    def local_slice(self, S):
        c = self.counts(S)
        return sum(c[:self.rank]), c[self.rank]

    def all_gather_cat(self, t):
        """[n, ...] per rank (same n everywhere) -> [world*n, ...] in rank order."""
        if self.world == 1:
            return t
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out

    def owner_of(self, idx, S):
        """Rank owning global sample index ``idx`` (device tensor; no host synchronisation)."""
        c = self.counts(S)
        ends = torch.tensor([sum(c[:i + 1]) for i in range(self.world)], device=idx.device, dtype=idx.dtype)
        return torch.bucketize(idx, ends, right=True).clamp_(max=self.world - 1)


def default_chunks(P, S_local, S_max=None, sms=148):
    """Chunks per particle: enough CTAs to fill the GPU twice, at least 64 samples per chunk, and few enough samples
    per chunk for the kernel's shared-memory staging.  With several ranks every rank must use the same value, so it
    is derived from the largest local sample count."""
    S_ref = S_max if S_max is not None else S_local
    n = max(1, min(-(-2 * sms // max(P, 1)), -(-S_ref // 64)))
    return max(n, -(-S_ref // 16384))


def split_softmax_update(cost, x, mu, temp, step, H, Dfull, c0=0, Dw=None, SigmaR=None, n_chunks=None, split=None,
                         S_global=None, weights_out=None, want_grad=False):
    """cost [P,S_local], x [P,S_local,H,Dfull], mu [P,H,Dw] (updated in place).
    -> dict(weights [P,S_local], lse [P,2], best_cost [P], best_idx [P] int32 (global sample index), grad|None)."""
    lib, st = _lib.lib(), _lib.stream_ptr()
    _lib.require_f32(cost, x, mu)
    P, S_local = cost.shape
    Dw = Dfull if Dw is None else Dw
    split = split or SampleSplit(world=1, rank=0)
    S_global = S_local if S_global is None else S_global
    offset, cnt = split.local_slice(S_global) if split.world > 1 else (0, S_local)
    assert cnt == S_local, f'rank {split.rank} should hold {cnt} samples, got {S_local}'
    S_max = max(split.counts(S_global)) if split.world > 1 else S_local
    if n_chunks is None:
        n_chunks = default_chunks(P, S_local, S_max)
    REC = lib.mpb_softmax_record_len(H, Dw)
    dev = cost.device
    rec = torch.empty(n_chunks, P, REC, device=dev, dtype=torch.float32)
    cost = cost.contiguous()        # bound to a name: a temporary would be freed (and reusable) before the launch
    _lib.check(lib.mpb_softmax_partial(_lib.ptr(cost), _lib.ptr(x), _lib.ptr(mu), _lib.ptr(rec), float(temp), P, S_local,
                                       H, Dfull, c0, Dw, n_chunks, offset, st))
    rec_all = split.all_gather_cat(rec)
    lse = torch.empty(P, 2, device=dev, dtype=torch.float32)
    best_cost = torch.empty(P, device=dev, dtype=torch.float32)
    best_idx = torch.empty(P, device=dev, dtype=torch.int32)
    grad = torch.empty(P, H, Dw, device=dev, dtype=torch.float32) if (want_grad or SigmaR is not None) else None    # Sigma_R is applied to the merged mean
    _lib.check(lib.mpb_softmax_combine(_lib.ptr(rec_all), rec_all.shape[0], _lib.ptr(mu), _lib.ptr(grad), _lib.ptr(lse),
                                       _lib.ptr(best_cost), _lib.ptr(best_idx), float(step), _lib.ptr(SigmaR), P, H, Dw, st))
    weights = weights_out if weights_out is not None else torch.empty(P, S_local, device=dev, dtype=torch.float32)
    _lib.check(lib.mpb_softmax_weights(_lib.ptr(cost), _lib.ptr(lse), _lib.ptr(weights), float(temp), P, S_local, st))
    return dict(weights=weights, lse=lse, best_cost=best_cost, best_idx=best_idx, grad=grad if want_grad else None, records=rec_all)
