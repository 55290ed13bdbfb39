"""Frame-pair sharding for multi-GPU runs.

The hot path has no cross-pair state (main.cc:302-559 reads frames 1,2 of one Camera only), so a
sequence of pairs is partitioned into contiguous blocks, one block per rank (one process per GPU),
with NO data-path collective; the only communication is the final gather of the per-pair result
records (v, w, k, LM iterations: 8 doubles per pair).  Works with any torch.distributed backend
(nccl on the GPU box, gloo in the CPU tests).
"""
import numpy as np

RECORD = 8   # v[3], w[3], k, lm_iterations


def shard_range(num_pairs, rank, world_size):
    """Contiguous block [lo, hi) of pair indices owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(num_pairs, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def run_shard(process_pair, num_pairs, rank, world_size):
    """Calls process_pair(pair_index) -> (v, w, k, iterations) for every pair of this rank's block;
    returns (lo, records[n_local, RECORD])."""
    lo, hi = shard_range(num_pairs, rank, world_size)
    rec = np.zeros((hi - lo, RECORD))
    for j, p in enumerate(range(lo, hi)):
        v, w, k, it = process_pair(p)
        rec[j, 0:3] = v; rec[j, 3:6] = w; rec[j, 6] = k; rec[j, 7] = it
    return lo, rec


def run_shard_sequence(ctx, make_pair, num_pairs, rank, world_size, const_acc, gs_mode, K4, gamma, batch=8, layout=0):
    """This rank's block through the pipelined sequence entry point (rsdsfm_refine_rectify_sequence), `batch`
    pairs per call so that only `batch` pairs have to be resident at a time.  make_pair(pair_index) -> dict
    with flow, inliers3, alpha, alpha_k, image, m, v, w, k (numpy = host buffers, torch CUDA = device
    buffers; optionally `out`).  Returns (lo, records[n_local, RECORD], results) with the per-pair result
    dicts of capi.Context.refine_rectify_sequence."""
    lo, hi = shard_range(num_pairs, rank, world_size)
    rec = np.zeros((hi - lo, RECORD))
    results = []
    for b0 in range(lo, hi, batch):
        block = [make_pair(p) for p in range(b0, min(b0 + batch, hi))]
        out = ctx.refine_rectify_sequence(block, const_acc, gs_mode, K4, gamma, layout=layout)
        for j, r in enumerate(out):
            i = b0 - lo + j
            rec[i, 0:3] = r["v"]; rec[i, 3:6] = r["w"]; rec[i, 6] = r["k"]; rec[i, 7] = r["summary"]["iterations"]
        results.extend(out)
    return lo, rec, results


def gather_records(local, num_pairs, dist=None, device="cpu"):
    """The final gather: every rank receives the records of all pairs, in pair order.
    `dist` is torch.distributed (initialised) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        assert local.shape[0] == num_pairs
        return local.copy()
    import torch
    world = dist.get_world_size()
    cap = -(-num_pairs // world)                      # ceil: blocks differ by at most one pair
    buf = torch.zeros((cap, RECORD), dtype=torch.float64, device=device)
    buf[:local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    full = np.zeros((num_pairs, RECORD))
    for r in range(world):
        lo, hi = shard_range(num_pairs, r, world)
        full[lo:hi] = out[r][:hi - lo].cpu().numpy()
    return full


def max_over_ranks(value, dist=None, device="cpu"):
    """Elapsed time of a multi-GPU region = max over ranks (never wall clock of one rank)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
