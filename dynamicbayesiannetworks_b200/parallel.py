"""Multi-GPU plumbing: one process per GPU, the flat (node, chain) unit list split over the ranks (`shard_units`; the
weak-scaling bench keeps every node on every rank and splits the chains, `shard_chains`), one all-reduce of the count
matrices per thinning window.

Nodes are independent regressions (network.py:413-416) and chains are independent by construction, so the sampling
itself needs no collective.  Only the edge-count / changepoint-histogram block (engine.counts()) is summed over
ranks.  torch.distributed is used for the rendezvous and the NCCL (GPU) / gloo (CPU tests) all-reduce.
"""
import numpy as np


def shard_chains(total_chains, rank, world_size):
    """Contiguous split of the chain index space: returns (chain_offset, n_local).  Every rank keeps all nodes, so
    the prefix table is replicated and unit u = node * n_local + chain_local; Philox counters use the GLOBAL chain
    id (chain_offset + chain_local), which makes the union of the shards identical to an unsharded run."""
    base, rem = divmod(int(total_chains), int(world_size))
    n_local = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, n_local


def shard_units(total_units, rank, world_size):
    """Contiguous split of the flat unit list u = node * n_chains + chain: returns (first_unit, n_local).
    Works for any number of chains, one included (the reference's own use: network.py:413-416 loops over the nodes of a
    single chain), and splits e.g. 500 nodes x 9 chains into 563 / 562 units per rank on 8 ranks.  Philox counters are
    keyed by the grid's (chain, node), so the union of the shards is identical to an unsharded run."""
    base, rem = divmod(int(total_units), int(world_size))
    n_local = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n_local


def default_device():
    """CUDA device of this process when the caller names none: under torch.distributed (one process per GPU) the
    process's LOCAL_RANK / current torch device, so that the ranks of a sharded run do not all land on cuda:0; else 0."""
    import os
    rank, world = dist_rank_world()
    if world > 1:
        if 'LOCAL_RANK' in os.environ:
            return int(os.environ['LOCAL_RANK'])
        try:
            import torch
            if torch.cuda.is_available():
                return int(torch.cuda.current_device())
        except ImportError:
            pass
        return rank
    return 0


class _DevicePtr:
    """Minimal __cuda_array_interface__ view of library-owned device memory (no copy)."""

    def __init__(self, ptr, n_elems, typestr='<i8'):
        self.__cuda_array_interface__ = {'shape': (int(n_elems),), 'typestr': typestr, 'data': (int(ptr), False),
                                         'version': 2}


def counts_tensor(engine):
    """torch int64 view (device memory of the engine) of the contiguous count block."""
    import torch
    ptr, n = engine.counts_device_ptr()
    return torch.as_tensor(_DevicePtr(ptr, n), device='cuda:%d' % engine.device)


def nccl_small_footprint_options(max_ctas=2):
    """ProcessGroupNCCL options for the group whose all-reduce OVERLAPS the chain kernel.  The only collective on the
    path is a 1.7 MB all-reduce per thinning window, running beside the next window's (issue-bound, all-SM) chain
    kernel; with NCCL's default CTA count it costs 8 % of the step (1.79 vs 1.66 ms at N = 2 and N = 8), with two
    CTAs 1.3 % (`profiles/r1_v13/nccl_diag.txt`).  A caller that WAITS for the reduced counts (the
    end-to-end read) should use a second group with default options (`dist.new_group()`): alone on the GPU the wide
    all-reduce is the faster one."""
    import torch.distributed as dist
    opts = dist.ProcessGroupNCCL.Options()
    opts.config.max_ctas = int(max_ctas)
    opts.config.min_ctas = 1
    return opts


def allreduce_counts(local, out=None, async_op=False, group=None):
    """Sum the count block over all ranks.  `local` is a torch tensor (CUDA for NCCL, CPU for gloo); the sum is
    written to `out` (a copy) so that the local accumulators keep counting only their own shard."""
    import torch
    import torch.distributed as dist
    if out is None:
        out = torch.empty_like(local)
    out.copy_(local)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        work = dist.all_reduce(out, op=dist.ReduceOp.SUM, async_op=async_op, group=group)
        if async_op:
            return out, work
    return (out, None) if async_op else out


def dist_rank_world():
    """(rank, world_size) of the default process group, (0, 1) without one."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def global_counts(engine):
    """All-reduce the engine's count block over the ranks (NCCL, on device memory) and return it split into the named
    matrices, like DbnEngine.counts()."""
    from .engine import split_counts
    engine.sync()                 # the chain kernels ran on the engine's own stream
    total = allreduce_counts(counts_tensor(engine))
    return split_counts(total.cpu().numpy().astype(np.int64), engine.n, engine.N)


def edge_scores_from_counts(edge, nonempty):
    """proposed_adj_matrix: edge[i][j] / #kept non-empty parent sets of node i (scores.py:184-192).  A node whose
    every kept sample had an empty parent set makes the reference divide by zero; 0.0 is returned here."""
    edge = np.asarray(edge, dtype=np.float64)
    den = np.asarray(nonempty, dtype=np.float64)[:, None]
    return np.divide(edge, den, out=np.zeros_like(edge), where=den > 0)


def credible_scores_from_counts(pos, neg, kept):
    """proposed_adj_matrix of the full-parents globally coupled model: max(#(mu_j >= 0), #(mu_j <= 0)) / #kept samples
    (credible_score, scores.py:196-205, applied to the thinned mu chain by network.py:339-347); diagonal = 0."""
    pos = np.asarray(pos, dtype=np.float64)
    neg = np.asarray(neg, dtype=np.float64)
    den = np.asarray(kept, dtype=np.float64)[:, None]
    out = np.divide(np.maximum(pos, neg), den, out=np.zeros_like(pos), where=den > 0)
    np.fill_diagonal(out, 0.0)
    return out
