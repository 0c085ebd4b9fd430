#!/usr/bin/env python
"""BASELINE configs[4] (500 genes, T = 10000, fp-glob-dbn, 9 chains per node = 4500 units) sharded over the ranks as flat
(node, chain) unit ranges: STRONG scaling of one iteration sweep.  Every rank builds the (replicated) prefix table, runs
its unit range; time = max over ranks (CUDA events), rank 0 prints one JSON line.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_c5_dist.py [iterations=2]
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamicbayesiannetworks_b200.engine import DbnEngine          # noqa: E402
from dynamicbayesiannetworks_b200.parallel import shard_units      # noqa: E402
from dynamicbayesiannetworks_b200.synth import make_synthetic     # noqa: E402

N_GENES, T, CHAINS = 500, 10000, 9


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
    data, cps, _ = make_synthetic(N_GENES, T, n_changepoints=5, max_fan_in=4, seed=20260120)
    first, count = shard_units(N_GENES * CHAINS, rank, world)
    eng = DbnEngine(n_chains=CHAINS, device=rank, seed=11, unit_range=(first, count))
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    eng.set_stream(s.cuda_stream)
    eng.build_stats([data])
    eng.chain_init('fp_glob_coup_nh_dbn')
    eng.set_changepoints([c + 1 for c in cps] + [eng.N + 2])
    eng.run_fp(1)
    torch.cuda.synchronize()
    eng.reset_counters()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.run_fp(iters)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device='cuda')
    c = eng.counters()
    flops = torch.tensor([float(c['algorithmic_flops'])], device='cuda')
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(flops, op=dist.ReduceOp.SUM)
    if rank == 0:
        units = N_GENES * CHAINS
        print(json.dumps(dict(case='c5 synthetic 500 genes, T=10000, fp-glob-dbn, %d units over %d GPUs (strong scaling)' % (units, world),
                              n_gpus=world, units=units, units_rank0=count, iterations=iters, ms_per_iteration=ms.item() / iters,
                              unit_iterations_per_sec=units * iters / ms.item() * 1e3,
                              algorithmic_tflops=flops.item() / ms.item() / 1e9)), flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
