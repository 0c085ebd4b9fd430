"""torchrun check of the multi-GPU Network path: every rank must return the adjacency matrix of an unsharded run.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_network_check.py"""
import glob
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamicbayesiannetworks_b200 import Network   # noqa: E402

g = dict(np.load(glob.glob(os.path.join(ROOT, 'tests', 'golden', 'chain_mtor_varying_nh_dbn.npz'))[0]))
data = [g['data_seg%d' % s] for s in range(int(g['n_data_segments']))]
rank = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
import dynamicbayesiannetworks_b200.parallel as P
CASES = [(m, 64) for m in ('varying_nh_dbn', 'seq_coup_nh_dbn', 'glob_coup_nh_dbn', 'var_glob_coup_nh_dbn', 'fp_glob_coup_nh_dbn',
                           'fp_varying_nh_dbn', 'fp_var_glob_coup_nh_dbn', 'fp_seq_coup_nh_dbn')]
CASES += [('varying_nh_dbn', 1), ('fp_glob_coup_nh_dbn', 1), ('glob_coup_nh_dbn', 3)]   # the flat unit list is split, one chain included
for method, n_chains in CASES:
    n_iter = 60 if method.startswith('fp') else 400
    # device=None: the rank's own GPU is picked from LOCAL_RANK (parallel.default_device)
    sharded = Network(data, n_iter, 100 if n_iter > 100 else 10, 1, n_chains=n_chains, keep_chain=False)
    sharded.infer_network(method)
    adj = np.array(sharded.proposed_adj_matrix)
    # reference: the same chains unsharded, on this rank's GPU (the Network is told there is no process group)
    saved = P.dist_rank_world
    P.dist_rank_world = lambda: (0, 1)
    single = Network(data, n_iter, 100 if n_iter > 100 else 10, 1, n_chains=n_chains, device=rank, keep_chain=False)
    single.infer_network(method)
    P.dist_rank_world = saved
    same = np.array_equal(adj, np.array(single.proposed_adj_matrix))
    print('rank %d %-24s %3d chains  sharded == unsharded: %s' % (dist.get_rank(), method, n_chains, same), flush=True)
    assert same
dist.barrier()
dist.destroy_process_group()
