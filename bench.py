#!/usr/bin/env python
"""Benchmark of the RJ-MCMC proposal-scoring hot path (contract: see the task brief / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload = BASELINE.json configs[3]: synthetic 100-gene, T=2000 nh-dbn, 1024 chains per GPU, fan-in 4.
One step = one thinning window (10 MCMC iterations) of every (node, chain) unit = one chain-kernel launch
(+ one all-reduce of the count matrices when N > 1).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GENES, T_SERIES, CHAINS_PER_GPU, FAN_IN, WINDOW = 100, 2000, 1024, 4, 10
DATA_SEED = 20260119
METHOD = 'varying_nh_dbn'
METRIC = 'mcmc_node_chain_iterations_per_sec'
UNIT = 'iterations/s'
# diagnostic only (never a bench value): DBN_BENCH_DIAG_NOREDUCE=1 drops the count all-reduce from the N > 1 step, to
# separate the collective's cost from everything else that changes with N; the JSON line then carries "diagnostic"
NCCL_OVERLAP_CTAS = int(os.environ.get('DBN_BENCH_NCCL_CTAS', '2'))
DIAG_NOREDUCE = os.environ.get('DBN_BENCH_DIAG_NOREDUCE') == '1'


def workload():
    return {'workload': 'synthetic nh-dbn: %d genes, T=%d, %d chains/GPU, fan-in %d (BASELINE configs[3])' % (
        N_GENES, T_SERIES, CHAINS_PER_GPU, FAN_IN),
        'n_genes': N_GENES, 'T': T_SERIES, 'chains_per_gpu': CHAINS_PER_GPU, 'fan_in': FAN_IN, 'method': 'nh-dbn',
        'iterations_per_step': WINDOW, 'parallelism': 'chains sharded over ranks, no data-path collective',
        'l2': 'inputs larger than L2 (prefix table 246 MB + boundary-row cache 162 MB vs 126 MB L2)'}


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------
# CPU arms (oracle port) — the only places bench.py executes oracle/
# ------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, node, n_iter = args
    from oracle import dbn_oracle as O
    from dynamicbayesiannetworks_b200.synth import make_synthetic
    data, _, _ = make_synthetic(N_GENES, T_SERIES, 3, 4, DATA_SEED)
    X, Y = O.lagged_matrices([data])
    t0 = time.perf_counter()
    N = X.shape[0]
    tr = O.run_chain(METHOD, X, Y, node, n_iter, O.RngDraws(np.random.default_rng(seed)), fan_in=FAN_IN,
                     init_tau=[T_SERIES // 4, T_SERIES // 2, 3 * T_SERIES // 4, N + 2])
    return n_iter, tr['n_evals'], time.perf_counter() - t0


def cpu_port_rate(budget_s=12.0, iters_per_task=40):
    """Gram-form NumPy oracle (oracle/dbn_oracle.py:run_chain) on the same workload, one (node, chain) unit per task,
    all host cores.  Returns (unit-iterations/s, evals/s, cores, description)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    for v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[v] = '1'      # one process per core; k x k BLAS calls gain nothing from threads
    ctx = mp.get_context('spawn')
    done_it = done_ev = 0
    t_start = time.perf_counter()
    with ctx.Pool(cores) as pool:
        rnd = 0
        while True:
            tasks = [(1000 + rnd * cores + j, (rnd * cores + j) % N_GENES, iters_per_task) for j in range(cores)]
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, tasks)
            done_it += sum(r[0] for r in res); done_ev += sum(r[1] for r in res)
            rnd += 1
            if rnd == 1:
                t_start = t0  # exclude pool start-up / first import from the rate
                t_first = time.perf_counter() - t0
                done_it0, done_ev0 = done_it, done_ev
            if time.perf_counter() - t_start > budget_s:
                break
    wall = time.perf_counter() - t_start
    sample = '%d units x %d iterations of the NumPy Gram-form oracle chain (nh-dbn, c4 data, chains started at pi={} and the 3 true changepoints; the GPU chains run at 8 segments on average), %d processes' % (
        rnd * cores, iters_per_task, cores)
    return done_it / wall, done_ev / wall, cores, sample


def cpu_literal_rate(budget_s=6.0):
    """The reference's own T_h x T_h formulation (oracle log_ml_literal) on c4-shaped states: evals/s."""
    from oracle import dbn_oracle as O
    from dynamicbayesiannetworks_b200.synth import make_synthetic
    data, _, _ = make_synthetic(N_GENES, T_SERIES, 3, 4, DATA_SEED)
    X, Y = O.lagged_matrices([data])
    N = X.shape[0]
    rng = np.random.default_rng(3)
    n_ev, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        node = int(rng.integers(0, N_GENES))
        pi = [int(v) for v in rng.choice([j for j in range(N_GENES) if j != node], size=3, replace=False)]
        tau = sorted(int(v) for v in rng.choice(np.arange(2, N + 1), size=3, replace=False)) + [N + 2]
        Xs, ys = O.segment_design(X, Y[:, node], pi, tau)
        O.log_ml_literal(Xs, ys, [np.zeros(4)] * 4, 0.005, 0.005, 1.0, N)
        n_ev += 1
    return n_ev / (time.perf_counter() - t0)


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    rates = []
    for _ in range(max(1, args.steps)):
        it_s, ev_s, cores, sample = cpu_port_rate(budget_s=8.0)
        rates.append((it_s, ev_s))
    it_s = float(np.mean([r[0] for r in rates])); ev_s = float(np.mean([r[1] for r in rates]))
    line = {'impl': 'reference', 'metric': METRIC, 'value': it_s, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': None, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic', 'config': workload(), 'ml_evals_per_sec': ev_s,
            'cpu_baseline': {'value': it_s, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': it_s, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        args.steps = min(args.steps, 3)
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from dynamicbayesiannetworks_b200.engine import DbnEngine
    from dynamicbayesiannetworks_b200.synth import make_synthetic
    from dynamicbayesiannetworks_b200 import parallel

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # default group: 2 CTAs, for the all-reduce that overlaps the chain kernel; `wide`: NCCL defaults, for the
        # end-to-end read that waits for the reduced counts
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank),
                                pg_options=parallel.nccl_small_footprint_options(NCCL_OVERLAP_CTAS))
        wide = dist.new_group(backend='nccl')
    K, W = args.steps, max(3, args.warmup)

    data, _, _ = make_synthetic(N_GENES, T_SERIES, 3, 4, DATA_SEED)
    data = torch.from_numpy(data).pin_memory().numpy()      # the host series lives in pinned memory (e2e copies from it)
    eng = DbnEngine(n_chains=CHAINS_PER_GPU, device=local_rank, seed=0xDB20B200, chain_offset=rank * CHAINS_PER_GPU)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    eng.build_stats([data])
    eng.chain_init(METHOD, fan_in=FAN_IN, burn_in=0)
    local_counts = parallel.counts_tensor(eng)
    global_counts = [torch.empty_like(local_counts), torch.empty_like(local_counts)]
    pending = [None, None]
    units = N_GENES * CHAINS_PER_GPU
    step_no = [0]

    def step():
        """One thinning window on every unit, then the all-reduce of the count block over the ranks.  The block is
        snapshotted on the kernel's stream and reduced on NCCL's stream (async_op), so the next window's kernel does
        not wait for the collective; a snapshot buffer is reused only after its reduction has finished."""
        eng.run(WINDOW)
        if world > 1 and not DIAG_NOREDUCE:
            j = step_no[0] & 1
            step_no[0] += 1
            if pending[j] is not None:
                pending[j].wait()
            _, pending[j] = parallel.allreduce_counts(local_counts, out=global_counts[j], async_op=True)

    def drain():
        for j in (0, 1):
            if pending[j] is not None:
                pending[j].wait()
                pending[j] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    drain()
    barrier()
    eng.reset_counters()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(K):
        step()
    drain()          # the last reductions are inside the timed region
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    ctr = eng.counters()
    if world > 1:
        t = torch.tensor([ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        c = torch.tensor([ctr['ml_evals'], ctr['algorithmic_bytes'], ctr['algorithmic_flops'], ctr['segment_evals']],
                         device='cuda', dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        tot_evals = float(c[0].item())
    else:
        tot_evals = float(ctr['ml_evals'])
    value = units * world * WINDOW * K / (ms * 1e-3)

    # ---- end-to-end through the C ABI with HOST buffers: per step, H2D of the series + prefix build + one window +
    # D2H of the count matrices (the chains carry on across steps, so the work per step equals the device-timed one)
    Ke = max(5, min(K, 40))
    h2d = data.nbytes
    if world > 1:           # untimed: the first collective of a group sets up its communicator
        parallel.allreduce_counts(local_counts, out=global_counts[0], group=wide)
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        eng.build_stats([data])
        eng.run(WINDOW)
        if world > 1:       # the caller reads the counts of ALL chains: reduce over the ranks, then device -> host
            g = parallel.allreduce_counts(local_counts, out=global_counts[0], group=wide)
            host_counts = g.cpu().numpy()
            d2h = host_counts.nbytes
        else:
            cnt = eng.counts()
            d2h = sum(v.nbytes for v in cnt.values())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = units * world * WINDOW * Ke / e2e_s

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    per_launch_bytes = ctr['algorithmic_bytes'] / K
    avg_launch_s = ms * 1e-3 / K
    achieved = per_launch_bytes / avg_launch_s / 1e9
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get('chain_kernel_dram_bytes_per_launch')
        except Exception:
            traffic = None
    fp64_peak = eng.fp64_peak_gflops()
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': workload(),
        'ml_evals_per_sec': tot_evals / (ms * 1e-3),
        'ml_evals_per_iteration': ctr['ml_evals'] / max(1, ctr['unit_iterations']),
        'accept_rate': {'pi': ctr['pi_accepts'] / max(1, ctr['unit_iterations']),
                        'tau': ctr['tau_accepts'] / max(1, ctr['unit_iterations'])},
        'segment_factorisations_per_iteration': ctr['segment_factorisations'] / max(1, ctr['unit_iterations']),
        'roofline': {'kernel': 'chain_fast_kernel<nh-dbn, KM=5>', 'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'],
                     'unit': 'GB/s', 'frac': achieved / peaks['hbm_gbs'], 'traffic': traffic, 'peak_source': peak_kind,
                     'algorithmic_bytes_per_launch': per_launch_bytes,
                     'fp64': {'achieved_gflops': ctr['algorithmic_flops'] / K / avg_launch_s / 1e9,
                              'peak_gflops_measured_dfma': fp64_peak}},
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'steps': Ke, 'what': 'dbn_build_stats(host series) + dbn_run(10 iterations) + (N > 1: all-reduce of the counts) + counts to host, per step'},
        'gpu_launches': K * (1 if world == 1 else 1),
        'clocks': clocks,
    }
    if DIAG_NOREDUCE:
        line['diagnostic'] = 'count all-reduce skipped: NOT a bench value'
    if not args.no_cpu_baseline and world == 1:
        it_s, ev_s, cores, sample = cpu_port_rate()
        line['cpu_baseline'] = {'value': it_s, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample,
                                'ml_evals_per_sec': ev_s,
                                'literal_T_h_x_T_h_form_ml_evals_per_sec_1core': cpu_literal_rate()}
    else:
        line['cpu_baseline'] = None
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
