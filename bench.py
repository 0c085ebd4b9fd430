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
BURN_IN_ITERATIONS = 200      # the timed region starts after 200 iterations, whatever --warmup says
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
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))

    def stop(self, t0=None, t1=None):
        """Summary of the samples taken in [t0, t1] (the timed region); nvidia-smi needs ~100 ms to deliver its first
        sample, so the sampler is started before the warm-up and, if the timed region was shorter than one sampling
        period, the samples taken under the identical warm-up load are used and the line says so."""
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        rows = [r for ts, r in self.rows if t0 is None or (t0 <= ts <= t1 + 0.02)]
        window = 'timed region'
        if not rows:
            rows, window = [r for _, r in self.rows], 'warm-up + timed region (same load)'
        for r in rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'samples': len(sm), 'reasons': sorted(reasons), 'window': window}


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


def _import_dyban_ref():
    """The UNMODIFIED reference installed by `pip install --target baseline/_ref` (git-ignored, travels to the GPU box),
    with the two import shims of SURVEY Appendix E (matplotlib stub, np.asscalar).  None when it is not there."""
    import tempfile
    import types
    ref = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref, 'dyban')):
        return None
    for m in ('matplotlib', 'matplotlib.pyplot'):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    if not hasattr(np, 'asscalar'):
        np.asscalar = lambda a: np.asarray(a).item()
    if ref not in sys.path:
        sys.path.insert(0, ref)
    os.environ.setdefault('TQDM_DISABLE', '1')
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix='dyban_ref_'))     # dyban/scores.py opens ./output.log at import time
    try:
        import dyban
        import dyban.network  # noqa: F401
    except Exception:
        return None
    finally:
        os.chdir(cwd)
    return dyban


def _literal_worker(task):
    """One node of a literal-reference run: (kind, node, n_iter) -> (unit-iterations, log-ML evaluations, seconds)."""
    import contextlib
    import io
    import random
    kind, node, n_iter = task
    dyban = _import_dyban_ref()
    import dyban.moves as mv
    import dyban.marginalLikelihood as ml
    calls = [0]
    for name in ('calculateMarginalLikelihoodWithChangepoints', 'calculateMarginalLikelihood'):
        orig = getattr(mv, name)
        def counted(*a, _o=orig, **k):
            calls[0] += 1
            return _o(*a, **k)
        setattr(mv, name, counted)
    if kind == 'c4':
        # BASELINE.md line B: the literal T_h x T_h code cannot evaluate T >= 344 (gamma overflows, pi ** (-T / 2)
        # underflows, marginalLikelihood.py:116-118); guard exactly those scalars, keep all of its O(T_h^3) NumPy work
        import math as _m

        class _Math:
            pi = _m.pi

            @staticmethod
            def log(x):
                x = float(x)
                return _m.log(x) if 0.0 < x < float('inf') else 0.0

            @staticmethod
            def exp(x):
                x = float(x)
                return _m.exp(min(x, 700.0)) if x == x else 1.0

            def __getattr__(self, name):
                return getattr(_m, name)
        ml.math = _Math()
        ml.gamma = lambda x: 1.0
        mv.math = _Math()
        from dynamicbayesiannetworks_b200.synth import make_synthetic
        data = [make_synthetic(N_GENES, T_SERIES, 3, 4, DATA_SEED)[0]]
        method = 'varying_nh_dbn'
    else:
        g = np.load(os.path.join(ROOT, 'tests', 'golden', 'chain_%s_varying_nh_dbn.npz' % ('yeast' if kind == 'c2' else 'mtor')))
        data = [g['data_seg%d' % s_] for s_ in range(int(g['n_data_segments']))]
        method = 'varying_nh_dbn' if kind == 'c2' else 'seq_coup_nh_dbn'
    np.random.seed(100 + node); random.seed(100 + node)
    net = dyban.Network(data, n_iter, 0, 1)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all='ignore'):
        net.set_network_configuration(node)
        net.fit(method)
    return n_iter, calls[0], time.perf_counter() - t0


def literal_reference_rates():
    """BASELINE.md lines A and B: dyban's OWN code (baseline/_ref) timed on the host cores.  A: unmodified, on the shipped
    yeast (c2, nh-dbn) and mTOR (c3, seq-dbn) series, one node per process.  B: the c4 series with the overflow guards.
    Returns a dict for the JSON line, or {'unavailable': why}."""
    import multiprocessing as mp
    if _import_dyban_ref() is None:
        return {'unavailable': 'baseline/_ref/dyban is not installed (pip install --target baseline/_ref /root/reference)'}
    cores = os.cpu_count() or 1
    for v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[v] = '1'
    out = {'cores': cores, 'what': "dyban's own regressors (unmodified install under baseline/_ref), one node per process"}
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores) as pool:
        for kind, n_nodes, n_iter in (('c2', 5, 400), ('c3', 11, 60), ('c4', min(cores, 8), 3)):
            t0 = time.perf_counter()
            res = pool.map(_literal_worker, [(kind, node, n_iter) for node in range(n_nodes)])
            wall = time.perf_counter() - t0
            out[kind] = {'unit_iterations_per_sec': sum(r[0] for r in res) / wall, 'ml_evals_per_sec': sum(r[1] for r in res) / wall,
                         'units': n_nodes, 'iterations': n_iter, 'wall_s': wall,
                         'series': {'c2': 'yeast nh-dbn (N = 33)', 'c3': 'mTOR seq-dbn (N = 119)',
                                    'c4': 'synthetic nh-dbn (N = 1999), gamma / pi ** (-T / 2) guarded'}[kind]}
    return out


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
            'gpu_launches': 0,
            'note': ("value = the Gram-form NumPy port of the path (oracle/dbn_oracle.py), which is ~1000x faster than dyban's own "
                     "T_h x T_h code on this workload; dyban's own code is timed in 'literal_reference' (BASELINE.md lines A, B)"),
            'literal_reference': literal_reference_rates()}
    print(json.dumps(line))


def c5_sample(device, fp64_peak):
    """BASELINE configs[4] (500 genes, T = 10000, full-parents globally coupled, design width k = 500) on one GPU: a bounded
    sample -- the first 592 of the 4500 (node, chain) units of the 9-chain workload (four full waves of one block per SM),
    started at the true changepoints, 1 untimed + 2 timed iterations -- of the one contraction-shaped kernel of the path.  Roofline: FP64 (k^3/3 + 4k^2 flops per segment evaluation, counted on
    the device), against the DFMA peak measured in this run."""
    import torch
    from dynamicbayesiannetworks_b200.engine import DbnEngine
    from dynamicbayesiannetworks_b200.synth import make_synthetic
    n, T = 500, 10000
    data, cps, _ = make_synthetic(n, T, n_changepoints=5, max_fan_in=4, seed=20260120)
    chains, units = 9, 592
    eng = DbnEngine(n_chains=chains, device=device, seed=11, unit_range=(0, units))
    try:
        eng.set_stream(torch.cuda.current_stream().cuda_stream)
        eng.build_stats([data])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.rebuild_stats(); e1.record(); torch.cuda.synchronize()
        prefix_ms = e0.elapsed_time(e1)
        info = eng.stats_info()
        eng.chain_init('fp_glob_coup_nh_dbn')
        eng.set_changepoints([c + 1 for c in cps] + [eng.N + 2])
        eng.run_fp(1)
        torch.cuda.synchronize()
        eng.reset_counters()
        iters = 2
        e0.record(); eng.run_fp(iters); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        c = eng.counters()
    finally:
        eng.close()
    gf = c['algorithmic_flops'] / ms / 1e6
    gf_exec = c['executed_level3_flops'] / ms / 1e6
    return {'workload': 'synthetic fp-glob-dbn: 500 genes, T=10000, %d chains per node (BASELINE configs[4]): units 0..%d of %d, %d timed iterations'
                        % (chains, units - 1, n * chains, iters),
            'metric': METRIC, 'value': units * iters / ms * 1e3, 'unit': UNIT, 'ms_per_iteration': ms / iters,
            'ml_evals_per_sec': c['ml_evals'] / ms * 1e3,
            'roofline': {'kernel': 'fp_glob_kernel<glob, BIG>', 'bound': 'fp64', 'achieved': gf / 1e3, 'peak': fp64_peak / 1e3,
                         'unit': 'TFLOP/s', 'frac': gf / fp64_peak, 'peak_source': 'DFMA microbenchmark in this run',
                         'executed_tflops': gf_exec / 1e3, 'executed_frac': gf_exec / fp64_peak,
                         'executed_note': 'k^3/3 per Cholesky factor executed, plus 2 k^3/3 where the inverse factor and I - A^-1 are formed '
                                          '(the muSampler precision is an explicit k x k matrix); frac counts SURVEY 8d flops only',
                         'algorithmic_gbs': c['algorithmic_bytes'] / ms / 1e6},
            'prefix_build': {'table_gb': info['algorithmic_bytes'] / 1e9, 'ms': prefix_ms, 'gbs': info['algorithmic_bytes'] / prefix_ms / 1e6}}


# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak (the contract): 1024 chains per GPU; strong: 1024 chains in total, the flat (node, chain) unit list split over the ranks')
    ap.add_argument('--no-c5', action='store_true', help='skip the BASELINE configs[4] sample (extra.c5)')
    ap.add_argument('--handles', type=int, default=2,
                    help='library handles per GPU for the device-resident measurement: each takes a contiguous range of the GPU\'s '
                         '(node, chain) units and its own stream, so the launches of one range fill the partial last wave of the '
                         'other (102 400 units are 2.7 waves of resident blocks; measured +12 %%, scripts/two_stream_ab.py). 1 = one handle')
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
        wide = dist.new_group(backend='nccl')   # (a high-priority NCCL stream for this group was measured at N = 2: no difference)
    K = args.steps
    W = max(3, args.warmup, (BURN_IN_ITERATIONS + WINDOW - 1) // WINDOW)   # >= 200 untimed iterations: a stationary timed region

    data, _, _ = make_synthetic(N_GENES, T_SERIES, 3, 4, DATA_SEED)
    data = torch.from_numpy(data).pin_memory().numpy()      # the host series lives in pinned memory (e2e copies from it)
    strong = args.scaling == 'strong' and world > 1
    if strong:      # 1024 chains in total, rank r takes a contiguous slice of the 102 400 units
        first, count = parallel.shard_units(N_GENES * CHAINS_PER_GPU, rank, world)
        eng = DbnEngine(n_chains=CHAINS_PER_GPU, device=local_rank, seed=0xDB20B200, unit_range=(first, count))
    else:
        eng = DbnEngine(n_chains=CHAINS_PER_GPU, device=local_rank, seed=0xDB20B200, chain_offset=rank * CHAINS_PER_GPU)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    eng.build_stats([data])
    eng.chain_init(METHOD, fan_in=FAN_IN, burn_in=0)
    local_counts = parallel.counts_tensor(eng)
    global_counts = [torch.empty_like(local_counts), torch.empty_like(local_counts)]
    units = N_GENES * CHAINS_PER_GPU
    total_units = units if strong else units * world
    step_no = [0]

    # ---- the handles of the device-resident loop.  `eng` (the whole range of this GPU on `stream`) serves the end-to-end
    # measurement; with --handles H > 1 the timed loop runs H handles of its own, each on a contiguous sub-range of this GPU's
    # units and on its own stream.  Units are independent and the Philox counters are keyed by (chain, node), so the chains are
    # the same chains; what changes is that consecutive windows of different ranges overlap on the device.
    class Part:
        pass
    parts = []
    H = max(1, args.handles)
    units_here = count if strong else units
    if units_here <= 148 * 2 * 128:     # at most one wave of resident blocks (2 x 128 threads per SM): nothing to overlap, one handle
        H = 1
    for p_i in range(H):
        pt = Part()
        if H == 1:
            pt.eng, pt.stream = eng, stream
        else:
            base_first, base_count = (first, count) if strong else (0, units)
            f_, c_ = parallel.shard_units(base_count, p_i, H)
            pt.eng = DbnEngine(n_chains=CHAINS_PER_GPU, device=local_rank, seed=0xDB20B200,
                               chain_offset=0 if strong else rank * CHAINS_PER_GPU, unit_range=(base_first + f_, c_))
            pt.stream = torch.cuda.Stream()
            pt.eng.set_stream(pt.stream.cuda_stream)
            pt.eng.build_stats([data])
            pt.eng.chain_init(METHOD, fan_in=FAN_IN, burn_in=0)
        pt.local = local_counts if H == 1 else parallel.counts_tensor(pt.eng)
        pt.glob = global_counts if H == 1 else [torch.empty_like(pt.local), torch.empty_like(pt.local)]
        pt.pending = [None, None]
        parts.append(pt)

    def step():
        """One thinning window on every unit, then the all-reduce of the count block over the ranks.  The block is
        snapshotted on the kernel's stream and reduced on NCCL's stream (async_op), so the next window's kernel does
        not wait for the collective; a snapshot buffer is reused only after its reduction has finished."""
        j = step_no[0] & 1
        step_no[0] += 1
        for pt in parts:
            with torch.cuda.stream(pt.stream):
                pt.eng.run(WINDOW)
                if world > 1 and not DIAG_NOREDUCE:
                    if pt.pending[j] is not None:
                        pt.pending[j].wait()
                    _, pt.pending[j] = parallel.allreduce_counts(pt.local, out=pt.glob[j], async_op=True)

    def drain():
        for pt in parts:
            with torch.cuda.stream(pt.stream):
                for j in (0, 1):
                    if pt.pending[j] is not None:
                        pt.pending[j].wait()
                        pt.pending[j] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(W):
        step()
    drain()
    barrier()
    for pt in parts:
        pt.eng.reset_counters()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in parts]
    barrier()
    t_begin = time.perf_counter()
    ev0.record()
    for pt in parts:
        pt.stream.wait_event(ev0)
    for _ in range(K):
        step()
    drain()          # the last reductions are inside the timed region
    for pt, e1 in zip(parts, ev1):
        e1.record(pt.stream)
    barrier()
    t_end = time.perf_counter()
    ms = max(ev0.elapsed_time(e1) for e1 in ev1)   # the device time of the slowest handle's stream
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    ctr = {}
    for pt in parts:
        for k_, v_ in pt.eng.counters().items():
            ctr[k_] = ctr.get(k_, 0) + v_
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
    value = total_units * WINDOW * K / (ms * 1e-3)

    # ---- end-to-end through the C ABI with HOST buffers: per step and per handle, H2D of the series + prefix build + one
    # window + D2H of the count matrices (the chains carry on across steps, so the work per step equals the device-timed one).
    # The handles are the ones of the device-resident loop (each its own table, its own unit range, its own stream) and every
    # handle's step is strictly sequential (build -> run -> read, the host waits for the read before it queues that handle's
    # next build); the host queues one handle's next step before it waits for the other's read, so the tail of one handle's
    # window overlaps the other handle's copies, prefix build and first wave -- the way a caller drives two handles per device
    # (INTEGRATION.md section 3).  h2d / d2h count every handle's copies.
    Ke = max(5, min(K, 40))
    h2d = data.nbytes * len(parts)
    for pt in parts:
        pt.host = torch.empty(pt.local.shape, dtype=pt.local.dtype, pin_memory=True)
        pt.host_flat = pt.host.numpy().reshape(-1)
    if world > 1:           # untimed: the first collective of a group sets up its communicator
        parallel.allreduce_counts(parts[0].local, out=parts[0].glob[0], group=wide)

    # the reduce the caller waits for runs on the wide group (NCCL's default CTA count); the 2-CTA group of the overlapped
    # all-reduce was measured here too (DBN_E2E_GROUP=small, N = 2, two handles): 1.15e9 against 1.24e9
    e2e_group = (None if os.environ.get('DBN_E2E_GROUP', 'wide') == 'small' else wide) if world > 1 else None

    def e2e_queue(pt):
        with torch.cuda.stream(pt.stream):
            pt.eng.build_stats([data])
            pt.eng.run(WINDOW)
            if world > 1:   # the caller reads the counts of ALL chains on rank 0: reduce over the ranks, then device -> pinned host
                pt.glob[0].copy_(pt.local)
                dist.reduce(pt.glob[0], dst=0, op=dist.ReduceOp.SUM, group=e2e_group)
                if rank == 0:
                    pt.host.copy_(pt.glob[0], non_blocking=True)

    def e2e_read(pt):
        if world > 1:
            if rank == 0:
                pt.stream.synchronize()
            return pt.host.numel() * pt.host.element_size()
        cnt = pt.eng.counts(out=pt.host_flat)      # pinned destination: written by the copy engine directly
        return sum(v.nbytes for v in cnt.values())

    barrier()
    t0 = time.perf_counter()
    for s_ in range(Ke):
        d2h = 0
        for pt in parts:
            if s_ > 0:
                d2h += e2e_read(pt)     # the result of this handle's previous step
            e2e_queue(pt)
    d2h = sum(e2e_read(pt) for pt in parts)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = total_units * WINDOW * Ke / e2e_s
    if H > 1:
        for pt in parts:
            pt.eng.close()
        parts = []

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    per_launch_bytes = ctr['algorithmic_bytes'] / K
    avg_launch_s = ms * 1e-3 / K
    achieved = per_launch_bytes / avg_launch_s / 1e9
    traffic = issue_active = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                tj = json.load(f)
            traffic = tj.get('chain_kernel_dram_bytes_per_launch')
            issue_active = tj.get('chain_kernel_issue_active_pct')
        except Exception:
            traffic = None
    fp64_peak = eng.fp64_peak_gflops()
    cfg = workload()
    cfg['handles_per_gpu'] = H
    if H > 1:
        cfg['handles'] = ('%d library handles per GPU, each a contiguous range of the GPU\'s (node, chain) units on its own stream: the '
                          'launches of one range fill the partial last wave of the other (device-resident loop and e2e alike)' % H)
    if strong:
        cfg['parallelism'] = 'flat (node, chain) unit list split over ranks (1024 chains in total), no data-path collective'
        cfg['chains_total'] = CHAINS_PER_GPU
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'strong' if strong else 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': cfg,
        'ml_evals_per_sec': tot_evals / (ms * 1e-3),
        'ml_evals_per_iteration': ctr['ml_evals'] / max(1, ctr['unit_iterations']),
        'accept_rate': {'pi': ctr['pi_accepts'] / max(1, ctr['unit_iterations']),
                        'tau': ctr['tau_accepts'] / max(1, ctr['unit_iterations'])},
        'segment_factorisations_per_iteration': ctr['segment_factorisations'] / max(1, ctr['unit_iterations']),
        'burn_in_iterations': W * WINDOW,
        # bound: the nearest roofline of the two the contract names.  frac = ALGORITHMIC bytes (16 (k(k+1)/2 + k + 1)
        # per executed segment evaluation, counted on the device) / time / measured HBM peak.  The kernel does not move
        # those bytes: unchanged segments come from the per-unit boundary-row cache, so the DRAM traffic ncu measures
        # (traffic, dram_frac) is lower, and what limits the kernel is dependent-issue latency at 8 warps per SM (255 registers per thread) plus
        # instruction delivery (limiter, issue_active_pct; profiles/r2_notes.md).
        'roofline': {'kernel': 'chain_fast_kernel<nh-dbn, KM=5>', 'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'],
                     'unit': 'GB/s', 'frac': achieved / peaks['hbm_gbs'], 'traffic': traffic, 'peak_source': peak_kind,
                     'dram_frac': (traffic / avg_launch_s / 1e9 / peaks['hbm_gbs']) if traffic else None,
                     'limiter': 'latency / issue (FP64 dependency chains at 8 warps per SM, ~100 KB loop body), not HBM',
                     'issue_active_pct': issue_active,
                     'algorithmic_bytes_per_launch': per_launch_bytes,
                     'per_launch_means': 'one window of all units of the GPU = one launch per handle (%d); traffic was captured on the one-handle launch of the same units' % H,
                     'fp64': {'achieved_gflops': ctr['algorithmic_flops'] / K / avg_launch_s / 1e9,
                              'peak_gflops_measured_dfma': fp64_peak}},
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'steps': Ke, 'handles': H,
                'what': 'per step and per handle (%d per GPU, each a unit range with its own table and stream): dbn_build_stats(pinned host series) + '
                        'dbn_run(10 iterations) + (N > 1: reduce of the counts to rank 0) + counts to pinned host; a handle\'s step is sequential, '
                        'the host queues one handle\'s next step before waiting for the other\'s read' % H},
        'gpu_launches': K * H,
        'handles_per_gpu': H,
        'clocks': clocks,
    }
    if DIAG_NOREDUCE:
        line['diagnostic'] = 'count all-reduce skipped: NOT a bench value'
    if world == 1 and not args.no_c5:
        eng.close()
        del local_counts, global_counts
        torch.cuda.empty_cache()
        try:
            line['extra'] = {'c5': c5_sample(local_rank, fp64_peak)}
        except Exception as exc:      # the sample must never cost the contract line
            line['extra'] = {'c5': {'unavailable': '%s: %s' % (type(exc).__name__, exc)}}
    if not args.no_cpu_baseline and world == 1:
        it_s, ev_s, cores, sample = cpu_port_rate()
        line['cpu_baseline'] = {'value': it_s, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample,
                                'ml_evals_per_sec': ev_s,
                                'literal_T_h_x_T_h_form_ml_evals_per_sec_1core': cpu_literal_rate(),
                                'note': ("kind 'port' = the Gram-form NumPy restatement, ~1000x faster than dyban's own T_h x T_h "
                                         "code on this workload; dyban's own code is in 'literal_reference'"),
                                'literal_reference': literal_reference_rates()}
    else:
        line['cpu_baseline'] = None
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
