"""Does a thin kernel get SM room BESIDE the resident blocks of the h / nh chain kernel (2 x 128 threads x 240 registers per SM leave
4096 registers), or does it wait for chain blocks to retire?  A torch fill_ (128-thread blocks) of the table's size is timed alone
and while a chain window is running on another stream."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from dynamicbayesiannetworks_b200.engine import DbnEngine
from dynamicbayesiannetworks_b200.synth import make_synthetic
data, _, _ = make_synthetic(100, 2000, 3, 4, 20260119)
e = DbnEngine(n_chains=1024)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
e.set_stream(sb.cuda_stream)
e.build_stats([data]); e.chain_init('varying_nh_dbn', fan_in=4, burn_in=0)
for _ in range(25): e.run(10)
torch.cuda.synchronize()
x = torch.empty(246_000_000 // 8, dtype=torch.float64, device='cuda')
y = torch.empty(1 << 20, dtype=torch.float64, device='cuda')
ev = lambda: torch.cuda.Event(enable_timing=True)
def timed(fn, busy):
    ts = []
    for _ in range(8):
        torch.cuda.synchronize()
        k0, k1 = ev(), ev()
        if busy:
            k0.record(sb); e.run(10); k1.record(sb)
            sa.wait_event(k0)                     # the probe is queued once the window has been launched
            with torch.cuda.stream(sa):
                torch.cuda._sleep(400_000)        # ~0.2 ms into the window: every SM holds two chain blocks
        a0, a1 = ev(), ev()
        with torch.cuda.stream(sa):
            a0.record(sa); fn(); a1.record(sa)
        torch.cuda.synchronize()
        ts.append((a0.elapsed_time(a1), k0.elapsed_time(k1) if busy else 0.0))
    return ts
for name, fn in (('fill_ 246 MB', lambda: x.fill_(1.0)), ('fill_ 8 MB', lambda: y.fill_(1.0)), ('add_ 8 MB', lambda: y.add_(1.0))):
    alone = timed(fn, False); busy = timed(fn, True)
    print('%-14s alone %.3f ms | beside a running chain window %.3f ms (window %.3f ms)' % (
        name, np.median([t[0] for t in alone]), np.median([t[0] for t in busy]), np.median([t[1] for t in busy])))
