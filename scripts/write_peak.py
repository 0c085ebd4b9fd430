"""Development aid: what a pure WRITE stream reaches on this box (torch fill_ of 245.6 MB / 1 GiB), next to the
copy figure of MEASURED_PEAKS.json (read + write bytes).  The prefix build (kernel 1) is a pure write stream."""
import torch
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for nbytes in (245_600_000, 1 << 30, 4 << 30):
    x = torch.empty(nbytes // 8, dtype=torch.float64, device='cuda')
    y = torch.empty_like(x)
    for name, fn, moved in (('fill_', lambda: x.fill_(1.5), nbytes), ('copy_', lambda: y.copy_(x), 2 * nbytes)):
        ts = []
        for _ in range(20):
            ev0.record(); fn(); ev1.record(); torch.cuda.synchronize()
            ts.append(ev0.elapsed_time(ev1))
        print('%-6s %6.1f MB: best %.4f ms -> %.0f GB/s (bytes moved / time)' % (name, nbytes / 1e6, min(ts), moved / min(ts) / 1e6))
