"""Pure-write and pure-read HBM bandwidth on this board (torch fill / sum over 4 GiB), next to the copy figure of MEASURED_PEAKS.json:
the first-layer kernel only writes, the last-layer kernel only reads."""
import torch
n = 1 << 30
x = torch.empty(n, dtype=torch.float32, device="cuda")
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
w = timed(lambda: x.fill_(1.0))
r = timed(lambda: x.sum())
y = torch.empty_like(x)
c = timed(lambda: y.copy_(x))
print({"write_GBs": round(4 * n / w / 1e6, 1), "read_GBs": round(4 * n / r / 1e6, 1), "copy_GBs": round(8 * n / c / 1e6, 1)})
