"""HBM bandwidth probes next to MEASURED_PEAKS.json's copy figure: pure write (fill), pure read (sum), copy."""
import torch

dev = torch.device("cuda:0")
n = 1 << 30                      # 4 GiB of fp32
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)


def t(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best


w = t(lambda: a.fill_(1.0))
r = t(lambda: a.sum())
c = t(lambda: b.copy_(a))
m = t(lambda: torch.cuda.memset if False else a.zero_())
gb = n * 4 / 1e9
print(f"HBM_PROBE write(fill) {gb / w * 1e3:.0f} GB/s | write(zero_) {gb / m * 1e3:.0f} GB/s | read(sum) {gb / r * 1e3:.0f} GB/s | copy {2 * gb / c * 1e3:.0f} GB/s (read+write)")
