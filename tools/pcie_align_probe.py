"""D2H copy-engine throughput against the alignment of the two pointers and of the length (pinned destination)."""
import torch

dev = torch.device("cuda", 0)
n = 5_000_000 // 4 * 4
src = torch.zeros(n + 4096, dtype=torch.uint8, device=dev)
dst = torch.zeros(n + 4096, dtype=torch.uint8).pin_memory()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for off_s, off_d, ln in ((0, 0, n), (4, 4, n), (16, 16, n), (64, 64, n), (128, 128, n), (256, 256, n), (4, 4, n - 4), (0, 0, n - 4),
                         (0, 0, n - 60), (4, 0, n), (0, 4, n), (36, 36, n), (2052, 2052, n)):
    best = 1e9
    for _ in range(5):
        e0.record()
        dst[off_d:off_d + ln].copy_(src[off_s:off_s + ln], non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"src+{off_s:5d} dst+{off_d:5d} len {ln}: {best:.3f} ms = {ln / best / 1e6:.1f} GB/s", flush=True)
