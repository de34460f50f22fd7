"""Measure D2H copy bandwidth: 1-D contiguous vs 2-D with various row widths (pinned host)."""
import torch, time
dev = torch.device("cuda")
total = 1 << 30  # bytes
src = torch.empty(total, dtype=torch.uint8, device=dev)
dst = torch.empty(total, dtype=torch.uint8, pin_memory=True)
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n
print("1D 1GiB: %.1f GB/s" % (total / t(lambda: dst.copy_(src, non_blocking=True)) / 1e9))
for row in (8 << 10, 32 << 10, 128 << 10, 512 << 10, 2 << 20):
    rows = total // (2 * row)
    s2 = src[: rows * row].view(rows, row)
    d2 = dst[: rows * 2 * row].view(rows, 2 * row)[:, :row]  # pitch = 2*row on the host
    print("2D row %7d B x %6d rows: %.1f GB/s" % (row, rows, rows * row / t(lambda: d2.copy_(s2, non_blocking=True)) / 1e9))
