"""D2H bandwidth into pinned host memory: one cudaMemcpyAsync at a time against 2 / 4 concurrent copies on separate
streams, and against a kernel storing straight into mapped pinned memory (what the rows-to-host path could do instead)."""
import time

import torch

dev = torch.device("cuda")
total = 2 << 30
src = torch.empty(total, dtype=torch.uint8, device=dev)
dst = torch.empty(total, dtype=torch.uint8, pin_memory=True)


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


for k in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(k)]
    part = total // k

    def go():
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                dst[i * part:(i + 1) * part].copy_(src[i * part:(i + 1) * part], non_blocking=True)

    print(f"{k} concurrent copies of {part >> 20} MiB: {total / timed(go) / 1e9:.1f} GB/s")
for chunk in (64 << 20, 256 << 20):
    n = total // chunk

    def go2():
        for i in range(n):
            dst[i * chunk:(i + 1) * chunk].copy_(src[i * chunk:(i + 1) * chunk], non_blocking=True)

    print(f"back-to-back copies of {chunk >> 20} MiB on one stream: {total / timed(go2) / 1e9:.1f} GB/s")
