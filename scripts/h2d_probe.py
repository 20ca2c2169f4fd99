"""PCIe probe for the e2e leg: pinned H2D bandwidth on this box at the copy sizes ndtb_register_scans issues (one scan =
1.6 MB) and in one piece, alone and from 3 streams at once."""
import time
import torch

dev = torch.device("cuda:0")
n, per = 256, 100000 * 16
hs = [torch.empty(per, dtype=torch.uint8).pin_memory() for _ in range(n)]
big_h = torch.empty(n * per, dtype=torch.uint8).pin_memory()
d = torch.empty(n * per, dtype=torch.uint8, device=dev)
ds = [torch.empty(n * per, dtype=torch.uint8, device=dev) for _ in range(3)]


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def small():
    for i, h in enumerate(hs):
        d[i * per:(i + 1) * per].copy_(h, non_blocking=True)


def big():
    d.copy_(big_h, non_blocking=True)


sts = [torch.cuda.Stream(dev) for _ in range(3)]


def three():
    for k, st in enumerate(sts):
        with torch.cuda.stream(st):
            for i, h in enumerate(hs):
                ds[k][i * per:(i + 1) * per].copy_(h, non_blocking=True)


gb = n * per / 1e9
print(f"H2D pinned, {n} x 1.6 MB copies on one stream: {gb / timed(small):.1f} GB/s")
print(f"H2D pinned, one {gb:.2f} GB copy:              {gb / timed(big):.1f} GB/s")
print(f"H2D pinned, 3 streams x {n} x 1.6 MB:          {3 * gb / timed(three):.1f} GB/s")
h_out = torch.empty(n * per, dtype=torch.uint8).pin_memory()
print(f"D2H pinned, one {gb:.2f} GB copy:              {gb / timed(lambda: h_out.copy_(d, non_blocking=True)):.1f} GB/s")
