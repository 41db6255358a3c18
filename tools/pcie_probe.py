#!/usr/bin/env python
"""Measure the PCIe ceilings the end-to-end path sits under: pinned H2D, D2H, and both at once."""
import torch
n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def both():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)


t = timed(lambda: d_in.copy_(h_in, non_blocking=True)); print(f"H2D alone  {n / t / 1e6:.1f} GB/s")
t = timed(lambda: h_out.copy_(d_out, non_blocking=True)); print(f"D2H alone  {n / t / 1e6:.1f} GB/s")
t = timed(both); print(f"both       {n / t / 1e6:.1f} GB/s each way ({2 * n / t / 1e6:.1f} aggregate)")


def chunked(nseg):
    """The e2e copy pattern without kernels: nseg H2D copies on one stream, nseg D2H copies on another."""
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    c = n // nseg
    with torch.cuda.stream(s1):
        for i in range(nseg):
            d_in[i * c:(i + 1) * c].copy_(h_in[i * c:(i + 1) * c], non_blocking=True)
    with torch.cuda.stream(s2):
        for i in range(nseg):
            h_out[i * c:(i + 1) * c].copy_(d_out[i * c:(i + 1) * c], non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)


for nseg in (1, 16, 64):
    t = timed(lambda: chunked(nseg))
    print(f"both, {nseg:2d} chunks each way: {n / t / 1e6:.1f} GB/s each way")
