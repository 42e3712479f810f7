"""Shared timing helper for the GPU probes."""
import torch

flush = None


def bench(name, fn, flops=0, bytes_=0, n=10):
    """Two timings per kernel: `cold` = one launch after an L2 flush (CUDA events; includes any CPU launch gap when
    the host has not run ahead), `graph` = 10 launches captured in a CUDA graph and replayed (no host in the loop,
    operands L2-warm as in the sampler loop)."""
    global flush
    if flush is None:
        flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    for _ in range(2): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.sum()   # evict L2 with clean lines
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / n
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            for _ in range(10): fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(3): g.replay()
    e1.record(); torch.cuda.synchronize()
    mg = e0.elapsed_time(e1) / 30
    s = f"{name:44s} cold {ms*1e3:7.1f} us  graph {mg*1e3:7.1f} us"
    if flops: s += f"  {flops/mg/1e9:7.1f} TFLOP/s"
    if bytes_: s += f"  {bytes_/mg/1e6:7.1f} GB/s"
    print(s, flush=True)

