"""Host<->device copy rates of the box with pinned buffers (what bounds the e2e leg of bench.py)."""
import torch
n = 10_000_000
hx = torch.empty(n * 3, dtype=torch.float64).pin_memory()      # 240 MB, the step's inputs
ho = torch.empty(n * 5 // 2, dtype=torch.float64).pin_memory()  # 200 MB, the step's outputs
dx = torch.empty_like(hx, device="cuda")
do = torch.empty_like(ho, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t = timed(lambda: dx.copy_(hx, non_blocking=True))
print("H2D 240 MB: %.3f ms  %.1f GB/s" % (t, 0.24 / t * 1e3))
t = timed(lambda: ho.copy_(do, non_blocking=True))
print("D2H 200 MB: %.3f ms  %.1f GB/s" % (t, 0.20 / t * 1e3))


def both():
    e = torch.cuda.Event()
    e.record()
    with torch.cuda.stream(s1):
        s1.wait_event(e)
        dx.copy_(hx, non_blocking=True)
    with torch.cuda.stream(s2):
        s2.wait_event(e)
        ho.copy_(do, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)


t = timed(both)
print("H2D 240 MB + D2H 200 MB concurrently: %.3f ms  (%.1f GB/s aggregate)" % (t, 0.44 / t * 1e3))
