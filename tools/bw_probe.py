"""HBM bandwidth by direction (development probe): write-only (fill), read-only (reduction), copy, on buffers far larger than
the 126 MB L2.  The 1x1 expansion convs write 4x what they read; their bound is the WRITE figure, not the copy figure."""
import torch

def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

for mb in (134, 537, 1074):
    n = (mb * 1000 * 1000 // 2) // 4096 * 4096
    a = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    b = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    a.normal_()
    t_fill = timed(lambda: a.zero_())
    t_copy = timed(lambda: b.copy_(a))
    t_read = timed(lambda: torch.sum(a.view(torch.int16).view(-1, 4096), dim=1, dtype=torch.int32))
    gb = n * 2 / 1e9
    print(f"{mb:5d} MB  fill {gb / t_fill:6.2f} TB/s ({t_fill * 1e3:6.1f} us)   copy {2 * gb / t_copy:6.2f} TB/s read+write ({t_copy * 1e3:6.1f} us)   "
          f"read {gb / t_read:6.2f} TB/s ({t_read * 1e3:6.1f} us)", flush=True)
