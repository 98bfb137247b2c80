"""Measure pinned H2D bandwidth on this box: one 39 MB copy vs 16 keyframe-sized copies (the
end-to-end path's per-step traffic)."""
import torch, time
dev = torch.device("cuda:0")
def bw(chunks, reps=20):
    hs = [torch.empty(n, dtype=torch.uint8).pin_memory() for n in chunks]
    ds = [torch.empty(n, dtype=torch.uint8, device=dev) for n in chunks]
    st = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        for h, d in zip(hs, ds): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(reps):
            for h, d in zip(hs, ds): d.copy_(h, non_blocking=True)
        e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return sum(chunks) / ms / 1e6, ms
P = 640 * 480
print("one 39MB copy  : %.1f GB/s (%.3f ms)" % bw([P * 16 * 8]))
print("16 frame copies: %.1f GB/s (%.3f ms)" % bw([P * 12, P * 4] * 8))
print("8 frame copies : %.1f GB/s (%.3f ms)" % bw([P * 16] * 8))
print("1 GB copy      : %.1f GB/s (%.3f ms)" % bw([1 << 30], reps=3))
