"""Pinned host <-> device copy bandwidth on this box (the ceiling of the host-buffer API, kzgpu_trace): one direction and both at once."""
import torch, time
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return 3 * n / (time.perf_counter() - t0) / 1e9
run(True, True)
print(f"H2D {run(True, False):.1f} GB/s   D2H {run(False, True):.1f} GB/s   both at once {run(True, True):.1f} GB/s per direction")
