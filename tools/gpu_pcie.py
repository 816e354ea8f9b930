"""PCIe ceiling for the e2e figure: pinned H2D of the C2 input (3.2 GB) and D2H of its features (2.64 GB), alone and together."""
import torch, time
dev = torch.device("cuda", 0)
n_in, n_out = 3_200_000_000, 2_642_400_000
h_in = torch.empty(n_in, dtype=torch.uint8, pin_memory=True); d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
h_out = torch.empty(n_out, dtype=torch.uint8, pin_memory=True); d_out = torch.empty(n_out, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if h2d:
        with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    if d2h:
        with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter() - t0
for _ in range(2): run(True, True)
a = min(run(True, False) for _ in range(3)); b = min(run(False, True) for _ in range(3)); c = min(run(True, True) for _ in range(3))
print(f"H2D alone {a*1e3:.1f} ms ({n_in/a/1e9:.1f} GB/s); D2H alone {b*1e3:.1f} ms ({n_out/b/1e9:.1f} GB/s); both {c*1e3:.1f} ms -> ceiling {100000/c/1e6:.2f} M audio-s/s")
