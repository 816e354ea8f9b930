"""Peer-to-peer sanity numbers for a multi-GPU box: topology, cudaMemcpyPeer bandwidth, and a kernel pulling / pushing over
peer memory (torch tensors on two devices in one process; plumbing-level check, not part of the library)."""
import subprocess, time, torch
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
n = torch.cuda.device_count()
print("devices", n, "can_access_peer(0,1)", torch.cuda.can_device_access_peer(0, 1) if n > 1 else None)
if n > 1:
    a = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    b = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:1")
    for name, fn in (("copy 0->1 (push by 0's stream)", lambda: b.copy_(a)), ("copy 1->0", lambda: a.copy_(b))):
        fn(); torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        t0 = time.perf_counter()
        for _ in range(10): fn()
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        dt = (time.perf_counter() - t0) / 10
        print(f"{name}: {a.numel() / dt / 1e9:.1f} GB/s")
    # small transfer latency-ish: 753 KB
    s0 = torch.empty(753040, dtype=torch.uint8, device="cuda:0"); s1 = torch.empty(753040, dtype=torch.uint8, device="cuda:1")
    s1.copy_(s0); torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    t0 = time.perf_counter()
    for _ in range(200): s1.copy_(s0)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    print(f"753 KB peer copy: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us each")
