"""Multi-GPU MLP epoch timing (torchrun, one rank per GPU): configs[2] per rank; mode 0 = overlapped NCCL all-reduces,
   mode 4 = fused peer-memory exchange (szb_comm_peer_exchange).
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/gpu_mlp_multi.py [modes]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import streamz_b200 as sz
from streamz_b200 import _native as N
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("gloo")
ctx = sz.Context(int(os.environ["LOCAL_RANK"]))
uid = [sz.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(uid[0], rank, world)
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
nwin, batch = 1_000_000, 4096
g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
src = torch.randn((nwin, 60), generator=g, device=dev).contiguous()
labels = torch.randint(0, 100, (nwin,), generator=g, device=dev, dtype=torch.int32)
perm = np.random.default_rng(5).permutation(nwin).astype(np.uint32)
loss, used = C.c_double(), C.c_uint64()
modes = [int(m) for m in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 4]
for mode in modes:
    peer = ctx.comm_peer_exchange(bool(mode & 4), {4: "two-shot", 5: "one-shot"}.get(mode, "auto"))   # 4 / 5: fused peer exchange; 0: NCCL
    net = sz.SimpleNeuralNet(60, 512, 256, 100, seed=7, ctx=ctx)
    def epoch(n_rows):
        N.check(N.lib.szb_net_train_epoch_dev(net._h, C.c_void_p(src.data_ptr()), C.c_void_p(labels.data_ptr()), nwin, N.ptr(perm),
                                              n_rows, batch, 0.01, 0.2, 99, 0, None, C.byref(loss), C.byref(used)))
    epoch(batch * 8)
    dist.barrier()
    tr = np.zeros(8, np.float64)
    N.check(N.lib.szb_comm_peer_trace(ctx.handle, 1, N.ptr(tr)))
    ctx.timer_start(); t0 = time.perf_counter()
    epoch(nwin)
    wall = (time.perf_counter() - t0) * 1e3
    ms = ctx.timer_stop()
    N.check(N.lib.szb_comm_peer_trace(ctx.handle, 0, N.ptr(tr)))
    if peer:
        print(f"  rank {rank} exchange phases, mean us per step over {int(tr[7])} steps (CTA 0): scatter {tr[0]/1e3:.2f}, publish1 {tr[1]/1e3:.2f}, "
              f"wait1 {tr[2]/1e3:.2f}, reduce+bcast {tr[3]/1e3:.2f}, publish2 {tr[4]/1e3:.2f}, wait2 {tr[5]/1e3:.2f}  (sum {tr[:6].sum()/1e3:.2f})", flush=True)
    t = torch.tensor([ms, wall], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"world {world} mode {mode} (peer exchange {peer}): {t[0].item():.2f} ms/epoch device, {t[1].item():.2f} ms wall, "
              f"{t[0].item() * 1e3 / 245:.1f} us/step, {world * nwin / t[0].item() / 1e3:.1f} M windows/s, mean loss {loss.value / max(1, used.value):.5f}", flush=True)
dist.destroy_process_group()
