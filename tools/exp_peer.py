"""Micro-benchmark of peer-memory access paths (run under torchrun with 2+ ranks)."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temgymcore_b200 import distributed as D

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", local)
H = W = 1024
pi = D.PeerImage(H, W)
peer = (rank + 1) % world
pview = torch.as_tensor(D._CudaBuf(pi._ptrs[peer], (H, W), "<c16", pi), device=dev)
src = torch.randn((H, W), dtype=torch.complex128, device=dev)
other = torch.randn((H, W), dtype=torch.complex128, device=dev)
loc = torch.empty_like(src)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {
    "memcpy local->local (16.8 MB)": timed(lambda: loc.copy_(src)),
    "memcpy local->own IPC image": timed(lambda: pi.image.copy_(src)),
    "memcpy local->peer image (copy engine/kernel)": timed(lambda: pview.copy_(src)),
    "SM stores: add(out=local)": timed(lambda: torch.add(src, other, out=loc)),
    "SM stores: add(out=own IPC image)": timed(lambda: torch.add(src, other, out=pi.image)),
    "SM stores: add(out=peer image)": timed(lambda: torch.add(src, other, out=pview)),
    "SM loads: add(peer image, local)": timed(lambda: torch.add(pview, other, out=loc)),
    "peer barrier alone": timed(lambda: pi.barrier()),
}
if rank == 0:
    for k, v in res.items():
        print(f"{k:50s} {v*1e3:9.1f} us", flush=True)
    print("can_device_access_peer:", torch.cuda.can_device_access_peer(0, 1), flush=True)
pi.close()
dist.destroy_process_group()
