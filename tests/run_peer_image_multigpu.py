"""Multi-GPU check of the fused field sum over peer memory (PeerImage).  Not collected by pytest;
run on a multi-GPU box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/run_peer_image_multigpu.py

Every rank computes its row block; the kernels store it into every rank's image over NVLink; after the
device-side barrier each rank compares its full image with (a) the single-GPU result it computes itself
and (b) the NCCL broadcast + all-gather path.  Prints timings of both exchange styles."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dataclasses import fields, replace  # noqa: E402

from temgymcore_b200 import distributed as D  # noqa: E402
from temgymcore_b200.gaussian import _field_sum_grid, beamlet_polynomials  # noqa: E402
from tests import models as M  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", local)
    ok = True
    for name, (g, model) in {"c2": M.aperture_diffraction_case(10_000, (1024, 1024)),
                             "general": M.biprism_case(2000, (1024, 1024), general=True,
                                                       rng=np.random.default_rng(3))}.items():
        gd = replace(g, **{f.name: torch.as_tensor(getattr(g, f.name), device=dev) for f in fields(g)})
        grid = model[-1]
        H, W = grid.shape
        poly, n, _ = beamlet_polynomials(gd, model)
        for method in ("auto", "sfu"):
            single = _field_sum_grid(poly, n, grid, dev, cull_bits=0, method=method)
            pi = D.PeerImage(H, W)
            try:
                def fused():
                    return D.make_gaussian_image_sharded(gd, model, cull_bits=0, method=method, peer_image=pi)

                def nccl():
                    return D.make_gaussian_image_sharded(gd, model, cull_bits=0, method=method)
                res = {}
                for label, fn in (("fused_peer", fused), ("nccl_gather", nccl)):
                    for _ in range(3):
                        out = fn()
                    torch.cuda.synchronize()
                    dist.barrier()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(10):
                        out = fn()
                    e1.record()
                    torch.cuda.synchronize()
                    t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    res[label] = float(t)
                    err = float((out - single).abs().max() / single.abs().max())
                    # row blocks are tile-aligned: the sharded sums reproduce the single-GPU bits on the SFU
                    # path; the tensor path pre-scales per call (power of two), equal to rounding
                    ok &= err < 1e-6
                    if rank == 0:
                        print(f"{name:8s} {method:5s} {label:12s} world={world} {res[label]:.3f} ms/image "
                              f"max|diff|/max = {err:.2e}", flush=True)
            finally:
                pi.close()
    # the graph-captured step (PeerImagePlan: ray kernel + coefficients + field sum with peer stores + ONE barrier,
    # ping-pong images) on both C2 paths and on C3-like narrow beamlets through the culled SFU kernel
    from temgymcore_b200.gaussian import make_gaussian_image_device
    for name, (g, model), kw in (("c2", M.aperture_diffraction_case(10_000, (1024, 1024)), dict(cull_bits=0, method="auto")),
                                 ("c2", M.aperture_diffraction_case(10_000, (1024, 1024)), dict(cull_bits=0, method="sfu")),
                                 ("c3_20k", M.biprism_case(20_000, (2048, 2048)), dict(method="auto"))):
        gd = replace(g, **{f.name: torch.as_tensor(getattr(g, f.name), device=dev) for f in fields(g)})
        grid = model[-1]
        single = make_gaussian_image_device(gd, model, **kw)
        with D.PeerImage(grid.shape[0], grid.shape[1]) as pi:
            plan = D.PeerImagePlan(gd, model, pi, **kw)
            outs = [plan.run().clone() for _ in range(5)]          # odd count: both buffers, ends on buffer 0
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                plan.run()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            errs = [float((o - single).norm() / single.norm()) for o in outs]
            ok &= max(errs) < 1e-6 and pi.status()[1] == 0
            if rank == 0:
                print(f"{name:8s} {kw['method']:5s} graph_plan   world={world} {float(t):.3f} ms/image "
                      f"rel-L2 vs single GPU = {max(errs):.2e}", flush=True)
    # complex64 images and the explicit tensor method: the f64 -> c64 conversion kernel and the c64 branch of the
    # split-reduce kernel issue the peer stores there
    g, model = M.aperture_diffraction_case(3000, (512, 768))
    gd = replace(g, **{f.name: torch.as_tensor(getattr(g, f.name), device=dev) for f in fields(g)})
    grid = model[-1]
    poly, n, _ = beamlet_polynomials(gd, model)
    for method in ("tensor", "sfu"):
        single = _field_sum_grid(poly, n, grid, dev, cull_bits=0, method=method, out_dtype=torch.complex64)
        pi = D.PeerImage(grid.shape[0], grid.shape[1], dtype=torch.complex64)
        try:
            out = D.make_gaussian_image_sharded(gd, model, cull_bits=0, method=method, peer_image=pi)
            torch.cuda.synchronize()
            err = float((out - single).abs().max() / single.abs().max())
            ok &= out.dtype == torch.complex64 and err < 1e-6
            if rank == 0:
                print(f"c64      {method:6s} fused_peer   world={world} max|diff|/max = {err:.2e}", flush=True)
        finally:
            pi.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("PEER_IMAGE_OK" if int(flag) else "PEER_IMAGE_FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
