#!/usr/bin/env bash
# 2-GPU job: peer-image correctness + timings, bench at N=2, plus the quick 1-GPU tests that changed
mkdir -p gpurun_out
echo "== quick tests"
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "concentric or make_rays or decompose_q_inv_device or container or c2_full_size_tensor or peer_image" > gpurun_out/j3_pytest.log 2>&1
tail -15 gpurun_out/j3_pytest.log
echo "== peer image, 2 ranks"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_peer_image_multigpu.py > gpurun_out/j3_peer.log 2>&1
grep -E "world=|PEER_IMAGE|Error|error" gpurun_out/j3_peer.log | head -40
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/j3_bench_n2.json 2> gpurun_out/j3_bench_n2.err
grep -E '"section": "(headline|row_sharded_single_image|e2e)"' gpurun_out/j3_bench_n2.json | cut -c1-2500
tail -n 1 gpurun_out/j3_bench_n2.json | cut -c1-600; tail -5 gpurun_out/j3_bench_n2.err
