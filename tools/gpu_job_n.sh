#!/usr/bin/env bash
# multi-GPU job: N = $1
N=$1
mkdir -p gpurun_out
echo "== peer image, $N ranks"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/run_peer_image_multigpu.py > gpurun_out/jn${N}_peer.log 2>&1
grep -E "world=|PEER_IMAGE|Error|error" gpurun_out/jn${N}_peer.log | head -40
echo "== bench N=$N"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/jn${N}_bench.json 2> gpurun_out/jn${N}_bench.err
grep -E '"section": "(headline|row_sharded_single_image|e2e)"' gpurun_out/jn${N}_bench.json | cut -c1-3000
tail -n 1 gpurun_out/jn${N}_bench.json | cut -c1-300; tail -3 gpurun_out/jn${N}_bench.err
