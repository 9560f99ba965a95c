#!/bin/bash
# 8-GPU validation of the sharded product path: default bench (C3 per GPU = configs[4]'s 65 536 samples in total).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
NCCL_DEBUG=WARN timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 2 --warmup 1 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
echo "rc=$?"; tail -c 600 gpurun_out/bench_8gpu.err; cut -c1-400 gpurun_out/bench_8gpu.json
