#!/bin/bash
# default bench on N GPUs (N = $1) through the sharded product path
N=${1:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "rc=$?"; tail -c 300 gpurun_out/bench_${N}gpu.err; cut -c1-200 gpurun_out/bench_${N}gpu.json
