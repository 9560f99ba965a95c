#!/bin/bash
# 2-GPU pass: the sharded product path equals the single-GPU result bit for bit; default bench on 2 GPUs; C5 strong scaling.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_noise.py -m gpu -q -rP -k "two_rank or sharding" > gpurun_out/pytest_2gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_2gpu.log; tail -3 gpurun_out/pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 400 gpurun_out/bench_2gpu.err; cut -c1-300 gpurun_out/bench_2gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload C5 --steps 1 --warmup 1 > gpurun_out/bench_c5_2gpu.json 2> gpurun_out/bench_c5_2gpu.err
tail -c 400 gpurun_out/bench_c5_2gpu.err; cut -c1-300 gpurun_out/bench_c5_2gpu.json
