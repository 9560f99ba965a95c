#!/bin/bash
# The driver's round-end commands, verbatim: reference arm, then the bench, both with --steps 20 --warmup 5 on one GPU.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
( time python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/dress_ref.json 2> gpurun_out/dress_ref.err; echo "ref rc=$?"; tail -3 gpurun_out/dress_ref.err
( time python3 bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/dress_bench.json 2> gpurun_out/dress_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/dress_bench.err
cut -c1-300 gpurun_out/dress_bench.json
