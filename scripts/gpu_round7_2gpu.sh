#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -s -k "two_gpu or sharded" ) > gpurun_out/pytest_2gpu.log 2>&1
grep -n "passed\|failed\|skipped" gpurun_out/pytest_2gpu.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 1500 gpurun_out/bench_2gpu.json
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 ) > gpurun_out/bench_ref_2gpu.json 2> gpurun_out/bench_ref_2gpu.err
tail -c 300 gpurun_out/bench_ref_2gpu.json
