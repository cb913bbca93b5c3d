#!/bin/bash
# 2-GPU run: torchrun bench (headline + c5 sharded + c4 hypothesis sharding over NCCL), the reference arm under torchrun,
# and the C++ ShardedEngine test spanning both GPUs
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/s13_gpus.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/s13_bench.err
echo "rc=$?"; tail -c 1500 gpurun_out/s13_bench.err; wc -l gpurun_out/r02_bench_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_n2.json 2>> gpurun_out/s13_bench.err
echo "rc=$?"; wc -l gpurun_out/r02_bench_reference_n2.json
python -m pytest tests/test_facade_cpp.py -m gpu -x -q -k sharded 2>&1 | tail -5
