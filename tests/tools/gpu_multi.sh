#!/bin/bash
# N-GPU run (gpurun --gpus N): torchrun bench (headline + c5 strong-scaled + c4 hypothesis sharding over NCCL), the reference arm
# under torchrun, and the C++ ShardedEngine test spanning the GPUs.   usage: gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_gpus.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/multi_bench.err
echo "rc=$?"; tail -c 600 gpurun_out/multi_bench.err; wc -l gpurun_out/r02_bench_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_n$N.json 2>> gpurun_out/multi_bench.err
echo "rc=$?"; wc -l gpurun_out/r02_bench_reference_n$N.json
python -m pytest tests/test_facade_cpp.py -m gpu -x -q -k sharded 2>&1 | tail -3
