#!/usr/bin/env bash
# Round 2's multi-GPU pass (run with gpurun --gpus 8): bench.py under torchrun at 2 / 4 / 8 ranks (sample split, and tile split at 8),
# the one-process N-GPU context on the soups with the device-side builder, GPU tests on a multi-GPU box.
set -u
O=gpurun_out; T=${1:-r2}; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > $O/gpus_$T.txt
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 3 --warmup 3 > $O/bench_${n}gpu_$T.json 2> $O/bench_${n}gpu_$T.err
  cut -c1-200 $O/bench_${n}gpu_$T.json
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 3 --warmup 3 --split tiles > $O/bench_8gpu_tiles_$T.json 2> $O/bench_8gpu_tiles_$T.err
cut -c1-200 $O/bench_8gpu_tiles_$T.json
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_1gpu_samebox_$T.json 2> /dev/null; cut -c1-200 $O/bench_1gpu_samebox_$T.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29529 bench.py --gpus 8 --steps 2 --warmup 3 --workload c4 > $O/bench_8gpu_c4_$T.json 2> $O/bench_8gpu_c4_$T.err
cut -c1-200 $O/bench_8gpu_c4_$T.json
DSRT_BUILD_TIMING=1 python tools/run_configs.py --gpus 8 --only C5 --soup-min 8 --soup-max 8 --device-build > $O/cfg8_dev_$T.jsonl 2> $O/cfg8_dev_$T.err
DSRT_BUILD_TIMING=1 python tools/run_configs.py --gpus 8 --only C5 --soup-min 64 --soup-max 64 --device-build >> $O/cfg8_dev_$T.jsonl 2>> $O/cfg8_dev_$T.err; cut -c1-330 $O/cfg8_dev_$T.jsonl
python tools/run_configs.py --gpus 8 --only C2,C5 --soup-min 8 --soup-max 8 > $O/cfg8_host_$T.jsonl 2> $O/cfg8_host_$T.err; cut -c1-330 $O/cfg8_host_$T.jsonl
timeout 900 python -m pytest tests -m gpu -q -k "multi_device or window or shim" 2>&1 | tail -3
