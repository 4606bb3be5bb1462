#!/usr/bin/env bash
# final build on a 2-GPU box: multi-device / window / shim GPU tests, bench.py under torchrun at 2 ranks, same-box single GPU
set -u
O=gpurun_out; T=r2f
timeout 900 python -m pytest tests -m gpu -q -k "multi_device or window or shim" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_2gpu_$T.json 2> $O/bench_2gpu_$T.err
cut -c1-200 $O/bench_2gpu_$T.json
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_1gpu_samebox_$T.json 2> /dev/null; cut -c1-200 $O/bench_1gpu_samebox_$T.json
python tools/run_configs.py --gpus 2 --only C2 > $O/cfg2_host_$T.jsonl 2> $O/cfg2_host_$T.err; cut -c1-330 $O/cfg2_host_$T.jsonl
