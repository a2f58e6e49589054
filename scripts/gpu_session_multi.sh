#!/bin/bash
# Multi-GPU session on one box: scripts/gpu_session_multi.sh N [tag]  -> gpurun_out/*_<tag>_nN.json
N=${1:-2}
TAG=${2:-r01}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29501 bench.py --gpus $N --steps 200 --warmup 5 2>&1 | tail -1 > $O/bench_${TAG}_n$N.json
timeout 300 python scripts/config5_bench.py --gpus $N --graphs 64 --no-parity 2>&1 | tail -1 > $O/config5_${TAG}_n$N.json
timeout 200 $TR --master-port 29511 scripts/h2n_strips.py --steps 50 2>&1 | tail -1 > $O/h2n_strips_${TAG}_n$N.json
timeout 200 $TR --master-port 29513 scripts/h2n_strips.py --steps 50 --halo nccl 2>&1 | tail -1 > $O/h2n_strips_nccl_${TAG}_n$N.json
timeout 200 $TR --master-port 29512 scripts/resize_strips.py --steps 50 2>&1 | tail -2 > $O/resize_strips_${TAG}_n$N.json
for f in bench config5 h2n_strips h2n_strips_nccl resize_strips; do echo "== $f"; cut -c1-260 $O/${f}_${TAG}_n$N.json; done
