#!/usr/bin/env bash
# Multi-GPU checks in one call:   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_round_multi.sh r02 2'
# (N = 2, 4 or 8).  Bench lines under torchrun, the Python driver under torchrun vs a single process, and the native driver's
# single-process --devices partition vs one device.  Everything lands in gpurun_out/<tag>_*.
set -u
TAG=${1:-rXX}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build_multi.log 2>&1 || { echo "build failed"; exit 1; }
PORT=29531
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 5 --warmup 3 \
    > $OUT/${TAG}_bench_n${N}.json 2> $OUT/${TAG}_bench_n${N}.err; tail -c 400 $OUT/${TAG}_bench_n${N}.json
echo "== Python driver under torchrun (cdens slab gather through the J path, integral all-reduce)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((PORT + 1)) tools/dist_driver_check.py \
    > $OUT/${TAG}_dist_driver_check.txt 2>&1; tail -5 $OUT/${TAG}_dist_driver_check.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((PORT + 2)) tools/dist_integral_check.py \
    > $OUT/${TAG}_dist_integral_check.txt 2>&1; tail -3 $OUT/${TAG}_dist_integral_check.txt
echo "== native driver: --devices all vs one device (reports and files at print precision)"
timeout 900 python tools/native_multi_check.py > $OUT/${TAG}_native_multi_check.txt 2>&1; tail -8 $OUT/${TAG}_native_multi_check.txt
