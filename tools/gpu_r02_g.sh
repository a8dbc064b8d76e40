#!/usr/bin/env bash
# Round-2 multi-GPU call:  /usr/local/graft/bin/gpurun --gpus N --timeout 900 -- 'bash tools/gpu_r02_g.sh r02g N'
# The headline bench under torchrun (whole 256^3 grid, cost-balanced partition, strong scaling; stages.integral_nccl = the NCCL-reduced
# plane integral), the NCCL integral test, the Python driver under torchrun and the native driver's --devices mode.
set -u
TAG=${1:-r02g}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
PORT=29631
echo "== pytest -m gpu: kernel parity (tiled scans, 30-bit sort)"
( timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu ) > $OUT/${TAG}_pytest_parity.log 2>&1; tail -2 $OUT/${TAG}_pytest_parity.log
echo "== bench N=$N (strong scaling through gimic_b200_partition_*)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 5 --warmup 3 \
    > $OUT/${TAG}_bench_n${N}.json 2> $OUT/${TAG}_bench_n${N}.err
python - $OUT/${TAG}_bench_n${N}.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); s=d["stage_ms_per_step"]; b=d["balance"]
    print("   ms/step %.2f  pts/s %.3fM  e2e %.3fM  plan %.2f  contract(rank0) %.2f" % (d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, s["ms_plan"], s["ms_contract"]))
    print("   per-rank ms", [round(x,2) for x in b["per_rank_ms"]], " flops max/mean %.4f" % b["flops_max_over_mean"], " points", b["per_rank_points"])
    print("   integral_nccl:", d["stages"].get("integral_nccl"))
except Exception as e:
    print("   no result:", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
echo "== NCCL integral test + drivers on $N GPUs"
( timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "nccl_integral" ) > $OUT/${TAG}_pytest_nccl.log 2>&1; tail -2 $OUT/${TAG}_pytest_nccl.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((PORT + 1)) tools/dist_driver_check.py \
    > $OUT/${TAG}_dist_driver_check.txt 2>&1; tail -3 $OUT/${TAG}_dist_driver_check.txt
timeout 400 python tools/native_multi_check.py > $OUT/${TAG}_native_multi_check.txt 2>&1; tail -6 $OUT/${TAG}_native_multi_check.txt
