#!/usr/bin/env bash
# Strong-scaling bench line only:  /usr/local/graft/bin/gpurun --gpus N --timeout 600 -- 'bash tools/gpu_r02_h.sh r02h N'
set -u
TAG=${1:-r02h}
N=${2:-8}
OUT=gpurun_out
mkdir -p $OUT
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus $N --steps 5 --warmup 3 \
    > $OUT/${TAG}_bench_n${N}.json 2> $OUT/${TAG}_bench_n${N}.err
python - $OUT/${TAG}_bench_n${N}.json <<'PY'
import json,sys
try:
    txt=open(sys.argv[1]).read(); d=json.loads(txt[txt.index('{"metric"'):]); s=d["stage_ms_per_step"]; b=d["balance"]
    print("   ms/step %.2f  pts/s %.3fM  e2e %.3fM  plan %.2f  contract(rank0) %.2f" % (d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, s["ms_plan"], s["ms_contract"]))
    print("   per-rank ms", [round(x,2) for x in b["per_rank_ms"]], " flops max/mean %.4f" % b["flops_max_over_mean"])
    print("   integral_nccl:", d["stages"].get("integral_nccl"))
except Exception as e:
    print("   no result:", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
