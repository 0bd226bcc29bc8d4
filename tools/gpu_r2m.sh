#!/bin/bash
# multi-GPU session: sharded parity tests of this world size + the bench line; usage: gpu_r2m.sh TAG N "pytest -k expr"
O=gpurun_out; mkdir -p $O; T=${1:-r2m}; N=${2:-2}; K="${3:-}"
nvidia-smi -L > $O/${T}_gpus.txt
if [ -n "$K" ]; then
  ( timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "$K" 2>&1 | tail -40 ) > $O/${T}_pytest_multi.log
else
  ( timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -40 ) > $O/${T}_pytest_multi.log
fi
tail -5 $O/${T}_pytest_multi.log
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > $O/${T}_bench_n$N.json 2> $O/${T}_bench_n$N.err
tail -c 2500 $O/${T}_bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("$O/${T}_bench_n$N.json").read().strip().splitlines()[-1])
    print("cfg3 N=$N ms/step %.4f warm %.4f value %.3e e2e %.4f"%(d["ms_per_step"], d["ms_per_step_warm"], d["value"], d["e2e"]["ms_per_step"]), {k:round(v,4) for k,v in d["phases_ms_max_over_ranks"].items()}, d["rank_parity"], d["first_step_loss"], d["first_step_loss_single_gpu"])
    print("rank", {k:d["rank"][k] for k in ("value","ms_per_step","mrr")})
    c=d.get("cfg5",{})
    print("cfg5", {k:c.get(k) for k in ("value","ms_per_step","ms_per_step_warm","rank_parity","first_step_loss","first_step_loss_single_gpu","error")}, c.get("e2e"))
    print("cfg5 phases", {k:round(v,4) for k,v in (c.get("phases_ms_max_over_ranks") or {}).items()})
except Exception as e:
    print("ERR", e)
PY
