#!/bin/bash
# 8-GPU A/B on cfg5 (train only): pieces per step with NCCL, NCCL channel limit
O=gpurun_out; mkdir -p $O; T=${1:-r2s}; N=${2:-8}
run() { name=$1; shift; ch=$1; shift
  env "$@" NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload cfg5 --steps 10 --warmup 3 --no-cpu --no-rank --no-sub --chunks $ch > $O/${T}_$name.json 2> $O/${T}_$name.err
  python -c "
import json
d=json.loads(open('$O/${T}_$name.json').read().strip().splitlines()[-1])
print('$name', 'ms/step %.4f warm %.4f e2e %.4f'%(d['ms_per_step'], d['ms_per_step_warm'], d['e2e']['ms_per_step']), {k:round(v,3) for k,v in d['phases_ms_max_over_ranks'].items()})" 2>/dev/null || tail -3 $O/${T}_$name.err
}
run nccl_c1 1 X=1
run nccl_c3 3 X=1
run nccl_c2_ch8 2 NCCL_MAX_NCHANNELS=8
