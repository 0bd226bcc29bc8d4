"""N-process sweep of the owner-push knobs on the cfg5 shape: per-phase step times for each setting.
torchrun --nproc-per-node N tools/push_sweep.py"""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from emgraph_b200.distributed import ShardedKGE

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
E, R, k, eta, B = 4594485, 822, 256, 64, 10308
sk = ShardedKGE("DistMult", k, eta, "nll", "adam", E, R, B, lr=5e-4, seed=0, device=local)
dev = sk.eng.tdev
g = torch.Generator(device=dev).manual_seed(100 + rank)
sk.ent.tensor.uniform_(-0.03, 0.03, generator=g); sk.rel.uniform_(-0.05, 0.05)
pos = torch.stack([torch.randint(0, E, (B * 8,), device=dev, generator=g), torch.randint(0, R, (B * 8,), device=dev, generator=g),
                   torch.randint(0, E, (B * 8,), device=dev, generator=g)], 1).to(torch.int32).contiguous()
flush = torch.empty(128 << 20, dtype=torch.float32, device=dev)
settings = [("8", "0"), ("4", "0"), ("2", "0"), ("1", "0"), ("8", "1"), ("8", "64"), ("8", "1024"), ("4", "1024"), ("2", "1024"), ("2", "1")]
for ctas, group in settings:
    os.environ["KGE_PUSH_CTAS"], os.environ["KGE_PUSH_GROUP"] = ctas, group
    for i in range(2):
        sk.train_step(pos[(i % 8) * B:(i % 8 + 1) * B])
    torch.cuda.synchronize(); dist.barrier()
    sk.timing = True
    for i in range(6):
        flush.fill_(float(i))
        sk.train_step(pos[(i % 8) * B:(i % 8 + 1) * B])
    ph = sk.phase_times(); sk.timing = False
    names = sorted(ph)
    t = torch.tensor([ph[n] for n in names], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        d = dict(zip(names, [round(v, 3) for v in t.tolist()]))
        print("ctas/sm %s group %s: push %.3f barrier %.3f total(sum of max) %.3f  %s" % (ctas, group, d["push"], d["push_barrier"], sum(d.values()), d), flush=True)
dist.destroy_process_group()
