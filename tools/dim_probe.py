#!/usr/bin/env python
"""One rank of the dimension-sharded step on ONE GPU (no NCCL): the local kernels of rank 0 of a W-rank job on a
dataset shape -- corruption generator + sort, partial sums, backward, segmented reduction + optimizer -- timed
per phase with CUDA events, L2 flushed before every step.  The all-reduce is replaced by a scale of the rank's own
sums (same bytes read/written).  Sizes the per-GPU time of the W-GPU job before spending multi-GPU minutes.

    python tools/dim_probe.py --workload cfg5 --world 8 --steps 20
"""
import argparse
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from emgraph_b200 import _lib  # noqa: E402
from emgraph_b200 import distributed as D  # noqa: E402
from emgraph_b200.engine import get_engine, internal_k, model_id  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg5")
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--chunks", type=int, default=1)
    ap.add_argument("--pipeline", type=int, default=0)
    ap.add_argument("--uniform", action="store_true")
    ap.add_argument("--sorted", type=int, default=0, help="1: phase 1 over the sorted slot list")
    args = ap.parse_args()
    w = bench.WORKLOADS[args.workload]
    W = args.world
    eng = get_engine(0)
    dev = eng.tdev
    E, R, k, eta = w["E"], w["R"], w["k"], w["eta"]
    B = int(math.ceil(w["N"] / w["batches"]))
    n = B * W
    kc = D.dim_width(k, W)
    Kc = internal_k(w["model"], kc)
    need = (args.steps + 4) * n
    X = bench.synth_triples(E, R, min(w["N"], need), seed=0, zipf=not args.uniform)
    Xd = torch.from_numpy(X).to(dev)
    nb = max(1, X.shape[0] // n)
    g = torch.Generator(device=dev).manual_seed(1)
    lim = math.sqrt(6.0 / (E + internal_k(w["model"], k)))
    ent = torch.empty((E, Kc), device=dev).uniform_(-lim, lim, generator=g)
    rel = torch.empty((R, Kc), device=dev).uniform_(-0.1, 0.1, generator=g)
    st = dict(ent_m=torch.zeros_like(ent), ent_v=torch.zeros_like(ent), rel_m=torch.zeros_like(rel), rel_v=torch.zeros_like(rel))
    loss = torch.zeros(1, device=dev)
    bounds = D.chunk_bounds(n, args.chunks)
    sums_flat = torch.zeros((1 + eta) * n, device=dev)
    sums = [sums_flat[(1 + eta) * lo:(1 + eta) * hi] for lo, hi in bounds]
    flush = torch.empty(bench.L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev)
    kw = dict(model=model_id(w["model"]), loss=_lib.LOSS_IDS[w["loss"]], opt=0, k=kc, k_model=k, eta=eta, margin=w["margin"], lr=w["lr"], seed=0)
    names = ["emit+partial", "allreduce_standin", "backward", "reduce_apply"]
    acc = {nm: 0.0 for nm in names}
    total = 0.0
    for s in range(args.steps + 3):
        b = s % nb
        pos = Xd[b * n:(b + 1) * n]
        a = eng.train_args(ent=ent, rel=rel, pos=pos, loss_out=loss, step=s + 1, flags=(_lib.F_PIPELINE if args.pipeline else 0), **kw, **st)
        flush.fill_(float(s))
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        if args.sorted:
            eng.train_partial_sorted(a, sums_flat, len(bounds))
        else:
            for c, (lo, hi) in enumerate(bounds):
                eng.train_partial(a, sums[c], lo, hi)
        ev[1].record()
        for t in sums:
            t.mul_(float(W))  # stand-in for the sum over ranks (every rank would hold comparable partial sums)
        ev[2].record()
        for c, (lo, hi) in enumerate(bounds):
            eng.train_backward(a, sums[c], lo, hi)
        ev[3].record()
        eng.train_reduce(a)
        ev[4].record()
        torch.cuda.synchronize()
        if s >= 3:
            for i, nm in enumerate(names):
                acc[nm] += ev[i].elapsed_time(ev[i + 1])
            total += ev[0].elapsed_time(ev[4])
    steps = args.steps
    out = {"workload": args.workload, "world": W, "global_batch": n, "Kc": Kc, "chunks": args.chunks, "pipeline": args.pipeline, "sorted": args.sorted,
           "ms_per_step": total / steps, "phases_ms": {nm: v / steps for nm, v in acc.items()}, "loss": float(loss.item()),
           "triples_per_s_job": n * (1 + eta) / (total / steps * 1e-3)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
