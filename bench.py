#!/usr/bin/env python
"""Benchmark of the KGE hot path: train step (score + eta-negative corruption + loss + backward +
sparse optimizer) and filtered ranking, on synthetic triples of the BASELINE.json dataset shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--impl b200|reference]

Prints ONE JSON line (rank 0).  Contract (see DESIGN.md "Measurement"):
  value      train triples/sec (eta negatives incl.), inputs resident in HBM, L2 flushed between the
             timed steps, CUDA-event timed on the launching stream, max over ranks
  e2e        the same metric through the public API with HOST buffers: every step copies its batch
             from pinned host memory and reads the batch loss back
  roofline   dominant training kernel: algorithmic bytes / live CUDA-event kernel time vs the measured
             HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the reference-equivalent CPU op graph (oracle/torch_port.py) on the host cores, on a
             bounded sample
  rank       the second half of BASELINE.json's metric: filtered-rank test triples/sec (same keys)
  cfg5       the same train measurement on the Wikidata5M-shaped config (the one BASELINE.json asks to
             scale over 1/2/4/8 GPUs), carried by EVERY line so the driver's 1->8 sweep records it
  others     one-line summaries of the remaining BASELINE configs (N = 1 only)
  rank_parity  ranks of the seeded, untrained model: identical for every N (a correctness witness of
             the multi-GPU path inside the scaling record itself)
`--impl reference` times only the CPU arm (all host threads) on the same config.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# SURVEY.md section 8 config table (BASELINE.json configs[0..4] == cfg1..cfg5)
WORKLOADS = {
    "cfg1": dict(desc="TransE k=100 eta=20 pairwise adam, WN18-shaped", model="TransE", k=100, eta=20, loss="pairwise",
                 margin=1.0, opt="adam", lr=1e-4, E=40943, R=18, N=141442, T=5000, batches=64),
    "cfg2": dict(desc="DistMult k=200 eta=10 pairwise margin=5, FB15k-237-shaped", model="DistMult", k=200, eta=10,
                 loss="pairwise", margin=5.0, opt="adam", lr=5e-4, E=14541, R=237, N=272115, T=20466, batches=64),
    "cfg3": dict(desc="ComplEx k=200 eta=20 nll adam, FB15k-237-shaped, 20k test triples", model="ComplEx", k=200, eta=20,
                 loss="nll", margin=1.0, opt="adam", lr=5e-4, E=14541, R=237, N=272115, T=20000, batches=64),
    "cfg4": dict(desc="HolE k=256 eta=20 multiclass_nll, YAGO3-10-shaped", model="HolE", k=256, eta=20, loss="multiclass_nll",
                 margin=1.0, opt="adam", lr=5e-4, E=123182, R=37, N=1079040, T=5000, batches=100),
    "cfg5": dict(desc="DistMult k=256 eta=64 nll adam, Wikidata5M-shaped", model="DistMult", k=256, eta=64, loss="nll",
                 margin=1.0, opt="adam", lr=5e-4, E=4594485, R=822, N=20614279, T=5133, batches=2000),
}
TRAIN_METRIC = "train triples/sec (eta negatives incl.)"
RANK_METRIC = "filtered-rank test triples/sec"
L2_FLUSH_BYTES = 512 << 20
TIMED_REGION_MS = 100.0  # every timed region is at least this long: steps are repeated (config.inner_repeat)
TABLE_SEED = 2


def internal_k(model, k):
    return 2 * k if model in ("ComplEx", "HolE") else k


# ------------------------------------------------------------------------------------------------
# synthetic graph of a dataset shape: Zipf(1.0) subjects/objects truncated to E and
# randomly permuted, uniform relations, no self loops; the first E triples are a covering chain so
# every entity occurs.  Duplicates are left in (the filter de-duplicates; training does not care).
# ------------------------------------------------------------------------------------------------
def synth_triples(E, R, n, seed, zipf=True):
    rng = np.random.Generator(np.random.PCG64(seed))
    if zipf:
        cdf = np.cumsum(1.0 / np.arange(1, E + 1))
        cdf /= cdf[-1]
        perm = rng.permutation(E).astype(np.int32)

        def draw(m):
            return perm[np.minimum(np.searchsorted(cdf, rng.random(m)), E - 1)]
    else:
        def draw(m):
            return rng.integers(0, E, size=m, dtype=np.int32)
    out = np.empty((n, 3), np.int32)
    c = min(E, n)
    out[:c, 0] = np.arange(c)
    out[:c, 2] = (np.arange(c) + 1) % E
    if n > c:
        s, o = draw(n - c), draw(n - c)
        clash = s == o
        o[clash] = (o[clash] + 1) % E
        out[c:, 0], out[c:, 2] = s, o
    out[:, 1] = rng.integers(0, R, size=n, dtype=np.int32)
    # interleave the chain with the random part so that batches look alike
    return out[rng.permutation(n)]


def glorot(rows, cols, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    lim = math.sqrt(6.0 / (rows + cols))
    return rng.uniform(-lim, lim, size=(rows, cols)).astype(np.float32)


def seeded_table(rows, model, k, dev, seed, c0=0, c1=None, pad_to=None):
    """Glorot-uniform table drawn on the device in blocks of 4 columns, each block from its own generator seed, so that
    ANY column range of the table can be produced without the rest: rank r of a column-sharded run draws exactly the
    values the single-GPU run holds in those columns (rank_parity below depends on it).  Returns columns [c0,c1) of each
    half, zero-padded to pad_to columns per half: [rows, halves*width]."""
    import torch
    halves = 2 if model in ("ComplEx", "HolE") else 1
    c1 = k if c1 is None else c1
    width = (c1 - c0) if pad_to is None else pad_to
    K = internal_k(model, k)
    lim = math.sqrt(6.0 / (rows + K))
    out = torch.zeros((rows, halves * width), dtype=torch.float32, device=dev)
    assert c0 % 4 == 0
    for h in range(halves):
        for b in range(c0 // 4, (c1 + 3) // 4):
            g = torch.Generator(device=dev).manual_seed(seed * 1000003 + h * 50021 + b)
            blk = torch.empty((rows, 4), dtype=torch.float32, device=dev).uniform_(-lim, lim, generator=g)
            lo, hi = 4 * b, min(4 * b + 4, c1)
            out[:, h * width + lo - c0:h * width + hi - c0] = blk[:, :hi - lo]
    return out


def batch_positives(w):
    return int(math.ceil(w["N"] / w["batches"]))


def make_dataset(w, need_train, with_test, zipf=True):
    """(X, test): synthetic triples of the shape + T test triples sampled from them.  The size does NOT depend on the number
    of GPUs or steps (the whole set up to 4 M triples, batches wrap around), so that the test set, the filter and with them
    the rank_parity digest are the same at every N."""
    n = min(w["N"], 4_000_000) if with_test else min(w["N"], need_train)
    X = synth_triples(w["E"], w["R"], n, seed=0, zipf=zipf)
    test = None
    if with_test:
        rng = np.random.Generator(np.random.PCG64(1))
        test = X[rng.permutation(X.shape[0])[: w["T"]]].copy()
    return X, test


def config_for(args, name, w, world):
    """The SAME dictionary in both arms (GPU and --impl reference) and at every N except for `parallelism`."""
    K = internal_k(w["model"], w["k"])
    return {"workload": "%s: %s" % (name, w["desc"]), "batch_positives_per_gpu": batch_positives(w), "global_batch": batch_positives(w) * world,
            "eta": w["eta"], "E": w["E"], "R": w["R"], "K": K, "test_triples": w["T"],
            "entity_popularity": "uniform" if args.uniform else "zipf(1.0)", "optimizer": w["opt"],
            "l2": "GPU arm: flushed before every timed step (%d MiB write); CPU arm: n/a" % (L2_FLUSH_BYTES >> 20),
            "inner_repeat": "every one of the --steps is the mean of R back-to-back-submitted, individually timed steps, R chosen so "
                            "that the timed region is >= %d ms (R is in measurement.inner_repeat)" % int(TIMED_REGION_MS),
            "parallelism": "1 GPU" if world == 1 else
                           "%d GPUs: tables + optimizer state column-sharded (dimension-parallel), global batch on every GPU, one "
                           "all-reduce of (1+eta) partial scores per positive" % world}


# ------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x1: "gpu_idle"}

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join()
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d["bf16_tflops"]), bf16_sus=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src="fallback (B200_PROFILING.md)")


def load_traffic(workload, kernel):
    """DRAM bytes per launch of `kernel` from an `ncu --set full` capture of this command: the capture of the current
    gpurun session when tools/ncu_traffic.py left one in gpurun_out/ (its session tag is returned too), else the
    committed capture of the round (profiles/traffic_r02.json, then r01).  (bytes, source) or (None, None)."""
    for p in (os.path.join(ROOT, "gpurun_out", "traffic_session.json"), os.path.join(ROOT, "profiles", "traffic_r02.json"),
              os.path.join(ROOT, "profiles", "traffic_r01.json")):
        try:
            d = json.load(open(p))
            return float(d[workload][kernel]), "%s (session %s)" % (os.path.relpath(p, ROOT), d.get("session", "r01"))
        except Exception:
            continue
    return None, None


def measure_tf32_peak(dev):
    """Dense TF32 tensor throughput of THIS GPU, measured the way MEASURED_PEAKS.json measures bf16: cuBLAS fp32 GEMM
    with TF32 tensor cores, 8192^3, best of 10 after warm-up (TFLOP/s).  The ranking roofline divides by it."""
    import torch
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn((n, n), device=dev)
        b = torch.randn((n, n), device=dev)
        for _ in range(3):
            a @ b
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference-equivalent op graph (oracle/torch_port.py) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_arm(w, X, test, filt_for_rank, steps, warmup, budget_s, rank_budget_s, do_rank=True):
    import torch
    from oracle import torch_port as tp
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    K = internal_k(w["model"], w["k"])
    B = batch_positives(w)
    ent, rel = glorot(w["E"], K, 2), glorot(w["R"], K, 3)
    tr = tp.CpuTrainer(w["model"], w["k"], w["loss"], w["eta"], ent, rel, margin=w["margin"], lr=w["lr"], optimizer=w["opt"])
    g = torch.Generator().manual_seed(0)
    Xt = torch.as_tensor(X.astype(np.int64))
    nb = max(1, Xt.shape[0] // B)
    t_begin = time.perf_counter()
    for i in range(warmup):
        tr.step(Xt[(i % nb) * B:(i % nb + 1) * B], rng=g)
        if time.perf_counter() - t_begin > budget_s * 0.3:
            break
    done, t0 = 0, time.perf_counter()
    while done < steps and (time.perf_counter() - t0) < budget_s:
        b = (warmup + done) % nb
        tr.step(Xt[b * B:(b + 1) * B], rng=g)
        done += 1
    dt = time.perf_counter() - t0
    train = dict(value=done * B * (1 + w["eta"]) / dt, ms_per_step=1e3 * dt / max(done, 1), steps_done=done,
                 sample="%d of %d requested steps of %d positives x (1+%d), dense Keras-style fresh-state %s, torch-CPU %d threads"
                 % (done, steps, B, w["eta"], w["opt"], cores))
    rank = None
    if do_rank:
        rk = tp.CpuRanker(w["model"], w["k"], ent, rel, filt_for_rank)
        rk.rank(test[0])
        n, t0 = 0, time.perf_counter()
        while n < test.shape[0] and (time.perf_counter() - t0) < rank_budget_s:
            rk.rank(test[n])
            n += 1
        dt = time.perf_counter() - t0
        rank = dict(value=n / dt, sample="the first %d of the GPU arm's %d test triples, per-triple 2E-corruption sweep, dict filter, torch-CPU %d threads"
                    % (n, test.shape[0], cores), n=n)
    return train, rank, cores


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-rank", action="store_true", help="skip the ranking half")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sub", action="store_true", help="skip the cfg5 sub-record and the other configs' summaries")
    ap.add_argument("--rank-steps", type=int, default=3)
    ap.add_argument("--rank-tc", type=int, default=-1, help="1/0 force the tensor-core ranking sweep on/off")
    ap.add_argument("--uniform", action="store_true", help="uniform instead of Zipf entity popularity")
    ap.add_argument("--chunks", type=int, default=2, help="multi-GPU: pieces a step is cut into (all-reduce / kernel overlap)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    w = dict(WORKLOADS[args.workload])
    rank_id = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if rank_id != 0:
            return 0
        return reference_main(args, w)
    if world > 1:
        return multi_gpu_main(args, w, rank_id, world)
    if args.gpus > 1:
        raise SystemExit("--gpus %d needs one process per GPU: launch with `python -m torch.distributed.run --nnodes=1 "
                         "--nproc-per-node %d --master-addr 127.0.0.1 --master-port P bench.py --gpus %d ...`" % (args.gpus, args.gpus, args.gpus))
    return single_gpu_main(args, w)


def reference_main(args, w):
    B = batch_positives(w)
    X, test = make_dataset(w, (args.steps + args.warmup + 1) * B, not args.no_rank, zipf=not args.uniform)
    train, rank, cores = cpu_arm(w, X, test if test is not None else X[:8], X, args.steps, args.warmup, budget_s=150.0, rank_budget_s=20.0,
                                 do_rank=not args.no_rank)
    line = {
        "impl": "reference", "metric": TRAIN_METRIC, "value": train["value"], "unit": "triples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": train["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_for(args, args.workload, w, max(1, args.gpus)),
        "cpu_baseline": {"value": train["value"], "unit": "triples/s", "cores": cores, "kind": "port", "sample": train["sample"]},
        "e2e": {"value": train["value"], "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if rank is not None:
        line["rank"] = {"metric": RANK_METRIC, "value": rank["value"], "unit": "test triples/s",
                        "e2e": {"value": rank["value"], "unit": "test triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                        "cpu_baseline": {"value": rank["value"], "unit": "test triples/s", "cores": cores, "kind": "port", "sample": rank["sample"]}}
    print(json.dumps(line), flush=True)
    return 0


def inner_repeat(steps, est_ms):
    return max(1, int(math.ceil(TIMED_REGION_MS / max(steps * est_ms, 1e-6))))


def build_model(w, dev):
    """The public-API model of a workload with the seeded tables injected (initializer='constant')."""
    from emgraph_b200 import models
    E, R, k = w["E"], w["R"], w["k"]
    ent0 = seeded_table(E, w["model"], k, dev, TABLE_SEED).cpu().numpy()
    rel0 = seeded_table(R, w["model"], k, dev, TABLE_SEED + 1).cpu().numpy()
    cls = models.MODEL_REGISTRY[w["model"]]
    model = cls(k=k, eta=w["eta"], epochs=1, batches_count=w["batches"], seed=0, optimizer=w["opt"], optimizer_params={"lr": w["lr"]},
                loss=w["loss"], loss_params={"margin": w["margin"]}, initializer="constant",
                initializer_params={"entity": ent0, "relation": rel0})
    return model


def rank_pretrain_single(args, w, eng, ent, rel, X, test):
    """Filtered 's,o' ranks of the seeded, untrained tables on one GPU (the witness multi-GPU runs must reproduce)."""
    import torch
    from emgraph_b200 import _lib
    from emgraph_b200.engine import model_id
    from emgraph_b200.evaluation import EvalDataset
    ds = EvalDataset(test, X)
    ds.build_filter(eng, w["E"], w["R"])
    use_tc = ((w["model"] != "TransE") if args.rank_tc < 0 else bool(args.rank_tc)) and eng.has_tensor_core_rank()
    ranks = eng.rank(model_id(w["model"]), w["k"], ent, rel, ds.test_device(eng.tdev), side=0, strategy=0, filtered=True,
                     use_tensor_cores=use_tc)
    torch.cuda.synchronize()
    return ranks.cpu().numpy()


def ranks_digest(r):
    import hashlib
    r = np.ascontiguousarray(np.asarray(r, np.int32))
    return {"mrr_pretrain": float(np.mean(1.0 / r.reshape(-1).astype(np.float64))), "sha1": hashlib.sha1(r.tobytes()).hexdigest()[:16], "n": int(r.size)}


def train_bench_single(args, name, w, steps, warmup, detailed):
    """Single-GPU train measurement of one workload through the public model API.  detailed: also per-kernel phases,
    roofline, warm-L2 and synchronous e2e numbers (the main record); the sub-records keep value + e2e."""
    import torch
    from emgraph_b200.engine import get_engine
    eng = get_engine(0)
    dev = eng.tdev
    peaks = load_peaks()
    E, R, k, eta = w["E"], w["R"], w["k"], w["eta"]
    K = internal_k(w["model"], k)
    B = batch_positives(w)
    X, test = make_dataset(w, (steps * 4 + warmup + 16) * B, with_test=True, zipf=not args.uniform)
    nb = max(1, X.shape[0] // B)
    model = build_model(w, dev)
    f = model._fit_prepare(E, R)
    # ranks of the untrained seeded tables: the witness every multi-GPU line must reproduce bit for bit
    parity = ranks_digest(rank_pretrain_single(args, w, eng, f["ent"], f["rel"], X, test))
    pipeline_on = os.environ.get("KGE_PIPELINE", "1")[:1] != "0"
    Xd = torch.from_numpy(X).to(dev)
    Xh = torch.from_numpy(X).pin_memory()
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev)

    def batch(i):
        b = i % nb
        return b * B, (b + 1) * B

    it = 0
    first_loss = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for s_ in range(warmup):
        lo, hi = batch(it)
        if s_ == warmup - 1:
            e0.record()
        model._fit_step_device(Xd[lo:hi])
        if s_ == warmup - 1:
            e1.record()
        if it == 0:
            first_loss = float(f["loss_dev"].item())
        it += 1
    torch.cuda.synchronize()
    est = e0.elapsed_time(e1)
    rep = inner_repeat(steps, est)
    n_timed = steps * rep
    sampler = ClockSampler(0)
    sampler.start()

    # ---- value: device-resident inputs, L2 flushed before every timed step.  In-order steps: with pipelined steps
    # (KGE_F_PIPELINE) the corruption generator + sort of step t+1 could slip into the UNTIMED flush between two
    # steps, so the pipeline is switched off here; the back-to-back loops below (no untimed gaps) keep it on.
    f["pipeline"] = False
    l0 = eng.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_timed)]
    torch.cuda.synchronize()
    for s in range(n_timed):
        flush.fill_(float(s))
        lo, hi = batch(it)
        evs[s][0].record()
        model._fit_step_device(Xd[lo:hi])
        evs[s][1].record()
        it += 1
    torch.cuda.synchronize()
    launches_timed = eng.launches - l0
    t_cold_ms = sum(a.elapsed_time(b) for a, b in evs)
    triples_per_step = B * (1 + eta)
    value = n_timed * triples_per_step / (t_cold_ms * 1e-3)
    step_ms = t_cold_ms / n_timed
    out = {"value": value, "ms_per_step": step_ms, "inner_repeat": rep, "timed_steps": n_timed, "timed_region_ms": t_cold_ms,
           "launches": launches_timed, "first_step_loss": first_loss, "rank_parity": parity}

    # ---- warm: steps back to back (what a training loop sees; tables stay in L2 when they fit), pipelined
    f["pipeline"] = pipeline_on
    for _ in range(2):
        lo, hi = batch(it)
        model._fit_step_device(Xd[lo:hi])
        it += 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for s in range(n_timed):
        lo, hi = batch(it)
        model._fit_step_device(Xd[lo:hi])
        it += 1
    e1.record()
    torch.cuda.synchronize()
    t_warm_ms = e0.elapsed_time(e1)
    out["value_warm_l2"] = n_timed * triples_per_step / (t_warm_ms * 1e-3)
    out["ms_per_step_warm"] = t_warm_ms / n_timed

    if detailed:
        # ---- per-kernel time inside the real step (library-side CUDA events on the launching stream, L2
        # flushed before every step): emit | fwd_bwd | reduce_apply (after the hidden sort) | span/hub reduction
        f["pipeline"] = False
        eng.set_timing(True)
        for s in range(min(n_timed, 200)):
            flush.fill_(float(s))
            lo, hi = batch(it)
            model._fit_step_device(Xd[lo:hi])
            it += 1
        torch.cuda.synchronize()
        phases, n_ph = eng.get_timing_ex()
        eng.set_timing(False)
        f["pipeline"] = pipeline_on
        # reduce_apply = the level-1 reduction kernel alone, from its launch point on the stream (an event recorded behind
        # the wait for the side-stream sort) to its end; the wait itself is reported as sort_wait
        t_emit, t_fb, t_apply, t_span = phases["emit"], phases["fwd_bwd"], phases["reduce_apply"], phases["spans"]
        t_wait = phases["sort_wait"]
        S = (3 + eta) * B
        keys = torch.empty(S, dtype=torch.int32, device=dev)
        uniq = []
        for j in range(3):
            lo, hi = batch(it + j)
            a = eng.train_args(ent=f["ent"], rel=f["rel"], pos=Xd[lo:hi], loss_out=f["loss_dev"], side=0, step=f["step"] + 1 + j, **f["kw"], **f["st"])
            eng.train_emit(a, keys)
            uniq.append(int(torch.unique(keys).numel()))
        n_unique = float(np.mean(uniq))
        # roofline of the dominant training kernel.  Algorithmic bytes per launch:
        #   fwd_bwd      : (3+eta) rows gathered + 5 rows + eta coefficients/flags written, per positive
        #   reduce_apply : one row-sized read + 13 B of key/slot/coefficient per slot, plus w,m,v read and
        #                  written once per DISTINCT touched row (Adam: 6 row-sized accesses)
        n_state = {"adam": 6, "adagrad": 4, "momentum": 4, "sgd": 2}[w["opt"]]
        bytes_fb = ((3 + eta) * 4 * K + 5 * 4 * K + 5 * eta) * B
        bytes_apply = (3 + eta) * B * (4 * K + 13) + n_state * 4 * K * n_unique
        t_red = t_apply + t_span
        dom = "kge_fwd_bwd_kernel" if t_fb >= t_red else "kge_reduce_apply_kernel (+ span/hub reduction)"
        dom_t, dom_b = (t_fb, bytes_fb) if t_fb >= t_red else (t_red, bytes_apply)
        ach = dom_b / (dom_t * 1e-3) / 1e9
        traffic, tsrc = load_traffic(name, dom.split(" ")[0])
        out["roofline"] = {
            "bound": "hbm", "kernel": dom, "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"],
            "traffic": traffic, "traffic_source": tsrc, "peak_source": peaks["src"], "algorithmic_bytes_per_launch": dom_b, "kernel_ms": dom_t,
            "phases_ms": {"emit": t_emit, "fwd_bwd": t_fb, "sort_wait": t_wait, "reduce_apply": t_apply, "span_hub": t_span,
                          "sort_done_after_emit": phases.get("sort_after_emit", 0.0), "timed_steps": n_ph},
            "kernels": {"fwd_bwd": {"bytes": bytes_fb, "GBps": bytes_fb / (t_fb * 1e-3) / 1e9, "frac": bytes_fb / (t_fb * 1e-3) / 1e9 / peaks["hbm"]},
                        "reduce_apply": {"bytes": bytes_apply, "GBps": bytes_apply / (t_red * 1e-3) / 1e9,
                                         "frac": bytes_apply / (t_red * 1e-3) / 1e9 / peaks["hbm"]}},
            "distinct_rows_per_step": n_unique,
            "step_algorithmic_GBps": (bytes_fb + bytes_apply) / (step_ms * 1e-3) / 1e9,
            "step_frac": (bytes_fb + bytes_apply) / (step_ms * 1e-3) / 1e9 / peaks["hbm"],
            "note": "tables of cfg1-3 are L2-resident: the bytes are what the kernel loads/stores, mostly served by L2"
                    if (E * K * 4 * 3) < (100 << 20) else "tables exceed L2: HBM-bound"}

    # ---- e2e: host batches through the public step (pinned H2D of the batch + D2H of the loss, every step).
    # pipelined (what fit(host_batches) runs): the call returns once the step is queued and hands back the loss of the
    # previous step, so the GPU never waits for the host; same copies, same kernels, same order.
    last_loss = 0.0
    if detailed:
        for _ in range(3):
            lo, hi = batch(it)
            model._fit_step_host(Xh[lo:hi])
            it += 1
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in range(n_timed):
            lo, hi = batch(it)
            last_loss = model._fit_step_host(Xh[lo:hi])
            it += 1
        torch.cuda.synchronize()
        t_e2e_sync = time.perf_counter() - t0
        assert math.isfinite(last_loss), "training diverged in the benchmark"
    for _ in range(3):
        lo, hi = batch(it)
        model._fit_step_host_pipelined(Xh[lo:hi])
        it += 1
    model._fit_host_flush()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n_losses = 0
    for s in range(n_timed):
        lo, hi = batch(it)
        lv = model._fit_step_host_pipelined(Xh[lo:hi])
        if lv is not None:
            last_loss = lv
            n_losses += 1
        it += 1
    last_loss = model._fit_host_flush()
    n_losses += 1
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    assert math.isfinite(last_loss) and n_losses == n_timed, "training diverged in the benchmark / a loss was not read back"
    out["clocks"] = sampler.stop()
    out["e2e"] = {"value": n_timed * triples_per_step / t_e2e, "unit": "triples/s", "h2d_bytes_per_step": B * 12, "d2h_bytes_per_step": 4,
                  "ms_per_step": 1e3 * t_e2e / n_timed,
                  "api": "EmbeddingModel._fit_step_host_pipelined -> kge_train_step_host_async / kge_train_host_wait (the loss of step t "
                         "is read while step t+1 runs; every step copies its batch in and its loss out)"}
    if detailed:
        out["e2e"]["synchronous"] = {"value": n_timed * triples_per_step / t_e2e_sync, "ms_per_step": 1e3 * t_e2e_sync / n_timed,
                                     "api": "EmbeddingModel._fit_step_host -> kge_train_step_host (returns its own loss)"}
    out["_ctx"] = (eng, model, f, X, test)
    return out


def single_gpu_main(args, w):
    import torch
    torch.cuda.set_device(0)
    steps, warmup = args.steps, args.warmup
    do_rank = not args.no_rank
    r = train_bench_single(args, args.workload, w, steps, warmup, detailed=True)
    eng, model, f, X, test = r.pop("_ctx")
    peaks = load_peaks()
    pipeline_on = os.environ.get("KGE_PIPELINE", "1")[:1] != "0"
    cfg = config_for(args, args.workload, w, 1)
    line = {
        "metric": TRAIN_METRIC, "value": r["value"], "unit": "triples/s", "n_gpus": 1, "steps": steps, "warmup": warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg,
        "measurement": {"inner_repeat": r["inner_repeat"], "timed_steps": r["timed_steps"], "timed_region_ms": r["timed_region_ms"],
                        "step_pipelining": ("value: off (in-order steps, every step's work inside its own timed window); value_warm_l2 and "
                                            "e2e: corruption generation + sort of step t+1 overlap step t") if pipeline_on else "off",
                        "optimizer": "stateful sparse " + w["opt"]},
        "value_warm_l2": r["value_warm_l2"], "ms_per_step_warm": r["ms_per_step_warm"],
        "e2e": r["e2e"], "gpu_launches": r["launches"], "clocks": r["clocks"], "roofline": r["roofline"],
        "rank_parity": dict(r["rank_parity"], equal_to_single_gpu=True, note="this IS the single-GPU run: the digest every N must print"),
        "first_step_loss": r["first_step_loss"],
    }
    if do_rank:
        line["rank"] = bench_rank(args, w, eng, model, f, X, test, peaks)
    del model, f
    torch.cuda.empty_cache()
    if not args.no_sub:
        if args.workload != "cfg5":
            line["cfg5"] = sub_record_single(args, "cfg5", max(10, steps // 2), warmup)
        line["others"] = {nm: sub_record_single(args, nm, max(10, steps // 2), warmup, brief=True)
                          for nm in ("cfg1", "cfg2", "cfg4") if nm != args.workload}
    # ---- CPU baseline (reference-equivalent op graph on the host cores), bounded sample
    if not args.no_cpu:
        B = batch_positives(w)
        tr, rk, cores = cpu_arm(w, X[: min(X.shape[0], 40 * B)], test if do_rank else X[:8], X if do_rank else None, steps=12, warmup=1,
                                budget_s=15.0, rank_budget_s=12.0, do_rank=do_rank)
        line["cpu_baseline"] = {"value": tr["value"], "unit": "triples/s", "cores": cores, "kind": "port", "sample": tr["sample"]}
        if do_rank and rk is not None:
            line["rank"]["cpu_baseline"] = {"value": rk["value"], "unit": "test triples/s", "cores": cores, "kind": "port", "sample": rk["sample"]}
    print(json.dumps(line), flush=True)
    return 0


def sub_record_single(args, name, steps, warmup, brief=False):
    """Train measurement of another BASELINE config on this GPU (same protocol as the main record)."""
    import torch
    w = dict(WORKLOADS[name])
    try:
        r = train_bench_single(args, name, w, steps, warmup, detailed=False)
    except Exception as e:  # the main record must survive a failing sub-record
        return {"error": "%s: %s" % (type(e).__name__, e)}
    r.pop("_ctx")
    torch.cuda.empty_cache()
    rec = {"workload": "%s: %s" % (name, w["desc"]), "metric": TRAIN_METRIC, "value": r["value"], "unit": "triples/s", "ms_per_step": r["ms_per_step"],
           "e2e": {"value": r["e2e"]["value"], "ms_per_step": r["e2e"]["ms_per_step"], "h2d_bytes_per_step": r["e2e"]["h2d_bytes_per_step"],
                   "d2h_bytes_per_step": 4}, "value_warm_l2": r["value_warm_l2"], "timed_steps": r["timed_steps"], "n_gpus": 1,
           "batch_positives_per_gpu": batch_positives(w)}
    if not brief:
        rec.update(rank_parity=dict(r["rank_parity"], equal_to_single_gpu=True), first_step_loss=r["first_step_loss"], scaling="weak",
                   nvlink_bytes_per_gpu_per_step=0, phases_ms=None)
    return rec


# ------------------------------------------------------------------------------------------------
# multi-GPU: one process per GPU, tables column-sharded, the global batch on every GPU
# ------------------------------------------------------------------------------------------------
def sharded_record(args, name, w, rank, world, steps, warmup, detailed):
    """Train measurement of one workload on `world` GPUs (weak scaling: B positives per GPU and step) + the
    correctness witnesses: pre-training ranks and first-step loss against a single-GPU run of the same seeded tables
    and the same global batch, computed on rank 0 inside this run."""
    import torch
    import torch.distributed as dist
    from emgraph_b200 import _lib
    from emgraph_b200 import distributed as D
    from emgraph_b200.engine import model_id
    from emgraph_b200.evaluation import EvalDataset
    local = int(os.environ.get("LOCAL_RANK", rank))
    dev = torch.device("cuda", local)
    E, R, k, eta = w["E"], w["R"], w["k"], w["eta"]
    K = internal_k(w["model"], k)
    B = batch_positives(w)
    n = B * world
    X, test = make_dataset(w, (steps * 4 + warmup + 16) * n, with_test=True, zipf=not args.uniform)
    kc = D.dim_width(k, world)
    c0, c1 = D.dim_range(k, world, rank)
    ent_slice = seeded_table(E, w["model"], k, dev, TABLE_SEED, c0, c1, pad_to=kc)
    rel_slice = seeded_table(R, w["model"], k, dev, TABLE_SEED + 1, c0, c1, pad_to=kc)
    sk = D.ShardedKGE(w["model"], k, eta, w["loss"], w["opt"], E, R, B, lr=w["lr"], margin=w["margin"], seed=0, device=local,
                      chunks=args.chunks, ent_slice=ent_slice, rel_slice=rel_slice)
    eng = sk.eng
    Xd = torch.from_numpy(X).to(dev)
    Xh = torch.from_numpy(X).pin_memory()
    nb = max(1, X.shape[0] // n)
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev)
    use_tc = ((w["model"] != "TransE") if args.rank_tc < 0 else bool(args.rank_tc)) and eng.has_tensor_core_rank()

    def gbatch(i):
        b = i % nb
        return b * n, (b + 1) * n

    # ---- witnesses.  Rank 0 holds the whole seeded table for a moment and runs the single-GPU kernels on it.
    ds = EvalDataset(test, X)
    ds.build_filter(eng, E, R)
    test_d = ds.test_device(dev)
    ranks_sh = sk.rank(test_d, side=0, strategy=0, filtered=True, use_tensor_cores=use_tc).cpu().numpy()
    full_ent = sk._gather_cols(sk.ent)
    full_rel = sk._gather_cols(sk.rel)
    single = {}
    if rank == 0:
        r1 = eng.rank(model_id(w["model"]), k, full_ent, full_rel, test_d, side=0, strategy=0, filtered=True, use_tensor_cores=use_tc).cpu().numpy()
        single["ranks_equal"] = bool(np.array_equal(r1, ranks_sh))
        single["digest"] = ranks_digest(r1)
        lo, hi = gbatch(0)
        loss1 = torch.zeros(1, device=dev)
        a = eng.train_args(model=model_id(w["model"]), loss=_lib.LOSS_IDS[w["loss"]], opt=_lib.OPT_IDS[w["opt"]], k=k, eta=eta, ent=full_ent,
                           rel=full_rel, pos=Xd[lo:hi].contiguous(), loss_out=loss1, flags=_lib.F_NO_UPDATE, margin=w["margin"], lr=w["lr"],
                           seed=0, step=1)
        eng.train_step(a)
        single["first_step_loss"] = float(loss1.item())
    del full_ent, full_rel
    torch.cuda.empty_cache()
    dist.barrier()

    it = 0
    first_loss = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for s_ in range(warmup):
        lo, hi = gbatch(it)
        if s_ == warmup - 1:
            e0.record()
        sk.train_step(Xd[lo:hi], pos_is_global=True)
        if s_ == warmup - 1:
            e1.record()
        if it == 0:
            first_loss = float(sk.loss_dev.item())
        it += 1
    torch.cuda.synchronize()
    est = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(est, op=dist.ReduceOp.MAX)
    rep = inner_repeat(steps, float(est.item()))
    n_timed = steps * rep
    dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()

    # ---- value: L2 flushed before every step, in-order steps (no prologue hiding in the untimed flush), max over ranks
    sk.pipeline = False
    l0 = eng.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_timed)]
    torch.cuda.synchronize()
    dist.barrier()
    for s in range(n_timed):
        flush.fill_(float(s))
        lo, hi = gbatch(it)
        evs[s][0].record()
        sk.train_step(Xd[lo:hi], pos_is_global=True)
        evs[s][1].record()
        it += 1
    torch.cuda.synchronize()
    dist.barrier()
    launches = (eng.launches - l0) * world
    t_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev)
    dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    t_ms = float(t_ms.item())
    triples_per_step = n * (1 + eta)
    out = {"value": n_timed * triples_per_step / (t_ms * 1e-3), "ms_per_step": t_ms / n_timed, "inner_repeat": rep, "timed_steps": n_timed,
           "timed_region_ms": t_ms, "launches": launches}

    # ---- warm: back to back, pipelined (emit + sort of step t+1 beside step t), one event pair, max over ranks
    sk.pipeline = True
    for _ in range(2):
        lo, hi = gbatch(it)
        sk.train_step(Xd[lo:hi], pos_is_global=True)
        it += 1
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(n_timed):
        lo, hi = gbatch(it)
        sk.train_step(Xd[lo:hi], pos_is_global=True)
        it += 1
    e1.record()
    torch.cuda.synchronize()
    tw = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    out["value_warm_l2"] = n_timed * triples_per_step / (float(tw.item()) * 1e-3)
    out["ms_per_step_warm"] = float(tw.item()) / n_timed

    # ---- e2e: every rank copies ITS batch from pinned host memory every step, the global loss is read back every step
    # (one step late: the loss of step t is read while step t+1 runs)
    def my_batch(i):
        lo, _ = gbatch(i)
        return lo + rank * B, lo + (rank + 1) * B

    for _ in range(3):
        lo, hi = my_batch(it)
        sk.train_step_host(Xh[lo:hi])
        it += 1
    sk.host_flush()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    n_losses, last = 0, 0.0
    for s in range(n_timed):
        lo, hi = my_batch(it)
        lv = sk.train_step_host(Xh[lo:hi])
        if lv is not None:
            last = lv
            n_losses += 1
        it += 1
    last = sk.host_flush()
    n_losses += 1
    torch.cuda.synchronize()
    dist.barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
    dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    t_e2e = float(t_e2e.item())
    assert math.isfinite(last) and n_losses == n_timed, "training diverged in the benchmark / a loss was not read back"
    out["clocks"] = sampler.stop()
    out["e2e"] = {"value": n_timed * triples_per_step / t_e2e, "unit": "triples/s", "h2d_bytes_per_step": B * 12 * world,
                  "d2h_bytes_per_step": 4 * world, "ms_per_step": 1e3 * t_e2e / n_timed,
                  "api": "ShardedKGE.train_step_host (every rank copies its own batch in, NCCL all-gather of the batches, loss read back one step late)"}

    # ---- per-phase time inside the real step (CUDA events on the launching stream, L2 flushed, max over ranks)
    sk.pipeline = False
    sk.timing = True
    for s in range(min(n_timed, 20)):
        flush.fill_(float(s))
        lo, hi = gbatch(it)
        sk.train_step(Xd[lo:hi], pos_is_global=True)
        it += 1
    ph = sk.phase_times()
    sk.timing = False
    names = sorted(ph)
    pt = torch.tensor([ph[k_] for k_ in names], device=dev)
    dist.all_reduce(pt, op=dist.ReduceOp.MAX)
    ph = {k_: float(v) for k_, v in zip(names, pt.tolist())}
    out["phases_ms_max_over_ranks"] = ph
    payload = 4 * (1 + eta) * n
    out["nvlink_bytes_per_gpu_per_step"] = int(2 * (world - 1) / world * payload)
    out["allreduce_payload_bytes"] = payload
    # witnesses
    flags = torch.tensor([1.0 if single.get("ranks_equal", False) else 0.0, single.get("first_step_loss", 0.0)], device=dev, dtype=torch.float64)
    dist.broadcast(flags, 0)
    l_single = float(flags[1].item())
    out["rank_parity"] = dict(ranks_digest(ranks_sh), equal_to_single_gpu=bool(flags[0].item() == 1.0),
                              how="rank 0 gathered the seeded tables and ran the single-GPU sweep on the same test set in this run")
    out["first_step_loss"] = first_loss
    out["first_step_loss_single_gpu"] = l_single
    out["first_step_loss_rel_err"] = abs(first_loss - l_single) / max(abs(l_single), 1e-30)
    assert out["first_step_loss_rel_err"] < 1e-5, "sharded first-step loss %r != single-GPU loss %r on the concatenated batch" % (first_loss, l_single)
    assert out["rank_parity"]["equal_to_single_gpu"], "sharded pre-training ranks differ from the single-GPU ranks"

    if detailed:
        # per-GPU algorithmic HBM bytes of the step: both phase kernels gather (3+eta) row slices per
        # positive of the GLOBAL batch, the reduction reads one slice per slot and w, m, v of every distinct touched row
        Kc = sk.Kc
        S = (3 + eta) * n
        keys = torch.empty(S, dtype=torch.int32, device=dev)
        lo, hi = gbatch(it)
        a = sk.make_args(Xd[lo:hi], step=sk.step + 1)
        eng.train_emit(a, keys)
        n_unique = int(torch.unique(keys).numel())
        n_state = {"adam": 6, "adagrad": 4, "momentum": 4, "sgd": 2}[w["opt"]]
        b_part = ((3 + eta) * 4 * Kc + 4 * (1 + eta) + 5 * eta) * n
        b_bwd = ((3 + eta) * 4 * Kc + 5 * 4 * Kc + 4 * (1 + eta) + 10 * eta) * n
        b_apply = S * (4 * Kc + 13) + n_state * 4 * Kc * n_unique
        peaks = load_peaks()
        step_ms = out["ms_per_step"]
        t_apply = ph.get("reduce_apply", step_ms)
        out["roofline"] = {"bound": "hbm", "kernel": "kge_reduce_apply_group_kernel (+ span/hub reduction) on the column slice" if Kc <= 64
                           else "kge_reduce_apply_kernel (+ span/hub reduction) on the column slice",
                           "achieved": b_apply / (t_apply * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s per GPU",
                           "frac": b_apply / (t_apply * 1e-3) / 1e9 / peaks["hbm"], "traffic": None, "peak_source": peaks["src"],
                           "algorithmic_bytes_per_launch": b_apply, "kernel_ms": t_apply, "distinct_rows_per_step": n_unique,
                           "bytes_per_gpu": {"partial": b_part, "backward": b_bwd, "reduce_apply": b_apply},
                           "step_algorithmic_GBps_per_gpu": (b_part + b_bwd + b_apply) / (step_ms * 1e-3) / 1e9,
                           "step_frac": (b_part + b_bwd + b_apply) / (step_ms * 1e-3) / 1e9 / peaks["hbm"],
                           "nvlink": {"bytes_per_gpu_per_step": out["nvlink_bytes_per_gpu_per_step"], "peak_GBps": 770.0,
                                      "time_at_peak_ms": out["nvlink_bytes_per_gpu_per_step"] / 770e9 * 1e3}}
    out["_ctx"] = (sk, ds, X, test, use_tc)
    return out


def multi_gpu_main(args, w, rank, world):
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    # NCCL's kernels on a high-priority stream: the all-reduce of one piece of the partial scores must get SM slots while the
    # phase kernel of the next piece is filling the machine
    opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
    dev = torch.device("cuda", local)
    steps, warmup = args.steps, args.warmup
    do_rank = not args.no_rank
    r = sharded_record(args, args.workload, w, rank, world, steps, warmup, detailed=True)
    sk, ds, X, test, use_tc = r.pop("_ctx")
    eng = sk.eng
    E, R, k, eta = w["E"], w["R"], w["k"], w["eta"]
    K = internal_k(w["model"], k)
    line = {
        "metric": TRAIN_METRIC, "value": r["value"], "unit": "triples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_for(args, args.workload, w, world),
        "measurement": {"inner_repeat": r["inner_repeat"], "timed_steps": r["timed_steps"], "timed_region_ms": r["timed_region_ms"],
                        "chunks": args.chunks, "optimizer": "stateful sparse " + w["opt"],
                        "step_pipelining": "value: off (in-order steps); value_warm_l2: corruption generation + sort of step t+1 overlap step t"},
        "value_warm_l2": r["value_warm_l2"], "ms_per_step_warm": r["ms_per_step_warm"], "e2e": r["e2e"],
        "gpu_launches": r["launches"], "clocks": r["clocks"], "roofline": r["roofline"], "rank_parity": r["rank_parity"],
        "first_step_loss": r["first_step_loss"], "first_step_loss_single_gpu": r["first_step_loss_single_gpu"],
        "phases_ms_max_over_ranks": r["phases_ms_max_over_ranks"], "nvlink_bytes_per_gpu_per_step": r["nvlink_bytes_per_gpu_per_step"],
    }

    if do_rank:
        T = w["T"]
        flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev)
        t_filter = 0.0
        test_d = ds.test_device(dev)
        test_h = ds.test_host_pinned()
        for _ in range(2):
            ranks = sk.rank(test_d, side=0, strategy=0, filtered=True, use_tensor_cores=use_tc)
        torch.cuda.synchronize()
        dist.barrier()
        nr = max(1, args.rank_steps)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nr)]
        l0 = eng.launches
        for s in range(nr):
            flush.fill_(float(s))
            evs[s][0].record()
            ranks = sk.rank(test_d, side=0, strategy=0, filtered=True, use_tensor_cores=use_tc)
            evs[s][1].record()
        torch.cuda.synchronize()
        dist.barrier()
        r_launches = (eng.launches - l0) * world
        tr = torch.tensor([sum(a.elapsed_time(b) for a, b in evs) / nr], device=dev)
        dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        tr = float(tr.item())
        stage_t = torch.empty((T, 3), dtype=torch.int32, device=dev)
        out_h = torch.empty((T, 2), dtype=torch.int32).pin_memory()
        dist.barrier()
        t0 = time.perf_counter()
        for s in range(nr):
            stage_t.copy_(test_h, non_blocking=True)
            out_h.copy_(sk.rank(stage_t, side=0, strategy=0, filtered=True, use_tensor_cores=use_tc))
            torch.cuda.synchronize()
        dist.barrier()
        te = torch.tensor([(time.perf_counter() - t0) / nr], device=dev)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        te = float(te.item())
        rk = out_h.numpy()
        assert rk.min() >= 1
        flops = 4.0 * E * K * T
        tf32 = measure_tf32_peak(dev) if use_tc else None
        peak = (tf32 if tf32 else 148 * 128 * 2 * 1.965e9 / 1e12) * world
        line["rank"] = {"metric": RANK_METRIC, "value": T / (tr * 1e-3), "unit": "test triples/s", "ms_per_step": tr, "steps": nr, "T": T,
                        "scaling": "strong (the trained column slices are transposed once into row-range shards -- cached, outside the timed "
                                   "sweep -- every GPU sweeps its rows, counts all-reduced)", "corrupt_side": "s,o",
                        "filter_triples": int(X.shape[0]), "filter_build_ms": 1e3 * t_filter, "tensor_cores": bool(use_tc),
                        "mrr": float(np.mean(1.0 / rk.reshape(-1))),
                        "e2e": {"value": T / te, "unit": "test triples/s", "h2d_bytes_per_step": T * 12 * world, "d2h_bytes_per_step": T * 8 * world,
                                "api": "ShardedKGE.rank (host test triples, host ranks)"},
                        "gpu_launches": r_launches,
                        "roofline": {"bound": "tensor" if use_tc else "fp32-alu", "achieved": flops / (tr * 1e-3) / 1e12, "peak": peak,
                                     "unit": "TFLOP/s (logical; x3 TF32 MMAs issued), whole job", "frac": flops / (tr * 1e-3) / 1e12 / peak,
                                     "peak_source": "cuBLAS TF32 GEMM 8192^3 measured in this run x %d GPUs" % world if use_tc else "fp32 FMA issue rate",
                                     "traffic": None, "kernel_ms": tr}}
    del sk, ds
    torch.cuda.empty_cache()
    if not args.no_sub and args.workload != "cfg5":
        w5 = dict(WORKLOADS["cfg5"])
        try:
            r5 = sharded_record(args, "cfg5", w5, rank, world, max(10, steps // 2), warmup, detailed=True)
            r5.pop("_ctx")
            line["cfg5"] = {"workload": "cfg5: %s" % w5["desc"], "metric": TRAIN_METRIC, "unit": "triples/s", "n_gpus": world, "scaling": "weak",
                            "batch_positives_per_gpu": batch_positives(w5),
                            **{k_: r5[k_] for k_ in ("value", "ms_per_step", "value_warm_l2", "ms_per_step_warm", "timed_steps", "rank_parity",
                                                     "first_step_loss", "first_step_loss_single_gpu", "phases_ms_max_over_ranks",
                                                     "nvlink_bytes_per_gpu_per_step", "roofline")},
                            "e2e": {k_: r5["e2e"][k_] for k_ in ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step")}}
        except Exception as e:
            line["cfg5"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0


def bench_rank(args, w, eng, model, f, X, test, peaks):
    """Filtered 's,o' ranking of T test triples against all E entities; filter = all synthetic triples."""
    import torch
    from emgraph_b200.evaluation import EvalDataset
    dev = eng.tdev
    E, R, k = w["E"], w["R"], w["k"]
    K = internal_k(w["model"], k)
    T = test.shape[0]
    model._fit_finish()
    ent, rel = model._device_params()
    use_tc = (w["model"] != "TransE") if args.rank_tc < 0 else bool(args.rank_tc)
    use_tc = use_tc and eng.has_tensor_core_rank()
    model.engine_params["rank_tensor_cores"] = use_tc
    ds = EvalDataset(test, X)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ds.build_filter(eng, E, R)
    torch.cuda.synchronize()
    t_filter = time.perf_counter() - t0
    model.set_filter_for_eval()
    model.configure_evaluation_protocol({"corrupt_side": "s,o", "ranking_strategy": "worst"})
    test_d = ds.test_device(dev)
    mid = model._model_id()
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev)
    counts = torch.empty((T, 2, 4), dtype=torch.int32, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(2):
        if i == 1:
            e0.record()
        eng.rank_counts(mid, k, ent, rel, test_d, side=0, filtered=True, use_tensor_cores=use_tc, counts=counts)
        eng.rank_finalize(counts, side=0, strategy=0, filtered=True)
        if i == 1:
            e1.record()
    torch.cuda.synchronize()
    n = max(1, args.rank_steps) * inner_repeat(max(1, args.rank_steps), e0.elapsed_time(e1))
    l0 = eng.launches
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n)]
    # the tensor-core sweep runs into the board's power cap (sw_power_cap, SM clock ~1.55 GHz): the ranking record carries
    # the clocks and reasons of ITS timed region, the top-level "clocks" are those of the training steps
    rank_sampler = ClockSampler(dev.index or 0)
    rank_sampler.start()
    for s in range(n):
        flush.fill_(float(s))
        evs[s][0].record()
        eng.rank_counts(mid, k, ent, rel, test_d, side=0, filtered=True, use_tensor_cores=use_tc, counts=counts)
        evs[s][1].record()
        ranks = eng.rank_finalize(counts, side=0, strategy=0, filtered=True)
        evs[s][2].record()
    torch.cuda.synchronize()
    rank_clocks = rank_sampler.stop()
    launches = eng.launches - l0
    t_ms = sum(e[0].elapsed_time(e[2]) for e in evs) / n
    t_sweep_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / n
    value = T / (t_ms * 1e-3)
    # e2e: the public evaluation hook with host buffers (pinned H2D of the test triples, D2H of the ranks)
    model.get_ranks(ds)
    t0 = time.perf_counter()
    for _ in range(n):
        r_host = model.get_ranks(ds)
    t_e2e = (time.perf_counter() - t0) / n
    assert r_host.shape == (T, 2) and r_host.min() >= 1
    np.testing.assert_array_equal(r_host, ranks.cpu().numpy())
    mrr = float(np.mean(1.0 / r_host.reshape(-1)))
    flops = 4.0 * E * K * T  # 2 sides x E candidates x K MACs x 2
    if w["model"] == "TransE":
        # fp32 CUDA-core sweep: per (query, candidate, column) two ALU instructions -- FADD d = q - e, then FADD
        # acc += |d| (the abs is a source modifier; checked in SASS) -- against the fp32 pipe's issue rate of
        # 128 lanes per SM per clock (the FMA peak counts 2 flop per lane-slot; an add fills a slot all the same)
        ach = 2.0 * 2 * E * K * T / (t_sweep_ms * 1e-3) / 1e12
        peak = 148 * 128 * 1.965e9 / 1e12
        roof = {"bound": "fp32-alu", "achieved": ach, "peak": peak, "unit": "T lane-instr/s (FADD sub + FADD |.| accumulate per element)",
                "frac": ach / peak, "traffic": None, "peak_source": "148 SMs x 128 fp32 lanes x 1.965 GHz (max SM clock)"}
    elif use_tc:
        # three tensor-core MMAs per logical MAC (hi*hi + hi*lo + lo*hi); the denominator is the dense TF32 rate of this GPU
        # measured in this run (cuBLAS fp32 GEMM with TF32 tensor cores, same recipe as MEASURED_PEAKS.json's bf16)
        peak = measure_tf32_peak(dev)
        ach = flops / (t_sweep_ms * 1e-3) / 1e12
        traffic, tsrc = load_traffic(args.workload, "kge_rank_tc_kernel")
        roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s (logical fp32-accurate flops; x3 tensor-core MMAs issued)",
                "frac": ach / peak, "frac_issued_3x": 3 * ach / peak, "traffic": traffic, "traffic_source": tsrc,
                "peak_source": "cuBLAS TF32 GEMM 8192^3, best of 10, measured in this run (bf16 peak / 2 would be %.0f)" % (peaks["bf16"] / 2.0)}
    else:
        peak = 148 * 128 * 2 * 1.965e9 / 1e12
        ach = flops / (t_sweep_ms * 1e-3) / 1e12
        roof = {"bound": "fp32-alu", "achieved": ach, "peak": peak, "unit": "TFLOP/s fp32 FMA", "frac": ach / peak, "traffic": None}
    v1 = os.environ.get("KGE_SWEEP_V1", "0")[:1] == "1"
    roof["kernel"] = "kge_rank_sweep_tc" if use_tc else ("kge_rank_sweep2_kernel" if w["model"] == "TransE" and not v1 else "kge_rank_sweep_kernel")
    roof["kernel_ms"] = t_sweep_ms
    return {"metric": RANK_METRIC, "value": value, "unit": "test triples/s", "ms_per_step": t_ms, "steps": n, "T": T,
            "corrupt_side": "s,o", "filter_triples": int(X.shape[0]), "filter_build_ms": 1e3 * t_filter, "tensor_cores": bool(use_tc),
            "mrr": mrr, "e2e": {"value": T / t_e2e, "unit": "test triples/s", "h2d_bytes_per_step": T * 12, "d2h_bytes_per_step": T * 8,
                                "api": "EmbeddingModel.get_ranks -> kge_rank_host"},
            "gpu_launches": launches, "roofline": roof, "clocks": rank_clocks}


if __name__ == "__main__":
    sys.exit(main() or 0)
