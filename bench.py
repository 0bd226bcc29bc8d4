#!/usr/bin/env python
"""Benchmark of the KGE hot path: train step (score + eta-negative corruption + loss + backward +
sparse optimizer) and filtered ranking, on synthetic triples of the BASELINE.json dataset shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--impl b200|reference]

Prints ONE JSON line (rank 0).  Contract (see DESIGN.md "Measurement"):
  value      train triples/sec (eta negatives incl.), inputs resident in HBM, L2 flushed between the
             timed steps, CUDA-event timed on the launching stream, max over ranks
  e2e        the same metric through the public API with HOST buffers: every step copies its batch
             from pinned host memory and reads the batch loss back (EmbeddingModel._fit_step_host ->
             C ABI kge_train_step_host)
  roofline   dominant training kernel: algorithmic bytes / live CUDA-event kernel time vs the measured
             HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the reference-equivalent CPU op graph (oracle/torch_port.py) on the host cores, on a
             bounded sample
  rank       the second half of BASELINE.json's metric: filtered-rank test triples/sec (same keys)
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
            self._stop.wait(0.05)

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
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic_r01.json);
    None when the workload / kernel was not captured."""
    p = os.path.join(ROOT, "profiles", "traffic_r01.json")
    try:
        return float(json.load(open(p))[workload][kernel])
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference-equivalent op graph (oracle/torch_port.py) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_arm(w, X, test, filt_for_rank, steps, warmup, budget_s, rank_budget_s, do_rank=True):
    import torch
    from oracle import torch_port as tp
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    K = internal_k(w["model"], w["k"])
    B = int(math.ceil(w["N"] / w["batches"]))
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
        rank = dict(value=n / dt, sample="%d of %d test triples, per-triple 2E-corruption sweep, dict filter, torch-CPU %d threads"
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
    ap.add_argument("--rank-steps", type=int, default=3)
    ap.add_argument("--rank-tc", type=int, default=-1, help="1/0 force the tensor-core ranking sweep on/off")
    ap.add_argument("--uniform", action="store_true", help="uniform instead of Zipf entity popularity")
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
    B = int(math.ceil(w["N"] / w["batches"]))
    need = (args.steps + args.warmup + 1) * B
    X = synth_triples(w["E"], w["R"], min(w["N"], max(need, B)), seed=0, zipf=not args.uniform)
    test = X[:: max(1, X.shape[0] // 512)][:512]
    train, rank, cores = cpu_arm(w, X, test, X, args.steps, args.warmup, budget_s=150.0, rank_budget_s=20.0, do_rank=not args.no_rank)
    line = {
        "impl": "reference", "metric": TRAIN_METRIC, "value": train["value"], "unit": "triples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": train["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: %s" % (args.workload, w["desc"]), "batch_positives": B, "eta": w["eta"], "E": w["E"],
                   "K": internal_k(w["model"], w["k"])},
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


def single_gpu_main(args, w):
    import torch
    from emgraph_b200 import _lib, models
    from emgraph_b200.engine import get_engine, make_table
    from emgraph_b200.evaluation import EvalDataset

    torch.cuda.set_device(0)
    eng = get_engine(0)
    dev = eng.tdev
    peaks = load_peaks()
    E, R, k, eta = w["E"], w["R"], w["k"], w["eta"]
    K = internal_k(w["model"], k)
    B = int(math.ceil(w["N"] / w["batches"]))
    steps, warmup = args.steps, args.warmup
    do_rank = not args.no_rank
    n_train_needed = (steps + warmup) * B
    X = synth_triples(E, R, w["N"] if do_rank else min(w["N"], n_train_needed), seed=0, zipf=not args.uniform)
    nb = max(1, X.shape[0] // B)
    rng = np.random.Generator(np.random.PCG64(1))
    test = X[rng.permutation(X.shape[0])[: w["T"]]].copy() if do_rank else None

    cls = models.MODEL_REGISTRY[w["model"]]
    model = cls(k=k, eta=eta, epochs=1, batches_count=w["batches"], seed=0, optimizer=w["opt"], optimizer_params={"lr": w["lr"]},
                loss=w["loss"], loss_params={"margin": w["margin"]},
                initializer="constant", initializer_params={"entity": glorot(E, K, 2), "relation": glorot(R, K, 3)})
    f = model._fit_prepare(E, R)
    pipeline_on = os.environ.get("KGE_PIPELINE", "1")[:1] != "0"
    Xd = torch.from_numpy(X).to(dev)
    Xh = torch.from_numpy(X).pin_memory()
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev)

    def batch(i):
        b = i % nb
        return b * B, (b + 1) * B

    it = 0
    for _ in range(warmup):
        lo, hi = batch(it)
        model._fit_step_device(Xd[lo:hi])
        it += 1
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()

    # ---- value: device-resident inputs, L2 flushed before every timed step.  In-order steps: with pipelined steps
    # (KGE_F_PIPELINE) the corruption generator + sort of step t+1 could slip into the UNTIMED flush between two
    # steps, so the pipeline is switched off here; the back-to-back loops below (no untimed gaps) keep it on.
    f["pipeline"] = False
    l0 = eng.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize()
    for s in range(steps):
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
    value = steps * triples_per_step / (t_cold_ms * 1e-3)

    # ---- warm: K steps back to back (what a training loop sees; tables stay in L2 when they fit), pipelined
    f["pipeline"] = pipeline_on
    for _ in range(2):
        lo, hi = batch(it)
        model._fit_step_device(Xd[lo:hi])
        it += 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for s in range(steps):
        lo, hi = batch(it)
        model._fit_step_device(Xd[lo:hi])
        it += 1
    e1.record()
    torch.cuda.synchronize()
    t_warm_ms = e0.elapsed_time(e1)
    value_warm = steps * triples_per_step / (t_warm_ms * 1e-3)

    # ---- per-kernel time inside the real step (library-side CUDA events on the launching stream, L2
    # flushed before every step): emit | fwd_bwd | reduce_apply (after the hidden sort) | span/hub reduction
    f["pipeline"] = False
    eng.set_timing(True)
    for s in range(steps):
        flush.fill_(float(s))
        lo, hi = batch(it)
        model._fit_step_device(Xd[lo:hi])
        it += 1
    torch.cuda.synchronize()
    phases, n_timed = eng.get_timing()
    eng.set_timing(False)
    f["pipeline"] = pipeline_on
    t_emit, t_fb, t_apply, t_span = phases["emit"], phases["fwd_bwd"], phases["reduce_apply"], phases["spans"]
    # distinct rows touched per step (entities + relations), from the library's own sort keys
    S = (3 + eta) * B
    keys = torch.empty(S, dtype=torch.int32, device=dev)
    uniq = []
    for j in range(3):
        lo, hi = batch(it + j)
        a = eng.train_args(ent=f["ent"], rel=f["rel"], pos=Xd[lo:hi], loss_out=f["loss_dev"], side=0, step=f["step"] + 1 + j, **f["kw"], **f["st"])
        eng.train_emit(a, keys)
        uniq.append(int(torch.unique(keys).numel()))
    n_unique = float(np.mean(uniq))

    # ---- e2e: host batches through the public step (pinned H2D of the batch + D2H of the loss, every step).
    # (a) synchronous: every call returns its own loss (what the reference's loop does);
    # (b) pipelined (what fit(host_batches) runs): the call returns once the step is queued and hands back the
    #     loss of the previous step, so the GPU never waits for the host; same copies, same kernels, same order.
    for _ in range(3):
        lo, hi = batch(it)
        model._fit_step_host(Xh[lo:hi])
        it += 1
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    last_loss = 0.0
    for s in range(steps):
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
    for s in range(steps):
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
    e2e_value = steps * triples_per_step / t_e2e
    assert math.isfinite(last_loss) and n_losses == steps, "training diverged in the benchmark / a loss was not read back"
    clocks = sampler.stop()

    # roofline of the dominant training kernel.  Algorithmic bytes per launch:
    #   fwd_bwd      : (3+eta) rows gathered + 5 rows + eta coefficients/flags written, per positive
    #   reduce_apply : one row-sized read + 13 B of key/slot/coefficient per slot, plus w,m,v read and
    #                  written once per DISTINCT touched row (Adam: 6 row-sized accesses)
    n_state = {"adam": 6, "adagrad": 4, "momentum": 4, "sgd": 2}[w["opt"]]
    bytes_fb = ((3 + eta) * 4 * K + 5 * 4 * K + 5 * eta) * B
    bytes_apply = (3 + eta) * B * (4 * K + 13) + n_state * 4 * K * n_unique
    bytes_survey = 36 * K * (3 + eta) * B  # SURVEY 8(d) per-positive figure (no duplicate-row reuse), for reference
    t_red = t_apply + t_span
    dom = "kge_fwd_bwd_kernel" if t_fb >= t_red else "kge_reduce_apply_kernel (+ span/hub reduction)"
    dom_t, dom_b = (t_fb, bytes_fb) if t_fb >= t_red else (t_red, bytes_apply)
    ach = dom_b / (dom_t * 1e-3) / 1e9
    step_ms = t_cold_ms / steps
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"],
                "traffic": load_traffic(args.workload, dom.split(" ")[0]), "traffic_source": "ncu --set full capture, profiles/traffic_r01.json",
                "peak_source": peaks["src"], "algorithmic_bytes_per_launch": dom_b, "kernel_ms": dom_t,
                "phases_ms": {"emit": t_emit, "fwd_bwd": t_fb, "reduce_apply": t_apply, "span_hub": t_span,
                              "sort_done_after_emit": phases.get("sort_after_emit", 0.0), "timed_steps": n_timed},
                "kernels": {"fwd_bwd": {"bytes": bytes_fb, "GBps": bytes_fb / (t_fb * 1e-3) / 1e9, "frac": bytes_fb / (t_fb * 1e-3) / 1e9 / peaks["hbm"]},
                            "reduce_apply": {"bytes": bytes_apply, "GBps": bytes_apply / (t_red * 1e-3) / 1e9,
                                             "frac": bytes_apply / (t_red * 1e-3) / 1e9 / peaks["hbm"]}},
                "distinct_rows_per_step": n_unique,
                "step_algorithmic_GBps": (bytes_fb + bytes_apply) / (step_ms * 1e-3) / 1e9,
                "step_frac": (bytes_fb + bytes_apply) / (step_ms * 1e-3) / 1e9 / peaks["hbm"],
                "step_survey_accounting_GBps": bytes_survey / (step_ms * 1e-3) / 1e9,
                "note": "tables of cfg1-3 are L2-resident: the bytes are what the kernel loads/stores, mostly served by L2"
                        if (E * K * 4 * 3) < (100 << 20) else "tables exceed L2: HBM-bound"}

    line = {
        "metric": TRAIN_METRIC, "value": value, "unit": "triples/s", "n_gpus": 1, "steps": steps, "warmup": warmup,
        "ms_per_step": t_cold_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "%s: %s" % (args.workload, w["desc"]), "batch_positives": B, "eta": eta, "E": E, "R": R, "K": K,
                   "entity_popularity": "uniform" if args.uniform else "zipf(1.0)", "optimizer": "stateful sparse " + w["opt"],
                   "l2": "flushed before every timed step (%d MiB write)" % (L2_FLUSH_BYTES >> 20), "parallelism": "1 GPU",
                   "step_pipelining": ("value: off (in-order steps, every step's work inside its own timed window); value_warm_l2 and "
                                       "e2e: corruption generation + sort of step t+1 overlap step t") if pipeline_on else "off"},
        "value_warm_l2": value_warm, "ms_per_step_warm": t_warm_ms / steps,
        "e2e": {"value": e2e_value, "unit": "triples/s", "h2d_bytes_per_step": B * 12, "d2h_bytes_per_step": 4,
                "ms_per_step": 1e3 * t_e2e / steps,
                "api": "EmbeddingModel._fit_step_host_pipelined -> kge_train_step_host_async / kge_train_host_wait (the loss of step t "
                       "is read while step t+1 runs; every step copies its batch in and its loss out)",
                "synchronous": {"value": steps * triples_per_step / t_e2e_sync, "ms_per_step": 1e3 * t_e2e_sync / steps,
                                "api": "EmbeddingModel._fit_step_host -> kge_train_step_host (returns its own loss)"}},
        "gpu_launches": launches_timed, "clocks": clocks, "roofline": roofline,
    }

    # ---- ranking half of the metric
    if do_rank:
        line["rank"] = bench_rank(args, w, eng, model, f, X, test, peaks)

    # ---- CPU baseline (reference-equivalent op graph on the host cores), bounded sample
    if not args.no_cpu:
        tr, rk, cores = cpu_arm(w, X[: min(X.shape[0], 40 * B)], test if do_rank else X[:8], X if do_rank else None, steps=12, warmup=1,
                                budget_s=15.0, rank_budget_s=12.0, do_rank=do_rank)
        line["cpu_baseline"] = {"value": tr["value"], "unit": "triples/s", "cores": cores, "kind": "port", "sample": tr["sample"]}
        if do_rank and rk is not None:
            line["rank"]["cpu_baseline"] = {"value": rk["value"], "unit": "test triples/s", "cores": cores, "kind": "port", "sample": rk["sample"]}
    print(json.dumps(line), flush=True)
    return 0


def multi_gpu_main(args, w, rank, world):
    """Row-sharded table over `world` GPUs.  Weak scaling for training: every
    rank takes its own batch of B positives per step; ranking shards the entity sweep (fixed T)."""
    import torch
    import torch.distributed as dist
    from emgraph_b200 import _lib
    from emgraph_b200.distributed import ShardedKGE, batch_slice
    from emgraph_b200.evaluation import EvalDataset

    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = load_peaks()
    E, R, k, eta = w["E"], w["R"], w["k"], w["eta"]
    K = internal_k(w["model"], k)
    B = int(math.ceil(w["N"] / w["batches"]))
    steps, warmup = args.steps, args.warmup
    do_rank = not args.no_rank
    need = (steps * 2 + warmup + 8) * B * world
    X = synth_triples(E, R, w["N"] if do_rank else min(w["N"], need), seed=0, zipf=not args.uniform)
    sk = ShardedKGE(w["model"], k, eta, w["loss"], w["opt"], E, R, B, lr=w["lr"], margin=w["margin"], seed=0, device=local)
    eng, dev = sk.eng, sk.eng.tdev
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    lim_e, lim_r = math.sqrt(6.0 / (E + K)), math.sqrt(6.0 / (R + K))
    sk.ent.tensor.uniform_(-lim_e, lim_e, generator=g)
    sk.rel.copy_(torch.from_numpy(glorot(R, K, 3)).to(dev))
    dist.barrier()
    Xd = torch.from_numpy(X).to(dev)
    Xh = torch.from_numpy(X).pin_memory()
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev)
    it = 0

    def my_batch(i):
        return batch_slice(X.shape[0], world, rank, i, B)

    for _ in range(warmup):
        lo, hi = my_batch(it)
        sk.train_step(Xd[lo:hi])
        it += 1
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize()
    dist.barrier()
    for s in range(steps):
        flush.fill_(float(s))
        lo, hi = my_batch(it)
        evs[s][0].record()
        sk.train_step(Xd[lo:hi])
        evs[s][1].record()
        it += 1
    torch.cuda.synchronize()
    dist.barrier()
    launches = (eng.launches - l0) * world
    t_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev)
    dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    t_ms = float(t_ms.item())
    triples_per_step = world * B * (1 + eta)
    value = steps * triples_per_step / (t_ms * 1e-3)

    # e2e: every step the rank's batch comes from pinned host memory and the global loss is read back
    stage = torch.empty((B, 3), dtype=torch.int32, device=dev)
    for _ in range(2):
        lo, hi = my_batch(it)
        stage.copy_(Xh[lo:hi], non_blocking=True)
        float(sk.train_step(stage).item())
        it += 1
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for s in range(steps):
        lo, hi = my_batch(it)
        stage.copy_(Xh[lo:hi], non_blocking=True)
        last = float(sk.train_step(stage).item())
        it += 1
    torch.cuda.synchronize()
    dist.barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
    dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    t_e2e = float(t_e2e.item())
    assert math.isfinite(last), "training diverged in the benchmark"
    clocks = sampler.stop()

    # per-phase time inside the real step (CUDA events on the launching stream, L2 flushed, max over ranks)
    sk.timing = True
    for s in range(min(steps, 10)):
        flush.fill_(float(s))
        lo, hi = my_batch(it)
        sk.train_step(Xd[lo:hi])
        it += 1
    ph = sk.phase_times()
    sk.timing = False
    names = sorted(ph)
    pt = torch.tensor([ph[k_] for k_ in names], device=dev)
    dist.all_reduce(pt, op=dist.ReduceOp.MAX)
    ph = {k_: float(v) for k_, v in zip(names, pt.tolist())}
    # roofline of the exchange: rows pushed to the other ranks cross NVLink once ((world-1)/world of the
    # entity slots), plus the all-gathered [Qo|Qs|coef|keep] tails
    nv_bytes = (2 + eta) * 4 * K * B * (world - 1) / world
    t_push = ph.get("push", 0.0) + ph.get("push_barrier", 0.0)
    roofline = {"bound": "nvlink", "kernel": "kge_push_rows_kernel (owner-side row push, peer stores)",
                "achieved": nv_bytes / (max(t_push, 1e-6) * 1e-3) / 1e9,
                "peak": 770.0, "unit": "GB/s per direction per GPU (measured peer copy, B200_PROFILING.md)",
                "frac": nv_bytes / (max(t_push, 1e-6) * 1e-3) / 1e9 / 770.0, "traffic": None, "kernel_ms": t_push,
                "algorithmic_bytes_per_launch": nv_bytes, "phases_ms_max_over_ranks": ph,
                "tails_allgather_bytes": (world - 1) * sk.tail_stride * 4}

    line = {
        "metric": TRAIN_METRIC, "value": value, "unit": "triples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": t_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "%s: %s" % (args.workload, w["desc"]), "batch_positives_per_gpu": B, "global_batch": B * world, "eta": eta,
                   "E": E, "R": R, "K": K, "entity_popularity": "uniform" if args.uniform else "zipf(1.0)",
                   "optimizer": "stateful sparse " + w["opt"], "l2": "flushed before every timed step (%d MiB write)" % (L2_FLUSH_BYTES >> 20),
                   "parallelism": "dp%d, entity table + optimizer state row-sharded, owner-push row exchange over peer memory" % world},
        "e2e": {"value": steps * triples_per_step / t_e2e, "unit": "triples/s", "h2d_bytes_per_step": B * 12 * world,
                "d2h_bytes_per_step": 4 * world, "ms_per_step": 1e3 * t_e2e / steps, "api": "ShardedKGE.train_step (host batch, loss read back)"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
    }

    if do_rank:
        T = w["T"]
        rng = np.random.Generator(np.random.PCG64(1))
        test = X[rng.permutation(X.shape[0])[:T]].copy()
        ds = EvalDataset(test, X)
        t0 = time.perf_counter()
        ds.build_filter(eng, E, R)
        torch.cuda.synchronize()
        t_filter = time.perf_counter() - t0
        use_tc = ((w["model"] != "TransE") if args.rank_tc < 0 else bool(args.rank_tc)) and eng.has_tensor_core_rank()
        test_d = ds.test_device(dev)
        test_h = ds.test_host_pinned()
        for _ in range(2):
            ranks = sk.rank(test_d, side=0, strategy=0, filtered=True, use_tensor_cores=use_tc)
        torch.cuda.synchronize()
        dist.barrier()
        n = max(1, args.rank_steps)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        l0 = eng.launches
        for s in range(n):
            flush.fill_(float(s))
            evs[s][0].record()
            ranks = sk.rank(test_d, side=0, strategy=0, filtered=True, use_tensor_cores=use_tc)
            evs[s][1].record()
        torch.cuda.synchronize()
        dist.barrier()
        r_launches = (eng.launches - l0) * world
        tr = torch.tensor([sum(a.elapsed_time(b) for a, b in evs) / n], device=dev)
        dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        tr = float(tr.item())
        stage_t = torch.empty((T, 3), dtype=torch.int32, device=dev)
        out_h = torch.empty((T, 2), dtype=torch.int32).pin_memory()
        dist.barrier()
        t0 = time.perf_counter()
        for s in range(n):
            stage_t.copy_(test_h, non_blocking=True)
            out_h.copy_(sk.rank(stage_t, side=0, strategy=0, filtered=True, use_tensor_cores=use_tc))
            torch.cuda.synchronize()
        dist.barrier()
        te = torch.tensor([(time.perf_counter() - t0) / n], device=dev)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        te = float(te.item())
        rk = out_h.numpy()
        assert rk.min() >= 1
        flops = 4.0 * E * K * T
        peak = peaks["bf16"] / 2.0 * world
        line["rank"] = {"metric": RANK_METRIC, "value": T / (tr * 1e-3), "unit": "test triples/s", "ms_per_step": tr, "steps": n, "T": T,
                        "scaling": "strong (entity sweep sharded by row range, counts all-reduced)", "corrupt_side": "s,o",
                        "filter_triples": int(X.shape[0]), "filter_build_ms": 1e3 * t_filter, "tensor_cores": bool(use_tc),
                        "mrr": float(np.mean(1.0 / rk.reshape(-1))),
                        "e2e": {"value": T / te, "unit": "test triples/s", "h2d_bytes_per_step": T * 12 * world, "d2h_bytes_per_step": T * 8 * world,
                                "api": "ShardedKGE.rank (host test triples, host ranks)"},
                        "gpu_launches": r_launches,
                        "roofline": {"bound": "tensor" if use_tc else "fp32-alu", "achieved": flops / (tr * 1e-3) / 1e12, "peak": peak,
                                     "unit": "TFLOP/s (logical; x3 TF32 MMAs issued), whole job", "frac": flops / (tr * 1e-3) / 1e12 / peak,
                                     "traffic": None, "kernel_ms": tr}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0


def bench_rank(args, w, eng, model, f, X, test, peaks):
    """Filtered 's,o' ranking of T test triples against all E entities; filter = all synthetic triples."""
    import torch
    from emgraph_b200 import _lib
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
    n = max(1, args.rank_steps)
    counts = torch.empty((T, 2, 4), dtype=torch.int32, device=dev)
    for _ in range(2):
        eng.rank_counts(mid, k, ent, rel, test_d, side=0, filtered=True, use_tensor_cores=use_tc, counts=counts)
        eng.rank_finalize(counts, side=0, strategy=0, filtered=True)
    torch.cuda.synchronize()
    l0 = eng.launches
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n)]
    for s in range(n):
        flush.fill_(float(s))
        evs[s][0].record()
        eng.rank_counts(mid, k, ent, rel, test_d, side=0, filtered=True, use_tensor_cores=use_tc, counts=counts)
        evs[s][1].record()
        ranks = eng.rank_finalize(counts, side=0, strategy=0, filtered=True)
        evs[s][2].record()
    torch.cuda.synchronize()
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
        # 3xTF32: three tensor-core MMAs per logical MAC; TF32 dense peak = 1/2 of the measured bf16 peak
        peak = peaks["bf16"] / 2.0
        ach = flops / (t_sweep_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s (logical fp32-accurate flops; x3 TF32 MMAs issued)",
                "frac": ach / peak, "frac_issued_3x": 3 * ach / peak, "traffic": load_traffic(args.workload, "kge_rank_tc_kernel"), "peak_source": peaks["src"] + " bf16/2"}
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
            "gpu_launches": launches, "roofline": roof}


if __name__ == "__main__":
    sys.exit(main() or 0)
