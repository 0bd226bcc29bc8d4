"""Generate tests/golden/*.npz by executing the reference's OWN Python (see oracle/ref_shim.py).

Run in the builder container only (needs /root/reference):  python oracle/make_golden.py
The outputs are results (scores, losses, gradients, ranks), not reference code; they travel to the
GPU box with the repo, where /root/reference does not exist.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import kge_oracle as ko  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

TRAIN_CASES = [
    # name, model, loss, k, eta, E, R, n, side, loss_params, emb_params
    ("transe_l1_pairwise", "TransE", "pairwise", 12, 5, 64, 4, 48, "s,o", {"margin": 1.0}, {}),
    ("transe_l2_nll", "TransE", "nll", 16, 3, 64, 4, 40, "s,o", {}, {"norm": 2}),
    ("transe_l1_multiclass", "TransE", "multiclass_nll", 10, 4, 50, 3, 33, "o", {}, {}),
    ("distmult_pairwise_m5", "DistMult", "pairwise", 16, 6, 64, 5, 48, "s,o", {"margin": 5.0}, {}),
    ("distmult_nll", "DistMult", "nll", 20, 3, 40, 5, 31, "s", {}, {}),
    ("distmult_multiclass", "DistMult", "multiclass_nll", 8, 7, 64, 5, 48, "s,o", {}, {}),
    ("complex_nll", "ComplEx", "nll", 12, 5, 64, 6, 48, "s,o", {}, {}),
    ("complex_pairwise", "ComplEx", "pairwise", 8, 4, 30, 2, 17, "s,o", {"margin": 2.0}, {}),
    ("complex_multiclass", "ComplEx", "multiclass_nll", 6, 9, 64, 6, 40, "s,o", {}, {}),
    ("hole_multiclass", "HolE", "multiclass_nll", 16, 5, 64, 3, 48, "s,o", {}, {}),
    ("hole_nll", "HolE", "nll", 10, 2, 48, 3, 25, "s,o", {}, {}),
    ("hole_pairwise", "HolE", "pairwise", 6, 3, 48, 3, 25, "o", {"margin": 0.5}, {}),
    # section 8(f).4 plug-ins: losses/absolute_margin.py:54-70, losses/self_adversarial.py:78-112
    ("distmult_absmargin", "DistMult", "absolute_margin", 12, 5, 64, 5, 40, "s,o", {"margin": 0.5}, {}),
    ("transe_l1_absmargin", "TransE", "absolute_margin", 10, 4, 50, 3, 33, "s,o", {"margin": 3.0}, {}),
    ("complex_selfadv", "ComplEx", "self_adversarial", 12, 6, 64, 6, 40, "s,o", {"margin": 3.0, "alpha": 0.5}, {}),
    ("distmult_selfadv", "DistMult", "self_adversarial", 16, 9, 64, 5, 48, "s,o", {"margin": 1.0, "alpha": 1.5}, {}),
    ("transe_l1_selfadv", "TransE", "self_adversarial", 10, 4, 50, 3, 33, "o", {"margin": 2.0, "alpha": 0.5}, {}),
    ("hole_selfadv", "HolE", "self_adversarial", 8, 5, 48, 3, 30, "s,o", {}, {}),
]

# LP regulariser (regularizers/lp.py:81-113): name, base train case fields..., regularizer_params
REG_CASES = [
    ("distmult_nll_l2", "DistMult", "nll", 12, 4, 64, 5, 40, "s,o", {}, {}, {"p": 2, "lambda": 1e-2}),
    ("complex_pairwise_l3", "ComplEx", "pairwise", 8, 4, 48, 4, 30, "s,o", {"margin": 1.0}, {}, {"p": 3, "lambda": [1e-2, 5e-2]}),
    ("transe_l1_pairwise_l1", "TransE", "pairwise", 10, 3, 50, 3, 25, "s,o", {"margin": 2.0}, {}, {"p": 1, "lambda": 1e-3}),
]
# embedding_model_params['non_linearity'] (models/EmbeddingModel.py:679-689, :801-812)
NL_CASES = [
    ("distmult_nll_tanh", "DistMult", "nll", 12, 4, 64, 5, 40, "s,o", {}, {"non_linearity": "tanh"}, None),
    ("complex_pairwise_sigmoid", "ComplEx", "pairwise", 8, 4, 48, 4, 30, "s,o", {"margin": 0.3}, {"non_linearity": "sigmoid"}, None),
    ("transe_l2_multiclass_softplus", "TransE", "multiclass_nll", 10, 5, 50, 3, 25, "s,o", {}, {"non_linearity": "softplus", "norm": 2}, None),
    ("hole_selfadv_tanh", "HolE", "self_adversarial", 8, 4, 48, 3, 30, "s,o", {"margin": 0.5}, {"non_linearity": "tanh"}, None),
]
TRAIN_CASES = [c + (None,) for c in TRAIN_CASES] + REG_CASES + NL_CASES

# a LIST-valued corrupt_side: one loss term per side, summed before the single step (models/EmbeddingModel.py:780-816)
# name, model, loss, k, eta, E, R, n, sides, loss_params, emb_params, regularizer_params
MULTISIDE_CASES = [
    ("complex_nll_s_o", "ComplEx", "nll", 8, 4, 48, 4, 30, ["s", "o"], {}, {}, None),
    ("distmult_multiclass_so_s", "DistMult", "multiclass_nll", 12, 5, 64, 5, 40, ["s,o", "s"], {}, {}, None),
    ("transe_l1_pairwise_o_s_so", "TransE", "pairwise", 10, 3, 50, 3, 25, ["o", "s", "s+o"], {"margin": 2.0}, {}, None),
    ("hole_selfadv_s_o_l2", "HolE", "self_adversarial", 8, 4, 48, 3, 30, ["s", "o"], {"margin": 1.0, "alpha": 0.5}, {},
     {"p": 2, "lambda": 1e-2}),
]

# Wide rows on a larger table (the vectorised multi-chunk kernel paths, K >= 256, E = 5000), executed by the reference's own
# code.  Tables come from np.random.RandomState(seed) (a frozen stream) and the gradients are stored for the touched rows
# only, so the fixtures stay small: name, model, loss, k, eta, E, R, n, emb_params
WIDE_CASES = [
    ("transe_l1_pairwise_k256", "TransE", "pairwise", 256, 6, 5000, 9, 96, {}),
    ("distmult_nll_k256", "DistMult", "nll", 256, 8, 5000, 9, 96, {}),
    ("complex_nll_k200", "ComplEx", "nll", 200, 6, 5000, 9, 96, {}),
    ("hole_multiclass_k256", "HolE", "multiclass_nll", 256, 5, 5000, 9, 64, {}),
    ("transe_l2_selfadv_k300", "TransE", "self_adversarial", 300, 4, 5000, 9, 64, {"norm": 2}),
]


def wide_tables(seed, E, R, K):
    """The tables of a wide case, regenerated by the tests from the seed (RandomState streams are frozen)."""
    rs = np.random.RandomState(seed)
    ent = (rs.standard_normal((E, K)) * 0.3).astype(np.float32)
    rel = (rs.standard_normal((R, K)) * 0.3).astype(np.float32)
    return ent, rel


RANK_CASES = [
    # name, model, k, E, R, F, T, emb_params, scale
    ("rank_transe_l1", "TransE", 10, 120, 4, 900, 60, {}, 0.5),
    ("rank_transe_l2", "TransE", 8, 90, 3, 600, 40, {"norm": 2}, 0.5),
    ("rank_distmult", "DistMult", 16, 150, 5, 1200, 60, {}, 0.7),
    ("rank_complex", "ComplEx", 12, 130, 4, 1000, 60, {}, 0.7),
    ("rank_hole", "HolE", 8, 100, 3, 800, 50, {}, 1.0),
    ("rank_distmult_tanh", "DistMult", 12, 110, 4, 800, 40, {"non_linearity": "tanh"}, 0.6),
    ("rank_transe_l1_sigmoid", "TransE", 8, 90, 3, 600, 40, {"non_linearity": "sigmoid"}, 0.4),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    assert ref_shim.available(), "needs /root/reference"
    only = [a for a in sys.argv[1:] if not a.startswith("-")]  # optional name substrings: regenerate a subset

    def wanted(name):
        return not only or any(o in name for o in only)

    for ci, (name, model, loss, k, eta, E, R, n, side, lp, ep, rp) in enumerate(TRAIN_CASES):
        if not wanted("train_" + name):
            continue
        rng = np.random.Generator(np.random.PCG64(1000 + ci))
        K = ko.internal_k(model, k)
        ent = (rng.normal(size=(E, K)) * 0.6).astype(np.float32)
        rel = (rng.normal(size=(R, K)) * 0.6).astype(np.float32)
        pos = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)
        keep = ko.side_mask(side, n * eta, rng)
        repl = rng.integers(0, E, n * eta).astype(np.int32)
        ref = ref_shim.ref_train_forward_backward(model, k, eta, loss, ent, rel, pos, keep, repl, lp, ep, side,
                                                  "LP" if rp else None, rp)
        lam = (rp or {}).get("lambda", 0.0)
        lam = [lam, lam] if np.isscalar(lam) else list(lam)
        np.savez_compressed(
            os.path.join(OUT, "train_%s.npz" % name),
            model=model, loss_name=loss, k=k, eta=eta, side=side,
            margin=float(lp.get("margin", 3.0 if loss == "self_adversarial" else 1.0)), alpha=float(lp.get("alpha", 0.5)),
            reg_p=int((rp or {}).get("p", 0)), reg_lambda_ent=float(lam[0]), reg_lambda_rel=float(lam[1]),
            norm=int(ep.get("norm", 1)), nl=str(ep.get("non_linearity", "linear")), ent=ent, rel=rel, pos=pos, keep_subj=keep, repl=repl,
            loss=np.float32(ref["loss"]), scores_pos=ref["scores_pos"], scores_neg=ref["scores_neg"],
            neg=ref["neg"].astype(np.int32), grad_ent=ref["grad_ent"], grad_rel=ref["grad_rel"])
        print("train", name, "loss", ref["loss"])
    for ci, (name, model, loss, k, eta, E, R, n, sides, lp, ep, rp) in enumerate(MULTISIDE_CASES):
        if not wanted("train_multiside_" + name):
            continue
        rng = np.random.Generator(np.random.PCG64(3000 + ci))
        K = ko.internal_k(model, k)
        ent = (rng.normal(size=(E, K)) * 0.6).astype(np.float32)
        rel = (rng.normal(size=(R, K)) * 0.6).astype(np.float32)
        pos = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)
        keeps = [ko.side_mask(sd, n * eta, rng) for sd in sides]
        repls = [rng.integers(0, E, n * eta).astype(np.int32) for _ in sides]
        ref = ref_shim.ref_train_forward_backward_sides(model, k, eta, loss, ent, rel, pos, sides, keeps, repls, lp, ep,
                                                        "LP" if rp else None, rp)
        # stored in the stacked layout the engine consumes (oracle/kge_oracle.py:stack_sides)
        pos2, keep2, repl2 = ko.stack_sides(pos, eta, sides, keeps, repls)
        S = len(sides)
        sneg = np.stack([x.reshape(eta, n) for x in ref["scores_neg"]], 1).reshape(-1)
        neg2 = np.stack([x.reshape(eta, n, 3) for x in ref["neg"]], 1).reshape(-1, 3)
        lam = (rp or {}).get("lambda", 0.0)
        lam = [lam, lam] if np.isscalar(lam) else list(lam)
        np.savez_compressed(
            os.path.join(OUT, "train_multiside_%s.npz" % name),
            model=model, loss_name=loss, k=k, eta=eta, side="|".join(sides), n_sides=S,
            margin=float(lp.get("margin", 3.0 if loss == "self_adversarial" else 1.0)), alpha=float(lp.get("alpha", 0.5)),
            reg_p=int((rp or {}).get("p", 0)), reg_lambda_ent=float(lam[0]), reg_lambda_rel=float(lam[1]),
            norm=int(ep.get("norm", 1)), nl=str(ep.get("non_linearity", "linear")), ent=ent, rel=rel, pos=pos2, pos_single=pos,
            keep_subj=keep2, repl=repl2, loss=np.float32(ref["loss"]), scores_pos=np.tile(ref["scores_pos"], S), scores_neg=sneg,
            neg=neg2.astype(np.int32), grad_ent=ref["grad_ent"], grad_rel=ref["grad_rel"])
        print("train multiside", name, "loss", ref["loss"])
    for ci, (name, model, loss, k, eta, E, R, n, ep) in enumerate(WIDE_CASES):
        if not wanted("wide_" + name):
            continue
        seed = 5000 + ci
        K = ko.internal_k(model, k)
        ent, rel = wide_tables(seed, E, R, K)
        rng = np.random.Generator(np.random.PCG64(seed))
        pos = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)
        pos[: n // 4, 0] = pos[0, 0]  # a hub entity
        keep = ko.side_mask("s,o", n * eta, rng)
        repl = rng.integers(0, E, n * eta).astype(np.int32)
        ref = ref_shim.ref_train_forward_backward(model, k, eta, loss, ent, rel, pos, keep, repl, {}, ep, "s,o", None, None)
        rows = np.flatnonzero(np.abs(ref["grad_ent"]).max(1) > 0).astype(np.int32)
        touched = np.unique(np.concatenate([pos[:, 0], pos[:, 2], repl]))
        assert np.all(np.isin(rows, touched))
        np.savez_compressed(
            os.path.join(OUT, "wide_%s.npz" % name), model=model, loss_name=loss, k=k, eta=eta, E=E, R=R, table_seed=seed,
            margin=3.0 if loss == "self_adversarial" else 1.0, alpha=0.5, norm=int(ep.get("norm", 1)), pos=pos, keep_subj=keep, repl=repl,
            loss=np.float32(ref["loss"]), scores_pos=ref["scores_pos"], scores_neg=ref["scores_neg"],
            grad_rows=touched.astype(np.int32), grad_ent_rows=ref["grad_ent"][touched], grad_rel=ref["grad_rel"])
        print("wide", name, "loss", ref["loss"], "touched rows", touched.size)
    if wanted("split_cases"):
        # train_test_split_no_unseen (evaluation/protocol.py:24-407), both algorithms, executed by the reference
        rng = np.random.Generator(np.random.PCG64(4000))
        out, ci = {}, 0
        for trial in range(12):
            E, R, n = int(rng.integers(5, 40)), int(rng.integers(1, 5)), int(rng.integers(10, 200))
            X = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(str)
            X[:, 1] = np.char.add("r", X[:, 1])
            for bc in (False, True):
                for dup, filt in ((False, None), (True, None), (False, ["r0"])):
                    ts, seed = int(rng.integers(1, max(2, n // 3))), int(rng.integers(0, 1000))
                    try:
                        a, b = ref_shim.load().protocol.train_test_split_no_unseen(
                            X, test_size=ts, seed=seed, allow_duplication=dup, filtered_test_predicates=filt, backward_compatible=bc)
                        err = ""
                    except Exception as e:  # noqa: BLE001
                        a = b = np.zeros((0, 3), dtype=X.dtype)
                        err = str(e)
                    out.update({"X_%d" % ci: X, "par_%d" % ci: np.array([ts, seed, int(dup), int(bc), int(filt is not None)]),
                                "train_%d" % ci: a, "test_%d" % ci: b, "err_%d" % ci: np.array(err)})
                    ci += 1
        out["n_cases"] = np.array(ci)
        np.savez_compressed(os.path.join(OUT, "split_cases.npz"), **out)
        print("split cases", ci, "of which raising", sum(1 for i in range(ci) if str(out["err_%d" % i])))
    for ci, (name, model, k, E, R, F, T, ep, scale) in enumerate(RANK_CASES):
        if not wanted(name):
            continue
        rng = np.random.Generator(np.random.PCG64(2000 + ci))
        K = ko.internal_k(model, k)
        ent = (rng.normal(size=(E, K)) * scale).astype(np.float32)
        rel = (rng.normal(size=(R, K)) * scale).astype(np.float32)
        filt = ko.synthetic_triples(E, R, F, seed=300 + ci)
        # duplicates in the filter must be harmless (SQL UNION / DISTINCT)
        filt = np.concatenate([filt, filt[:17]], axis=0)
        test = filt[rng.permutation(F)[:T]].copy()
        # a few test triples that are NOT in the filter: self must still be filtered
        test[:5, 2] = (test[:5, 2] + 7) % E
        out = dict(model=model, k=k, norm=int(ep.get("norm", 1)), nl=str(ep.get("non_linearity", "linear")), ent=ent, rel=rel, filt=filt,
                   test=test)
        for side in ("s,o", "s+o", "s", "o"):
            for strat in ("worst", "best", "middle"):
                for fl in (0, 1):
                    r = ref_shim.ref_ranks(model, k, ent, rel, test, filt if fl else None, side, strat, ep)
                    out["ranks_%s_%s_%d" % (side.replace(",", "c").replace("+", "p"), strat, fl)] = r.astype(np.int32)
        np.savez_compressed(os.path.join(OUT, "%s.npz" % name), **out)
        print("rank", name, out["ranks_sco_worst_1"][:4].tolist())


if __name__ == "__main__":
    main()
