"""CPU: the host logic of fit() / predict() end to end, with the engine replaced by tests/fake_engine.py (every step
computed by the oracle on CPU tensors).  What is checked is the Python around the kernels -- batch slicing, step / seed
counters, flags, side stacking, loss bookkeeping, re-fit seeding, the sgd schedule plumbing -- against the oracle's
restatement of the loop (oracle/kge_oracle.py:fit_emulation).  The kernels themselves are checked on the GPU."""
import numpy as np
import pytest
import torch

from emgraph_b200 import _lib, models
from fake_engine import FakeEngine
from oracle import kge_oracle as ko
from toy_graph import TOY_CASES, TOY_QUERY, TOY_X, toy_fit_emulation


@pytest.fixture
def fake(monkeypatch):
    eng = FakeEngine()
    monkeypatch.setattr(models, "get_engine", lambda device=None: eng)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    return eng


def _toy_model(model, bc, margin, reg, **kw):
    extra = dict(regularizer="LP", regularizer_params={"lambda": reg[0], "p": reg[1]}) if reg else {}
    return getattr(models, model)(batches_count=bc, seed=555, epochs=20, k=10, loss="pairwise", loss_params={"margin": margin},
                                  optimizer="adagrad", optimizer_params={"lr": 0.1}, **extra, **kw)


@pytest.mark.parametrize("host_batches,host_pipeline", [(False, True), (True, True), (True, False)])
@pytest.mark.parametrize("model,bc,margin,reg", TOY_CASES)
def test_fit_loop_equals_the_emulation(fake, model, bc, margin, reg, host_batches, host_pipeline):
    m = _toy_model(model, bc, margin, reg, engine_params={"host_batches": host_batches, "host_pipeline": host_pipeline})
    m.fit(TOY_X)
    y_ref, losses_ref = toy_fit_emulation(model, bc, margin, reg)
    np.testing.assert_array_equal(m.predict(TOY_QUERY), y_ref)
    np.testing.assert_allclose(m.loss_history, np.asarray(losses_ref) / 16.0, rtol=1e-6)
    assert [c["step"] for c in fake.calls] == list(range(1, 20 * bc + 1)) and all(c["seed"] == 555 for c in fake.calls)
    assert all(c["n"] == 8 // bc for c in fake.calls) and m._opt_step == 20 * bc
    # batches are resident slices / library-copied host buffers: every step may be software-pipelined
    assert all(c["flags"] & _lib.F_PIPELINE for c in fake.calls)
    # a second fit starts from the seed again (models/EmbeddingModel.py:1285-1290)
    m.fit(TOY_X)
    np.testing.assert_array_equal(m.predict(TOY_QUERY), y_ref)


def test_engine_params_switch_pipelining_off(fake):
    m = _toy_model("DistMult", 2, 5.0, None, engine_params={"pipeline": False})
    m.fit(TOY_X)
    assert not any(c["flags"] & _lib.F_PIPELINE for c in fake.calls)


def test_side_list_is_one_stacked_step_per_batch(fake):
    """corrupt_side=['s', 'o', 's,o'] -> one step per batch over the positives stacked three times, per-negative side
    codes, in-order step (the stacked batch is made on torch's stream); equals the oracle on the stacked batch."""
    r2i, e2i = ko.create_mappings(TOY_X)
    Xi = ko.to_idx(TOY_X, e2i, r2i)
    E, R, k, eta = len(e2i), len(r2i), 6, 3
    m = models.ComplEx(k=k, eta=eta, epochs=3, batches_count=2, seed=9, loss="multiclass_nll", optimizer="adam", optimizer_params={"lr": 0.05},
                       embedding_model_params={"corrupt_side": ["s", "o", "s,o"]})
    m.fit(TOY_X)
    assert len(fake.calls) == 6 and all(c["n"] == 12 and c["keep_codes"] and not (c["flags"] & _lib.F_PIPELINE) for c in fake.calls)
    np.testing.assert_array_equal(fake.calls[0]["pos"], np.tile(Xi[:4], (3, 1)))
    # the same loop on the oracle
    K = 2 * k
    rnd = np.random.RandomState(9)
    ent = rnd.uniform(-np.sqrt(6 / (E + K)), np.sqrt(6 / (E + K)), size=(E, K)).astype(np.float32)
    rel = rnd.uniform(-np.sqrt(6 / (R + K)), np.sqrt(6 / (R + K)), size=(R, K)).astype(np.float32)
    state = ((np.zeros_like(ent), np.zeros_like(ent)), (np.zeros_like(rel), np.zeros_like(rel)))
    step = 0
    for _ in range(3):
        for b in range(2):
            step += 1
            pos = np.tile(Xi[b * 4:(b + 1) * 4], (3, 1))
            codes = np.tile(np.repeat(np.array([0, 1, 2], np.uint8), 4), eta)
            repl, keep = ko.draw_corruptions(9, step, 12, eta, E, "s,o", keep_codes=codes)
            o = ko.train_step("ComplEx", k, "multiclass_nll", eta, ent, rel, pos, keep, repl, opt="adam", lr=0.05, state=state, step=step)
            ent, rel, state = o["ent_new"], o["rel_new"], (o["state_ent"], o["state_rel"])
    np.testing.assert_array_equal(m.trained_model_params[0], ent)
    np.testing.assert_array_equal(m.trained_model_params[1], rel)


def test_batch_corruption_entities_keep_the_in_order_step(fake):
    m = models.DistMult(k=4, eta=2, epochs=1, batches_count=2, seed=1, embedding_model_params={"negative_corruption_entities": "batch"})
    m.fit(TOY_X)
    assert all(c["neg_entities"] and not (c["flags"] & _lib.F_PIPELINE) for c in fake.calls)


def test_sgd_schedule_reaches_every_step(fake):
    from emgraph_b200.optimizers import SGDSchedule
    params = {"lr": 0.1, "decay_cycle": 3, "end_lr": 0.01, "cosine_decay": True}
    m = models.TransE(k=4, eta=2, epochs=2, batches_count=4, seed=1, optimizer="sgd", optimizer_params=params, loss="nll")
    m.fit(TOY_X)
    sched = SGDSchedule(params, 4)
    want = [float(sched(b + 1, e)) for e in (1, 2) for b in range(4)]
    np.testing.assert_allclose([c["lr"] for c in fake.calls], want, rtol=0, atol=0)


def test_divergence_raises_like_the_reference(fake):
    """models/EmbeddingModel.py:1422-1427: a NaN / Inf batch loss aborts the fit with ValueError."""
    m = models.DistMult(k=4, eta=2, epochs=3, batches_count=1, seed=0, optimizer="sgd", optimizer_params={"lr": 1e18}, loss="pairwise")
    with np.errstate(all="ignore"), pytest.raises(ValueError, match="Please change the hyperparameters"):
        m.fit(TOY_X)


def test_normalize_ent_emb_after_every_batch(fake):
    m = models.TransE(k=4, eta=2, epochs=2, batches_count=2, seed=0, optimizer="sgd", optimizer_params={"lr": 5.0}, loss="pairwise",
                      embedding_model_params={"normalize_ent_emb": True})
    m.fit(TOY_X)
    assert np.all(np.linalg.norm(m.trained_model_params[0], axis=1) <= 1.0 + 1e-6)


def _synthetic(E=40, R=3, n=500, seed=5):
    tri = ko.synthetic_triples(E, R, n, seed=seed)
    X = np.empty(tri.shape, dtype=object)
    X[:, 0] = ["e%03d" % v for v in tri[:, 0]]
    X[:, 1] = ["r%d" % v for v in tri[:, 1]]
    X[:, 2] = ["e%03d" % v for v in tri[:, 2]]
    return tri, X.astype(str)


def test_evaluate_performance_host_path(fake):
    """evaluate_performance -> set_filter_for_eval / configure_evaluation_protocol / get_ranks / end_evaluation
    (evaluation/protocol.py:726-979): label mapping, unseen-entity filtering, output shapes, filter hand-over."""
    from emgraph_b200.evaluation import evaluate_performance, hits_at_n_score, mrr_score
    tri, X = _synthetic()
    m = models.DistMult(k=6, eta=2, epochs=5, batches_count=4, seed=0, optimizer="adam", optimizer_params={"lr": 0.05})
    m.fit(X[:400])
    test = np.concatenate([X[400:430], np.array([["never-seen", "r0", "e001"]])])
    ent, rel = m.trained_model_params
    seen = tri[400:430]
    for side in ("s,o", "s+o", "o"):
        got = evaluate_performance(test, m, filter_triples=X, corrupt_side=side)
        exp = ko.ranks("DistMult", 6, ent, rel, seen, tri, side, "worst")
        np.testing.assert_array_equal(got, exp)  # the unseen triple was dropped, ids follow the sorted label order
    raw = evaluate_performance(test, m, corrupt_side="s,o", ranking_strategy="best")
    np.testing.assert_array_equal(raw, ko.ranks("DistMult", 6, ent, rel, seen, None, "s,o", "best"))
    assert 0 < mrr_score(raw) <= 1 and 0 <= hits_at_n_score(raw, 10) <= 1
    assert m.is_filtered is False and m.eval_config == {}  # end_evaluation cleaned up
    with pytest.raises(AssertionError):
        evaluate_performance(test, m, corrupt_side="x")


def test_early_stopping_host_logic(fake):
    """models/EmbeddingModel.py:824-1020: validation every check_interval epochs after burn_in, stop after
    stop_interval checks without improvement, return with the best parameters."""
    tri, X = _synthetic(seed=6)
    m = models.ComplEx(k=4, eta=2, epochs=40, batches_count=2, seed=0, optimizer="adam", optimizer_params={"lr": 0.5}, loss="nll")
    m.fit(X[:400], early_stopping=True,
          early_stopping_params={"x_valid": X[400:440], "criteria": "mrr", "burn_in": 2, "check_interval": 2, "stop_interval": 2,
                                 "x_filter": X})
    hist = m.early_stopping_history
    assert [e for e, _ in hist] == list(range(2, 2 * len(hist) + 1, 2)) and fake.rank_calls == len(hist)
    if m.early_stopping_epoch is not None:  # stopped early: the last two checks did not beat the best one
        best = max(v for _, v in hist)
        assert all(v <= best for _, v in hist[-2:]) and len(m.loss_history) == m.early_stopping_epoch < 40
    with pytest.raises(KeyError):
        models.ComplEx(k=4, epochs=1, batches_count=1).fit(X[:50], early_stopping=True, early_stopping_params={})


def test_select_best_model_ranking_end_to_end(fake):
    from emgraph_b200.evaluation import select_best_model_ranking
    _, X = _synthetic(seed=7)
    grid = {"batches_count": [2], "seed": 0, "epochs": [8], "k": [4, 8], "eta": [2], "loss": ["nll"], "loss_params": {},
            "embedding_model_params": {}, "regularizer": [None], "regularizer_params": {}, "optimizer": ["adam"],
            "optimizer_params": {"lr": [1e-9, 5e-2]}}
    best, params, mrr_valid, ranks_test, res, hist = select_best_model_ranking(models.DistMult, X[:400], X[400:450], X[450:], grid,
                                                                              retrain_best_model=True)
    assert len(hist) == 4 and params["optimizer_params"]["lr"] in (1e-9, 5e-2) and best.is_fitted
    assert {(h["model_params"]["k"], h["model_params"]["optimizer_params"]["lr"]) for h in hist} == {(4, 1e-9), (4, 5e-2), (8, 1e-9), (8, 5e-2)}
    assert mrr_valid == max(h["results"]["mrr"] for h in hist) and ranks_test.shape[1] == 2 and np.isfinite(res["mrr"])


@pytest.mark.parametrize("opt", ["adam", "adagrad", "momentum"])
def test_resume_continues_from_saved_optimizer_state(fake, tmp_path, opt):
    """engine_params['resume'] (checkpoint + optimizer-state resume): 2 epochs, save with the optimizer
    state, restore, 2 more epochs == 4 epochs in one go, bit for bit (parameters, per-row state and the global step
    -- which also keys the corruption stream -- carry over)."""
    from emgraph_b200 import restore_model, save_model
    _, X = _synthetic(E=30, n=240, seed=8)
    kw = dict(k=4, eta=3, batches_count=3, seed=4, optimizer=opt, optimizer_params={"lr": 0.05}, loss="nll")
    straight = models.DistMult(epochs=4, **kw)
    straight.fit(X)
    first = models.DistMult(epochs=2, **kw)
    first.fit(X)
    path = save_model(first, str(tmp_path / "m.pkl"), save_optimizer_state=True)
    second = restore_model(path)
    assert second._opt_step == 6
    second.engine_params["resume"] = True
    second.fit(X)
    assert second._opt_step == 12
    np.testing.assert_array_equal(second.trained_model_params[0], straight.trained_model_params[0])
    np.testing.assert_array_equal(second.trained_model_params[1], straight.trained_model_params[1])
    np.testing.assert_allclose(second.loss_history, straight.loss_history[2:], rtol=1e-6)
    # a different entity set cannot be resumed
    with pytest.raises(ValueError, match="resume needs the entities"):
        second.fit(X[(X[:, 0] != "e000") & (X[:, 2] != "e000")])
    # without the flag a re-fit starts over from the seed
    second.engine_params["resume"] = False
    second.fit(X)
    np.testing.assert_array_equal(second.trained_model_params[0], first.trained_model_params[0])


def test_entities_subset_host_logic(fake):
    """evaluate_performance(entities_subset=...) (evaluation/protocol.py:940-944, models/EmbeddingModel.py:1845-1857,
    :1898-1940): the host re-labels the entities so that the subset occupies the first rows of a permuted table,
    re-labels test and filter triples alike, sweeps those rows only and tells rank_finalize whether the test triple's
    own entity was a candidate.  Checked against the oracle's subset ranking; the stand-in engine implements the
    documented contract of kge_rank_counts / kge_rank_finalize."""
    from emgraph_b200.evaluation import evaluate_performance
    rng = np.random.default_rng(31)
    E, R, k = 60, 3, 5
    ent = (rng.normal(size=(E, k)) * 0.6).astype(np.float32)
    rel = (rng.normal(size=(R, k)) * 0.6).astype(np.float32)
    tri = ko.synthetic_triples(E, R, 400, seed=8)
    X = np.empty(tri.shape, dtype=object)
    X[:, 0] = ["e%05d" % v for v in tri[:, 0]]
    X[:, 1] = ["r%03d" % v for v in tri[:, 1]]
    X[:, 2] = ["e%05d" % v for v in tri[:, 2]]
    X = X.astype(str)
    m = models.DistMult(k=k, eta=2, epochs=1, batches_count=1, seed=0, optimizer="sgd", optimizer_params={"lr": 0.0}, loss="nll",
                        initializer="constant", initializer_params={"entity": ent, "relation": rel})
    m.fit(X)
    sel = rng.permutation(len(tri))[:25]
    subset_ids = np.sort(rng.permutation(E)[:17])
    subset_labels = ["e%05d" % v for v in subset_ids] + ["not-an-entity"]
    for side in ("s,o", "s+o", "o", "s"):
        for strat in ("worst", "middle", "best"):
            for filt in (None, tri):
                got = evaluate_performance(X[sel], m, filter_triples=None if filt is None else X, entities_subset=subset_labels,
                                           corrupt_side=side, ranking_strategy=strat)
                exp = ko.ranks("DistMult", k, ent, rel, tri[sel], filt, side, strat, subset=subset_ids)
                np.testing.assert_array_equal(got, exp, err_msg="%s %s %s" % (side, strat, filt is not None))
    # the full sweep afterwards uses the un-permuted filter again
    np.testing.assert_array_equal(evaluate_performance(X[sel], m, filter_triples=X), ko.ranks("DistMult", k, ent, rel, tri[sel], tri))


def test_gpu_tier_property_tests_also_hold_on_the_stand_in_engine(fake, tmp_path):
    """The bodies of the GPU-tier end-to-end tests (toy-graph fit/predict vs the emulation, re-fit determinism, resume)
    are host logic over the engine: run them here too, so a Python-level mistake in them shows up in the CPU tier."""
    import test_zzz_gpu_reference_properties as gpu_tests
    for case in TOY_CASES:
        gpu_tests.test_fit_predict_toy_graph(None, *case)
    gpu_tests.test_refit_is_deterministic(None)
    gpu_tests.test_resume_equals_training_in_one_go(None, tmp_path)
