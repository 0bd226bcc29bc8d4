"""CPU: the parameter-grid helpers and the control flow of select_best_model_ranking
(reference evaluation/protocol.py:1044-1703).  The cases follow the reference's own tests
(tests/emgraph/evaluation/test_protocol.py:679-1043); the expected values are the reference's."""
from itertools import islice

import numpy as np

from emgraph_b200 import model_selection as ms

GRID = {
    "batches_count": [50], "epochs": [4000], "k": [100, 200], "eta": [5, 10, 15], "loss": ["pairwise", "nll"],
    "loss_params": {"margin": [2]}, "embedding_model_params": {}, "regularizer": ["LP", None],
    "regularizer_params": {"p": [1, 3], "lambda": [1e-4, 1e-5]}, "optimizer": ["adagrad", "adam"],
    "optimizer_params": {"lr": [0.01, 0.001, 0.0001]}, "verbose": [False], "model_name": ["ComplEx"],
}


def _grid():
    return {k: (dict(v) if isinstance(v, dict) else v) for k, v in GRID.items()}


def test_remove_unused_params():  # reference test_protocol.py:679-721
    p1 = {"batches_count": 50, "epochs": 4000, "k": 200, "eta": 15, "loss": "nll", "loss_params": {"margin": 2},
          "embedding_model_params": {}, "regularizer": "LP", "regularizer_params": {"p": 1, "lambda": 1e-5},
          "optimizer": "adam", "optimizer_params": {"lr": 0.001}, "verbose": False, "model_name": "ComplEx"}
    out = ms._remove_unused_params(p1)
    assert out["loss_params"] == {} and out["embedding_model_params"] == {}
    assert out["regularizer_params"] == {"p": 1, "lambda": 1e-5} and out["optimizer_params"] == {"lr": 0.001}
    assert p1["loss_params"] == {"margin": 2}  # input untouched
    p2 = dict(p1, loss="self_adversarial", regularizer=None, model_name="unknown_model")
    del p2["embedding_model_params"]
    out = ms._remove_unused_params(p2)
    assert out["loss_params"] == {"margin": 2} and out["regularizer_params"] == {} and out["optimizer_params"] == {"lr": 0.001}


def test_flatten_unflatten_are_inverse():  # reference test_protocol.py:724-796
    nested = {"k": 5, "loss_params": {"margin": 2, "alpha": 1}, "optimizer_params": {"lr": 0.1}, "embedding_model_params": {}}
    flat = ms._flatten_nested_keys(nested)
    assert flat == {"k": 5, ("loss_params", "margin"): 2, ("loss_params", "alpha"): 1, ("optimizer_params", "lr"): 0.1}
    back = ms._unflatten_nested_keys(flat)
    assert back == {"k": 5, "loss_params": {"margin": 2, "alpha": 1}, "optimizer_params": {"lr": 0.1}}
    assert ms._flatten_nested_keys(back) == flat


def test_param_hash_ignores_unused_parameters():  # reference test_protocol.py:799-911
    a = {"loss": "nll", "loss_params": {"margin": 2}, "regularizer": None, "regularizer_params": {"p": 1}, "k": 10}
    b = {"loss": "nll", "loss_params": {"margin": 7}, "regularizer": None, "regularizer_params": {"p": 3}, "k": 10}
    c = dict(a, k=11)
    assert ms._get_param_hash(a) == ms._get_param_hash(b) != ms._get_param_hash(c)
    assert ms._get_param_hash(ms._flatten_nested_keys(a)) == ms._get_param_hash(a)
    h = ms.ParamHistory()
    h.add(a)
    assert b in h and c not in h
    # list-valued parameters (corrupt_side lists) hash too
    d = {"model_name": "ComplEx", "embedding_model_params": {"corrupt_side": ["s", "o"]}}
    assert ms._get_param_hash(d) == ms._get_param_hash({"model_name": "ComplEx", "embedding_model_params": {"corrupt_side": ["s", "o"]}})


def test_next_hyperparam_counts():  # reference test_protocol.py:952-977: 360 distinct configurations
    combos = list(ms._next_hyperparam(_grid()))
    assert len(combos) == 360
    assert len(set(frozenset(ms._flatten_nested_keys(c).items()) for c in combos)) == 360
    assert all(type(c) is dict and all(type(k) is str for k in c) for c in combos)


def test_next_hyperparam_random_is_distinct():  # reference test_protocol.py:980-1005
    np.random.seed(0)
    combos = list(islice(ms._next_hyperparam_random(_grid()), 200))
    assert len(set(frozenset(ms._flatten_nested_keys(c).items()) for c in combos)) == 200


def test_sample_parameters_and_scalars_into_lists():  # reference test_protocol.py:914-949, :1008-1043
    np.random.seed(0)
    g = _grid()
    g["eta"] = lambda: np.random.choice([5, 10, 15])
    g["optimizer_params"] = {"lr": lambda: np.random.uniform(0.001, 0.1)}
    g["verbose"], g["model_name"] = False, "ComplEx"
    for _ in range(10):
        p = ms._sample_parameters(g)
        assert p["batches_count"] == 50 and p["k"] in (100, 200) and p["eta"] in (5, 10, 15)
        assert p["regularizer"] in ("LP", None) and p["optimizer"] in ("adagrad", "adam")
        assert 0.001 < p["optimizer_params"]["lr"] < 0.1 and p["model_name"] == "ComplEx" and not p["verbose"]
    eta_fn = g["eta"]
    grid = {"batches_count": 50, "epochs": [4000], "eta": eta_fn, "loss": "nll", "loss_params": {"margin": 2},
            "embedding_model_params": {}, "regularizer": ["LP", None], "optimizer_params": {"lr": "wrong"}, "verbose": False}
    ms._scalars_into_lists(grid)
    assert grid == {"batches_count": [50], "epochs": [4000], "eta": eta_fn, "loss": ["nll"], "loss_params": {"margin": [2]},
                    "embedding_model_params": {}, "regularizer": ["LP", None], "optimizer_params": {"lr": ["wrong"]},
                    "verbose": [False]}


class _FakeModel:
    """Stands in for a fitted EmbeddingModel: MRR is a known function of the parameters."""
    name = "TransE"
    fits = []

    def __init__(self, k=1, optimizer_params=None, **kw):
        if k == 13:
            raise ValueError("unlucky k")
        self.k, self.lr, self.kw = k, (optimizer_params or {}).get("lr", 0.0), kw

    def fit(self, X, early_stopping=False, early_stopping_params=None):
        _FakeModel.fits.append((X.shape[0], early_stopping, dict(early_stopping_params or {})))


def test_select_best_model_ranking_control_flow(monkeypatch):
    """Grid + random search, exception bookkeeping, filter assembly, retraining (evaluation/protocol.py:1517-1703)."""
    from emgraph_b200 import evaluation as ev
    calls = []

    def fake_eval(X, model, filter_triples=None, **kw):
        calls.append((X.shape[0], None if filter_triples is None else filter_triples.shape[0], kw["corrupt_side"]))
        best = model.k == 4 and model.lr == 0.1
        return np.full((X.shape[0], 2), 1 if best else 10)

    monkeypatch.setattr(ev, "evaluate_performance", fake_eval)
    Xtr, Xva, Xte = np.zeros((30, 3), dtype="<U4"), np.zeros((7, 3), dtype="<U4"), np.zeros((5, 3), dtype="<U4")
    _FakeModel.fits = []
    grid = {"k": [2, 4, 13], "seed": 0, "optimizer": "adam", "optimizer_params": {"lr": [0.1, 0.2]}, "loss": "nll",
            "loss_params": {"margin": [1, 2]}}
    best, params, mrr, ranks, res, hist = ms.select_best_model_ranking(_FakeModel, Xtr, Xva, Xte, grid, retrain_best_model=True,
                                                                      early_stopping=True)
    assert best.k == 4 and params["optimizer_params"] == {"lr": 0.1} and mrr == 1.0
    assert len(hist) == 6  # 3 k x 2 lr; the two margins collapse (nll reads no margin)
    assert sum("exception" in h["results"] for h in hist) == 2 and all(h["model_name"] == "TransE" for h in hist)
    assert res == {"mrr": 1.0, "mr": 1.0, "hits_1": 1.0, "hits_3": 1.0, "hits_10": 1.0} and ranks.shape == (5, 2)
    # selection on X_valid with the train+valid+test filter, the final evaluation on X_test
    assert calls[0] == (7, 42, "s,o") and calls[-1] == (5, 42, "s,o")
    assert _FakeModel.fits[-1][0] == 37 and _FakeModel.fits[0][1] is True and _FakeModel.fits[0][2]["x_valid"] is Xva
    # random search: exactly max_combinations distinct configurations, callables sampled
    calls.clear()
    grid = {"k": [2, 4], "optimizer_params": {"lr": lambda: float(np.random.uniform(0.3, 0.4))}}
    out = ms.select_best_model_ranking(_FakeModel, Xtr, Xva, Xte, grid, max_combinations=5, use_filter=False,
                                       use_test_for_selection=True, corrupt_side="o")
    assert len(out[5]) == 5 and all(0.3 <= h["model_params"]["optimizer_params"]["lr"] <= 0.4 for h in out[5])
    assert calls[0] == (5, None, "o")
    # nothing trainable -> NaN summary, no model
    out = ms.select_best_model_ranking(_FakeModel, Xtr, Xva, Xte, {"k": [13]})
    assert out[0] is None and out[3] == [] and np.isnan(out[4]["mrr"]) and len(out[5]) == 1
