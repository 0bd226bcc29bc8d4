"""CPU: the reference's OWN test cases for the path (tests/emgraph/evaluation/test_protocol.py, tests/emgraph/models/test_models.py,
tests/emgraph/utils/test_model_utils.py), restated against emgraph_b200 with the reference's constructor arguments and assertions.
The reference loads WN18 / WN18RR / YAGO3-10 from the network; here the same calls run on a seeded synthetic graph with labels,
split with train_test_split_no_unseen, and the engine is tests/fake_engine.py (every step and every rank computed by the oracle),
so what is pinned is the host side of the drop-in: argument handling, warnings, errors, shapes, protocol identities.  The same
kernels' numbers are pinned in the GPU tier.  Every test names the reference test it restates (file:line)."""
import os
import warnings
from collections import namedtuple

import numpy as np
import pytest
import torch

from emgraph_b200 import evaluation, models, restore_model, save_model
from emgraph_b200.evaluation import (evaluate_performance, filter_unseen_entities, generate_corruptions_for_eval,
                                     generate_corruptions_for_fit, hits_at_n_score, mr_score, mrr_score, train_test_split_no_unseen)
from emgraph_b200.models import ComplEx, DistMult, TransE, create_mappings, reset_entity_threshold, set_entity_threshold, to_idx
from fake_engine import FakeEngine
from oracle import kge_oracle as ko


@pytest.fixture
def fake(monkeypatch):
    eng = FakeEngine()
    monkeypatch.setattr(models, "get_engine", lambda device=None: eng)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    return eng


def _graph(E=90, R=4, n=1400, seed=3):
    """A labelled stand-in for BaseDataset.load_dataset(...): {'train', 'valid', 'test'} with no unseen entity outside train."""
    tri = np.unique(ko.synthetic_triples(E, R, n, seed=seed, zipf=True), axis=0)
    X = np.empty(tri.shape, dtype=object)
    X[:, 0] = ["e%03d" % v for v in tri[:, 0]]
    X[:, 1] = ["r%d" % v for v in tri[:, 1]]
    X[:, 2] = ["e%03d" % v for v in tri[:, 2]]
    X = X.astype(str)
    rest, test = train_test_split_no_unseen(X, test_size=60, seed=0)
    train, valid = train_test_split_no_unseen(rest, test_size=60, seed=1)
    return {"train": train, "valid": valid, "test": test}


TOY = np.array([["a", "y", "b"], ["b", "y", "a"], ["a", "y", "c"], ["c", "y", "a"], ["a", "y", "d"], ["c", "y", "d"], ["b", "y", "c"],
                ["f", "y", "e"]])
TOY5 = np.array([["a", "y", "b"], ["b", "y", "a"], ["a", "y", "c"], ["c", "z", "a"], ["a", "z", "d"]])


# ------------------------------------------------------------------ tests/emgraph/evaluation/test_protocol.py
def test_evaluate_performance_too_many_entities_warning(fake, monkeypatch):
    """test_protocol.py:35-77 (issue #186): a warning when the corruption pool is at or above TOO_MANY_ENTITIES_TH -- without an
    entity list and with a long one -- and none for a short list or a small graph.  The threshold (50 000) is lowered to the
    synthetic graph's scale."""
    X = _graph()
    model = TransE(batches_count=20, seed=0, epochs=1, k=5, eta=1, verbose=True)
    model.fit(X["train"])
    n_ent = len(model.ent_to_idx)
    monkeypatch.setattr(evaluation, "TOO_MANY_ENTITIES_TH", n_ent - 10)
    with pytest.warns(UserWarning):
        evaluate_performance(X["test"][::10], model, verbose=True, corrupt_side="o")
    entities_subset = np.union1d(np.unique(X["train"][:, 0]), np.unique(X["train"][:, 2]))[:n_ent - 10]
    with pytest.warns(UserWarning):
        evaluate_performance(X["test"][::10], model, verbose=True, corrupt_side="o", entities_subset=entities_subset)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        evaluate_performance(X["test"][::10], model, verbose=True, corrupt_side="o", entities_subset=entities_subset[:10])
        monkeypatch.setattr(evaluation, "TOO_MANY_ENTITIES_TH", 50000)
        evaluate_performance(X["test"][::10], model, verbose=True, corrupt_side="o")


def _complex_nll(**kw):
    return ComplEx(batches_count=10, seed=0, epochs=1, k=20, eta=10, loss="nll", optimizer="adam", optimizer_params={"lr": 0.01},
                   verbose=True, **kw)


def test_evaluate_performance_filter_without_xtest(fake):
    """test_protocol.py:80-104: the filter does not contain the test triples."""
    X = _graph()
    model = _complex_nll(regularizer=None)
    model.fit(X["train"])
    X_filter = np.concatenate((X["train"], X["valid"]))
    ranks = evaluate_performance(X["test"][::5], model, X_filter, verbose=True, corrupt_side="s,o")
    assert ranks.shape == (len(X["test"][::5]), 2) and mrr_score(ranks) > 0


def test_evaluate_performance_ranking_against_specified_entities(fake):
    """test_protocol.py:107-134: ranks against a subset never exceed the subset's size."""
    X = _graph()
    model = _complex_nll()
    model.fit(X["train"])
    X_filter = np.concatenate((X["train"], X["valid"], X["test"]))
    entities_subset = np.concatenate([X["test"][::5, 0], X["test"][::5, 2]], 0)
    ranks = evaluate_performance(X["test"][::5], model=model, filter_triples=X_filter, corrupt_side="s+o", verbose=True,
                                 entities_subset=entities_subset)
    assert np.sum(ranks.reshape(-1) > len(entities_subset)) == 0


def test_evaluate_performance_ranking_against_shuffled_all_entities(fake):
    """test_protocol.py:137-172: the default protocol against all entities == the same protocol with entities_subset = all
    entities in another order.  (The reference passes random.shuffle's return value, None; both readings are checked.)"""
    X = _graph()
    model = _complex_nll()
    model.fit(X["train"])
    X_filter = np.concatenate((X["train"], X["valid"], X["test"]))
    ranks_all = evaluate_performance(X["test"][::5], model, X_filter, verbose=True, corrupt_side="s,o")
    ranks_none = evaluate_performance(X["test"][::5], model, X_filter, verbose=True, corrupt_side="s,o", entities_subset=None)
    shuffled = list(model.ent_to_idx.keys())
    np.random.RandomState(0).shuffle(shuffled)
    ranks_shuffled = evaluate_performance(X["test"][::5], model, X_filter, verbose=True, corrupt_side="s,o", entities_subset=shuffled)
    assert mrr_score(ranks_all) == mrr_score(ranks_none) == mrr_score(ranks_shuffled)
    np.testing.assert_array_equal(ranks_all, ranks_shuffled)


@pytest.mark.parametrize("with_filter", [False, True])
def test_evaluate_performance_default_protocol(fake, with_filter):
    """test_protocol.py:175-232 (without filter) and :235-301 (with): 'o' ranks followed by 's' ranks have the mean rank of the
    's,o' protocol."""
    wn = _graph()
    X_filter = np.concatenate((wn["train"], wn["valid"], wn["test"])) if with_filter else None
    model = TransE(batches_count=10, seed=0, epochs=1, k=50, eta=10, verbose=True,
                   embedding_model_params={"normalize_ent_emb": False, "norm": 1}, loss="self_adversarial",
                   loss_params={"margin": 1, "alpha": 0.5}, optimizer="adam", optimizer_params={"lr": 0.0005})
    model.fit(wn["train"])
    ranks_sep = []
    ranks_sep.extend(evaluate_performance(wn["test"][::2], model, X_filter, verbose=True, corrupt_side="o"))
    ranks_sep.extend(evaluate_performance(wn["test"][::2], model, X_filter, verbose=True, corrupt_side="s"))
    ranks = evaluate_performance(wn["test"][::2], model, X_filter, verbose=True, corrupt_side="s,o")
    np.testing.assert_equal(mr_score(ranks_sep), mr_score(ranks))
    assert mrr_score(ranks) is not np.inf and hits_at_n_score(ranks, 10) <= 1.0
    # use_default_protocol forces 's,o' whatever corrupt_side says (evaluation/protocol.py:871-873)
    np.testing.assert_array_equal(evaluate_performance(wn["test"][::2], model, X_filter, corrupt_side="o", use_default_protocol=True), ranks)


def test_evaluate_performance_so_side_corruptions(fake):
    """test_protocol.py:304-325 and :328-352: 's+o' gives one rank per triple, with and without a filter."""
    X = _graph()
    model = ComplEx(batches_count=10, seed=0, epochs=2, k=16, eta=10, loss="nll", optimizer="adam", optimizer_params={"lr": 0.01}, verbose=True)
    model.fit(X["train"])
    ranks = evaluate_performance(X["test"][::2], model=model, verbose=True, corrupt_side="s+o")
    assert ranks.shape == (len(X["test"][::2]),) and np.isfinite(mrr_score(ranks)) and 0 <= hits_at_n_score(ranks, n=10) <= 1
    X_filter = np.concatenate((X["train"], X["valid"], X["test"]))
    ranks_f = evaluate_performance(X["test"][::2], model, X_filter, verbose=True, corrupt_side="s+o")
    assert ranks_f.shape == ranks.shape and np.all(ranks_f <= ranks) and ranks_f.min() >= 1


@pytest.mark.parametrize("cls,kw", [(ComplEx, dict(k=15, optimizer_params={"lr": 0.1}, eta=10, loss="nll", optimizer="adagrad")),
                                    (TransE, dict(k=10, eta=5, optimizer_params={"lr": 0.1}, loss="pairwise", loss_params={"margin": 5},
                                                  optimizer="adagrad"))])
def test_evaluate_performance_trained_models(fake, cls, kw):
    """test_protocol.py:355-381 (ComplEx nll adagrad) and :384-414 (TransE pairwise adagrad): train on train+valid, rank a test
    prefix against the full filter (both skipped in the reference's CI for their run time; run here at toy scale)."""
    X = _graph()
    model = cls(batches_count=10, seed=0, epochs=3, verbose=True, **kw)
    model.fit(np.concatenate((X["train"], X["valid"])))
    filter_triples = np.concatenate((X["train"], X["valid"], X["test"]))
    ranks = evaluate_performance(X["test"][:20], model=model, filter_triples=filter_triples, verbose=True)
    assert ranks.shape == (20, 2) and ranks.min() >= 1 and ranks.max() <= len(model.ent_to_idx)
    assert 0 < mrr_score(ranks) <= 1 and 0 <= hits_at_n_score(ranks, n=10) <= 1


FIVE = np.array([["a", "x", "b"], ["c", "x", "d"], ["e", "x", "f"], ["b", "y", "h"], ["a", "y", "l"]])


def test_generate_corruptions_for_eval():
    """test_protocol.py:419-455, the expected array verbatim."""
    rel_to_idx, ent_to_idx = create_mappings(FIVE)
    X = to_idx(FIVE, ent_to_idx=ent_to_idx, rel_to_idx=rel_to_idx)
    all_ent = np.array(list(ent_to_idx.values()), dtype=np.int64)
    x_n_actual = generate_corruptions_for_eval(np.array([X[0]]), all_ent)
    x_n_expected = np.array([[0, 0, 0], [0, 0, 1], [0, 0, 2], [0, 0, 3], [0, 0, 4], [0, 0, 5], [0, 0, 6], [0, 0, 7],
                             [0, 0, 1], [1, 0, 1], [2, 0, 1], [3, 0, 1], [4, 0, 1], [5, 0, 1], [6, 0, 1], [7, 0, 1]])
    np.testing.assert_array_equal(x_n_actual, x_n_expected)


def test_to_idx():
    """test_protocol.py:492-499."""
    X = np.array([["a", "x", "b"], ["c", "y", "d"]])
    rel_to_idx, ent_to_idx = create_mappings(X)
    np.testing.assert_array_equal(to_idx(X, ent_to_idx=ent_to_idx, rel_to_idx=rel_to_idx), [[0, 0, 1], [2, 1, 3]])


def test_filter_unseen_entities():
    """test_protocol.py:515-530: the model is anything with an ent_to_idx dictionary."""
    base_model = namedtuple("test_model", "ent_to_idx")
    X = np.array([["a", "x", "b"], ["c", "y", "d"], ["e", "y", "d"]])
    model = base_model({"a": 1, "b": 2, "c": 3, "d": 4})
    np.testing.assert_array_equal(filter_unseen_entities(X, model), np.array([["a", "x", "b"], ["c", "y", "d"]]))


@pytest.mark.parametrize("side", ["s,o", "s", "o"])
def test_generate_corruptions_for_fit_sides(side):
    """test_protocol.py:534-610.  The reference's expected arrays are what TensorFlow's generator draws for seed 0 and cannot be
    reproduced by any other library; what the three tests establish -- shape, only the chosen side is replaced, replacements come
    from range(entities_size), the same seed gives the same corruptions -- is checked."""
    rel_to_idx, ent_to_idx = create_mappings(FIVE)
    X = to_idx(FIVE, ent_to_idx=ent_to_idx, rel_to_idx=rel_to_idx)
    X_corr = generate_corruptions_for_fit(X, eta=1, corrupt_side=side, entities_size=len(X), rnd=0)
    assert X_corr.shape == X.shape
    np.testing.assert_array_equal(X_corr[:, 1], X[:, 1])
    s_changed, o_changed = X_corr[:, 0] != X[:, 0], X_corr[:, 2] != X[:, 2]
    assert not np.any(s_changed & o_changed)
    if side == "s":
        assert not o_changed.any() and np.all((X_corr[:, 0] >= 0) & (X_corr[:, 0] < len(X)))
    if side == "o":
        assert not s_changed.any() and np.all((X_corr[:, 2] >= 0) & (X_corr[:, 2] < len(X)))
    np.testing.assert_array_equal(X_corr, generate_corruptions_for_fit(X, eta=1, corrupt_side=side, entities_size=len(X), rnd=0))
    assert generate_corruptions_for_fit(X, eta=3, corrupt_side=side, entities_size=len(X), rnd=0).shape == (15, 3)


def test_train_test_split():
    """test_protocol.py:613-648, the expected split verbatim (backward_compatible=True)."""
    X = np.array([["a", "y", "b"], ["a", "y", "c"], ["c", "y", "a"], ["d", "y", "e"], ["e", "y", "f"], ["f", "y", "c"], ["f", "y", "c"]])
    X_train, X_test = train_test_split_no_unseen(X, test_size=2, seed=0, backward_compatible=True)
    np.testing.assert_array_equal(X_train, np.array([["a", "y", "b"], ["c", "y", "a"], ["d", "y", "e"], ["e", "y", "f"], ["f", "y", "c"]]))
    np.testing.assert_array_equal(X_test, np.array([["a", "y", "c"], ["f", "y", "c"]]))


def test_train_test_split_fast():
    """test_protocol.py:651-685 on a synthetic graph in place of FB15k-237: sizes add up, the training side keeps every entity and
    relation, an impossible split raises the reference's message, and allow_duplication makes it possible."""
    g = _graph(E=60, R=3, n=900, seed=5)
    x_all = np.concatenate([g["train"], g["valid"], g["test"]], 0)
    unique_entities = len(set(x_all[:, 0]).union(x_all[:, 2]))
    unique_rels = len(set(x_all[:, 1]))
    x_train, x_test = train_test_split_no_unseen(x_all, 0.5)
    assert x_train.shape[0] + x_test.shape[0] == x_all.shape[0]
    assert len(set(x_train[:, 0]).union(x_train[:, 2])) == unique_entities and len(set(x_train[:, 1])) == unique_rels
    with pytest.raises(Exception) as e:
        train_test_split_no_unseen(x_all, 0.99, allow_duplication=False)
    assert str(e.value) == ("Cannot create a test split of the desired size. "
                            "Some entities will not occur in both training and test set. "
                            "Set allow_duplication=True,"
                            "remove filter on test predicates or "
                            "set test_size to a smaller value.")
    x_train, x_test = train_test_split_no_unseen(x_all, 0.99, allow_duplication=True)
    assert x_train.shape[0] + x_test.shape[0] > x_all.shape[0]
    assert len(set(x_train[:, 0]).union(x_train[:, 2])) == unique_entities and len(set(x_train[:, 1])) == unique_rels


# ------------------------------------------------------------------ tests/emgraph/models/test_models.py
def test_large_graph_mode(fake):
    """test_models.py:38-60: set_entity_threshold(10) switches the reference to host-paged embeddings (SGD only).  Here the table
    stays in HBM (sharded when it must be), so the same script runs the normal path and gives the numbers of the normal mode."""
    X = _graph()
    kw = dict(batches_count=20, seed=555, epochs=1, k=10, loss="multiclass_nll", loss_params={"margin": 5}, verbose=True,
              optimizer="sgd", optimizer_params={"lr": 0.001})
    normal = ComplEx(**kw)
    normal.fit(X["train"])
    set_entity_threshold(10)
    try:
        model = ComplEx(**kw)
        model.fit(X["train"])
        X_filter = np.concatenate((X["train"], X["valid"], X["test"]), axis=0)
        ranks = evaluate_performance(X["test"][::5], model, X_filter, verbose=True, corrupt_side="s,o")
        y = model.predict(X["test"][:1])
    finally:
        reset_entity_threshold()
    assert ranks.shape[1] == 2 and y.shape == (1,)
    np.testing.assert_array_equal(y, normal.predict(X["test"][:1]))
    assert models.ENTITY_THRESHOLD == 5e5


def test_output_sizes(fake):
    """test_models.py:63-98: embedding matrix sizes match the data (entities, relations, k), in both modes."""
    X = _graph()

    def perform_test():
        k = 5
        unique_entities = np.unique(np.concatenate([X["train"][:, 0], X["train"][:, 2]], 0))
        unique_relations = np.unique(X["train"][:, 1])
        model = TransE(batches_count=20, seed=555, epochs=1, k=k, loss="multiclass_nll", loss_params={"margin": 5}, verbose=True,
                       optimizer="sgd", optimizer_params={"lr": 0.001})
        model.fit(X["train"])
        assert model.trained_model_params[0].shape[0] == len(unique_entities)
        assert model.trained_model_params[1].shape[0] == len(unique_relations)
        assert model.trained_model_params[0].shape[1] == k
        assert model.trained_model_params[1].shape[1] == k

    perform_test()
    set_entity_threshold(10)
    try:
        perform_test()
    finally:
        reset_entity_threshold()


def test_large_graph_mode_adam(fake):
    """test_models.py:101-120: the reference refuses adam in large-graph mode (and the test swallows the exception); the sparse
    row-wise Adam here has no such limit, the fit simply runs."""
    X = _graph()
    set_entity_threshold(10)
    try:
        model = ComplEx(batches_count=20, seed=555, epochs=1, k=10, loss="multiclass_nll", loss_params={"margin": 5}, verbose=True,
                        optimizer="adam", optimizer_params={"lr": 0.001})
        model.fit(X["train"])
    finally:
        reset_entity_threshold()
    assert model.is_fitted and np.all(np.isfinite(model.trained_model_params[0]))


@pytest.mark.parametrize("with_filter", [True, False])
def test_fit_predict_TransE_early_stopping(fake, with_filter):
    """test_models.py:123-151 (with x_filter) and :154-180 (without): positional early_stopping arguments, a strided validation set."""
    X = _graph()
    model = TransE(batches_count=1, seed=555, epochs=7, k=10, loss="pairwise", loss_params={"margin": 5}, verbose=True,
                   optimizer="adagrad", optimizer_params={"lr": 0.1})
    es = {"x_valid": X["valid"][::4], "criteria": "mrr", "stop_interval": 2, "burn_in": 1, "check_interval": 2}
    if with_filter:
        es["x_filter"] = np.concatenate((X["train"], X["valid"], X["test"]))
    model.fit(X["train"], True, es)
    y = model.predict(X["test"][:1])
    assert y.shape == (1,) and np.isfinite(y).all()


def test_retrain(fake):
    """test_models.py:338-367."""
    model = ComplEx(batches_count=1, seed=555, epochs=20, k=10, loss="pairwise", loss_params={"margin": 1}, regularizer="LP",
                    regularizer_params={"lambda": 0.1, "p": 2}, optimizer="adagrad", optimizer_params={"lr": 0.1})
    model.fit(TOY)
    y_pred_1st = model.predict(np.array([["f", "y", "e"], ["b", "y", "d"]]))
    model.fit(TOY)
    y_pred_2nd = model.predict(np.array([["f", "y", "e"], ["b", "y", "d"]]))
    np.testing.assert_array_equal(y_pred_1st, y_pred_2nd)


@pytest.mark.parametrize("cls,kw", [(TransE, dict(k=20, loss="pairwise", loss_params={"margin": 5})),
                                    (ComplEx, dict(k=10, loss="pairwise", loss_params={"margin": 1}, regularizer="LP",
                                                   regularizer_params={"lambda": 0.1, "p": 2}))])
def test_fit_predict_wn18(fake, cls, kw):
    """test_models.py:370-386 (TransE) and :412-428 (ComplEx + LP): one batch per epoch over the whole training set, predict one
    test triple."""
    X = _graph()
    model = cls(batches_count=1, seed=555, epochs=5, verbose=True, optimizer="adagrad", optimizer_params={"lr": 0.1}, **kw)
    model.fit(X["train"])
    y = model.predict(X["test"][:1])
    assert y.shape == (1,) and np.isfinite(y).all()


def test_missing_entity_ComplEx(fake):
    """test_models.py:389-409: unknown subject, predicate or object -> ValueError."""
    model = ComplEx(batches_count=1, seed=555, epochs=2, k=5)
    model.fit(TOY)
    with pytest.raises(ValueError):
        model.predict(["a", "y", "zzzzzzzzzzz"])
    with pytest.raises(ValueError):
        model.predict(["a", "xxxxxxxxxx", "e"])
    with pytest.raises(ValueError):
        model.predict(["zzzzzzzz", "y", "e"])


def _distmult(epochs=1):
    return DistMult(batches_count=2, seed=555, epochs=epochs, k=10, loss="pairwise", loss_params={"margin": 5}, optimizer="adagrad",
                    optimizer_params={"lr": 0.1})


def test_lookup_embeddings(fake):
    """test_models.py:431-455."""
    model = _distmult(20)
    model.fit(TOY)
    emb = model.get_embeddings(["a", "b"], embedding_type="entity")
    assert emb.shape == (2, 10)
    np.testing.assert_array_equal(emb, model.trained_model_params[0][[model.ent_to_idx["a"], model.ent_to_idx["b"]]])
    assert model.get_embeddings(["y"], embedding_type="relation").shape == (1, 10)


def test_is_fitted_on(fake):
    """test_models.py:458-506."""
    model = _distmult()
    model.fit(TOY5)
    X1 = np.array([["a", "y", "b"], ["b", "y", "a"], ["a", "y", "c"], ["c", "z", "a"], ["g", "z", "d"]])
    X2 = np.array([["a", "y", "b"], ["b", "y", "a"], ["a", "y", "c"], ["c", "z", "a"], ["a", "x", "d"]])
    assert model.is_fitted_on(TOY5) is True
    assert model.is_fitted_on(X1) is False
    assert model.is_fitted_on(X2) is False


def test_predict(fake):
    """test_models.py:967-992: labels and ids give the same predictions."""
    model = _distmult()
    model.fit(TOY5)
    preds1 = model.predict(TOY5)
    preds2 = model.predict(to_idx(TOY5, model.ent_to_idx, model.rel_to_idx), from_idx=True)
    np.testing.assert_array_equal(preds1, preds2)


def test_predict_twice(fake):
    """test_models.py:995-1024."""
    model = _distmult()
    model.fit(TOY5)
    preds1 = model.predict(np.array([["a", "y", "b"], ["b", "y", "a"]]))
    preds2 = model.predict(np.array([["a", "y", "c"], ["c", "z", "a"]]))
    assert not np.array_equal(preds1, preds2)


def test_predict_before_fit_raises():
    """models/EmbeddingModel.py:2115-2118 / :2197-2200: RuntimeError('Model has not been fitted.')."""
    model = _distmult()
    with pytest.raises(RuntimeError):
        model.predict(TOY5)
    with pytest.raises(RuntimeError):
        model.is_fitted_on(TOY5)
    with pytest.raises(RuntimeError):
        model.get_embeddings(["a"])


# ------------------------------------------------------------------ tests/emgraph/utils/test_model_utils.py
@pytest.mark.parametrize("model_name", ["ComplEx", "TransE", "DistMult"])
def test_save_and_restore_model(fake, model_name, tmp_path):
    """test_model_utils.py:18-80."""
    model = getattr(models, model_name)(batches_count=2, seed=555, epochs=20, k=10, optimizer="adagrad", optimizer_params={"lr": 0.1})
    model.fit(TOY)
    example_name = str(tmp_path / "helloworld.pkl")
    save_model(model, model_name_path=example_name)
    loaded_model = restore_model(model_name_path=example_name)
    assert loaded_model is not None
    assert loaded_model.all_params == model.all_params
    assert loaded_model.is_fitted == model.is_fitted
    assert loaded_model.ent_to_idx == model.ent_to_idx
    assert loaded_model.rel_to_idx == model.rel_to_idx
    for i in range(len(loaded_model.trained_model_params)):
        np.testing.assert_array_equal(loaded_model.trained_model_params[i], model.trained_model_params[i])
    q = np.array([["f", "y", "e"], ["b", "y", "d"]])
    np.testing.assert_array_equal(loaded_model.predict(q), model.predict(q))
    np.testing.assert_array_equal(loaded_model.get_embeddings(["a", "b"], embedding_type="entity"),
                                  model.get_embeddings(["a", "b"], embedding_type="entity"))
    os.remove(example_name)


def test_restore_model_errors():
    """test_model_utils.py:83-85."""
    with pytest.raises(FileNotFoundError):
        restore_model(model_name_path="filenotfound.model")


# ------------------------------------------------------------------ tests/emgraph/models/test_initializers.py, test_regularizers.py, test_misc.py
def _table(initializer, params, rows, cols, which="entity", seed=0):
    m = DistMult(k=cols, seed=seed, initializer=initializer, initializer_params=params)
    return m._init_table(rows, cols, which)


def test_random_normal():
    """test_initializers.py:12-22: mean / std of the 'normal' initializer (the reference compares its numpy and TF variants to one
    decimal; here the one implementation is compared with the requested moments)."""
    v = _table("normal", {"mean": 0.5, "std": 0.1}, 100, 10)
    assert v.shape == (100, 10) and v.dtype == np.float32
    assert abs(float(np.mean(v)) - 0.5) < 0.02 and abs(float(np.std(v)) - 0.1) < 0.02


def test_glorot_uniform():
    """test_initializers.py:40-52 (and :25-37: the TF path of the reference draws uniformly whatever 'uniform' says):
    values fill (-sqrt(6 / (rows + cols)), +sqrt(6 / (rows + cols)))."""
    for params in ({"uniform": True}, {"uniform": False}):
        v = _table("glorot_uniform", params, 20, 100)
        lim = np.sqrt(6.0 / 120)
        assert -lim <= np.min(v) < -lim + 0.005 and lim - 0.005 < np.max(v) <= lim
        assert abs(float(np.mean(v))) < 0.02


def test_random_uniform():
    """test_initializers.py:55-67."""
    v = _table("uniform", {"low": 0.1, "high": 0.4}, 100, 10)
    assert 0.1 <= np.min(v) < 0.105 and 0.395 < np.max(v) <= 0.4


def test_constant(fake):
    """test_initializers.py:70-97: the constant initializer hands back the arrays it was given -- and a model built on it starts
    from them (lr = 0 leaves them untouched)."""
    rs = np.random.RandomState(117)
    ent_init = rs.normal(1, 1, size=(300, 30))
    rel_init = rs.normal(2, 2, size=(10, 30))
    params = {"entity": ent_init, "relation": rel_init}
    np.testing.assert_array_equal(_table("constant", params, 300, 30, "entity"), ent_init.astype(np.float32))
    np.testing.assert_array_equal(_table("constant", params, 10, 30, "relation"), rel_init.astype(np.float32))
    with pytest.raises(AssertionError):
        _table("constant", params, 301, 30, "entity")
    with pytest.raises(Exception):
        _table("constant", {"entity": ent_init}, 10, 30, "relation")
    E, R = 6, 1
    m = DistMult(k=30, eta=1, epochs=1, batches_count=1, seed=0, optimizer="sgd", optimizer_params={"lr": 0.0}, initializer="constant",
                 initializer_params={"entity": ent_init[:E], "relation": rel_init[:R]})
    m.fit(TOY)
    np.testing.assert_array_equal(m.trained_model_params[0], ent_init[:E].astype(np.float32))
    np.testing.assert_array_equal(m.trained_model_params[1], rel_init[:R].astype(np.float32))


@pytest.mark.parametrize("p,lam,expected", [(1, 1.0, 9.0), (1, [2.0, 3.0], 24.0), (2, 1.0, 15.0), (2, [2.0, 3.0], 42.0)])
def test_lp_regularizer(p, lam, expected):
    """test_regularizers.py:7-22 (L1) and :25-40 (L2): lambda_i * sum(|param_i|^p) over the trainable tables, lambda a scalar or one
    weight per table -- through the constructor's parsing and the oracle's training step (the penalty is what the regulariser adds
    to the loss of a step)."""
    kw = EmbeddingModelReg.parse("LP", {"lambda": lam, "p": p})
    ent = np.array([[1, -1, 1], [0, 0, 0]], dtype=np.float32)  # p1 (+ a zero row so that a corruption has somewhere to go)
    rel = np.array([[2, -2, 2]], dtype=np.float32)             # p2
    pos = np.array([[0, 0, 0]])
    args = ("DistMult", 3, "nll", 1, ent, rel, pos, np.array([1], np.uint8), np.array([1]))
    with_reg = ko.train_step(*args, **kw)["loss"]
    without = ko.train_step(*args)["loss"]
    np.testing.assert_allclose(with_reg - without, expected, rtol=1e-6)


class EmbeddingModelReg:
    @staticmethod
    def parse(regularizer, params):
        return models.EmbeddingModel._parse_regularizer(regularizer, params)


def test_lp_regularizer_argument_checks():
    """regularizers/lp.py:68-104: p must be an integer, lambda a scalar or one weight per table."""
    with pytest.raises(Exception):
        EmbeddingModelReg.parse("LP", {"lambda": 1.0, "p": 1.5})
    with pytest.raises(ValueError):
        EmbeddingModelReg.parse("LP", {"lambda": [1.0, 2.0, 3.0], "p": 2})
    assert EmbeddingModelReg.parse(None, {}) == dict(reg_p=0, reg_lambda_ent=0.0, reg_lambda_rel=0.0)


def test_get_entity_triples():
    """test_misc.py:6-29."""
    from emgraph_b200 import get_entity_triples
    X = np.array([["a", "y", "b"], ["a", "y", "c"], ["c", "y", "a"], ["d", "y", "e"], ["e", "y", "f"], ["f", "y", "c"]])
    XN = np.array([["a", "y", "c"], ["c", "y", "a"], ["f", "y", "c"]])
    assert np.all(get_entity_triples("c", X) == XN)
