"""Hyper-parameter search on top of fit() and the ranking kernel: ``select_best_model_ranking``
(reference evaluation/protocol.py:1317-1703) and the parameter-grid helpers its tests exercise
(:1044-1315; tests/emgraph/evaluation/test_protocol.py:679-1153).

Every candidate is trained by ``model.fit`` (the fused train step) and scored by ``evaluate_performance`` (the
filtered ranking sweep), so a search is a loop of two kernel paths; nothing here touches the device itself."""
from __future__ import annotations

import itertools
from collections.abc import Iterable

import numpy as np

# Parameters each plug-in reads (the reference keeps these lists on its registries: losses/*.py, training/*.py,
# regularizers/lp.py, initializers/*.py decorators; models/*.py `register_model`).  A nested parameter that
# the chosen plug-in does not read is dropped, so that e.g. {"regularizer": None, "regularizer_params": {"p": 1}}
# and {"regularizer": None, "regularizer_params": {"p": 3}} count as ONE configuration.
_COMMON_MODEL = ("negative_corruption_entities", "corrupt_side", "corrupt_sides", "non_linearity")
EXTERNAL_PARAMS = {
    "loss": {"pairwise": ("margin",), "nll": (), "multiclass_nll": (), "absolute_margin": ("margin",),
             "self_adversarial": ("margin", "alpha")},
    "regularizer": {"LP": ("p", "lambda")},
    "optimizer": {"adam": ("lr",), "adagrad": ("lr",), "momentum": ("lr", "momentum"),
                  "sgd": ("lr", "decay_cycle", "end_lr", "sine_decay", "expand_factor", "decay_lr_rate")},
    "initializer": {"constant": ("entity", "relation"), "glorot_uniform": ("uniform",), "xavier": ("uniform",),
                    "normal": ("mean", "std"), "uniform": ("low", "high")},
    # the reference's lists (models/*.py) plus the keys its own _get_model_loss reads from
    # embedding_model_params (models/EmbeddingModel.py:679, :780): dropping those would silently change the model
    "model_name": {"TransE": ("norm", "normalize_ent_emb") + _COMMON_MODEL,
                   "DistMult": ("normalize_ent_emb",) + _COMMON_MODEL,
                   "ComplEx": _COMMON_MODEL, "HolE": _COMMON_MODEL},
}
_NESTED_OF = {"loss": "loss_params", "regularizer": "regularizer_params", "optimizer": "optimizer_params",
              "initializer": "initializer_params", "model_name": "embedding_model_params"}


def _remove_unused_params(params):
    """Copy of `params` whose nested *_params dictionaries keep only what the selected plug-in reads
    (evaluation/protocol.py:1044-1095; an unknown / None plug-in keeps nothing)."""
    out = dict(params)
    for selector, nested in _NESTED_OF.items():
        if selector not in out or nested not in out:
            continue
        try:
            known = EXTERNAL_PARAMS[selector].get(out[selector])
        except TypeError:  # unhashable selector value
            known = None
        out[nested] = {} if known is None else {k: v for k, v in out[nested].items() if k in known}
    return out


def _flatten_nested_keys(dictionary):
    """{"a": {"b": 1}, "c": 2} -> {("a", "b"): 1, "c": 2} (one level, evaluation/protocol.py:1098-1122)."""
    flat = {}
    for k, v in dictionary.items():
        if type(v) is dict:
            for k2, v2 in v.items():
                flat[(k, k2)] = v2
        else:
            flat[k] = v
    return flat


def _unflatten_nested_keys(dictionary):
    """Inverse of _flatten_nested_keys (evaluation/protocol.py:1125-1149)."""
    out = {}
    for k, v in dictionary.items():
        if type(k) is tuple:
            out.setdefault(k[0], {})[k[1]] = v
        else:
            out[k] = v
    return out


def _freeze(v):
    if isinstance(v, (list, tuple)):
        return tuple(_freeze(x) for x in v)
    if isinstance(v, np.ndarray):
        return (v.shape, v.tobytes())
    if isinstance(v, dict):
        return tuple(sorted((k, _freeze(x)) for k, x in v.items()))
    return v


def _get_param_hash(param):
    """Hash of a configuration after unused nested parameters are dropped (evaluation/protocol.py:1152-1173).
    `param` may be nested or flattened."""
    flat = _flatten_nested_keys(_remove_unused_params(_unflatten_nested_keys(param)))
    return hash(frozenset((k, _freeze(v)) for k, v in flat.items()))


class ParamHistory(object):
    """Set of configurations already seen, compared after dropping unused parameters (:1176-1212)."""

    def __init__(self):
        self.param_hash_history = set()

    def add(self, param):
        self.param_hash_history.add(_get_param_hash(param))

    def __contains__(self, other):
        return _get_param_hash(other) in self.param_hash_history


def _next_hyperparam(param_grid):
    """Every distinct configuration of a grid of lists, in itertools.product order (:1215-1243)."""
    seen = ParamHistory()
    flat = _flatten_nested_keys(param_grid)
    names = list(flat.keys())
    for values in itertools.product(*[flat[n] for n in names]):
        cand = dict(zip(names, values))
        if cand in seen:
            continue
        seen.add(cand)
        yield _remove_unused_params(_unflatten_nested_keys(cand))


def _sample_parameters(param_grid):
    """One random configuration: callables are called, lists sampled with np.random.choice (:1246-1269)."""
    out = {}
    for k, v in param_grid.items():
        if callable(v):
            out[k] = v()
        elif type(v) is dict:
            out[k] = _sample_parameters(v)
        elif isinstance(v, Iterable) and type(v) is not str:
            vals = list(v)
            pick = vals[np.random.randint(len(vals))] if any(x is None or isinstance(x, (list, dict, str)) for x in vals) \
                else np.random.choice(vals)
            out[k] = pick
        else:
            out[k] = v
    return out


def _next_hyperparam_random(param_grid):
    """Endless stream of distinct random configurations (:1272-1295)."""
    seen = ParamHistory()
    while True:
        cand = _sample_parameters(param_grid)
        if cand in seen:
            continue
        seen.add(cand)
        yield _remove_unused_params(cand)


def _scalars_into_lists(param_grid):
    """In place: scalars (and strings) become one-element lists, nested dictionaries recursively (:1298-1314)."""
    for k, v in param_grid.items():
        if type(v) is dict:
            _scalars_into_lists(v)
        elif type(v) is str or not (callable(v) or isinstance(v, Iterable)):
            param_grid[k] = [v]


def _summary(ranks):
    from .evaluation import hits_at_n_score, mr_score, mrr_score
    return {"mrr": mrr_score(ranks), "mr": mr_score(ranks), "hits_1": hits_at_n_score(ranks, n=1),
            "hits_3": hits_at_n_score(ranks, n=3), "hits_10": hits_at_n_score(ranks, n=10)}


def select_best_model_ranking(model_class, X_train, X_valid, X_test, param_grid, max_combinations=None,
                              param_grid_random_seed=0, use_filter=True, early_stopping=False, early_stopping_params=None,
                              use_test_for_selection=False, entities_subset=None, corrupt_side="s,o",
                              use_default_protocol=False, retrain_best_model=False, verbose=False):
    """Grid search (``max_combinations=None``) or random search over ``param_grid`` by validation MRR
    (evaluation/protocol.py:1317-1703).

    Returns ``(best_model, best_params, best_mrr_train, ranks_test, test_evaluation, experimental_history)``;
    a configuration whose training raises is recorded with ``{"exception": str(e)}`` and skipped, as in the
    reference.  Every fit and every evaluation runs on the GPU engine."""
    from .evaluation import evaluate_performance
    if use_default_protocol:
        corrupt_side = "s,o"
    early_stopping_params = {} if early_stopping_params is None else early_stopping_params
    param_grid["model_name"] = model_class.name
    _scalars_into_lists(param_grid)
    if max_combinations is not None:
        np.random.seed(param_grid_random_seed)
        combos = itertools.islice(_next_hyperparam_random(param_grid), max_combinations)
    else:
        combos = _next_hyperparam(param_grid)
    if early_stopping and "x_valid" not in early_stopping_params:
        early_stopping_params["x_valid"] = X_valid
    X_filter = np.concatenate((X_train, X_valid, X_test)) if use_filter else None
    selection = X_test if use_test_for_selection else X_valid
    eval_kw = dict(filter_triples=X_filter, verbose=verbose, entities_subset=entities_subset,
                   use_default_protocol=use_default_protocol, corrupt_side=corrupt_side)

    best_model, best_params, best_mrr_train = None, None, 0
    history = []
    for params in combos:
        entry = {"model_name": params["model_name"], "model_params": params}
        del params["model_name"]
        try:
            model = model_class(**params)
            model.fit(X_train, early_stopping, early_stopping_params)
            res = _summary(evaluate_performance(selection, model=model, **eval_kw))
            entry["results"] = res
            if verbose:
                print("mr: {mr} mrr: {mrr} hits 1: {hits_1} hits 3: {hits_3} hits 10: {hits_10}".format(**res)
                      + ", model: {}, params: {}".format(type(model).__name__, params))
            if res["mrr"] > best_mrr_train:
                best_model, best_params, best_mrr_train = model, params, res["mrr"]
        except Exception as e:  # noqa: BLE001 -- the reference records the failure and carries on (:1633-1643)
            entry["results"] = {"exception": str(e)}
            if verbose:
                print("Exception occurred for parameters:{}\n{}".format(params, e))
        history.append(entry)

    if best_model is None:
        nan = float("nan")
        return None, None, best_mrr_train, [], {"mrr": nan, "mr": nan, "hits_1": nan, "hits_3": nan, "hits_10": nan}, history
    if retrain_best_model:
        best_model.fit(np.concatenate((X_train, X_valid)), early_stopping, early_stopping_params)
    ranks_test = evaluate_performance(X_test, model=best_model, **eval_kw)
    test_evaluation = _summary(ranks_test)
    if verbose:
        print("Best model test results: {}, model: {}, params: {}".format(test_evaluation, type(best_model).__name__, best_params))
    return best_model, best_params, best_mrr_train, ranks_test, test_evaluation, history
