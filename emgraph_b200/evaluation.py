"""Host-side mirror of ``emgraph.evaluation`` for the hot path: ``evaluate_performance`` (filtered
ranking, evaluation/protocol.py:726-979) and the rank metrics (evaluation/metrics.py).  The filter
index is built on the device (replaces the temp-file SQLite DB of datasets/sqlite_adapter.py)."""
from __future__ import annotations

import warnings

import numpy as np

from .engine import to_dev_i32
from .models import EmbeddingModel, _fixed_width, _sorted_factorize, create_mappings, to_idx  # noqa: F401  (re-exported like the reference)

from .model_selection import select_best_model_ranking  # noqa: E402,F401  (evaluation/__init__.py:6 exports it)

TOO_MANY_ENTITIES_TH = 50000  # evaluation/protocol.py:21


class EvalDataset:
    """What NumpyDatasetAdapter + SQLiteAdapter carry during evaluation: mapped test triples and the
    mapped filter triples (datasets/numpy_adapter.py:79-131, :229-247)."""

    def __init__(self, test_idx, filter_idx=None):
        self.test_idx = np.ascontiguousarray(test_idx, dtype=np.int32).reshape(-1, 3)
        self.filter_idx = None if filter_idx is None else np.ascontiguousarray(filter_idx, dtype=np.int32).reshape(-1, 3)
        self._test_dev = None
        self._test_pin = None
        self._engine = None
        self._perm_built = False

    def get_size(self, dataset_type="test"):
        return self.test_idx.shape[0] if dataset_type == "test" else (0 if self.filter_idx is None else self.filter_idx.shape[0])

    def test_device(self, device):
        if self._test_dev is None:
            self._test_dev = to_dev_i32(self.test_idx, device)
        return self._test_dev

    def test_host_pinned(self):
        """The mapped test triples as a pinned int32 CPU tensor (source of the per-call H2D copy)."""
        if self._test_pin is None:
            import torch
            self._test_pin = torch.from_numpy(self.test_idx).pin_memory()
        return self._test_pin

    def build_filter(self, engine, E, R, perm=None):
        """Device filter index; built once per handle (the reference builds its SQLite DB once per
        evaluate_performance call, datasets/numpy_adapter.py:229-247).  perm: optional device int64 [E]
        re-labelling of the entities (entities_subset ranking sweeps a permuted table)."""
        if self._engine is engine and perm is None and not self._perm_built:
            return
        f = self.filter_idx if self.filter_idx is not None else np.zeros((0, 3), np.int32)
        fd = to_dev_i32(f, engine.tdev)
        if perm is not None and fd.shape[0]:
            import torch
            fl = fd.long()
            fd = torch.stack([perm[fl[:, 0]], fl[:, 1], perm[fl[:, 2]]], 1).to(torch.int32).contiguous()
        engine.filter_build(fd, E, R)
        self._engine = engine
        self._perm_built = perm is not None

    def cleanup(self):
        if self._engine is not None:
            self._engine.filter_clear()
            self._engine = None
        self._test_dev = None


def filter_unseen_entities(X, model, verbose=False):
    """evaluation/protocol.py:1014-1041: drop triples whose subject or object the model never saw."""
    X = np.asarray(X)
    index = getattr(model, "_ent_index", None)
    if index is not None:
        keep = index.contains(X[:, 0]) & index.contains(X[:, 2])
    else:  # anything that carries an ent_to_idx dictionary (the reference's own test passes a namedtuple)
        known = np.array(list(model.ent_to_idx.keys()))
        keep = np.isin(X[:, 0], known) & np.isin(X[:, 2], known)
    n_removed = int(X.shape[0] - keep.sum())
    if n_removed > 0:
        if verbose:
            print("Removing {} triples containing unseen entities. ".format(n_removed))
        return X[keep]
    return X


def check_filter_size(model, corruption_entities):
    """evaluation/protocol.py:982-1011."""
    n = len(model._ent_index) if corruption_entities is None else len(corruption_entities)
    if n >= TOO_MANY_ENTITIES_TH:
        warnings.warn(
            "You are attempting to use %d distinct entities to generate synthetic negatives in the evaluation "
            "protocol. This may be unnecessary and will lead to a 'harder' task. Besides, it will lead to a much "
            "slower evaluation procedure." % n)


def evaluate_performance(X, model, filter_triples=None, verbose=False, filter_unseen=True, entities_subset=None,
                         corrupt_side="s,o", ranking_strategy="worst", use_default_protocol=False):
    """Rank every test triple against all-entity corruptions (evaluation/protocol.py:726-979).

    Returns ``ndarray [T]`` or, for ``corrupt_side='s,o'``, ``[T,2]`` (col 0 subject-side rank, col 1
    object-side rank, protocol.py:810-817)."""
    dataset_handle = None
    try:
        if use_default_protocol:
            corrupt_side = "s,o"
        assert corrupt_side in ["s", "o", "s+o", "s,o"], "Invalid value for corrupt_side."
        if isinstance(X, np.ndarray):
            if filter_unseen:
                X = filter_unseen_entities(X, model, verbose=verbose)
            test_idx = to_idx(X, model._ent_index, model._rel_index) if X.shape[0] else np.zeros((0, 3), np.int32)
        elif isinstance(X, EvalDataset):
            dataset_handle = X
            test_idx = None
        else:
            raise ValueError("X must be either a numpy array or an EmgraphBaseDatasetAdaptor.")
        filt_idx = None
        if filter_triples is not None:
            if isinstance(X, EvalDataset):
                # a prepared dataset brings its own filter; filter_triples only switches it on (evaluation/protocol.py:905-915)
                if not isinstance(filter_triples, bool):
                    raise Exception("Expected a boolean type")
                if filter_triples is True:
                    if X.filter_idx is None:
                        raise Exception("Filtered evaluation requested but the dataset holds no filter triples")
                    model.set_filter_for_eval()
            elif isinstance(filter_triples, np.ndarray):
                if filter_unseen:
                    filter_triples = filter_unseen_entities(filter_triples, model, verbose=verbose)
                filt_idx = to_idx(filter_triples, model._ent_index, model._rel_index)
                model.set_filter_for_eval()
            else:
                raise Exception("Invalid datatype for filter. Expected a numpy array or preset data in the adapter.")
        if dataset_handle is None:
            dataset_handle = EvalDataset(test_idx, filt_idx)
        eval_dict = {}
        check_filter_size(model, entities_subset)
        if entities_subset is not None:  # evaluation/protocol.py:940-944
            idx_entities = model._ent_index.lookup_known(np.asarray(entities_subset))
            eval_dict["corruption_entities"] = idx_entities
        eval_dict["corrupt_side"] = corrupt_side
        assert ranking_strategy in ["worst", "best", "middle"], "Invalid ranking_strategy!"
        eval_dict["ranking_strategy"] = ranking_strategy
        model.configure_evaluation_protocol(eval_dict)
        ranks = model.get_ranks(dataset_handle)
        model.end_evaluation()
        return np.array(ranks)
    except BaseException as e:
        model.end_evaluation()
        if dataset_handle is not None:
            dataset_handle.cleanup()
        raise e


# ------------------------------------------------------------------------------------------------
# host-side protocol utilities a reference script imports next to evaluate_performance
# (evaluation/__init__.py:6-16).  The engine itself never materialises corruptions: kge_emit_kernel draws the
# training negatives and the ranking sweeps enumerate the candidates in place; these return the same triples as
# arrays for callers that want to look at them.
# ------------------------------------------------------------------------------------------------
def generate_corruptions_for_eval(X, entities_for_corruption, corrupt_side="s,o"):
    """Every corruption of the triple(s) X over `entities_for_corruption` (evaluation/protocol.py:448-528): for
    's,o' / 's+o' the object sweep [(s, p, e) for all e] followed by the subject sweep [(e, p, o) for all e];
    'o' / 's' one sweep.  int ids in, ndarray [m, 3] out."""
    if corrupt_side == "s,o":
        corrupt_side = "s+o"
    if corrupt_side not in ("s+o", "s", "o"):
        raise ValueError("Invalid argument value for corruption side passed for evaluation")
    X = np.asarray(X).reshape(-1, 3)
    ents = np.asarray(entities_for_corruption).reshape(-1)
    n, m = X.shape[0], ents.shape[0]
    out = []
    if corrupt_side in ("s+o", "o"):
        o_sweep = np.empty((n, m, 3), dtype=X.dtype)
        o_sweep[:, :, 0] = X[:, None, 0]
        o_sweep[:, :, 1] = X[:, None, 1]
        o_sweep[:, :, 2] = ents[None, :]
        out.append(o_sweep.reshape(-1, 3))
    if corrupt_side in ("s+o", "s"):
        s_sweep = np.empty((n, m, 3), dtype=X.dtype)
        s_sweep[:, :, 0] = ents[None, :]
        s_sweep[:, :, 1] = X[:, None, 1]
        s_sweep[:, :, 2] = X[:, None, 2]
        out.append(s_sweep.reshape(-1, 3))
    return np.concatenate(out, axis=0)


def generate_corruptions_for_fit(X, entities_list=None, eta=1, corrupt_side="s,o", entities_size=0, rnd=None):
    """Training corruptions as arrays (evaluation/protocol.py:531-659): X tiled eta times (row j*n+i corrupts
    positive i), per row either the subject or the object replaced -- 's,o' / 's+o': a fair coin per row, 'o' the
    object, 's' the subject -- by a uniform draw from range(entities_size), or from entities_list when given.  No
    check for accidental positives, as in the reference.  `rnd`: seed or numpy RandomState (the reference draws
    from TensorFlow's stream, which no other library reproduces; the engine's own stream is Philox)."""
    if corrupt_side not in ("s+o", "s", "o", "s,o"):
        raise ValueError("Invalid argument value {} for corruption side passed for evaluation.".format(corrupt_side))
    rs = rnd if isinstance(rnd, np.random.RandomState) else np.random.RandomState(rnd)
    X = np.asarray(X).reshape(-1, 3)
    ds = np.tile(X, (eta, 1))
    rows = ds.shape[0]
    if corrupt_side in ("s+o", "s,o"):
        keep_subj = rs.randint(0, 2, size=rows).astype(bool)
    else:
        keep_subj = np.full(rows, corrupt_side == "o")
    if entities_list is not None:
        pool = np.asarray(entities_list).reshape(-1)
        repl = pool[rs.randint(0, pool.shape[0], size=rows)]
    else:
        repl = rs.randint(0, int(entities_size), size=rows)
    out = ds.copy()
    out[:, 0] = np.where(keep_subj, ds[:, 0], repl)
    out[:, 2] = np.where(keep_subj, repl, ds[:, 2])
    return out


_SPLIT_ERROR = ("Cannot create a test split of the desired size. Some entities will not occur in both training and test set. "
                "Set allow_duplication=True,{}remove filter on test predicates or set test_size to a smaller value.")


def _label_codes(labels):
    """Dense integer code per label (equal labels, equal codes) -- the hashing pass of fit()'s id mapping."""
    if labels.shape[0] == 0:
        return np.zeros(0, np.int64)
    return _sorted_factorize(_fixed_width(np.asarray(labels)))[1]


def _split_shuffled(X, test_size, seed, allow_duplication, filtered_test_predicates):
    """evaluation/protocol.py:24-181: walk the candidates in one random order and move a triple to the test set
    whenever every entity and relation of it still occurs elsewhere.  Integer ids and count arrays instead of
    label-keyed dictionaries; the random draws are made in the reference's order from numpy's global stream, so a
    given seed yields the reference's split."""
    np.random.seed(seed)
    if filtered_test_predicates:
        is_cand = np.isin(X[:, 1], filtered_test_predicates)
        cand, fixed_train = X[is_cand], X[~is_cand]
    else:
        cand, fixed_train = X, None
    n = cand.shape[0]
    ent_inv = _label_codes(np.concatenate([cand[:, 0], cand[:, 2]]))
    s_id, o_id = ent_inv[:n], ent_inv[n:]
    p_id = _label_codes(cand[:, 1])
    ent_left = np.bincount(ent_inv).tolist()  # occurrences still in the training part
    rel_left = np.bincount(p_id).tolist()
    s_id, o_id, p_id = s_id.tolist(), o_id.tolist(), p_id.tolist()
    order = np.random.permutation(np.arange(n))
    test_idx, train_idx = [], []
    for at, idx in enumerate(order.tolist()):
        s, p, o = s_id[idx], p_id[idx], o_id[idx]
        ent_left[s] -= 1
        rel_left[p] -= 1
        ent_left[o] -= 1
        if ent_left[s] > 0 and rel_left[p] > 0 and ent_left[o] > 0:
            test_idx.append(idx)
            if len(test_idx) == test_size:
                train_idx.extend(order[at + 1:].tolist())
                break
        else:  # taking it out would leave an entity or a relation unseen: it stays in the training set
            ent_left[s] += 1
            rel_left[p] += 1
            ent_left[o] += 1
            train_idx.append(idx)
    if len(test_idx) != test_size:
        if not allow_duplication:
            raise Exception(_SPLIT_ERROR.format(""))
        test_idx.extend(np.random.choice(test_idx, size=(test_size - len(test_idx))).tolist())
    X_train = cand[train_idx] if fixed_train is None else np.concatenate([fixed_train, cand[train_idx]])
    X_test = cand[test_idx]
    return np.random.permutation(X_train), np.random.permutation(X_test)


def _split_random_search(X, test_size, seed, allow_duplication, filtered_test_predicates):
    """evaluation/protocol.py:184-320 (backward_compatible=True): draw candidates one at a time from
    RandomState(seed); a drawn triple joins the test set when its subject, object and relation each still occur
    more than once in their own role."""
    rnd = np.random.RandomState(seed)
    s_id, o_id, p_id = _label_codes(X[:, 0]), _label_codes(X[:, 2]), _label_codes(X[:, 1])
    s_left, o_left, p_left = np.bincount(s_id), np.bincount(o_id), np.bincount(p_id)
    pool = np.where(np.isin(X[:, 1], filtered_test_predicates))[0] if filtered_test_predicates else np.arange(len(X))
    chosen = []
    taken = set()
    budget = len(X) * 10
    tries = 0
    while len(chosen) < test_size:
        i = int(rnd.choice(pool))
        s, p, o = s_id[i], p_id[i], o_id[i]
        if s_left[s] > 1 and o_left[o] > 1 and p_left[p] > 1:
            s_left[s] -= 1
            o_left[o] -= 1
            p_left[p] -= 1
            if allow_duplication:
                chosen.append(i)
            elif i not in taken:
                taken.add(i)
                chosen.append(i)
        tries += 1
        if tries == budget:
            raise Exception(_SPLIT_ERROR.format("change seed values, ") if not allow_duplication else
                            "Cannot create a test split of the desired size. Some entities will not occur in both training and "
                            "test set. Change seed values, remove filter on test predicates or set test_size to a smaller value.")
    idx_test = np.asarray(chosen, dtype=int) if allow_duplication else np.unique(np.asarray(chosen, dtype=int))
    idx_train = np.setdiff1d(np.arange(len(X)), idx_test)
    return X[idx_train, :], X[idx_test, :]


def train_test_split_no_unseen(X, test_size=100, seed=0, allow_duplication=False, filtered_test_predicates=None,
                               backward_compatible=False):
    """Split X [n, 3] into (train, test) such that every entity and relation of the test set also occurs in the
    training set (evaluation/protocol.py:323-407).  test_size: a count, or a fraction of len(X) when float."""
    X = np.asarray(X)
    if type(test_size) is float:
        test_size = int(len(X) * test_size)
    split = _split_random_search if backward_compatible else _split_shuffled
    return split(X, test_size, seed, allow_duplication, filtered_test_predicates)


# ------------------------------------------------------------------------------------------------
# metrics (evaluation/metrics.py:11-67, :70-130, :133-164, :167-222)
# ------------------------------------------------------------------------------------------------
def hits_at_n_score(ranks, n):
    if isinstance(ranks, list):
        ranks = np.asarray(ranks)
    ranks = ranks.reshape(-1)
    return np.sum(ranks <= n) / len(ranks)


def mrr_score(ranks):
    if isinstance(ranks, list):
        ranks = np.asarray(ranks)
    ranks = ranks.reshape(-1)
    return np.sum(1 / ranks) / len(ranks)


def mr_score(ranks):
    if isinstance(ranks, list):
        ranks = np.asarray(ranks)
    ranks = ranks.reshape(-1)
    return np.sum(ranks) / len(ranks)


def rank_score(y_true, y_pred, pos_lab=1):
    idx = np.argsort(y_pred)[::-1]
    y_ord = np.asarray(y_true)[idx]
    return int(np.where(y_ord == pos_lab)[0][0] + 1)
