"""Host-side mirror of ``emgraph.evaluation`` for the hot path: ``evaluate_performance`` (filtered
ranking, evaluation/protocol.py:726-979) and the rank metrics (evaluation/metrics.py).  The filter
index is built on the device (replaces the temp-file SQLite DB of datasets/sqlite_adapter.py)."""
from __future__ import annotations

import warnings

import numpy as np

from .engine import to_dev_i32
from .models import EmbeddingModel, create_mappings, to_idx  # noqa: F401  (re-exported like the reference)

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
    keep = model._ent_index.contains(X[:, 0]) & model._ent_index.contains(X[:, 2])
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
            if isinstance(filter_triples, np.ndarray):
                if filter_unseen:
                    filter_triples = filter_unseen_entities(filter_triples, model, verbose=verbose)
                filt_idx = to_idx(filter_triples, model._ent_index, model._rel_index)
                model.set_filter_for_eval()
            elif isinstance(X, EvalDataset):
                if not isinstance(filter_triples, bool):
                    raise Exception("Expected a boolean type")
                if filter_triples is True:
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
