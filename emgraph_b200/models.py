"""Host-side mirror of ``emgraph.models`` for the hot path: TransE / DistMult / ComplEx / HolE with
the reference constructor signature, ``fit`` / ``predict`` / ``get_embeddings`` and the four methods
``evaluate_performance`` calls.  All arithmetic runs in libkge_b200.so (CUDA, sm_100a).

Reference interface mirrored (paths relative to the reference root):
  models/EmbeddingModel.py:184-314 (ctor, registries, error conventions), :1113-1492 (fit),
  :2101-2147 (predict), :455-488 (get_embeddings), :1494-1518, :2035-2099 (evaluation hooks);
  models/TransE.py:142-165, DistMult.py:50-71, ComplEx.py:93-113, HolE.py:37-57 (signatures);
  utils/constants.py (defaults).
Documented deviations: every batch is trained once and its loss comes from
the same pass (F6); optimizer state persists across batches unless ``engine_params
['reset_state']`` (F5); ``get_ranks`` ranks every test triple (F3).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .optimizers import SGDSchedule
from .engine import get_engine, internal_k, model_id, to_dev_i32

# utils/constants.py
DEFAULT_EMBEDDING_SIZE = 100
DEFAULT_ETA = 2
DEFAULT_EPOCH = 100
DEFAULT_BATCH_COUNT = 100
DEFAULT_SEED = 0
DEFAULT_OPTIM = "adam"
DEFAULT_LOSS = "nll"
DEFAULT_LR = 0.0005
DEFAULT_MOMENTUM = 0.9
DEFAULT_REGULARIZER = None
DEFAULT_INITIALIZER = "glorot_uniform"
DEFAULT_VERBOSE = False
DEFAULT_NORM_TRANSE = 1
DEFAULT_CORRUPT_SIDE_TRAIN = ["s,o"]
DEFAULT_CORRUPT_SIDE_EVAL = "s,o"
DEFAULT_CORRUPTION_ENTITIES = "all"
DEFAULT_RANK_COMPARE_STRATEGY = "worst"
DEFAULT_MARGIN = 1  # losses/_loss_constants.py:8

MODEL_REGISTRY = {}

# models/EmbeddingModel.py:37-58: number of entities above which the reference pages the table through host memory
# ("large graph mode", SGD only).  The table lives in HBM here (sharded over GPUs when it must be, DESIGN.md
# section 7), so the threshold changes nothing; the two functions exist so that scripts calling them keep running.
ENTITY_THRESHOLD = 5e5


def set_entity_threshold(threshold):
    global ENTITY_THRESHOLD
    ENTITY_THRESHOLD = threshold


def reset_entity_threshold():
    global ENTITY_THRESHOLD
    ENTITY_THRESHOLD = 5e5

DEFAULT_BURN_IN_EARLY_STOPPING = 100  # utils/constants.py:12
DEFAULT_CHECK_INTERVAL_EARLY_STOPPING = 10  # utils/constants.py:15
DEFAULT_STOP_INTERVAL_EARLY_STOPPING = 3  # utils/constants.py:18
DEFAULT_CRITERIA_EARLY_STOPPING = "mrr"  # utils/constants.py:21
DEFAULT_MARGIN_ADVERSARIAL = 3  # losses/_loss_constants.py:12
DEFAULT_ALPHA_ADVERSARIAL = 0.5  # losses/_loss_constants.py:10
SUPPORTED_LOSSES = ("pairwise", "nll", "multiclass_nll", "absolute_margin", "self_adversarial")
TILED_POSITIVE_LOSSES = ("pairwise", "nll", "absolute_margin")  # require_same_size_pos_neg (losses/utils.py:24)
SUPPORTED_OPTIMIZERS = ("adam", "adagrad", "momentum", "sgd")
SUPPORTED_INITIALIZERS = ("glorot_uniform", "xavier", "normal", "uniform", "constant")
SUPPORTED_REGULARIZERS = ("LP",)  # regularizers/lp.py
DEFAULT_LAMBDA = 1e-5  # regularizers/_regularizer_constants.py:8
DEFAULT_REG_NORM = 2  # regularizers/_regularizer_constants.py:10


def register_model(name):
    def deco(cls):
        MODEL_REGISTRY[name] = cls
        cls.name = name
        return cls
    return deco


# ------------------------------------------------------------------------------------------------
# id mapping (evaluation/protocol.py:429-445, :662-723) -- same sorted-unique ids, vectorised
# ------------------------------------------------------------------------------------------------
class LabelIndex:
    """label -> id with ids assigned in np.unique (sorted) order; dict view built lazily."""

    def __init__(self, labels_sorted):
        self.labels = np.asarray(labels_sorted)
        self._dict = None
        self._hash = None  # lazily: hash table over the labels for bulk lookups

    def __len__(self):
        return len(self.labels)

    _HASH_MIN = 4096  # below this many lookups the binary search over the labels is as fast

    def _hashed(self):
        """(hashes of the labels in id order, pandas Index over them or None): built once, None when hashing does not
        apply (labels not fixed-width str / bytes, or two labels share a hash)."""
        if self._hash is None:
            self._hash = False
            if self.labels.ndim == 1 and self.labels.dtype.kind in "US" and len(self.labels) > 0:
                h = _hash_fixed_width(self.labels)
                if np.unique(h).shape[0] == h.shape[0]:
                    try:
                        import pandas as pd
                        self._hash = (h, pd.Index(h))
                    except ImportError:
                        order = np.argsort(h, kind="stable")
                        self._hash = (h, (h[order], order))
        return self._hash or None

    def _positions(self, x):
        """(candidate id of every element of x, whether the label there IS that element).  Large fixed-width str / bytes
        inputs are hashed (one pass, then an integer table) instead of binary-searched label by label; the label
        comparison at the end makes the result exact either way."""
        n = len(self.labels)
        if self.labels.dtype.kind in "US":
            x = _fixed_width(x)
        if x.size >= self._HASH_MIN and x.ndim == 1 and x.dtype.kind == self.labels.dtype.kind and x.dtype.kind in "US":
            tab = self._hashed()
            if tab is not None:
                hx = _hash_fixed_width(x if x.dtype == self.labels.dtype else x.astype(self.labels.dtype))
                if isinstance(tab[1], tuple):
                    hs, order = tab[1]
                    p = np.minimum(np.searchsorted(hs, hx), n - 1)
                    pos = order[p]
                else:
                    pos = tab[1].get_indexer(hx)
                    pos = np.where(pos < 0, 0, pos)
                return pos, self.labels[pos] == x
        pos = np.minimum(np.searchsorted(self.labels, x), n - 1)
        return pos, self.labels[pos] == x

    def lookup(self, x, what):
        x = np.asarray(x)
        if len(self.labels) == 0:
            raise ValueError(_UNSEEN_MSG.format(concept_type=what))
        try:
            pos, ok = self._positions(x)
        except TypeError:
            raise ValueError(_UNSEEN_MSG.format(concept_type="concepts"))
        if not np.all(ok):
            raise ValueError(_UNSEEN_MSG.format(concept_type=what))
        return pos.astype(np.int32)

    def contains(self, x):
        x = np.asarray(x)
        if len(self.labels) == 0:
            return np.zeros(x.shape, bool)
        try:
            return self._positions(x)[1]
        except TypeError:
            return np.zeros(x.shape, bool)

    def lookup_known(self, x):
        """ids of the labels in x the index knows, sorted by id; unknown labels are dropped (the reference's
        ``[idx for uri, idx in ent_to_idx.items() if uri in x]``, evaluation/protocol.py:941-944)."""
        x = np.asarray(x).reshape(-1)
        if x.size == 0 or len(self.labels) == 0:
            return np.zeros(0, np.int32)
        x = x[self.contains(x)]
        return np.unique(self.lookup(x, "entities")).astype(np.int32)

    def as_dict(self):
        if self._dict is None:
            self._dict = dict(zip(self.labels.tolist(), range(len(self.labels))))
        return self._dict


_UNSEEN_MSG = (
    "Input triples include one or more {concept_type} not present in the training set. "
    "Please filter all concepts in X that do not occur in the training test "
    "(set filter_unseen=True in evaluate_performance) or retrain the model on a "
    "training set that includes all the desired concept types."
)


def create_mappings(X):
    """evaluation/protocol.py:429-445 -> (rel_to_idx, ent_to_idx) dicts."""
    X = np.asarray(X)
    ent = LabelIndex(np.unique(np.concatenate((X[:, 0], X[:, 2]))))
    rel = LabelIndex(np.unique(X[:, 1]))
    return rel.as_dict(), ent.as_dict()


def to_idx(X, ent_to_idx, rel_to_idx):
    """evaluation/protocol.py:706-723; unseen label -> ValueError."""
    X = np.asarray(X)
    if X.ndim == 1:
        X = X[np.newaxis, :]
    ei = ent_to_idx if isinstance(ent_to_idx, LabelIndex) else _index_from_dict(ent_to_idx)
    ri = rel_to_idx if isinstance(rel_to_idx, LabelIndex) else _index_from_dict(rel_to_idx)
    s = ei.lookup(X[:, 0], "entities")
    p = ri.lookup(X[:, 1], "relations")
    o = ei.lookup(X[:, 2], "entities")
    return np.stack([s, p, o], axis=1)


def _fixed_width(x):
    """Object arrays whose elements are all `str` (what `DataFrame.values` / `read_csv` hand over) as fixed-width
    unicode arrays: same labels, same (code-point) order, but sortable and hashable without Python-level comparisons.
    Anything else is returned unchanged."""
    if x.dtype != object or x.size == 0:
        return x
    try:
        import pandas as pd
        if pd.api.types.infer_dtype(x.reshape(-1), skipna=False) == "string":
            return x.astype(str)
    except ImportError:
        pass
    return x


def _hash_fixed_width(a):
    """64-bit FNV-1a-style hash of every element of a 1-D fixed-width str / bytes array, eight bytes at a time."""
    unit = 4 if a.dtype.kind == "U" else 1
    chars = a.dtype.itemsize // unit
    if chars == 0:
        return np.zeros(a.shape[0], np.uint64)
    if (chars * unit) % 8:  # pad the items to whole 8-byte words (zero padding, like the dtype's own)
        chars += (8 - (chars * unit) % 8) // unit
        a = a.astype("%s%d" % ("<U" if unit == 4 else "S", chars))
    lanes = np.ascontiguousarray(a).view(np.uint64).reshape(a.shape[0], -1)
    h = np.full(a.shape[0], 0xCBF29CE484222325, np.uint64)
    for j in range(lanes.shape[1]):
        h ^= lanes[:, j]
        h *= np.uint64(0x100000001B3)
        h ^= h >> np.uint64(29)
    return h


def _sorted_factorize(values):
    """(sorted unique labels, code of every value in that order) == np.unique(values, return_inverse=True).

    Fixed-width str / bytes arrays (what np.loadtxt / np.array of labels give) take a hashing pass instead of a sort of
    all n labels: 64-bit hashes -> integer factorisation -> one representative label per hash, VERIFIED against every
    input (a hash collision falls back to the sort) -> only the distinct labels are sorted."""
    values = np.asarray(values)
    n = values.shape[0]
    if values.ndim == 1 and values.dtype.kind in "US" and n > 0:
        h = _hash_fixed_width(values)
        try:
            import pandas as pd
            codes, _ = pd.factorize(h)  # codes in order of first appearance
        except ImportError:
            _, first_u, inv = np.unique(h, return_index=True, return_inverse=True)
            order_fa = np.argsort(first_u, kind="stable")  # re-number in order of first appearance
            renum = np.empty(order_fa.shape[0], np.int64)
            renum[order_fa] = np.arange(order_fa.shape[0])
            codes = renum[inv.reshape(-1)]
        n_codes = int(codes.max()) + 1
        first = np.empty(n_codes, np.int64)
        first[codes[::-1]] = np.arange(n - 1, -1, -1)  # the last write wins: position of the first occurrence
        reps = values[first]
        if np.array_equal(reps[codes], values):  # no two different labels share a hash
            order = np.argsort(reps, kind="stable")
            rank = np.empty(n_codes, np.int64)
            rank[order] = np.arange(n_codes)
            return reps[order], rank[codes]
    uniq, inv = np.unique(values, return_inverse=True)
    return uniq, inv.reshape(-1)


def index_training_triples(X):
    """create_mappings + to_idx of a training set (evaluation/protocol.py:429-445, :662-723) in one pass: the ids are the
    positions of the labels in np.unique (sorted) order, exactly as the reference assigns them, but every label is
    hashed once instead of being binary-searched (or looked up through np.vectorize(dict.get)) after a full sort.
    Returns (entity LabelIndex, relation LabelIndex, ids [n,3] int32)."""
    X = np.asarray(X)
    n = X.shape[0]
    ent_labels, ent_codes = _sorted_factorize(_fixed_width(np.concatenate((X[:, 0], X[:, 2]))))
    rel_labels, rel_codes = _sorted_factorize(_fixed_width(X[:, 1]))
    Xi = np.stack([ent_codes[:n], rel_codes, ent_codes[n:]], axis=1).astype(np.int32)
    return LabelIndex(ent_labels), LabelIndex(rel_labels), Xi


def _index_from_dict(d):
    keys = np.asarray(list(d.keys()))
    vals = np.asarray(list(d.values()))
    order = np.argsort(vals)
    li = LabelIndex(keys[order])
    if not np.array_equal(vals[order], np.arange(len(vals))) or np.any(li.labels[:-1] > li.labels[1:]):
        # arbitrary user mapping: fall back to an explicit permutation
        srt = np.argsort(keys)
        li = _PermutedIndex(keys[srt], vals[srt])
    return li


class _PermutedIndex(LabelIndex):
    def __init__(self, labels_sorted, ids):
        super().__init__(labels_sorted)
        self.ids = np.asarray(ids)

    def lookup(self, x, what):
        return self.ids[super().lookup(x, what)].astype(np.int32)


# ------------------------------------------------------------------------------------------------
# model base
# ------------------------------------------------------------------------------------------------
class EmbeddingModel:
    name = None

    def __init__(self, k=DEFAULT_EMBEDDING_SIZE, eta=DEFAULT_ETA, epochs=DEFAULT_EPOCH,
                 batches_count=DEFAULT_BATCH_COUNT, seed=DEFAULT_SEED, embedding_model_params=None,
                 optimizer=DEFAULT_OPTIM, optimizer_params=None, loss=DEFAULT_LOSS, loss_params=None,
                 regularizer=DEFAULT_REGULARIZER, regularizer_params=None, initializer=DEFAULT_INITIALIZER,
                 initializer_params=None, large_graphs=False, verbose=DEFAULT_VERBOSE, engine_params=None):
        embedding_model_params = {} if embedding_model_params is None else embedding_model_params
        optimizer_params = {"lr": DEFAULT_LR} if optimizer_params is None else optimizer_params
        loss_params = {} if loss_params is None else loss_params
        regularizer_params = {} if regularizer_params is None else regularizer_params
        initializer_params = {"uniform": False} if initializer_params is None else initializer_params
        if loss == "bce":  # models/EmbeddingModel.py:206-210
            raise ValueError("Invalid Model - Loss combination. ConvE model can be used with BCE loss only and vice versa.")
        self.all_params = {
            "k": k, "eta": eta, "epochs": epochs, "batches_count": batches_count, "seed": seed,
            "embedding_model_params": embedding_model_params, "optimizer": optimizer,
            "optimizer_params": optimizer_params, "loss": loss, "loss_params": loss_params,
            "regularizer": regularizer, "regularizer_params": regularizer_params,
            "initializer": initializer, "initializer_params": initializer_params, "verbose": verbose,
        }
        self.seed = seed
        self.rnd = np.random.RandomState(seed)
        self.k = k
        self.internal_k = internal_k(self.name, k)
        self.epochs = epochs
        self.eta = eta
        self.batches_count = batches_count
        self.embedding_model_params = embedding_model_params
        self.loss_params = loss_params
        self.optimizer_params = optimizer_params
        self.regularizer_params = regularizer_params
        self.initializer_params = initializer_params
        self.engine_params = dict(engine_params or {})
        self.dealing_with_large_graphs = large_graphs  # accepted, ignored: the table lives in HBM
        self.verbose = verbose
        if loss not in SUPPORTED_LOSSES:
            raise ValueError("Unsupported loss function: {}".format(loss))
        if regularizer is not None and regularizer not in SUPPORTED_REGULARIZERS:
            raise ValueError("Unsupported regularizer: {}".format(regularizer))
        self._reg = self._parse_regularizer(regularizer, regularizer_params)
        if optimizer not in SUPPORTED_OPTIMIZERS:
            raise ValueError("Unsupported optimizer: {}".format(optimizer))
        if initializer not in SUPPORTED_INITIALIZERS:
            raise ValueError("Unsupported initializer: {}".format(initializer))
        self.loss = loss
        self.optimizer = optimizer
        self.regularizer = regularizer
        self.initializer = initializer
        self.trained_model_params = []
        self.is_fitted = False
        self.is_filtered = False
        self.eval_config = {}
        self.eval_dataset_handle = None
        self.is_calibrated = False
        self.calibration_parameters = []
        self._ent_index = None
        self._rel_index = None
        self._dev = None  # device-resident parameters {ent, rel}
        self.loss_history = []

    # ---- mappings exposed like the reference attributes (utils/model_utils.py:63-72 reads them)
    @property
    def ent_to_idx(self):
        return {} if self._ent_index is None else self._ent_index.as_dict()

    @ent_to_idx.setter
    def ent_to_idx(self, d):
        self._ent_index = _index_from_dict(d)

    @property
    def rel_to_idx(self):
        return {} if self._rel_index is None else self._rel_index.as_dict()

    @rel_to_idx.setter
    def rel_to_idx(self, d):
        self._rel_index = _index_from_dict(d)

    @staticmethod
    def _parse_regularizer(regularizer, params):
        """LP regulariser hyper-parameters (regularizers/lp.py:41-104): p int, lambda scalar or [ent, rel]."""
        if regularizer is None:
            return dict(reg_p=0, reg_lambda_ent=0.0, reg_lambda_rel=0.0)
        p = params.get("p", DEFAULT_REG_NORM)
        if not isinstance(p, (int, np.integer)):
            raise Exception("Invalid value for regularizer parameter p:{}. Supported type int, np.int32 or np.int64".format(p))
        lam = params.get("lambda", DEFAULT_LAMBDA)
        if np.isscalar(lam):
            lam = [lam, lam]
        elif not (isinstance(lam, list) and len(lam) == 2):
            raise ValueError("Regularizer weight must be a scalar or a list with length equal to number of params passes")
        return dict(reg_p=int(p), reg_lambda_ent=float(lam[0]), reg_lambda_rel=float(lam[1]))

    def get_embedding_model_params(self, output_dict):  # models/EmbeddingModel.py:347-358
        output_dict["model_params"] = [np.asarray(p) for p in self.trained_model_params]
        output_dict["large_graph"] = self.dealing_with_large_graphs
        output_dict["calibration_parameters"] = self.calibration_parameters

    def restore_model_params(self, in_dict):  # models/EmbeddingModel.py:360-383
        self.trained_model_params = [np.asarray(p, dtype=np.float32) for p in in_dict["model_params"]]
        self.calibration_parameters = in_dict.get("calibration_parameters", [])
        self.dealing_with_large_graphs = in_dict.get("large_graph", False)
        self._dev = None  # uploaded lazily on the first predict / get_ranks
        st = in_dict.get("b200_optimizer_state")
        if st:
            self._opt_state = {k_: torch.from_numpy(np.ascontiguousarray(v)) for k_, v in st.items()}
            self._opt_step = int(in_dict.get("b200_optimizer_step", 0))

    def is_fitted_on(self, X):  # models/EmbeddingModel.py:2188-2210
        """Heuristic of the reference: same NUMBER of distinct entities and relations as the training set."""
        if not self.is_fitted:
            raise RuntimeError("Model has not been fitted.")
        X = np.asarray(X)
        unique_ent = np.unique(np.concatenate((X[:, 0], X[:, 2])))
        unique_rel = np.unique(X[:, 1])
        return len(unique_ent) == len(self._ent_index) and len(unique_rel) == len(self._rel_index)

    def get_hyperparameter_dict(self):  # models/EmbeddingModel.py:338
        return self.all_params

    # ---- ids
    def _model_id(self):
        return model_id(self.name, int(self.embedding_model_params.get("norm", DEFAULT_NORM_TRANSE)))

    def _non_linearity(self):  # models/EmbeddingModel.py:679-689
        nl = self.embedding_model_params.get("non_linearity", "linear")
        if nl not in _lib.NL_IDS:
            raise ValueError("Invalid non-linearity")
        return _lib.NL_IDS[nl]

    def _train_sides(self):
        # the reference reads 'corrupt_side' though ctors document 'corrupt_sides' (SURVEY F9): accept both
        sides = self.embedding_model_params.get("corrupt_side", self.embedding_model_params.get("corrupt_sides", DEFAULT_CORRUPT_SIDE_TRAIN))
        if not isinstance(sides, list):
            sides = [sides]
        for s in sides:
            if s not in _lib.TRAIN_SIDE_IDS:
                raise ValueError("Invalid argument value {} for corruption side passed for evaluation.".format(s))
        return sides

    # ---- parameter initialisation (models/EmbeddingModel.py:547-601, initializers/*.py)
    def _init_table(self, rows, cols, which):
        ini, p = self.initializer, self.initializer_params
        if ini in ("glorot_uniform", "xavier"):
            if p.get("uniform", False) or True:  # TF path always uniform
                lim = np.sqrt(6.0 / (rows + cols))
                return self.rnd.uniform(-lim, lim, size=(rows, cols)).astype(np.float32)
        if ini == "normal":
            return self.rnd.normal(p.get("mean", 0), p.get("std", 0.05), size=(rows, cols)).astype(np.float32)
        if ini == "uniform":
            return self.rnd.uniform(p.get("low", -0.05), p.get("high", 0.05), size=(rows, cols)).astype(np.float32)
        if ini == "constant":
            key = "entity" if which == "entity" else "relation"
            try:
                arr = np.asarray(p[key], dtype=np.float32)
            except KeyError:
                raise Exception("Initial {} value not passed to the initializer!".format(key))
            assert arr.shape == (rows, cols), "Invalid shape for {} initializer!".format(key)
            return np.ascontiguousarray(arr)
        raise ValueError("Unsupported initializer: {}".format(ini))

    # ---- training (models/EmbeddingModel.py:1113-1492)
    def fit(self, X, early_stopping=False, early_stopping_params={}, focusE_numeric_edge_values=None,
            tensorboard_logs_path=None):
        if not isinstance(X, np.ndarray):
            raise ValueError("Invalid type for input X. Expected ndarray/EmgraphDataset object, got {}".format(type(X)))
        if X.ndim != 2 or X.shape[1] != 3:
            raise ValueError("Invalid size for input X. Expected (n,3):  got {}".format(X.shape))
        if focusE_numeric_edge_values is not None:
            raise NotImplementedError("FocusE edge weights are outside the B200 hot-path scope")
        ent_index, rel_index, Xi = index_training_triples(X)
        # engine_params['resume']: continue from the current parameters, the saved sparse-optimizer state and the
        # global step (no reference counterpart: its optimizers are re-created every batch)
        self._resume = bool(self.engine_params.get("resume", False)) and self.is_fitted
        if self._resume and not (type(self._ent_index) is LabelIndex and type(self._rel_index) is LabelIndex  # ids in sorted order
                                 and np.array_equal(self._ent_index.labels, ent_index.labels)
                                 and np.array_equal(self._rel_index.labels, rel_index.labels)):
            self._resume = False
            raise ValueError("resume needs the entities and relations the model was fitted on")
        self._ent_index, self._rel_index = ent_index, rel_index
        self.early_stopping_params = early_stopping_params
        try:
            self._fit_idx(Xi, len(self._ent_index), len(self._rel_index), early_stopping=early_stopping)
        finally:
            self._after_training()
        return self

    def _after_training(self):  # models/EmbeddingModel.py:1022-1042
        if self.eval_dataset_handle is not None:
            self.eval_dataset_handle.cleanup()
            self.eval_dataset_handle = None
        self.is_filtered = False
        self.eval_config = {}

    # ---- early stopping on the ranking kernel (models/EmbeddingModel.py:824-1020)
    def _initialize_early_stopping(self):
        from .evaluation import EvalDataset  # late import: evaluation imports this module
        try:
            x_valid = self.early_stopping_params["x_valid"]
        except KeyError:
            raise KeyError("x_valid must be passed for early fitting.")
        if isinstance(x_valid, np.ndarray):
            if x_valid.ndim <= 1 or np.shape(x_valid)[1] != 3:
                raise ValueError("Invalid size for input x_valid. Expected (n,3):  got {}".format(np.shape(x_valid)))
            valid_idx = to_idx(x_valid, self._ent_index, self._rel_index)
        elif isinstance(x_valid, EvalDataset):
            valid_idx = x_valid.test_idx
        else:
            raise ValueError("Invalid type for input X. Expected ndarray/EmgraphDataset object, got {}".format(type(x_valid)))
        self.early_stopping_criteria = self.early_stopping_params.get("criteria", DEFAULT_CRITERIA_EARLY_STOPPING)
        if self.early_stopping_criteria not in ["hits10", "hits1", "hits3", "mrr"]:
            raise ValueError("Unsupported early stopping criteria.")
        ce = self.early_stopping_params.get("corruption_entities", DEFAULT_CORRUPTION_ENTITIES)
        if isinstance(ce, list):  # raw labels -> entity ids (:872-884)
            ce = self._ent_index.lookup_known(np.asarray(ce))
        elif isinstance(ce, str) and ce not in ("all", "batch"):
            raise ValueError("Invalid value for corruption_entities: {}".format(ce))
        self.eval_config["corruption_entities"] = ce
        self.eval_config["corrupt_side"] = self.early_stopping_params.get("corrupt_side", DEFAULT_CORRUPT_SIDE_EVAL)
        self.early_stopping_best_value = None
        self.early_stopping_stop_counter = 0
        self.early_stopping_epoch = None
        filt_idx = None
        if "x_filter" in self.early_stopping_params:
            x_filter = self.early_stopping_params["x_filter"]
            if isinstance(x_filter, np.ndarray):
                if x_filter.ndim <= 1 or np.shape(x_filter)[1] != 3:
                    raise ValueError("Invalid size for input x_valid. Expected (n,3):  got {}".format(np.shape(x_filter)))
                filt_idx = to_idx(x_filter, self._ent_index, self._rel_index)
            elif isinstance(x_valid, EvalDataset):
                filt_idx = x_valid.filter_idx
            self.set_filter_for_eval()
        self.eval_dataset_handle = EvalDataset(valid_idx, filt_idx)

    def _save_trained_params(self):  # models/EmbeddingModel.py:385-401 (device snapshot of the best parameters)
        f = self._fit
        self._best_params = (f["ent"].clone(), f["rel"].clone())

    def _perform_early_stopping_test(self, epoch):
        p = self.early_stopping_params
        if not (epoch >= p.get("burn_in", DEFAULT_BURN_IN_EARLY_STOPPING)
                and epoch % p.get("check_interval", DEFAULT_CHECK_INTERVAL_EARLY_STOPPING) == 0):
            return False
        from .evaluation import hits_at_n_score, mrr_score
        f = self._fit
        ranks = self._rank_with_params(f["eng"], f["ent"], f["rel"], self.eval_dataset_handle)
        crit = self.early_stopping_criteria
        current = mrr_score(ranks) if crit == "mrr" else hits_at_n_score(ranks, int(crit[4:]))
        self.early_stopping_history.append((epoch, float(current)))
        if self.early_stopping_best_value is None:  # first validation
            self.early_stopping_best_value = current
            self.early_stopping_first_value = current
        elif self.early_stopping_best_value >= current:
            self.early_stopping_stop_counter += 1
            if self.early_stopping_stop_counter == p.get("stop_interval", DEFAULT_STOP_INTERVAL_EARLY_STOPPING):
                # the best value never moved from the first one: keep the current parameters (:991-996)
                if self.early_stopping_best_value == self.early_stopping_first_value:
                    self._save_trained_params()
                if self.verbose:
                    print("Early stopping at epoch:{}".format(epoch))
                    print("Best {}: {:10f}".format(crit, self.early_stopping_best_value))
                self.early_stopping_epoch = epoch
                return True
        else:
            self.early_stopping_best_value = current
            self.early_stopping_stop_counter = 0
            self._save_trained_params()
        return False

    def _reseed_if_refit(self):
        """Re-fitting a fitted model starts from the seed again (models/EmbeddingModel.py:1285-1290), so that
        fit(X); fit(X) reproduces the first fit (reference tests/emgraph/models/test_models.py:338-367) and
        select_best_model_ranking(retrain_best_model=True) retrains deterministically."""
        if self.is_fitted:
            self.rnd = np.random.RandomState(self.seed)

    def _fit_prepare(self, E, R):
        """Allocate device parameters + optimizer state and freeze the per-step arguments."""
        self._reseed_if_refit()
        eng = get_engine(self.engine_params.get("device"))
        dev = eng.tdev
        K = self.internal_k
        resume = bool(getattr(self, "_resume", False))
        if resume:
            ent = torch.from_numpy(np.ascontiguousarray(self.trained_model_params[0], dtype=np.float32)).to(dev)
            rel = torch.from_numpy(np.ascontiguousarray(self.trained_model_params[1], dtype=np.float32)).to(dev)
            if tuple(ent.shape) != (E, K) or tuple(rel.shape) != (R, K):
                raise ValueError("resume: parameter shapes {} / {} do not fit this model".format(tuple(ent.shape), tuple(rel.shape)))
        else:
            ent = torch.from_numpy(self._init_table(E, K, "entity")).to(dev)
            rel = torch.from_numpy(self._init_table(R, K, "relation")).to(dev)
        opt = _lib.OPT_IDS[self.optimizer]
        reset = bool(self.engine_params.get("reset_state", False))
        st = {}
        if not reset:
            if opt == 0:
                st = dict(ent_m=torch.zeros_like(ent), ent_v=torch.zeros_like(ent), rel_m=torch.zeros_like(rel), rel_v=torch.zeros_like(rel))
            elif opt == 1:
                st = dict(ent_m=torch.full_like(ent, 0.1), rel_m=torch.full_like(rel, 0.1))
            elif opt == 2:
                st = dict(ent_m=torch.zeros_like(ent), rel_m=torch.zeros_like(rel))
            saved = getattr(self, "_opt_state", None) or {}
            if resume and st and set(saved) == set(st) and all(tuple(saved[k_].shape) == tuple(st[k_].shape) for k_ in st):
                st = {k_: saved[k_].detach().clone().to(dev) for k_ in st}  # rows' m / v / accumulator as they were left
        step0 = int(getattr(self, "_opt_step", 0)) if resume else 0
        # embedding_model_params['negative_corruption_entities'] (models/EmbeddingModel.py:732-777)
        nce = self.embedding_model_params.get("negative_corruption_entities", DEFAULT_CORRUPTION_ENTITIES)
        neg = {}
        self._neg_batch = False
        if isinstance(nce, str):
            if nce == "batch":
                self._neg_batch = True
            elif nce != "all":
                raise ValueError("Invalid value for negative_corruption_entities: {}".format(nce))
        elif isinstance(nce, list):
            ids = self._ent_index.lookup_known(np.asarray(nce)) if self._ent_index is not None else np.asarray(nce)
            if len(ids) == 0:
                raise ValueError("negative_corruption_entities: none of the supplied entities is known to the model")
            neg["neg_entities"] = to_dev_i32(np.asarray(ids, dtype=np.int32), dev)
        elif isinstance(nce, (int, np.integer)):
            if not 0 < int(nce) <= E:
                raise ValueError("negative_corruption_entities must be in (0, {}]".format(E))
            neg["neg_entities_n"] = int(nce)
        else:
            raise ValueError("Invalid type for negative_corruption_entities: {}".format(type(nce)))
        torch.cuda.synchronize(dev)  # parameters, state and entity lists are resident before the first step
        self._fit = dict(
            eng=eng, ent=ent, rel=rel, st=st, step=step0, sides=self._train_sides(), neg=neg,
            pipeline=bool(self.engine_params.get("pipeline", True)),
            loss_dev=torch.zeros(1, dtype=torch.float32, device=dev),
            loss_host=torch.zeros(1, dtype=torch.float32).pin_memory(),
            kw=dict(model=self._model_id(), loss=_lib.LOSS_IDS[self.loss], opt=opt, k=self.k, eta=self.eta,
                    flags=_lib.F_RESET_STATE if reset else 0, margin=float(self.loss_params.get("margin", DEFAULT_MARGIN_ADVERSARIAL if self.loss == "self_adversarial" else DEFAULT_MARGIN)),
                    alpha=float(self.loss_params.get("alpha", DEFAULT_ALPHA_ADVERSARIAL)), **self._reg,
                    non_linearity=self._non_linearity(),
                    lr=float(self.optimizer_params.get("lr", DEFAULT_LR)),
                    momentum=float(self.optimizer_params.get("momentum", DEFAULT_MOMENTUM)), seed=int(self.seed)))
        return self._fit

    # A list-valued corrupt_side (models/EmbeddingModel.py:780-816) sums one loss term per side, each over fresh
    # corruptions of the same positives, into ONE optimizer step.  Every loss is a sum of per-positive terms,
    # so the step runs as one batch that holds the positives once per side (n' = S*n) with the side of every
    # negative fixed through keep_subj codes (0: 's', 1: 'o', 2: per-negative coin) -- negative row
    # j*n' + s*n + i belongs to side s.  Pinned against the reference by tests/golden/train_multiside_*.npz.
    _SIDE_CODE = {"s": 0, "o": 1, "s,o": 2, "s+o": 2}

    def _stack_sides(self, pos):
        """(stacked positives, keep_subj codes) for the model's side list; pos is a [n,3] int32 tensor."""
        f = self._fit
        sides, n = f["sides"], pos.shape[0]
        cache = f.setdefault("side_keep", {})
        if n not in cache:
            codes = torch.tensor([self._SIDE_CODE[s] for s in sides], dtype=torch.uint8)
            cache[n] = codes.repeat_interleave(n).repeat(self.eta).to(f["eng"].tdev)
        return pos.repeat(len(sides), 1), cache[n]

    def _step_kw(self, keep_subj=None):
        """Per-step scalar arguments.  KGE_F_PIPELINE (corruption generation + radix sort of step t+1 overlap step
        t, include/kge_b200.h) is requested when everything the corruption generator reads was resident before
        the previous step was submitted: batches are slices of the resident training array (or host buffers the
        library copies in itself); per-step device tensors made on torch's stream (the stacked batch of a side
        list, the entity list of negative_corruption_entities='batch') keep the in-order step."""
        f = self._fit
        kw = f["kw"]
        if f["pipeline"] and keep_subj is None and not self._neg_batch:
            kw = dict(kw, flags=kw["flags"] | _lib.F_PIPELINE)
        return kw

    def _fit_step_device(self, pos_dev, side="s,o", keep_subj=None, loss_out=None):
        """One optimisation step on a device-resident batch; enqueue only, the batch loss is written to `loss_out`
        (a 1-element fp32 device tensor; default f['loss_dev'])."""
        f = self._fit
        f["step"] += 1
        loss_out = f["loss_dev"] if loss_out is None else loss_out
        if keep_subj is None and not self._neg_batch:
            # the common step: only the batch, the counters and the loss slot change -> refresh the cached block
            kw = self._step_kw()
            a = f.get("args_dev")
            if a is not None and f.get("args_dev_side") == side:
                f["eng"].train_args_update(a, pos=pos_dev, step=f["step"], lr=kw["lr"], loss_out=loss_out, flags=kw["flags"])
            else:
                a = f["eng"].train_args(ent=f["ent"], rel=f["rel"], pos=pos_dev, loss_out=loss_out, side=_lib.TRAIN_SIDE_IDS[side],
                                        step=f["step"], **kw, **f["st"], **f["neg"])
                f["args_dev"], f["args_dev_side"] = a, side
            f["eng"].train_step(a)
            return
        neg = f["neg"]
        if self._neg_batch:  # corruptions drawn from the batch's own entities (evaluation/protocol.py:620-641)
            neg = dict(neg_entities=torch.unique(pos_dev[:, [0, 2]]).to(torch.int32))
        a = f["eng"].train_args(ent=f["ent"], rel=f["rel"], pos=pos_dev, loss_out=loss_out,
                                keep_subj=keep_subj,
                                side=_lib.TRAIN_SIDE_IDS[side], step=f["step"], **self._step_kw(keep_subj), **f["st"], **neg)
        f["eng"].train_step(a)

    def _host_step_args(self, pos_host, side, keep_subj):
        """Argument block of a host-buffer step (pos=None: the library copies the batch in itself); the common step
        refreshes a cached block, like _fit_step_device."""
        f = self._fit
        if keep_subj is None and not self._neg_batch:
            kw = self._step_kw()
            a = f.get("args_host")
            if a is not None and f.get("args_host_side") == side:
                return f["eng"].train_args_update(a, pos=None, step=f["step"], lr=kw["lr"], loss_out=f["loss_dev"], flags=kw["flags"])
            a = f["eng"].train_args(ent=f["ent"], rel=f["rel"], pos=None, loss_out=f["loss_dev"], n_pos=pos_host.shape[0],
                                    side=_lib.TRAIN_SIDE_IDS[side], step=f["step"], **kw, **f["st"], **f["neg"])
            f["args_host"], f["args_host_side"] = a, side
            return a
        neg = f["neg"]
        if self._neg_batch:
            neg = dict(neg_entities=to_dev_i32(np.unique(pos_host[:, [0, 2]].numpy()), f["eng"].tdev))
        return f["eng"].train_args(ent=f["ent"], rel=f["rel"], pos=None, loss_out=f["loss_dev"], keep_subj=keep_subj,
                                   n_pos=pos_host.shape[0], side=_lib.TRAIN_SIDE_IDS[side], step=f["step"], **self._step_kw(keep_subj),
                                   **f["st"], **neg)

    def _fit_step_host(self, pos_host, side="s,o", keep_subj=None):
        """One optimisation step fed like the reference feeds it: the batch comes from (pinned) host
        memory and the batch loss is read back (models/EmbeddingModel.py:1329-1337, :1421)."""
        f = self._fit
        f["step"] += 1
        a = self._host_step_args(pos_host, side, keep_subj)
        f["eng"].train_step_host(a, pos_host, f["loss_host"])
        return float(f["loss_host"][0])

    def _fit_step_host_pipelined(self, pos_host, side="s,o", keep_subj=None):
        """Like _fit_step_host, but the call returns as soon as the step is queued and hands back the loss of
        the step submitted ONE call earlier (None on the first call): the GPU always has the next step queued
        while the host checks the previous loss.  Every step still copies its batch in and its loss out;
        _fit_host_flush() returns the loss of the last step."""
        f = self._fit
        f["step"] += 1
        a = self._host_step_args(pos_host, side, keep_subj)  # the library copies the block during the call
        ring = f.setdefault("loss_ring", torch.zeros(4, dtype=torch.float32).pin_memory())
        i = f["step"] % 4
        ticket = f["eng"].train_step_host_async(a, pos_host, ring[i:i + 1])
        prev = f.get("pending")
        f["pending"] = (ticket, i, a, pos_host)  # keeps the args and the batch alive while in flight
        if prev is None:
            return None
        f["eng"].train_host_wait(prev[0])
        return float(ring[prev[1]])

    def _fit_host_flush(self):
        f = self._fit
        prev = f.pop("pending", None)
        if prev is None:
            return None
        f["eng"].train_host_wait(prev[0])
        return float(f["loss_ring"][prev[1]])

    # ---- multi-GPU fit: engine_params={"n_gpus": N}.  The tables and the optimizer state are split by column range over
    # N GPUs of one node (emgraph_b200/distributed.py:ShardedKGE) and every step runs the SAME global batch, corruption
    # stream and update as the single-GPU fit, so the trained parameters agree to rounding.  Two ways in:
    #   * the script runs under torchrun (torch.distributed initialised, world size == N): every rank calls fit() with the
    #     same arguments (SPMD) and ends up with the same model object
    #   * a plain `python script.py`: fit() spawns N worker processes (one per GPU, NCCL over 127.0.0.1), hands them the
    #     model's hyper-parameters and the id triples, and takes back the trained tables and the optimizer state
    # Replaces the reference's host-paged "large graph" training (models/EmbeddingModel.py:645-666, :1251-1281), which is
    # single-process and SGD-only.
    def _sharded_unsupported(self, early_stopping):
        f = []
        if early_stopping:
            f.append("early_stopping")
        if len(self._train_sides()) > 1:
            f.append("a list-valued corrupt_side")
        if self.embedding_model_params.get("negative_corruption_entities", DEFAULT_CORRUPTION_ENTITIES) != "all":
            f.append("negative_corruption_entities")
        if self._reg.get("reg_p", 0):
            f.append("the LP regulariser")
        if self.embedding_model_params.get("normalize_ent_emb", False):
            f.append("normalize_ent_emb")
        if self.engine_params.get("host_batches", False) or self.engine_params.get("reset_state", False):
            f.append("host_batches / reset_state")
        if f:
            raise NotImplementedError("engine_params['n_gpus'] > 1 does not support: " + ", ".join(f))

    def _fit_sharded_spmd(self, Xi, E, R):
        """Every rank of an initialised process group runs this with the same arguments."""
        import torch.distributed as dist
        from .distributed import ShardedKGE, slice_columns
        self._reseed_if_refit()
        K = self.internal_k
        resume = bool(getattr(self, "_resume", False))
        if resume:
            ent0 = np.ascontiguousarray(self.trained_model_params[0], dtype=np.float32)
            rel0 = np.ascontiguousarray(self.trained_model_params[1], dtype=np.float32)
            if ent0.shape != (E, K) or rel0.shape != (R, K):
                raise ValueError("resume: parameter shapes {} / {} do not fit this model".format(ent0.shape, rel0.shape))
        else:
            ent0, rel0 = self._init_table(E, K, "entity"), self._init_table(R, K, "relation")  # same seed on every rank
        N = Xi.shape[0]
        batch_size = int(np.ceil(N / self.batches_count))
        world = dist.get_world_size()
        norm = int(self.embedding_model_params.get("norm", DEFAULT_NORM_TRANSE)) if self.name == "TransE" else 1
        sk = ShardedKGE(self.name, self.k, self.eta, self.loss, self.optimizer, E, R, -(-batch_size // world), norm=norm,
                        margin=float(self.loss_params.get("margin", DEFAULT_MARGIN_ADVERSARIAL if self.loss == "self_adversarial" else DEFAULT_MARGIN)),
                        alpha=float(self.loss_params.get("alpha", DEFAULT_ALPHA_ADVERSARIAL)), seed=int(self.seed), init_ent=ent0, init_rel=rel0,
                        device=self.engine_params.get("device"), chunks=int(self.engine_params.get("chunks", 2)),
                        non_linearity=self.embedding_model_params.get("non_linearity", "linear"), side=self._train_sides()[0],
                        optimizer_params=self.optimizer_params, pipeline=bool(self.engine_params.get("pipeline", True)))
        del ent0, rel0
        if resume:
            saved = getattr(self, "_opt_state", None) or {}
            if set(saved) == set(sk.state):
                for nm, t in sk.state.items():  # the full-model state (merged at restore) -> this rank's columns
                    full = saved[nm].detach().cpu().numpy() if isinstance(saved[nm], torch.Tensor) else np.asarray(saved[nm])
                    t.copy_(torch.from_numpy(slice_columns(full, self.name, self.k, sk.world, sk.rank_id)).to(t.device))
            sk.step = int(getattr(self, "_opt_step", 0))
        dev = sk.eng.tdev
        Xd = to_dev_i32(Xi, dev)
        torch.cuda.synchronize(dev) if dev.type == "cuda" else None
        loss_steps = torch.zeros(self.batches_count, dtype=torch.float32, device=dev)
        sched = SGDSchedule(self.optimizer_params, self.batches_count) if self.optimizer == "sgd" else None
        denom = batch_size * (self.eta if self.loss in TILED_POSITIVE_LOSSES else 1) * self.batches_count
        self.loss_history = []
        for epoch in range(1, self.epochs + 1):
            loss_steps.zero_()
            for b in range(self.batches_count):
                lo, hi = b * batch_size, min(N, (b + 1) * batch_size)
                if hi <= lo:
                    continue
                if sched is not None:
                    sk.kw["lr"] = float(sched(b + 1, epoch))
                loss_steps[b:b + 1].copy_(sk.train_step(Xd[lo:hi], pos_is_global=True))
            el = float(loss_steps.double().sum().item())
            if not np.isfinite(el):  # models/EmbeddingModel.py:1422-1427
                raise ValueError("Loss is {}. Please change the hyperparameters.".format(el))
            self.loss_history.append(el / denom)
            if self.verbose and sk.rank_id == 0:
                print("Average Loss: {:10f} -- epoch {}/{}".format(self.loss_history[-1], epoch, self.epochs))
        self._sharded = sk
        self._best_params = None
        self._dev = None
        self._opt_step = sk.step
        # the whole model on every rank (host): predict / get_embeddings / save_model work as after a single-GPU fit;
        # the optimizer state stays sharded (save_model writes one file per rank)
        self.trained_model_params = [sk.gather_entities(), sk.gather_relations()]
        self._opt_state = {nm: sk.gather_state(nm) for nm in sk.state}
        self.is_fitted = True

    def _fit_sharded(self, Xi, E, R, n_gpus, early_stopping):
        import torch.distributed as dist
        self._sharded_unsupported(early_stopping)
        if dist.is_available() and dist.is_initialized():
            if dist.get_world_size() != n_gpus:
                raise ValueError("engine_params['n_gpus']={} but the process group has {} ranks".format(n_gpus, dist.get_world_size()))
            return self._fit_sharded_spmd(Xi, E, R)
        from .distributed import spawn_fit
        out = spawn_fit(self, Xi, E, R, n_gpus)
        self.trained_model_params = [out["ent"], out["rel"]]
        self._opt_state = {nm: torch.from_numpy(v) for nm, v in out["state"].items()}
        self._opt_step = int(out["step"])
        self.loss_history = list(out["loss_history"])
        self._dev = None
        self._best_params = None
        self.is_fitted = True

    def _fit_idx(self, Xi, E, R, early_stopping=False):
        n_gpus = int(self.engine_params.get("n_gpus", 1))
        if n_gpus > 1:
            return self._fit_sharded(Xi, E, R, n_gpus, early_stopping)
        f = self._fit_prepare(E, R)
        self._best_params = None
        self.early_stopping_history = []
        if early_stopping:
            self._initialize_early_stopping()
        eng, ent, rel = f["eng"], f["ent"], f["rel"]
        N = Xi.shape[0]
        batch_size = int(np.ceil(N / self.batches_count))
        host_batches = bool(self.engine_params.get("host_batches", False))
        pipelined = bool(self.engine_params.get("host_pipeline", True))
        if host_batches:
            Xh = torch.from_numpy(np.ascontiguousarray(Xi, dtype=np.int32)).pin_memory()
        else:
            Xd = to_dev_i32(Xi, eng.tdev)
            torch.cuda.synchronize(eng.tdev)  # KGE_F_PIPELINE: the batches are resident before the first step
        # every step of an epoch writes its batch loss into its own slot: no per-step accumulation kernels between
        # two steps, one float64 sum per epoch (and per NaN check)
        loss_steps = torch.zeros(self.batches_count, dtype=torch.float32, device=eng.tdev)
        normalize = bool(self.embedding_model_params.get("normalize_ent_emb", False))
        check_every = int(self.engine_params.get("nan_check_every", self.batches_count))
        self.loss_history = []
        multi_side = len(f["sides"]) > 1
        sched = SGDSchedule(self.optimizer_params, self.batches_count) if self.optimizer == "sgd" else None
        denom = batch_size * (self.eta if self.loss in TILED_POSITIVE_LOSSES else 1) * self.batches_count  # :1343-1344, :1453-1457
        for epoch in range(1, self.epochs + 1):
            loss_steps.zero_()
            host_loss = 0.0
            for b in range(self.batches_count):
                lo, hi = b * batch_size, min(N, (b + 1) * batch_size)
                if hi <= lo:
                    continue
                if sched is not None:  # sgd: decayed rate of this batch (training/sgd.py:127-185)
                    f["kw"]["lr"] = float(sched(b + 1, epoch))
                side, keep = f["sides"][0], None
                if host_batches:
                    pos_b = Xh[lo:hi]
                    if multi_side:  # one step over the positives stacked once per side
                        pos_b, keep = self._stack_sides(pos_b)
                        pos_b, side = pos_b.pin_memory(), "s,o"
                    lv = self._fit_step_host_pipelined(pos_b, side, keep) if pipelined else self._fit_step_host(pos_b, side, keep)
                    if lv is not None:
                        if not np.isfinite(lv):  # models/EmbeddingModel.py:1422-1427
                            raise ValueError("Loss is {}. Please change the hyperparameters.".format(lv))
                        host_loss += lv
                else:
                    pos_b = Xd[lo:hi]
                    if multi_side:
                        pos_b, keep = self._stack_sides(pos_b)
                        side = "s,o"
                    self._fit_step_device(pos_b, side, keep, loss_out=loss_steps[b:b + 1])
                if normalize:
                    if host_batches and pipelined:
                        lv = self._fit_host_flush()
                        host_loss += lv if lv is not None else 0.0
                    eng.normalize_rows(ent)
                if not host_batches and f["step"] % check_every == 0 and not bool(torch.isfinite(loss_steps).all().item()):
                    raise ValueError("Loss is {}. Please change the hyperparameters.".format(float(loss_steps.double().sum().item())))
            if host_batches and pipelined:  # the epoch's last step
                lv = self._fit_host_flush()
                host_loss += lv if lv is not None else 0.0
            el = host_loss if host_batches else float(loss_steps.double().sum().item())
            if not np.isfinite(el):  # models/EmbeddingModel.py:1422-1427
                raise ValueError("Loss is {}. Please change the hyperparameters.".format(el))
            self.loss_history.append(el / denom)
            if self.verbose:
                msg = "Average Loss: {:10f} -- epoch {}/{}".format(self.loss_history[-1], epoch, self.epochs)
                if early_stopping and self.early_stopping_best_value is not None:
                    msg += " -- Best validation ({}): {:5f}".format(self.early_stopping_criteria, self.early_stopping_best_value)
                print(msg)
            if early_stopping and self._perform_early_stopping_test(epoch):
                # the reference returns with the parameters of the best validation (:1471-1476)
                if self._best_params is not None:
                    f["ent"], f["rel"] = self._best_params
                self._fit_finish()
                return
        self._fit_finish()

    def _fit_finish(self):
        f = self._fit
        self._best_params = None
        self._dev = {"ent": f["ent"], "rel": f["rel"]}
        self._opt_state = f["st"]
        self._opt_step = f["step"]
        self.trained_model_params = [f["ent"].cpu().numpy(), f["rel"].cpu().numpy()]
        self.is_fitted = True

    def _device_params(self):
        """Device copies of the trained parameters (uploaded once; restored models upload lazily)."""
        if self._dev is None:
            eng = get_engine(self.engine_params.get("device"))
            e, r = self.trained_model_params[0], self.trained_model_params[1]
            self._dev = {"ent": torch.from_numpy(np.ascontiguousarray(e, dtype=np.float32)).to(eng.tdev),
                         "rel": torch.from_numpy(np.ascontiguousarray(r, dtype=np.float32)).to(eng.tdev)}
        return self._dev["ent"], self._dev["rel"]

    # ---- predict (models/EmbeddingModel.py:2101-2147)
    def predict(self, X, from_idx=False):
        if not self.is_fitted:
            raise RuntimeError("Model has not been fitted.")
        X = np.asarray(X)
        if X.ndim == 1:
            X = X[np.newaxis, :]
        Xi = X.astype(np.int32) if from_idx else to_idx(X, self._ent_index, self._rel_index)
        eng = get_engine(self.engine_params.get("device"))
        ent, rel = self._device_params()
        if from_idx and Xi.size:
            # the reference's gather raises on an id outside the tables (tf.nn.embedding_lookup, models/EmbeddingModel.py:504);
            # the kernel does not range-check, so the ids the user hands in are validated here
            E, R = int(ent.shape[0]), int(rel.shape[0])
            if Xi.min() < 0 or Xi[:, 0].max() >= E or Xi[:, 2].max() >= E or Xi[:, 1].max() >= R:
                raise ValueError("predict(from_idx=True): ids outside the fitted tables (entities 0..%d, relations 0..%d)" % (E - 1, R - 1))
        # the non-linearity is applied to the returned scores as the reference does (models/EmbeddingModel.py:2135-2147; its
        # tanh branch reads a non-existent attribute and raises -- the intended tanh(score) is returned here)
        out = eng.score(self._model_id(), self.k, ent, rel, to_dev_i32(Xi, eng.tdev), non_linearity=self._non_linearity())
        return out.cpu().numpy()

    # ---- get_embeddings (models/EmbeddingModel.py:455-488)
    def get_embeddings(self, entities, embedding_type="entity"):
        if not self.is_fitted:
            raise RuntimeError("Model has not been fitted.")
        if embedding_type == "entity":
            idxs = self._ent_index.lookup(np.asarray(entities), "entities")
            return self.trained_model_params[0][idxs]
        if embedding_type == "relation":
            idxs = self._rel_index.lookup(np.asarray(entities), "relations")
            return self.trained_model_params[1][idxs]
        raise ValueError("Invalid entity type: {}".format(embedding_type))

    # ---- evaluation hooks used by evaluate_performance (models/EmbeddingModel.py:1494-1518, :2035-2099)
    def set_filter_for_eval(self):
        self.is_filtered = True

    def configure_evaluation_protocol(self, config=None):
        if config is None:
            config = {"corruption_entities": DEFAULT_CORRUPTION_ENTITIES, "corrupt_side": DEFAULT_CORRUPT_SIDE_EVAL}
        self.eval_config = config

    def end_evaluation(self):
        if self.is_filtered and self.eval_dataset_handle is not None:
            self.eval_dataset_handle.cleanup()
            self.eval_dataset_handle = None
        self.is_filtered = False
        self.eval_config = {}

    def get_ranks(self, dataset_handle):
        """dataset_handle: evaluation.EvalDataset (test ids + optional filter ids)."""
        if not self.is_fitted:
            raise RuntimeError("Model has not been fitted.")
        self.eval_dataset_handle = dataset_handle
        eng = get_engine(self.engine_params.get("device"))
        ent, rel = self._device_params()
        return self._rank_with_params(eng, ent, rel, dataset_handle)

    def _rank_with_params(self, eng, ent, rel, dataset_handle):
        """Ranks of the handle's triples under eval_config with the given device parameters (shared by
        get_ranks and the early-stopping check, which ranks with the parameters being trained)."""
        side = self.eval_config.get("corrupt_side", DEFAULT_CORRUPT_SIDE_EVAL)
        strategy = self.eval_config.get("ranking_strategy", DEFAULT_RANK_COMPARE_STRATEGY)
        assert side in _lib.RANK_SIDE_IDS, "Invalid value for corrupt_side."
        assert strategy in ("worst", "best", "middle"), "Invalid score comparision type!"
        ce = self.eval_config.get("corruption_entities", DEFAULT_CORRUPTION_ENTITIES)
        filtered = bool(self.is_filtered)
        mid = self._model_id()
        use_tc = bool(self.engine_params.get("rank_tensor_cores", False)) and self.name != "TransE"
        E, R = ent.shape[0], rel.shape[0]
        if ce is None or isinstance(ce, str):
            if filtered:
                dataset_handle.build_filter(eng, E, R)
            # host buffers in, host buffers out (one H2D of the test triples, one D2H of the ranks)
            test_h = dataset_handle.test_host_pinned()
            T = test_h.shape[0]
            ranks_h = torch.empty((T, 2) if side == "s,o" else (T,), dtype=torch.int32).pin_memory()
            eng.rank_host(mid, self.k, ent, rel, test_h, ranks_h, side=_lib.RANK_SIDE_IDS[side],
                          strategy=_lib.STRATEGY_IDS[strategy], filtered=filtered, use_tensor_cores=use_tc,
                          non_linearity=self._non_linearity())
            return ranks_h.numpy().copy()
        # ---- entities_subset: corruptions only from `ce` (models/EmbeddingModel.py:1845-1857, :1898-1940).
        # The entities are re-labelled so that the subset occupies rows [0,|C|) of a permuted table; the sweep
        # then covers exactly those rows and the (re-labelled) filter applies unchanged.
        dev = eng.tdev
        C = torch.unique(torch.as_tensor(np.asarray(ce, dtype=np.int64).reshape(-1), device=dev))
        assert C.numel() > 0 and int(C.min()) >= 0 and int(C.max()) < E, "corruption_entities out of range"
        in_c = torch.zeros(E, dtype=torch.bool, device=dev)
        in_c[C] = True
        order = torch.cat([C, (~in_c).nonzero().squeeze(1)])  # new row -> old id
        perm = torch.empty(E, dtype=torch.int64, device=dev)
        perm[order] = torch.arange(E, device=dev)  # old id -> new row
        ent_p = ent.index_select(0, order).contiguous()
        test = dataset_handle.test_device(dev).long()
        T = test.shape[0]
        if T == 0:
            return np.zeros((0, 2) if side == "s,o" else (0,), np.int32)
        test_p = torch.stack([perm[test[:, 0]], test[:, 1], perm[test[:, 2]]], 1).to(torch.int32).contiguous()
        if filtered:
            dataset_handle.build_filter(eng, E, R, perm=perm)
        nC = int(C.numel())
        counts = eng.rank_counts(mid, self.k, ent_p, rel, test_p, side=_lib.RANK_SIDE_IDS[side], filtered=filtered,
                                 use_tensor_cores=use_tc, ent_local=ent_p[:nC], row_begin=0, row_end=nC,
                                 non_linearity=self._non_linearity())
        self_cand = torch.stack([in_c[test[:, 0]], in_c[test[:, 2]]], 1).to(torch.uint8).contiguous()
        ranks = eng.rank_finalize(counts, side=_lib.RANK_SIDE_IDS[side], strategy=_lib.STRATEGY_IDS[strategy], filtered=filtered,
                                  self_is_candidate=self_cand)
        return ranks.cpu().numpy()


@register_model("TransE")
class TransE(EmbeddingModel):
    """f = -||s + p - o||_n  (models/TransE.py:190-216)."""


@register_model("DistMult")
class DistMult(EmbeddingModel):
    """f = sum s*p*o  (models/DistMult.py:181-201)."""


@register_model("ComplEx")
class ComplEx(EmbeddingModel):
    """f = Re<p, s, conj(o)>, rows [re(k) | im(k)]  (models/ComplEx.py:267-298)."""


@register_model("HolE")
class HolE(EmbeddingModel):
    """f = (2/k) * ComplEx score  (models/HolE.py:169-189)."""
