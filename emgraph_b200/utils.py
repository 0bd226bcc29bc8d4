"""Mirror of ``emgraph.utils.model_utils`` save/restore (utils/model_utils.py:22-160): the same pickle
dictionary (class name, hyper-parameters, label maps, NumPy parameter arrays), so a file written by either
library restores in the other.  Extension: ``save_optimizer_state=True`` also stores the sparse optimizer's
per-row state (the reference has none to save: its optimizers are re-created every batch)."""
from __future__ import annotations

import glob
import importlib
import os
import pickle
from time import gmtime, strftime

import numpy as np

DEFAULT_MODEL_NAMES = "{0}.model.pkl"  # utils/model_utils.py:16


def save_model(model, model_name_path=None, protocol=pickle.HIGHEST_PROTOCOL, save_optimizer_state=False):
    """utils/model_utils.py:22-87."""
    obj = {
        "class_name": model.__class__.__name__,
        "hyperparams": model.all_params,
        "is_fitted": model.is_fitted,
        "ent_to_idx": model.ent_to_idx,
        "rel_to_idx": model.rel_to_idx,
        "is_calibrated": model.is_calibrated,
    }
    model.get_embedding_model_params(obj)
    if model_name_path is None:
        model_name_path = DEFAULT_MODEL_NAMES.format(strftime("%Y_%m_%d-%H_%M_%S", gmtime()))
    sk = getattr(model, "_sharded", None)
    sharded = save_optimizer_state and sk is not None and _dist_ranks() > 1
    if sharded:
        # a model trained under torchrun: every rank writes the optimizer state of ITS column slice next to the main file
        # (<path>.opt.<rank>-of-<world>.npz); rank 0 writes the reference-format pickle with the whole parameter tables
        np.savez(shard_path(model_name_path, sk.rank_id, sk.world), step=int(sk.step), world=sk.world, rank=sk.rank_id, k=sk.k,
                 model=sk.model, **{nm: t.detach().cpu().numpy() for nm, t in sk.state.items()})
        obj["b200_optimizer_shards"] = int(sk.world)
        obj["b200_optimizer_step"] = int(sk.step)
    elif save_optimizer_state and getattr(model, "_opt_state", None):
        obj["b200_optimizer_state"] = {k: v.detach().cpu().numpy() for k, v in model._opt_state.items()}
        obj["b200_optimizer_step"] = int(getattr(model, "_opt_step", 0))
    if not sharded or sk.rank_id == 0:
        with open(model_name_path, "wb") as fw:
            pickle.dump(obj, fw, protocol=protocol)
    if sharded:
        import torch.distributed as dist
        dist.barrier()
    return model_name_path


def _dist_ranks():
    try:
        import torch.distributed as dist
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    except Exception:
        return 1


def shard_path(model_name_path, rank, world):
    return "{}.opt.{}-of-{}.npz".format(model_name_path, rank, world)


def load_optimizer_shards(model_name_path, world):
    """Merge the per-rank optimizer-state files of a column-sharded model into full-model tables (any later fit --
    one GPU or a different number of GPUs -- slices them again)."""
    from .distributed import merge_columns
    parts = [np.load(shard_path(model_name_path, r, world)) for r in range(world)]
    model, k = str(parts[0]["model"]), int(parts[0]["k"])
    names = [nm for nm in parts[0].files if nm in ("ent_m", "ent_v", "rel_m", "rel_v")]
    return {nm: merge_columns([p[nm] for p in parts], model, k) for nm in names}, int(parts[0]["step"])


def restore_model(model_name_path=None):
    """utils/model_utils.py:90-160."""
    if model_name_path is None:
        default_models = sorted(glob.glob("*.model.pkl"))
        if len(default_models) == 0:
            raise Exception("No default model found. Please specify model_name_path...")
        model_name_path = default_models[len(default_models) - 1]
    try:
        with open(model_name_path, "rb") as fr:
            restored_obj = pickle.load(fr)
        module = importlib.import_module("emgraph_b200.models")
        class_ = getattr(module, restored_obj["class_name"])
        model = class_(**restored_obj["hyperparams"])
        model.is_fitted = restored_obj["is_fitted"]
        model.ent_to_idx = restored_obj["ent_to_idx"]
        model.rel_to_idx = restored_obj["rel_to_idx"]
        model.is_calibrated = restored_obj.get("is_calibrated", False)
        model.restore_model_params(restored_obj)
        if restored_obj.get("b200_optimizer_shards"):
            import torch
            st, step = load_optimizer_shards(model_name_path, int(restored_obj["b200_optimizer_shards"]))
            model._opt_state = {k_: torch.from_numpy(np.ascontiguousarray(v)) for k_, v in st.items()}
            model._opt_step = step
    except pickle.UnpicklingError as e:
        raise Exception("Error unpickling model {} : {}.".format(model_name_path, e))
    except (IOError, FileNotFoundError):
        raise FileNotFoundError("No model found: {}.".format(model_name_path))
    return model


# ------------------------------------------------------------------------------------------------
# small data-format helpers a reference script uses on the way to fit() (utils/__init__.py)
# ------------------------------------------------------------------------------------------------
def dataframe_to_triples(X, schema):
    """Rows of a DataFrame as triples (utils/model_utils.py:326-365).  schema: (subject column, relation name, object
    column) tuples; one triple per row and schema entry, in schema order."""
    schema = [tuple(t) for t in schema]
    missing = {c for s, _, o in schema for c in (s, o)} - set(X.columns)
    if missing:
        raise Exception("Subject/Object {} are not in data frame headers".format(missing))
    blocks = []
    for s, p, o in schema:
        block = np.empty((len(X), 3), dtype=object)
        block[:, 0], block[:, 1], block[:, 2] = X[s].to_numpy(), p, X[o].to_numpy()
        blocks.append(block)
    # one homogeneous (string) array, as np.array(list of mixed lists) gives in the reference
    return np.array(np.concatenate(blocks).tolist()) if blocks else np.array([])


def get_entity_triples(entity, graph):
    """All triples of `graph` [n, 3] with `entity` as subject or object, in graph order (utils/misc.py:28-56)."""
    graph = np.asarray(graph)
    return graph[(graph[:, 0] == entity) | (graph[:, 2] == entity)]
