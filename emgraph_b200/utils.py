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
    if save_optimizer_state and getattr(model, "_opt_state", None):
        obj["b200_optimizer_state"] = {k: v.detach().cpu().numpy() for k, v in model._opt_state.items()}
        obj["b200_optimizer_step"] = int(getattr(model, "_opt_step", 0))
    if model_name_path is None:
        model_name_path = DEFAULT_MODEL_NAMES.format(strftime("%Y_%m_%d-%H_%M_%S", gmtime()))
    with open(model_name_path, "wb") as fw:
        pickle.dump(obj, fw, protocol=protocol)
    return model_name_path


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
