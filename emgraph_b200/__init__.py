"""emgraph_b200 -- B200-native KGE train step + filtered ranking behind the emgraph API.

    from emgraph_b200.models import TransE, DistMult, ComplEx, HolE
    from emgraph_b200.evaluation import evaluate_performance, mrr_score, hits_at_n_score

All arithmetic runs in hand-written sm_100a CUDA kernels (emgraph_b200/csrc) reached through the C
ABI of include/kge_b200.h; there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _lib, evaluation, models, utils  # noqa: F401
from .models import ComplEx, DistMult, EmbeddingModel, HolE, TransE, reset_entity_threshold, set_entity_threshold  # noqa: F401
from .utils import dataframe_to_triples, get_entity_triples, restore_model, save_model  # noqa: F401
