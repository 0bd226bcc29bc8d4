"""The toy graph of the reference's own fit/predict tests (tests/emgraph/models/test_models.py:218-335) and the
oracle's emulation of the engine's fit loop on it; shared by the CPU and the GPU tests."""
import numpy as np

from oracle import kge_oracle as ko

TOY_X = np.array([["a", "y", "b"], ["b", "y", "a"], ["a", "y", "c"], ["c", "y", "a"], ["a", "y", "d"], ["c", "y", "d"],
                  ["b", "y", "c"], ["f", "y", "e"]])
TOY_QUERY = np.array([["f", "y", "e"], ["b", "y", "d"]])
# model, batches_count, pairwise margin, LP (lambda, p) or None -- the reference's constructor arguments
TOY_CASES = [("TransE", 1, 5.0, None), ("DistMult", 2, 5.0, None), ("ComplEx", 1, 1.0, (0.1, 2)), ("HolE", 1, 1.0, (0.1, 2))]


def toy_fit_emulation(model, bc, margin, reg, seed=555):
    """(scores of TOY_QUERY, per-epoch summed loss) after k=10, eta=2, 20 epochs of adagrad lr=0.1."""
    r2i, e2i = ko.create_mappings(TOY_X)
    Xi = ko.to_idx(TOY_X, e2i, r2i)
    lam, p = reg if reg else (0.0, 0)
    ent, rel, losses = ko.fit_emulation(model, 10, 2, 20, bc, seed, "pairwise", "adagrad", 0.1, Xi, len(e2i), len(r2i),
                                        margin=margin, reg_p=p, reg_lambda_ent=lam, reg_lambda_rel=lam)
    return ko.score(model, 10, ent, rel, ko.to_idx(TOY_QUERY, e2i, r2i)), losses
