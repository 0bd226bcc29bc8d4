"""CPU: the C-ABI library loads and exports every symbol include/kge_b200.h declares; host-side
logic (id mapping, constructor error conventions, metrics) mirrors the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from emgraph_b200 import _lib as L
    if not os.path.exists(L.LIB_PATH):
        from emgraph_b200.build import build
        build(verbose=False)
    return L


def test_header_symbols_exported_and_bound():
    L = _lib()
    hdr = open(os.path.join(ROOT, "include", "kge_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(kge_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = C.CDLL(L.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "missing export " + name
        assert name in L.SYMBOLS, "ctypes binding missing for " + name
    assert set(L.SYMBOLS) == declared
    assert L.load().kge_abi_version() == L.ABI_VERSION


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _lib()
    h = C.c_void_p()
    rc = L.load().kge_ctx_create(0, C.byref(h))
    assert rc != 0 and b"no CPU fallback" in L.load().kge_last_error()
    from emgraph_b200.engine import get_engine
    with pytest.raises(L.KgeError):
        get_engine(0)
    from emgraph_b200.models import DistMult
    m = DistMult(k=4, eta=1, epochs=1, batches_count=1)
    with pytest.raises(L.KgeError):
        m.fit(np.array([["a", "x", "b"], ["b", "x", "c"]]))


def test_struct_layout_matches_header():
    """ctypes mirror of kge_table / kge_train_args against sizes computed by the C compiler."""
    import subprocess
    import tempfile
    L = _lib()
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "kge_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(kge_table), sizeof(kge_train_args), offsetof(kge_train_args, ent), offsetof(kge_train_args, pos), offsetof(kge_train_args, dbg_grad_rel));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "t.c")
        open(p, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), p, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    st, sa, o_ent, o_pos, o_last = (int(v) for v in out)
    assert C.sizeof(L.KgeTable) == st
    assert C.sizeof(L.KgeTrainArgs) == sa
    assert L.KgeTrainArgs.ent.offset == o_ent
    assert L.KgeTrainArgs.pos.offset == o_pos
    assert L.KgeTrainArgs.dbg_grad_rel.offset == o_last


def test_mappings_match_reference_goldens():
    from emgraph_b200.models import create_mappings, to_idx
    # reference tests/emgraph/evaluation/test_protocol.py:490-496
    X = np.array([["a", "x", "b"], ["c", "y", "d"]])
    rel_to_idx, ent_to_idx = create_mappings(X)
    np.testing.assert_array_equal(to_idx(X, ent_to_idx=ent_to_idx, rel_to_idx=rel_to_idx), [[0, 0, 1], [2, 1, 3]])
    assert ent_to_idx == {"a": 0, "b": 1, "c": 2, "d": 3} and rel_to_idx == {"x": 0, "y": 1}
    with pytest.raises(ValueError):
        to_idx(np.array([["a", "x", "zz"]]), ent_to_idx, rel_to_idx)
    with pytest.raises(ValueError):
        to_idx(np.array([["a", "q", "b"]]), ent_to_idx, rel_to_idx)
    # 1-d input is promoted (protocol.py:721-722)
    np.testing.assert_array_equal(to_idx(np.array(["c", "x", "a"]), ent_to_idx, rel_to_idx), [[2, 0, 0]])
    # agrees with the oracle's dict-based restatement on random labels
    from oracle import kge_oracle as ko
    rng = np.random.default_rng(0)
    Xr = rng.integers(0, 50, size=(200, 3)).astype(str)
    r1, e1 = create_mappings(Xr)
    r2, e2 = ko.create_mappings(Xr)
    assert r1 == r2 and e1 == e2
    np.testing.assert_array_equal(to_idx(Xr, e1, r1), ko.to_idx(Xr, e2, r2))


def test_constructor_error_conventions():
    # reference models/EmbeddingModel.py:206-210, :257-298; tests/emgraph/models/test_models.py:22-35
    from emgraph_b200.models import ComplEx, DistMult, HolE, TransE
    for cls in (TransE, DistMult, ComplEx, HolE):
        with pytest.raises(ValueError):
            cls(loss="bce")
        with pytest.raises(ValueError):
            cls(loss="nope")
        with pytest.raises(ValueError):
            cls(optimizer="nope")
        with pytest.raises(ValueError):
            cls(initializer="nope")
        with pytest.raises(ValueError):
            cls(regularizer="nope")
        m = cls(k=10, eta=3)
        assert m.internal_k == (20 if cls in (ComplEx, HolE) else 10)
        assert m.get_hyperparameter_dict()["k"] == 10 and not m.is_fitted
        with pytest.raises(RuntimeError):
            m.predict(np.array([["a", "b", "c"]]))
        with pytest.raises(RuntimeError):
            m.get_ranks(None)
        with pytest.raises(ValueError):
            m.fit([["a", "b", "c"]])


def test_metrics_match_reference_goldens():
    from emgraph_b200.evaluation import hits_at_n_score, mr_score, mrr_score, rank_score
    assert rank_score(np.array([0, 0, 1, 0]), np.array([0.434, 0.65, 0.21, 0.84])) == 4
    assert hits_at_n_score(np.array([1, 12, 6, 2]), n=3) == 0.5
    np.testing.assert_almost_equal(mrr_score([1, 12, 6, 2]), 0.4375)
    assert mr_score(np.array([[1, 12], [6, 2]])) == 5.25


def test_evaluate_performance_argument_checks():
    from emgraph_b200.evaluation import evaluate_performance
    from emgraph_b200.models import DistMult
    m = DistMult(k=4)
    with pytest.raises(AssertionError):
        evaluate_performance(np.zeros((0, 3)), m, corrupt_side="x")
