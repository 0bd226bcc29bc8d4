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
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "kge_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(kge_table), sizeof(kge_train_args), offsetof(kge_train_args, ent), offsetof(kge_train_args, pos), offsetof(kge_train_args, dbg_grad_rel), offsetof(kge_train_args, stage), offsetof(kge_train_args, alpha), offsetof(kge_train_args, neg_entities_n));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "t.c")
        open(p, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), p, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    st, sa, o_ent, o_pos, o_last, o_stage, o_alpha, o_negn = (int(v) for v in out)
    assert L.KgeTrainArgs.stage.offset == o_stage
    assert L.KgeTrainArgs.alpha.offset == o_alpha
    assert L.KgeTrainArgs.neg_entities_n.offset == o_negn
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


def test_regularizer_and_loss_hyperparameters():
    # regularizers/lp.py:41-104, losses/_loss_constants.py:8-12
    from emgraph_b200.models import ComplEx
    m = ComplEx(k=4, regularizer="LP", regularizer_params={"p": 3, "lambda": [1e-3, 1e-2]})
    assert m._reg == dict(reg_p=3, reg_lambda_ent=1e-3, reg_lambda_rel=1e-2)
    assert ComplEx(k=4, regularizer="LP", regularizer_params={})._reg == dict(reg_p=2, reg_lambda_ent=1e-5, reg_lambda_rel=1e-5)
    assert ComplEx(k=4)._reg["reg_p"] == 0
    with pytest.raises(Exception):
        ComplEx(k=4, regularizer="LP", regularizer_params={"p": 1.5})
    with pytest.raises(ValueError):
        ComplEx(k=4, regularizer="LP", regularizer_params={"lambda": [1.0, 2.0, 3.0]})
    for loss in ("absolute_margin", "self_adversarial"):
        assert ComplEx(k=4, loss=loss).loss == loss


def test_save_restore_roundtrip_host(tmp_path):
    # utils/model_utils.py:22-160 : same pickle dictionary as the reference
    import pickle
    from emgraph_b200 import restore_model, save_model
    from emgraph_b200.models import DistMult, LabelIndex
    m = DistMult(k=3, eta=2, epochs=7, batches_count=2, seed=5, loss="pairwise", loss_params={"margin": 2.0})
    m._ent_index = LabelIndex(np.array(["a", "b", "c"]))
    m._rel_index = LabelIndex(np.array(["x"]))
    m.trained_model_params = [np.arange(9, dtype=np.float32).reshape(3, 3), np.ones((1, 3), np.float32)]
    m.is_fitted = True
    path = str(tmp_path / "m.pkl")
    save_model(m, path)
    obj = pickle.load(open(path, "rb"))
    assert set(obj) >= {"class_name", "hyperparams", "is_fitted", "ent_to_idx", "rel_to_idx", "is_calibrated", "model_params",
                        "large_graph", "calibration_parameters"}
    assert obj["class_name"] == "DistMult" and obj["ent_to_idx"] == {"a": 0, "b": 1, "c": 2}
    r = restore_model(path)
    assert type(r) is DistMult and r.is_fitted and r.all_params == m.all_params
    assert r.ent_to_idx == m.ent_to_idx and r.rel_to_idx == m.rel_to_idx
    np.testing.assert_array_equal(r.trained_model_params[0], m.trained_model_params[0])
    np.testing.assert_array_equal(r.get_embeddings(np.array(["c", "a"])), m.trained_model_params[0][[2, 0]])
    assert r.is_fitted_on(np.array([["a", "x", "b"], ["b", "x", "c"]])) and not r.is_fitted_on(np.array([["a", "x", "b"]]))
    with pytest.raises(FileNotFoundError):
        restore_model(str(tmp_path / "missing.pkl"))


def test_lookup_known_drops_unknown_labels():
    from emgraph_b200.models import LabelIndex
    li = LabelIndex(np.array(["a", "b", "d", "e"]))
    np.testing.assert_array_equal(li.lookup_known(np.array(["e", "zz", "a", "a"])), [0, 3])
    assert li.lookup_known(np.array(["q"])).size == 0


def test_early_stopping_argument_checks():
    from emgraph_b200.models import DistMult, LabelIndex
    m = DistMult(k=3)
    m._ent_index = LabelIndex(np.array(["a", "b"]))
    m._rel_index = LabelIndex(np.array(["x"]))
    m.early_stopping_params = {}
    with pytest.raises(KeyError):
        m._initialize_early_stopping()
    m.early_stopping_params = {"x_valid": np.array([["a", "x", "b"]]), "criteria": "nope"}
    with pytest.raises(ValueError):
        m._initialize_early_stopping()
    m.early_stopping_params = {"x_valid": np.array(["a", "x", "b"])}
    with pytest.raises(ValueError):
        m._initialize_early_stopping()


def test_sgd_schedule_reference_goldens():
    # reference tests/emgraph/models/test_optimizers.py:6-79 (exact values, compared with == there too)
    from emgraph_b200.optimizers import SGDSchedule
    s = SGDSchedule({"lr": 0.001}, 10)
    v = [s(b, e) for e in range(1, 11) for b in range(1, 11)][-1]
    assert v == 0.001
    s = SGDSchedule({"lr": 0.001, "decay_lr_rate": 2, "cosine_decay": False, "decay_cycle": 10}, 10)
    v = [s(b, e) for e in range(1, 11) for b in range(1, 11)][-1]
    assert v == 0.001 and s(1, 11) == 0.0005
    s = SGDSchedule({"lr": 0.001, "end_lr": 0.00001, "decay_lr_rate": 2, "expand_factor": 2, "cosine_decay": True, "decay_cycle": 10}, 10)
    seen = {}
    for e in range(1, 31):
        for b in range(1, 11):
            seen[(e, b)] = s(b, e)
    assert seen[(11, 1)] == 0.0005 and seen[(6, 1)] == 0.000505 and seen[(21, 1)] == 0.000255
    assert s(1, 31) == 0.00025


def test_quantised_comparison_thresholds_mirror():
    """Host mirror of csrc/kge_rank.cu:kge_quant_thresholds -- the TransE sweep compares y = score*1e5 against
    two fp32 thresholds instead of quantising every candidate: trunc(y) >= n <=> y >= t_ge and
    trunc(y) > n <=> y >= t_gt must hold for every fp32 y (reference models/EmbeddingModel.py:2010-2029)."""
    INF = np.float32(np.inf)

    def at_least(m, strictly):
        f = np.float32(m)  # round to nearest, like __ll2float_rn
        if (float(f) <= m) if strictly else (float(f) < m):
            f = np.nextafter(f, INF)
        return f

    def thresholds(n):
        m0, m1 = n, n + 1
        t_ge = at_least(m0, False) if m0 > 0 else at_least(m0 - 1, True)
        t_gt = np.float32(np.nan) if n == 2**31 - 1 else (at_least(m1, False) if m1 > 0 else at_least(m1 - 1, True))  # NaN: never
        return t_ge, t_gt

    def trunc_sat(y):  # F2I.TRUNC saturates
        if not np.isfinite(y):
            return 2**31 - 1 if y > 0 else -2**31
        return int(max(-2**31, min(2**31 - 1, int(np.trunc(np.float64(y))))))

    rng = np.random.default_rng(0)
    ns = [0, 1, -1, 2, -2, 7, -7, 2**24 - 1, 2**24, 2**24 + 1, 2**24 + 3, -2**24 - 1, -2**24 - 3, 33554433, -33554435,
          2**31 - 1, 2**31 - 2, 2**31 - 200, -2**31 + 1, -2**31 + 300] + [int(v) for v in rng.integers(-2**31 + 2, 2**31 - 2, 40)] \
        + [int(v) for v in rng.integers(-3000000, 3000000, 40)]
    for n in ns:
        t_ge, t_gt = thresholds(n)
        ys = []
        for c in (n - 2, n - 1, n, n + 1, n + 2):
            y = np.float32(c)
            lo = hi = y
            for _ in range(6):  # the fp32 neighbourhood of every nearby integer
                lo, hi = np.nextafter(lo, -INF), np.nextafter(hi, INF)
                ys += [lo, hi]
            ys.append(y)
        ys += list(rng.normal(scale=max(4.0, abs(n) * 1e-3), size=50).astype(np.float32) + np.float32(n))
        ys += [np.float32(0.0), np.float32(-0.0), np.float32(0.5), np.float32(-0.5), np.float32(-0.99999994), np.float32(3e9), INF]
        for y in ys:
            q = trunc_sat(y)
            assert (q >= n) == bool(y >= t_ge), (n, float(y), q, float(t_ge))
            assert (q > n) == bool(y >= t_gt), (n, float(y), q, float(t_gt))


def test_integration_stub_matches_the_binding():
    """INTEGRATION.md shows the ctypes stub a reference maintainer would add: its kge_train_args fields must be the
    binding's (same names, same order), and every entry point it calls must be declared in include/kge_b200.h."""
    import re
    from emgraph_b200 import _lib as L
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    block = doc[doc.index("class KgeTrainArgs(C.Structure)"):doc.index("lib.kge_last_error.restype")]
    names = re.findall(r'\("(\w+)",\s*(?:C\.|KgeTable)', block)
    assert names == [f[0] for f in L.KgeTrainArgs._fields_]
    header = open(os.path.join(root, "include", "kge_b200.h")).read()
    for fn in set(re.findall(r"lib\.(kge_\w+)", doc)):
        assert re.search(r"\b%s\s*\(" % fn, header), fn


def test_header_is_plain_c(tmp_path):
    """include/kge_b200.h is the boundary a non-Python host binds (cgo / JNI / FFI): it must compile as C99 and as
    C++ on its own, and a C translation unit that references every declared entry point must link against the
    library (no torch, no C++ types in the signatures)."""
    import re
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(root, "include", "kge_b200.h")
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr], check=True)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr], check=True)
    names = sorted(set(re.findall(r"^(?:int|int64_t|const char\*)\s+(kge_[a-z0-9_]+)\s*\(", open(hdr).read(), re.M)))
    src = tmp_path / "link_all.c"
    src.write_text('#include "kge_b200.h"\n#include <stdio.h>\nint main(void) {\n  void* p[] = {%s};\n'
                   '  printf("%%d %%d\\n", kge_abi_version(), (int)(sizeof(p) / sizeof(p[0])));\n  return kge_abi_version() == KGE_ABI_VERSION ? 0 : 1;\n}\n'
                   % ", ".join("(void*)%s" % n for n in names))
    lib_dir = os.path.join(root, "emgraph_b200")
    exe = tmp_path / "link_all"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(root, "include"), str(src), "-o", str(exe), "-L", lib_dir,
                    "-l:libkge_b200.so", "-Wl,-rpath," + lib_dir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[1]) == len(names) >= 30


def test_refit_starts_from_the_seed_again():
    """models/EmbeddingModel.py:1285-1290: a fitted model re-seeds its RNG when fit() is called again (host logic only;
    the GPU test tests/test_zzz_gpu_reference_properties.py::test_refit_is_deterministic runs the real thing)."""
    from emgraph_b200 import models
    m = models.ComplEx(k=4, seed=555)
    first = (m._init_table(6, 8, "entity"), m._init_table(1, 8, "relation"))
    m._reseed_if_refit()  # not fitted yet: the stream simply continues
    assert not np.array_equal(m._init_table(6, 8, "entity"), first[0])
    m.is_fitted = True
    m._reseed_if_refit()
    np.testing.assert_array_equal(m._init_table(6, 8, "entity"), first[0])
    np.testing.assert_array_equal(m._init_table(1, 8, "relation"), first[1])


def test_train_test_split_matches_the_reference():
    """train_test_split_no_unseen against (a) the reference's own golden (tests/emgraph/evaluation/test_protocol.py:608-639),
    (b) 72 splits the reference itself produced (tests/golden/split_cases.npz, oracle/make_golden.py), error messages
    included, and (c) the properties of reference test_protocol.py:642-677."""
    from emgraph_b200.evaluation import train_test_split_no_unseen
    X = np.array([["a", "y", "b"], ["a", "y", "c"], ["c", "y", "a"], ["d", "y", "e"], ["e", "y", "f"], ["f", "y", "c"], ["f", "y", "c"]])
    tr, te = train_test_split_no_unseen(X, test_size=2, seed=0, backward_compatible=True)
    np.testing.assert_array_equal(tr, [["a", "y", "b"], ["c", "y", "a"], ["d", "y", "e"], ["e", "y", "f"], ["f", "y", "c"]])
    np.testing.assert_array_equal(te, [["a", "y", "c"], ["f", "y", "c"]])
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "split_cases.npz"))
    raised = 0
    for i in range(int(g["n_cases"])):
        ts, seed, dup, bc, filt = [int(v) for v in g["par_%d" % i]]
        kw = dict(test_size=ts, seed=seed, allow_duplication=bool(dup), filtered_test_predicates=["r0"] if filt else None,
                  backward_compatible=bool(bc))
        err = str(g["err_%d" % i])
        if err:
            raised += 1
            with pytest.raises(Exception) as e:
                train_test_split_no_unseen(g["X_%d" % i], **kw)
            assert str(e.value) == err
        else:
            tr, te = train_test_split_no_unseen(g["X_%d" % i], **kw)
            np.testing.assert_array_equal(tr, g["train_%d" % i])
            np.testing.assert_array_equal(te, g["test_%d" % i])
    assert raised >= 1
    # a fraction as test_size; nothing unseen; sizes add up; too large a request raises unless duplicates are allowed
    rng = np.random.default_rng(1)
    X = np.stack([rng.integers(0, 300, 4000), rng.integers(0, 6, 4000), rng.integers(0, 300, 4000)], 1).astype(str)
    tr, te = train_test_split_no_unseen(X, 0.5)
    assert tr.shape[0] + te.shape[0] == 4000 and te.shape[0] == 2000
    assert set(te[:, 0]) | set(te[:, 2]) <= set(tr[:, 0]) | set(tr[:, 2]) and set(te[:, 1]) <= set(tr[:, 1])
    with pytest.raises(Exception, match="Cannot create a test split of the desired size"):
        train_test_split_no_unseen(X, 0.99)
    tr, te = train_test_split_no_unseen(X, 0.99, allow_duplication=True)
    assert tr.shape[0] + te.shape[0] > 4000


def test_corruption_utilities():
    """generate_corruptions_for_eval layout (reference tests/emgraph/evaluation/test_protocol.py:418-455) and the
    invariants of generate_corruptions_for_fit (reference :530-605: one side of every row is the positive's)."""
    from emgraph_b200.evaluation import generate_corruptions_for_eval, generate_corruptions_for_fit
    got = generate_corruptions_for_eval(np.array([[0, 0, 1]]), np.arange(8))
    np.testing.assert_array_equal(got, [[0, 0, e] for e in range(8)] + [[e, 0, 1] for e in range(8)])
    np.testing.assert_array_equal(generate_corruptions_for_eval(np.array([[0, 0, 1]]), np.arange(3), "s"), [[e, 0, 1] for e in range(3)])
    np.testing.assert_array_equal(generate_corruptions_for_eval(np.array([[0, 0, 1]]), np.arange(3), "o"), [[0, 0, e] for e in range(3)])
    with pytest.raises(ValueError):
        generate_corruptions_for_eval(np.array([[0, 0, 1]]), np.arange(3), "x")
    X = np.array([[0, 0, 1], [2, 1, 3], [4, 0, 5]])
    for side in ("s,o", "s+o", "s", "o"):
        neg = generate_corruptions_for_fit(X, eta=4, corrupt_side=side, entities_size=50, rnd=7)
        pos = np.tile(X, (4, 1))
        assert neg.shape == (12, 3) and np.array_equal(neg[:, 1], pos[:, 1]) and neg.min() >= 0 and neg[:, [0, 2]].max() < 50
        same_s, same_o = neg[:, 0] == pos[:, 0], neg[:, 2] == pos[:, 2]
        assert np.all(same_s | same_o)
        if side == "s":
            assert np.all(same_o)
        if side == "o":
            assert np.all(same_s)
    neg = generate_corruptions_for_fit(X, entities_list=[7, 9], eta=50, corrupt_side="o", rnd=np.random.RandomState(0))
    assert set(np.unique(neg[:, 2])) == {7, 9} and np.array_equal(neg[:, 0], np.tile(X[:, 0], 50))
    np.testing.assert_array_equal(generate_corruptions_for_fit(X, eta=2, entities_size=9, rnd=3), generate_corruptions_for_fit(X, eta=2, entities_size=9, rnd=3))


def test_data_format_helpers():
    """reference tests/emgraph/models/test_misc.py:6-28 and tests/emgraph/utils/test_model_utils.py:120-140."""
    import pandas as pd
    from emgraph_b200.utils import dataframe_to_triples, get_entity_triples
    X = np.array([["a", "y", "b"], ["a", "y", "c"], ["c", "y", "a"], ["d", "y", "e"], ["e", "y", "f"], ["f", "y", "c"]])
    np.testing.assert_array_equal(get_entity_triples("c", X), [["a", "y", "c"], ["c", "y", "a"], ["f", "y", "c"]])
    assert get_entity_triples("zz", X).shape == (0, 3)
    df = pd.DataFrame({"species": ["setosa", "virginica"], "sepal_length": [5.1, 6.3], "petal_width": [0.2, 1.8]})
    t = dataframe_to_triples(df, [("species", "has_sepal_length", "sepal_length"), ("species", "has_petal_width", "petal_width")])
    np.testing.assert_array_equal(t[0], ["setosa", "has_sepal_length", "5.1"])
    assert t.shape == (4, 3) and t[3].tolist() == ["virginica", "has_petal_width", "1.8"]
    with pytest.raises(Exception, match="not in data frame headers"):
        dataframe_to_triples(df, [("species", "has_sepal_length", "abc")])


def test_index_training_triples_equals_unique_plus_to_idx():
    """fit()'s one-pass id mapping == the reference's create_mappings + to_idx (sorted-unique ids,
    evaluation/protocol.py:429-445, :662-723) for ASCII, non-ASCII, numeric-looking and object-dtype labels."""
    from emgraph_b200 import models
    from oracle import kge_oracle as ko
    rng = np.random.default_rng(0)
    pools = [np.array(["e%03d" % i for i in range(50)]), np.array(["é", "z", "a", "Ω", "ab", "aB", "10", "9", "", " x"]),
             np.array(["Q1", "Q10", "Q2", "q1", "P31"], dtype=object)]
    for pool in pools:
        for n in (0, 1, 7, 500):
            X = np.stack([pool[rng.integers(0, len(pool), n)], pool[rng.integers(0, min(3, len(pool)), n)],
                          pool[rng.integers(0, len(pool), n)]], 1) if n else np.zeros((0, 3), dtype=pool.dtype)
            ei, ri, Xi = models.index_training_triples(X)
            r2i, e2i = ko.create_mappings(X) if n else ({}, {})
            assert list(ei.labels) == list(e2i.keys()) and list(ri.labels) == list(r2i.keys())
            if n:
                np.testing.assert_array_equal(Xi, ko.to_idx(X, e2i, r2i))
                np.testing.assert_array_equal(Xi, models.to_idx(X, ei, ri))  # later lookups agree with the training ids
            assert Xi.dtype == np.int32 and Xi.shape == (n, 3)


def test_sorted_factorize_is_exact(monkeypatch):
    """The hashing pass behind fit()'s id mapping equals np.unique(return_inverse=True) -- for odd item widths, bytes,
    empty strings, labels differing only in the last character, with and without pandas -- and a hash collision
    (forced here) falls back to the sort instead of merging two labels."""
    import builtins
    from emgraph_b200 import models
    rng = np.random.default_rng(2)
    cases = [np.array(["a", "b", "", "ab", "ba", "aa"])[rng.integers(0, 6, 300)],
             np.array(["entity_%d" % i for i in range(1000)])[rng.integers(0, 1000, 5000)],       # width 10: padded to 12 chars
             np.array([b"x1", b"x2", b"y"])[rng.integers(0, 3, 50)],
             np.array(["Ωmega", "omega", "Omega", "omegb"])[rng.integers(0, 4, 64)],
             np.array(["%033d" % i for i in range(40)])[rng.integers(0, 40, 200)]]
    for v in cases:
        u, inv = np.unique(v, return_inverse=True)
        for no_pandas in (False, True):
            if no_pandas:
                real_import = builtins.__import__
                monkeypatch.setattr(builtins, "__import__", lambda name, *a, **k: (_ for _ in ()).throw(ImportError(name))
                                    if name == "pandas" else real_import(name, *a, **k))
            labels, codes = models._sorted_factorize(v)
            if no_pandas:
                monkeypatch.undo()
            np.testing.assert_array_equal(labels, u)
            np.testing.assert_array_equal(codes, inv.reshape(-1))
            assert labels.dtype == v.dtype
    monkeypatch.setattr(models, "_hash_fixed_width", lambda a: np.zeros(a.shape[0], np.uint64))  # everything collides
    labels, codes = models._sorted_factorize(cases[1])
    u, inv = np.unique(cases[1], return_inverse=True)
    np.testing.assert_array_equal(labels, u)
    np.testing.assert_array_equal(codes, inv.reshape(-1))


def test_label_index_bulk_lookup_is_exact(monkeypatch):
    """LabelIndex.lookup / contains on large inputs take the hashed path: same ids as the binary search, unseen labels
    still raise (evaluation/protocol.py:684-701), wider / narrower query dtypes and truncation look-alikes included."""
    import builtins
    from emgraph_b200 import models
    rng = np.random.default_rng(3)
    labels = np.unique(np.array(["ent%05d" % i for i in rng.integers(0, 90000, 30000)]))
    idx = models.LabelIndex(labels)
    q = labels[rng.integers(0, len(labels), 20000)]
    want = np.searchsorted(labels, q).astype(np.int32)
    for no_pandas in (False, True):
        idx._hash = None
        if no_pandas:
            real_import = builtins.__import__
            monkeypatch.setattr(builtins, "__import__", lambda name, *a, **k: (_ for _ in ()).throw(ImportError(name))
                                if name == "pandas" else real_import(name, *a, **k))
        got = idx.lookup(q, "entities")
        wide = idx.lookup(q.astype("<U20"), "entities")
        has = idx.contains(np.concatenate([q[:5000], np.array(["ent00000x", "zzz", ""])]))
        if no_pandas:
            monkeypatch.undo()
        np.testing.assert_array_equal(got, want)
        np.testing.assert_array_equal(wide, want)
        assert has[:5000].all() and not has[5000:].any()
    assert idx._hashed() is not None
    bad = q.astype("<U20").copy()
    bad[7] = q[7] + "-longer-than-any-label"  # truncating to the index's width would look like a known label
    with pytest.raises(ValueError, match="not present in the training set"):
        idx.lookup(bad, "entities")
    with pytest.raises(ValueError):
        idx.lookup(np.concatenate([q, np.array(["nope"])]), "entities")
    # small inputs and non-string labels keep the binary search
    np.testing.assert_array_equal(idx.lookup(q[:10], "entities"), want[:10])
    ints = models.LabelIndex(np.arange(0, 100000, 3))
    np.testing.assert_array_equal(ints.lookup(np.arange(0, 30000, 3), "entities"), np.arange(10000))
    assert ints._hashed() is None


def test_object_dtype_labels_take_the_fast_path_and_keep_their_ids():
    """Labels arriving as object arrays of str (DataFrame.values) map to the same ids as their fixed-width twins; a
    mixed-type object array keeps numpy's own ordering (no silent stringification)."""
    from emgraph_b200 import models
    rng = np.random.default_rng(5)
    pool = np.array(["n%04d" % i for i in range(3000)])
    Xs = np.stack([pool[rng.integers(0, 3000, 9000)], pool[rng.integers(0, 7, 9000)], pool[rng.integers(0, 3000, 9000)]], 1)
    Xo = Xs.astype(object)
    e1, r1, i1 = models.index_training_triples(Xs)
    e2, r2, i2 = models.index_training_triples(Xo)
    np.testing.assert_array_equal(i1, i2)
    assert list(e1.labels) == list(e2.labels) and e2.labels.dtype.kind == "U"
    np.testing.assert_array_equal(models.to_idx(Xo, e2, r2), i1)          # object queries against a str index
    np.testing.assert_array_equal(models.to_idx(Xo[:20], e1, r1), i1[:20])  # small inputs too
    assert e2.contains(np.array(["n0001", "nope"], dtype=object)).tolist() == [True, False]
    with pytest.raises(ValueError):
        models.to_idx(np.array([["n0001", "n0000", "unknown"]], dtype=object), e2, r2)
    mixed = np.array([[1, "r", 2], [2, "r", 10]], dtype=object)
    e3, _, i3 = models.index_training_triples(mixed)
    assert list(e3.labels) == [1, 2, 10] and i3[:, 2].tolist() == [1, 2]


def test_cached_argument_block_equals_a_fresh_one(monkeypatch):
    """fit()'s device-batch step refreshes a cached kge_train_args block instead of rebuilding it: what reaches the C ABI
    must be byte-identical to a freshly built block at every step -- changing batch (pointer, size), step counter,
    learning rate (sgd schedule), loss slot and flags (pipelining toggled) included."""
    import torch
    from emgraph_b200 import _lib as L
    from emgraph_b200 import engine as en
    from emgraph_b200 import models
    monkeypatch.setattr(en, "_chk_f32", lambda t, n: None)
    monkeypatch.setattr(en, "_chk_i32", lambda t, n: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    seen = []

    class Recorder(en.Engine):
        tdev = torch.device("cpu")

        def __init__(self):
            self.launches = 0

        def train_step(self, a):
            seen.append(bytes(a))

    eng = Recorder()
    monkeypatch.setattr(models, "get_engine", lambda device=None: eng)
    m = models.ComplEx(k=8, eta=3, epochs=1, batches_count=4, seed=5, optimizer="adam", optimizer_params={"lr": 1e-3},
                       embedding_model_params={"negative_corruption_entities": 17})
    f = m._fit_prepare(40, 3)
    X = torch.zeros(100, 3, dtype=torch.int32)
    slots = torch.zeros(6, dtype=torch.float32)
    for i, (lo, hi) in enumerate([(0, 30), (30, 60), (60, 90), (90, 100), (0, 30), (30, 60)]):
        if i == 2:
            f["kw"]["lr"] = 5e-4
        if i == 4:
            f["pipeline"] = False
        pos, out = X[lo:hi], slots[i:i + 1]
        m._fit_step_device(pos, loss_out=out)
        fresh = eng.train_args(ent=f["ent"], rel=f["rel"], pos=pos, loss_out=out, side=0, step=f["step"], **m._step_kw(), **f["st"], **f["neg"])
        assert seen[-1] == bytes(fresh), i
        assert fresh.step == i + 1 and fresh.n_pos == hi - lo and bool(fresh.flags & L.F_PIPELINE) == (i < 4)
    assert len(set(seen)) == 6


def test_cached_host_step_argument_block_equals_a_fresh_one(monkeypatch):
    """Same guarantee for the host-buffer steps (synchronous and ticketed): the refreshed cached block is byte-identical
    to a freshly built one at every call."""
    import torch
    from emgraph_b200 import engine as en
    from emgraph_b200 import models
    monkeypatch.setattr(en, "_chk_f32", lambda t, n: None)
    monkeypatch.setattr(en, "_chk_i32", lambda t, n: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    seen = []

    class Recorder(en.Engine):
        tdev = torch.device("cpu")

        def __init__(self):
            self.launches = 0

        def train_step_host(self, a, pos_host, loss_host):
            a.n_pos = pos_host.shape[0]
            seen.append(bytes(a))

        def train_step_host_async(self, a, pos_host, loss_slot):
            a.n_pos = pos_host.shape[0]
            seen.append(bytes(a))
            return len(seen) % 4

        def train_host_wait(self, ticket):
            pass

    eng = Recorder()
    monkeypatch.setattr(models, "get_engine", lambda device=None: eng)
    m = models.DistMult(k=8, eta=3, epochs=1, batches_count=4, seed=5, optimizer="momentum", optimizer_params={"lr": 1e-3, "momentum": 0.8})
    f = m._fit_prepare(40, 3)
    X = torch.zeros(100, 3, dtype=torch.int32)
    for i, (lo, hi) in enumerate([(0, 30), (30, 60), (90, 100), (0, 30), (30, 60), (60, 90)]):
        if i == 3:
            f["kw"]["lr"] = 5e-4
            f["pipeline"] = False
        (m._fit_step_host if i % 2 else m._fit_step_host_pipelined)(X[lo:hi])
        fresh = eng.train_args(ent=f["ent"], rel=f["rel"], pos=None, loss_out=f["loss_dev"], n_pos=hi - lo, side=0, step=f["step"],
                               **m._step_kw(), **f["st"], **f["neg"])
        fresh.n_pos = hi - lo
        assert seen[-1] == bytes(fresh), i
    assert len(set(seen)) == 6
