"""Host-side logic of the row-sharded multi-GPU path, exercised with world_size 2 on the gloo backend
(no GPU): shard ranges, batch dealing, Philox stream bases, and that the owner selection applied to an
all-gathered key list covers every entity slot exactly once (relation slots on every rank)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from emgraph_b200 import distributed as D


def test_shard_ranges_cover_rows_exactly():
    for E in (1, 7, 8, 14541, 4594485):
        for world in (1, 2, 3, 4, 8):
            seen = np.zeros(E, np.int32)
            for r in range(world):
                b, e = D.shard_range(E, world, r)
                assert 0 <= b <= e <= E and e - b <= D.rows_per_shard(E, world)
                seen[b:e] += 1
                if e > b:
                    assert np.all(D.owner_of(np.arange(b, e), E, world) == r)
            assert np.all(seen == 1)


def test_batches_are_disjoint_across_ranks_within_a_step():
    n_total, n_per = 10000, 128
    for world in (2, 4, 8):
        for step in (0, 1, 5, 77):
            sl = [D.batch_slice(n_total, world, r, step, n_per) for r in range(world)]
            assert all(hi - lo == n_per and hi <= n_total for lo, hi in sl)
            assert len({lo for lo, _ in sl}) == world
    # negative index bases never overlap
    assert [D.neg_index_base(r, 20, 128) for r in range(3)] == [0, 2560, 5120]


def test_push_schedule_visits_every_block_once_and_spreads_destinations():
    for S, world in ((67 * 10308, 8), (23 * 4252, 2), (100, 4), (31, 3), (32 * 5, 5)):
        bpr = (S + 31) // 32
        for group in (0, 1, 3, 64):
            firsts = []
            for own in range(world):
                order = D.push_block_order(S, world, own, group)
                assert len(order) == bpr * world and len(set(order)) == bpr * world
                assert all(0 <= r < world and 0 <= b < bpr for r, b in order)
                firsts.append(order[0][0])
            assert sorted(firsts) == list(range(world))  # no two owners start on the same destination
    # default grouping: one destination region at a time
    order = D.push_block_order(320, 4, 1, 0)
    assert [r for r, _ in order] == [2] * 10 + [3] * 10 + [0] * 10 + [1] * 10


def test_grad_tail_layout_alignment():
    for eta, n, K in ((64, 10308, 256), (20, 4252, 400), (5, 3, 8), (7, 33, 10)):
        head, tail, stride = D.grad_tail_layout(eta, n, K)
        assert head == 3 * n * K and stride >= tail and stride % 4 == 0 and stride - tail < 4
        assert tail * 4 >= 2 * n * K * 4 + eta * n * 4 + eta * n  # rows + coefficients + one byte per negative


def _worker(rank, world, port, E, R, eta, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(10 + rank)
        pos = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)
        repl = rng.integers(0, E, eta * n).astype(np.int32)
        # slot keys in the library's layout: subjects | objects | replacements | E + relation
        keys_local = torch.from_numpy(np.concatenate([pos[:, 0], pos[:, 2], repl, E + pos[:, 1]]).astype(np.int32))
        S = (3 + eta) * n
        assert keys_local.numel() == S
        keys_all = torch.empty(S * world, dtype=torch.int32)
        dist.all_gather_into_tensor(keys_all, keys_local)
        ka = keys_all.numpy()
        np.testing.assert_array_equal(ka[rank * S:(rank + 1) * S], keys_local.numpy())
        mine = D.owned_mask(ka, E, world, rank)
        cover = torch.from_numpy(mine.astype(np.int32))
        dist.all_reduce(cover)
        ent_slot = ka < E
        ok = bool(np.all(cover.numpy()[ent_slot] == 1) and np.all(cover.numpy()[~ent_slot] == world))
        # the counters of a sharded ranking sweep add up: per-shard candidate counts -> all-reduce
        b, e = D.shard_range(E, world, rank)
        cnt = torch.tensor([e - b], dtype=torch.int32)
        dist.all_reduce(cnt)
        q.put((rank, ok and int(cnt.item()) == E))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gloo_world2_key_gather_and_owner_selection():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1001, 7, 5, 64, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


# ------------------------------------------------------------------------------------------------
# owner-compute exchange (the round-2 plan): the decomposition must be exact
# ------------------------------------------------------------------------------------------------
import pytest  # noqa: E402

from oracle import kge_oracle as ko  # noqa: E402
from oracle import sharded_oracle as so  # noqa: E402


def _rank_batches(rng, E, R, n, eta, W):
    out = []
    for _ in range(W):
        pos = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)
        out.append((pos, rng.integers(0, 2, n * eta).astype(np.uint8), rng.integers(0, E, n * eta).astype(np.int32)))
    return out


@pytest.mark.parametrize("model,loss,norm,nl", [("DistMult", "nll", 1, "linear"), ("ComplEx", "multiclass_nll", 1, "linear"),
                                                ("HolE", "self_adversarial", 1, "tanh"), ("TransE", "pairwise", 1, "linear"),
                                                ("TransE", "nll", 2, "linear"), ("ComplEx", "absolute_margin", 1, "sigmoid")])
def test_owner_compute_step_equals_the_oracle_step(model, loss, norm, nl):
    """Folded queries shipped to the owners, owner-side scores, per-owner partial sums of c*dF/dQ sent back: the
    summed loss and gradients of W ranks equal the single-process oracle step on the same batches."""
    rng = np.random.default_rng(7)
    E, R, k, eta, n, W = 97, 5, 6, 5, 23, 4
    K = ko.internal_k(model, k)
    ent = (rng.normal(size=(E, K)) * 0.5).astype(np.float32)
    rel = (rng.normal(size=(R, K)) * 0.5).astype(np.float32)
    batches = _rank_batches(rng, E, R, n, eta, W)
    got = so.owner_compute_step(model, k, loss, eta, ent, rel, batches, W, margin=2.0, norm=norm, nl=nl)
    exp_loss, exp_ge, exp_gr = 0.0, np.zeros((E, K)), np.zeros((R, K))
    for pos, keep, repl in batches:
        o = ko.train_step(model, k, loss, eta, ent, rel, pos, keep, repl, margin=2.0, norm=norm, dtype=np.float64, nl=nl)
        exp_loss += float(o["loss"])
        exp_ge += o["grad_ent"]
        exp_gr += o["grad_rel"]
    np.testing.assert_allclose(got["loss"], exp_loss, rtol=1e-10)
    np.testing.assert_allclose(got["grad_ent"], exp_ge, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(got["grad_rel"], exp_gr, rtol=1e-9, atol=1e-11)


def test_owner_compute_volume_against_the_row_push():
    """Per-rank NVLink volume of the two exchanges at a cfg5-like shape (eta = 64, 8 ranks, 1 KiB rows): shipping
    queries and partial sums moves under half of what shipping rows moves (DESIGN.md section 7 sizes round 2 with
    this: ~33 rows per positive against ~74)."""
    rng = np.random.default_rng(3)
    E, R, k, eta, n, W = 20000, 9, 256, 64, 120, 8
    ent = np.zeros((E, k), np.float32)
    rel = np.zeros((R, k), np.float32)
    batches = _rank_batches(rng, E, R, n, eta, W)
    oc = so.owner_compute_step("DistMult", k, "nll", eta, ent, rel, batches, W)
    push = so.push_step_bytes(eta, n, k, W, E, batches)
    row = 4 * k
    assert oc["bytes"]["queries_in"] == 2 * n * (W - 1) * row
    assert oc["bytes"]["partials_in"] <= 2 * n * (W - 1) * row
    assert abs(push["rows_in"] - (2 + eta) * n * row * (W - 1) / W) < 0.05 * push["rows_in"]
    assert 30 * n * row < oc["bytes_total"] < 36 * n * row
    assert 72 * n * row < push["total"] < 76 * n * row


# ------------------------------------------------------------------------------------------------
# dimension-sharded path (the product's multi-GPU exchange): decomposition exact, column bookkeeping, and the
# ShardedKGE driver end to end on 2 gloo ranks with the oracle-backed stand-in engine
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model,loss,norm,nl,k,W", [("DistMult", "nll", 1, "linear", 10, 4), ("ComplEx", "multiclass_nll", 1, "linear", 12, 8),
                                                   ("HolE", "self_adversarial", 1, "tanh", 6, 2), ("TransE", "pairwise", 1, "linear", 9, 3),
                                                   ("TransE", "nll", 2, "linear", 16, 4), ("ComplEx", "absolute_margin", 1, "sigmoid", 5, 2)])
def test_dim_sharded_step_equals_the_oracle_step(model, loss, norm, nl, k, W):
    """Column-slice partial sums, one sum per scored triple over the ranks, slice-local backward: loss and gradients
    of the W slices equal the single-process oracle step (the decomposition is exact, not an approximation)."""
    rng = np.random.default_rng(0)
    E, R, eta, n = 97, 5, 5, 23
    K = ko.internal_k(model, k)
    ent = (rng.normal(size=(E, K)) * 0.5).astype(np.float32)
    rel = (rng.normal(size=(R, K)) * 0.5).astype(np.float32)
    pos = np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32)
    keep = rng.integers(0, 2, n * eta).astype(np.uint8)
    repl = rng.integers(0, E, n * eta).astype(np.int32)
    g = so.dim_sharded_step(model, k, loss, eta, ent, rel, pos, keep, repl, W, margin=2.0, norm=norm, nl=nl)
    o = ko.train_step(model, k, loss, eta, ent, rel, pos, keep, repl, margin=2.0, norm=norm, dtype=np.float64, nl=nl)
    np.testing.assert_allclose(g["loss"], o["loss"], rtol=1e-12)
    np.testing.assert_allclose(g["grad_ent"], o["grad_ent"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(g["grad_rel"], o["grad_rel"], rtol=1e-9, atol=1e-12)
    assert g["bytes_per_rank"] == 4 * (1 + eta) * n  # the whole exchange: one float per scored triple


def test_column_slices_round_trip_and_match_the_oracle_layout():
    rng = np.random.default_rng(1)
    for model in ("TransE", "DistMult", "ComplEx", "HolE"):
        for k, W in ((200, 8), (256, 8), (100, 2), (10, 4), (7, 3), (4, 8), (512, 1)):
            K = ko.internal_k(model, k)
            full = rng.normal(size=(5, K)).astype(np.float32)
            kc = D.dim_width(k, W)
            assert kc % 4 == 0 and kc * W >= k
            parts = [D.slice_columns(full, model, k, W, r) for r in range(W)]
            for r, p in enumerate(parts):
                ref, kc2, cr = so.dim_slice(full, model, k, W, r)
                assert kc2 == kc and cr == D.dim_range(k, W, r)
                np.testing.assert_array_equal(p, ref)
            np.testing.assert_array_equal(D.merge_columns(parts, model, k), full)
            cat = np.concatenate(parts, 1)
            np.testing.assert_array_equal(cat[:, D.merge_index(model, k, W)], full)
    assert D.chunk_bounds(10, 3) == [(0, 4), (4, 7), (7, 10)] and D.chunk_bounds(2, 5) == [(0, 1), (1, 2)]


def _dim_worker(rank, world, port, cfg, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fake_engine import FakeEngine
        fake = FakeEngine()
        D.get_engine = lambda device=None: fake  # the stand-in engine computes every phase with the oracle
        model, loss, k, eta, E, R, n = cfg["model"], cfg["loss"], cfg["k"], cfg["eta"], cfg["E"], cfg["R"], cfg["n"]
        sk = D.ShardedKGE(model, k, eta, loss, "adam", E, R, n, lr=1e-2, seed=11, init_ent=cfg["ent"], init_rel=cfg["rel"], chunks=2)
        losses = []
        for step in range(2):
            pos = torch.from_numpy(cfg["pos"][step][rank])
            losses.append(float(sk.train_step(pos)[0]))
        ent_new, rel_new = sk.gather_entities(), sk.gather_relations()
        fake.filter_build(torch.from_numpy(cfg["filt"]), E, R)
        ranks = sk.rank(torch.from_numpy(cfg["test"]), side=0, strategy=0, filtered=True).numpy()
        rows = sk.row_shard().numpy()
        q.put((rank, dict(losses=losses, ent=ent_new, rel=rel_new, ranks=ranks, rows=rows, range=(sk.row_begin, sk.row_end))))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
@pytest.mark.parametrize("model,loss,k", [("ComplEx", "nll", 10), ("TransE", "pairwise", 8)])
def test_gloo_world2_sharded_driver_matches_single_process_oracle(model, loss, k):
    """ShardedKGE on 2 ranks (gloo, stand-in engine): batch all-gather, the per-chunk all-reduce of partial sums, the
    column all-gather and the all-to-all that builds the ranking shards -- two optimisation steps and a filtered
    ranking equal the single-process oracle on the concatenated global batches and the same Philox stream."""
    world, E, R, eta, n = 2, 61, 4, 3, 17
    rng = np.random.default_rng(5)
    K = ko.internal_k(model, k)
    ent = (rng.normal(size=(E, K)) * 0.4).astype(np.float32)
    rel = (rng.normal(size=(R, K)) * 0.4).astype(np.float32)
    pos = [[np.stack([rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)], 1).astype(np.int32) for _ in range(world)]
           for _ in range(2)]
    filt = ko.synthetic_triples(E, R, 300, seed=5)
    test = filt[:12]
    cfg = dict(model=model, loss=loss, k=k, eta=eta, E=E, R=R, n=n, ent=ent, rel=rel, pos=pos, filt=filt, test=test)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + ((os.getpid() * 7 + k) % 2000)
    procs = [ctx.Process(target=_dim_worker, args=(r, world, port, cfg, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=200) for _ in procs)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    e_o, r_o, state = ent, rel, None
    for step in range(2):
        P = np.concatenate(pos[step], 0)
        repl, keep = ko.draw_corruptions(11, step + 1, P.shape[0], eta, E, "s,o")
        o = ko.train_step(model, k, loss, eta, e_o, r_o, P, keep, repl, opt="adam", lr=1e-2, step=step + 1,
                          state=state if state is not None else ((np.zeros_like(ent), np.zeros_like(ent)), (np.zeros_like(rel), np.zeros_like(rel))))
        e_o, r_o, state = o["ent_new"], o["rel_new"], (o["state_ent"], o["state_rel"])
        for r in range(world):
            np.testing.assert_allclose(res[r]["losses"][step], o["loss"], rtol=1e-5)
    exp = ko.ranks(model, k, res[0]["ent"], res[0]["rel"], test, filt, "s,o", "worst")
    for r in range(world):
        np.testing.assert_array_equal(res[r]["ent"], res[0]["ent"])
        big = np.abs(state[0][0]) > 1e-4
        np.testing.assert_allclose(res[r]["ent"][big], e_o[big], rtol=2e-4, atol=2e-6)
        np.testing.assert_allclose(res[r]["rel"], r_o, rtol=2e-3, atol=2e-5)
        b, e = res[r]["range"]
        np.testing.assert_array_equal(res[r]["rows"][:e - b], res[0]["ent"][b:e])  # the transposed shard holds whole rows
        np.testing.assert_array_equal(res[r]["ranks"], exp)


# ------------------------------------------------------------------------------------------------
# fit(engine_params={"n_gpus": 2}) under an initialised process group + the sharded optimizer-state checkpoint
# ------------------------------------------------------------------------------------------------
def _fit_worker(rank, world, port, tmp, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fake_engine import FakeEngine
        from emgraph_b200 import models, utils
        fake = FakeEngine()
        D.get_engine = lambda device=None: fake
        models.get_engine = lambda device=None: fake
        from toy_graph import TOY_QUERY, TOY_X
        kw = dict(k=10, eta=2, batches_count=2, seed=555, optimizer="adam", optimizer_params={"lr": 0.05}, loss="nll",
                  engine_params={"n_gpus": world})
        m = models.ComplEx(epochs=4, **kw)
        m.fit(TOY_X)
        y4 = m.predict(TOY_QUERY)
        # 2 + 2 epochs through the sharded checkpoint == 4 epochs
        a = models.ComplEx(epochs=2, **kw)
        a.fit(TOY_X)
        path = os.path.join(tmp, "m.pkl")
        utils.save_model(a, path, save_optimizer_state=True)
        shard_files = sorted(f for f in os.listdir(tmp) if ".opt." in f)
        b = utils.restore_model(path)
        np.testing.assert_array_equal(b._opt_state["ent_v"].numpy(), a._opt_state["ent_v"].numpy())  # merged shards == gathered state
        b.engine_params = {"n_gpus": world, "resume": True}
        b.fit(TOY_X)
        q.put((rank, dict(ent=m.trained_model_params[0], rel=m.trained_model_params[1], y4=y4, y22=b.predict(TOY_QUERY),
                          ent22=b.trained_model_params[0], losses=m.loss_history, shard_files=shard_files, step=b._opt_step)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gloo_world2_fit_n_gpus_equals_the_fit_emulation_and_resumes_from_sharded_state(tmp_path):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + ((os.getpid() * 13) % 2000)
    procs = [ctx.Process(target=_fit_worker, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=250) for _ in procs)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    from toy_graph import TOY_QUERY, TOY_X
    r2i, e2i = ko.create_mappings(TOY_X)
    Xi = ko.to_idx(TOY_X, e2i, r2i)
    ent, rel, losses = ko.fit_emulation("ComplEx", 10, 2, 4, 2, 555, "nll", "adam", 0.05, Xi, len(e2i), len(r2i))
    y = ko.score("ComplEx", 10, ent, rel, ko.to_idx(TOY_QUERY, e2i, r2i))
    for r in range(world):
        np.testing.assert_allclose(res[r]["ent"], ent, rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(res[r]["y4"], y, rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(res[r]["y22"], res[r]["y4"], rtol=1e-5, atol=1e-6)  # resumed == uninterrupted
        np.testing.assert_allclose(res[r]["ent22"], res[r]["ent"], rtol=1e-5, atol=1e-6)
        assert res[r]["step"] == 8 and res[r]["shard_files"] == ["m.pkl.opt.0-of-2.npz", "m.pkl.opt.1-of-2.npz"]
    np.testing.assert_array_equal(res[0]["ent"], res[1]["ent"])
