"""Oracle self-consistency (CPU): float64 finite-difference gradient checks of every cell and
loss, updater closed forms, top-k semantics.  The reference has no tests of its own
(SURVEY.md §4), so these are the manufactured pins for oracle/sbr_oracle.py."""
import numpy as np
import pytest

from oracle import sbr_oracle as O


def make_batch(rng, B, T, N, K=1, n_extra=0):
    lens = rng.randint(1, T + 1, size=B)
    lens[0] = T
    X = np.zeros((B, T, K), dtype=np.int32)
    mask = np.zeros((B, T))
    for b in range(B):
        X[b, :lens[b], 0] = rng.randint(0, N, size=lens[b])
        for k in range(1, K):
            X[b, :lens[b], k] = N + rng.randint(0, n_extra, size=lens[b])
        mask[b, :lens[b]] = 1
    return X, mask, lens


def loss_kwargs(spec, rng, B, N, X, lens):
    if spec.loss == "CCE":
        return dict(Y=rng.randint(0, N, size=B), pop=rng.uniform(0.5, 2.0, size=B))
    if spec.loss in O.SAMPLING_LOSSES:
        return dict(Y=rng.randint(0, N, size=B), samples=rng.randint(0, N, size=5),
                    pop=rng.uniform(0.5, 2.0, size=B))
    in_seqs = [list(X[b, :lens[b], 0]) for b in range(B)]
    targets = [list(rng.choice(N, size=2, replace=False)) for _ in range(B)]
    Y, W = O.margin_targets(N, in_seqs, targets)
    return dict(Ymat=Y, Wmat=W)


def fd_check(spec, seed=0, B=3, T=4, n_probe=40, eps=1e-6, K=1):
    rng = np.random.RandomState(seed)
    N = spec.n_items
    vals = O.init_params(spec, rng, np.float64)
    # non-trivial biases / inits so every gradient path is exercised
    for v in vals:
        if not v.any():
            v[...] = rng.normal(0, 0.1, size=v.shape)
    X, mask, lens = make_batch(rng, B, T, N, K, spec.n_extra_ids)
    kw = loss_kwargs(spec, rng, B, N, X, lens)
    cost, grads = O.loss_and_grads(spec, vals, X, mask, **kw)
    assert np.isfinite(cost)
    worst = 0.0
    for pi, (v, g) in enumerate(zip(vals, grads)):
        flat = v.reshape(-1)
        idxs = rng.choice(flat.size, size=min(n_probe, flat.size), replace=False)
        if pi == 0 and spec.embedding == 0:
            # make sure some probed W_in rows were actually used
            used = np.unique(X[mask > 0][:, 0])[:4]
            idxs = np.concatenate([idxs, used * v.shape[1]])
        for i in idxs:
            old = flat[i]
            flat[i] = old + eps
            cp, _ = O.loss_and_grads(spec, vals, X, mask, **kw)
            flat[i] = old - eps
            cm, _ = O.loss_and_grads(spec, vals, X, mask, **kw)
            flat[i] = old
            num = (cp - cm) / (2 * eps)
            ana = g.reshape(-1)[i]
            # mixed tolerance: central differences carry ~1e-10 absolute round-off
            err = max(0.0, abs(num - ana) - 2e-9) / max(1e-7, abs(num) + abs(ana))
            worst = max(worst, err)
    return worst


@pytest.mark.parametrize("cell", ["GRU", "LSTM", "Vanilla"])
def test_cce_gradients_single_layer(cell):
    spec = O.Spec(n_items=11, cell=cell, layers=(5,), loss="CCE", regularization=0.01)
    assert fd_check(spec) < 2e-5


@pytest.mark.parametrize("cell", ["GRU", "LSTM"])
def test_cce_gradients_two_layers(cell):
    spec = O.Spec(n_items=9, cell=cell, layers=(4, 3), loss="CCE")
    assert fd_check(spec, seed=1) < 2e-5


@pytest.mark.parametrize("cell", ["GRU", "LSTM"])
def test_embedding_path_gradients(cell):
    spec = O.Spec(n_items=9, cell=cell, layers=(4,), embedding=3, loss="CCE", regularization=-0.02)
    assert fd_check(spec, seed=2) < 2e-5


def test_rating_feature_ids_gradients():
    spec = O.Spec(n_items=9, cell="GRU", layers=(4,), n_extra_ids=10, ids_per_step=2, loss="CCE")
    assert fd_check(spec, seed=3, K=2) < 2e-5


@pytest.mark.parametrize("loss", ["BPR", "BPRI", "TOP1", "Blackout"])
@pytest.mark.parametrize("tanh", [False, True])
def test_sampling_loss_gradients(loss, tanh):
    if loss == "Blackout" and tanh:
        pytest.skip("last_layer_tanh is not used by the Blackout loss (rnn_sampling.py:68-72)")
    spec = O.Spec(n_items=13, cell="GRU", layers=(4,), loss=loss, last_layer_tanh=tanh)
    assert fd_check(spec, seed=4) < 2e-5


@pytest.mark.parametrize("loss", ["hinge", "logit", "logsig"])
def test_margin_loss_gradients(loss):
    spec = O.Spec(n_items=13, cell="LSTM", layers=(4,), loss=loss)
    assert fd_check(spec, seed=5) < 2e-5


def test_masked_rows_keep_state_and_final_is_last_valid():
    """hid_out[-1] equals the state at each row's last valid step (sparse_lstm.py:417-425,485-486)."""
    rng = np.random.RandomState(7)
    spec = O.Spec(n_items=10, cell="LSTM", layers=(6,))
    vals = O.init_params(spec, rng)
    X, mask, lens = make_batch(rng, 4, 6, 10)
    P = O.as_dict(spec, vals)
    h_full, _ = O.forward_stack(spec, P, X, mask)
    for b in range(4):
        L = lens[b]
        h_b, _ = O.forward_stack(spec, P, X[b:b + 1, :L], mask[b:b + 1, :L])
        np.testing.assert_allclose(h_full[b], h_b[0], rtol=0, atol=1e-14)


def test_padding_ids_do_not_matter():
    rng = np.random.RandomState(8)
    spec = O.Spec(n_items=10, cell="GRU", layers=(5,))
    vals = O.init_params(spec, rng)
    X, mask, lens = make_batch(rng, 4, 6, 10)
    kw = dict(Y=rng.randint(0, 10, 4), pop=np.ones(4))
    c0, g0 = O.loss_and_grads(spec, vals, X, mask, **kw)
    X2 = X.copy()
    X2[mask == 0] = 7
    c1, g1 = O.loss_and_grads(spec, vals, X2, mask, **kw)
    assert c0 == c1
    for a, b in zip(g0, g1):
        np.testing.assert_array_equal(a, b)


def test_grad_clip_only_affects_backward():
    rng = np.random.RandomState(9)
    a = O.Spec(n_items=10, cell="GRU", layers=(5,), grad_clip=100.0)
    b = O.Spec(n_items=10, cell="GRU", layers=(5,), grad_clip=1e-4)
    vals = O.init_params(a, rng)
    X, mask, _ = make_batch(rng, 4, 6, 10)
    kw = dict(Y=rng.randint(0, 10, 4), pop=np.ones(4))
    ca, ga = O.loss_and_grads(a, vals, X, mask, **kw)
    cb, gb = O.loss_and_grads(b, vals, X, mask, **kw)
    assert ca == cb
    names = [n for n, _ in O.param_names_shapes(a)]
    assert abs(gb[names.index("l0.W_hid_to_resetgate")]).max() < abs(ga[names.index("l0.W_hid_to_resetgate")]).max()
    # the output layer sits above the clip sites: untouched
    np.testing.assert_array_equal(ga[names.index("out.W")], gb[names.index("out.W")])


def test_param_order_matches_checkpoint_layout():
    spec = O.Spec(n_items=7, cell="GRU", layers=(3,))
    names = [n for n, _ in O.param_names_shapes(spec)]
    assert names[:3] == ["l0.W_in_to_updategate", "l0.W_hid_to_updategate", "l0.b_updategate"]
    assert names[3].endswith("resetgate") and names[6].endswith("hidden_update")
    assert names[-3:] == ["l0.hid_init", "out.W", "out.b"]
    spec = O.Spec(n_items=7, cell="LSTM", layers=(3,), embedding=2)
    ns = O.param_names_shapes(spec)
    assert ns[0] == ("emb.W", (7, 2))
    assert ns[1] == ("l0.W_in_to_ingate", (2, 3))
    assert [n for n, _ in ns][13:18] == ["l0.W_cell_to_ingate", "l0.W_cell_to_forgetgate",
                                          "l0.W_cell_to_outgate", "l0.cell_init", "l0.hid_init"]


def test_adam_first_step_closed_form():
    p = [np.array([1.0, -2.0, 3.0])]
    g = [np.array([0.5, -0.25, 0.0])]
    u = O.Updater("adam", lr=1e-3)
    u.step(p, g)
    # first step: m = .1 g, v = .001 g^2, a = lr*sqrt(.001)/.1 -> p -= lr * g/(|g| + 1e-8*sqrt(1000)...)
    a = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    exp = np.array([1.0, -2.0, 3.0]) - a * (0.1 * g[0]) / (np.sqrt(0.001 * g[0] ** 2) + 1e-8)
    np.testing.assert_allclose(p[0], exp, rtol=1e-14)
    assert p[0][2] == 3.0


def test_adam_moves_rows_with_zero_gradient_after_first_touch():
    """Dense optimiser semantics (SURVEY §7 hard parts): m decays, so a row whose gradient is 0
    this step still moves."""
    p = [np.array([1.0, 1.0])]
    u = O.Updater("adam", lr=1e-2)
    u.step(p, [np.array([1.0, 0.0])])
    before = p[0].copy()
    u.step(p, [np.array([0.0, 0.0])])
    assert p[0][0] != before[0] and p[0][1] == before[1]


@pytest.mark.parametrize("kind", ["adagrad", "rmsprop", "adadelta", "nesterov"])
def test_other_updaters_one_step(kind):
    p = [np.array([1.0, -1.0])]
    g = np.array([0.3, -0.7])
    u = O.Updater(kind, lr=0.1, rho=0.9)
    u.step(p, [g])
    if kind == "adagrad":
        exp = np.array([1.0, -1.0]) - 0.1 * g / np.sqrt(g * g + 1e-6)
    elif kind == "rmsprop":
        exp = np.array([1.0, -1.0]) - 0.1 * g / np.sqrt(0.1 * g * g + 1e-6)
    elif kind == "adadelta":
        exp = np.array([1.0, -1.0]) - 0.1 * g * np.sqrt(1e-6) / np.sqrt(0.1 * g * g + 1e-6)
    else:
        v = -0.1 * g
        exp = np.array([1.0, -1.0]) + 0.9 * v - 0.1 * g
    np.testing.assert_allclose(p[0], exp, rtol=1e-13)


def test_topk_sorted_best_first_and_exclusion_multiplies():
    s = np.array([[0.1, 0.5, 0.3, 0.05, 0.05]])
    np.testing.assert_array_equal(O.top_k(s, 3)[0], [1, 2, 0])
    spec = O.Spec(n_items=5, cell="GRU", layers=(3,), loss="hinge")
    vals = O.init_params(spec, np.random.RandomState(0))
    X = np.array([[[1], [2]]]); mask = np.ones((1, 2))
    ex = np.zeros((1, 5)); ex[0, [1, 2]] = 1
    out = O.test_scores(spec, vals, X, mask, exclude=ex)
    assert out[0, 1] == 0 and out[0, 2] == 0   # margin quirk: excluded -> 0, not -inf


def test_margin_targets_match_reference_fill():
    Y, W = O.margin_targets(10, [[1, 2, 3]], [[4, 5]], balance=2.0)
    w = 2.0 * 2 / (10 - 2 - 3)
    assert W[0, 0] == w and W[0, 4] == -1 and W[0, 1] == 0
    assert Y[0, 4] == 1 and Y[0, 1] == 0 and Y[0, 0] == 0


@pytest.mark.parametrize("cell,layers,emb", [("GRU", (4,), 0), ("LSTM", (4, 3), 0), ("GRU", (3, 3), 3), ("Vanilla", (5,), 0)])
def test_bidirectional_gradients(cell, layers, emb):
    """--r_bi (recurrent_layers.py:72-78): forward + backwards layer per depth, concatenated; finite differences."""
    spec = O.Spec(n_items=9, cell=cell, layers=layers, embedding=emb, loss="CCE", bidirectional=True)
    names = [n for n, _ in O.param_names_shapes(spec)]
    assert "l0b.hid_init" in names and names.index("l0.hid_init") < names.index("l0b.W_in_to_" + ("ingate" if cell == "LSTM" else ("updategate" if cell == "GRU" else "hidden_update")))
    assert dict(O.param_names_shapes(spec))["out.W"] == (2 * layers[-1], 9)
    assert fd_check(spec, seed=6, B=3, T=5) < 2e-5


@pytest.mark.parametrize("layers,emb,bi", [((5, 4), 0, False), ((4,), 3, False), ((4, 3, 3), 0, True)])
def test_dense_vanilla_layers(layers, emb, bi):
    """Vanilla layers with a dense input are Lasagne RecurrentLayers (recurrent_layers.py:98-99): rectifier, parameter
    order hid_init, W_in_to_hid, b, W_hid_to_hid, Uniform(-0.01, 0.01) weights; layer 0 without embedding stays the
    in-tree tanh copy."""
    spec = O.Spec(n_items=9, cell="Vanilla", layers=layers, embedding=emb, loss="CCE", bidirectional=bi)
    names = [n for n, _ in O.param_names_shapes(spec)]
    first_dense = "l0." if emb else "l1."
    i = names.index(first_dense + "hid_init")
    assert names[i:i + 4] == [first_dense + x for x in ("hid_init", "W_in_to_hid", "b", "W_hid_to_hid")]
    if not emb:
        assert "l0.W_in_to_hidden_update" in names and O.dense_vanilla(spec, 0) is False
    vals = O.init_params(spec, np.random.RandomState(0))
    W = dict(zip(names, vals))[first_dense + "W_hid_to_hid"]
    assert np.abs(W).max() <= 0.01 and W.std() > 0.004
    assert fd_check(spec, seed=8, B=3, T=5) < 2e-5


def test_backwards_layer_equals_forward_layer_on_reversed_rows():
    """A backwards layer over left-aligned, right-padded rows computes exactly what a forward layer computes on the rows
    with their valid prefix reversed; its output, un-reversed the same way, is the aligned output.  (This identity is
    how the CUDA path runs bidirectional stacks on its forward-only scan kernels.)"""
    rng = np.random.RandomState(3)
    spec = O.Spec(n_items=12, cell="LSTM", layers=(5,), loss="CCE")
    vals = O.init_params(spec, rng)
    for v in vals:
        if not v.any():
            v[...] = rng.normal(0, 0.1, size=v.shape)
    P = O.as_dict(spec, vals)
    X, mask, lens = make_batch(rng, 4, 6, 12)
    T = X.shape[1]
    maskT = mask.T
    W_in, _, b, _ = O._stack(spec, P, 0)
    Xt = np.transpose(X, (1, 0, 2))
    Xg = W_in[Xt, :].sum(axis=-2) + b
    hs_b, _ = O._layer_forward(spec, 0, P, Xg, maskT, backwards=True)
    Xr = np.zeros_like(X)
    for i in range(X.shape[0]):
        Xr[i, :lens[i]] = X[i, :lens[i]][::-1]
    Xgr = W_in[np.transpose(Xr, (1, 0, 2)), :].sum(axis=-2) + b
    hs_f, _ = O._layer_forward(spec, 0, P, Xgr, maskT, backwards=False)
    for i in range(X.shape[0]):
        np.testing.assert_allclose(hs_b[:lens[i], i], hs_f[:lens[i], i][::-1], rtol=0, atol=1e-15)
        np.testing.assert_allclose(hs_b[0, i], hs_f[lens[i] - 1, i], rtol=0, atol=1e-15)      # final state of the backwards scan
