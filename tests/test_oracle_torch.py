"""Independent second derivation of the oracle's gradients with torch.autograd (CPU, float64).

The forward graph below is written straight from the reference's step functions
(sparse_lstm.py:377-425, :764-805) with a custom grad_clip op (theano.gradient.grad_clip:
identity forward, elementwise clamp of the incoming gradient backward), so it pins the
clip sites that finite differences cannot see."""
import numpy as np
import pytest
import torch

from oracle import sbr_oracle as O
from tests.test_oracle import make_batch


class GradClip(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bound):
        ctx.bound = bound
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.clamp(-ctx.bound, ctx.bound), None


def torch_cost(spec, vals, X, mask, Y, pop):
    P = {n: torch.tensor(v, dtype=torch.float64, requires_grad=True)
         for (n, _), v in zip(O.param_names_shapes(spec), vals)}
    H = spec.layers[0]
    gc = spec.grad_clip
    Xt = torch.tensor(X).long()
    B, T, K = Xt.shape
    m = torch.tensor(mask, dtype=torch.float64)
    if spec.cell == "LSTM":
        order = ["ingate", "forgetgate", "cell", "outgate"]
    else:
        order = ["resetgate", "updategate", "hidden_update"]
    W_in = torch.cat([P["l0.W_in_to_" + g] for g in order], 1)
    W_hid = torch.cat([P["l0.W_hid_to_" + g] for g in order], 1)
    b = torch.cat([P["l0.b_" + g] for g in order], 0)
    Xg = W_in[Xt].sum(-2) + b                      # [B,T,GH]
    h = P["l0.hid_init"].expand(B, H)
    if spec.cell == "LSTM":
        c = P["l0.cell_init"].expand(B, H)
    for t in range(T):
        mt = m[:, t:t + 1]
        if spec.cell == "LSTM":
            gates = GradClip.apply(Xg[:, t] + h @ W_hid, gc)
            i = torch.sigmoid(gates[:, :H] + c * P["l0.W_cell_to_ingate"])
            f = torch.sigmoid(gates[:, H:2 * H] + c * P["l0.W_cell_to_forgetgate"])
            g = torch.tanh(gates[:, 2 * H:3 * H])
            cn = f * c + i * g
            o = torch.sigmoid(gates[:, 3 * H:] + cn * P["l0.W_cell_to_outgate"])
            hn = o * torch.tanh(cn)
            c = torch.where(mt > 0, cn, c)
            h = torch.where(mt > 0, hn, h)
        else:
            hid_in = GradClip.apply(h @ W_hid, gc)
            x = GradClip.apply(Xg[:, t], gc)
            r = torch.sigmoid(hid_in[:, :H] + x[:, :H])
            u = torch.sigmoid(hid_in[:, H:2 * H] + x[:, H:2 * H])
            q = GradClip.apply(x[:, 2 * H:] + r * hid_in[:, 2 * H:], gc)
            hn = (1 - u) * h + u * torch.tanh(q)
            h = torch.where(mt > 0, hn, h)
    z = h @ P["out.W"] + P["out.b"]
    logp = torch.log_softmax(z, dim=1)
    cost = (-logp[torch.arange(B), torch.tensor(Y).long()] / torch.tensor(pop)).mean()
    cost.backward()
    return cost.item(), [P[n].grad.numpy() if P[n].grad is not None else np.zeros(s)
                         for n, s in O.param_names_shapes(spec)]


@pytest.mark.parametrize("cell", ["GRU", "LSTM"])
@pytest.mark.parametrize("gc", [100.0, 2e-3])
def test_oracle_matches_torch_autograd(cell, gc):
    rng = np.random.RandomState(11)
    spec = O.Spec(n_items=17, cell=cell, layers=(6,), grad_clip=gc)
    vals = O.init_params(spec, rng)
    for v in vals:
        if not v.any():
            v[...] = rng.normal(0, 0.1, size=v.shape)
    X, mask, _ = make_batch(rng, 5, 7, 17)
    Y = rng.randint(0, 17, 5)
    pop = rng.uniform(0.5, 2, 5)
    c0, g0 = O.loss_and_grads(spec, vals, X, mask, Y=Y, pop=pop)
    c1, g1 = torch_cost(spec, vals, X, mask, Y, pop)
    assert abs(c0 - c1) < 1e-12
    if gc < 1:
        # make sure the clip really bites in this configuration
        s100 = O.Spec(n_items=17, cell=cell, layers=(6,), grad_clip=100.0)
        _, gfree = O.loss_and_grads(s100, vals, X, mask, Y=Y, pop=pop)
        assert max(abs(a - b).max() for a, b in zip(gfree, g0)) > 1e-6
    for (n, _), a, b in zip(O.param_names_shapes(spec), g0, g1):
        np.testing.assert_allclose(a, b.reshape(a.shape), rtol=1e-9, atol=1e-13, err_msg=n)


# ------------------------------------------------------------------------------------------------------------
# Losses and updaters: a second derivation that does not share code with the oracle.  The torch graphs below are
# written from the reference's own expressions (rnn_sampling.py:68-91,137; sparse_lstm.py:41-54; rnn_margin.py:61-68,
# 109) on top of a hidden state h; torch.autograd supplies the gradients.  The updaters are torch.optim where torch
# implements the same rule (Adagrad, RMSprop, Adadelta, SGD with Nesterov momentum match lasagne up to the documented
# reparametrisations) and a hand-written restatement of lasagne.updates.adam (its bias correction is folded into the
# step size, which torch.optim.Adam does differently: eps is applied after the correction there).
# ------------------------------------------------------------------------------------------------------------
def _torch_head_sampling(spec, h, W, b, Y, samples, pop):
    B = h.shape[0]
    cells = torch.cat([torch.tensor(Y).long(), torch.tensor(samples).long()])
    pred = h @ W[:, cells] + b[cells]                      # BlackoutLayer, sparse_lstm.py:49-54
    if spec.loss == "Blackout":
        p = torch.softmax(pred, dim=1)
        pos = -torch.log(p[torch.arange(B), torch.arange(B)])
        neg = torch.log(1 - p)
        loss = pos - neg[:, B:].sum(dim=-1)
    else:
        if spec.last_layer_tanh:
            pred = torch.tanh(pred)
        diff = (pred - torch.diag(pred)[:, None])[:, B:]
        if spec.loss == "BPR":
            loss = -torch.log(torch.sigmoid(-diff)).mean(dim=-1)
        elif spec.loss == "BPRI":
            loss = torch.log(torch.sigmoid(diff)).mean(dim=-1)
        else:
            loss = (torch.sigmoid(diff) + torch.sigmoid(pred[:, B:] ** 2)).mean(dim=-1)
    return (loss / torch.tensor(pop)).mean()


def _torch_head_margin(spec, h, W, b, Ymat, Wmat):
    pred = h @ W + b
    Yt, Wt = torch.tensor(Ymat), torch.tensor(Wmat)
    if spec.loss == "hinge":
        x = (pred - Yt) * Wt
        loss = (0.5 * (x + x.abs())).sum(dim=-1)           # T.nnet.relu = 0.5 (x + |x|)
    elif spec.loss == "logit":
        loss = (torch.sigmoid(pred - Yt) * Wt).sum(dim=-1)
    else:
        loss = -torch.log(torch.sigmoid((Yt - pred) * Wt)).sum(dim=-1)
    return loss.mean()


@pytest.mark.parametrize("loss,tanh", [("BPR", False), ("BPR", True), ("BPRI", False), ("TOP1", True), ("Blackout", False)])
def test_sampling_losses_match_torch_autograd(loss, tanh):
    rng = np.random.RandomState(5)
    N, H, B, S = 23, 7, 6, 4
    spec = O.Spec(n_items=N, cell="GRU", layers=(H,), loss=loss, last_layer_tanh=tanh)
    P = {"out.W": rng.normal(0, 0.4, (H, N)), "out.b": rng.normal(0, 0.2, N)}
    h = rng.normal(0, 0.8, (B, H))
    Y = rng.randint(0, N, B)
    samples = rng.randint(0, N, S)
    samples[0] = Y[2]                                     # a sample colliding with a target: no filtering in the reference
    pop = rng.uniform(0.5, 2, B)
    c0, dh0, dW0, db0 = O.sampling_loss(spec, P, h, Y, samples, pop)
    ht = torch.tensor(h, requires_grad=True)
    Wt = torch.tensor(P["out.W"], requires_grad=True)
    bt = torch.tensor(P["out.b"], requires_grad=True)
    c1 = _torch_head_sampling(spec, ht, Wt, bt, Y, samples, pop)
    c1.backward()
    assert abs(float(c0) - c1.item()) < 1e-12
    np.testing.assert_allclose(dh0, ht.grad.numpy(), rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(dW0, Wt.grad.numpy(), rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(db0, bt.grad.numpy(), rtol=1e-9, atol=1e-13)


@pytest.mark.parametrize("loss", ["hinge", "logit", "logsig"])
def test_margin_losses_match_torch_autograd(loss):
    rng = np.random.RandomState(6)
    N, H, B = 19, 5, 4
    spec = O.Spec(n_items=N, cell="GRU", layers=(H,), loss=loss)
    P = {"out.W": rng.normal(0, 0.5, (H, N)), "out.b": rng.normal(0, 0.2, N)}
    h = rng.normal(0, 0.8, (B, H))
    in_seqs = [list(rng.choice(N, size=3 + b, replace=False)) for b in range(B)]
    targets = [list(rng.choice(N, size=1 + b % 2, replace=False)) for b in range(B)]
    Ymat, Wmat = O.margin_targets(N, in_seqs, targets)
    c0, dh0, dW0, db0 = O.margin_loss(spec, P, h, Ymat, Wmat)
    ht = torch.tensor(h, requires_grad=True)
    Wt = torch.tensor(P["out.W"], requires_grad=True)
    bt = torch.tensor(P["out.b"], requires_grad=True)
    c1 = _torch_head_margin(spec, ht, Wt, bt, Ymat, Wmat)
    c1.backward()
    assert abs(float(c0) - c1.item()) < 1e-12
    np.testing.assert_allclose(dh0, ht.grad.numpy(), rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(dW0, Wt.grad.numpy(), rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(db0, bt.grad.numpy(), rtol=1e-9, atol=1e-13)


def _lasagne_adam_by_hand(p, grads, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """lasagne.updates.adam restated independently of the oracle: one shared t, the bias
    correction folded into the step size a_t, eps added to sqrt(v) WITHOUT correction."""
    m = np.zeros_like(p); v = np.zeros_like(p)
    for t, g in enumerate(grads, 1):
        a_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        p = p - a_t * m / (np.sqrt(v) + eps)
    return p


@pytest.mark.parametrize("kind", ["adam", "adagrad", "rmsprop", "adadelta", "nesterov"])
def test_updaters_match_an_independent_implementation(kind):
    rng = np.random.RandomState(8)
    p0 = rng.normal(0, 1, (4, 3))
    grads = [rng.normal(0, 1, (4, 3)) for _ in range(6)]
    lr, rho = 0.05, 0.9
    upd = O.Updater(kind, lr=lr, rho=rho)
    p = [p0.copy()]
    for g in grads:
        upd.step(p, [g])
    if kind == "adam":
        ref = _lasagne_adam_by_hand(p0.copy(), grads, lr=lr)
    else:
        pt = torch.tensor(p0.copy(), requires_grad=True)
        if kind == "adagrad":       # lasagne: p -= lr g / sqrt(acc + eps); torch: lr g / (sqrt(acc) + eps) -> eps -> 0 on both sides
            opt = torch.optim.Adagrad([pt], lr=lr, eps=0.0)
        elif kind == "rmsprop":
            opt = torch.optim.RMSprop([pt], lr=lr, alpha=rho, eps=0.0)
        elif kind == "adadelta":
            opt = torch.optim.Adadelta([pt], lr=lr, rho=rho, eps=1e-6)
        else:                       # lasagne nesterov_momentum: v = mu v - lr g; p += mu v - lr g  == SGD(nesterov) on the scaled velocity
            opt = torch.optim.SGD([pt], lr=lr, momentum=rho, nesterov=True)
        for g in grads:
            opt.zero_grad()
            pt.grad = torch.tensor(g)
            opt.step()
        ref = pt.detach().numpy()
    # adagrad / rmsprop place eps INSIDE the square root in lasagne (sqrt(acc + 1e-6)) and outside in torch (here 0):
    # elements whose accumulator is ~1e-6 differ at the 1e-4 level, everything else agrees to rounding
    tol = 1e-3 if kind in ("adagrad", "rmsprop") else 1e-12
    np.testing.assert_allclose(p[0], ref, rtol=tol, atol=tol)
    if kind in ("adagrad", "rmsprop"):
        # exact check against the documented lasagne rule, restated here without the oracle
        q, acc = p0.copy(), np.zeros_like(p0)
        for g in grads:
            acc = acc + g * g if kind == "adagrad" else rho * acc + (1 - rho) * g * g
            q = q - lr * g / np.sqrt(acc + 1e-6)
        np.testing.assert_allclose(p[0], q, rtol=1e-12, atol=1e-12)


# --------------------------------------------------------------------------------------------------------------
# Bidirectional stacks against torch.nn.GRU / torch.nn.LSTM over packed sequences: an implementation that shares
# nothing with the oracle's go-backwards scan pins what `--r_bi` means on right-padded rows (recurrent_layers.py:72-78,
# Lasagne `backwards=True` + mask): per depth [forward | backward] features, final state = forward state after the
# last item | backward state after the first item.
# --------------------------------------------------------------------------------------------------------------
def _torch_rnn_from_oracle(spec, P):
    """torch module carrying the oracle's weights.  GRU: torch's z gate is 1 - update gate (negated weights), its
    candidate bias inside the reset product is zero; LSTM: no peepholes (the test zeroes them)."""
    H_list, nd = spec.layers, 2
    mods = []
    n_in = spec.n_in
    for li, H in enumerate(H_list):
        cls = torch.nn.GRU if spec.cell == "GRU" else torch.nn.LSTM
        m = cls(n_in, H, num_layers=1, batch_first=True, bidirectional=True).double()
        for d, pre in enumerate(O.layer_prefixes(spec, li)):
            sfx = "_reverse" if d else ""
            g = lambda name: torch.tensor(P[pre + name])
            if spec.cell == "GRU":
                w_ih = torch.cat([g("W_in_to_resetgate"), -g("W_in_to_updategate"), g("W_in_to_hidden_update")], 1).T
                w_hh = torch.cat([g("W_hid_to_resetgate"), -g("W_hid_to_updategate"), g("W_hid_to_hidden_update")], 1).T
                b_ih = torch.cat([g("b_resetgate"), -g("b_updategate"), g("b_hidden_update")])
            else:
                names = ["ingate", "forgetgate", "cell", "outgate"]
                w_ih = torch.cat([g("W_in_to_" + n) for n in names], 1).T
                w_hh = torch.cat([g("W_hid_to_" + n) for n in names], 1).T
                b_ih = torch.cat([g("b_" + n) for n in names])
            with torch.no_grad():
                getattr(m, "weight_ih_l0" + sfx).copy_(w_ih)
                getattr(m, "weight_hh_l0" + sfx).copy_(w_hh)
                getattr(m, "bias_ih_l0" + sfx).copy_(b_ih)
                getattr(m, "bias_hh_l0" + sfx).zero_()
        mods.append(m)
        n_in = nd * H
    return mods


@pytest.mark.parametrize("cell,layers", [("GRU", (5,)), ("LSTM", (4,)), ("GRU", (4, 3)), ("LSTM", (3, 4))])
def test_bidirectional_stack_matches_torch_packed_rnn(cell, layers):
    rng = np.random.RandomState(21)
    spec = O.Spec(n_items=13, cell=cell, layers=layers, loss="CCE", bidirectional=True)
    vals = O.init_params(spec, rng)
    for (name, _), v in zip(O.param_names_shapes(spec), vals):
        if "W_cell_to" in name:
            v[...] = 0.0                                  # torch's LSTM has no peepholes
        elif not v.any():
            v[...] = rng.normal(0, 0.2, size=v.shape)     # learned initial states and biases away from zero
    P = O.as_dict(spec, vals)
    X, mask, lens = make_batch(rng, 6, 7, 13)
    lens = np.maximum(lens, 1)
    mask = (np.arange(7)[None, :] < lens[:, None]).astype(np.float64)
    h_last, _ = O.forward_stack(spec, P, X, mask)

    mods = _torch_rnn_from_oracle(spec, P)
    B = X.shape[0]
    inp = torch.nn.functional.one_hot(torch.tensor(X[:, :, 0]).long(), spec.n_in).double()
    for li, m in enumerate(mods):
        H = spec.layers[li]
        pres = O.layer_prefixes(spec, li)
        h0 = torch.stack([torch.tensor(P[p + "hid_init"]).expand(B, H) for p in pres]).contiguous()
        packed = torch.nn.utils.rnn.pack_padded_sequence(inp, torch.tensor(lens), batch_first=True, enforce_sorted=False)
        if cell == "LSTM":
            c0 = torch.stack([torch.tensor(P[p + "cell_init"]).expand(B, H) for p in pres]).contiguous()
            out, (hn, _) = m(packed, (h0, c0))
        else:
            out, hn = m(packed, h0)
        inp, _ = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=X.shape[1])
    ref = torch.cat([hn[0], hn[1]], dim=1).detach().numpy()
    np.testing.assert_allclose(h_last, ref, rtol=0, atol=1e-10)


def test_stacked_vanilla_matches_torch_rnn_tanh_then_relu():
    """A Vanilla stack mixes the in-tree tanh layer (layer 0, sparse input, sparse_lstm.py:960-961) and Lasagne's
    RecurrentLayer (rectifier) above it (recurrent_layers.py:98-99): torch.nn.RNN('tanh') feeding torch.nn.RNN('relu')."""
    rng = np.random.RandomState(22)
    spec = O.Spec(n_items=11, cell="Vanilla", layers=(5, 4), loss="CCE")
    vals = O.init_params(spec, rng)
    for v in vals:
        if not v.any():
            v[...] = rng.normal(0, 0.2, size=v.shape)
        elif np.abs(v).max() <= 0.011:
            v[...] = rng.normal(0, 0.3, size=v.shape)     # Uniform(-0.01, 0.01) would leave the rectifier barely exercised
    P = O.as_dict(spec, vals)
    X, mask, lens = make_batch(rng, 5, 6, 11)
    lens = np.maximum(lens, 1)
    mask = (np.arange(6)[None, :] < lens[:, None]).astype(np.float64)
    h_last, _ = O.forward_stack(spec, P, X, mask)
    B = X.shape[0]
    inp = torch.nn.functional.one_hot(torch.tensor(X[:, :, 0]).long(), spec.n_in).double()
    layer_params = [("l0.W_in_to_hidden_update", "l0.W_hid_to_hidden_update", "l0.b_hidden_update", "l0.hid_init", "tanh"),
                    ("l1.W_in_to_hid", "l1.W_hid_to_hid", "l1.b", "l1.hid_init", "relu")]
    n_in = spec.n_in
    for (w_in, w_hid, b, h_init, nl), H in zip(layer_params, spec.layers):
        m = torch.nn.RNN(n_in, H, nonlinearity=nl, batch_first=True).double()
        with torch.no_grad():
            m.weight_ih_l0.copy_(torch.tensor(P[w_in]).T)
            m.weight_hh_l0.copy_(torch.tensor(P[w_hid]).T)
            m.bias_ih_l0.copy_(torch.tensor(P[b]))
            m.bias_hh_l0.zero_()
        packed = torch.nn.utils.rnn.pack_padded_sequence(inp, torch.tensor(lens), batch_first=True, enforce_sorted=False)
        out, hn = m(packed, torch.tensor(P[h_init]).expand(B, H).unsqueeze(0).contiguous())
        inp, _ = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=X.shape[1])
        n_in = H
    np.testing.assert_allclose(h_last, hn[0].detach().numpy(), rtol=0, atol=1e-10)
    assert (h_last == 0).any() and (h_last > 0).any()          # the rectifier really clips some units


@pytest.mark.parametrize("cell,layers", [("GRU", (4, 3)), ("LSTM", (3,))])
def test_bidirectional_gradients_match_torch_packed_rnn(cell, layers):
    """Cost and every recurrent / output gradient of a bidirectional CCE step against torch.autograd through
    torch.nn.GRU / LSTM over packed sequences (grad_clip far away; peepholes zero for the LSTM)."""
    rng = np.random.RandomState(23)
    spec = O.Spec(n_items=11, cell=cell, layers=layers, loss="CCE", bidirectional=True, grad_clip=1e6)
    vals = O.init_params(spec, rng)
    for (name, _), v in zip(O.param_names_shapes(spec), vals):
        if "W_cell_to" in name:
            v[...] = 0.0
        elif not v.any():
            v[...] = rng.normal(0, 0.2, size=v.shape)
    P = O.as_dict(spec, vals)
    X, mask, lens = make_batch(rng, 5, 6, 11)
    lens = np.maximum(lens, 1)
    mask = (np.arange(6)[None, :] < lens[:, None]).astype(np.float64)
    B = X.shape[0]
    Y = rng.randint(0, 11, B)
    pop = rng.uniform(0.5, 2.0, B)
    cost, grads = O.loss_and_grads(spec, vals, X, mask, Y=Y, pop=pop)
    G = dict(zip([n for n, _ in O.param_names_shapes(spec)], grads))

    mods = _torch_rnn_from_oracle(spec, P)
    for m in mods:
        for p_ in m.parameters():
            p_.requires_grad_(True)
    W_out = torch.tensor(P["out.W"], requires_grad=True)
    b_out = torch.tensor(P["out.b"], requires_grad=True)
    inits = {}
    inp = torch.nn.functional.one_hot(torch.tensor(X[:, :, 0]).long(), spec.n_in).double()
    for li, m in enumerate(mods):
        H = spec.layers[li]
        pres = O.layer_prefixes(spec, li)
        for p_ in pres:
            inits[p_ + "hid_init"] = torch.tensor(P[p_ + "hid_init"], requires_grad=True)
            if cell == "LSTM":
                inits[p_ + "cell_init"] = torch.tensor(P[p_ + "cell_init"], requires_grad=True)
        h0 = torch.stack([inits[p_ + "hid_init"].expand(B, H) for p_ in pres]).contiguous()
        packed = torch.nn.utils.rnn.pack_padded_sequence(inp, torch.tensor(lens), batch_first=True, enforce_sorted=False)
        if cell == "LSTM":
            c0 = torch.stack([inits[p_ + "cell_init"].expand(B, H) for p_ in pres]).contiguous()
            out, (hn, _) = m(packed, (h0, c0))
        else:
            out, hn = m(packed, h0)
        inp, _ = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=X.shape[1])
    h = torch.cat([hn[0], hn[1]], dim=1)
    logp = torch.log_softmax(h @ W_out + b_out, dim=1)
    tcost = (-logp[torch.arange(B), torch.tensor(Y).long()] / torch.tensor(pop)).mean()
    tcost.backward()
    assert abs(tcost.item() - cost) < 1e-10
    np.testing.assert_allclose(G["out.W"], W_out.grad.numpy(), atol=1e-9)
    np.testing.assert_allclose(G["out.b"], b_out.grad.numpy(), atol=1e-9)
    for li, m in enumerate(mods):
        H = spec.layers[li]
        for d, pre in enumerate(O.layer_prefixes(spec, li)):
            sfx = "_reverse" if d else ""
            g_hh = getattr(m, "weight_hh_l0" + sfx).grad.numpy().T          # [H, G*H], torch gate order
            g_ih = getattr(m, "weight_ih_l0" + sfx).grad.numpy().T
            g_b = getattr(m, "bias_ih_l0" + sfx).grad.numpy()
            if cell == "GRU":
                names, signs = ["resetgate", "updategate", "hidden_update"], [1.0, -1.0, 1.0]
            else:
                names, signs = ["ingate", "forgetgate", "cell", "outgate"], [1.0, 1.0, 1.0, 1.0]
            for gi, (n, sg) in enumerate(zip(names, signs)):
                np.testing.assert_allclose(G[pre + "W_hid_to_" + n], sg * g_hh[:, gi * H:(gi + 1) * H], atol=1e-9, err_msg=pre + n)
                np.testing.assert_allclose(G[pre + "W_in_to_" + n], sg * g_ih[:, gi * H:(gi + 1) * H], atol=1e-9, err_msg=pre + n)
                np.testing.assert_allclose(G[pre + "b_" + n], sg * g_b[gi * H:(gi + 1) * H], atol=1e-9, err_msg=pre + n)
            np.testing.assert_allclose(G[pre + "hid_init"], inits[pre + "hid_init"].grad.numpy(), atol=1e-9)
            if cell == "LSTM":
                np.testing.assert_allclose(G[pre + "cell_init"], inits[pre + "cell_init"].grad.numpy(), atol=1e-9)


def test_embedding_rating_ids_and_stack_match_torch():
    """--r_emb with two ids per step (item + rating feature, rnn_base.py:578-593): EmbeddingLayer + flatten(outdim=3)
    (recurrent_layers.py:47) feeding a two-layer GRU -- torch.nn.Embedding, concatenation, two torch.nn.GRU modules."""
    rng = np.random.RandomState(24)
    spec = O.Spec(n_items=9, cell="GRU", layers=(4, 3), loss="CCE", embedding=3, n_extra_ids=5, ids_per_step=2, grad_clip=1e6)
    vals = O.init_params(spec, rng)
    for v in vals:
        if not v.any():
            v[...] = rng.normal(0, 0.2, size=v.shape)
    P = O.as_dict(spec, vals)
    X, mask, lens = make_batch(rng, 5, 6, 9, 2, 5)
    lens = np.maximum(lens, 1)
    mask = (np.arange(6)[None, :] < lens[:, None]).astype(np.float64)
    B, T, K = X.shape
    Y = rng.randint(0, 9, B)
    pop = rng.uniform(0.5, 2.0, B)
    cost, grads = O.loss_and_grads(spec, vals, X, mask, Y=Y, pop=pop)
    G = dict(zip([n for n, _ in O.param_names_shapes(spec)], grads))

    emb = torch.tensor(P["emb.W"], requires_grad=True)
    inp = emb[torch.tensor(X).long()].reshape(B, T, K * spec.embedding)
    n_in = K * spec.embedding
    hn = None
    for li, H in enumerate(spec.layers):
        pre = "l%d." % li
        g = lambda name: torch.tensor(P[pre + name])
        m = torch.nn.GRU(n_in, H, batch_first=True).double()
        with torch.no_grad():
            m.weight_ih_l0.copy_(torch.cat([g("W_in_to_resetgate"), -g("W_in_to_updategate"), g("W_in_to_hidden_update")], 1).T)
            m.weight_hh_l0.copy_(torch.cat([g("W_hid_to_resetgate"), -g("W_hid_to_updategate"), g("W_hid_to_hidden_update")], 1).T)
            m.bias_ih_l0.copy_(torch.cat([g("b_resetgate"), -g("b_updategate"), g("b_hidden_update")]))
            m.bias_hh_l0.zero_()
        packed = torch.nn.utils.rnn.pack_padded_sequence(inp, torch.tensor(lens), batch_first=True, enforce_sorted=False)
        out, hn = m(packed, g("hid_init").expand(B, H).unsqueeze(0).contiguous())
        inp, _ = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=T)
        n_in = H
    logp = torch.log_softmax(hn[0] @ torch.tensor(P["out.W"]) + torch.tensor(P["out.b"]), dim=1)
    tcost = (-logp[torch.arange(B), torch.tensor(Y).long()] / torch.tensor(pop)).mean()
    tcost.backward()
    assert abs(tcost.item() - cost) < 1e-10
    np.testing.assert_allclose(G["emb.W"], emb.grad.numpy(), atol=1e-9)
