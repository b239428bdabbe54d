"""CPU oracle for the RNN sequence-recommender training step.

TEST INFRASTRUCTURE ONLY.  This module is a numpy restatement of the reference's
Theano/Lasagne graph for the hot path; only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it.  The product
path (``sbr_b200`` + ``libsbr_b200.so``) never does.

PARITY UNPINNED: the reference (rdevooght/sequence-based-recommendations @0dfefed) ships
no tests, golden vectors or seeds, and Theano/Lasagne cannot be installed here (Python 2
only, no network).  The pins that exist are manufactured: (1) float64 finite-difference
gradient checks of every cell and loss in ``tests/test_oracle.py``; (2) an independent
``torch.autograd`` re-derivation of the same graph in ``tests/test_oracle_torch.py`` (cells with the grad_clip sites,
sampling and margin losses, updaters) and, for everything torch ships its own implementation of, torch's own modules:
bidirectional and stacked GRU / LSTM over packed sequences (states and gradients), the tanh-then-rectifier Vanilla
stack, the embedding + rating-id input path;
(3) frozen fixtures under ``tests/golden/`` made by ``tests/golden/make_golden.py``.

Every function cites the reference file:line it follows (paths relative to
/root/reference).  Third-party semantics (Lasagne master / Theano>=0.8.2, neither
vendored) are restated from their published behaviour, see SURVEY.md Appendix A.

Parameter layout: the oracle keeps the reference's *checkpoint* layout -- a python list of
arrays in ``lasagne.layers.get_all_param_values`` order (rnn_base.py:470-479), per-gate
matrices -- so that it is the reference for ``sbr_get_param`` / ``sbr_set_param`` too.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

GATES = {"LSTM": 4, "GRU": 3, "Vanilla": 1}
SAMPLING_LOSSES = ("BPR", "BPRI", "TOP1", "Blackout")
MARGIN_LOSSES = ("hinge", "logit", "logsig")


@dataclasses.dataclass
class Spec:
    """Static description of one model (what rnn_*._prepare_networks builds)."""
    n_items: int
    cell: str = "GRU"                 # recurrent_layers.py:9   (--r_t)
    layers: Tuple[int, ...] = (50,)   # recurrent_layers.py:10  (--r_l)
    embedding: int = 0                # recurrent_layers.py:12  (--r_emb)
    n_extra_ids: int = 0              # optional-feature id rows after n_items (rnn_base.py:636-642)
    ids_per_step: int = 1             # K = rnn_base._input_size() when --mf is off (rnn_base.py:615-622)
    grad_clip: float = 100.0          # recurrent_layers.py:19 (always 100)
    loss: str = "CCE"                 # command_parser.py:43
    regularization: float = 0.0       # rnn_one_hot.py:73-77 (output bias only)
    last_layer_tanh: bool = False     # rnn_sampling.py:19
    last_layer_init: float = 1.0      # rnn_sampling.py:19,131
    bidirectional: bool = False       # recurrent_layers.py:11,72-78 (--r_bi): every level = concat(forward, backward layer)

    @property
    def n_in(self) -> int:
        return self.n_items + self.n_extra_ids

    @property
    def G(self) -> int:
        return GATES[self.cell]


# --------------------------------------------------------------------------------------
# Parameter list (checkpoint order)
# --------------------------------------------------------------------------------------

def layer_prefixes(spec: Spec, li: int) -> List[str]:
    """Parameter-name prefixes of the directional layers of depth li: ['l0.'] or ['l0.', 'l0b.']."""
    return ["l%d." % li, "l%db." % li] if spec.bidirectional else ["l%d." % li]


def param_names_shapes(spec: Spec) -> List[Tuple[str, Tuple[int, ...]]]:
    """Names and shapes in lasagne.layers.get_all_params order.

    LSTM gate creation order in, forget, cell, out (sparse_lstm.py:241-253), then peepholes
    (:258-266), cell_init, hid_init (:270-279).  GRU creation order update, reset,
    hidden_update (sparse_lstm.py:660-671), hid_init (:674-676).  Vanilla: one triple + hid_init
    (sparse_lstm.py:1028-1039).  Embedding first (recurrent_layers.py:47-50); output layer last
    (rnn_one_hot.py:65, rnn_margin.py:103, rnn_sampling.py:131).
    """
    out: List[Tuple[str, Tuple[int, ...]]] = []
    if spec.embedding > 0:
        out.append(("emb.W", (spec.n_in, spec.embedding)))
        n_inputs = spec.embedding * spec.ids_per_step
    else:
        n_inputs = spec.n_in
    for li, H in enumerate(spec.layers):
        if spec.cell == "LSTM":
            gate_names = ["ingate", "forgetgate", "cell", "outgate"]
        elif spec.cell == "GRU":
            gate_names = ["updategate", "resetgate", "hidden_update"]
        else:
            gate_names = ["hidden_update"]
        # bidirectional: the forward layer's parameters, then the backward layer's, per depth (recurrent_layers.py:73-76)
        for pre in layer_prefixes(spec, li):
            if dense_vanilla(spec, li):
                # lasagne.layers.RecurrentLayer (recurrent_layers.py:98-99): CustomRecurrentLayer.get_params lists its
                # own hid_init first, then input_to_hidden (W, b), then hidden_to_hidden (W; no bias)
                out.append((pre + "hid_init", (1, H)))
                out.append((pre + "W_in_to_hid", (n_inputs, H)))
                out.append((pre + "b", (H,)))
                out.append((pre + "W_hid_to_hid", (H, H)))
                continue
            for g in gate_names:
                out.append((pre + "W_in_to_" + g, (n_inputs, H)))
                out.append((pre + "W_hid_to_" + g, (H, H)))
                out.append((pre + "b_" + g, (H,)))
            if spec.cell == "LSTM":
                out.append((pre + "W_cell_to_ingate", (H,)))
                out.append((pre + "W_cell_to_forgetgate", (H,)))
                out.append((pre + "W_cell_to_outgate", (H,)))
                out.append((pre + "cell_init", (1, H)))
            out.append((pre + "hid_init", (1, H)))
        n_inputs = H * (2 if spec.bidirectional else 1)
    out.append(("out.W", (spec.layers[-1] * (2 if spec.bidirectional else 1), spec.n_items)))
    out.append(("out.b", (spec.n_items,)))
    return out


def dense_vanilla(spec: Spec, li: int) -> bool:
    """A Vanilla layer whose input is dense -- deeper layers of a stack, or any layer behind an embedding -- is Lasagne's
    own RecurrentLayer, not the in-tree VanillaLayerOHEInput (recurrent_layers.py:86-87 vs :98-99).  Lasagne is a
    third-party dependency that is not in the reference tree (the reference's README pins Lasagne 0.2.dev1); its
    published RecurrentLayer: h_t = rectify(x_t W_in + b + h_{t-1} W_hid), one grad_clip on the summed pre-activation,
    W_in_to_hid / W_hid_to_hid ~ Uniform(-0.01, 0.01), b = 0, learned hid_init = 0.  The in-tree sparse copy uses tanh
    (sparse_lstm.py:960-961), so a Vanilla stack mixes tanh (layer 0 without embedding) and rectifiers (the rest)."""
    return spec.cell == "Vanilla" and (li > 0 or spec.embedding > 0)


def init_params(spec: Spec, rng: np.random.RandomState, dtype=np.float64) -> List[np.ndarray]:
    """Lasagne-equivalent initialisation, drawn in add_param order.

    Gate W_in/W_hid/W_cell ~ Normal(std 0.1), b = 0 (lasagne Gate defaults used at
    sparse_lstm.py:156-159,590-593,960-961); cell_init/hid_init = 0; EmbeddingLayer.W ~
    Normal(std 0.01); Dense/Blackout W ~ GlorotUniform(gain) (rnn_sampling.py:131), b = 0.
    """
    vals = []
    for name, shape in param_names_shapes(spec):
        leaf = name.split(".")[1]
        if name == "emb.W":
            v = rng.normal(0.0, 0.01, size=shape)
        elif name == "out.W":
            gain = spec.last_layer_init if spec.loss in SAMPLING_LOSSES else 1.0
            std = gain * np.sqrt(2.0 / (shape[0] + shape[1]))
            a = np.sqrt(3.0) * std
            v = rng.uniform(-a, a, size=shape)
        elif leaf in ("W_in_to_hid", "W_hid_to_hid"):
            v = rng.uniform(-0.01, 0.01, size=shape)      # lasagne.init.Uniform() of RecurrentLayer
        elif leaf.startswith("W_"):
            v = rng.normal(0.0, 0.1, size=shape)
        else:
            v = np.zeros(shape)
        vals.append(np.asarray(v, dtype=dtype))
    return vals


def _stack(spec: Spec, P: Dict[str, np.ndarray], li):
    """Stack per-gate matrices the way get_output_for does.

    LSTM stacks [in, forget, cell, out] (sparse_lstm.py:348-360); GRU stacks
    [reset, update, hidden] (sparse_lstm.py:737-749) although the params were *created*
    update-first; Vanilla has a single block (sparse_lstm.py:1098-1104).
    """
    pre = li if isinstance(li, str) else "l%d." % li
    if pre + "W_in_to_hid" in P:                      # Lasagne RecurrentLayer (dense_vanilla)
        return P[pre + "W_in_to_hid"], P[pre + "W_hid_to_hid"], P[pre + "b"], None
    if spec.cell == "LSTM":
        order = ["ingate", "forgetgate", "cell", "outgate"]
    elif spec.cell == "GRU":
        order = ["resetgate", "updategate", "hidden_update"]
    else:
        order = ["hidden_update"]
    W_in = np.concatenate([P[pre + "W_in_to_" + g] for g in order], axis=1)
    W_hid = np.concatenate([P[pre + "W_hid_to_" + g] for g in order], axis=1)
    b = np.concatenate([P[pre + "b_" + g] for g in order], axis=0)
    return W_in, W_hid, b, order


def as_dict(spec: Spec, values: Sequence[np.ndarray]) -> Dict[str, np.ndarray]:
    names = [n for n, _ in param_names_shapes(spec)]
    assert len(names) == len(values), (len(names), len(values))
    return dict(zip(names, values))


def _sigmoid(x):
    # T.nnet.sigmoid; written to stay finite for large |x|
    return np.where(x >= 0, 1.0 / (1.0 + np.exp(-np.abs(x))), np.exp(-np.abs(x)) / (1.0 + np.exp(-np.abs(x))))


# --------------------------------------------------------------------------------------
# Recurrent stack forward / backward
# --------------------------------------------------------------------------------------

def _layer_forward(spec: Spec, li: int, P, Xg, mask, pre=None, backwards=False):
    """One recurrent layer over time-major Xg [T,B,G*H], mask [T,B].

    LSTM step: sparse_lstm.py:377-415; GRU step: :764-796; Vanilla step: :1120-1143; masked
    switch: :417-425 / :798-805 / :1145-1152; learned init broadcast over batch: :438-445 /
    :817-819.
    Returns hs [T,B,H] (state after each step) and a cache for the backward pass.
    """
    pre = pre or "l%d." % li
    _, W_hid, _, _ = _stack(spec, P, pre)
    T, B, _ = Xg.shape
    H = spec.layers[li]
    dt = Xg.dtype
    h = np.repeat(P[pre + "hid_init"], B, axis=0).astype(dt)
    c = np.repeat(P[pre + "cell_init"], B, axis=0).astype(dt) if spec.cell == "LSTM" else None
    if spec.cell == "LSTM":
        w_ci, w_cf, w_co = (P[pre + "W_cell_to_ingate"], P[pre + "W_cell_to_forgetgate"], P[pre + "W_cell_to_outgate"])
    hs = np.zeros((T, B, H), dtype=dt)
    cache = [None] * T
    # a backwards layer scans the padded arrays from the last position to the first (theano.scan go_backwards, as
    # Lasagne does for backwards=True): the padding comes first and carries the initial state, then the row's items
    # in reverse; hs[t] stays aligned with the input position t (Lasagne reverses the scan output back)
    for t in (range(T - 1, -1, -1) if backwards else range(T)):
        m = mask[t][:, None] > 0
        if spec.cell == "LSTM":
            gates = Xg[t] + h @ W_hid
            i = _sigmoid(gates[:, 0:H] + c * w_ci)
            f = _sigmoid(gates[:, H:2 * H] + c * w_cf)
            g = np.tanh(gates[:, 2 * H:3 * H])
            c_new = f * c + i * g
            o = _sigmoid(gates[:, 3 * H:4 * H] + c_new * w_co)
            h_new = o * np.tanh(c_new)
            cache[t] = (h, c, i, f, g, o, c_new, m)
            c = np.where(m, c_new, c)
            h = np.where(m, h_new, h)
        elif spec.cell == "GRU":
            a = h @ W_hid
            x = Xg[t]
            r = _sigmoid(a[:, 0:H] + x[:, 0:H])
            u = _sigmoid(a[:, H:2 * H] + x[:, H:2 * H])
            cand = np.tanh(x[:, 2 * H:3 * H] + r * a[:, 2 * H:3 * H])
            h_new = (1 - u) * h + u * cand
            cache[t] = (h, a, r, u, cand, m)
            h = np.where(m, h_new, h)
        else:
            z = Xg[t] + h @ W_hid
            h_new = np.maximum(z, 0) if dense_vanilla(spec, li) else np.tanh(z)
            cache[t] = (h, h_new, m)
            h = np.where(m, h_new, h)
        hs[t] = h
    return hs, cache


def _layer_backward(spec: Spec, li: int, P, cache, dhs, G_out, pre=None, backwards=False):
    """BPTT through one layer (derived from the forward above).

    dhs [T,B,H]: external gradient arriving at the layer output of every step (zero except the
    last step for the top layer, because of only_return_final, sparse_lstm.py:485-486).
    Accumulates parameter gradients into G_out and returns dXg [T,B,G*H].
    grad_clip sites: LSTM sparse_lstm.py:386-388; GRU :768-772 and :789-791; Vanilla
    :1125-1129 and :1138-1140.
    """
    pre = pre or "l%d." % li
    _, W_hid, _, order = _stack(spec, P, pre)
    T = len(cache)
    H = spec.layers[li]
    G = spec.G
    B = dhs.shape[1]
    gc = spec.grad_clip
    dt = dhs.dtype
    dXg = np.zeros((T, B, G * H), dtype=dt)
    dW_hid = np.zeros_like(W_hid)
    dh = np.zeros((B, H), dtype=dt)
    dc = np.zeros((B, H), dtype=dt)

    def clip(x):
        return np.clip(x, -gc, gc) if gc else x

    if spec.cell == "LSTM":
        w_ci, w_cf, w_co = (P[pre + "W_cell_to_ingate"], P[pre + "W_cell_to_forgetgate"], P[pre + "W_cell_to_outgate"])
        dw_ci = np.zeros(H, dtype=dt); dw_cf = np.zeros(H, dtype=dt); dw_co = np.zeros(H, dtype=dt)
    for t in (range(T) if backwards else range(T - 1, -1, -1)):      # reverse of the processing order
        dh = dh + dhs[t]
        if spec.cell == "LSTM":
            h_prev, c_prev, i, f, g, o, c_new, m = cache[t]
            tc = np.tanh(c_new)
            do_pre = dh * tc * o * (1 - o)
            dct = dc + dh * o * (1 - tc * tc) + do_pre * w_co
            di_pre = dct * g * i * (1 - i)
            df_pre = dct * c_prev * f * (1 - f)
            dg_pre = dct * i * (1 - g * g)
            mz = m.astype(dt)
            do_pre, di_pre, df_pre, dg_pre = do_pre * mz, di_pre * mz, df_pre * mz, dg_pre * mz
            dw_co += (do_pre * c_new).sum(0)
            dw_ci += (di_pre * c_prev).sum(0)
            dw_cf += (df_pre * c_prev).sum(0)
            dgates = clip(np.concatenate([di_pre, df_pre, dg_pre, do_pre], axis=1))
            dXg[t] = dgates
            dW_hid += h_prev.T @ dgates
            dc = np.where(m, dct * f + di_pre * w_ci + df_pre * w_cf, dc)
            dh = np.where(m, dgates @ W_hid.T, dh)
        elif spec.cell == "GRU":
            h_prev, a, r, u, cand, m = cache[t]
            mz = m.astype(dt)
            du_pre = dh * (cand - h_prev) * u * (1 - u) * mz
            dq = clip(dh * u * (1 - cand * cand)) * mz
            dr_pre = dq * a[:, 2 * H:3 * H] * r * (1 - r)
            da_c = dq * r
            dx = clip(np.concatenate([dr_pre, du_pre, dq], axis=1))
            da = clip(np.concatenate([dr_pre, du_pre, da_c], axis=1))
            dXg[t] = dx
            dW_hid += h_prev.T @ da
            dh = np.where(m, dh * (1 - u) + da @ W_hid.T, dh)
        else:
            h_prev, h_new, m = cache[t]
            mz = m.astype(dt)
            dq = clip(dh * ((h_new > 0).astype(dt) if dense_vanilla(spec, li) else 1 - h_new * h_new)) * mz
            dx = clip(dq)
            da = clip(dq)
            dXg[t] = dx
            dW_hid += h_prev.T @ da
            dh = np.where(m, da @ W_hid.T, dh)
    # split stacked dW_hid back into per-gate gradients
    if order is None:
        G_out[pre + "W_hid_to_hid"] = dW_hid
    for gi, gname in enumerate(order or ()):
        G_out[pre + "W_hid_to_" + gname] = dW_hid[:, gi * H:(gi + 1) * H]
    if spec.cell == "LSTM":
        G_out[pre + "W_cell_to_ingate"] = dw_ci
        G_out[pre + "W_cell_to_forgetgate"] = dw_cf
        G_out[pre + "W_cell_to_outgate"] = dw_co
        G_out[pre + "cell_init"] = dc.sum(0, keepdims=True)
    G_out[pre + "hid_init"] = dh.sum(0, keepdims=True)
    return dXg


def forward_stack(spec: Spec, P, X, mask):
    """Input ids -> final hidden state.  X [B,T,K] int, mask [B,T].

    OHE gather-sum for layer 0 without embedding (sparse_lstm.py:368,755,1111); embedding +
    dense precompute otherwise (recurrent_layers.py:47-50, Lasagne precompute_input);
    time-major dimshuffle (sparse_lstm.py:343-344).  Bidirectional (recurrent_layers.py:72-78): every depth is a
    forward and a backwards layer over the same input, concatenated along the features; the final state is
    [forward state after the last position | backward state after the first position] (only_return_final takes the
    last SCAN step of each direction).
    """
    X = np.asarray(X)
    if X.ndim == 2:
        X = X[:, :, None]
    B, T, K = X.shape
    dt = P["out.b"].dtype
    maskT = np.asarray(mask, dtype=dt).T            # [T,B]
    Xt = np.transpose(X, (1, 0, 2))                  # [T,B,K]
    caches = []
    inp = None
    finals = []
    for li in range(len(spec.layers)):
        outs, level = [], []
        finals = []
        for di, pre in enumerate(layer_prefixes(spec, li)):
            W_in, _, b, _ = _stack(spec, P, pre)
            if li == 0 and spec.embedding == 0:
                Xg = W_in[Xt, :].sum(axis=-2) + b
            elif li == 0:
                inp = P["emb.W"][Xt, :].reshape(T, B, K * spec.embedding)
                Xg = inp @ W_in + b
            else:
                Xg = inp @ W_in + b
            hs, cache = _layer_forward(spec, li, P, Xg, maskT, pre=pre, backwards=(di == 1))
            outs.append(hs)
            finals.append(hs[0] if di == 1 else hs[-1])
            level.append(cache)
        caches.append((inp, level))
        inp = np.concatenate(outs, axis=-1) if len(outs) > 1 else outs[0]
    h_last = np.concatenate(finals, axis=-1) if len(finals) > 1 else finals[0]
    return h_last, (Xt, maskT, caches)


def backward_stack(spec: Spec, P, fwd_cache, dh_last, G_out):
    """Back-propagate dh_last [B,H_out] through the stack; fills G_out with every stack gradient."""
    Xt, maskT, caches = fwd_cache
    T, B, K = Xt.shape
    dt = dh_last.dtype
    L = len(spec.layers)
    nd = 2 if spec.bidirectional else 1
    Htop = spec.layers[-1]
    dhs_dir = []
    for di in range(nd):
        d = np.zeros((T, B, Htop), dtype=dt)
        d[0 if di == 1 else -1] = dh_last[:, di * Htop:(di + 1) * Htop]
        dhs_dir.append(d)
    for li in range(L - 1, -1, -1):
        inp, level = caches[li]
        H = spec.layers[li]
        dinp_total = None
        for di, pre in enumerate(layer_prefixes(spec, li)):
            W_in, _, _, order = _stack(spec, P, pre)
            dXg = _layer_backward(spec, li, P, level[di], dhs_dir[di], G_out, pre=pre, backwards=(di == 1))
            db = dXg.sum(axis=(0, 1))
            if li == 0 and spec.embedding == 0:
                dW_in = np.zeros_like(W_in)
                # AdvancedIncSubtensor1: duplicates accumulate (grad of sparse_lstm.py:368)
                np.add.at(dW_in, Xt.reshape(-1), np.repeat(dXg.reshape(T * B, -1), K, axis=0))
                dinp = None
            else:
                flat_in = inp.reshape(T * B, -1)
                dW_in = flat_in.T @ dXg.reshape(T * B, -1)
                dinp = (dXg.reshape(T * B, -1) @ W_in.T).reshape(inp.shape)
            if order is None:
                G_out[pre + "W_in_to_hid"] = dW_in
                G_out[pre + "b"] = db
            for gi, gname in enumerate(order or ()):
                G_out[pre + "W_in_to_" + gname] = dW_in[:, gi * H:(gi + 1) * H]
                G_out[pre + "b_" + gname] = db[gi * H:(gi + 1) * H]
            if dinp is not None:
                dinp_total = dinp if dinp_total is None else dinp_total + dinp
        if li == 0 and spec.embedding > 0:
            dE = np.zeros_like(P["emb.W"])
            np.add.at(dE, Xt.reshape(-1), dinp_total.reshape(T * B * K, spec.embedding))
            G_out["emb.W"] = dE
        if li > 0:
            Hb = spec.layers[li - 1]
            dhs_dir = [dinp_total[:, :, di * Hb:(di + 1) * Hb] for di in range(nd)]


# --------------------------------------------------------------------------------------
# Output layers and losses
# --------------------------------------------------------------------------------------

def cce_loss(spec: Spec, P, h, Y, pop):
    """softmax Dense + categorical_crossentropy / pop, mean over rows (rnn_one_hot.py:65-71),
    bias-only L2 (reg>0) / L1 (reg<0) (rnn_one_hot.py:73-77).  Returns cost, dh, dW, db."""
    W, b = P["out.W"], P["out.b"]
    B = h.shape[0]
    z = h @ W + b
    z = z - z.max(axis=1, keepdims=True)
    lse = np.log(np.exp(z).sum(axis=1, keepdims=True))
    logp = z - lse
    rows = np.arange(B)
    cost = (-logp[rows, Y] / pop).mean()
    dz = np.exp(logp)
    dz[rows, Y] -= 1.0
    dz = dz / (pop[:, None] * B)
    db = dz.sum(0)
    reg = spec.regularization
    if reg > 0:
        cost = cost + reg * (b * b).sum()
        db = db + 2.0 * reg * b
    elif reg < 0:
        cost = cost - reg * np.abs(b).sum()
        db = db - reg * np.sign(b)
    return cost, dz @ W.T, h.T @ dz, db


def sampling_loss(spec: Spec, P, h, Y, samples, pop):
    """BlackoutLayer column gather (sparse_lstm.py:41-54: cells = [targets; samples], no
    collision filtering) + BPR/BPRI/TOP1/Blackout (rnn_sampling.py:68-91), cost =
    mean(loss/pop) (rnn_sampling.py:137).  Returns cost, dh, dW(full [H,N]), db(full [N])."""
    W, b = P["out.W"], P["out.b"]
    B = h.shape[0]
    S = len(samples)
    cells = np.concatenate([np.asarray(Y), np.asarray(samples)]).astype(np.int64)
    A0 = h @ W[:, cells] + b[cells]
    rows = np.arange(B)
    A = np.tanh(A0) if (spec.last_layer_tanh and spec.loss != "Blackout") else A0
    dA = np.zeros_like(A)
    wrow = 1.0 / (pop * B)
    if spec.loss in ("BPR", "BPRI", "TOP1"):
        d = A[:, B:] - A[rows, rows][:, None]
        sd = _sigmoid(d)
        if spec.loss == "BPR":
            # -log(sigmoid(-d)) = softplus(d)
            loss = (np.maximum(d, 0) + np.log1p(np.exp(-np.abs(d)))).mean(axis=1)
            gd = sd / S
            gn = 0.0
        elif spec.loss == "BPRI":
            loss = (np.minimum(d, 0) - np.log1p(np.exp(-np.abs(d)))).mean(axis=1)
            gd = (1 - sd) / S
            gn = 0.0
        else:
            n = A[:, B:]
            sn = _sigmoid(n * n)
            loss = (sd + sn).mean(axis=1)
            gd = sd * (1 - sd) / S
            gn = sn * (1 - sn) * 2 * n / S
        dA[:, B:] = (gd + gn) * wrow[:, None]
        dA[rows, rows] -= gd.sum(axis=1) * wrow
        if spec.last_layer_tanh:
            dA = dA * (1 - A * A)
    else:  # Blackout: softmax over ALL B+S columns (rnn_sampling.py:68-72)
        z = A - A.max(axis=1, keepdims=True)
        Pm = np.exp(z) / np.exp(z).sum(axis=1, keepdims=True)
        loss = -np.log(Pm[rows, rows]) - np.log(1 - Pm[:, B:]).sum(axis=1)
        g = np.zeros_like(Pm)
        g[:, B:] = 1.0 / (1 - Pm[:, B:])
        g[rows, rows] += -1.0 / Pm[rows, rows]
        dA = Pm * (g - (g * Pm).sum(axis=1, keepdims=True)) * wrow[:, None]
    cost = (loss / pop).mean()
    dW = np.zeros_like(W)
    np.add.at(dW.T, cells, (h.T @ dA).T)
    db = np.zeros_like(b)
    np.add.at(db, cells, dA.sum(0))
    dh = dA @ W[:, cells].T
    return cost, dh, dW, db


def margin_loss(spec: Spec, P, h, Ymat, Wmat):
    """Linear Dense (rnn_margin.py:103) + hinge/logit/logsig (rnn_margin.py:61-68), summed over
    items, mean over rows (rnn_margin.py:109).  Ymat/Wmat dense [B,N] as built at :121-149."""
    W, b = P["out.W"], P["out.b"]
    B = h.shape[0]
    pred = h @ W + b
    if spec.loss == "hinge":
        z = (pred - Ymat) * Wmat
        loss = np.maximum(z, 0).sum(axis=1)
        # theano relu = 0.5*(x+|x|): slope 0.5 exactly at 0
        dpred = np.where(z > 0, 1.0, np.where(z == 0, 0.5, 0.0)) * Wmat
    elif spec.loss == "logit":
        s = _sigmoid(pred - Ymat)
        loss = (s * Wmat).sum(axis=1)
        dpred = s * (1 - s) * Wmat
    elif spec.loss == "logsig":
        z = (Ymat - pred) * Wmat
        loss = -(np.minimum(z, 0) - np.log1p(np.exp(-np.abs(z)))).sum(axis=1)
        dpred = (1 - _sigmoid(z)) * Wmat
    else:
        raise ValueError(spec.loss)
    cost = loss.mean()
    dpred = dpred / B
    return cost, dpred @ W.T, h.T @ dpred, dpred.sum(0)


def margin_targets(n_items, in_seqs, targets, balance=1.0, interactions_are_unique=True,
                   default_target=None, dtype=np.float64):
    """Dense Y / weight matrices exactly as RNNMargin._prepare_input fills them
    (rnn_margin.py:121-149): w = balance*n_t/(N-n_t-len); targets -> Y=1,w=-1; seen -> Y=0,w=0."""
    B = len(in_seqs)
    Y = np.zeros((B, n_items), dtype=dtype)
    Wm = np.zeros((B, n_items), dtype=dtype)
    for i in range(B):
        w = balance * len(targets[i]) / (n_items - len(targets[i]) - len(in_seqs[i]))
        Wm[i, :] = w
        Wm[i, list(targets[i])] = -1
        if interactions_are_unique:
            Wm[i, list(in_seqs[i])] = 0
        Y[i, :] = 0.0 if default_target is None else default_target
        Y[i, list(targets[i])] = 1
        if interactions_are_unique:
            Y[i, list(in_seqs[i])] = 0
    return Y, Wm


# --------------------------------------------------------------------------------------
# Updaters (lasagne.updates; call sites update_manager.py:32-82)
# --------------------------------------------------------------------------------------

class Updater:
    def __init__(self, kind="adam", lr=1e-3, rho=0.9, beta1=0.9, beta2=0.999):
        self.kind, self.lr, self.rho, self.beta1, self.beta2 = kind, lr, rho, beta1, beta2
        self.t = 0
        self.state: Optional[List[Dict[str, np.ndarray]]] = None

    def step(self, params: List[np.ndarray], grads: List[np.ndarray]):
        if self.state is None:
            self.state = [dict(a=np.zeros_like(p), b=np.zeros_like(p)) for p in params]
        dt = params[0].dtype.type
        lr = dt(self.lr)
        self.t += 1
        for p, g, s in zip(params, grads, self.state):
            g = g.reshape(p.shape)
            if self.kind == "adam":          # lasagne.updates.adam, epsilon 1e-8
                b1, b2 = dt(self.beta1), dt(self.beta2)
                a_t = lr * np.sqrt(dt(1) - b2 ** dt(self.t)) / (dt(1) - b1 ** dt(self.t))
                s["a"][...] = b1 * s["a"] + (dt(1) - b1) * g
                s["b"][...] = b2 * s["b"] + (dt(1) - b2) * g * g
                p -= a_t * s["a"] / (np.sqrt(s["b"]) + dt(1e-8))
            elif self.kind == "adagrad":     # epsilon 1e-6
                s["a"] += g * g
                p -= lr * g / np.sqrt(s["a"] + dt(1e-6))
            elif self.kind == "rmsprop":
                rho = dt(self.rho)
                s["a"][...] = rho * s["a"] + (dt(1) - rho) * g * g
                p -= lr * g / np.sqrt(s["a"] + dt(1e-6))
            elif self.kind == "adadelta":
                rho = dt(self.rho)
                s["a"][...] = rho * s["a"] + (dt(1) - rho) * g * g
                upd = g * np.sqrt(s["b"] + dt(1e-6)) / np.sqrt(s["a"] + dt(1e-6))
                p -= lr * upd
                s["b"][...] = rho * s["b"] + (dt(1) - rho) * upd * upd
            elif self.kind == "nesterov":    # lasagne.updates.nesterov_momentum
                mu = dt(self.rho)
                s["a"][...] = mu * s["a"] - lr * g
                p += mu * s["a"] - lr * g
            else:
                raise ValueError(self.kind)


# --------------------------------------------------------------------------------------
# Whole-step entry points (what theano.function compiles, rnn_base.py:175-213)
# --------------------------------------------------------------------------------------

def loss_and_grads(spec: Spec, values: List[np.ndarray], X, mask, *, Y=None, pop=None,
                   samples=None, Ymat=None, Wmat=None):
    """cost and the gradient list (same order as ``values``) for one mini-batch."""
    P = as_dict(spec, values)
    dt = values[0].dtype
    h, cache = forward_stack(spec, P, X, mask)
    if spec.loss == "CCE":
        cost, dh, dW, db = cce_loss(spec, P, h, np.asarray(Y), np.asarray(pop, dtype=dt))
    elif spec.loss in SAMPLING_LOSSES:
        cost, dh, dW, db = sampling_loss(spec, P, h, Y, samples, np.asarray(pop, dtype=dt))
    else:
        cost, dh, dW, db = margin_loss(spec, P, h, np.asarray(Ymat, dtype=dt), np.asarray(Wmat, dtype=dt))
    G: Dict[str, np.ndarray] = {"out.W": dW, "out.b": db}
    backward_stack(spec, P, cache, dh.astype(dt), G)
    grads = [np.asarray(G[n], dtype=dt).reshape(s) for n, s in param_names_shapes(spec)]
    return dt.type(cost), grads


def train_step(spec: Spec, values: List[np.ndarray], updater: Updater, X, mask, **kw):
    """train_function(*batch): cost at the *pre-update* parameters, then in-place update
    (rnn_base.py:185,290)."""
    cost, grads = loss_and_grads(spec, values, X, mask, **kw)
    updater.step(values, grads)
    return cost


def scores(spec: Spec, values: List[np.ndarray], X, mask):
    """predict_function / deterministic output (rnn_base.py:188-194): softmax probabilities for
    CCE (rnn_one_hot.py:65), raw linear scores for margin (rnn_margin.py:103) and for the
    deterministic BlackoutLayer (sparse_lstm.py:37-40)."""
    P = as_dict(spec, values)
    h, _ = forward_stack(spec, P, X, mask)
    z = h @ P["out.W"] + P["out.b"]
    if spec.loss == "CCE":
        z = z - z.max(axis=1, keepdims=True)
        e = np.exp(z)
        return e / e.sum(axis=1, keepdims=True)
    return z


def test_scores(spec: Spec, values, X, mask, exclude=None, interactions_are_unique=True):
    """test_function output before top-k (rnn_base.py:196-213; sampling override applies a
    softmax first, rnn_sampling.py:140-157); excluded items are multiplied by 0, not -inf."""
    out = scores(spec, values, X, mask)
    if spec.loss in SAMPLING_LOSSES:
        z = out - out.max(axis=1, keepdims=True)
        e = np.exp(z)
        out = e / e.sum(axis=1, keepdims=True)
    if interactions_are_unique and exclude is not None:
        out = out * (1 - np.asarray(exclude, dtype=out.dtype))
    return out


def top_k(row_scores: np.ndarray, k: int = 10) -> np.ndarray:
    """np.argpartition(-output, range(k))[:k] (rnn_base.py:159,207): the k best ids, best first."""
    return np.argpartition(-row_scores, list(range(k)), axis=-1)[..., :k]


def recall_at_k(goals: Sequence[Sequence[int]], predictions: Sequence[Sequence[int]], k: int = 10) -> float:
    """Evaluator.average_recall (helpers/evaluation.py:116-124)."""
    tot = 0.0
    for goal, pred in zip(goals, predictions):
        if len(goal) > 0:
            tot += float(len(set(goal) & set(list(pred)[:k]))) / len(goal)
    return tot / len(goals)
