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
