"""Golden fixture (tests/golden/c1_golden.npz, BASELINE.json config 1).

CPU: the oracle reproduces the frozen costs / final parameters / top-10 exactly (regression pin).
GPU: the CUDA path, through the C ABI, reproduces them within the parity tolerance: per-step cost
1e-4 absolute, parameters 2e-4, recall@10 and sps identical, top-10 lists >= 95% identical entries
(fp32 vs float64 near-ties may swap neighbours)."""
import os

import numpy as np
import pytest

from oracle import sbr_oracle as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_golden.npz"))
SPEC = O.Spec(n_items=500, cell="GRU", layers=(100,), loss="CCE")
NP = int(G["n_params"])


def _goals():
    off = G["val_goal_off"]
    return [list(G["val_goal_flat"][off[i]:off[i + 1]]) for i in range(len(off) - 1)]


def _seen():
    off = G["val_seen_off"]
    return [list(G["val_seen_flat"][off[i]:off[i + 1]]) for i in range(len(off) - 1)]


def test_oracle_reproduces_golden():
    vals = [G["init_%02d" % i].astype(np.float64) for i in range(NP)]
    upd = O.Updater("adam", lr=1e-3)
    for s in range(8):
        c = O.train_step(SPEC, vals, upd, G["X_%d" % s], G["mask_%d" % s], Y=G["Y_%d" % s], pop=G["pop_%d" % s].astype(np.float64))
        assert abs(float(c) - float(G["costs"][s])) < 1e-12
    for i in range(NP):
        np.testing.assert_allclose(vals[i], G["final_%02d" % i], rtol=0, atol=1e-6)
    ex = np.zeros((len(G["val_X"]), 500))
    for i, s in enumerate(_seen()):
        ex[i, s] = 1
    top = O.top_k(O.test_scores(SPEC, vals, G["val_X"], G["val_mask"], exclude=ex), 10)
    np.testing.assert_array_equal(top, G["val_top10"])
    assert O.recall_at_k(_goals(), top, 10) == pytest.approx(float(G["val_recall10"]))


def test_float32_oracle_stays_within_parity_tolerance():
    """What a floatX=float32 Theano run would see: same fixture within 1e-4."""
    vals = [G["init_%02d" % i].astype(np.float32) for i in range(NP)]
    upd = O.Updater("adam", lr=1e-3)
    for s in range(8):
        c = O.train_step(SPEC, vals, upd, G["X_%d" % s], G["mask_%d" % s], Y=G["Y_%d" % s], pop=G["pop_%d" % s])
        assert abs(float(c) - float(G["costs"][s])) < 1e-4


@pytest.mark.gpu
def test_cuda_path_reproduces_golden():
    from sbr_b200 import _capi
    eng = _capi.Engine(n_items=500, cell="GRU", layers=(100,), loss="CCE", max_length=20, batch_size=20)
    try:
        eng.set_all_param_values([G["init_%02d" % i] for i in range(NP)])
        for s in range(8):
            c = eng.train_step_cce(G["X_%d" % s], G["mask_%d" % s], G["Y_%d" % s], G["pop_%d" % s])
            assert abs(float(c) - float(G["costs"][s])) < 1e-4, (s, c, G["costs"][s])
        for i, v in enumerate(eng.get_all_param_values()):
            assert np.abs(v - G["final_%02d" % i]).max() < 2e-4
        top = eng.topk(G["val_X"], G["val_mask"], k=10, exclude=_seen())
        assert (top == G["val_top10"]).mean() >= 0.95
        goals = _goals()
        assert O.recall_at_k(goals, top, 10) == pytest.approx(float(G["val_recall10"]), abs=1e-4)
        assert float(np.mean([g[0] in t for g, t in zip(goals, top)])) == pytest.approx(float(G["val_sps"]), abs=1e-4)
    finally:
        eng.close()
