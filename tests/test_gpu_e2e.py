"""GPU: the reference-facing Python API end to end -- RNNOneHot.train_function / test_function /
top_k_recommendations / save / load, the train.py and test.py command lines on a synthetic dataset in
the reference's on-disk format, the size-independent properties at BASELINE's full C2 size, and the
2-rank NCCL data-parallel path (skipped with fewer than 2 GPUs)."""
import glob
import os
import random
import subprocess
import sys

import numpy as np
import pytest

from oracle import sbr_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    from sbr_b200.helpers import synthetic
    from sbr_b200.helpers.data_handling import DataHandler
    d = tmp_path_factory.mktemp("ds")
    return DataHandler(synthetic.write_dataset(str(d / "c1"), 200, 500, seed=1234, uniform_len=(5, 40)))


def _predictor(dataset, cls=None, **kw):
    from sbr_b200.neural_networks.recurrent_layers import RecurrentLayers
    from sbr_b200.neural_networks.rnn_one_hot import RNNOneHot
    from sbr_b200.neural_networks.update_manager import Adam
    cls = cls or RNNOneHot
    p = cls(recurrent_layer=RecurrentLayers(layer_type=kw.pop("cell", "GRU"), layers=kw.pop("layers", [100]),
                                            bidirectional=kw.pop("bidirectional", False), embedding_size=kw.pop("emb", 0)),
            updater=Adam(), max_length=20, batch_size=16, use_ratings_features=kw.pop("rf", False),
            use_movies_features=False, use_users_features=False, **kw)
    p.prepare_model(dataset)
    p.set_dataset(dataset)
    return p


def test_public_api_matches_oracle_per_step_loss_and_recall(dataset):
    """C1: K training steps from identical init on identical batches -> per-step loss within 1e-4 and
    the same recall@10 / sps on the validation set (north_star parity statement)."""
    p = _predictor(dataset)
    p._compile_train_function()
    p._compile_test_function()
    spec = O.Spec(n_items=500, cell="GRU", layers=(100,), loss="CCE")
    vals = O.init_params(spec, np.random.RandomState(1), np.float64)
    p.engine.set_all_param_values(vals)
    upd = O.Updater("adam", lr=1e-3)
    random.seed(1234); np.random.seed(1234)
    gen = p._gen_mini_batch(dataset.training_set())
    for step in range(25):
        batch = next(gen)
        c = p.train_function(*batch)
        X, mask, Y, pop, _ = batch
        c_ref = O.train_step(spec, vals, upd, X, mask, Y=Y, pop=pop.astype(np.float64))
        assert abs(float(c) - float(c_ref)) < 1e-4, (step, c, c_ref)
    metrics = p._compute_validation_metrics({m: [] for m in p.metrics})
    goals, tops = [], []
    for (X, mask, Y, pop, seen), goal in p._gen_mini_batch(dataset.validation_set(epochs=1), test=True):
        ex = np.zeros((1, 500)); ex[0, seen[0]] = 1
        tops.append(O.top_k(O.test_scores(spec, vals, X, mask, exclude=ex), 10)[0])
        goals.append(goal)
    assert metrics["recall"][-1] == pytest.approx(O.recall_at_k(goals, tops, 10), abs=1e-4)
    assert metrics["sps"][-1] == pytest.approx(float(np.mean([g[0] in t for g, t in zip(goals, tops)])), abs=1e-4)
    # top_k_recommendations: seen items are -inf, not 0 (rnn_base.py:154-156)
    seq, user = next(dataset.test_set(epochs=1))
    rec = p.top_k_recommendations(seq[:len(seq) // 2], k=10)
    assert len(rec) == 10 and not set(rec) & set(seq[:len(seq) // 2, 0].astype(int))
    p.engine.close()


def test_a_fresh_predictor_is_initialised_and_learns(dataset):
    """A predictor straight from the constructor has Lasagne's initial weights (not the zero arena): one training step
    gives non-zero gradients for the recurrent and the output weights, and a few hundred steps beat the cost of the
    first one."""
    p = _predictor(dataset, cell="LSTM", layers=[50])
    p._compile_train_function()
    vals = dict(zip([n for n, _ in p.engine.param_infos()], p.engine.get_all_param_values()))
    assert vals["l0.W_hid_to_ingate"].std() == pytest.approx(0.1, rel=0.2) and np.abs(vals["out.W"]).max() > 0
    random.seed(5); np.random.seed(5)
    gen = p._gen_mini_batch(dataset.training_set())
    p.engine.set_skip_update(True)
    p.train_function(*next(gen))
    grads = dict(zip([n for n, _ in p.engine.param_infos()], p.engine.get_all_grads()))
    assert np.abs(grads["l0.W_hid_to_ingate"]).max() > 0 and np.abs(grads["out.W"]).max() > 0
    p.engine.set_skip_update(False)
    costs = [float(p.train_function(*next(gen))) for _ in range(300)]
    assert np.mean(costs[-20:]) < np.mean(costs[:20])
    p.engine.close()


@pytest.mark.parametrize("cell,layers,emb", [("GRU", [24, 16], 0), ("Vanilla", [32, 16], 0), ("LSTM", [24], 8)])
def test_bidirectional_and_stacked_predictors_match_the_oracle(dataset, cell, layers, emb):
    """--r_bi / --r_l a-b / --r_emb through the public API: per-step loss of the first Adam steps against the oracle."""
    p = _predictor(dataset, cell=cell, layers=list(layers), bidirectional=True, emb=emb)
    p._compile_train_function()
    spec = O.Spec(n_items=500, cell=cell, layers=tuple(layers), loss="CCE", bidirectional=True, embedding=emb)
    assert [n for n, _ in p.engine.param_infos()] == [n for n, _ in O.param_names_shapes(spec)]
    vals = [v.astype(np.float64) for v in p.engine.get_all_param_values()]      # the predictor's own initial weights
    upd = O.Updater("adam", lr=1e-3)
    random.seed(99); np.random.seed(99)
    gen = p._gen_mini_batch(dataset.training_set())
    for step in range(8):
        batch = next(gen)
        c = p.train_function(*batch)
        X, mask, Y, pop, _ = batch
        c_ref = O.train_step(spec, vals, upd, X, mask, Y=Y, pop=pop.astype(np.float64))
        assert abs(float(c) - float(c_ref)) < 1e-4, (step, c, c_ref)
    p.engine.close()


def test_save_load_roundtrip_and_python2_readable_pickle(dataset, tmp_path):
    import pickle
    p = _predictor(dataset, cell="LSTM", layers=[24, 16])
    f = str(tmp_path / "models" / p._get_model_filename(0.5))
    p.save(f)
    with open(f, "rb") as fh:
        blob = fh.read()
    assert blob[:2] == b"\x80\x02"                       # pickle protocol 2
    params = pickle.loads(blob)
    names = [n for n, _ in O.param_names_shapes(O.Spec(n_items=500, cell="LSTM", layers=(24, 16)))]
    assert len(params) == len(names) == len(p.engine.param_infos())
    assert [n for n, _ in p.engine.param_infos()] == names
    before = p.engine.get_all_param_values()
    p.engine.set_all_param_values([np.zeros_like(v) for v in before])
    p.load(f)
    for a, b in zip(before, p.engine.get_all_param_values()):
        np.testing.assert_array_equal(a, b)
    assert p.load_last(str(tmp_path / "models") + "/") == 0.5
    p.engine.close()


def test_train_and_test_command_lines(dataset):
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, os.path.join(ROOT, "train.py"), "-d", dataset.dirname, "-m", "RNN", "--r_t", "GRU", "--r_l", "32",
           "--max_length", "20", "-b", "16", "--max_iter", "40", "--progress", "20", "--save", "All", "--seed", "3",
           "--metrics", "sps,recall"]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stderr.splitlines() if l and l[0].isdigit()]
    assert len(lines) == 2 and lines[0].split()[0] == "20"          # machine-readable progress lines on stderr
    models = glob.glob(dataset.dirname + "models/rnn_cce_*")
    assert len(models) == 2
    cmd = [sys.executable, os.path.join(ROOT, "test.py"), "-d", dataset.dirname, "-m", "RNN", "--r_t", "GRU", "--r_l", "32",
           "--max_length", "20", "-b", "16"]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.count("sps@10:") == 2 and "recall@10:" in out.stdout


def test_full_size_c2_properties():
    """BASELINE config 2 at full size (LSTM 1x200, 3706 items, B=128, T=200): properties that do not
    need the oracle -- padding ids are inert, row permutation leaves the cost unchanged, a zero
    learning rate leaves the parameters bit-identical, and the cost starts at ~ln(N)."""
    from sbr_b200 import _capi
    rng = np.random.RandomState(2)
    B, T, N = 128, 200, 3706
    lens = np.sort(rng.randint(2, T + 1, B))
    X = np.zeros((B, T, 1), np.int32); mask = np.zeros((B, T), np.float32)
    base = rng.randint(0, N, T)
    for b in range(B):
        X[b, :lens[b], 0] = base[:lens[b]]            # nested prefixes, like the reference batches
        mask[b, :lens[b]] = 1
    Y = rng.randint(0, N, B).astype(np.int32); pop = rng.uniform(0.5, 2, B).astype(np.float32)
    eng = _capi.Engine(n_items=N, cell="LSTM", layers=(200,), loss="CCE", max_length=T, batch_size=B, lr=0.0)
    try:
        spec = O.Spec(n_items=N, cell="LSTM", layers=(200,))
        eng.set_all_param_values(O.init_params(spec, np.random.RandomState(1), np.float32))
        p0 = eng.get_all_param_values()
        c0 = eng.train_step_cce(X, mask, Y, pop)
        assert abs(float(c0) - np.mean(np.log(N) / pop)) < 0.5
        X2 = X.copy(); X2[mask == 0] = 1234
        assert eng.train_step_cce(X2, mask, Y, pop) == c0
        perm = rng.permutation(B)
        c1 = eng.train_step_cce(X[perm], mask[perm], Y[perm], pop[perm])
        assert abs(float(c1) - float(c0)) < 2e-5
        for a, b in zip(p0, eng.get_all_param_values()):
            np.testing.assert_array_equal(a, b)
        eng.set_skip_update(True)
        eng.train_step_cce(X, mask, Y, pop)
        g = eng.get_all_grads()
        names = [n for n, _ in eng.param_infos()]
        used = np.unique(base[:lens.max()])
        gW = g[names.index("l0.W_in_to_ingate")]
        unused = np.setdiff1d(np.arange(N), used)
        assert not gW[unused].any() and np.abs(gW[used]).sum() > 0       # scatter touches exactly the used rows
        assert abs(g[names.index("out.b")].sum()) < 1e-5                 # softmax gradient sums to zero
    finally:
        eng.close()


_RANK_SCRIPT = r"""
import os, sys, numpy as np, pickle
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from sbr_b200 import _capi
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
box = [_capi.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
d = pickle.load(open(sys.argv[2], "rb"))
loss = d["loss"]
Bl = d["B"] // world
eng = _capi.Engine(n_items=d["N"], cell="LSTM", layers=(48,), loss=loss, max_length=d["T"], batch_size=Bl, device=rank,
                   n_ranks=world, rank=rank, nccl_id=box[0], global_batch=d["B"], n_samples=8)
eng.set_all_param_values(d["vals"])
lo = rank * Bl
costs = []
for X, mask, Y, pop, samples in d["batches"]:
    if loss == "CCE":
        c = eng.train_step_cce(X[lo:lo+Bl], mask[lo:lo+Bl], Y[lo:lo+Bl], pop[lo:lo+Bl])
    else:
        c = eng.train_step_sampled(X[lo:lo+Bl], mask[lo:lo+Bl], Y[lo:lo+Bl], samples, pop[lo:lo+Bl], Y_all=Y, row_offset=lo)
    costs.append(float(c))
pickle.dump((costs, eng.get_all_param_values()), open(sys.argv[2] + ".out%d" % rank, "wb"))
eng.close()
dist.barrier()
dist.destroy_process_group()
"""


@pytest.mark.parametrize("loss", ["CCE", "Blackout"])
def test_two_rank_nccl_matches_single_rank(tmp_path, loss):
    from sbr_b200 import _capi
    import pickle
    if _capi.load_library().sbr_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.RandomState(4)
    N, T, B = 300, 12, 16
    spec = O.Spec(n_items=N, cell="LSTM", layers=(48,), loss=loss)
    vals = O.init_params(spec, rng, np.float32)
    batches = []
    for _ in range(4):
        lens = rng.randint(1, T + 1, B)
        X = np.zeros((B, T, 1), np.int32); mask = np.zeros((B, T), np.float32)
        for b in range(B):
            X[b, :lens[b], 0] = rng.randint(0, N, lens[b]); mask[b, :lens[b]] = 1
        batches.append((X, mask, rng.randint(0, N, B).astype(np.int32), rng.uniform(0.5, 2, B).astype(np.float32),
                        rng.randint(0, N, 8).astype(np.int32)))
    f = str(tmp_path / "job.pkl")
    pickle.dump(dict(N=N, T=T, B=B, vals=vals, batches=batches, loss=loss), open(f, "wb"))
    script = str(tmp_path / "rank.py")
    open(script, "w").write(_RANK_SCRIPT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", script, ROOT, f]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    r0 = pickle.load(open(f + ".out0", "rb")); r1 = pickle.load(open(f + ".out1", "rb"))
    # replicas stay identical; the all-reduced cost is the global cost
    for a, b in zip(r0[1], r1[1]):
        np.testing.assert_array_equal(a, b)
    assert r0[0] == r1[0]
    v64 = [v.astype(np.float64) for v in vals]
    upd = O.Updater("adam", lr=1e-3)
    for s, (X, mask, Y, pop, samples) in enumerate(batches):
        kw = dict(Y=Y, pop=pop.astype(np.float64))
        if loss != "CCE":
            kw["samples"] = samples
        c_ref = O.train_step(spec, v64, upd, X, mask, **kw)
        assert abs(r0[0][s] - float(c_ref)) < 1e-4
    for a, b in zip(v64, r0[1]):
        assert np.abs(a - b).max() < 2e-4


def test_device_side_batch_assembly_gives_identical_training(dataset):
    """rows sent as (sequence, start, length) triples and assembled on the device train exactly like the
    reference-style X / mask batches (same RNG state -> same rows -> identical costs; parameters equal up to the order of the atomic gradient sums)."""
    from sbr_b200.neural_networks.recurrent_layers import RecurrentLayers
    from sbr_b200.neural_networks.rnn_one_hot import RNNOneHot
    from sbr_b200.neural_networks.update_manager import Adam

    def run(device_batches):
        random.seed(5); np.random.seed(5)
        p = RNNOneHot(recurrent_layer=RecurrentLayers(layer_type="GRU", layers=[48]), updater=Adam(), max_length=20, batch_size=16,
                      use_ratings_features=True, use_movies_features=False, use_users_features=False, init_seed=3,
                      device_batches=device_batches)
        p.prepare_model(dataset)
        p.set_dataset(dataset)
        p._compile_train_function()
        if device_batches:
            assert p._upload_training_sequences(dataset)
        gen = p._gen_mini_batch(dataset.training_set())
        costs = []
        for _ in range(12):
            batch = next(gen)
            if device_batches:
                assert len(batch) == 1
            costs.append(float(p.train_function(*batch)))
        vals = p.engine.get_all_param_values()
        p.engine.close()
        return costs, vals

    c_rows, v_rows = run(True)
    c_dense, v_dense = run(False)
    assert c_rows[0] == c_dense[0]                      # same rows, same parameters: the first cost is bit-identical
    np.testing.assert_allclose(c_rows, c_dense, rtol=0, atol=1e-5)
    for a, b in zip(v_rows, v_dense):       # gradient sums are accumulated with floating-point atomics: equal up to their order
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-5)
