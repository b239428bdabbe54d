"""CPU: host-side mirror of the reference API -- batch construction invariants, input preparation,
data format, metrics, checkpoint names, command-line flags."""
import os
import random
import re

import numpy as np
import pytest

from oracle import sbr_oracle as O
from sbr_b200.helpers import command_parser as cp
from sbr_b200.helpers import evaluation, synthetic
from sbr_b200.helpers.data_handling import DataHandler
from sbr_b200.helpers.early_stopping import StopAfterN, WaitWorstCaseTimesX
from sbr_b200.neural_networks.recurrent_layers import RecurrentLayers
from sbr_b200.neural_networks.rnn_margin import RNNMargin
from sbr_b200.neural_networks.rnn_one_hot import RNNOneHot
from sbr_b200.neural_networks.rnn_sampling import RNNSampling
from sbr_b200.neural_networks.sequence_noise import SequenceNoise
from sbr_b200.neural_networks.target_selection import SelectTargets
from sbr_b200.neural_networks.update_manager import Adam


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    d = tmp_path_factory.mktemp("ds")
    return DataHandler(synthetic.write_dataset(str(d / "c1"), 120, 300, seed=7, uniform_len=(5, 40)))


def _pred(cls, dataset, T=20, B=16, **kw):
    p = cls(recurrent_layer=RecurrentLayers(layer_type="GRU", layers=[32]), updater=Adam(), max_length=T,
            batch_size=B, use_movies_features=False, use_users_features=False,
            use_ratings_features=kw.pop("rf", False), **kw)
    p.n_items = dataset.n_items
    if hasattr(p, "sampling"):
        p.effective_sampling = int(p.sampling)
    p.set_dataset(dataset)
    return p


def test_dataset_format_and_stats(dataset):
    assert dataset.n_items == 300 and dataset.n_users == 120
    assert dataset.training_set.n_users + dataset.validation_set.n_users + dataset.test_set.n_users == 120
    seqs = list(dataset.test_set(epochs=1))
    assert len(seqs) == dataset.test_set.n_users
    seq, user = seqs[0]
    assert seq.shape[1] == 2 and len(np.unique(seq[:, 0])) == len(seq)   # unique interactions
    assert set(np.unique(seq[:, 1])) <= {1., 2., 3., 4., 5.}
    pop = dataset.item_popularity
    assert pop.shape == (300,) and pop.sum() == dataset.training_set.n_interactions


def test_mini_batches_are_nested_prefixes(dataset):
    random.seed(3); np.random.seed(3)
    p = _pred(RNNOneHot, dataset)
    gen = p._gen_mini_batch(dataset.training_set())
    for _ in range(5):
        X, mask, Y, pop, seen = next(gen)
        assert X.shape == (16, 20, 1) and X.dtype == np.int32 and mask.shape == (16, 20) and mask.dtype == np.float32
        lens = mask.sum(1).astype(int)
        assert (lens >= 1).all()
        for b in range(16):
            assert mask[b, :lens[b]].all() and not mask[b, lens[b]:].any()      # left aligned
            assert (X[b, lens[b]:] == 0).all()
            assert Y[b] not in X[b, :lens[b], 0]                                  # target comes after the input
            np.testing.assert_array_equal(seen[b], X[b, :lens[b], 0])
        assert (pop == 1).all()                                                   # diversity_bias 0


def test_truncation_keeps_the_last_max_length_items(dataset):
    random.seed(4); np.random.seed(4)
    p = _pred(RNNOneHot, dataset, T=6, B=8)
    X, mask, Y, pop, seen = next(p._gen_mini_batch(dataset.training_set()))
    assert mask.sum(1).max() <= 6


def test_test_mode_splits_in_the_middle(dataset):
    p = _pred(RNNOneHot, dataset)
    users = list(dataset.validation_set(epochs=1))
    out = list(p._gen_mini_batch(dataset.validation_set(epochs=1), test=True))
    assert len(out) == len(users)
    (X, mask, Y, pop, seen), goal = out[0]
    seq = users[0][0]
    half = int(len(seq) / 2)
    assert goal == [int(i) for i in seq[half:, 0]] and Y[0] == goal[0]
    n = int(mask.sum())
    np.testing.assert_array_equal(X[0, :n, 0], seq[max(0, half - 20):half, 0].astype(np.int32))


def test_rating_feature_ids(dataset):
    p = _pred(RNNOneHot, dataset, rf=True)
    f = p._features_of(np.array([[7, 0.5], [9, 5.0], [1, 3.0]]))
    assert f.shape == (3, 2)
    np.testing.assert_array_equal(f[:, 0], [7, 9, 1])
    np.testing.assert_array_equal(f[:, 1] - 300, [0, 9, 5])     # round(rating*2) - 1
    assert p._input_size() == 2 and p._n_optional_features() == 10


def test_sampling_and_margin_inputs(dataset):
    random.seed(5); np.random.seed(5)
    ps = _pred(RNNSampling, dataset, loss_function="BPR", sampling=32)
    X, mask, Y, samples, pop, seen = next(ps._gen_mini_batch(dataset.training_set()))
    assert samples.shape == (32,) and samples.dtype == np.int32 and samples.max() < 300
    pm = _pred(RNNMargin, dataset, loss_function="hinge", target_selection=SelectTargets(n_targets=3))
    batch = next(pm._gen_mini_batch(dataset.training_set()))
    X, mask, (off, ids), w, seen = batch
    assert off[0] == 0 and off[-1] == len(ids) and (np.diff(off) >= 1).all() and (np.diff(off) <= 3).all()
    Yd, Wd = pm.dense_targets(batch)
    in_seqs = [list(s) for s in seen]
    targets = [list(ids[off[i]:off[i + 1]]) for i in range(16)]
    Yr, Wr = O.margin_targets(300, in_seqs, targets, balance=1.0, dtype=np.float32)
    np.testing.assert_allclose(Yd, Yr)
    np.testing.assert_allclose(Wd, Wr, rtol=1e-6)


def test_target_selection_and_noise():
    seq = np.array([[i, 3.0] for i in range(10)], dtype=np.float64)
    assert SelectTargets(n_targets=2)(seq[4:]).tolist() == [[4, 3.0], [5, 3.0]]
    random.seed(0)
    sh = SelectTargets(n_targets=3, shuffle=True)(seq[4:])
    assert len(sh) == 3 and set(sh[:, 0]) <= set(range(4, 10))
    assert SelectTargets(n_targets=3, shuffle=True)(seq[4:], test=True)[:, 0].tolist() == [4, 5, 6]
    noise = SequenceNoise(dropout=0.5)
    np.random.seed(0)
    out, _ = next(noise(iter([(seq, "u")])))
    assert 2 <= len(out) < 10 and noise.name == "do0.5"
    assert SequenceNoise().name == ""


def test_model_filenames_follow_the_reference_scheme(dataset):
    p = _pred(RNNOneHot, dataset)
    assert p._get_model_filename(1.5) == "rnn_cce_db0.0_r0.0_ml20_bs16_ne1.5_GRU_gc100_h32_Ua_lr0.001_b10.9_b20.999_nt1_nf"
    assert re.search(r'_ne([0-9]+(\.[0-9]+)?)_', p._get_model_filename(2.25)).group(1) == "2.25"
    ps = _pred(RNNSampling, dataset, loss_function="TOP1", sampling=32)
    assert ps._get_model_filename(3).startswith("rnn_sampling_TOP1_s32_ini1.0_db0.0_ml20_bs16_ne3_")
    pm = _pred(RNNMargin, dataset, loss_function="logit")
    assert pm._get_model_filename(3).startswith("rnn_multitarget_logit_b1.0_ml20")
    assert RecurrentLayers(layer_type="LSTM", layers=[100, 50], embedding_size=8).name == "gc100_e8h100-50"


def test_command_line_flags_cover_the_reference_parser():
    args = cp.command_parser(cp.predictor_command_parser, argv=[])
    ref_defaults = dict(method='RNN', batch_size=16, learning_rate=0.01, regularization=0., gradient_clipping=100,
                        loss='CCE', sampling=32.0, diversity_bias=0.0, max_length=30, update_manager='adam', u_l=0.001,
                        u_rho=0.9, u_b1=0.9, u_b2=0.999, recurrent_layer_type='GRU', r_l='50', r_emb=0, n_targets=1,
                        target_bias=-1., n_dropout=0., clusters=-1, balance=1., min_access=0.05)
    for k, v in ref_defaults.items():
        assert getattr(args, k) == v, k
    a = cp.command_parser(cp.predictor_command_parser, argv="--loss hinge --r_t LSTM --r_l 64-32 --n_targets 2 -b 8".split())
    pred = cp.get_predictor(a)
    assert isinstance(pred, RNNMargin) and pred.recurrent_layer.layers == [64, 32] and pred.batch_size == 8
    assert isinstance(cp.get_predictor(cp.command_parser(cp.predictor_command_parser, argv=["--loss", "BPR"])), RNNSampling)
    with pytest.raises(NotImplementedError):
        cp.get_predictor(cp.command_parser(cp.predictor_command_parser, argv=["-m", "POP"]))
    with pytest.raises(ValueError):
        cp.get_predictor(cp.command_parser(cp.predictor_command_parser, argv=["--loss", "nope"]))


def test_evaluator_metrics(dataset):
    ev = evaluation.Evaluator(dataset, k=3)
    ev.add_instance([1, 2, 3], [3, 9, 1, 2])
    ev.add_instance([5], [6, 7, 8])
    assert ev.average_recall() == pytest.approx((2 / 3 + 0) / 2)
    assert ev.average_recall() == pytest.approx(O.recall_at_k([[1, 2, 3], [5]], [[3, 9, 1, 2], [6, 7, 8]], 3))
    assert ev.sps() == 0.5 and ev.user_coverage() == 0.5 and ev.item_coverage() == 2
    assert ev.average_precision() == pytest.approx((2 / 3) / 2)
    dcg = 1 / np.log2(2) + 1 / np.log2(4)
    ideal = 1 / np.log2(2) + 1 / np.log2(3) + 1 / np.log2(4)
    assert ev.average_ndcg() == pytest.approx(dcg / ideal / 2)


def test_early_stopping_rules():
    s = StopAfterN(n=2)
    assert not s([1, 2, 3], [0.1, 0.2, 0.3])
    assert s([1, 2, 3, 4], [0.3, 0.3, 0.2, 0.1])
    w = WaitWorstCaseTimesX(x=2., min_wait=1.)
    assert not w([1, 2, 3], [0.1, 0.2, 0.3])
    assert w([1, 2, 3, 9], [0.1, 0.2, 0.3, 0.1])
    assert StopAfterN(n=1, higher_is_better=False)([1, 2, 3], [0.1, 0.2, 0.3])


def test_pareto_front(dataset):
    p = _pred(RNNOneHot, dataset)
    m = {"sps": [0.1, 0.3, 0.2], "recall": [0.3, 0.1, 0.2]}
    assert p.get_pareto_front(m, ["sps", "recall"]) == [0, 1, 2]
    assert p.get_pareto_front({"sps": [0.1, 0.3, 0.2]}, ["sps"]) == [1]


# ---------------------------------------------------------------------------------------------
# threaded_generator: the reference's (disabled) prefetch helper, rnn_base.py:34-56 / :273-274
# ---------------------------------------------------------------------------------------------
def test_threaded_generator_preserves_order_and_content():
    from sbr_b200.neural_networks.rnn_base import threaded_generator
    src = [(i, np.full(3, i)) for i in range(500)]
    out = list(threaded_generator(iter(src), num_cached=7))
    assert [i for i, _ in out] == list(range(500))
    assert all((a == i).all() for i, a in out)
    assert list(threaded_generator(iter([]), num_cached=3)) == []


def test_threaded_generator_reraises_producer_errors_in_place():
    from sbr_b200.neural_networks.rnn_base import threaded_generator

    def gen():
        yield 1
        yield 2
        raise KeyError("boom")

    g = threaded_generator(gen(), num_cached=2)
    assert next(g) == 1 and next(g) == 2
    with pytest.raises(KeyError):
        next(g)
    with pytest.raises(StopIteration):
        next(g)


def test_threaded_generator_close_stops_an_endless_producer():
    import itertools
    import time as _time
    from sbr_b200.neural_networks.rnn_base import threaded_generator
    g = threaded_generator(itertools.count(), num_cached=4)
    assert [next(g) for _ in range(10)] == list(range(10))
    g.close()
    _time.sleep(0.2)
    assert not g._thread.is_alive()
    with pytest.raises(StopIteration):
        next(g)


def test_prefetch_flag_reaches_the_predictor():
    from sbr_b200.helpers import command_parser as cp
    args = cp.command_parser(cp.predictor_command_parser, argv=["--prefetch", "16"])
    assert args.prefetch == 16
    p = cp.get_predictor(args)
    assert p.prefetch_batches == 16
    args = cp.command_parser(cp.predictor_command_parser, argv=[])
    assert cp.get_predictor(args).prefetch_batches == 0


@pytest.mark.parametrize("rf", [False, True])
def test_pre_encoded_rows_give_the_same_batches(rf, tmp_path):
    """_gen_mini_batch encodes a user's sequence once and hands the rows views of it; the padded tensors must be
    identical to encoding every row on its own (the reference's per-row loop, rnn_one_hot.py:90-101)."""
    from sbr_b200.neural_networks.rnn_one_hot import RNNOneHot
    from sbr_b200.neural_networks.rnn_margin import RNNMargin
    from sbr_b200.neural_networks.target_selection import SelectTargets
    d = str(tmp_path / "ds") + "/"
    synthetic.write_dataset(d, n_users=60, n_items=80, min_len=8, max_len=40, seed=3)
    ds = DataHandler(d)
    for cls, kw in ((RNNOneHot, {}), (RNNMargin, dict(target_selection=SelectTargets(n_targets=2)))):
        p = cls(max_length=12, batch_size=16, use_ratings_features=rf, use_movies_features=False,
                use_users_features=False, **kw)
        p.n_items = ds.n_items
        p.set_dataset(ds)
        captured = []
        orig = p._prepare_input
        p._prepare_input = lambda seqs: (captured.append([r[:3] for r in seqs]), orig(seqs))[1]
        random.seed(4); np.random.seed(4)
        gen = p._gen_mini_batch(ds.training_set())
        for _ in range(5):
            out = next(gen)
            ref = orig(captured[-1])            # same rows without the pre-encoded 4th element
            assert any(len(r) > 3 for r in captured[-1]) is False
            np.testing.assert_array_equal(out[0], ref[0])
            np.testing.assert_array_equal(out[1], ref[1])
            assert all((a == b).all() for a, b in zip(out[-1], ref[-1]))
            assert out[0].dtype == np.int32 and out[1].dtype == np.float32
            lens = out[1].sum(1).astype(int)
            assert ((out[1] == 1) == (np.arange(12)[None, :] < lens[:, None])).all()


@pytest.mark.parametrize("rf", [False, True])
def test_compact_batches_describe_the_same_rows(rf, dataset):
    """Device-side assembly: with the training sequences uploaded, a training batch is B (sequence,
    start, length) triples; expanding them against the uploaded CSR gives exactly the X / mask / Y / pop of the
    reference-style batch built from the same RNG state."""
    class FakeEngine(object):
        def dataset_upload(self, off, ids):
            self.off, self.ids = np.asarray(off), np.asarray(ids)

    p_dense = _pred(RNNOneHot, dataset, rf=rf)
    p_rows = _pred(RNNOneHot, dataset, rf=rf)
    assert p_rows._plain_first_target() or True
    p_rows.engine = FakeEngine()
    assert p_rows._upload_training_sequences(dataset)
    eng = p_rows.engine
    assert eng.ids.shape[1] == (2 if rf else 1) and eng.off[-1] == len(eng.ids)
    random.seed(11); np.random.seed(11)
    g_dense = p_dense._gen_mini_batch(dataset.training_set())
    dense = [next(g_dense) for _ in range(4)]
    random.seed(11); np.random.seed(11)
    g_rows = p_rows._gen_mini_batch(dataset.training_set())
    rows = [next(g_rows) for _ in range(4)]
    from sbr_b200.neural_networks.rnn_base import CompactBatch
    for (X, mask, Y, pop, seen), batch in zip(dense, rows):
        assert len(batch) == 1 and isinstance(batch[0], CompactBatch)
        cb = batch[0]
        assert len(cb) == X.shape[0]
        X2 = np.zeros_like(X)
        for b in range(len(cb)):
            lo = eng.off[cb.seq[b]] + cb.start[b]
            X2[b, :cb.length[b]] = eng.ids[lo:lo + cb.length[b]]
        np.testing.assert_array_equal(X2, X)
        np.testing.assert_array_equal(cb.length, mask.sum(1).astype(np.int32))
        np.testing.assert_array_equal(cb.Y, Y)
        np.testing.assert_array_equal(cb.pop, pop)
    # validation batches keep the reference form (they are not rows of the training set)
    (Xv, maskv, *_), goal = next(p_rows._gen_mini_batch(dataset.validation_set(epochs=1), test=True))
    assert Xv.shape[0] == 1 and maskv.sum() >= 1 and len(goal) >= 1


def test_vectorised_and_row_loop_compact_generators_agree(dataset):
    """The per-user numpy pass (`_gen_compact_batches`, used when a row's target is just the next item) and the row loop
    of `_gen_mini_batch` consume the RNGs identically and produce the same triples / targets."""
    class FakeEngine(object):
        def dataset_upload(self, off, ids):
            pass

    from sbr_b200.neural_networks.rnn_base import CompactBatch
    outs = []
    for force_loop in (False, True):
        p = _pred(RNNOneHot, dataset, B=32)
        p.engine = FakeEngine()
        assert p._upload_training_sequences(dataset)
        if force_loop:
            p._plain_first_target = lambda: False
        else:
            assert p._plain_first_target()
        random.seed(21); np.random.seed(21)
        g = p._gen_mini_batch(dataset.training_set())
        outs.append([next(g)[0] for _ in range(6)])
    for a, b in zip(*outs):
        assert isinstance(a, CompactBatch) and isinstance(b, CompactBatch)
        for f in ("seq", "start", "length", "Y", "pop"):
            np.testing.assert_array_equal(getattr(a, f), getattr(b, f))
