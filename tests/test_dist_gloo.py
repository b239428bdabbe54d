"""CPU, world_size 2 over gloo: the host-side data-parallel logic.

Every rank builds the same global mini-batch from the same seeds and takes its slice of rows
(RNNBase._split_rows); the library then scales each row's loss by 1/B_global and sums gradients over
ranks with one all-reduce.  Here the all-reduce is gloo and the per-rank arithmetic is the oracle:
sum over ranks of (local gradient * B_local / B_global) must equal the single-process gradient of the
global batch, and the ragged margin targets must re-base correctly."""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sbr_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmpdir, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sbr_b200.helpers import synthetic
        from sbr_b200.helpers.data_handling import DataHandler
        from sbr_b200.neural_networks.recurrent_layers import RecurrentLayers
        from sbr_b200.neural_networks.rnn_margin import RNNMargin
        from sbr_b200.neural_networks.rnn_one_hot import RNNOneHot
        from sbr_b200.neural_networks.target_selection import SelectTargets
        from sbr_b200.neural_networks.update_manager import Adam
        ds = DataHandler(os.path.join(tmpdir, "ds") + "/")
        spec = O.Spec(n_items=ds.n_items, cell="GRU", layers=(12,), loss="CCE")
        vals = O.init_params(spec, np.random.RandomState(5))

        def mk(cls, **kw):
            p = cls(recurrent_layer=RecurrentLayers(layer_type="GRU", layers=[12]), updater=Adam(), max_length=10,
                    batch_size=8, use_ratings_features=False, use_movies_features=False, use_users_features=False,
                    n_ranks=world, rank=rank, nccl_id=b"x" * 128, **kw)
            p.n_items = ds.n_items
            p.set_dataset(ds)
            return p

        # ---- CCE: gradient identity under row sharding
        p = mk(RNNOneHot)
        random.seed(11); np.random.seed(11)
        X, mask, Y, pop, _ = next(p._gen_mini_batch(ds.training_set()))
        sl = p._split_rows
        assert sl(X).shape[0] == 4
        c_loc, g_loc = O.loss_and_grads(spec, vals, sl(X), sl(mask), Y=sl(Y), pop=sl(pop).astype(np.float64))
        scale = p.local_batch / p.batch_size
        flat = torch.tensor(np.concatenate([g.ravel() for g in g_loc] + [[c_loc]]) * scale)
        dist.all_reduce(flat)
        c_glob, g_glob = O.loss_and_grads(spec, vals, X, mask, Y=Y, pop=pop.astype(np.float64))
        ref = np.concatenate([g.ravel() for g in g_glob] + [[c_glob]])
        err = float(np.abs(flat.numpy() - ref).max())
        # ---- every rank saw the same global batch
        h = torch.tensor([float(X.sum()), float(Y.sum()), float(mask.sum())], dtype=torch.float64)
        hs = [torch.zeros_like(h) for _ in range(world)]
        dist.all_gather(hs, h)
        same = all(bool((a == hs[0]).all()) for a in hs)
        # ---- margin: ragged targets are re-based per rank
        pm = mk(RNNMargin, loss_function="hinge", target_selection=SelectTargets(n_targets=2))
        random.seed(12); np.random.seed(12)
        Xm, mm, (off, ids), w, seen = next(pm._gen_mini_batch(ds.training_set()))
        captured = {}

        class FakeEngine(object):
            def train_step_margin(self, X, mask, off, ids, w, default, exclude_seen=True):
                captured.update(X=X, off=np.array(off), ids=np.array(ids), w=w)
                return np.float32(0)
        pm.engine = FakeEngine()
        pm._compile_train_function()
        pm.train_function(Xm, mm, (off, ids), w, seen)
        lo = rank * 4
        ok = captured["off"][0] == 0 and len(captured["off"]) == 5
        for i in range(4):
            a = captured["ids"][captured["off"][i]:captured["off"][i + 1]]
            b = ids[off[lo + i]:off[lo + i + 1]]
            ok = ok and np.array_equal(a, b)
        ok = ok and np.array_equal(captured["X"], Xm[lo:lo + 4]) and np.array_equal(captured["w"], w[lo:lo + 4])
        q.put((rank, err, same, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gradient_identity_and_sharding(tmp_path):
    from sbr_b200.helpers import synthetic
    synthetic.write_dataset(str(tmp_path / "ds"), 60, 80, seed=3, uniform_len=(6, 25))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=400) for _ in procs]     # a cold `import torch` in the spawned ranks can take minutes
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, same, ok in res:
        assert err < 1e-12, (rank, err)
        assert same and ok
