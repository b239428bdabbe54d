"""RNNMargin -- host mirror of neural_networks/rnn_margin.py:13-161: linear full-catalog output
with the hinge / logit / logsig multi-target losses.  The reference fills three dense [B, n_items]
host matrices per batch (rnn_margin.py:121-149); here the batch carries the ragged targets and
one scalar weight per row, and `sbr_train_step_margin` rebuilds the matrices on the device.
The dense form (`sbr_train_step_margin_dense`) is kept for drop-in callers."""
import numpy as np

from . import rnn_base as rnn


class RNNMargin(rnn.RNNBase):
    def __init__(self, loss_function="hinge", balance=1., popularity_based=False, min_access=0.05, n_targets=1, **kwargs):
        super().__init__(**kwargs)
        self.balance = balance
        self.popularity_based = popularity_based
        self.min_access = min_access
        self.n_targets = n_targets
        if loss_function is None:
            loss_function = "hinge"
        if loss_function not in ("hinge", "logit", "logsig"):
            raise ValueError('Unknown loss function')
        self.loss_function_name = loss_function
        self.loss_name = loss_function
        self.name = "RNN multi-targets"

    def _get_model_filename(self, epochs):
        filename = "rnn_multitarget_" + self.loss_function_name + "_b" + str(self.balance)
        if self.popularity_based:
            filename += '_pb_ma' + str(self.min_access)
        return filename + "_" + self._common_filename(epochs)

    def _compile_train_function(self):
        """train_function(X, mask, Y, weight, exclude) -> cost (rnn_margin.py:100).  Y / weight are either
        the reference's dense [B, n_items] matrices or the ragged pair made by _prepare_input:
        Y = (target_offsets, target_ids), weight = w_neg[B]."""
        def train_function(X, mask, Y, weight, exclude=None):
            sl = self._split_rows
            if isinstance(Y, tuple):
                off, ids = Y
                if self.n_ranks > 1:
                    lo = self.rank * self.local_batch
                    base = off[lo]
                    ids = ids[base:off[lo + self.local_batch]]
                    off = off[lo:lo + self.local_batch + 1] - base
                default = self._default_target() if self.popularity_based else None
                return self.engine.train_step_margin(sl(X), sl(mask), off, ids, sl(weight), default,
                                                     exclude_seen=self.interactions_are_unique)
            return self.engine.train_step_margin_dense(sl(X), sl(mask), sl(Y), sl(weight))
        self.train_function = train_function

    def _prepare_input(self, sequences):
        """(X, mask, (target_offsets, target_ids), w_neg, exclude): the ragged equivalent of the dense
        Y / weight of rnn_margin.py:121-149, w = balance * n_targets / (n_items - n_targets - len)."""
        X, mask, seen = self._fill_inputs(sequences)
        B = len(sequences)
        off = np.zeros(B + 1, dtype=np.int32)
        ids = []
        w = np.zeros(B, dtype=np.float32)
        for i, row in enumerate(sequences):
            in_seq, target = row[1], row[2]       # rows may carry a 4th element (pre-encoded ids, rnn_base._gen_mini_batch)
            t = np.asarray(target)[:, 0].astype(np.int32)
            ids.append(t)
            off[i + 1] = off[i] + len(t)
            w[i] = self.balance * len(t) / (self.n_items - len(t) - len(in_seq))
        return (X, mask, (off, np.concatenate(ids) if ids else np.zeros(0, np.int32)), w, seen)

    def dense_targets(self, batch):
        """The reference's dense (Y, weight) [B, n_items] for a batch made by _prepare_input."""
        X, mask, (off, ids), w, seen = batch
        B = len(w)
        Y = np.tile(self._default_target().astype(np.float32), (B, 1))
        W = np.repeat(w[:, None], self.n_items, axis=1).astype(np.float32)
        for i in range(B):
            t = ids[off[i]:off[i + 1]]
            W[i, t] = -1
            Y[i, t] = 1
            if self.interactions_are_unique:
                W[i, seen[i]] = 0
                Y[i, seen[i]] = 0
        return Y, W

    def _default_target(self):
        if not hasattr(self, '_default_target_cache'):
            if not self.popularity_based:
                self._default_target_cache = np.zeros(self.n_items)
            else:
                num_users = self.dataset.training_set.n_users
                view_prob = self.dataset.item_popularity / num_users
                self._default_target_cache = np.minimum(1 - view_prob, (1 - self.min_access) * view_prob / self.min_access)
        return self._default_target_cache
