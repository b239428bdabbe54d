"""RNNOneHot -- host mirror of neural_networks/rnn_one_hot.py:13-106: recurrent stack ->
Dense(n_items, softmax), cost = mean(categorical_crossentropy / popularity**diversity_bias)
(+ L2/L1 on the output bias).  The arithmetic is `sbr_train_step_cce` (include/sbr_b200.h)."""
import numpy as np

from . import rnn_base as rnn


class RNNOneHot(rnn.RNNBase):
    loss_name = "CCE"

    def __init__(self, diversity_bias=0.0, regularization=0.0, **kwargs):
        super().__init__(**kwargs)
        self.diversity_bias = np.float32(diversity_bias)
        self.regularization = regularization
        self.name = "RNN with categorical cross entropy"

    def _get_model_filename(self, epochs):
        return "rnn_cce_db" + str(self.diversity_bias) + "_r" + str(self.regularization) + "_" + self._common_filename(epochs)

    def _engine_extra_kwargs(self):
        return dict(regularization=float(self.regularization))

    _supports_device_batches = True

    def _compile_train_function(self):
        """train_function(X, mask, Y, pop, exclude) -> cost (rnn_one_hot.py:61, rnn_base.py:185).
        `exclude` is accepted and ignored, as in the reference (on_unused_input='ignore').  A CompactBatch (rows as
        (sequence, start, length) triples of the uploaded training set) takes the device-assembly entry point; every
        rank sends only its own rows."""
        def train_function(X, mask=None, Y=None, pop=None, exclude=None):
            sl = self._split_rows
            if isinstance(X, rnn.CompactBatch):
                return self.engine.train_step_cce_rows(sl(X.seq), sl(X.start), sl(X.length), sl(X.Y), sl(X.pop))
            return self.engine.train_step_cce(sl(X), sl(mask), sl(Y), sl(pop))
        self.train_function = train_function

    def _compact_from_triples(self, seq, start, length, Y):
        pop = np.power(self.dataset.item_popularity[Y], self.diversity_bias).astype(np.float32)
        return (rnn.CompactBatch(seq, start, length, Y, pop),)

    def _split_rows(self, arr):
        """This rank's rows of a global-batch array: row r goes to rank r % n_ranks.  The batch builder emits rows in
        ascending length (nested prefixes), so a contiguous split would hand the last rank the longest sequences on
        every step and the other ranks would sit in the all-reduce waiting for it.  The CCE loss is a plain sum over
        rows, so any partition gives the same gradient."""
        if self.n_ranks == 1:
            return arr
        return np.ascontiguousarray(arr[self.rank::self.n_ranks])

    def _prepare_input(self, sequences):
        """(X, mask, Y, pop, exclude) for a list of [user_id, input_sequence, targets]
        (rnn_one_hot.py:83-106); `exclude` is the ragged list of seen ids instead of a dense [B,N]."""
        Y = np.array([int(t[2][0][0]) for t in sequences], dtype=np.int32)       # first and only target
        pop = np.power(self.dataset.item_popularity[Y], self.diversity_bias).astype(np.float32)
        if sequences and all(len(r) > 4 for r in sequences):
            # device assembly: nothing but three integers per row leaves the host
            tri = np.array([r[4] for r in sequences], dtype=np.int32)
            return (rnn.CompactBatch(tri[:, 0].copy(), tri[:, 1].copy(), tri[:, 2].copy(), Y, pop),)
        X, mask, seen = self._fill_inputs(sequences)
        return (X, mask, Y, pop, seen)
