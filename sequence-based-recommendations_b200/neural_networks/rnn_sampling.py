"""RNNSampling -- host mirror of neural_networks/rnn_sampling.py:14-194: sampled-column output
(BlackoutLayer, sparse_lstm.py:23-56) with the BPR / BPRI / TOP1 / Blackout losses.  The
arithmetic is `sbr_train_step_sampled` (include/sbr_b200.h)."""
import random
from bisect import bisect

import numpy as np

from . import rnn_base as rnn


class RNNSampling(rnn.RNNBase):
    def __init__(self, loss_function="Blackout", sampling=32, last_layer_tanh=False, last_layer_init=1.,
                 diversity_bias=0.0, sampling_bias=0., **kwargs):
        super().__init__(**kwargs)
        self.last_layer_init = last_layer_init
        self.last_layer_tanh = last_layer_tanh
        self.diversity_bias = diversity_bias
        self.sampling = sampling
        self.sampling_bias = sampling_bias
        if loss_function is None:
            loss_function = "Blackout"
        if loss_function not in ("BPR", "BPRI", "TOP1", "Blackout"):
            raise ValueError("Unknown loss function")
        self.loss_function_name = loss_function
        self.loss_name = loss_function
        self.name = "RNN with sampling loss"

    _test_softmax = True   # the sampling test function applies a softmax first (rnn_sampling.py:140-157)

    def _get_model_filename(self, epochs):
        filename = "rnn_sampling_" + self.loss_function_name + "_"
        if self.sampling_bias > 0.:
            filename += "p" + str(self.sampling_bias)
        filename += "s" + str(self.sampling) + "_ini" + str(self.last_layer_init) + "_db" + str(self.diversity_bias)
        return filename + "_" + self._common_filename(epochs)

    def _prepare_networks(self, n_items):
        if self.sampling < 1:
            self.effective_sampling = int(self.sampling * n_items)
        else:
            self.effective_sampling = int(self.sampling)
        super()._prepare_networks(n_items)

    def _engine_extra_kwargs(self):
        return dict(n_samples=self.effective_sampling, last_layer_tanh=self.last_layer_tanh)

    def _compile_train_function(self):
        """train_function(X, mask, Y, samples, pop, exclude) -> cost (rnn_sampling.py:128).  Every rank
        passes the targets of the whole global batch: they are softmax columns of the Blackout loss
        (rnn_sampling.py:68-72) whatever row they belong to."""
        def train_function(X, mask, Y, samples, pop, exclude=None):
            sl = self._split_rows
            return self.engine.train_step_sampled(sl(X), sl(mask), sl(Y), samples, sl(pop), Y_all=Y,
                                                  row_offset=self.rank * self.local_batch)
        self.train_function = train_function

    def _popularity_sample(self):
        if not hasattr(self, '_cumsum'):
            self._cumsum = np.cumsum(np.power(self.dataset.item_popularity, self.sampling_bias))
        return bisect(self._cumsum, random.uniform(0, self._cumsum[-1]))

    def _prepare_input(self, sequences):
        """(X, mask, Y, samples, pop, exclude) (rnn_sampling.py:165-194): one shared vector of negative
        samples per batch, uniform or popularity**sampling_bias weighted."""
        X, mask, seen = self._fill_inputs(sequences)
        Y = np.array([int(t[2][0][0]) for t in sequences], dtype=np.int32)
        pop = np.power(self.dataset.item_popularity[Y], self.diversity_bias).astype(np.float32)
        if getattr(self, '_assembling_test_batch', False):
            # validation / test batches: the test function never reads the samples (rnn_sampling.py:140-157); no draw,
            # so that validation leaves the training RNG streams untouched
            samples = np.zeros(self.effective_sampling, dtype=np.int32)
        elif self.sampling_bias > 0:
            samples = np.array([self._popularity_sample() for _ in range(self.effective_sampling)], dtype=np.int32)
        else:
            samples = np.random.choice(self.n_items, self.effective_sampling).astype(np.int32)
        return (X, mask, Y, samples, pop, seen)
