"""Dataset access for the RNN path -- Python 3 host mirror of the reference's
helpers/data_handling.py (DataHandler :12-113, SequenceGenerator :115-174).

Same on-disk format (``data/stats``, ``data/{train,val,test}_set_sequences`` with lines
``user item rating item rating ...``, ``data/train_set_triplets``) and the same public surface
(``DataHandler.training_set(...)`` yields ``(sequence, user_id)``; ``.epochs``; ``item_popularity``).
Differences that matter for throughput: every file is parsed ONCE into numpy arrays (the
reference re-splits every text line on every epoch, data_handling.py:142-145), and a sequence is
a float64 array of shape [L, 2] (column 0 item id, column 1 rating) instead of a list of
``[item, rating]`` lists -- ``len(seq)``, ``seq[a:b]`` and ``row[0]`` behave the same.
"""
import os
import random

import numpy as np

DEFAULT_DIR = '../../data/'


class SequenceGenerator(object):
    """Iterates over the user sequences of one file (reference :115-174)."""

    def __init__(self, filename, shuffle=False):
        self.filename = filename
        self.shuffle = shuffle
        self.epochs = 0.
        self._seqs = None

    def load(self):
        seqs = []
        with open(self.filename, 'r') as f:
            for line in f:
                tok = line.split()
                if not tok:
                    continue
                vals = np.asarray(tok[1:], dtype=np.float64)
                n = len(vals) // 2
                seqs.append((tok[0], vals[:2 * n].reshape(n, 2)))
        self._seqs = seqs

    @property
    def lines(self):
        if self._seqs is None:
            self.load()
        return self._seqs

    def __call__(self, min_length=2, max_length=None, length_choice='max', subsequence='contiguous', epochs=np.inf):
        seqs = self.lines
        counter = 0
        self.epochs = 0.
        while counter < epochs:
            counter += 1
            print("Opening file ({})".format(counter))
            if self.shuffle:
                random.shuffle(seqs)
            for j, (user_id, sequence) in enumerate(seqs):
                self.epochs = counter - 1 + j / len(seqs)
                cap = len(sequence) if max_length is None else max_length
                if len(sequence) < min_length:
                    continue
                if length_choice == 'random':
                    length = np.random.randint(min_length, min(cap, len(sequence)) + 1)
                elif length_choice == 'max':
                    length = min(cap, len(sequence))
                else:
                    raise ValueError('Unrecognised length_choice option. Authorised values are "random" and "max" ')
                if length < len(sequence):
                    if subsequence == 'random':
                        keep = sorted(random.sample(range(len(sequence)), length))
                        sequence = sequence[keep]
                    elif subsequence == 'contiguous':
                        start = np.random.randint(0, len(sequence) - length + 1)
                        sequence = sequence[start:start + length]
                    elif subsequence == 'begining':
                        sequence = sequence[:length]
                    else:
                        raise ValueError('Unrecognised subsequence option. Authorised values are "random", '
                                         '"contiguous" and "begining".')
                yield sequence, user_id


class DataHandler(object):
    """Directory layout and statistics of one dataset (reference :12-113)."""

    def __init__(self, dirname, extended_training_set=False, shuffle_training=False):
        self.dirname = self._get_path(dirname)
        self.extended_training_set = extended_training_set
        train = 'data/train_set_sequences+' if extended_training_set else 'data/train_set_sequences'
        self.training_set = SequenceGenerator(self.dirname + train, shuffle=shuffle_training)
        self.validation_set = SequenceGenerator(self.dirname + 'data/val_set_sequences')
        self.test_set = SequenceGenerator(self.dirname + 'data/test_set_sequences')
        self._load_stats()

    def training_set_triplets(self):
        with open(self.dirname + 'data/train_set_triplets') as f:
            for line in f:
                u, i, r = line.split()[:3]
                yield {'user_id': int(u), 'item_id': int(i), 'rating': float(r)}

    @property
    def item_popularity(self):
        """Occurrences of each item in the training set (cached to .npy like the reference :59-74)."""
        ts = self.training_set
        if not hasattr(ts, '_item_pop'):
            cache = self.dirname + 'data/training_set_item_popularity.npy'
            if os.path.isfile(cache):
                ts._item_pop = np.load(cache)
            else:
                pop = np.zeros(self.n_items)
                with open(self.dirname + 'data/train_set_triplets') as f:
                    for line in f:
                        pop[int(line.split()[1])] += 1
                ts._item_pop = pop
                np.save(cache, pop)
        return ts._item_pop

    def _get_path(self, dirname):
        here, there = os.path.exists(dirname), os.path.exists(DEFAULT_DIR + dirname + '/')
        if here and there:
            print('WARNING: ambiguous directory name, both "' + dirname + '" and "' + DEFAULT_DIR + dirname +
                  '" exist. "' + dirname + '" is used.')
        if here:
            return dirname if dirname.endswith('/') else dirname + '/'
        if there:
            return DEFAULT_DIR + dirname + '/'
        raise ValueError('Dataset not found')

    def _load_stats(self):
        with open(self.dirname + 'data/stats', 'r') as f:
            f.readline()
            rows = [list(map(int, f.readline().split()[1:5])) for _ in range(4)]
        self.n_users, self.n_items, self.n_interactions, self.longest_sequence = rows[0]
        for part, row in zip((self.training_set, self.validation_set, self.test_set), rows[1:]):
            part.n_users, part.n_items, part.n_interactions, part.longest_sequence = row
        if self.extended_training_set:
            self.training_set.n_users, self.training_set.n_items = self.n_users, self.n_items
            self.training_set.n_interactions += (self.validation_set.n_interactions + self.test_set.n_interactions) // 2
