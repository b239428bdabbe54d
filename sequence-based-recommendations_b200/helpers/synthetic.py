"""Synthetic datasets in the reference's on-disk format.

MovieLens-1M-shaped generator: sequence lengths ~ shifted log-normal (mean ~165, min 20), items
~ Zipf(alpha) over the catalog without repeats inside a user (the reference assumes unique
interactions, rnn_base.py:66,154), ratings uniform in {1..5}; 10% validation / 10% test users as in
preprocess.py:18-19.  Files written: data/stats (5 lines, data_handling.py:89-97),
data/{train,val,test}_set_sequences (data_handling.py:142-145), data/train_set_triplets
(data_handling.py:52-57).
"""
import os

import numpy as np


def _zipf_probs(n_items, alpha):
    p = 1.0 / np.power(np.arange(1, n_items + 1, dtype=np.float64), alpha)
    return p / p.sum()


def sample_sequences(n_users, n_items, mean_len=165.0, min_len=20, max_len=None, alpha=1.0, seed=1234,
                     uniform_len=None):
    """List of int32 item-id arrays, one per user, no repeated item inside a user."""
    rng = np.random.default_rng(seed)
    probs = _zipf_probs(n_items, alpha)
    perm = rng.permutation(n_items)          # popularity rank -> item id
    if uniform_len is not None:
        lens = rng.integers(uniform_len[0], uniform_len[1] + 1, size=n_users)
    else:
        sigma = 1.0
        mu = np.log(max(mean_len - min_len, 1.0)) - sigma * sigma / 2
        lens = (min_len + rng.lognormal(mu, sigma, size=n_users)).astype(np.int64)
    cap = n_items - 1 if max_len is None else min(max_len, n_items - 1)
    lens = np.clip(lens, 2, cap)
    seqs = []
    if n_items > 20000:
        # large catalogs (BASELINE configs C3-C5): inverse-CDF draws with rejection of repeats -- O(L log N) per user
        # instead of the O(N) Gumbel pass below (the two differ only in the tail of very long sequences)
        cdf = np.cumsum(probs)
        for L in lens:
            got = np.zeros(0, dtype=np.int64)
            while len(got) < L:
                cand = np.searchsorted(cdf, rng.random(2 * int(L) + 16), side='right')
                cand = np.minimum(cand, n_items - 1)
                allc = np.concatenate([got, cand])
                _, first = np.unique(allc, return_index=True)
                got = allc[np.sort(first)]
            seqs.append(perm[got[:L]].astype(np.int32))
        return seqs
    for L in lens:
        # Gumbel top-L == sampling without replacement proportional to probs
        keys = np.log(probs) + rng.gumbel(size=n_items)
        top = np.argpartition(-keys, L - 1)[:L]
        rng.shuffle(top)
        seqs.append(perm[top].astype(np.int32))
    return seqs


def write_dataset(dirname, n_users, n_items, seed=1234, **kw):
    """Create <dirname>/{data,models,results} and return dirname (with trailing slash)."""
    if not dirname.endswith('/'):
        dirname += '/'
    for sub in ('data', 'models', 'results'):
        os.makedirs(dirname + sub, exist_ok=True)
    seqs = sample_sequences(n_users, n_items, seed=seed, **kw)
    rng = np.random.default_rng(seed + 1)
    ratings = [rng.integers(1, 6, size=len(s)) for s in seqs]
    order = rng.permutation(n_users)
    n_val = max(1, n_users // 10)
    n_test = max(1, n_users // 10)
    parts = {'val': order[:n_val], 'test': order[n_val:n_val + n_test], 'train': order[n_val + n_test:]}

    def stats(ids):
        inter = int(sum(len(seqs[u]) for u in ids))
        items = len(set(int(i) for u in ids for i in seqs[u]))
        longest = int(max(len(seqs[u]) for u in ids))
        return [len(ids), items, inter, longest]

    for name, ids in parts.items():
        with open(dirname + 'data/%s_set_sequences' % name, 'w') as f:
            for u in ids:
                f.write(str(int(u)) + ' ' + ' '.join('%d %d' % (i, r) for i, r in zip(seqs[u], ratings[u])) + '\n')
    with open(dirname + 'data/train_set_triplets', 'w') as f:
        for u in parts['train']:
            for i, r in zip(seqs[u], ratings[u]):
                f.write('%d %d %d\n' % (u, i, r))
    with open(dirname + 'data/stats', 'w') as f:
        f.write('set n_users n_items n_interactions longest_sequence\n')
        full = stats(order)
        full[1] = n_items
        f.write('Full ' + ' '.join(map(str, full)) + '\n')
        for name in ('train', 'val', 'test'):
            f.write(name.capitalize() + ' ' + ' '.join(map(str, stats(parts[name]))) + '\n')
    return dirname
