"""Optional perturbations of the training sequences -- host mirror of
neural_networks/sequence_noise.py:4-95 (all off by default)."""
import numpy as np


def sequence_noise_command_parser(parser):
    parser.add_argument('--n_dropout', help='Dropout probability', default=0., type=float)
    parser.add_argument('--n_swap', help="Probability of swapping two consecutive items", default=0., type=float)
    parser.add_argument('--n_shuf', help="Probability of swapping two random items", default=0., type=float)
    parser.add_argument('--n_shuf_std', help="The distance between the two items to be swapped is drawn from a normal "
                        "distribution whose std is defined by this parameter", default=5., type=float)
    parser.add_argument('--n_ratings', help='Probability of changing the rating.', default=0., type=float)


def get_sequence_noise(args):
    return SequenceNoise(dropout=args.n_dropout, swap=args.n_swap, ratings_perturb=args.n_ratings, shuf=args.n_shuf,
                         shuf_std=args.n_shuf_std)


class SequenceNoise(object):
    def __init__(self, dropout=0., swap=0., ratings_perturb=0., shuf=0., shuf_std=0.):
        self.dropout, self.swap, self.ratings_perturb, self.shuf, self.shuf_std = dropout, swap, ratings_perturb, shuf, shuf_std
        for value, label in ((dropout, 'Dropout'), (swap, 'Swapping probability'),
                             (ratings_perturb, 'Rating perturbation probability')):
            if value < 0. or value >= 1.:
                raise ValueError(label + ' should be in [0,1)')
        parts = []
        if dropout > 0:
            parts.append("do" + str(dropout))
        if swap > 0:
            parts.append("sw" + str(swap))
        if ratings_perturb > 0:
            parts.append("rp" + str(ratings_perturb))
        if shuf > 0:
            parts.append("sh" + str(shuf) + "-" + str(shuf_std))
        self.name = "_".join(parts)

    @property
    def active(self):
        return self.dropout > 0 or self.swap > 0 or self.shuf > 0 or self.ratings_perturb > 0

    def __call__(self, sequence_generator):
        for sequence, user in sequence_generator:
            if not self.active:
                yield sequence, user
                continue
            seq = np.array(sequence, dtype=np.float64, copy=True)
            if self.dropout > 0.:
                seq = seq[np.random.random(len(seq)) >= self.dropout]
                if len(seq) < 2:
                    continue
            if self.swap > 0.:
                i = 0
                while i < len(seq) - 1:
                    if np.random.random() < self.swap:
                        seq[[i, i + 1]] = seq[[i + 1, i]]
                        i += 1          # never move the same item twice
                    i += 1
            if self.shuf > 0.:
                for i in range(len(seq)):
                    if np.random.random() < self.shuf:
                        other = max(0, min(len(seq) - 1, int(np.random.randn() * self.shuf_std) + i))
                        seq[[i, other]] = seq[[other, i]]
            if self.ratings_perturb > 0:
                for i in range(len(seq)):
                    if np.random.random() < self.ratings_perturb:
                        if np.random.random() < 0.5:
                            seq[i, 1] = min(5, seq[i, 1] + 0.5)
                        else:
                            seq[i, 1] = max(1, seq[i, 1] - 0.5)
            yield seq, user
