"""Early stopping rules -- host mirror of the reference's helpers/early_stopping.py:4-85
(same flags, same decisions)."""


def early_stopping_command_parser(parser):
    parser.add_argument('--es_m', dest='early_stopping_method', choices=['WorstTimesX', 'StopAfterN', 'None'],
                        help='Early stopping method', default='None')
    parser.add_argument('--es_n', help='N parameter (for StopAfterN)', default=5, type=int)
    parser.add_argument('--es_x', help='X parameter (for WorstTimesX)', default=2., type=float)
    parser.add_argument('--es_min_wait', help='Mininum wait before stopping (for WorstTimesX)', default=1., type=float)
    parser.add_argument('--es_LiB', help='Lower is better for validation score.', action='store_true')


def get_early_stopper(args):
    hib = not args.es_LiB
    if args.early_stopping_method == 'StopAfterN':
        return StopAfterN(n=args.es_n, higher_is_better=hib)
    if args.early_stopping_method == 'WorstTimesX':
        return WaitWorstCaseTimesX(x=args.es_x, min_wait=args.es_min_wait, higher_is_better=hib)
    return None


class EarlyStopperBase(object):
    def __init__(self, higher_is_better=True):
        self.higher_is_better = higher_is_better

    def __call__(self, epochs, val_costs):
        scores = list(val_costs) if self.higher_is_better else [-v for v in val_costs]
        return self.decideStopping(epochs, scores)

    def decideStopping(self, epochs, val_costs):
        raise NotImplementedError


class StopAfterN(EarlyStopperBase):
    """Stop after n consecutive evaluations without improvement."""

    def __init__(self, n=3, **kwargs):
        super().__init__(**kwargs)
        self.n = n

    def decideStopping(self, epochs, val_costs):
        if len(val_costs) <= self.n:
            return False
        tail = val_costs[-self.n - 1:]
        return all(later <= earlier for earlier, later in zip(tail[:-1], tail[1:]))


class WaitWorstCaseTimesX(EarlyStopperBase):
    """Stop once the wait since the best score exceeds x times the longest wait between two bests."""

    def __init__(self, x=2., min_wait=1., **kwargs):
        super().__init__(**kwargs)
        self.x = x
        self.min_wait = min_wait

    def decideStopping(self, epochs, val_costs):
        best, best_epoch, longest = val_costs[0], epochs[0], 0
        for epoch, score in zip(epochs[1:], val_costs[1:]):
            if score > best:
                longest = max(longest, epoch - best_epoch)
                best, best_epoch = score, epoch
        waited = epochs[-1] - best_epoch
        if longest == 0:
            return waited > self.min_wait
        print('current wait : ', round(waited, 3), ' longest wait : ', round(longest, 3), ' ratio : ',
              waited / longest, ' / ', self.x)
        return waited > max(self.min_wait, longest * self.x)
