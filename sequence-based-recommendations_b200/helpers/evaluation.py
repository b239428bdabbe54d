"""Top-N metrics -- host mirror of the reference's helpers/evaluation.py:17-226 (Evaluator).

Same metric definitions (sps, recall, precision, ndcg, user/item coverage, blockbuster share,
assr); instances are (goal ids, predicted ids) pairs added one by one, exactly like the reference's
validation (rnn_base.py:358-371) and test loops (test.py:55-69)."""
import numpy as np


class Evaluator(object):
    def __init__(self, dataset, k=10):
        self.instances = []
        self.dataset = dataset
        self.k = k
        self.metrics = {'sps': self.short_term_prediction_success, 'recall': self.average_recall,
                        'precision': self.average_precision, 'ndcg': self.average_ndcg,
                        'item_coverage': self.item_coverage, 'user_coverage': self.user_coverage,
                        'assr': self.assr, 'blockbuster_share': self.blockbuster_share}

    def add_instance(self, goal, predictions):
        self.instances.append([list(goal), list(predictions)])

    def _top(self, prediction):
        return prediction[:min(len(prediction), self.k)]

    def _hits(self, goal, prediction):
        return set(goal) & set(self._top(prediction))

    def average_precision(self):
        tot = sum(len(self._hits(g, p)) / float(min(len(p), self.k)) for g, p in self.instances if len(p) > 0)
        return tot / len(self.instances)

    def average_recall(self):
        tot = sum(len(self._hits(g, p)) / float(len(g)) for g, p in self.instances if len(g) > 0)
        return tot / len(self.instances)

    def average_ndcg(self):
        tot = 0.
        for goal, prediction in self.instances:
            if len(prediction) == 0:
                continue
            gset = set(goal)
            dcg = ideal = 0.
            for rank, p in enumerate(self._top(prediction)):
                gain = 1. / np.log2(2 + rank)
                if rank < len(goal):
                    ideal += gain
                if p in gset:
                    dcg += gain
            tot += dcg / ideal
        return tot / len(self.instances)

    def short_term_prediction_success(self):
        return sum(int(g[0] in self._top(p)) for g, p in self.instances) / len(self.instances)

    def sps(self):
        return self.short_term_prediction_success()

    def user_coverage(self):
        return sum(int(len(self._hits(g, p)) > 0) for g, p in self.instances) / len(self.instances)

    def get_all_goals(self):
        return [g for goal, _ in self.instances for g in goal]

    def get_strict_goals(self):
        return [goal[0] for goal, _ in self.instances]

    def get_all_predictions(self):
        return [p for _, prediction in self.instances for p in self._top(prediction)]

    def get_correct_predictions(self):
        out = []
        for goal, prediction in self.instances:
            out.extend(self._hits(goal, prediction))
        return out

    def get_correct_strict_predictions(self):
        out = []
        for goal, prediction in self.instances:
            out.extend(set([goal[0]]) & set(self._top(prediction)))
        return out

    def item_coverage(self):
        return len(set(self.get_correct_predictions()))

    def blockbuster_share(self):
        """Share of the correct predictions that fall in the 1% most popular items."""
        correct = self.get_correct_predictions()
        if len(correct) == 0:
            return 0
        n_pop = self.dataset.n_items // 100
        pop_items = set(np.argpartition(-self.dataset.item_popularity, n_pop)[:n_pop].tolist())
        return len([i for i in correct if i in pop_items]) / len(correct)

    def get_rank_comparison(self):
        out = []
        for goal, prediction in self.instances:
            pos = np.argsort(prediction)[goal]
            out.extend(list(enumerate(pos)))
        return out

    def assr(self):
        nb = getattr(self, 'nb_of_dp', 0)
        return self.dataset.n_items / nb if nb and nb > 0 else 1
