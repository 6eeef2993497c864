"""Host-side evaluation protocol shared by recommendation models: the mirror of
/root/reference/RecModel/base_model.py:34-179 for the pieces the WMF path touches.

``eval_prec`` is implemented by the model on the device (fused kernel K3). ``eval_topn`` /
``compute_hit`` keep the reference's sampled Recall@N protocol, including its draw order from
the global NumPy RNG, and call the model's (device) ``rank``.
"""
import numpy as np


def iter_rows_two_matrices(A, B):
    """(row, A data, A indices, B data, B indices) for every row (base_model.py:10-20)."""
    for i in range(A.shape[0]):
        a0, a1 = A.indptr[i], A.indptr[i + 1]
        b0, b1 = B.indptr[i], B.indptr[i + 1]
        yield i, A.data[a0:a1], A.indices[a0:a1], B.data[b0:b1], B.indices[b0:b1]


class RecModel:
    """Base class: same evaluation scheme for every model (base_model.py:34-49)."""

    def train(self):
        pass

    def predict(self, user_item):
        pass

    def rank(self, items, user, topn=None):
        pass

    def compute_hit(self, elem, rand_sampled, topn, dtype="float32"):
        """Hits per cut-off for one user (base_model.py:51-98): one draw of rand_sampled+1
        candidate ids and one slot; every held-out item overwrites the slot and is ranked."""
        user, _, _, test_dat, test_idx = elem
        if len(test_dat) == 0:
            return np.zeros(topn.shape, dtype=dtype)
        cand = np.random.randint(0, self.num_items, size=(rand_sampled + 1))
        slot = np.random.randint(0, rand_sampled - (2 * topn.max()))
        hits = np.zeros(topn.shape, dtype=dtype)
        kmax = int(topn.max())
        for item in test_idx:
            cand[slot] = item
            top = self.rank(items=cand, users=user, topn=kmax)
            for pos in range(len(topn)):
                if item in top[:topn[pos]]:
                    hits[pos] += 1
        return hits

    def eval_topn(self, test_mat, train_mat=None, eval_mat=None, topn=[10], rand_sampled=1000, cores=1,
                  random_state=None, dtype="float32"):
        """Sampled Recall@N (base_model.py:100-148). ``cores`` is accepted for signature
        compatibility; ranking runs on the GPU, the protocol loop on one host thread, which
        keeps the reference's cores=1 RNG draw order."""
        super_mat = test_mat
        if train_mat is not None:
            super_mat = super_mat + train_mat
        if eval_mat is not None:
            super_mat = super_mat + eval_mat
        if random_state is not None:
            np.random.seed(random_state)
        if not isinstance(topn, np.ndarray):
            raise ValueError("Topn has to be a np.array")
        hits = np.zeros(topn.shape, dtype=dtype)
        for elem in iter_rows_two_matrices(super_mat, test_mat):
            hits += self.compute_hit(elem, rand_sampled=rand_sampled, topn=topn)
        recall = hits / len(test_mat.nonzero()[0])
        return {f"Recall@{topn[pos]}": recall[pos] for pos in range(len(topn))}
