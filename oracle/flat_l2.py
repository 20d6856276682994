"""TEST INFRASTRUCTURE ONLY (never imported by the product).

CPU oracle of the exact fingerprint search.  The reference searches with FAISS
(eval.py:37-151; ``index.search(q, k_probe)``, eval.py:306); its exact index type 'l2' is
``faiss.IndexFlatL2`` (faiss-gpu 1.7.2 pinned in requirements.txt -- a third-party dependency that is NOT in
/root/reference nor in this image).  Published semantics restated here: D[i, r] = r-th smallest squared
Euclidean distance |q_i - x_j|^2, I[i, r] = its database position; -1 / +inf beyond ntotal.  Computed in
float64; ties (FAISS leaves their order unspecified) are broken towards the lower index.
PARITY UNPINNED: the reference ships no search fixtures and FAISS cannot be run here."""
import numpy as np


def flat_l2_search(db: np.ndarray, q: np.ndarray, k: int):
    db64, q64 = db.astype(np.float64), q.astype(np.float64)
    d = (q64 * q64).sum(1)[:, None] - 2.0 * (q64 @ db64.T) + (db64 * db64).sum(1)[None, :]
    order = np.argsort(d, axis=1, kind="stable")[:, :k]
    D = np.take_along_axis(d, order, axis=1)
    if order.shape[1] < k:
        pad = k - order.shape[1]
        order = np.concatenate([order, -np.ones((q.shape[0], pad), dtype=order.dtype)], axis=1)
        D = np.concatenate([D, np.full((q.shape[0], pad), np.inf)], axis=1)
    return D, order.astype(np.int64)
