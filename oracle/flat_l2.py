"""TEST INFRASTRUCTURE ONLY (never imported by the product).

CPU oracle of the exact fingerprint search.  The reference searches with FAISS
(eval.py:37-151; ``index.search(q, k_probe)``, eval.py:306); its exact index type 'l2' is
``faiss.IndexFlatL2`` (faiss-gpu 1.7.2 pinned in requirements.txt -- a third-party dependency that is NOT in
/root/reference nor in this image).  Published semantics restated here: D[i, r] = r-th smallest squared
Euclidean distance |q_i - x_j|^2, I[i, r] = its database position; -1 / +inf beyond ntotal.  Computed in
float64; ties (FAISS leaves their order unspecified) are broken towards the lower index.
The reference ships no search fixtures and FAISS cannot be run here, so FAISS ITSELF WAS NEVER RUN; this oracle is
pinned instead to tests/golden/search_expected.npz -- an independent float64 brute force (direct sum of squared
differences, lexsort by (distance, id); tests/golden/make_search_golden.py) over a database written in the reference's
memmap layout (tests/test_oracle_golden.py)."""
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


def song_level_ranking(db: np.ndarray, q: np.ndarray, k_probe: int, file_of: np.ndarray, query_file: int = -1,
                       first_valid: int = 0):
    """eval.py:300-336 restated for one query sequence with integer file ids (the reference keys its histogram by the
    lookup string): candidates = every id of the segment-level top-k_probe lists; score = np.mean(np.sum(q_match *
    candidate_seq, axis=1)); hist[file] += score; ranking by descending score."""
    _, I = flat_l2_search(db, q, k_probe)
    cand = I[np.where(I >= 0)].flatten()
    hist = {}
    sl = q.shape[0]
    for cid in cand:
        if cid < first_valid or file_of[cid] == query_file:
            continue
        seq = db[cid:cid + sl].astype(np.float64)
        qm = q[:seq.shape[0]].astype(np.float64)
        hist[int(file_of[cid])] = hist.get(int(file_of[cid]), 0.0) + float(np.mean(np.sum(qm * seq, axis=1)))
    files = sorted(hist, key=hist.get, reverse=True)
    return np.array(files, dtype=np.int64), np.array([hist[f] for f in files])
