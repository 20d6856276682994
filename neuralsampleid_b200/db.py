"""Fingerprint database: the reference's on-disk layout and an exact GPU search.

Layout (test_fp.py:158-171, eval.py:154-196): ``{name}.mm`` float32 memmap (n, d), ``{name}_shape.npy``,
``{name}_lookup.json``.  Search: the reference builds a FAISS index over the memmap and calls
``index.search(q, k_probe)`` (eval.py:37-151, 306); ``FlatL2Index`` reproduces the exact variant
(index type 'l2' = ``faiss.IndexFlatL2``: squared L2 distances ascending, int64 ids) on the GPU:

    Y[q, j] = |d_j|^2 - 2 <q, d_j>      one tcgen05 GEMM per database chunk (database rows = weight operand,
                                        |d|^2 = per-column shift, a1 = -2 q), fp32-parity bf16x3 engine
    k smallest of every row             grafp_topk_rows_fwd (warp per (query, column split), threshold scan)
    merge chunks / splits, + |q|^2      grafp_topk_merge_fwd

FAISS is not in this image: the parity oracle is the exact float64 search (oracle/flat_l2.py)."""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib, ops
from ._lib import GrafpError, check
from ._prep import make_linear


def save_fingerprints(out_dir: str, name: str, emb: np.ndarray, lookup=None) -> None:
    """Writes (n, d) float32 fingerprints in the reference's layout (test_fp.py:158-171)."""
    os.makedirs(out_dir, exist_ok=True)
    emb = np.ascontiguousarray(emb, dtype=np.float32)
    mm = np.memmap(os.path.join(out_dir, name + ".mm"), dtype="float32", mode="w+", shape=emb.shape)
    mm[:] = emb
    mm.flush()
    np.save(os.path.join(out_dir, name + "_shape.npy"), np.array(emb.shape))
    if lookup is not None:
        with open(os.path.join(out_dir, name + "_lookup.json"), "w") as f:
            json.dump(lookup, f)


def load_fingerprints(source_dir: str, name: str) -> Tuple[np.ndarray, np.ndarray]:
    """load_memmap_data (eval.py:154-196): (memmap (n, d) float32 with NaN -> 0, shape)."""
    shape = np.load(os.path.join(source_dir, name + "_shape.npy"))
    data = np.memmap(os.path.join(source_dir, name + ".mm"), dtype="float32", mode="r+",
                     shape=(int(shape[0]), int(shape[1])))
    data[np.isnan(data)] = 0.0
    return data, shape


@torch.no_grad()
def create_fp_db(gsim, batches, out: torch.Tensor) -> int:
    """Fingerprints of a stream of host batches into a host array: the loop of the reference's ``create_ref_db`` /
    ``create_query_db`` (test_fp.py:92-171: ``model(x, x)`` per chunk, ``z_i`` appended) as a pipeline over a captured
    model (``graphed.GraphedSimCLR``): the host -> device copy of batch i+1 and the device -> host copy of batch i-1
    overlap the kernels of batch i (double-buffered staging on both sides).

    ``gsim``: one captured model, or a LIST of captures of the same model (each with its own buffers): batch i then runs
    on lane i % len(gsim), every lane on its own stream.  At the reference's call shape (chunks of <= 128 segments,
    generate.py:40-46) the kernels of one chunk leave most SMs idle from stage 3 on (64- and 32-CTA grids), so two or
    three lanes overlap; large captures (thousands of segments) fill the machine and want one lane.
    ``batches``: iterable of host tensors (b, n_mels, n_frames) float32 with b <= the captured batch (pinned memory for
    asynchronous copies); ``out``: host tensor (>= total, d) float32 (pinned likewise), filled in order.  Returns the
    number of fingerprints written.  Rows of a short batch beyond b are computed on stale input and dropped (segments
    are independent: SURVEY section 8e)."""
    lanes = list(gsim) if isinstance(gsim, (list, tuple)) else [gsim]
    L = len(lanes)
    dev = lanes[0].input.device
    cap = lanes[0].input.shape[0]
    cur = torch.cuda.current_stream(dev)
    s_comp = [cur] + [torch.cuda.Stream(device=dev) for _ in range(L - 1)]
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    D = 2 * L                                   # staging depth: two batches per lane
    xin = [torch.empty_like(lanes[0].input) for _ in range(D)]
    zout = [torch.empty_like(lanes[0].z) for _ in range(D)]
    e_in = [torch.cuda.Event() for _ in range(D)]
    e_in_free = [torch.cuda.Event() for _ in range(D)]
    e_comp = [torch.cuda.Event() for _ in range(D)]
    e_out_free = [torch.cuda.Event() for _ in range(D)]
    for s in s_comp[1:] + [s_in, s_out]:
        s.wait_stream(cur)
    n = 0
    for i, x in enumerate(batches):
        b = int(x.shape[0])
        if b > cap or tuple(x.shape[1:]) != tuple(lanes[0].input.shape[1:]):
            raise GrafpError("create_fp_db: batch of shape %s does not fit the captured model %s"
                             % (tuple(x.shape), tuple(lanes[0].input.shape)))
        if n + b > out.shape[0]:
            raise GrafpError("create_fp_db: output holds %d rows, batch %d needs %d" % (out.shape[0], i, n + b))
        q, lane = i % D, i % L
        g, sc = lanes[lane], s_comp[lane]
        with torch.cuda.stream(s_in):
            if i >= D:
                s_in.wait_event(e_in_free[q])
            xin[q][:b].copy_(x, non_blocking=True)
            e_in[q].record(s_in)
        with torch.cuda.stream(sc):
            sc.wait_event(e_in[q])
            g.input.copy_(xin[q], non_blocking=True)
            e_in_free[q].record(sc)
            g.replay()
            if i >= D:
                sc.wait_event(e_out_free[q])
            zout[q].copy_(g.z, non_blocking=True)
            e_comp[q].record(sc)
        with torch.cuda.stream(s_out):
            s_out.wait_event(e_comp[q])
            out[n:n + b].copy_(zout[q][:b], non_blocking=True)
            e_out_free[q].record(s_out)
        n += b
    for s in s_comp[1:] + [s_in, s_out]:
        cur.wait_stream(s)
    return n


class FlatL2Index:
    """Exact squared-L2 index with the ``faiss.IndexFlatL2`` calling convention used by eval.py:
    ``add(x)``, ``search(q, k) -> (D, I)`` (float32 (nq, k) ascending, int64 (nq, k); -1 / inf when k > ntotal)."""

    CHUNK = 65536          # database rows per GEMM (Y chunk = nq x 65536 fp32)
    QBLOCK = 2048          # queries per pass

    def __init__(self, d: int, device="cuda:0"):
        self.d = int(d)
        self.device = torch.device(device)
        self.ntotal = 0
        self._chunks = []          # (prepared Linear over the chunk rows padded to 32, valid rows, first id)
        self._rows = []            # the added blocks as given (device), for reconstruct / sequence scores
        self._flat = None

    def add(self, x) -> None:
        x = torch.as_tensor(np.asarray(x, dtype=np.float32) if not torch.is_tensor(x) else x)
        if x.dim() != 2 or x.shape[1] != self.d:
            raise ValueError("expected (n, %d) vectors" % self.d)
        x = x.to(self.device, torch.float32).contiguous()
        self._rows.append(x)
        self._flat = None
        lib = _lib.load()
        for c0 in range(0, x.shape[0], self.CHUNK):
            rows = x[c0:c0 + self.CHUNK]
            n = rows.shape[0]
            npad = (n + 31) // 32 * 32                                 # tcgen05 tiles need n % 32 == 0
            w = torch.zeros((npad, self.d), device=self.device, dtype=torch.float32)
            w[:n] = rows
            norms = torch.full((npad,), float("inf"), device=self.device, dtype=torch.float32)   # padding never wins
            with torch.cuda.device(self.device):
                check(lib.grafp_row_sumsq(C.c_void_p(w.data_ptr()), n, self.d, C.c_void_p(norms.data_ptr()),
                                          ops._stream(w)), "row_sumsq")
            self._chunks.append((make_linear(w, None, norms), n, self.ntotal))
            self.ntotal += n

    def search(self, q, k: int, engine: Optional[int] = None):
        if not 1 <= k <= 32:
            raise GrafpError("FlatL2Index.search: 1 <= k <= 32")
        as_numpy = not torch.is_tensor(q)
        q = torch.as_tensor(np.asarray(q, dtype=np.float32)) if as_numpy else q
        q = q.to(self.device, torch.float32).contiguous()
        nq = q.shape[0]
        D = torch.full((nq, k), float("inf"), device=self.device, dtype=torch.float32)
        I = torch.full((nq, k), -1, device=self.device, dtype=torch.int64)
        lib = _lib.load()
        P = C.c_void_p
        for q0 in range(0, nq, self.QBLOCK):
            qb = q[q0:q0 + self.QBLOCK]
            nb = qb.shape[0]
            if nb == 0 or not self._chunks:
                continue
            qn = torch.empty((nb,), device=self.device, dtype=torch.float32)
            q2 = (qb * -2.0).contiguous()                               # exact scaling
            splits = max(1, min(16, (148 * 8) // max(nb, 1)))           # enough warps to fill the GPU for few queries
            parts = len(self._chunks) * splits
            pv = torch.empty((nb, parts, k), device=self.device, dtype=torch.float32)
            pi = torch.empty((nb, parts, k), device=self.device, dtype=torch.int64)
            with torch.cuda.device(self.device):
                st = ops._stream(qb)
                check(lib.grafp_row_sumsq(P(qb.data_ptr()), nb, self.d, P(qn.data_ptr()), st), "row_sumsq")
                for ci, (lin, n, first) in enumerate(self._chunks):
                    y = ops.linear(q2, lin, engine=engine)              # (nb, npad) = |d|^2 - 2 q.d  (inf in the padding)
                    sl_v, sl_i = pv[:, ci * splits:(ci + 1) * splits], pi[:, ci * splits:(ci + 1) * splits]
                    # the partial buffers are (nb, parts, k): write this chunk's `splits` lists through a
                    # temporary contiguous block, then place them
                    tv = torch.empty((nb, splits, k), device=self.device, dtype=torch.float32)
                    ti = torch.empty((nb, splits, k), device=self.device, dtype=torch.int64)
                    check(lib.grafp_topk_rows_fwd(P(y.data_ptr()), y.stride(0), nb, y.shape[1], first, k, splits,
                                                  P(tv.data_ptr()), P(ti.data_ptr()), st), "topk_rows")
                    sl_v.copy_(tv)
                    sl_i.copy_(ti)
                check(lib.grafp_topk_merge_fwd(P(pv.data_ptr()), P(pi.data_ptr()), nb, parts, k, P(qn.data_ptr()),
                                               P(D[q0:q0 + nb].data_ptr()), P(I[q0:q0 + nb].data_ptr()), st), "topk_merge")
        if as_numpy:
            return D.cpu().numpy(), I.cpu().numpy()
        return D, I

    def vectors(self) -> torch.Tensor:
        """The indexed vectors as one (ntotal, d) device tensor (what eval.py keeps as ``fake_recon_index``)."""
        if self._flat is None:
            self._flat = (torch.cat(self._rows) if len(self._rows) != 1 else self._rows[0]) if self._rows else \
                torch.empty((0, self.d), device=self.device, dtype=torch.float32)
        return self._flat

    def sequence_scores(self, q, cand) -> torch.Tensor:
        """Song-level match scores of eval.py:322-331 for a query sequence q (sl, d) and candidate start ids cand
        (nc,) int64: mean over the sequence of <q[t], db[cand + t]> (shorter at the end of the database)."""
        q = torch.as_tensor(q).to(self.device, torch.float32).contiguous()
        cand = torch.as_tensor(cand).to(self.device, torch.int64).contiguous()
        db = self.vectors()
        out = torch.empty((cand.numel(),), device=self.device, dtype=torch.float32)
        P = C.c_void_p
        with torch.cuda.device(self.device):
            check(_lib.load().grafp_sequence_score_fwd(P(q.data_ptr()), q.shape[0], self.d, P(db.data_ptr()), db.shape[0],
                                                       P(cand.data_ptr()), cand.numel(), P(out.data_ptr()),
                                                       ops._stream(q)), "sequence_score")
        return out


def song_level_ranking(index: FlatL2Index, q, k_probe: int, file_of: np.ndarray, query_file: int = -1,
                       first_valid: int = 0):
    """eval.py:300-336 for one query sequence: segment-level top-k_probe search, every returned id is a candidate start,
    each candidate adds its sequence score to the histogram bin of the file it belongs to (ids below ``first_valid`` --
    the dummy database -- and the query's own file are skipped); returns (file ids sorted by score descending, scores).
    ``file_of`` (ntotal,) int array: file index of every database row."""
    _, I = index.search(q, k_probe)
    I = I.cpu().numpy() if torch.is_tensor(I) else I
    cand = I[np.where(I >= 0)].flatten()
    keep = (cand >= first_valid) & (file_of[cand] != query_file)
    cand = cand[keep]
    if cand.size == 0:
        return np.zeros((0,), dtype=np.int64), np.zeros((0,), dtype=np.float64)
    scores = index.sequence_scores(q, cand).cpu().numpy().astype(np.float64)
    files = file_of[cand]
    uniq, inv = np.unique(files, return_inverse=True)
    hist = np.bincount(inv, weights=scores, minlength=uniq.size)
    order = np.argsort(-hist, kind="stable")
    return uniq[order], hist[order]
