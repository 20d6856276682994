#!/usr/bin/env python
"""Fixture for the exact fingerprint search (reference eval.py:37-151 index type 'l2').

FAISS (faiss-gpu 1.7.2 in the reference's requirements.txt) is NOT in this image and was NEVER RUN to make this
fixture.  What pins the search instead is an INDEPENDENT exact computation: a database written in the reference's
on-disk layout exactly as test_fp.py:158-171 writes it ({name}.mm float32 memmap, {name}_shape.npy,
{name}_lookup.json), read back the way eval.py:154-196 reads it, and searched by a float64 brute force that shares no
code with oracle/flat_l2.py or the product: per query, sum_c (q_c - x_c)^2 accumulated in float64 with math.fsum-grade
pairwise numpy sums, candidates ordered by (distance, id) with a lexsort.  IndexFlatL2's published semantics: squared
Euclidean distances ascending with int64 database positions.

Run:  python tests/golden/make_search_golden.py   (writes tests/golden/search_db/ and search_expected.npz)"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "search_db")


def main():
    rng = np.random.Generator(np.random.PCG64(2024))
    n, d, nq, k = 1536, 128, 24, 20
    fp = rng.standard_normal((n, d)).astype(np.float32)
    fp /= np.linalg.norm(fp, axis=1, keepdims=True)
    fp[700] = fp[33]                                       # an exact duplicate pair: tie -> lower id first
    lookup = ["song_%03d" % (i // 48) for i in range(n)]
    os.makedirs(OUT, exist_ok=True)
    # --- written as test_fp.py:158-171 does ---
    arr_shape = (len(fp), fp.shape[-1])
    arr = np.memmap(os.path.join(OUT, "ref_db.mm"), dtype="float32", mode="w+", shape=arr_shape)
    arr[:] = fp[:]
    arr.flush()
    del arr
    np.save(os.path.join(OUT, "ref_db_shape.npy"), arr_shape)
    json.dump(lookup, open(os.path.join(OUT, "ref_db_lookup.json"), "w"))
    # --- read back as eval.py:154-160 (load_memmap_data) does ---
    shape = tuple(np.load(os.path.join(OUT, "ref_db_shape.npy")))
    db = np.memmap(os.path.join(OUT, "ref_db.mm"), dtype="float32", mode="r", shape=shape)
    qi = rng.integers(0, n, size=nq)
    q = np.asarray(db)[qi] + 0.05 * rng.standard_normal((nq, d)).astype(np.float32)
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    q[0] = np.asarray(db)[33]                              # hits the duplicate pair exactly
    D = np.empty((nq, k), dtype=np.float64)
    I = np.empty((nq, k), dtype=np.int64)
    x64 = np.asarray(db).astype(np.float64)
    for i in range(nq):
        diff = x64 - q[i].astype(np.float64)[None, :]
        dist = np.einsum("nc,nc->n", diff, diff)           # sum_c (q_c - x_c)^2, no |q|^2 - 2qx + |x|^2 expansion
        order = np.lexsort((np.arange(n), dist))[:k]
        D[i], I[i] = dist[order], order
    np.savez_compressed(os.path.join(HERE, "search_expected.npz"), q=q, D=D, I=I, source=qi, k=np.int64(k))
    print("db", shape, "queries", q.shape, "top-1 is the source item for %d of %d" % (int((I[:, 0] == qi).sum()), nq))


if __name__ == "__main__":
    main()
