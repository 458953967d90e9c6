"""GPU tests of the device band index / candidate join (csrc/index_join.cu) through the C ABI.

Integer work, so the bar is EXACT equality with the reference's data model: a dict of sets per band, collision
counting in a dict, ordering by (-collisions, id) -- reference lshrs/core/main.py:1088-1111 and :614 (what
MockStorage / RedisStorage + LSHRS._candidate_counts do).
"""

from __future__ import annotations

import math

import numpy as np
import pytest

from oracle import lshrs_oracle as oracle

pytestmark = pytest.mark.gpu


class Model:
    """The reference's bucket semantics in plain Python."""

    def __init__(self, nb):
        self.nb = nb
        self.b = [dict() for _ in range(nb)]

    def add(self, sig, ids):
        for s, i in zip(sig, ids):
            for band in range(self.nb):
                self.b[band].setdefault(s[band].tobytes(), set()).add(int(i))

    def remove(self, ids):
        gone = set(int(i) for i in ids)
        for band in self.b:
            for m in band.values():
                m -= gone

    def query(self, sig):
        out = []
        for s in sig:
            counts = {}
            for band in range(self.nb):
                for m in self.b[band].get(s[band].tobytes(), ()):
                    counts[m] = counts.get(m, 0) + 1
            out.append(sorted(counts.items(), key=lambda kv: (-kv[1], kv[0])))
        return out


def _lists(ix, sig, collisions=True):
    ix.query(sig)
    offs, counts, ids, coll = ix.fetch(collisions=True)
    return [list(zip(ids[o:o + c].tolist(), coll[o:o + c].tolist())) for o, c in zip(offs[:-1].tolist(), counts.tolist())]


@pytest.mark.parametrize("nb, bpb, keyspace", [(16, 2, 40), (4, 1, 6), (3, 8, 200), (32, 4, 5000), (1, 3, 2)])
def test_join_equals_dict_of_sets(nb, bpb, keyspace):
    from lshrs_b200.storage.device import DeviceIndex

    rng = np.random.default_rng(nb * 131 + bpb)
    pool = rng.integers(0, 256, size=(keyspace, bpb), dtype=np.uint8)     # few distinct keys: large buckets

    def sigs(n):
        return pool[rng.integers(0, keyspace, size=(n, nb))]              # (n, nb, bpb)

    ix, model = DeviceIndex(nb, bpb, device=0), Model(nb)
    next_id = 0
    added = 0
    Q = sigs(300)
    assert _lists(ix, Q) == model.query(Q) == [[] for _ in range(300)]   # empty index
    for rnd, n in enumerate((1, 700, 3000, 129, 5000)):
        S = sigs(n)
        if rnd == 2:    # ids far apart and large: more radix passes
            ids = (rng.permutation(n).astype(np.int64) * 7_000_003 + (1 << 40))
        else:
            ids = np.arange(next_id, next_id + n, dtype=np.int64)
        next_id += n
        ix.add(S, ids)
        model.add(S, ids)
        added += n
        assert len(ix) == added
        got, want = _lists(ix, Q), model.query(Q)
        assert got == want, f"round {rnd}"
        if rnd == 1:    # the same vectors again, same ids: SET semantics, nothing changes
            ix.add(S, ids)
            model.add(S, ids)
            assert _lists(ix, Q) == want
            # the same ids under OTHER keys: an id may sit in several buckets of a band
            S2 = sigs(n)
            ix.add(S2, ids)
            model.add(S2, ids)
            added += 2 * n
            assert _lists(ix, Q) == model.query(Q)
        if rnd == 3:
            gone = rng.choice(next_id, size=400, replace=False)
            ix.remove(gone)
            model.remove(gone)
            assert _lists(ix, Q) == model.query(Q)
            back = gone[:50]
            Sb = sigs(50)
            ix.add(Sb, back)                                              # removed ids come back
            model.add(Sb, back)
            added += 50
            assert _lists(ix, Q) == model.query(Q)
    # get_top_k prefix
    ix.query(Q)
    for k in (1, 7, 100):
        ids, counts = ix.topk(k)
        want = model.query(Q)
        for q in range(Q.shape[0]):
            w = [i for i, _ in want[q][:k]]
            assert counts[q] == len(w) and ids[q, :len(w)].tolist() == w and (ids[q, len(w):] == -1).all()
    ix.clear()
    assert len(ix) == 0 and _lists(ix, Q) == [[] for _ in range(300)]
    ix.close()


def test_join_with_huge_buckets_takes_the_global_workspace():
    """More than 4096 candidate slots for one query: the sort leaves shared memory."""
    from lshrs_b200.storage.device import DeviceIndex

    rng = np.random.default_rng(5)
    nb, bpb, n = 8, 2, 60_000
    S = rng.integers(0, 2, size=(n, nb, bpb), dtype=np.uint8)          # 4 keys per band: buckets of ~15 000
    ids = rng.permutation(n).astype(np.int64)
    ix, model = DeviceIndex(nb, bpb, device=0), Model(nb)
    ix.add(S, ids)
    model.add(S, ids)
    Q = rng.integers(0, 2, size=(6, nb, bpb), dtype=np.uint8)
    total, maxc = ix.query(Q)
    assert maxc > 4096 * 8
    assert _lists(ix, Q) == model.query(Q)
    ix.close()


def test_index_rejects_what_it_cannot_hold():
    from lshrs_b200 import LshxError
    from lshrs_b200.storage.device import DeviceIndex

    with pytest.raises(LshxError, match="8 bytes"):
        DeviceIndex(4, 9, device=0)
    with pytest.raises(LshxError, match="num_bands"):
        DeviceIndex(300, 2, device=0)
    ix = DeviceIndex(2, 2, device=0)
    with pytest.raises(LshxError, match="2\\^56"):
        ix.add(np.zeros((1, 2, 2), np.uint8), np.array([-1], dtype=np.int64))
    assert len(ix) == 0
    ix.add(np.zeros((1, 2, 2), np.uint8), np.array([5], dtype=np.int64))
    assert _lists(ix, np.zeros((1, 2, 2), np.uint8)) == [[(5, 2)]]
    with pytest.raises(LshxError, match="no query result"):
        ix.add(np.zeros((1, 2, 2), np.uint8), np.array([6], dtype=np.int64))
        ix.fetch()
    ix.close()


@pytest.fixture(scope="module")
def indexed_50k():
    """50 000 indexed vectors of dimension 768 (clustered, so that buckets are shared)."""
    import torch

    from lshrs_b200 import LSHRS, InMemoryStorage

    rng = np.random.default_rng(11)
    n, dim = 50_000, 768
    centers = rng.standard_normal((n // 8, dim)).astype(np.float32)
    X = (np.repeat(centers, 8, axis=0) + 0.15 * rng.standard_normal((n, dim))).astype(np.float32)
    lsh = LSHRS(dim=dim, num_perm=256, storage=InMemoryStorage(), vector_fetch_fn=lambda ids: X[np.asarray(ids)],
                device_index=True, device=0)
    lsh.index(list(range(n)), X)
    corpus = torch.from_numpy(X).to("cuda:0")
    Q = (X[rng.integers(0, n, 1024)] + 0.05 * rng.standard_normal((1024, dim))).astype(np.float32)
    return lsh, X, corpus, Q


def test_query_batch_device_index_equals_storage_path(indexed_50k):
    lsh, X, corpus, Q = indexed_50k
    storage_all = lsh.query_batch(Q, top_k=None)
    assert lsh.query_batch(Q, top_k=None, device_index=True) == storage_all
    assert sum(len(x) for x in storage_all) > 5 * len(Q)            # the probes do find their clusters
    assert lsh.query_batch(Q, top_k=10, device_index=True) == lsh.query_batch(Q, top_k=10)
    for kw in (dict(top_k=10, top_p=0.2), dict(top_k=None, top_p=0.95), dict(top_k=3, top_p=1.0)):
        a = lsh.query_batch(Q, corpus=corpus, **kw)                  # storage path + device corpus
        b = lsh.query_batch(Q, corpus=corpus, device_index=True, **kw)
        c = lsh.query_batch(Q[:64], device_index=True, **kw)         # device lists + vector_fetch_fn gather
        assert a == b, kw
        assert [[i for i, _ in r] for r in c] == [[i for i, _ in r] for r in a[:64]]
        ids, scores, counts = lsh.query_batch(Q, corpus=corpus, device_index=True, as_arrays=True, **kw)
        assert [list(zip(ids[i, :counts[i]].tolist(), scores[i, :counts[i]].tolist())) for i in range(len(Q))] == a


def test_query_batch_device_index_equals_the_oracle_pipeline(indexed_50k):
    """Against the reference's own steps restated by the oracle: hash_vector, per-band buckets, dict counting,
    (-collisions, id) order, top_k_cosine on the fetched vectors, rank-fraction cut (main.py:524-658)."""
    lsh, X, corpus, Q = indexed_50k
    projs = lsh._hasher.projections
    got = lsh.query_batch(Q[:48], top_k=10, top_p=0.5, corpus=corpus, device_index=True)
    sig_all = oracle.hash_batch_vectorized(projs, X)
    buckets = [dict() for _ in range(16)]
    for i, s in enumerate(sig_all):
        for b in range(16):
            buckets[b].setdefault(s[b].tobytes(), set()).add(i)
    for qi in range(48):
        sq = oracle.hash_vector(projs, Q[qi])
        counts = {}
        for b, key in enumerate(sq):
            for m in buckets[b].get(key, ()):
                counts[m] = counts.get(m, 0) + 1
        ordered = [i for i, _ in sorted(counts.items(), key=lambda kv: (-kv[1], kv[0]))]
        if not ordered:
            assert got[qi] == []
            continue
        ref = oracle.top_k_cosine(Q[qi], X[ordered], k=len(ordered))
        limit = oracle.top_p_limit(len(ordered), 0.5, 10)
        want = [(ordered[p], s) for p, s in ref[:limit]]
        assert [i for i, _ in got[qi]] == [i for i, _ in want]
        np.testing.assert_allclose([s for _, s in got[qi]], [s for _, s in want], atol=1e-5, rtol=0)


def test_device_mirror_tracks_ingest_delete_clear(indexed_50k):
    from lshrs_b200 import LSHRS, InMemoryStorage

    _, X, _, Q = indexed_50k
    lsh = LSHRS(dim=768, num_perm=256, storage=InMemoryStorage(), device_index=True, device=0, buffer_size=160)
    lsh.index(list(range(2000)), X[:2000])
    for i in range(2000, 2040):
        lsh.ingest(i, X[i])                       # flushes every 10 vectors (160 operations)
    probes = X[np.r_[0:64, 2000:2040]]
    same = lambda **kw: (lsh.query_batch(probes, **kw), lsh.query_batch(probes, device_index=True, **kw))
    a, b = same(top_k=None)
    assert a == b and all(i in a[j] for j, i in enumerate(np.r_[0:64, 2000:2040]))
    lsh.delete(list(range(0, 64, 2)))
    a, b = same(top_k=None)
    assert a == b and 0 not in a[0]
    lsh.clear()
    a, b = same(top_k=5)
    assert a == b == [[] for _ in range(len(probes))]


@pytest.mark.parametrize("dim, nb, r", [(768, 16, 16), (128, 16, 4), (64, 5, 20), (33, 3, 64)])
def test_latency_path_equals_hash_plus_batched_join(dim, nb, r):
    """lshx_index_query_vectors (hash_small + one lookup/join/emit launch into mapped memory) returns what
    hash_batch_packed + lshx_index_query + lshx_index_fetch return, for 1 .. 32 vectors and any capacity; queries
    with more than 4096 bucket entries report -1."""
    import lshrs_b200
    from lshrs_b200.storage.device import DeviceIndex

    rng = np.random.default_rng(dim + nb)
    n = 3000
    centers = rng.standard_normal((n // 6, dim)).astype(np.float32)
    X = (np.repeat(centers, 6, axis=0) + 0.2 * rng.standard_normal((n, dim))).astype(np.float32)
    h = lshrs_b200.LSHHasher(nb, r, dim, seed=7, device=0)
    ix = DeviceIndex(nb, h.bytes_per_band, device=0)
    Q = (X[rng.integers(0, n, 32)] + 0.05 * rng.standard_normal((32, dim))).astype(np.float32)
    Q[5] = 0                                                        # zero vector: flagged, still answered
    ids0, coll0, counts0, zero0 = ix.query_vectors(h, Q[:3], 8)     # empty index
    assert counts0.tolist() == [0, 0, 0]
    ix.add(h.hash_batch_packed(X), np.arange(n, dtype=np.int64) * 3 + 1)
    most = min(32, 65536 // (4 * dim))                              # rows the hasher's latency path takes
    for nq, cap in ((1, 10), (1, 4096), (7, 3), (most, 64), (most, 1)):
        ids, coll, counts, zero = ix.query_vectors(h, Q[:nq], cap)
        want = _lists(ix, h.hash_batch_packed(Q[:nq]))
        assert zero.tolist() == [1 if i == 5 else 0 for i in range(nq)]
        for q in range(nq):
            assert counts[q] == len(want[q])
            take = min(cap, len(want[q]))
            assert list(zip(ids[q, :take].tolist(), coll[q, :take].tolist())) == want[q][:take]
    # one bucket with 5000 members: beyond the shared-memory sort of the latency path
    big = np.repeat(h.hash_batch_packed(Q[:1]), 5000, axis=0)
    ix.add(big, np.arange(10_000_000, 10_005_000, dtype=np.int64))
    ids, coll, counts, zero = ix.query_vectors(h, Q[:2], 16)
    ix.query(h.hash_batch_packed(Q[:2]))
    offs = ix.fetch()[0]
    slots1 = int(offs[2] - offs[1])                                 # bucket entries query 1 matches
    want1 = len(_lists(ix, h.hash_batch_packed(Q[1:2]))[0])
    assert counts[0] == -1 and counts[1] == (want1 if slots1 <= 4096 else -1)
    with pytest.raises(lshrs_b200.LshxError):
        ix.query_vectors(h, np.zeros((33, dim), np.float32), 8)
    with pytest.raises(lshrs_b200.LshxError):
        ix.query_vectors(h, Q[:1], 5000)
    h.set_kernel("ffma")
    with pytest.raises(lshrs_b200.LshxError, match="pinned"):
        ix.query_vectors(h, Q[:1], 8)
    ix.close()
    h.close()


def test_small_adds_go_to_a_delta_run_with_identical_results():
    """Above 32 768 entries a small add is sorted on its own (delta run) instead of re-sorting the segment; queries
    search both runs.  Same lists as the dict-of-sets model through adds, duplicates of main-run entries, removals
    that hit both runs, re-adds, the latency path, bucket reads (which fold the runs) and the fold by growth."""
    import lshrs_b200
    from lshrs_b200.storage.device import DeviceIndex

    rng = np.random.default_rng(77)
    dim, nb, r, n0 = 64, 12, 10, 40_000
    centers = rng.standard_normal((n0 // 8, dim)).astype(np.float32)

    def vectors(n):
        return (centers[rng.integers(0, centers.shape[0], n)] + 0.25 * rng.standard_normal((n, dim))).astype(np.float32)

    h = lshrs_b200.LSHHasher(nb, r, dim, seed=3, device=0)
    ix, model = DeviceIndex(nb, h.bytes_per_band, device=0), Model(nb)
    X0 = vectors(n0)
    S0 = h.hash_batch_packed(X0)
    ids0 = rng.permutation(n0).astype(np.int64) * 5 + 2
    ix.add(S0, ids0)
    model.add(S0, ids0)
    Qv = (X0[rng.integers(0, n0, 200)] + 0.05 * rng.standard_normal((200, dim))).astype(np.float32)
    Q = h.hash_batch_packed(Qv)

    def check(tag):
        want = model.query(Q)
        assert _lists(ix, Q) == want, tag
        ids, coll, counts, _ = ix.query_vectors(h, Qv[:16], 64)            # the latency path sees both runs too
        for q in range(16):
            if counts[q] >= 0:
                take = min(64, len(want[q]))
                assert counts[q] == len(want[q]) and \
                    list(zip(ids[q, :take].tolist(), coll[q, :take].tolist())) == want[q][:take], (tag, q)

    check("main run only")
    next_id = 10_000_000
    for rnd, m in enumerate((500, 700, 1, 2000)):
        Xn = vectors(m)
        Sn = h.hash_batch_packed(Xn)
        idn = np.arange(next_id, next_id + m, dtype=np.int64)
        next_id += m
        ix.add(Sn, idn)
        model.add(Sn, idn)
        check(f"delta round {rnd}")
        if rnd == 0:
            # entries of the MAIN run again (same ids, same keys): SET semantics across the runs
            ix.add(S0[:300], ids0[:300])
            model.add(S0[:300], ids0[:300])
            check("duplicates of main-run entries in the delta run")
        if rnd == 1:
            gone = np.concatenate([ids0[100:400], idn[:50]])               # ids in the main run, in both, in the delta
            ix.remove(gone)
            model.remove(gone)
            check("removal across both runs")
            ix.add(S0[150:250], ids0[150:250])                             # removed ids come back (delta run)
            model.add(S0[150:250], ids0[150:250])
            check("re-add after removal")
        if rnd == 2:
            # a bucket read folds the runs (one id-ascending range per bucket) -- and must agree with the model
            keys = Q[:5].reshape(5 * nb, -1)
            bands = np.tile(np.arange(nb, dtype=np.int32), 5)
            offs, flat = ix.get_buckets(bands, keys)
            for t in range(5 * nb):
                assert set(flat[offs[t]:offs[t + 1]].tolist()) == model.b[int(bands[t])].get(keys[t].tobytes(), set())
            check("after the fold by a bucket read")
    Xn = vectors(15_000)                                                   # more than a quarter of the main run: fold
    Sn = h.hash_batch_packed(Xn)
    idn = np.arange(next_id, next_id + 15_000, dtype=np.int64)
    ix.add(Sn, idn)
    model.add(Sn, idn)
    check("fold by growth")
    ix.close()
    h.close()
