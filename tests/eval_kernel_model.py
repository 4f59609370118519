"""Executable model of the selection algorithm in mmssl_b200/csrc/eval.cu (one user), used by the CPU suite to check
the algorithm itself -- key packing, threshold + candidate buffer, bitonic compaction, metric arithmetic -- against
the oracle while no GPU is at hand.  Mirrors the kernel statement by statement; appends within a sweep are shuffled
to model the unordered shared-memory atomics."""
import numpy as np

CAP, THREADS, MAXK = 512, 128, 64


def eval_key(score: np.float32, item: int) -> int:
    b = int(np.float32(score).view(np.uint32))
    b = (~b & 0xFFFFFFFF) if (b & 0x80000000) else (b | 0x80000000)
    return (b << 32) | (0xFFFFFFFF - item)


def key_score(key: int) -> np.float32:
    b = key >> 32
    b = (b & 0x7FFFFFFF) if (b & 0x80000000) else (~b & 0xFFFFFFFF)
    return np.uint32(b).view(np.float32)


def bitonic_desc(a: list, n: int) -> None:
    k = 2
    while k <= n:
        j = k >> 1
        while j > 0:
            for i in range(n):
                p = i ^ j
                if p > i:
                    x, y = a[i], a[p]
                    desc = (i & k) == 0
                    if (x < y) if desc else (x > y):
                        a[i], a[p] = y, x
            j >>= 1
        k <<= 1


def compact(keys: list, cnt: int, thr: int, kmax: int):
    n = cnt
    np2 = 64
    while np2 < n:
        np2 <<= 1
    for i in range(n, np2):
        keys[i] = 0
    bitonic_desc(keys, np2)
    if n >= kmax:
        thr, cnt = keys[kmax - 1], kmax
    return cnt, thr


def rank_one_user(rating: np.ndarray, train_sorted: np.ndarray, held_sorted: np.ndarray, Ks, rng: np.random.Generator):
    n_items = rating.shape[0]
    kmax = max(Ks)
    keys = [0] * CAP
    cnt, thr = 0, 0
    train = set(int(x) for x in train_sorted)
    for base in range(0, n_items, THREADS):
        js = np.arange(base, min(base + THREADS, n_items))
        rng.shuffle(js)
        for j in js:
            s = np.float32(rating[j]) + np.float32(0.0)
            key = eval_key(s, int(j))
            if key > thr and int(j) not in train:
                assert cnt < CAP
                keys[cnt] = key
                cnt += 1
        if cnt > CAP - THREADS:
            cnt, thr = compact(keys, cnt, thr, kmax)
    cnt, thr = compact(keys, cnt, thr, kmax)
    m = min(cnt, kmax)
    ranked = [0xFFFFFFFF - (keys[p] & 0xFFFFFFFF) for p in range(m)]
    held = set(int(x) for x in held_sorted)
    H = 0
    for p, it in enumerate(ranked):
        if it in held:
            H |= 1 << p
    disc = [1.0 / np.log2(float(i + 2)) for i in range(MAXK)]
    nh_all = bin(H).count("1")
    n_pos = float(len(held_sorted))
    out = np.zeros((4, len(Ks)))
    for q, K in enumerate(Ks):
        kk = min(K, m)
        Hk = H if kk >= 64 else H & ((1 << kk) - 1)
        nh = bin(Hk).count("1")
        dcg = idcg = 0.0
        for i in range(kk):
            if (Hk >> i) & 1:
                dcg += disc[i]
            if i < nh_all:
                idcg += disc[i]
        out[0, q] = nh / kk if kk > 0 else np.nan
        out[1, q] = nh / n_pos if n_pos > 0 else 0.0
        out[2, q] = dcg / idcg if idcg > 0 else 0.0
        out[3, q] = 1.0 if nh > 0 else 0.0
    return np.array(ranked, np.int64), out
