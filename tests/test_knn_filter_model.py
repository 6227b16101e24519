"""The tensor-core k-NN (map-merge_b200/csrc/knn_tc.cu) as a numpy model: a lower-bound filter with an error of up to
ES (||a'||^2 + ||b'||^2), k picks per row from per-chunk minima whose column rides in the five low mantissa bits, a threshold
from the EXACT distances of the picks, one candidate bit per column, exact evaluation in ascending column order.  The model
plants the worst the error bound allows (every accumulator pushed to either end of its interval) and checks that the result
is still the brute-force scan's (distance, index) order — i.e. that the kernel's logic does not depend on the filter being
accurate, only on `acc + (1 - ES) ||a'||^2 <= d`.  The CUDA kernels themselves are compared with the exact scan on the GPU
(tests/test_parity_gpu.py, tests/test_full_size_gpu.py); this file pins the argument on the CPU."""
import numpy as np
import pytest

ES = np.float32(3.1e-5)


def seq_dist(a, B):
    """acc += (a_t - b_t)^2 for t = 0 .. D-1 in FP32 — the exact scan's distance (matching.cpp:31-93 -> flann::L2_Simple)."""
    acc = np.zeros(len(B), np.float32)
    for t in range(B.shape[1]):
        diff = (a[t] - B[:, t]).astype(np.float32)
        acc = (acc + diff * diff).astype(np.float32)
    return acc


def brute(A, B, k):
    idx = np.full((len(A), k), -1, np.int64)
    dist = np.zeros((len(A), k), np.float32)
    for i, a in enumerate(A):
        d = seq_dist(a, B)
        order = np.lexsort((np.arange(len(B)), d))[:k]
        idx[i, :len(order)] = order
        dist[i, :len(order)] = d[order]
    return idx, dist


def model(A, B, k, rng, adversarial):
    mu = np.concatenate([A, B]).mean(0).astype(np.float32)
    Ac, Bc = (A - mu).astype(np.float64), (B - mu).astype(np.float64)
    na, nb = (Ac * Ac).sum(1), (Bc * Bc).sum(1)
    idx = np.full((len(A), k), -1, np.int64)
    dist = np.zeros((len(A), k), np.float32)
    evals = 0
    n_chunks = (len(B) + 31) // 32
    for i, a in enumerate(A):
        d_exact = seq_dist(a, B)
        # what the tensor core may deliver: v = acc + (1 - ES) na anywhere in [d - 2 ES (na + nb), d]
        slack = 2.0 * float(ES) * (na[i] + nb)
        if adversarial == "low":
            v = d_exact - slack
        elif adversarial == "high":
            v = d_exact.astype(np.float64)
        else:
            v = d_exact - slack * rng.random(len(B))
        na_low = np.float32((1.0 - float(ES)) * na[i])
        acc = (v - float(na_low)).astype(np.float32)
        # pass 0: per 32-column chunk the smallest accumulator, column in the five low mantissa bits; the k smallest chunks
        picks = []
        for ch in range(n_chunks):
            seg = acc[ch * 32:(ch + 1) * 32].copy()
            bits = (seg.view(np.uint32) & np.uint32(0xffffffe0)) | np.arange(len(seg), dtype=np.uint32)
            stuffed = bits.view(np.float32)
            m = stuffed.min()
            picks.append((m, ch * 32 + int(np.float32(m).view(np.uint32) & 31)))
        picks.sort(key=lambda t: t[0])
        cols = [c for _, c in picks[:k]]
        assert len(set(cols)) == len(cols) and all(0 <= c < len(B) for c in cols)
        # knn_thr_kernel
        if len(cols) < k:
            thr = np.float32(np.inf)
        else:
            dmax = np.float32(max(d_exact[c] for c in cols))
            thr = np.float32((dmax - na_low) + np.float32(1e-6) * (np.float32(1.0) + abs(dmax) + na_low))
        # pass 1 + knn_eval_kernel
        cand = np.nonzero(acc <= thr)[0]
        evals += len(cand)
        bd = [np.float32(np.inf)] * k
        bi = [-1] * k
        for j in cand:  # ascending column order, strict <
            d = d_exact[j]
            if d < bd[k - 1]:
                pos = 0
                while not d < bd[pos]:
                    pos += 1
                bd.insert(pos, d); bi.insert(pos, int(j))
                bd.pop(); bi.pop()
        for t in range(k):
            if bi[t] >= 0:
                idx[i, t] = bi[t]; dist[i, t] = bd[t]
    return idx, dist, evals / max(len(A), 1)


def clustered(rng, n, d=33, clusters=6, dup=0.1):
    centres = rng.uniform(0, 100, (clusters, d))
    x = centres[rng.integers(0, clusters, n)] + rng.normal(0, 0.05, (n, d))
    x = x.astype(np.float32)
    for _ in range(int(dup * n)):  # exact duplicates: ties that only the index order resolves
        x[rng.integers(0, n)] = x[rng.integers(0, n)]
    return x


@pytest.mark.parametrize("adversarial", ["low", "high", "random"])
@pytest.mark.parametrize("k", [1, 5])
def test_filter_model_reproduces_the_exact_scan(adversarial, k):
    rng = np.random.default_rng(5)
    A, B = clustered(rng, 96), clustered(rng, 330)
    want_i, want_d = brute(A, B, k)
    got_i, got_d, evals = model(A, B, k, rng, adversarial)
    np.testing.assert_array_equal(got_i, want_i)
    np.testing.assert_array_equal(got_d.view(np.uint32), want_d.view(np.uint32))
    assert evals >= k


def test_filter_model_with_fewer_columns_than_k():
    rng = np.random.default_rng(6)
    A, B = clustered(rng, 8), clustered(rng, 3)
    want_i, want_d = brute(A, B, 5)
    got_i, got_d, _ = model(A, B, 5, rng, "random")
    np.testing.assert_array_equal(got_i, want_i)
    np.testing.assert_array_equal(got_d.view(np.uint32), want_d.view(np.uint32))
