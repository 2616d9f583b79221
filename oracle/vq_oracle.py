"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the vector-quantization bottleneck.

A numpy restatement of the reference algorithm in
``/root/reference/modules/vector_quantization/quantize2_mask.py``; every function cites the lines
it follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module - it is the checker, never the product path.

Pinning: the reference ships no tests or golden vectors for this path, so the
oracle is pinned against outputs of the reference's own Python classes run in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/vq_*.npz``; checked by
``tests/test_oracle_golden.py``).
"""
import numpy as np


def bf16_round(a):
    """Round-to-nearest-even fp32 -> bf16 -> fp32 (the operand precision of the CUDA search)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return rounded.astype(np.uint32).view(np.float32).reshape(a.shape)


def compute_distances(x, weight):
    """quantize2_mask.py:29-48.  x [N,C] fp32, weight [K+1,C] fp32 (padding row excluded, :31).

    distances = (||x||^2 + ||e||^2) - 2 x e^T, evaluated in fp32 like torch.addmm(alpha=-2).
    """
    cb_t = weight[:-1, :].T.astype(np.float32)
    x = x.astype(np.float32)
    x_sq = (x * x).sum(axis=1, keepdims=True, dtype=np.float32)
    e_sq = (cb_t * cb_t).sum(axis=0, keepdims=True, dtype=np.float32)
    return (x_sq + e_sq) + np.float32(-2.0) * (x @ cb_t)


def find_nearest_embedding(x, weight):
    """quantize2_mask.py:50-55: argmin over codes, first minimum wins (torch/numpy semantics)."""
    return compute_distances(x, weight).argmin(axis=-1).astype(np.int64)


def nearest_fp64(x, weight):
    """Tie audit: distances in fp64 (without the row constant); returns (idx, best, second)."""
    cb = weight[:-1, :].astype(np.float64)
    d = (cb * cb).sum(axis=1)[None, :] - 2.0 * (x.astype(np.float64) @ cb.T)
    idx = d.argmin(axis=1)
    part = np.partition(d, 1, axis=1)
    return idx.astype(np.int64), part[:, 0], part[:, 1]


def audit_codes(x, weight, got, rel=1e-6, chunk=2048):
    """SURVEY.md 8d tie policy, as one function.  x [N,C] and weight [K+1,C] are the operands the
    search really saw (already bf16-rounded, held in fp32); got [N] are the codes under test.

    Every row is re-searched in fp64 (exact products of bf16 values, lowest index wins exact ties,
    like quantize2_mask.py:50-55 / torch.argmin).  A row whose code differs from the fp64 answer is
      * a rounding near-tie when the fp64 distance gap between the two codes is below
        rel * (||x||^2 + ||e||^2) - the fp32 accumulation cannot resolve it -, counted and reported;
      * a real mismatch otherwise.
    Returns dict(mismatch, near_tie, real, worst) with worst = max gap / (||x||^2 + ||e||^2)."""
    cb = weight[:-1, :].astype(np.float64)
    e_sq = (cb * cb).sum(axis=1)
    got = np.asarray(got).reshape(-1).astype(np.int64)
    n = x.shape[0]
    mismatch = near = real = 0
    worst = 0.0
    for r0 in range(0, n, chunk):
        xs = x[r0:r0 + chunk].astype(np.float64)
        d = e_sq[None, :] - 2.0 * (xs @ cb.T)
        best = d.argmin(axis=1)
        g = got[r0:r0 + chunk]
        bad = np.nonzero(best != g)[0]
        if bad.size == 0:
            continue
        gap = d[bad, g[bad]] - d[bad, best[bad]]
        scale = (xs[bad] * xs[bad]).sum(axis=1) + e_sq[best[bad]]
        ratio = gap / scale
        mismatch += int(bad.size)
        near += int((ratio < rel).sum())
        real += int((ratio >= rel).sum())
        worst = max(worst, float(ratio.max()))
    return dict(mismatch=mismatch, near_tie=near, real=real, worst=worst)


def update_buffers(x, idx, cluster_size_ema, embed_ema, decay, restart_rows=None):
    """quantize2_mask.py:66-105 (single process: no all_reduce).  Returns new (cs_ema, embed_ema).

    restart_rows [K,C] are the rows the reference would draw with randperm (:97); passing None
    disables the restart (restart_unused_codes=False).
    """
    K, C = embed_ema.shape
    counts = np.bincount(idx.reshape(-1), minlength=K).astype(np.float32)            # :77-83
    sums = np.zeros((K, C), np.float32)
    np.add.at(sums, idx.reshape(-1), x.reshape(-1, C).astype(np.float32))             # :84
    cs = cluster_size_ema * np.float32(decay) + counts * np.float32(1 - decay)        # :90
    em = embed_ema * np.float32(decay) + sums * np.float32(1 - decay)                 # :91
    if restart_rows is not None:
        usage = (cs >= 1).astype(np.float32)                                          # :102
        em = em * usage[:, None] + restart_rows * (1 - usage[:, None])                # :103
        cs = cs * usage + (1 - usage)                                                 # :104-105
    return cs.astype(np.float32), em.astype(np.float32)


def update_embedding(cluster_size_ema, embed_ema, eps=1e-5):
    """quantize2_mask.py:107-115: Laplace-smoothed normalisation -> new weight[:K]."""
    K = cluster_size_ema.shape[0]
    n = cluster_size_ema.sum(dtype=np.float32)
    norm = n * (cluster_size_ema + np.float32(eps)) / (n + np.float32(K * eps))
    return (embed_ema / norm[:, None]).astype(np.float32)


def vq_forward(x, weight, mask=None, beta=0.25):
    """quantize2_mask.py:117-132 + :157-191 on a flattened [N,C] latent (eval mode).

    Returns (x_q [N,C], loss scalar, idx [N]).  x_q is the straight-through value
    x + (weight[idx] - x) exactly as :182 evaluates it in fp32 (equal to weight[idx] up to 1 ulp).
    loss = beta*mean((xq-x)^2 m) + mean((xq-x)^2 m), mean over N*C (:172-179).
    """
    idx = find_nearest_embedding(x, weight)
    xq = weight[idx]
    d2 = (xq - x) ** 2
    if mask is not None:
        d2 = d2 * mask.reshape(-1, 1)
    m = d2.mean(dtype=np.float64)
    x = x.astype(np.float32)
    return x + (xq - x), np.float32(beta * m + m), idx
