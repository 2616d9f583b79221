"""Shared checkers of the GPU parity tests (test infrastructure: wraps the oracles, never the product)."""
import numpy as np


def audit_codes(xr, wr, got, what):
    """Codes vs the fp64 re-search on the same bf16 operands (oracle.vq_oracle.audit_codes): a code may
    differ from the fp64 answer only where the distance gap is below 1e-6 (||x||^2 + ||e||^2); such rounding
    near-ties are counted and printed (expected 0), anything larger fails.  xr [N,C], wr [K+1,C]: fp32 arrays
    holding the bf16-rounded operands; got [N]."""
    from oracle import vq_oracle as vo
    a = vo.audit_codes(np.asarray(xr, np.float32), np.asarray(wr, np.float32), np.asarray(got))
    n, k = xr.shape[0], wr.shape[0] - 1
    print(f"[vq audit] {what}: {a['mismatch']} of {n} codes differ from the fp64 search "
          f"({a['near_tie']} rounding near-ties, {a['real']} real, worst relative gap {a['worst']:.2e})")
    assert a["real"] == 0, f"{what}: {a['real']} real code mismatches (worst relative gap {a['worst']:.3e})"
    assert a["near_tie"] <= max(1, (n * k) >> 28), f"{what}: {a['near_tie']} near-tie mismatches"
    return a


def bf16_operands(x_rows, weight):
    """(rows, weight[K+1]) as the search kernel sees them: rows and the K code rows rounded to bf16 (held in fp32)."""
    from oracle import vq_oracle as vo
    x_rows = np.asarray(x_rows, np.float32)
    weight = np.asarray(weight, np.float32)
    return vo.bf16_round(x_rows), np.concatenate([vo.bf16_round(weight[:-1]), weight[-1:]], 0)


def grad_report(got, ref, floor_frac=1e-3):
    """got / ref: dicts name -> gradient tensor (any device).  Returns (worst, cosines, median) where worst is a
    list of (rel-RMS error, name) sorted descending - the error of a tensor is ||a-b|| / max(||b||, floor_frac *
    median norm), so tensors whose true gradient is ~0 (e.g. the attention k-bias: softmax is invariant to it) are
    judged on an absolute scale - and cosines a list of (cosine similarity, name) sorted ascending."""
    import torch
    norms = {n: float(r.double().norm()) for n, r in ref.items()}
    typical = float(np.median([v for v in norms.values() if v > 0]))
    worst, cosines = [], []
    for n, r in ref.items():
        assert n in got and got[n] is not None, f"no gradient for {n}"
        a, b = got[n].detach().double().cpu().flatten(), r.detach().double().cpu().flatten()
        worst.append((float((a - b).norm()) / max(norms[n], floor_frac * typical), n))
        if norms[n] > floor_frac * typical:
            cosines.append((float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)), n))
    worst.sort(reverse=True)
    cosines.sort()
    return worst, cosines, worst[len(worst) // 2][0]


def force_codes(monkeypatch, forced):
    """Teacher forcing for comparisons against a run whose codes are known (the reference's golden run): the
    product's search kernel still runs, then the rows whose code differs are overwritten with the forced code -
    gathered row and the loss partial sum corrected accordingly.  Returns a dict that receives 'differ' (how many
    rows the product had chosen differently) on every call."""
    import torch
    from dynamicvectorquantization_b200 import kernels as kn
    real = kn.vq_search_gather
    info = {}

    def forced_search(x_bf16, codebook, weight_f32, x_f32=None, row_mask=None, loss_acc=None, **kw):
        codes, xq_b, xq_f = real(x_bf16, codebook, weight_f32, x_f32=x_f32, row_mask=row_mask, loss_acc=loss_acc, **kw)
        want = forced.to(codes.device).reshape(-1)
        diff = (codes != want).nonzero().flatten()
        info["differ"] = int(diff.numel())
        if diff.numel():
            rows = (x_f32 if x_f32 is not None else x_bf16.float())[diff]
            m = 1.0 if row_mask is None else row_mask[diff]
            w_new, w_old = weight_f32[want[diff]], weight_f32[codes[diff]]
            if loss_acc is not None:
                loss_acc += ((((w_new - rows) ** 2).sum(1) - ((w_old - rows) ** 2).sum(1)) * m).sum()
            if xq_b is not None:
                xq_b[diff] = w_new.to(xq_b.dtype)
            if xq_f is not None:
                xq_f[diff] = w_new
        return want.clone(), xq_b, xq_f

    monkeypatch.setattr(kn, "vq_search_gather", forced_search)
    return info
