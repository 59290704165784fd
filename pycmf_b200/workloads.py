"""Synthetic workloads of BASELINE.json (`configs`), generated on the device in fixed row blocks so that
the data are identical for every shard count.  Used by bench.py and the full-size tests.
"""
import math

import numpy as np

ROW_BLOCK = 1000

# name -> shape / solver description (BASELINE.json configs[0..4]; c3 / c4 / c5 can be row-scaled)
CONFIGS = {
    "c1": dict(n=1000, d=500, l=20, k=10, solver="mu", sparse=False, x_link="linear", y_link="linear"),
    "c2": dict(n=20000, d=5000, l=50, k=32, solver="newton", sparse=False, x_link="linear", y_link="logit"),
    "c3": dict(n=2000000, d=200000, l=6, k=64, solver="mu", sparse=True, x_link="linear", y_link="linear",
               nnz_per_row=100),
    "c4": dict(n=2000000, d=200000, l=6, k=128, solver="newton", sparse=True, x_link="logit", y_link="logit",
               nnz_per_row=100, sg_sample_ratio=0.1),
    "c5": dict(n=200000, d=50000, l=1000, k=256, solver="mu", sparse=False, x_link="linear", y_link="linear"),
}

SOLVER_PARAMS = {
    "c1": dict(),
    # signed factors: with the non-negativity projection the reference's full Newton steps diverge on this
    # data (objective 1.6e4 -> 1e17 in 6 iterations, measured with the oracle); unconstrained they converge
    "c2": dict(alpha=0.5, l1_reg=0.0, l2_reg=0.1, U_non_negative=False, V_non_negative=False,
               Z_non_negative=False, hessian_pertubation=0.2),
    "c3": dict(),
    "c4": dict(alpha=0.5, l1_reg=0.0, l2_reg=0.1, U_non_negative=False, V_non_negative=False,
               Z_non_negative=False, hessian_pertubation=0.2, sg_sample_ratio=0.1),
    "c5": dict(),
}


def describe(name, scale=1.0):
    c = dict(CONFIGS[name])
    if scale != 1.0:
        c["n"] = max(ROW_BLOCK, int(c["n"] * scale) // ROW_BLOCK * ROW_BLOCK)
        if c["sparse"]:
            c["d"] = max(1000, int(c["d"] * scale))
    return c


def _gen(torch, device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def generate(be, name, r0, r1, scale=1.0, seed=1234):
    """Rows [r0, r1) of workload `name` on be.device.  Returns dict(X, Y, U, V, Z, sums) where X is a
    DenseMatrix / SparseMatrix of the compute dtype, factors are the non-negative random init of the
    reference (cmf.py:110-117 scaling, (V+V_)/2 merge) and `sums` = (sum X over these rows, sum Y)."""
    from .device import DenseMatrix, SparseMatrix
    torch = be.torch
    c = describe(name, scale)
    n, d, l, k = c["n"], c["d"], c["l"], c["k"]
    dev, dt = be.device, be.tdtype
    g0 = _gen(torch, dev, seed)
    Vt = 0.5 * torch.randn(d, k, generator=g0, device=dev, dtype=torch.float32).abs()
    Zt = 0.5 * torch.randn(l, k, generator=g0, device=dev, dtype=torch.float32)
    if c["sparse"]:
        rates = torch.tensor([0.096, 0.010, 0.053, 0.003, 0.049, 0.009], device=dev)[:l]
        Y = (torch.rand(d, l, generator=g0, device=dev) < rates).to(dt)
    elif c["y_link"] == "logit":
        Y = torch.sigmoid(Vt @ Zt.T).to(dt)
    else:
        Y = torch.randn(d, l, generator=g0, device=dev, dtype=torch.float32).abs().to(dt)
    V0a = torch.randn(d, k, generator=g0, device=dev, dtype=torch.float32).abs()
    V0b = torch.randn(d, k, generator=g0, device=dev, dtype=torch.float32).abs()
    Z0 = torch.randn(l, k, generator=g0, device=dev, dtype=torch.float32)
    if SOLVER_PARAMS[name].get("Z_non_negative", True):
        Z0 = Z0.abs()                                             # U0 / V0 always start non-negative
    if c["sparse"]:
        # Zipf-like column popularity (tf-idf-like skew), fixed permutation
        w = (torch.arange(d, device=dev, dtype=torch.float64) + 10.0) ** -0.9
        perm = torch.randperm(d, generator=g0, device=dev)
        col_cdf = torch.cumsum(w / w.sum(), 0)
    blocks_x, blocks_u, x_sum = [], [], 0.0
    rp_parts, ci_parts, vl_parts, nnz_off = [], [], [], 0
    for b0 in range((r0 // ROW_BLOCK) * ROW_BLOCK, r1, ROW_BLOCK):
        g = _gen(torch, dev, seed + 1 + b0 // ROW_BLOCK)
        rows = min(ROW_BLOCK, n - b0)
        lo, hi = max(r0, b0) - b0, min(r1, b0 + rows) - b0
        U0 = torch.randn(rows, k, generator=g, device=dev, dtype=torch.float32)
        U0 = U0.abs()
        if c["sparse"]:
            per_row = torch.poisson(torch.full((rows,), float(c["nnz_per_row"]), device=dev), generator=g
                                    ).clamp_(min=1).to(torch.int64)
            total = int(per_row.sum())
            u = torch.rand(total, generator=g, device=dev, dtype=torch.float64)
            cols = perm[torch.searchsorted(col_cdf, u).clamp_(max=d - 1)]
            vals = torch.exp(0.5 * torch.randn(total, generator=g, device=dev, dtype=torch.float32))
            row_of = torch.repeat_interleave(torch.arange(rows, device=dev), per_row)
            # sort by (row, col), drop duplicate columns inside a row
            key = row_of * d + cols
            key, order = torch.sort(key)
            keep = torch.ones_like(key, dtype=torch.bool)
            keep[1:] = key[1:] != key[:-1]
            key, vals = key[keep], vals[order][keep]
            row_of, cols = key // d, key % d
            sq = torch.zeros(rows, device=dev, dtype=torch.float32).index_add_(0, row_of, vals * vals)
            vals = vals / torch.sqrt(sq)[row_of]                 # row-L2-normalised, as TfidfVectorizer emits
            sel = (row_of >= lo) & (row_of < hi)
            row_of, cols, vals = row_of[sel] - lo, cols[sel], vals[sel]
            counts = torch.bincount(row_of, minlength=hi - lo)
            rp_parts.append(torch.cumsum(counts, 0) + nnz_off)
            nnz_off += int(counts.sum())
            ci_parts.append(cols.to(torch.int32))
            vl_parts.append(vals.to(dt))
            x_sum += float(vals.sum())
        else:
            if name == "c2":
                Ut = 0.5 * torch.randn(rows, k, generator=g, device=dev, dtype=torch.float32).abs()
                Xb = Ut @ Vt.T + 0.05 * torch.randn(rows, d, generator=g, device=dev, dtype=torch.float32).abs()
            else:
                Xb = torch.randn(rows, d, generator=g, device=dev, dtype=torch.float32).abs()
            Xb = Xb[lo:hi]
            x_sum += float(Xb.sum(dtype=torch.float64))
            blocks_x.append(Xb.to(dt))
        blocks_u.append(U0[lo:hi])
    if c["sparse"]:
        rowptr = torch.cat([torch.zeros(1, device=dev, dtype=torch.int64)] + rp_parts).to(torch.int32)
        colidx, vals = torch.cat(ci_parts), torch.cat(vl_parts)
        order = torch.sort(colidx.to(torch.int64), stable=True).indices
        row_ids = torch.repeat_interleave(torch.arange(r1 - r0, device=dev, dtype=torch.int32),
                                          (rowptr[1:] - rowptr[:-1]).to(torch.int64))
        colptr = torch.zeros(d + 1, device=dev, dtype=torch.int32)
        colptr[1:] = torch.cumsum(torch.bincount(colidx.to(torch.int64), minlength=d), 0).to(torch.int32)
        X = SparseMatrix((r1 - r0, d), rowptr, colidx, vals, colptr, row_ids[order].contiguous(),
                         vals[order].contiguous())
    else:
        X = be.dense(torch.cat(blocks_x, 0))
    return dict(X=X, Y=DenseMatrix(Y.contiguous()), U_raw=torch.cat(blocks_u, 0), V_raw=(V0a, V0b), Z_raw=Z0,
                x_sum=x_sum, y_sum=float(Y.sum(dtype=torch.float64)), shape=(n, d, l, k), config=c)


def finish_init(be, data, x_sum_total):
    """Scale the raw N(0,1) draws like the reference's random init: sqrt(mean / k) (cmf.py:111)."""
    n, d, l, k = data["shape"]
    sx = math.sqrt(abs(x_sum_total / (float(n) * d)) / k)
    sy = math.sqrt(abs(data["y_sum"] / (float(d) * l)) / k)
    dt = be.tdtype
    U = (sx * data["U_raw"]).to(dt).contiguous()
    V = ((sx * data["V_raw"][0] + sy * data["V_raw"][1]) / 2).to(dt).contiguous()
    Z = (sy * data["Z_raw"]).to(dt).contiguous()
    return U, V, Z


def algorithmic_work(name, c, dtype_bytes=4):
    """Per-iteration algorithmic flops / bytes (SURVEY 8d formulas)."""
    n, d, l, k = c["n"], c["d"], c["l"], c["k"]
    s = dtype_bytes
    if c["solver"] == "mu" and not c["sparse"]:
        flops = 4 * n * d * k + 4 * d * l * k + 4 * n * k * k + 4 * d * k * k + 4 * l * k * k
        byts = 2 * n * d * s + 2 * d * l * s + 3 * (n + d + l) * k * s
    elif c["solver"] == "mu":
        nnz = n * c["nnz_per_row"]
        flops = 4 * nnz * k + 4 * n * k * k + 4 * d * k * k + 4 * d * l * k
        byts = 2 * nnz * (s + 4) + 2 * (n + 1) * 4 + 3 * n * k * s + 6 * d * k * s
    else:
        flops = 8 * n * d * k + 4 * d * l * k + 4 * d * l * k * k + 6 * n * k * k + 6 * d * k * k
        byts = 2 * n * d * s + 2 * d * l * s + 3 * (n + d + l) * k * s
    return float(flops), float(byts)
