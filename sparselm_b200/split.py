"""SparseGroupLasso / AdaptiveSparseGroupLasso with ``standardize=True``.

Reference: the group norms become ``||X_g b_g||`` (_lasso.py:249-252) while the l1 term stays on
``b`` (:627-639), so the penalty  sum_j w1_j |b_j| + sum_g w2_g ||X_g b_g||  is separable neither
in ``b`` nor in the per-group whitened variables the other standardized estimators are solved in.
It is solved by the method of multipliers on the split

    s_g = R_g b_g / sqrt(n),      R_g^T R_g = X_g^T X_g  (the Cholesky factor of the Gram's diagonal block)

    min_{b, s}  1/(2n) ||y - X b||^2 + sum_j w1_j |b_j| + sum_g sqrt(n) w2_g ||s_g||
                + rho/2 || R b / sqrt(n) - s + u ||^2

whose inner problem in z = (b, s) has SEPARABLE penalties (l1 on the first p coordinates, group-l2
on the last p) on the Gram of the augmented design [[X, 0], [sqrt(rho) R, -sqrt(n rho) I]]:

    M = [[G + rho R^T R,  -rho sqrt(n) R^T], [-rho sqrt(n) R,  n rho I]],   c' = [c - rho sqrt(n) R^T u ; n rho u].

That is exactly the problem class of the engine (``slm_solve_batch`` with per-coordinate l1 weights
and group weights): every inner solve runs on the engine's kernels, the scaled multiplier ``u`` moves
by the constraint residual between solves.  Only the linear term depends on ``u`` -- the augmented
Gram and its Lipschitz constant are built once per training set -- but it differs from column to
column, so the columns of a grid are solved one after the other here (correctness first: this is
the rarely used corner of the estimator family, not the benchmarked path).

The adaptive variant repeats this with the reference's weight updates (_adaptive_lasso.py:712-726),
its group update fed by the problem's norm expression ``||X_g b_g||`` as in the reference.
"""

from __future__ import annotations

import ctypes

import numpy as np

from .engine import PenaltyGrid

__all__ = ["solve_split"]

RHO = 5.0          # multiplier penalty (in units of the data term's curvature)
MAX_OUTER = 80     # multiplier iterations per problem and pass
FEAS_TOL = 1e-9    # constraint residual ||R b / sqrt(n) - s||_inf (relative to max(1, ||s||_inf)) at which the multipliers stop


def _group_factor(engine, Gf, p, gptr, kscale):
    """Dense block-diagonal R [p, p] (upper-triangular blocks) with R_g^T R_g = kscale * G_gg."""
    torch = engine.torch
    pa = Gf.shape[-1]
    gptr = np.asarray(gptr, dtype=np.int64)
    Gn = len(gptr) - 1
    sizes = np.diff(gptr)
    wptr = np.concatenate([[0], np.cumsum(sizes * sizes)]).astype(np.int64)
    wtot = int(wptr[-1])
    gptr_dev = engine.to_device(gptr.astype(np.int32))
    wptr_dev = engine.to_device(wptr)
    W = torch.empty(max(wtot, 1), dtype=torch.float64, device=engine.device)
    scratch = torch.empty(max(wtot, 1), dtype=torch.float64, device=engine.device)
    info = torch.zeros(1, dtype=torch.int32, device=engine.device)
    engine._ck(engine.lib.slm_group_whiten_factors(engine.h, engine._ptr(Gf), pa, p, engine._ptr(gptr_dev),
                                                   engine._ptr(wptr_dev), Gn, ctypes.c_void_p(0), engine._ptr(W),
                                                   engine._ptr(scratch), engine._ptr(info), engine.stream),
               "slm_group_whiten_factors")
    if int(info.item()):
        raise ValueError("standardize=True: a group's block X_g^T X_g is not positive definite on a training set "
                         "(linearly dependent columns inside a group, or fewer rows than the group has features)")
    rows, cols, src = [], [], []
    for g in range(Gn):
        m, g0 = int(sizes[g]), int(gptr[g])
        i, j = np.triu_indices(m)
        rows.append(g0 + i)
        cols.append(g0 + j)
        src.append(wptr[g] + i * m + j)
    rows, cols, src = (torch.from_numpy(np.concatenate(a)).to(engine.device) for a in (rows, cols, src))
    R = torch.zeros((p, p), dtype=torch.float64, device=engine.device)
    R[rows, cols] = scratch[src] * float(np.sqrt(kscale))
    return R


def solve_split(engine, fd, G, keys, n_obs, fold_specs, tol=1e-10, max_iter=None, check_every=10, floor_rel=1e-14):
    """Same contract as ``Engine.solve`` for specs with ``split=True``: returns B [F, p, ldz] (solver
    feature order) and numpy [F, ldz] gap / primal / n_iter / status / n_pass."""
    torch = engine.torch
    dev = engine.device
    F = G.shape[0]
    p = fd.p
    pa = G.shape[-1]
    Ks = [len(fs) for fs in fold_specs]
    ldz = max(8, (max(Ks) + 7) // 8 * 8)
    s0 = next(fs[0] for fs in fold_specs if len(fs))
    gptr = np.arange(p + 1) if s0.gptr is None else np.asarray(s0.gptr, dtype=np.int64)
    Gn = len(gptr) - 1
    sizes = np.diff(gptr)
    gid = torch.from_numpy(np.repeat(np.arange(Gn), sizes)).to(dev)
    p2 = 2 * p
    pa2 = engine.padded_cols(p2)
    gptr2 = np.concatenate([np.arange(p), p + gptr]).astype(np.int32)  # p singleton groups, then the s groups
    B = torch.zeros((F, p, ldz), dtype=torch.float64, device=dev)
    out = {k: np.zeros((F, ldz)) for k in ("gap", "primal")}
    out.update(n_iter=np.zeros((F, ldz), dtype=np.int32), status=np.zeros((F, ldz), dtype=np.int32),
               n_pass=np.ones((F, ldz), dtype=np.int64))
    total_iters, n_unconverged = 0, 0
    # a multiplier iteration cannot push the constraint residual below the accuracy of its inner solutions
    # (a relative duality gap of eps leaves ~sqrt(eps) in the coefficients): the inner solves run three
    # decades tighter than the caller's tolerance, with a bounded iteration budget
    inner_tol = max(min(float(tol) * 1e-3, 1e-13), 1e-14)
    if max_iter is None:
        max_iter = 200000 if p2 <= 160 else 50000

    for f in range(F):
        if Ks[f] == 0:
            continue
        n = float(n_obs[f])
        sn = np.sqrt(n)
        rows_f = float(fd.n) if keys[f] == "full" else float(fd.n_train[keys[f]])
        kscale = rows_f / n if fd.extra.get("weighted") else 1.0  # weights normalised to sum to the row count
        Gf = G[f]
        R = _group_factor(engine, Gf, p, gptr, kscale)
        RtR = R.t() @ R
        M = torch.zeros((1, pa2, pa2), dtype=torch.float64, device=dev)
        M[0, :p, :p] = Gf[:p, :p] + RHO * RtR
        M[0, :p, p:p2] = -RHO * sn * R.t()
        M[0, p:p2, :p] = -RHO * sn * R
        M[0, p:p2, p:p2] = n * RHO * torch.eye(p, dtype=torch.float64, device=dev)
        cvec, yty = Gf[p, :p].clone(), float(Gf[p, p].item())
        L = engine.lipschitz_device(M, p2).clamp_min(1e-300) * (engine.LIPSCHITZ_MARGIN / n)

        for k, spec in enumerate(fold_specs[f]):
            ad = spec.adaptive
            gw = np.ones(Gn) if spec.gw is None else np.asarray(spec.gw, dtype=float)
            w1 = torch.full((p,), float(spec.lam1), dtype=torch.float64, device=dev)
            w2 = engine.to_device(np.asarray(spec.w2, dtype=float))
            max_pass = 1 if ad is None else int(ad["max_iter"])
            u = torch.zeros(p, dtype=torch.float64, device=dev)
            z = torch.zeros((1, p2, 8), dtype=torch.float64, device=dev)
            beta = torch.zeros(p, dtype=torch.float64, device=dev)
            iters, ok, resid, n_pass = 0, True, 0.0, 0
            for ps in range(max_pass):
                n_pass = ps + 1
                W1 = np.zeros((p2, 1))
                W1[:p, 0] = w1.cpu().numpy()
                W2 = np.zeros((p + Gn, 1))
                W2[p:, 0] = sn * w2.cpu().numpy()
                grid = PenaltyGrid(p=p2, lam1=np.zeros(1), gptr=gptr2, W2=W2)
                best_resid, stalls = float("inf"), 0
                for _outer in range(MAX_OUTER):
                    M[0, p2, :p] = cvec - RHO * sn * (R.t() @ u)
                    M[0, p2, p:p2] = n * RHO * u
                    M[0, :p2, p2] = M[0, p2, :p2]
                    M[0, p2, p2] = yty + n * RHO * float((u @ u).item())
                    res = engine.solve(M, p2, [n], L, [grid], B0=z, tol=inner_tol, floor_rel=floor_rel,
                                       max_iter=max_iter, check_every=check_every, newton=False,
                                       W1_init=engine.to_device(np.ascontiguousarray(np.repeat(W1, 8, axis=1))[None]))
                    z = res["B"]
                    iters += int(res["iters_run"])
                    ok = int(res["status"][0, 0]) == 0  # the last inner solve is the one that counts
                    beta, s = z[0, :p, 0], z[0, p:p2, 0]
                    r = (R @ beta) / sn - s
                    u = u + r
                    resid = float(r.abs().max().item())
                    scale = max(1.0, float(s.abs().max().item()))
                    if resid <= FEAS_TOL * scale:
                        break
                    if _outer >= 8 and resid > 0.7 * best_resid:  # stalled at the accuracy of the inner solves
                        stalls += 1
                        if stalls >= 3:
                            break
                    best_resid = min(best_resid, resid)
                # a group whose split variable is exactly zero is a zero group (R b = sqrt(n) s in the limit)
                snz = torch.zeros(Gn, dtype=torch.float64, device=dev).index_add_(0, gid, s.abs())
                beta = torch.where(snz[gid] > 0, beta, torch.zeros_like(beta))
                if ad is None:
                    break
                # reference updates (_adaptive_lasso.py:712-726): group norms = the problem's ||X_g b_g||
                rb = R @ beta
                norms = torch.sqrt(torch.zeros(Gn, dtype=torch.float64, device=dev).index_add_(0, gid, rb * rb))
                if ad.get("update_function") is not None:
                    fn = ad["update_function"]
                    w1n = float(ad["a1"]) * engine.to_device(np.asarray(fn(beta.cpu().numpy(), ad["eps"]), dtype=float))
                    w2n = engine.to_device(float(ad["a2"]) * gw * np.asarray(fn(norms.cpu().numpy(), ad["eps"]), dtype=float))
                else:
                    w1n = float(ad["a1"]) * (float(ad["alpha"]) / (beta.abs() + float(ad["eps"])))
                    w2n = engine.to_device(float(ad["a2"]) * gw) * (float(ad["alpha"]) / (norms + float(ad["eps"])))
                dn = float(torch.sqrt(((w1n - w1) ** 2).sum() + ((w2n - w2) ** 2).sum()).item())
                w1, w2 = w1n, w2n
                if dn <= float(ad["tol"]):
                    break
            B[f, :, k] = beta
            rb = R @ beta
            gn = torch.sqrt(torch.zeros(Gn, dtype=torch.float64, device=dev).index_add_(0, gid, rb * rb))
            quad = (yty - 2.0 * float((cvec @ beta).item()) + float((beta @ (Gf[:p, :p] @ beta)).item())) / (2.0 * n)
            out["primal"][f, k] = quad + float((w1 * beta.abs()).sum().item()) + float((w2 * gn).sum().item())
            out["gap"][f, k] = resid
            out["n_iter"][f, k] = min(iters, np.iinfo(np.int32).max)
            conv = resid <= 1e-7 * max(1.0, float(beta.abs().max().item()))  # parity level: coefficients to 1e-6
            out["status"][f, k] = 0 if conv else 1
            out["n_pass"][f, k] = n_pass
            total_iters += iters
            n_unconverged += 0 if conv else 1
    out.update(B=B, ldz=ldz, K=Ks, iters_run=total_iters, n_unconverged=n_unconverged, W1=None, W2=None, newton=None)
    return out
