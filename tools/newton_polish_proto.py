"""PROTOTYPE (not wired into the engine): second-order polish of slow columns of a batched solve.

Status at the end of round 1 -- measured on CPU on one training fold of C4 (AdaptiveOverlapGroupLasso,
p_ext = 1961, pass 2 of the reweighting), see DESIGN.md section 4 "Known weak spot":
  * the active groups of every column are settled after <= 3000 proximal-gradient iterations, the remaining
    10^4 iterations crawl along the flat directions of the duplicated columns;
  * moderately penalised columns: ONE Newton step on the active manifold takes the relative gap from 7e-10
    to 1e-14;
  * weakly penalised columns (the 20-40k-iteration tail): 7-13 damped Newton steps (one Cholesky of
    n_active <= 1961 each) reach a relative gap of 1e-10; the first ~6 steps are heavily damped
    (t = 0.002 ... 0.125) because groups with a small norm make the norm term strongly non-quadratic;
    a radial-crossing clamp (group -> 0 when the step crosses the origin) keeps the iteration moving from
    unsettled supports but can settle on a wrong support (proximal-gradient iterations must follow);
  * IRLS (majorise ||b_g|| by a quadratic) contracts by only ~3 % per factorisation: useless here;
  * the duality gap of a polished point is dominated by the dual infeasibility of the inactive groups,
    amplified by yty/(2nP) ~ 1e4: the polish must be followed by the engine's own iterations and certificate.
Projected cost on the GPU (3 ms per 2000^2 FP64 Cholesky): ~4 s per C4 search against 12.4 s today -- not
enough to justify an untuned second code path this round; the batched lock-step variant (one factorisation
launch for all slow columns) is the version worth building.


Why: the accelerated proximal-gradient iterations of the engine need ~sqrt(condition number)
iterations.  A group-penalised problem whose penalty is nearly flat on its support (the adaptive
passes weight an active group by alpha^2/||b_g||) on a Gram with exactly flat directions (the
duplicated columns of the overlap expansion, reference _lasso.py:440-461) reaches condition
numbers of 1e6 and more: the support settles after a few hundred iterations and the remaining
10^4 iterations crawl along the flat directions.  On the settled support the objective

    phi(b_A) = 1/(2n) b_A' G_AA b_A - c_A' b_A / n + sum_g w_g ||b_g|| + 1/2 sum_g d_g ||b_g||^2

is smooth (every active group has ||b_g|| > 0), so a damped Newton iteration on the active
coordinates finishes in a handful of steps.  The interior-point solvers behind the reference's
cvxpy call (_base.py:512-519) are second-order methods too.

What this module is: host-side orchestration on device tensors.  The dense factorisation is a
plain library call (``torch.linalg.cholesky_ex`` = cuSOLVER potrf on the GPU); everything is
written with torch operations, so the same code runs on CPU tensors in the tests.  It never
decides convergence: the polished point goes back into the engine's batch and the engine's own
duality-gap certificate (exact ``G b``, all groups, ``gap_final_kernel``) judges it.  A polish on a
wrong support only costs time -- every accepted step decreases the objective (Armijo test on the
*difference* of objective values, formed without cancellation).

Scope: pure group penalties (no l1 term) with optional ridge: GroupLasso, OverlapGroupLasso,
RidgedGroupLasso and their adaptive variants (also in whitened variables, standardize=True).
"""

from __future__ import annotations

import torch

__all__ = ["polish_column", "active_groups"]


def active_groups(b, gid, n_groups):
    """Boolean [n_groups]: groups with a non-zero coefficient."""
    nz = torch.zeros(n_groups, dtype=b.dtype, device=b.device)
    nz.index_add_(0, gid, (b != 0).to(b.dtype))
    return nz > 0


def _group_sums(v, g_of, n_groups):
    out = torch.zeros(n_groups, dtype=v.dtype, device=v.device)
    out.index_add_(0, g_of, v)
    return out


def polish_column(G, c, n, b, w2, d2, gid, scale, tol, max_factor=4, max_steps=12):
    """Damped Newton / chord iteration on the active groups of one column.

    G [p, p] (view of the Gram, any row stride), c [p], n = rows of the data term, b [p] start
    point, w2 [Gn] group weights, d2 [Gn] ridge weights or None, gid [p] int64 group of every
    feature.  ``scale`` = max(|P|, floor) of the engine's relative test and ``tol`` its tolerance:
    the iteration stops once the Newton decrement says phi - phi* <= 1e-3 * tol * scale.

    Returns (b_new [p], info) -- info: steps, factorizations, decrement, ok (False: the line search
    collapsed or the Hessian was not positive definite: the support is not settled; b_new still
    has an objective <= that of b).
    """
    n_groups = w2.shape[0]
    act = active_groups(b, gid, n_groups)[gid]
    idx = torch.nonzero(act).squeeze(1)
    m = int(idx.numel())
    info = {"steps": 0, "factorizations": 0, "decrement": float("nan"), "ok": False, "n_active": m}
    if m == 0:
        info["ok"] = True
        return b, info
    GA = G.index_select(0, idx).index_select(1, idx) / n
    cA = c.index_select(0, idx) / n
    g_of = gid.index_select(0, idx)
    w = w2.index_select(0, g_of)
    d = torch.zeros_like(w) if d2 is None else d2.index_select(0, g_of)
    same = g_of[:, None] == g_of[None, :]
    x = b.index_select(0, idx).clone()
    target = 1e-3 * tol * scale
    chol = None
    last_dec = None
    small_t = 0
    while info["steps"] < max_steps:
        nrm_g = torch.sqrt(_group_sums(x * x, g_of, n_groups))
        nrm = nrm_g.index_select(0, g_of)
        u = x / nrm
        Gx = GA @ x
        gs = Gx - cA                     # gradient of the quadratic part
        grad = gs + w * u + d * x
        if chol is None:
            if info["factorizations"] >= max_factor:
                break
            k = w / nrm
            H = GA - torch.where(same, (k * u)[:, None] * u[None, :], torch.zeros((), dtype=x.dtype, device=x.device))
            H.diagonal().add_(k + d)
            chol, err = torch.linalg.cholesky_ex(H)
            info["factorizations"] += 1
            fresh = True
            if int(err) != 0:
                break
        else:
            fresh = False
        dlt = -torch.cholesky_solve(grad[:, None], chol)[:, 0]
        gd = float(grad @ dlt)
        dec = -gd                        # Newton decrement^2 (chord: an estimate of it)
        info["decrement"] = dec
        if not dec > 0.0:
            info["ok"] = dec == 0.0
            break
        if (not fresh) and last_dec is not None and dec > 0.25 * last_dec:
            chol = None                  # the chord iteration stalls: refactor at this point
            last_dec = None
            continue
        # Armijo backtracking on phi(x + t dlt) - phi(x), every term formed as a difference
        Gd = GA @ dlt
        q1, q2 = float(gs @ dlt), float(dlt @ Gd)
        xd = x * dlt
        dd = dlt * dlt
        t = 1.0
        accepted = False
        for _ in range(40):
            dsq = _group_sums(2.0 * t * xd + (t * t) * dd, g_of, n_groups)   # ||x_g + t d_g||^2 - ||x_g||^2
            new_sq = nrm_g * nrm_g + dsq
            if bool((new_sq[nrm_g > 0] <= 0).any()):
                t *= 0.5
                continue
            new_nrm = torch.sqrt(torch.clamp(new_sq, min=0.0))
            wg = _group_sums(w, g_of, n_groups) / torch.clamp(_group_sums(torch.ones_like(w), g_of, n_groups), min=1.0)
            dpen = float((wg * dsq / (new_nrm + nrm_g + (nrm_g == 0))).sum())
            dridge = float((d * (t * xd + 0.5 * t * t * dd)).sum())
            dphi = t * q1 + 0.5 * t * t * q2 + dridge + dpen
            if dphi <= 1e-4 * t * gd:
                accepted = True
                break
            t *= 0.5
        if not accepted:
            break
        x = x + t * dlt
        info["steps"] += 1
        if t < 1.0:
            chol = None                  # far from the quadratic regime: new Hessian next step
            last_dec = None
            small_t = small_t + 1 if t < 1e-2 else 0
            if small_t >= 2:
                break
        else:
            last_dec = dec
            if 0.5 * dec <= target:
                info["ok"] = True
                break
    out = b.clone()
    out.index_copy_(0, idx, x)
    return out, info
