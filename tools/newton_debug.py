"""One slm_newton_step against a float64 torch model, intermediate by intermediate (Hessian, factor,
gradient, direction, step) -- read back from the engine's workspace.  Debug aid."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparselm_b200.engine import Engine  # noqa: E402

eng = Engine(0)
dev = eng.device
rng = np.random.default_rng(4)
n, p, Gn, F, K = 500, int(sys.argv[1]) if len(sys.argv) > 1 else 203, 29, 2, int(sys.argv[2]) if len(sys.argv) > 2 else 4
EMPTY = len(sys.argv) > 3 and sys.argv[3] == 'empty'
sizes = rng.multinomial(p - Gn, np.ones(Gn) / Gn) + 1
gptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
gid = np.repeat(np.arange(Gn), sizes)
pa = eng.padded_cols(p)
Gs = torch.zeros((F, pa, pa), dtype=torch.float64, device=dev)
for f in range(F):
    X = rng.standard_normal((n, p))
    w = np.zeros(p)
    for g in rng.choice(Gn, 6, replace=False):
        w[gptr[g]:gptr[g + 1]] = rng.standard_normal(sizes[g])
    y = X @ w + 0.3 * rng.standard_normal(n)
    Xa = np.hstack([X, y[:, None], np.ones((n, 1)), np.zeros((n, pa - p - 2))])
    Gs[f] = torch.from_numpy(Xa.T @ Xa).to(dev)
fold = np.array([0, 1, 1, 0, 1, 0, 1, 1, 0, 0, 1], dtype=np.int32)[:K]
nobs = torch.full((K,), float(n), dtype=torch.float64, device=dev)
w2 = torch.from_numpy(np.logspace(-1, -2, K)[:, None] * (0.5 + rng.random((K, Gn)))).to(dev)
d2 = torch.from_numpy((rng.random((K, Gn)) < 0.3) * 0.2).to(dev)
X0 = np.zeros((K, p))
for c in range(K):
    for g in rng.choice(Gn, 8, replace=False):
        X0[c, gptr[g]:gptr[g + 1]] = rng.standard_normal(sizes[g])
ldv = (p + 7) // 8 * 8
Xw = torch.zeros((K, ldv), dtype=torch.float64, device=dev)
Xw[:, :p] = torch.from_numpy(X0).to(dev)
nbytes = eng.lib.slm_newton_workspace(p, Gn, K, F)
work = torch.empty(nbytes, dtype=torch.uint8, device=dev) if EMPTY else torch.zeros(nbytes, dtype=torch.uint8, device=dev)
if EMPTY:
    work.view(torch.float64)[:] = float('nan')  # poison: anything read before it is written shows up
out = torch.zeros((K, 4), dtype=torch.float64, device=dev)
gp, gi = torch.from_numpy(gptr).to(dev), torch.from_numpy(gid.astype(np.int32)).to(dev)
x_before = Xw.clone()
eng._ck(eng.lib.slm_newton_step(eng.h, eng._ptr(Gs), pa * pa, pa, p, F, K, fold.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                eng._ptr(nobs), eng._ptr(Xw), eng._ptr(w2), eng._ptr(d2), eng._ptr(gp), eng._ptr(gi), Gn,
                                eng._ptr(work), nbytes, eng._ptr(out), eng.stream), "slm_newton_step")
torch.cuda.synchronize()
wd = work.view(torch.float64) if nbytes % 8 == 0 else work[: nbytes // 8 * 8].view(torch.float64)
ldh = ldv
npan = (p + 63) // 64
o = 0
H = wd[o:o + K * ldh * ldh].view(K, ldh, ldh); o += K * ldh * ldh
INV = wd[o:o + K * npan * 4096].view(K, npan, 64, 64); o += K * npan * 4096
o += 2 * K * 64 * ldh
names = ["U", "KK", "DP", "GS", "GRAD", "DIR"]
vec = {}
for nm in names:
    vec[nm] = wd[o:o + K * ldv].view(K, ldv)[:, :p]; o += K * ldv
print("out", out.cpu().numpy())
for nm in names:
    print(nm, "nan count", int(torch.isnan(vec[nm]).sum()))
print("H nan", int(torch.isnan(H[:, :p, :p]).sum()), "INV nan", int(torch.isnan(INV).sum()))
gid_t = torch.from_numpy(gid.astype(np.int64)).to(dev)
same = (gid_t[:, None] == gid_t[None, :]).to(torch.float64)
for c in range(K):
    x = x_before[c, :p]
    f = int(fold[c])
    nrm_g = torch.sqrt(torch.zeros(Gn, dtype=torch.float64, device=dev).index_add_(0, gid_t, x * x))
    act = (nrm_g > 0)[gid_t].to(torch.float64)
    safe = torch.where(nrm_g > 0, nrm_g, torch.ones_like(nrm_g))[gid_t]
    u = x / safe * act
    kk = w2[c][gid_t] * act / safe
    dp = d2[c][gid_t] * act
    GA = Gs[f, :p, :p] / n
    gs = (GA @ x - Gs[f, p, :p] / n) * act
    grad = gs + kk * x + dp * x
    Hm = GA * act[:, None] * act[None, :] - same * (kk * u)[:, None] * u[None, :]
    Hm = Hm + torch.diag(kk + dp + (1 - act))
    L = torch.linalg.cholesky(Hm)
    d = -torch.cholesky_solve(grad[:, None], L)[:, 0]
    Ldev = torch.tril(H[c, :p, :p])
    print(f"col {c}: grad err {float((vec['GRAD'][c] - grad).abs().max()):.2e} (|grad| {float(grad.abs().max()):.2e})  "
          f"L err {float((Ldev - L).abs().max()):.2e} (|L| {float(L.abs().max()):.2e})  "
          f"dir err {float((vec['DIR'][c] - d).abs().max()):.2e} (|d| {float(d.abs().max()):.2e})  "
          f"x step {float((Xw[c, :p] - x).abs().max()):.2e}")
    # where does L first go wrong (panel granularity)
    e = (Ldev - L).abs()
    for pn in range(npan):
        j0 = pn * 64
        blk = float(e[j0:, j0:j0 + 64].max())
        if blk > 1e-8:
            print(f"    first bad panel {pn}: diag-block err {float(e[j0:j0+64, j0:j0+64].max()):.2e} below err "
                  f"{float(e[j0+64:, j0:j0+64].max()) if j0 + 64 < p else 0:.2e}")
            break
