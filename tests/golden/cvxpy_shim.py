"""A numeric stand-in for the handful of cvxpy names the reference's convex estimators use.

TEST INFRASTRUCTURE (fixture generation only).  cvxpy and its conic solvers are not installed in
this image, so the reference's `problem.solve()` cannot run.  Everything AROUND that call is plain
Python that builds an expression tree: `CVXRegressor.generate_problem` -> `_generate_params` ->
`_generate_auxiliaries` -> `_generate_objective` (reference src/sparselm/model/_base.py:414-467),
the overlap expansion (_lasso.py:440-484), the adaptive re-weighting loop (_adaptive_lasso.py:
206-232) and its updates (:196-204, :364-374, :712-726).  This module provides just enough of
cvxpy's surface for that code to run UNMODIFIED:

    Variable, Parameter, Constant-folding arithmetic (+ - * @ unary -, indexing),
    norm1, norm2, sum_squares, multiply, hstack, Minimize, Problem(.solve/.value/.variables),
    installed_solvers, and the names used only in annotations (Expression, Constraint).

Every node evaluates numerically (`.value`) from the current values of its leaves, so the
reference's own objective expression can be evaluated at any beta, and the reference's own
`group_norms.value` / `adaptive_weights.value` drive its re-weighting loop.  `Problem.solve`
delegates to a hook (set by the fixture generator) that must put a minimiser into the Variable;
`analyze(objective)` turns the reference-built expression into a standard form

    sum_q c_q ||A_q b - r_q||^2  +  sum_r || w_r o (A_r b - r_r) ||_1  +  sum_i v_i ||A_i b - r_i||_2

by probing the affine sub-expressions, which the generator uses for a KKT certificate that is
independent of oracle/ (tests/golden/make_golden_reference.py).

No DCP analysis, no canonicalisation, no solver: this is not cvxpy.
"""

from __future__ import annotations

import numpy as np

__all__ = ["Variable", "Parameter", "Constant", "Expression", "Constraint", "Minimize", "Problem", "norm1", "norm2",
           "sum_squares", "multiply", "hstack", "installed_solvers", "analyze", "SOLVE_HOOK"]

SOLVE_HOOK = None  # callable(problem, **solve_kwargs) -> optimal value; must set the Variables


def installed_solvers():
    return ["SHIM"]


def _wrap(x):
    return x if isinstance(x, Expression) else Constant(x)


class Expression:
    __array_ufunc__ = None  # numpy defers to the reflected operators below (X @ beta, w @ norms, ...)
    __array_priority__ = 100

    # -- numeric evaluation ------------------------------------------------
    @property
    def value(self):
        return self._eval()

    def _eval(self):  # pragma: no cover
        raise NotImplementedError

    def children(self):
        return ()

    @property
    def shape(self):
        return np.shape(self.value)

    # -- arithmetic -----------------------------------------------------------
    def __add__(self, o):
        return BinOp("add", self, _wrap(o))

    def __radd__(self, o):
        return BinOp("add", _wrap(o), self)

    def __sub__(self, o):
        return BinOp("sub", self, _wrap(o))

    def __rsub__(self, o):
        return BinOp("sub", _wrap(o), self)

    def __mul__(self, o):
        return BinOp("mul", self, _wrap(o))

    def __rmul__(self, o):
        return BinOp("mul", _wrap(o), self)

    def __truediv__(self, o):
        return BinOp("div", self, _wrap(o))

    def __matmul__(self, o):
        return BinOp("matmul", self, _wrap(o))

    def __rmatmul__(self, o):
        return BinOp("matmul", _wrap(o), self)

    def __neg__(self):
        return BinOp("mul", Constant(-1.0), self)

    def __pow__(self, e):
        return BinOp("pow", self, _wrap(e))

    def __getitem__(self, key):
        return Index(self, key)

    def __len__(self):
        return self.shape[0]


class Constant(Expression):
    def __init__(self, value):
        self._value = np.asarray(value, dtype=float) if not np.isscalar(value) else float(value)

    def _eval(self):
        return self._value


class _Leaf(Expression):
    def __init__(self, shape=(), value=None, name=None):
        if isinstance(shape, (int, np.integer)):
            shape = (int(shape),)
        self._shape = tuple(int(s) for s in shape)
        self._value = None
        self.name = name
        if value is not None:
            self.value = value

    @property
    def shape(self):
        return self._shape

    def _check(self, v):
        return v

    @property
    def value(self):
        return self._value

    @value.setter
    def value(self, v):
        if v is None:
            self._value = None
            return
        v = np.asarray(v, dtype=float)
        if v.shape != self._shape:
            if v.size == int(np.prod(self._shape, dtype=int)) and v.ndim <= 1 and len(self._shape) <= 1:
                v = v.reshape(self._shape)
            else:
                raise ValueError(f"Invalid dimensions {v.shape} for value of shape {self._shape}")
        self._value = self._check(v if v.shape else float(v))

    def _eval(self):
        if self._value is None:
            raise ValueError("leaf has no value")
        return self._value


class Variable(_Leaf):
    def __init__(self, shape=(), name=None, **attrs):
        self.attributes = attrs
        super().__init__(shape, None, name)


class Parameter(_Leaf):
    """cvxpy.Parameter: value with optional sign/type attributes that are enforced on assignment
    (cvxpy raises ValueError when a value violates them)."""

    def __init__(self, shape=(), name=None, value=None, **attrs):
        self.attributes = attrs
        super().__init__(shape, value, name)

    def _check(self, v):
        a = self.attributes
        arr = np.asarray(v)
        tol = 1e-12
        if a.get("nonneg") and np.any(arr < -tol):
            raise ValueError("Parameter value must be nonnegative.")
        if a.get("pos") and np.any(arr <= 0):
            raise ValueError("Parameter value must be positive.")
        if a.get("nonpos") and np.any(arr > tol):
            raise ValueError("Parameter value must be nonpositive.")
        if a.get("neg") and np.any(arr >= 0):
            raise ValueError("Parameter value must be negative.")
        if a.get("integer") and np.any(arr != np.round(arr)):
            raise ValueError("Parameter value must be integer.")
        if a.get("boolean") and np.any((arr != 0) & (arr != 1)):
            raise ValueError("Parameter value must be boolean.")
        return v


class BinOp(Expression):
    def __init__(self, op, a, b):
        self.op, self.a, self.b = op, a, b

    def children(self):
        return (self.a, self.b)

    def _eval(self):
        a, b = self.a.value, self.b.value
        if self.op == "add":
            return a + b
        if self.op == "sub":
            return a - b
        if self.op == "mul":
            return a * b
        if self.op == "div":
            return a / b
        if self.op == "pow":
            return a ** b
        if self.op == "matmul":
            return np.asarray(a) @ np.asarray(b)
        raise ValueError(self.op)


class Index(Expression):
    def __init__(self, a, key):
        self.a, self.key = a, key

    def children(self):
        return (self.a,)

    def _eval(self):
        return np.asarray(self.a.value)[self.key]


class Atom(Expression):
    def __init__(self, kind, args):
        self.kind, self.args = kind, [_wrap(a) for a in args]

    def children(self):
        return tuple(self.args)

    def _eval(self):
        v = [a.value for a in self.args]
        if self.kind == "norm1":
            return float(np.sum(np.abs(v[0])))
        if self.kind == "norm2":
            return float(np.sqrt(np.sum(np.square(v[0]))))
        if self.kind == "sum_squares":
            return float(np.sum(np.square(v[0])))
        if self.kind == "multiply":
            return np.asarray(v[0]) * np.asarray(v[1])
        if self.kind == "hstack":
            return np.hstack([np.atleast_1d(x) for x in v])
        raise ValueError(self.kind)


def norm1(x):
    return Atom("norm1", [x])


def norm2(x):
    return Atom("norm2", [x])


def sum_squares(x):
    return Atom("sum_squares", [x])


def multiply(a, b):
    return Atom("multiply", [a, b])


def hstack(xs):
    return Atom("hstack", list(xs))


class Constraint:  # annotations only: the convex estimators generate no constraints
    pass


class Minimize:
    def __init__(self, expr):
        self.expr = _wrap(expr)

    @property
    def value(self):
        return self.expr.value


def _walk(e, seen, out):
    if id(e) in seen:
        return
    seen.add(id(e))
    if isinstance(e, Variable):
        out.append(e)
    for c in e.children():
        _walk(c, seen, out)


class Problem:
    def __init__(self, objective, constraints=None):
        self.objective = objective
        self.constraints = list(constraints or [])
        self.value = None
        self.status = None
        self.n_solves = 0

    def variables(self):
        out = []
        _walk(self.objective.expr, set(), out)
        return out

    def solve(self, solver=None, warm_start=False, **kwargs):
        if self.constraints:
            raise NotImplementedError("cvxpy shim: constrained problems are outside the convex estimators' path")
        if SOLVE_HOOK is None:
            raise RuntimeError("cvxpy shim: no SOLVE_HOOK installed (cvxpy itself is not available in this image)")
        self.value = float(SOLVE_HOOK(self, solver=solver, warm_start=warm_start, **kwargs))
        self.status = "optimal"
        self.n_solves += 1
        return self.value

    def __repr__(self):
        return "Problem(shim)"


# --------------------------------------------------------------------------- #
# standard form of an objective built from the atoms above
# --------------------------------------------------------------------------- #
def _depends(e, var):
    out = []
    _walk(e, set(), out)
    return any(v is var for v in out)


def _affine(e, var):
    """(A, r) with e(b) = A b - r for an expression that is affine in `var` (probed numerically:
    the subtree is evaluated at 0 and at the unit vectors; exact for affine maps)."""
    p = var.shape[0]
    keep = var.value
    try:
        var.value = np.zeros(p)
        e0 = np.atleast_1d(np.asarray(e.value, dtype=float)).copy()
        A = np.zeros((e0.size, p))
        for j in range(p):
            z = np.zeros(p)
            z[j] = 1.0
            var.value = z
            A[:, j] = np.atleast_1d(np.asarray(e.value, dtype=float)) - e0
        # affinity check at a random point
        z = np.random.default_rng(0).standard_normal(p)
        var.value = z
        chk = np.atleast_1d(np.asarray(e.value, dtype=float))
        if not np.allclose(chk, A @ z + e0, rtol=1e-10, atol=1e-10 * (1 + np.abs(chk).max())):
            raise ValueError("sub-expression is not affine in the variable")
    finally:
        var._value = keep
    return A, -e0


def analyze(expr, var):
    """Standard form of a scalar objective expression: dict(quad=[(c, A, r)], l1=[(w, A, r)],
    l2=[(v, A, r)]) meaning sum c||Ab-r||^2 + sum ||w o (Ab-r)||_1 + sum v||Ab-r||_2, plus
    const.  Weights are read from the CURRENT parameter values."""
    out = dict(quad=[], l1=[], l2=[], const=0.0)

    def scalar_terms(e, scale):
        """e is a scalar expression; add scale * e to `out`."""
        if not _depends(e, var):
            out["const"] += scale * float(e.value)
            return
        if isinstance(e, BinOp):
            if e.op == "add":
                scalar_terms(e.a, scale)
                scalar_terms(e.b, scale)
                return
            if e.op == "sub":
                scalar_terms(e.a, scale)
                scalar_terms(e.b, -scale)
                return
            if e.op == "mul":
                if not _depends(e.a, var):
                    scalar_terms(e.b, scale * float(e.a.value))
                    return
                if not _depends(e.b, var):
                    scalar_terms(e.a, scale * float(e.b.value))
                    return
            if e.op == "div" and not _depends(e.b, var):
                scalar_terms(e.a, scale / float(e.b.value))
                return
            if e.op == "matmul":
                # weights @ vector-of-atoms (either side may hold the weights)
                wexpr, vexpr = (e.a, e.b) if not _depends(e.a, var) else (e.b, e.a)
                if _depends(wexpr, var):
                    raise ValueError("bilinear term")
                w = np.atleast_1d(np.asarray(wexpr.value, dtype=float))
                atoms = vector_atoms(vexpr)
                if len(atoms) != len(w):
                    raise ValueError("weights / atoms length mismatch")
                for wi, (kind, sub, sc) in zip(w, atoms):
                    add_atom(kind, sub, scale * wi * sc)
                return
        if isinstance(e, Atom) and e.kind in ("norm1", "norm2", "sum_squares"):
            add_atom(e.kind, e.args[0], scale)
            return
        raise ValueError(f"cvxpy shim: cannot analyse scalar node {type(e).__name__}:{getattr(e, 'op', getattr(e, 'kind', ''))}")

    def vector_atoms(e):
        """e is a vector whose entries are scalar atoms: list of (kind, affine sub-expression, scale)."""
        if isinstance(e, Atom) and e.kind == "hstack":
            res = []
            for a in e.args:
                res += vector_atoms(a)
            return res
        if isinstance(e, Atom) and e.kind in ("norm1", "norm2", "sum_squares"):
            return [(e.kind, e.args[0], 1.0)]
        if isinstance(e, BinOp) and e.op == "mul":
            if not _depends(e.a, var):
                return [(k, s, sc * float(e.a.value)) for k, s, sc in vector_atoms(e.b)]
            if not _depends(e.b, var):
                return [(k, s, sc * float(e.b.value)) for k, s, sc in vector_atoms(e.a)]
        raise ValueError(f"cvxpy shim: cannot analyse vector node {type(e).__name__}")

    def add_atom(kind, sub, scale):
        if scale < 0:
            raise ValueError("negative weight on a convex atom")
        if kind == "norm1" and isinstance(sub, Atom) and sub.kind == "multiply":
            a0, a1 = sub.args
            wexpr, xexpr = (a0, a1) if not _depends(a0, var) else (a1, a0)
            A, r = _affine(xexpr, var)
            w = np.broadcast_to(np.asarray(wexpr.value, dtype=float), (A.shape[0],)) * scale
            out["l1"].append((np.array(w), A, r))
            return
        A, r = _affine(sub, var)
        if kind == "norm1":
            out["l1"].append((np.full(A.shape[0], scale), A, r))
        elif kind == "norm2":
            out["l2"].append((scale, A, r))
        else:
            out["quad"].append((scale, A, r))

    scalar_terms(_wrap(expr), 1.0)
    return out
