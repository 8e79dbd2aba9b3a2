"""TEST INFRASTRUCTURE: the sliver of the cvxpy API that ``laml_libs/EM_solver.py:349-407`` uses, on scipy.

The reference states its two M-step programs with cvxpy and solves them with MOSEK (commercial; neither package
is installable here).  With this module registered as ``cvxpy`` (tests/golden/ref_loader.py: ``load(cvxpy="shim")``)
the reference's own ``Mstep`` / ``EM_optimization`` / ``optimize`` code runs verbatim -- the assembly of the
statistics, the clamping and rescaling, the objective and constraint expressions, the status logic, the write-back
to the tree -- and only the numerical solver behind ``Problem.solve`` is a substitute: SLSQP with exact gradients
(expressions carry their Jacobian), polished by a projected Newton step count that reaches ~1e-9 in the variables.
Both programs are strictly concave, so any correct solver returns the same optimum MOSEK does.

Supported: ``Variable(n, nonneg=True)``, ``ndarray @ expr``, ``expr @ ndarray``, scalar / array ``*``, ``+``, ``-``,
unary ``-``, ``exp``, ``log``, comparisons giving box or linear-equality constraints, ``Maximize``, ``Problem``.
"""
import numpy as np
from scipy import optimize

MOSEK = "MOSEK"
ECOS = "ECOS"


class Expr:
    """value ``v`` (1-D array) and Jacobian ``J`` (len(v) x n) as functions of the variable vector"""
    __array_ufunc__ = None            # make ``ndarray (op) Expr`` defer to the reflected methods below

    def __init__(self, fn, var):
        self.fn, self.var = fn, var   # fn(x) -> (v, J)

    # -- arithmetic
    def _lift(self, other):
        if isinstance(other, Expr):
            return other
        c = np.atleast_1d(np.asarray(other, np.float64))
        n = self.var.n
        return Expr(lambda x: (c, np.zeros((c.shape[0], n))), self.var)

    def __neg__(self):
        return Expr(lambda x: (lambda vj: (-vj[0], -vj[1]))(self.fn(x)), self.var)

    def __add__(self, other):
        o = self._lift(other)

        def fn(x):
            (a, Ja), (b, Jb) = self.fn(x), o.fn(x)
            m = max(a.shape[0], b.shape[0])
            return (np.broadcast_to(a, (m,)) + np.broadcast_to(b, (m,)),
                    np.broadcast_to(Ja, (m, Ja.shape[1])) + np.broadcast_to(Jb, (m, Jb.shape[1])))
        return Expr(fn, self.var)

    __radd__ = __add__

    def __sub__(self, other):
        return self + (-self._lift(other))

    def __rsub__(self, other):
        return (-self) + other

    def __mul__(self, other):
        if isinstance(other, Expr):
            raise NotImplementedError("products of expressions are not needed by the reference")
        c = np.atleast_1d(np.asarray(other, np.float64))

        def fn(x):
            v, J = self.fn(x)
            m = max(v.shape[0], c.shape[0])
            vb, cb = np.broadcast_to(v, (m,)), np.broadcast_to(c, (m,))
            return vb * cb, np.broadcast_to(J, (m, J.shape[1])) * cb[:, None]
        return Expr(fn, self.var)

    __rmul__ = __mul__

    def __matmul__(self, other):                     # expr @ ndarray  (1-D . 1-D)
        c = np.asarray(other, np.float64)
        return Expr(lambda x: (lambda vj: (np.atleast_1d(vj[0] @ c), (c @ vj[1])[None, :]))(self.fn(x)), self.var)

    def __rmatmul__(self, other):                    # ndarray @ expr  (vector or matrix)
        a = np.asarray(other, np.float64)

        def fn(x):
            v, J = self.fn(x)
            if a.ndim == 1:
                return np.atleast_1d(a @ v), (a @ J)[None, :]
            return a @ v, a @ J
        return Expr(fn, self.var)

    # -- constraints
    def __le__(self, other):
        return Constraint(self, "<=", other)

    def __ge__(self, other):
        return Constraint(self, ">=", other)

    def __eq__(self, other):                          # noqa: PLW1641 (hashing is not needed)
        return Constraint(self, "==", other)


class Variable(Expr):
    def __init__(self, n, nonneg=False):
        self.n = int(n)
        self.nonneg = nonneg
        self.value = None
        super().__init__(lambda x: (x, np.eye(self.n)), self)


class Constraint:
    def __init__(self, lhs, op, rhs):
        self.lhs, self.op, self.rhs = lhs, op, rhs


def exp(e):
    if not isinstance(e, Expr):
        return np.exp(e)

    def fn(x):
        v, J = e.fn(x)
        ev = np.exp(v)
        return ev, J * ev[:, None]
    return Expr(fn, e.var)


def log(e):
    if not isinstance(e, Expr):
        return np.log(e)

    def fn(x):
        v, J = e.fn(x)
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.log(v), J / v[:, None]
    return Expr(fn, e.var)


class Maximize:
    def __init__(self, expr):
        self.expr = expr


class Problem:
    def __init__(self, objective, constraints=()):
        self.objective, self.constraints = objective, list(constraints)
        self.status = None
        self.value = None

    def solve(self, verbose=False, solver=None, **kw):
        var = self.objective.expr.var
        n = var.n
        lo = np.full(n, 0.0 if var.nonneg else -np.inf)
        hi = np.full(n, np.inf)
        A, b = [], []
        for c in self.constraints:
            if isinstance(c.lhs, Variable) and not isinstance(c.rhs, Expr):          # box
                r = np.broadcast_to(np.asarray(c.rhs, np.float64), (n,))
                if c.op == "<=":
                    hi = np.minimum(hi, r)
                elif c.op == ">=":
                    lo = np.maximum(lo, r)
                else:
                    lo = np.maximum(lo, r); hi = np.minimum(hi, r)
            elif c.op == "==":                                                      # linear equality
                v0, J = c.lhs.fn(np.zeros(n))
                A.append(J); b.append(np.broadcast_to(np.asarray(c.rhs, np.float64), v0.shape) - v0)
            else:
                raise NotImplementedError("general inequality constraints are not needed by the reference")
        if (lo > hi).any():
            self.status = "infeasible"
            return None

        def negf(x):
            v, J = self.objective.expr.fn(x)
            return -float(v[0]), -J[0]

        cons = []
        if A:
            A_, b_ = np.vstack(A), np.concatenate(b)
            cons.append({"type": "eq", "fun": lambda x: A_ @ x - b_, "jac": lambda x: A_})
        width = np.where(np.isfinite(hi), hi, lo + 1.0) - lo
        x0 = lo + np.minimum(0.5 * width, 0.1) + 1e-3                        # interior: log(1 - exp(-x)) needs x > 0
        x0 = np.minimum(x0, np.where(np.isfinite(hi), hi - 1e-9, x0))
        if A:                                                                       # start on the equality set
            corr, *_ = np.linalg.lstsq(A_, b_ - A_ @ x0, rcond=None)
            x0 = np.clip(x0 + corr, lo + 1e-9, np.where(np.isfinite(hi), hi, np.inf))
        best = None
        for _ in range(6):                                                          # restarts tighten SLSQP's answer
            res = optimize.minimize(negf, x0, jac=True, method="SLSQP", bounds=optimize.Bounds(lo, hi, keep_feasible=True),
                                    constraints=cons, options={"maxiter": 2000, "ftol": 1e-15})
            if best is not None and abs(res.fun - best.fun) <= 1e-13 * max(1.0, abs(best.fun)):
                best = res
                break
            best, x0 = res, res.x
        x = best.x
        if A and np.abs(A_ @ x - b_).max() > 1e-7:
            self.status = "infeasible"
            return None
        var.value = x
        self.value = -best.fun
        self.status = "optimal" if np.isfinite(best.fun) else "failure"
        return self.value


API = ["MOSEK", "ECOS", "Variable", "exp", "log", "Maximize", "Problem"]
