"""Model set-up: ``.mod`` text -> index bookkeeping + symbolic first-order model -> generated
CUDA / C model functions.

This replaces, for the likelihood path only, what the reference obtains from its external
preprocessor and loads as ``RuntimeGeneratedFunction``s (reference
``src/dynare_functions.jl:10-47, 145-157``; ``src/model.jl:120-139``) and the index sets built
by ``Model`` (``src/dynare_containers.jl:258-390, 557-602``).  It runs once per model on the
host; its output is compiled into ``libdynare_b200.so`` (the same idea as Dynare's ``use_dll``).

Conventions reproduced:
  * variable order = declaration order, then auxiliary variables for leads/lags beyond one period;
  * dynamic Jacobian column blocks ``[y(-1) | y | y(+1) | u]`` evaluated at the steady state
    with ``u = 0`` (``src/perturbations.jl:526-531``);
  * compressed Jacobian = columns with a positive lead/lag incidence, block by block, then the
    shock columns (``src/perturbations.jl:616-635``);
  * ``i_bkwrd_b`` (has a lag), ``i_fwrd_b`` (has a lead), ``i_static`` (current only);
  * Kalman state = ``sort(union(obs_idx, i_bkwrd_b))`` (``src/estimation/estimation.jl:384-390``).
"""
from __future__ import annotations

import json
import keyword
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import sympy as sp

from .modfile import ModFile, EstimatedParam, parse_mod

# estimated-parameter type codes shared with the C ABI (include/dynare_b200.h)
EST_PARAM, EST_SD_SHOCK, EST_VAR_SHOCK, EST_CORR_SHOCK, EST_SD_MEAS, EST_VAR_MEAS, EST_CORR_MEAS = range(7)

_FUNCS = {"exp": sp.exp, "log": sp.log, "ln": sp.log, "sqrt": sp.sqrt, "abs": sp.Abs,
          "sin": sp.sin, "cos": sp.cos, "tan": sp.tan}


def _tname(v: str, shift: int) -> str:
    if shift == 0:
        return v
    return f"{v}__{'p' if shift > 0 else 'm'}{abs(shift)}"


def _time_subst(expr: str, names: List[str]) -> str:
    """``x(+1)`` -> ``x__p1`` for every declared variable name."""
    if not names:
        return expr
    alt = "|".join(sorted(map(re.escape, names), key=len, reverse=True))
    pat = re.compile(rf"\b({alt})\s*\(\s*([+-]?\s*\d+)\s*\)")
    return pat.sub(lambda m: _tname(m.group(1), int(m.group(2).replace(" ", ""))), expr)


@dataclass
class ModelSpec:
    name: str
    endo: List[str]
    n_orig: int
    exo: List[str]
    params: List[str]
    param_values: np.ndarray
    linear: bool
    lli: np.ndarray                      # 3 x n bool: lag / current / lead incidence
    i_bkwrd_b: List[int]
    i_fwrd_b: List[int]
    i_both: List[int]
    i_static: List[int]
    i_dyn: List[int]                     # non-static, declaration order
    sigma_e0: np.ndarray
    sigma_m0: np.ndarray
    varobs: List[str]
    obs_idx: List[int]
    state_ids: List[int]
    obs_idx_state: List[int]
    lagged_state_ids: List[int]
    est: List[EstimatedParam]
    est_type: List[int]
    est_i: List[int]
    est_j: List[int]
    estimation_options: Dict[str, str]
    # symbolic content
    psym: List[sp.Symbol] = field(default_factory=list)
    ysym: List[sp.Symbol] = field(default_factory=list)
    ss_assign: List[Tuple[sp.Symbol, sp.Expr]] = field(default_factory=list)
    static_resid: List[sp.Expr] = field(default_factory=list)
    jac: Dict[Tuple[int, int], sp.Expr] = field(default_factory=dict)   # (row, compressed col) -> expr
    # numerical steady state (no steady_state_model block): Newton from `initval` on the static model
    ss_numeric: bool = False
    static_jac: Dict[Tuple[int, int], sp.Expr] = field(default_factory=dict)   # (row, variable) -> d static_resid / d y
    # the dynamic residuals before the steady state is substituted, with their lag / lead / shock symbols
    # (kept for tooling that writes the model out again, e.g. scripts/write_preprocessor_sample.py)
    dyn_resid: List[sp.Expr] = field(default_factory=list)
    dyn_syms: Dict[str, List[sp.Symbol]] = field(default_factory=dict)

    # sizes ------------------------------------------------------------------------------
    @property
    def n(self): return len(self.endo)
    @property
    def nx(self): return len(self.exo)
    @property
    def k(self): return len(self.i_bkwrd_b)
    @property
    def nf(self): return len(self.i_fwrd_b)
    @property
    def s(self): return len(self.i_static)
    @property
    def m(self): return self.n - self.s
    @property
    def ncol(self): return self.k + self.n + self.nf + self.nx
    @property
    def ns(self): return len(self.state_ids)
    @property
    def ny(self): return len(self.varobs)
    @property
    def nl(self): return len(self.lagged_state_ids)
    @property
    def np_est(self): return len(self.est)

    def theta_init(self) -> np.ndarray:
        """Initial / calibrated value of every estimated parameter, in ``estimated_params`` order."""
        out = []
        for e, t, i, j in zip(self.est, self.est_type, self.est_i, self.est_j):
            if e.init is not None:
                out.append(e.init)
            elif t == EST_PARAM:
                out.append(self.param_values[i])
            elif t == EST_SD_SHOCK:
                out.append(np.sqrt(self.sigma_e0[i, i]))
            elif t == EST_SD_MEAS:
                out.append(np.sqrt(self.sigma_m0[i, i]))
            else:
                out.append(e.mean)
        return np.array(out, dtype=np.float64)

    def describe(self) -> dict:
        """Plain-data description (what ``dyb_model_desc`` carries; also stored as JSON)."""
        return dict(
            name=self.name, n=self.n, n_orig=self.n_orig, nx=self.nx, npar=len(self.params),
            k=self.k, nf=self.nf, s=self.s, m=self.m, ncol=self.ncol, ns=self.ns, ny=self.ny,
            nl=self.nl, np_est=self.np_est, linear=self.linear,
            steady_state="numeric" if self.ss_numeric else "closed_form",
            endo=self.endo, exo=self.exo, params=self.params,
            param_values=[float(v) for v in self.param_values],
            lead_lag_incidence=self.lli.astype(int).tolist(),
            i_bkwrd_b=self.i_bkwrd_b, i_fwrd_b=self.i_fwrd_b, i_both=self.i_both,
            i_static=self.i_static, i_dyn=self.i_dyn,
            sigma_e0=self.sigma_e0.tolist(), sigma_m0=self.sigma_m0.tolist(),
            varobs=self.varobs, obs_idx=self.obs_idx, state_ids=self.state_ids,
            obs_idx_state=self.obs_idx_state, lagged_state_ids=self.lagged_state_ids,
            est_type=self.est_type, est_i=self.est_i, est_j=self.est_j,
            est_names=[(e.name1 if e.name2 is None else f"{e.name1}:{e.name2}") for e in self.est],
            priors=[dict(kind=e.kind, shape=e.shape, mean=e.mean,
                         stdev=(e.stdev if np.isfinite(e.stdev) else "inf"),
                         **({"domain": list(e.domain)} if e.domain is not None else {})) for e in self.est],
            theta_init=[float(v) for v in self.theta_init()],
            estimation_options=self.estimation_options,
        )


def build_spec(name: str, mf: ModFile) -> ModelSpec:
    endo = list(mf.endo)
    exo = list(mf.exo)
    n_orig = len(endo)
    names = endo + exo

    eq_txt = [_time_subst(f"({lhs}) - ({rhs})", names) for lhs, rhs in mf.equations]

    # --- auxiliary variables for leads / lags beyond one period (encounter order) ------------
    aux_src: Dict[str, Tuple[str, int]] = {}     # aux name -> (source name, +1/-1)
    aux_defs: List[str] = []
    tok = re.compile(r"\b(\w+?)__([pm])(\d+)\b")
    for _ in range(8):
        changed = False
        new_eqs = []
        for e in eq_txt:
            def repl(m):
                nonlocal changed
                v, d, h = m.group(1), m.group(2), int(m.group(3))
                if v in exo and h >= 1:
                    raise NotImplementedError(f"lead / lag on the exogenous variable {v!r} (needs an auxiliary endogenous "
                                              "variable, which the reference's preprocessor creates)")
                if h < 2 or v not in endo:
                    return m.group(0)
                sgn = 1 if d == "p" else -1
                kind = "LEAD" if sgn > 0 else "LAG"
                cur = v
                for j in range(1, h):
                    aux = f"AUX_ENDO_{kind}_{v}_{j}"
                    if aux not in aux_src:
                        aux_src[aux] = (cur, sgn)
                        endo.append(aux)
                        aux_defs.append(f"({aux}) - ({_tname(cur, sgn)})")
                    cur = aux
                changed = True
                return _tname(cur, sgn)
            new_eqs.append(tok.sub(repl, e))
        eq_txt = new_eqs
        if not changed:
            break
    eq_txt += aux_defs
    n = len(endo)
    nx = len(exo)
    if len(eq_txt) != n:
        raise ValueError(f"{name}: {len(eq_txt)} equations for {n} endogenous variables")

    # --- symbols ---------------------------------------------------------------------------
    psym = [sp.Symbol(p, real=True) for p in mf.params]
    ysym = [sp.Symbol(v, real=True) for v in endo]
    ym = [sp.Symbol(_tname(v, -1), real=True) for v in endo]
    yp = [sp.Symbol(_tname(v, +1), real=True) for v in endo]
    usym = [sp.Symbol(u, real=True) for u in exo]
    local = dict(_FUNCS)
    for s_ in psym + ysym + ym + yp + usym:
        local[s_.name] = s_

    # names that are Python keywords (fs2000 has a parameter called ``del``) cannot go through
    # sympify as they are; they are spelled ``<name>__kw`` in the text handed to the parser
    kw = [s_ for s_ in local if keyword.iskeyword(s_)]
    for s_ in kw:
        local[s_ + "__kw"] = local[s_]
    kw_pat = re.compile(r"\b(" + "|".join(map(re.escape, kw)) + r")\b") if kw else None

    def parse(txt: str, extra=None) -> sp.Expr:
        d = dict(local)
        if extra:
            d.update(extra)
        if kw_pat is not None:
            txt = kw_pat.sub(lambda m: m.group(1) + "__kw", txt)
        return sp.sympify(txt.replace("^", "**"), locals=d)

    # model-local variables (`# name = expr;`): substituted into the equations, in file order
    for lname, lexpr in mf.model_locals:
        local[lname] = parse(_time_subst(lexpr, names))
    resid = [parse(e) for e in eq_txt]
    for r in resid:
        bad = [s_ for s_ in r.free_symbols if s_.name not in local]
        if bad:
            raise ValueError(f"{name}: unknown symbols {bad}")

    # --- calibration -----------------------------------------------------------------------
    pval: Dict[sp.Symbol, float] = {}
    for pname, expr in mf.calibration:
        pval[local[pname]] = float(parse(expr).subs(pval))
    param_values = np.array([pval.get(p, np.nan) for p in psym], dtype=np.float64)

    # --- incidence and index sets ------------------------------------------------------------
    lli = np.zeros((3, n), dtype=bool)
    for r in resid:
        fs = r.free_symbols
        for j in range(n):
            lli[0, j] |= ym[j] in fs
            lli[1, j] |= ysym[j] in fs
            lli[2, j] |= yp[j] in fs
    i_bkwrd_b = [j for j in range(n) if lli[0, j]]
    i_fwrd_b = [j for j in range(n) if lli[2, j]]
    i_both = [j for j in range(n) if lli[0, j] and lli[2, j]]
    i_static = [j for j in range(n) if not lli[0, j] and not lli[2, j]]
    i_dyn = [j for j in range(n) if lli[0, j] or lli[2, j]]
    if not lli[1].all():
        raise NotImplementedError(f"{name}: a variable never appears at the current period")

    # --- steady state ----------------------------------------------------------------------
    ss_assign: List[Tuple[sp.Symbol, sp.Expr]] = []
    assigned = set()
    temps: Dict[str, sp.Symbol] = {}
    for lhs, rhs in mf.steady_state_model:
        if lhs in endo:
            target = local[lhs]
        else:
            target = temps.setdefault(lhs, sp.Symbol(f"ss_{lhs}", real=True))
        ss_assign.append((target, parse(rhs, temps)))
        assigned.add(lhs)
    # without a steady_state_model block a non-linear model gets its steady state from Newton's method on the static
    # model, started at the `initval` block (the reference: `steady` -> solve_steady_state!,
    # src/steady_state/SteadyState.jl:183-252); ss_assign then holds that starting point
    ss_numeric = not mf.linear and not mf.steady_state_model
    if ss_numeric:
        for lhs, rhs in mf.initval:
            if lhs in endo:
                ss_assign.append((local[lhs], parse(rhs)))
                assigned.add(lhs)
    for v in endo[:n_orig]:
        if v not in assigned:
            if mf.linear or not mf.steady_state_model:
                ss_assign.append((local[v], sp.Float(0.0)))
            else:
                raise ValueError(f"{name}: steady_state_model does not define {v}")
    for aux in endo[n_orig:]:
        ss_assign.append((local[aux], local[aux_src[aux][0]]))

    # --- static residuals and Jacobian at the steady state -----------------------------------
    to_ss = {}
    for j in range(n):
        to_ss[ym[j]] = ysym[j]
        to_ss[yp[j]] = ysym[j]
    for u in usym:
        to_ss[u] = sp.Integer(0)
    static_resid = [r.subs(to_ss) for r in resid]
    static_jac: Dict[Tuple[int, int], sp.Expr] = {}
    if ss_numeric:
        for i, r in enumerate(static_resid):
            fs = r.free_symbols
            for j, sym in enumerate(ysym):
                if sym in fs:
                    dd = sp.diff(r, sym)
                    if dd != 0:
                        static_jac[(i, j)] = dd

    k, nf = len(i_bkwrd_b), len(i_fwrd_b)
    cols: List[sp.Symbol] = [ym[j] for j in i_bkwrd_b] + list(ysym) + [yp[j] for j in i_fwrd_b] + usym
    jac: Dict[Tuple[int, int], sp.Expr] = {}
    for i, r in enumerate(resid):
        fs = r.free_symbols
        for c, sym in enumerate(cols):
            if sym in fs:
                d = sp.diff(r, sym).subs(to_ss)
                if d != 0:
                    jac[(i, c)] = d

    # --- shocks / measurement errors --------------------------------------------------------
    sigma_e0 = np.zeros((nx, nx))
    for u, expr in mf.shock_stderr.items():
        if u in exo:
            sigma_e0[exo.index(u), exo.index(u)] = float(parse(expr).subs(pval)) ** 2
    for u, expr in mf.shock_var.items():
        if u in exo:
            sigma_e0[exo.index(u), exo.index(u)] = float(parse(expr).subs(pval))
    for (a, b), expr in mf.shock_cov.items():
        if a in exo and b in exo:
            sigma_e0[exo.index(a), exo.index(b)] = sigma_e0[exo.index(b), exo.index(a)] = float(parse(expr).subs(pval))
    for (a, b), expr in mf.shock_corr.items():
        ia, ib = exo.index(a), exo.index(b)
        c = float(parse(expr).subs(pval)) * np.sqrt(sigma_e0[ia, ia] * sigma_e0[ib, ib])
        sigma_e0[ia, ib] = sigma_e0[ib, ia] = c
    ny = len(mf.varobs)
    sigma_m0 = np.zeros((ny, ny))
    # calibrated measurement errors: `var <observed variable>; stderr s;` etc. in the shocks block
    for u, expr in mf.shock_stderr.items():
        if u in mf.varobs:
            sigma_m0[mf.varobs.index(u), mf.varobs.index(u)] = float(parse(expr).subs(pval)) ** 2
    for u, expr in mf.shock_var.items():
        if u in mf.varobs:
            sigma_m0[mf.varobs.index(u), mf.varobs.index(u)] = float(parse(expr).subs(pval))
    for (a, b), expr in mf.shock_cov.items():
        if a in mf.varobs and b in mf.varobs:
            sigma_m0[mf.varobs.index(a), mf.varobs.index(b)] = sigma_m0[mf.varobs.index(b), mf.varobs.index(a)] = \
                float(parse(expr).subs(pval))

    # --- observation / state bookkeeping (estimation.jl:384-390) ---------------------------
    obs_idx = [endo.index(v) for v in mf.varobs]
    state_ids = sorted(set(obs_idx) | set(i_bkwrd_b))
    obs_idx_state = [state_ids.index(i) for i in obs_idx]
    lagged_state_ids = [p for p, i in enumerate(state_ids) if i in i_bkwrd_b]

    # --- estimated parameters (estimation.jl:517-601 ; dynare_containers.jl:800-841) ----------
    est_type, est_i, est_j = [], [], []
    for e in mf.estimated_params:
        if e.kind == "param":
            est_type.append(EST_PARAM); est_i.append(mf.params.index(e.name1)); est_j.append(0)
        elif e.kind == "stderr":
            if e.name1 in exo:
                est_type.append(EST_SD_SHOCK); est_i.append(exo.index(e.name1)); est_j.append(exo.index(e.name1))
            else:
                o = mf.varobs.index(e.name1)
                est_type.append(EST_SD_MEAS); est_i.append(o); est_j.append(o)
        elif e.kind == "var":
            if e.name1 in exo:
                est_type.append(EST_VAR_SHOCK); est_i.append(exo.index(e.name1)); est_j.append(exo.index(e.name1))
            else:
                o = mf.varobs.index(e.name1)
                est_type.append(EST_VAR_MEAS); est_i.append(o); est_j.append(o)
        elif e.kind == "corr":
            if e.name1 in exo:
                est_type.append(EST_CORR_SHOCK); est_i.append(exo.index(e.name1)); est_j.append(exo.index(e.name2))
            else:
                est_type.append(EST_CORR_MEAS); est_i.append(mf.varobs.index(e.name1)); est_j.append(mf.varobs.index(e.name2))

    return ModelSpec(
        name=name, endo=endo, n_orig=n_orig, exo=exo, params=list(mf.params),
        param_values=param_values, linear=mf.linear, lli=lli,
        i_bkwrd_b=i_bkwrd_b, i_fwrd_b=i_fwrd_b, i_both=i_both, i_static=i_static, i_dyn=i_dyn,
        sigma_e0=sigma_e0, sigma_m0=sigma_m0, varobs=list(mf.varobs), obs_idx=obs_idx,
        state_ids=state_ids, obs_idx_state=obs_idx_state, lagged_state_ids=lagged_state_ids,
        est=list(mf.estimated_params), est_type=est_type, est_i=est_i, est_j=est_j,
        estimation_options=dict(mf.estimation_options),
        psym=psym, ysym=ysym, ss_assign=ss_assign, static_resid=static_resid, jac=jac,
        ss_numeric=ss_numeric, static_jac=static_jac,
        dyn_resid=list(resid), dyn_syms=dict(lag=ym, lead=yp, exo=usym),
    )


def spec_from_file(name: str, path: str) -> ModelSpec:
    with open(path) as f:
        return build_spec(name, parse_mod(f.read()))


# =========================================================================================
# numeric evaluation on the host (used by tests and by the host-evaluated-Jacobian entry point)
# =========================================================================================
NEWTON_MAXIT, NEWTON_TOL, NEWTON_HALVINGS, NEWTON_ESCAPE, NEWTON_STEP_TOL = 50, 1e-13, 30, 1e8, 1e-6   # as in csrc/model_traits.cuh


class HostModel:
    """Lambdified steady state / static residuals / compressed Jacobian of a ``ModelSpec``."""

    def __init__(self, spec: ModelSpec):
        self.spec = spec
        p, y = spec.psym, spec.ysym
        # steady state: sequential substitution -> closed form in the parameters
        env: Dict[sp.Symbol, sp.Expr] = {}
        for tgt, expr in spec.ss_assign:
            env[tgt] = expr.subs(env)
        self._ss = sp.lambdify([p], [env[s_] for s_ in y], modules="numpy")
        self._res = sp.lambdify([p, y], spec.static_resid, modules="numpy")
        keys = sorted(spec.jac.keys(), key=lambda rc: (rc[1], rc[0]))
        self._jkeys = keys
        self._jac = sp.lambdify([p, y], [spec.jac[kx] for kx in keys], modules="numpy")

    def steady_state(self, params) -> np.ndarray:
        with np.errstate(all="ignore"):
            y = np.array(self._ss(list(params)), dtype=np.float64)
            if not self.spec.ss_numeric:
                return y
            return self._newton(list(params), y)

    def _newton(self, p, y):
        """The same iteration as newton_steady_state (csrc/model_traits.cuh): full Newton steps on the static model,
        halved while they do not reduce the sum of squared residuals; stops on max|f| <= 1e-13 (1 + max|y|)."""
        sp_ = self.spec
        n = sp_.n
        if not hasattr(self, "_sj"):
            self._sjkeys = sorted(sp_.static_jac.keys())
            self._sj = sp.lambdify([sp_.psym, sp_.ysym], [sp_.static_jac[kx] for kx in self._sjkeys], modules="numpy")
        y0max = max(1.0, np.abs(y).max())
        last_step = 0.0
        for _ in range(NEWTON_MAXIT):
            f = np.array(self._res(p, list(y)), dtype=np.float64)
            if not np.all(np.isfinite(f)) or np.abs(y).max() > NEWTON_ESCAPE * y0max:      # NaN, or a root at infinity
                return np.full(n, np.nan)
            if np.abs(f).max() <= NEWTON_TOL * (1.0 + np.abs(y).max()) and last_step <= NEWTON_STEP_TOL * (1.0 + np.abs(y).max()):
                return y
            J = np.zeros((n, n))
            for (r, c), v in zip(self._sjkeys, self._sj(p, list(y))):
                J[r, c] = v
            try:
                dy = np.linalg.solve(J, -f)
            except np.linalg.LinAlgError:
                return np.full(n, np.nan)
            g0 = f @ f
            t = 1.0
            for _h in range(NEWTON_HALVINGS):
                yt = y + t * dy
                ft = np.array(self._res(p, list(yt)), dtype=np.float64)
                if np.all(np.isfinite(ft)) and ft @ ft < g0:
                    break
                t *= 0.5
            else:
                return np.full(n, np.nan)
            last_step = np.abs(yt - y).max()
            y = yt
        return np.full(n, np.nan)

    def static_residuals(self, params, ybar) -> np.ndarray:
        with np.errstate(all="ignore"):
            return np.array(self._res(list(params), list(ybar)), dtype=np.float64)

    def jacobian(self, params, ybar) -> np.ndarray:
        """Compressed Jacobian, ``n x ncol`` (column blocks lags | current | leads | shocks)."""
        sp_ = self.spec
        J = np.zeros((sp_.n, sp_.ncol))
        with np.errstate(all="ignore"):
            vals = self._jac(list(params), list(ybar))
        for (r, c), v in zip(self._jkeys, vals):
            J[r, c] = v
        return J


# =========================================================================================
# code generation
# =========================================================================================
def _pow_opt(expr):
    from sympy.codegen.rewriting import create_expand_pow_optimization, optimize
    return optimize(expr, [create_expand_pow_optimization(4)])


class _Printer:
    def __init__(self):
        from sympy.printing.c import C99CodePrinter

        class P(C99CodePrinter):
            def _print_Rational(self, e):
                return f"({float(e.p)!r}/{float(e.q)!r})"

            def _print_Integer(self, e):
                return f"{float(int(e))!r}"

            def _print_Pow(self, e):
                b, ex = e.base, e.exp
                if ex.is_Integer and 2 <= int(ex) <= 4:
                    s_ = self._print(b)
                    if not (b.is_Symbol or b.is_Number):
                        s_ = f"({s_})"
                    return "(" + "*".join([s_] * int(ex)) + ")"
                if ex.is_Integer and -4 <= int(ex) <= -1:
                    return f"(1.0/({self._print(b ** (-ex))}))"
                if ex == sp.Rational(1, 2):
                    return f"sqrt({self._print(b)})"
                return f"pow({self._print(b)}, {self._print(ex)})"

        self.p = P()

    def __call__(self, e) -> str:
        return self.p.doprint(e)


def _lower(spec: ModelSpec):
    """Lower the symbolic model to three straight-line blocks of (name, C expression)."""
    pr = _Printer()
    psub = {s_: sp.Symbol(f"p[{i}]") for i, s_ in enumerate(spec.psym)}
    ysub = {s_: sp.Symbol(f"y[{i}]") for i, s_ in enumerate(spec.ysym)}
    sub = {**psub, **ysub}

    # steady state: keep the user's temporaries as locals
    ss_lines = []
    for tgt, expr in spec.ss_assign:
        e = expr.subs(sub)
        if tgt in ysub:
            ss_lines.append((str(ysub[tgt]), pr(e), False))
        else:
            sub[tgt] = sp.Symbol(tgt.name)
            ss_lines.append((tgt.name, pr(e), True))

    res_exprs = [r.subs(sub) for r in spec.static_resid]
    rrep, rred = sp.cse(res_exprs, symbols=sp.numbered_symbols("r_"), optimizations="basic")
    res_lines = [(str(a), pr(b), True) for a, b in rrep]
    res_out = [pr(e) for e in rred]

    keys = sorted(spec.jac.keys(), key=lambda rc: (rc[1], rc[0]))
    jexprs = [spec.jac[kx].subs(sub) for kx in keys]
    jrep, jred = sp.cse(jexprs, symbols=sp.numbered_symbols("t_"), optimizations="basic")
    jac_lines = [(str(a), pr(b), True) for a, b in jrep]
    jac_out = [(r, c, pr(e)) for (r, c), e in zip(keys, jred)]
    return ss_lines, res_lines, res_out, jac_lines, jac_out


def _lower_static_fj(spec: ModelSpec):
    """Static residuals and their Jacobian (n x n, column-major) as one straight-line block sharing subexpressions."""
    pr = _Printer()
    sub = {**{s_: sp.Symbol(f"p[{i}]") for i, s_ in enumerate(spec.psym)},
           **{s_: sp.Symbol(f"y[{i}]") for i, s_ in enumerate(spec.ysym)}}
    keys = sorted(spec.static_jac.keys(), key=lambda rc: (rc[1], rc[0]))
    exprs = [r.subs(sub) for r in spec.static_resid] + [spec.static_jac[kx].subs(sub) for kx in keys]
    rep, red = sp.cse(exprs, symbols=sp.numbered_symbols("s_"), optimizations="basic")
    lines = [(str(a), pr(b)) for a, b in rep]
    n = spec.n
    f_out = [pr(e) for e in red[:n]]
    j_out = [(r, c, pr(e)) for (r, c), e in zip(keys, red[n:])]
    return lines, f_out, j_out


def _int_list(v) -> str:
    v = list(v)
    return "{" + ", ".join(str(int(x)) for x in v) + "}" if v else "{0}"


def _dbl_list(v) -> str:
    v = [float(x) for x in np.asarray(v).ravel()]
    return "{" + ", ".join(repr(x) if np.isfinite(x) else "0.0" for x in v) + "}" if v else "{0.0}"



def plan_static_elim(S: ModelSpec, jperm):
    """Symbolic phase of the static elimination (reference: LRE ``remove_static!``, QR on the static
    columns): Householder reflectors restricted to the structurally non-zero rows, pivot rows chosen
    for least fill.  Zero entries contribute nothing to a reflector, so this is numerically the dense
    Householder QR of a row-permuted Jacobian; only the operation list is specialised to the model.
    Returns (initial pattern dict (r,c)->jac position, steps, pivot rows, remaining rows, final pattern)."""
    n, ncol, s = S.n, S.ncol, S.s
    keys = sorted(S.jac.keys(), key=lambda rc: (rc[1], rc[0]))
    pat = np.zeros((n, ncol), dtype=bool)
    init = {}
    for pos, (r, c) in enumerate(keys):
        pat[r, jperm[c]] = True
        init[(r, jperm[c])] = pos
    used, steps = [], []
    for j in range(s):
        cand = [r for r in range(n) if r not in used and pat[r, j]]
        if not cand:
            raise ValueError(f"{S.name}: static block structurally singular at column {j}")
        piv = min(cand, key=lambda r: (int(pat[r].sum()), r))
        rows = [piv] + [r for r in cand if r != piv]
        used.append(piv)
        cols = []
        if len(rows) > 1:
            for c in range(j + 1, ncol):
                nzr = [r for r in rows if pat[r, c]]
                if nzr:
                    cols.append((c, nzr))
                    pat[rows, c] = True
            pat[rows[1:], j] = False
        steps.append((j, rows, cols))
    rest = [r for r in range(n) if r not in used]
    return init, steps, used, rest, pat


def _emit_static_elim(S: ModelSpec, jperm, jac_lines, jac_out, A, arr):
    n, ncol, s = S.n, S.ncol, S.s
    init, steps, used, rest, pat = plan_static_elim(S, jperm)
    red = [(rr, c - s) for rr, r in enumerate(rest) for c in range(s, ncol) if pat[r, c]]
    top = [(q, c) for q, r in enumerate(used) for c in range(ncol) if pat[r, c]]
    A("")
    A("  // ---- structure-aware static elimination (symbolic phase done at model set-up) ----------------")
    A("  // Householder QR on the static columns restricted to structurally non-zero rows; outputs the")
    A("  // non-zeros of the reduced (dynamic) system, rows = remaining equations, columns [lags | dyn | leads | shocks],")
    A("  // then of the S pivot rows [R_s | lags | dyn | leads | shocks] used to recover the static variables.")
    A(f"  static constexpr int RED_NNZ = {len(red)};")
    A(f"  static constexpr int TOP_NNZ = {len(top)};")
    arr("red_row", [r for r, _ in red])
    arr("red_col", [c for _, c in red])
    arr("top_row", [r for r, _ in top])
    arr("top_col", [c for _, c in top])
    A("  template <class OT>")
    A("  DYB_HD static bool static_elim(const double* __restrict__ p, const double* __restrict__ y, OT& out) {")
    for nm, ex, _ in jac_lines:
        A(f"    const double {nm} = {ex};")
    keys = sorted(S.jac.keys(), key=lambda rc: (rc[1], rc[0]))
    expr_of = {(r, jperm[c]): ex for (r, c), (_, _, ex) in zip(keys, jac_out)}
    live = set()
    for (r, c), ex in sorted(expr_of.items(), key=lambda kv: (kv[0][1], kv[0][0])):
        A(f"    double w_{r}_{c} = {ex};")
        live.add((r, c))
    fill = set()
    for j, rows, cols in steps:
        for c, nzr in cols:
            for r in rows:
                if (r, c) not in live:
                    fill.add((r, c))
        if len(rows) > 1 and (rows[0], j) not in live:
            fill.add((rows[0], j))
    for (r, c) in sorted(fill, key=lambda rc: (rc[1], rc[0])):
        A(f"    double w_{r}_{c} = 0.0;   // fill-in")
    live |= fill
    A("    bool ok = true;")
    nfma = 0
    for j, rows, cols in steps:
        if len(rows) == 1:
            continue
        piv, oth = rows[0], rows[1:]
        A(f"    {{  // static column {j}: reflector over equations {rows}")
        A("      const double s2 = " + " + ".join(f"w_{r}_{j}*w_{r}_{j}" for r in oth) + ";")
        A(f"      const double xp = w_{piv}_{j};")
        A("      const double sn = fma(xp, xp, s2);")
        A("      ok = ok && (sn > 0.0);")
        A("      const double nrm = sqrt(sn);")
        A("      const double alpha = xp >= 0.0 ? -nrm : nrm;")
        A("      const double vp = xp - alpha;")
        A("      const double beta = 1.0 / (-alpha * vp);")
        for c, nzr in cols:
            terms = []
            for r in nzr:
                terms.append(f"vp*w_{r}_{c}" if r == piv else f"w_{r}_{j}*w_{r}_{c}")
            A(f"      {{ const double dt = beta*({' + '.join(terms)});")
            for r in rows:
                coef = "vp" if r == piv else f"w_{r}_{j}"
                A(f"        w_{r}_{c} = fma(-dt, {coef}, w_{r}_{c});")
                nfma += 2
            A("      }")
            # declare fill-in variables before the block that assigns them
        A(f"      w_{piv}_{j} = alpha;")
        A("    }")
    for e, (rr, c) in enumerate(red):
        A(f"    out[{e}] = w_{rest[rr]}_{c + s};")
    for e, (q, c) in enumerate(top):
        A(f"    out[{len(red) + e}] = w_{used[q]}_{c};")
    A("    return ok;")
    A("  }")
    A(f"  static constexpr int STATIC_ELIM_FMA = {nfma};")


def emit_cuda(spec: ModelSpec) -> str:
    """CUDA header: a traits struct with compile-time sizes/index sets and the device model function."""
    ss_lines, res_lines, res_out, jac_lines, jac_out = _lower(spec)
    S = spec
    pos_in_dyn = {v: i for i, v in enumerate(S.i_dyn)}
    L = []
    A = L.append
    A(f"// GENERATED by dynare_jl_b200.modelgen from model '{S.name}' -- do not edit.")
    A("// Replaces the preprocessor-generated SparseDynamicG1!/SteadyState2 functions the reference loads")
    A("// in src/dynare_functions.jl:10-47 and calls from src/model.jl:120-139.")
    A("#pragma once")
    A("#include \"../model_traits.cuh\"")
    A("namespace dyb {")
    A(f"struct Model_{S.name} {{")
    A(f"  static constexpr const char* name = \"{S.name}\";")
    for nm, val in [("N", S.n), ("NX", S.nx), ("NPAR", len(S.params)), ("K", S.k), ("NF", S.nf),
                    ("S", S.s), ("M", S.m), ("NCOL", S.ncol), ("NS", S.ns), ("NY", S.ny),
                    ("NL", S.nl), ("NP_DEFAULT", S.np_est)]:
        A(f"  static constexpr int {nm} = {val};")
    A(f"  static constexpr bool LINEAR = {'true' if S.linear else 'false'};")

    def arr(nm, v):
        v = list(v)
        A(f"  DYB_HD static constexpr int {nm}(int i) {{ constexpr int a[{max(len(v), 1)}] = {_int_list(v)}; return a[i]; }}")
    arr("i_bkwrd_b", S.i_bkwrd_b)
    arr("i_fwrd_b", S.i_fwrd_b)
    arr("i_static", S.i_static)
    arr("i_dyn", S.i_dyn)
    arr("bkwrd_in_dyn", [pos_in_dyn[v] for v in S.i_bkwrd_b])
    arr("fwrd_in_dyn", [pos_in_dyn[v] for v in S.i_fwrd_b])
    arr("state_ids", S.state_ids)
    arr("obs_idx_state", S.obs_idx_state)
    arr("lagged_state_ids", S.lagged_state_ids)
    # position of each lagged state inside i_bkwrd_b (column of g1_1)
    arr("lagged_in_bkwrd", [S.i_bkwrd_b.index(S.state_ids[p]) for p in S.lagged_state_ids])
    # variable -> (is static, position inside i_static / i_dyn)
    pos_in_static = {v: i for i, v in enumerate(S.i_static)}
    arr("var_static", [1 if v in pos_in_static else 0 for v in range(S.n)])
    arr("var_pos", [pos_in_static[v] if v in pos_in_static else pos_in_dyn[v] for v in range(S.n)])
    # compressed-Jacobian column -> working-matrix column [static | lags | dynamic current | leads | shocks]
    jperm = [S.s + j for j in range(S.k)]
    jperm += [pos_in_static[v] if v in pos_in_static else S.s + S.k + pos_in_dyn[v] for v in range(S.n)]
    jperm += [S.s + S.k + S.m + j for j in range(S.nf)]
    jperm += [S.s + S.k + S.m + S.nf + j for j in range(S.nx)]
    arr("jperm", jperm)
    # defaults taken from the model file (calibration, shocks block, estimated_params block)
    arr("est_type0", S.est_type)
    arr("est_i0", S.est_i)
    arr("est_j0", S.est_j)

    def darr(nm, v):
        v = [float(x) for x in np.asarray(v).ravel()]
        A(f"  DYB_HD static constexpr double {nm}(int i) {{ constexpr double a[{max(len(v), 1)}] = {_dbl_list(v)}; return a[i]; }}")
    darr("param0", S.param_values)
    darr("sigma_e0", S.sigma_e0.T)
    darr("sigma_m0", S.sigma_m0.T)
    A("")
    if S.ss_numeric:
        lines, f_out, j_out = _lower_static_fj(S)
        A("  // numerical steady state: Newton's method on the static model from the initval point")
        A("  // (reference: `steady` -> solve_steady_state!, src/steady_state/SteadyState.jl:183-252)")
        A("  static constexpr bool NUMERIC_SS = true;")
        A("  DYB_HD static void initval(const double* __restrict__ p, double* __restrict__ y) {")
        for nm, ex, is_local in ss_lines:
            A(f"    {'const double ' if is_local else ''}{nm} = {ex};")
        A("  }")
        A("  // f = static residuals, J = their Jacobian (N x N column-major, zero-initialised by the caller)")
        A("  DYB_HD static void static_fj(const double* __restrict__ p, const double* __restrict__ y,")
        A("                               double* __restrict__ f, double* __restrict__ J) {")
        for nm, ex in lines:
            A(f"    const double {nm} = {ex};")
        for i, ex in enumerate(f_out):
            A(f"    f[{i}] = {ex};")
        for r, c, ex in j_out:
            A(f"    J[{r + S.n * c}] = {ex};")
        A("  }")
        A("  DYB_HD static void steady_state(const double* __restrict__ p, double* __restrict__ y);")
    else:
        A("  // closed-form steady state (reference: generated steady_state!, src/dynare_functions.jl:39-41)")
        A("  DYB_HD static void steady_state(const double* __restrict__ p, double* __restrict__ y) {")
        for nm, ex, is_local in ss_lines:
            A(f"    {'const double ' if is_local else ''}{nm} = {ex};")
        A("  }")
    A("  // sum of squared static residuals at y (reference: check_steady_state, src/steady_state/SteadyState.jl:476-482)")
    A("  DYB_HD static double static_resid_ss(const double* __restrict__ p, const double* __restrict__ y) {")
    for nm, ex, _ in res_lines:
        A(f"    const double {nm} = {ex};")
    A("    double acc = 0.0, r;")
    for ex in res_out:
        A(f"    r = {ex}; acc += r*r;")
    A("    return acc;")
    A("  }")
    A("  // non-zeros of the compressed dynamic Jacobian, n x NCOL column-major, at the steady state, u = 0")
    A("  // (reference: get_dynamic_jacobian! + set_compressed_jacobian!, src/model.jl:120-139, src/perturbations.jl:616-635)")
    A("  // J must be zero-initialised by the caller; JT is any type with operator[](int).")
    A("  template <class JT>")
    A("  DYB_HD static void jacobian(const double* __restrict__ p, const double* __restrict__ y, JT& J) {")
    for nm, ex, _ in jac_lines:
        A(f"    const double {nm} = {ex};")
    for r, c, ex in jac_out:
        A(f"    J[{r + S.n * c}] = {ex};")
    A("  }")
    A(f"  static constexpr int JAC_NNZ = {len(jac_out)};")
    _emit_static_elim(S, jperm, jac_lines, jac_out, A, arr)
    A("};")
    if S.ss_numeric:
        A(f"DYB_HD void Model_{S.name}::steady_state(const double* __restrict__ p, double* __restrict__ y) {{")
        A("  initval(p, y);")
        A(f"  newton_steady_state<Model_{S.name}>(p, y);")
        A("}")
    A("}  // namespace dyb")
    return "\n".join(L) + "\n"


def emit_cuda_jacobian_fed(name: str, n: int, nx: int, i_static, i_bkwrd_b, i_fwrd_b, obs_idx) -> str:
    """CUDA header of a SIZES-ONLY model: the traits struct the register-resident kernels are templated on -- compile-time
    sizes and the reference's index sets (``Model.i_static / i_bkwrd_b / i_fwrd_b``, src/dynare_containers.jl:110-131,
    0-based) -- with NO model function: the host keeps evaluating steady state and compressed Jacobian per draw (as the
    reference does, src/perturbations.jl:526-561, 616-635) and ``jac_model_kernel`` (csrc/solve_coop.cuh) takes them from
    memory.  The hand-off patterns are dense (RED_NNZ = M (NCOL - S), TOP_NNZ = S NCOL).  ``JAC_FED`` marks the struct."""
    i_static = [int(v) for v in i_static]; i_bkwrd_b = [int(v) for v in i_bkwrd_b]
    i_fwrd_b = [int(v) for v in i_fwrd_b]; obs_idx = [int(v) for v in obs_idx]
    if sorted(set(i_static) | set(i_bkwrd_b) | set(i_fwrd_b)) != list(range(n)) or set(i_static) & (set(i_bkwrd_b) | set(i_fwrd_b)):
        raise ValueError("i_static, i_bkwrd_b and i_fwrd_b must cover 0..n-1, the static set disjoint from the other two")
    i_dyn = [v for v in range(n) if v not in set(i_static)]
    s_, k, nf, m = len(i_static), len(i_bkwrd_b), len(i_fwrd_b), len(i_dyn)
    ncol = k + n + nf + nx
    state_ids = sorted(set(obs_idx) | set(i_bkwrd_b))
    obs_idx_state = [state_ids.index(i) for i in obs_idx]
    lagged_state_ids = [p for p, i in enumerate(state_ids) if i in set(i_bkwrd_b)]
    pos_in_dyn = {v: i for i, v in enumerate(i_dyn)}
    pos_in_static = {v: i for i, v in enumerate(i_static)}
    ny, ns = len(obs_idx), len(state_ids)
    L = []
    A = L.append
    A(f"// GENERATED by dynare_jl_b200.modelgen.emit_cuda_jacobian_fed for '{name}' -- do not edit.")
    A("// Sizes and index sets only: steady state and Jacobian are evaluated by the host and uploaded per draw.")
    A("#pragma once")
    A("#include \"../model_traits.cuh\"")
    A("namespace dyb {")
    A(f"struct Model_{name} {{")
    A(f"  static constexpr const char* name = \"{name}\";")
    A("  static constexpr bool JAC_FED = true;")
    for nm, val in [("N", n), ("NX", nx), ("NPAR", 0), ("K", k), ("NF", nf), ("S", s_), ("M", m), ("NCOL", ncol),
                    ("NS", ns), ("NY", ny), ("NL", len(lagged_state_ids)), ("NP_DEFAULT", 0)]:
        A(f"  static constexpr int {nm} = {val};")
    A("  static constexpr bool LINEAR = true;")

    def arr(nm, v):
        v = list(v)
        A(f"  DYB_HD static constexpr int {nm}(int i) {{ constexpr int a[{max(len(v), 1)}] = {_int_list(v)}; return a[i]; }}")
    arr("i_bkwrd_b", i_bkwrd_b); arr("i_fwrd_b", i_fwrd_b); arr("i_static", i_static); arr("i_dyn", i_dyn)
    arr("bkwrd_in_dyn", [pos_in_dyn[v] for v in i_bkwrd_b]); arr("fwrd_in_dyn", [pos_in_dyn[v] for v in i_fwrd_b])
    arr("state_ids", state_ids); arr("obs_idx_state", obs_idx_state); arr("lagged_state_ids", lagged_state_ids)
    arr("lagged_in_bkwrd", [i_bkwrd_b.index(state_ids[p]) for p in lagged_state_ids])
    arr("var_static", [1 if v in pos_in_static else 0 for v in range(n)])
    arr("var_pos", [pos_in_static[v] if v in pos_in_static else pos_in_dyn[v] for v in range(n)])
    jperm = [s_ + j for j in range(k)]
    jperm += [pos_in_static[v] if v in pos_in_static else s_ + k + pos_in_dyn[v] for v in range(n)]
    jperm += [s_ + k + m + j for j in range(nf)]
    jperm += [s_ + k + m + nf + j for j in range(nx)]
    arr("jperm", jperm)
    arr("est_type0", []); arr("est_i0", []); arr("est_j0", [])
    A("  DYB_HD static constexpr double param0(int) { return 0.0; }")
    A("  DYB_HD static constexpr double sigma_e0(int) { return 0.0; }")
    A("  DYB_HD static constexpr double sigma_m0(int) { return 0.0; }")
    A("  // no model function: these exist so that the shared translation unit compiles; they are never launched")
    A("  DYB_HD static void steady_state(const double*, double* y) { for (int i = 0; i < N; ++i) y[i] = 0.0; }")
    A("  DYB_HD static double static_resid_ss(const double*, const double*) { return 0.0; }")
    A("  template <class JT> DYB_HD static void jacobian(const double*, const double*, JT&) {}")
    A("  static constexpr int JAC_NNZ = 0;")
    A(f"  static constexpr int RED_NNZ = {m * (ncol - s_)};")
    A(f"  static constexpr int TOP_NNZ = {s_ * ncol};")
    A(f"  DYB_HD static constexpr int red_row(int i) {{ return i / {ncol - s_}; }}")
    A(f"  DYB_HD static constexpr int red_col(int i) {{ return i % {ncol - s_}; }}")
    A(f"  DYB_HD static constexpr int top_row(int i) {{ return i / {ncol}; }}")
    A(f"  DYB_HD static constexpr int top_col(int i) {{ return i % {ncol}; }}")
    A("  template <class OT> DYB_HD static bool static_elim(const double*, const double*, OT&) { return false; }")
    A("  static constexpr int STATIC_ELIM_FMA = 0;")
    A("};")
    A("}  // namespace dyb")
    return "\n".join(L) + "\n"


def emit_c(spec: ModelSpec) -> str:
    if spec.ss_numeric:
        raise NotImplementedError("the C restatement covers models with a closed-form steady state only")
    """Plain-C model functions + descriptor for the CPU oracle port (oracle/c)."""
    ss_lines, res_lines, res_out, jac_lines, jac_out = _lower(spec)
    S = spec
    nm = S.name
    L = []
    A = L.append
    A(f"/* GENERATED by dynare_jl_b200.modelgen from model '{nm}' -- do not edit. Test infrastructure (CPU oracle). */")
    A("#include <math.h>")
    A("#include \"oracle_model.h\"")
    A(f"static void {nm}_steady_state(const double* p, double* y) {{")
    for n_, ex, is_local in ss_lines:
        A(f"  {'const double ' if is_local else ''}{n_} = {ex};")
    A("}")
    A(f"static double {nm}_static_resid_ss(const double* p, const double* y) {{")
    for n_, ex, _ in res_lines:
        A(f"  const double {n_} = {ex};")
    A("  double acc = 0.0, r;")
    for ex in res_out:
        A(f"  r = {ex}; acc += r*r;")
    A("  return acc;")
    A("}")
    A(f"static void {nm}_jacobian(const double* p, const double* y, double* J) {{")
    for n_, ex, _ in jac_lines:
        A(f"  const double {n_} = {ex};")
    for r, c, ex in jac_out:
        A(f"  J[{r + S.n * c}] = {ex};")
    A("}")
    for an, v in [("i_bkwrd_b", S.i_bkwrd_b), ("i_fwrd_b", S.i_fwrd_b), ("i_static", S.i_static),
                  ("i_dyn", S.i_dyn), ("state_ids", S.state_ids), ("obs_idx_state", S.obs_idx_state),
                  ("lagged_state_ids", S.lagged_state_ids), ("est_type", S.est_type),
                  ("est_i", S.est_i), ("est_j", S.est_j)]:
        A(f"static const int {nm}_{an}[] = {_int_list(v)};")
    A(f"static const double {nm}_params0[] = {_dbl_list(S.param_values)};")
    A(f"static const double {nm}_sigma_e0[] = {_dbl_list(S.sigma_e0.T)};")
    A(f"static const double {nm}_sigma_m0[] = {_dbl_list(S.sigma_m0.T)};")
    A(f"const oracle_model oracle_model_{nm} = {{")
    A(f"  \"{nm}\", {S.n}, {S.nx}, {len(S.params)}, {S.k}, {S.nf}, {S.s}, {S.ncol}, {S.ns}, {S.ny}, {S.nl}, {S.np_est},")
    A(f"  {nm}_i_bkwrd_b, {nm}_i_fwrd_b, {nm}_i_static, {nm}_i_dyn, {nm}_state_ids, {nm}_obs_idx_state, {nm}_lagged_state_ids,")
    A(f"  {nm}_est_type, {nm}_est_i, {nm}_est_j, {nm}_params0, {nm}_sigma_e0, {nm}_sigma_m0,")
    A(f"  {nm}_steady_state, {nm}_static_resid_ss, {nm}_jacobian")
    A("};")
    return "\n".join(L) + "\n"


def write_json(spec: ModelSpec, path: str) -> None:
    with open(path, "w") as f:
        json.dump(spec.describe(), f, indent=1)
