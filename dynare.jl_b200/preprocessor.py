"""Device model functions from the REFERENCE'S OWN front end.

Dynare.jl never reads model equations itself: its preprocessor (``DynarePreprocessor_jll``, ``src/DynarePreprocessor.jl:4-40``)
writes, under ``<modfile>/model/``,

  json/modfile.json                 symbol tables, ``model_info`` (lead/lag incidence), the sparse pattern of the dynamic
                                    Jacobian (``dynamic_g1_sparse_{rowval,colval,colptr}``) and the ``statements`` of the
                                    mod file (``param_init``, ``shocks``, ``estimated_params``, ``varobs``, ...)
                                    -- read by ``parseJSON`` / ``make_context`` (``src/DynareParser.jl:10-15, 169-218``)
  julia/SparseDynamicResidTT!.jl, SparseDynamicG1TT!.jl, SparseDynamicG1!.jl,
  julia/SparseStaticResidTT!.jl, SparseStaticResid!.jl, SteadyState2.jl
                                    straight-line Julia: ``T[i] = ...``, ``g1_v[k] = ...``, ``residual[i] = ...``,
                                    ``ys_[i] = ...`` -- loaded as RuntimeGeneratedFunctions
                                    (``src/dynare_functions.jl:10-47, 304-342``)

and a ``context`` is nothing but those two things.  ``spec_from_preprocessor`` turns the same directory into the
``ModelSpec`` that ``modelgen.emit_cuda`` compiles into device functions, so a model that Dynare.jl can load can be
handed to the engine WITHOUT going through this repo's own ``.mod`` reader (``modfile.py``):

    spec = preprocessor.spec_from_preprocessor("fs2000", "path/to/fs2000")       # the directory @dynare created
    lib  = preprocessor.build_model_library_from_preprocessor("fs2000", "path/to/fs2000", "build/fs2000")

The generated code is translated statement by statement (1-based ``y[j]`` / ``x[j]`` / ``params[j]`` / ``T[j]`` /
``steady_state[j]`` / ``ys_[j]`` references, ``^`` powers, ``get_power_deriv``) into symbolic expressions; temporary terms
are substituted; the dynamic Jacobian is evaluated at ``y3 = [ybar; ybar; ybar]``, ``x = 0`` as the reference does
(``src/perturbations.jl:526-531``) and its columns are compressed as ``set_compressed_jacobian!`` does (``:616-635``).
"""
from __future__ import annotations

import json
import os
import re
from typing import Dict, List, Tuple

import numpy as np
import sympy as sp

from .modelgen import (EST_CORR_MEAS, EST_CORR_SHOCK, EST_PARAM, EST_SD_MEAS, EST_SD_SHOCK, ModelSpec)
from .modfile import EstimatedParam

# prior_distribution codes of the preprocessor (src/estimation/priors.jl:482-529)
_SHAPE_OF_CODE = {1: "beta", 2: "gamma", 3: "normal", 4: "inv_gamma", 5: "uniform", 6: "inv_gamma2", 8: "weibull"}

_ASSIGN = re.compile(r"^\s*(T|residual|g1_v|ys_|[A-Za-z_]\w*)\s*(?:\[\s*(\d+)\s*\])?\s*=\s*(.+?);?\s*$")
_INDEXED = re.compile(r"\b(y|x|params|T|steady_state|ys_|exo_)\[\s*(\d+)\s*\]")


def _power_deriv(x, p, k):
    """``get_power_deriv(x, p, k)`` (src/dynare_functions.jl:53-64): k-th derivative of x^p, away from its x = 0 guard."""
    k = int(k)
    c = sp.Integer(1)
    for i in range(k):
        c = c * (p - i)
    return c * x ** (p - k)


_JL_FUNCS = {"exp": sp.exp, "log": sp.log, "sqrt": sp.sqrt, "abs": sp.Abs, "sin": sp.sin, "cos": sp.cos, "tan": sp.tan,
             "log10": lambda v: sp.log(v) / sp.log(10), "cbrt": lambda v: sp.real_root(v, 3), "erf": sp.erf,
             "get_power_deriv": _power_deriv, "max": sp.Max, "min": sp.Min}


class _Env:
    """Symbols behind the indexed Julia references; temporaries / locals are kept as substituted expressions."""

    def __init__(self, ysyms, xsyms, psyms, sssyms=None, ys_syms=None):
        self.tab = {"y": ysyms, "x": xsyms, "params": psyms, "steady_state": sssyms or [], "ys_": ys_syms or [],
                    "exo_": xsyms}
        self.T: Dict[int, sp.Expr] = {}
        self.locals: Dict[str, sp.Expr] = {}

    def expr(self, text: str) -> sp.Expr:
        names: Dict[str, sp.Expr] = dict(_JL_FUNCS)
        names.update(self.locals)

        def repl(m):
            arr, idx = m.group(1), int(m.group(2))
            key = f"{arr}__{idx}"
            if arr == "T":
                names[key] = self.T[idx]
            else:
                names[key] = self.tab[arr][idx - 1]
            return key
        py = _INDEXED.sub(repl, text).replace("^", "**")
        return sp.sympify(py, locals=names)


def _statements(path: str) -> List[Tuple[str, int, str]]:
    """(target, 1-based index or 0, right-hand side) of every assignment of a generated Julia file; [] when absent."""
    if not os.path.exists(path):
        return []
    out = []
    with open(path) as f:
        for ln in f:
            s = ln.strip()
            if (not s or s.startswith(("function", "end", "@", "return", "#", "using", "if ", "else"))):
                continue
            m = _ASSIGN.match(s)
            if m:
                out.append((m.group(1), int(m.group(2) or 0), m.group(3)))
    return out


def _num(text, pval: Dict[str, float]) -> float:
    """``dynare_parse_eval`` of a numeric JSON field: a literal or an expression in already calibrated parameters."""
    if isinstance(text, (int, float)):
        return float(text)
    t = str(text).strip().replace("^", "**")
    low = t.lower().strip("()")
    if low in ("nan",):
        return float("nan")
    if low in ("inf", "+inf"):
        return float("inf")
    if low == "-inf":
        return float("-inf")
    names = {k: sp.Float(v) for k, v in pval.items()}
    names.update({"exp": sp.exp, "log": sp.log, "sqrt": sp.sqrt})
    return float(sp.sympify(t, locals=names))


def spec_from_preprocessor(name: str, modpath: str) -> ModelSpec:
    """``modpath`` is the directory ``@dynare`` creates beside the mod file (it contains ``model/json/modfile.json``)."""
    with open(os.path.join(modpath, "model", "json", "modfile.json")) as f:
        mj = json.load(f)
    jl = os.path.join(modpath, "model", "julia")
    endo = [e["name"] for e in mj["endogenous"]]
    exo = [e["name"] for e in mj["exogenous"]]
    params = [e["name"] for e in mj["parameters"]]
    n, nx, n_orig = len(endo), len(exo), int(mj.get("orig_endo_nbr", len(endo)))
    if mj.get("exogenous_deterministic"):
        raise NotImplementedError("exogenous deterministic variables")

    # ---- lead / lag incidence and the index sets of Model (src/dynare_containers.jl:258-390) ---------------------------
    inc = np.array(mj["model_info"]["lead_lag_incidence"], dtype=int)        # one row [lag, current, lead] per variable
    if inc.shape == (3, n) and n != 3:
        inc = inc.T
    if inc.shape[1] != 3:
        raise NotImplementedError("models with more than one lead or lag in the transformed model")
    lli = (inc.T > 0)
    i_bkwrd_b = [j for j in range(n) if lli[0, j]]
    i_fwrd_b = [j for j in range(n) if lli[2, j]]
    i_both = [j for j in range(n) if lli[0, j] and lli[2, j]]
    i_static = [j for j in range(n) if not lli[0, j] and not lli[2, j]]
    i_dyn = [j for j in range(n) if lli[0, j] or lli[2, j]]

    # ---- symbols ----------------------------------------------------------------------------------------------------
    ysym = [sp.Symbol(v, real=True) for v in endo]
    psym = [sp.Symbol(p, real=True) for p in params]
    zero_x = [sp.Integer(0)] * nx
    y3_at_ss = ysym * 3                                       # y3 = [ybar; ybar; ybar]  (src/perturbations.jl:526-531)

    # ---- statements: calibration, shocks, estimated parameters, varobs, estimation options -----------------------------
    pval: Dict[str, float] = {}
    sigma_e0 = np.zeros((nx, nx))
    est: List[EstimatedParam] = []
    varobs: List[str] = [str(v) for v in mj.get("varobs", [])]
    shocks_meas: List[Tuple[str, str, object]] = []
    estimation_options: Dict[str, str] = {}
    init_rows: List[dict] = []
    for st in mj.get("statements", []):
        sn = st.get("statementName")
        if sn == "param_init":
            pval[st["name"]] = _num(st["value"], pval)
        elif sn == "varobs":
            varobs = [str(v) for v in st.get("vars", st.get("varobs", []))]
        elif sn == "shocks":
            for row in st.get("variance", []):
                shocks_meas.append(("var", row["name"], row["variance"]))
            for row in st.get("stderr", []):
                shocks_meas.append(("stderr", row["name"], row["stderr"]))
            for row in st.get("covariance", []):
                shocks_meas.append(("cov", (row["name"], row["name2"]), row["covariance"]))
            for row in st.get("correlation", []):
                shocks_meas.append(("corr", (row["name"], row["name2"]), row["correlation"]))
        elif sn == "estimated_params":
            for p in st["params"]:
                shape = _SHAPE_OF_CODE[int(p["prior_distribution"])]
                mean, std = _num(p.get("mean", "NaN"), pval), _num(p.get("std", "NaN"), pval)
                init = _num(p.get("init_val", "NaN"), pval)
                dom = None
                if shape == "uniform" and np.isnan(mean) and np.isnan(std):       # priors.jl:508-517
                    dom = (_num(p["p3"], pval), _num(p["p4"], pval))
                    mean, std = 0.5 * (dom[0] + dom[1]), (dom[1] - dom[0]) / 12 ** 0.5
                init = None if np.isnan(init) else init
                if "param" in p:
                    est.append(EstimatedParam("param", p["param"], None, shape, mean, std, init, dom))
                elif "var" in p:
                    est.append(EstimatedParam("stderr", p["var"], None, shape, mean, std, init, dom))
                else:
                    est.append(EstimatedParam("corr", p["var1"], p["var2"], shape, mean, std, init, dom))
        elif sn == "estimated_params_init":
            init_rows += st["params"]
        elif sn == "estimation":
            estimation_options = {k: str(v) for k, v in st.get("options", {}).items()}
    for p in init_rows:                                        # parse_estimated_parameter_init! (priors.jl:442-458)
        key = ("param", p["param"], None) if "param" in p else (("stderr", p["var"], None) if "var" in p else
                                                                ("corr", p["var1"], p["var2"]))
        for e in est:
            if (e.kind, e.name1, e.name2) == key:
                e.init = _num(p["init_val"], pval)
    ny = len(varobs)
    sigma_m0 = np.zeros((ny, ny))
    for kind, who, val in shocks_meas:                         # exogenous -> Sigma_e, observed endogenous -> Sigma_m
        v = _num(val, pval)
        if kind in ("var", "stderr"):
            S, idx = (sigma_e0, exo) if who in exo else (sigma_m0, varobs)
            if who in idx:
                S[idx.index(who), idx.index(who)] = v * v if kind == "stderr" else v
    for kind, who, val in shocks_meas:
        if kind in ("cov", "corr"):
            a, b = who
            S, idx = (sigma_e0, exo) if a in exo else (sigma_m0, varobs)
            ia, ib = idx.index(a), idx.index(b)
            v = _num(val, pval)
            S[ia, ib] = S[ib, ia] = v if kind == "cov" else v * np.sqrt(S[ia, ia] * S[ib, ib])
    param_values = np.array([pval.get(p, np.nan) for p in params], dtype=np.float64)

    # ---- steady state: SteadyState2.jl (src/dynare_functions.jl:333-342) -------------------------------------------------
    ss_assign: List[Tuple[sp.Symbol, sp.Expr]] = []
    env = _Env(ysym, zero_x, psym, ys_syms=ysym)
    temps: Dict[str, sp.Symbol] = {}
    stmts = _statements(os.path.join(jl, "SteadyState2.jl"))
    for tgt, idx, rhs in stmts:
        if tgt == "ys_":
            ss_assign.append((ysym[idx - 1], env.expr(rhs)))
        else:                                                  # a local of the steady_state_model block
            sym = temps.setdefault(tgt, sp.Symbol(f"ss_{tgt}", real=True))
            ss_assign.append((sym, env.expr(rhs)))
            env.locals[tgt] = sym
    assigned = {s.name for s, _ in ss_assign}
    # auxiliary variables not set by the block: SetAuxiliaryVariables.jl (src/dynare_functions.jl:44-45)
    aux_env = _Env(ysym, zero_x, psym)
    for tgt, idx, rhs in _statements(os.path.join(jl, "SetAuxiliaryVariables.jl")):
        if tgt == "y" and endo[idx - 1] not in assigned:
            ss_assign.append((ysym[idx - 1], aux_env.expr(rhs)))
            assigned.add(endo[idx - 1])
    missing_ss = [v for v in endo if v not in assigned]

    # ---- static residuals: SparseStaticResidTT! / SparseStaticResid! -----------------------------------------------------
    senv = _Env(ysym, zero_x, psym)
    for tgt, idx, rhs in _statements(os.path.join(jl, "SparseStaticResidTT!.jl")):
        if tgt == "T":
            senv.T[idx] = senv.expr(rhs)
    static_resid: List[sp.Expr] = [sp.Integer(0)] * n
    got = 0
    for tgt, idx, rhs in _statements(os.path.join(jl, "SparseStaticResid!.jl")):
        if tgt == "residual":
            static_resid[idx - 1] = senv.expr(rhs); got += 1
        elif tgt == "T":
            senv.T[idx] = senv.expr(rhs)

    # ---- dynamic Jacobian at the steady state: temporary terms, then g1_v in CSC order --------------------------------------
    denv = _Env(y3_at_ss, zero_x, psym, sssyms=ysym)
    for fn in ("SparseDynamicResidTT!.jl", "SparseDynamicG1TT!.jl"):
        for tgt, idx, rhs in _statements(os.path.join(jl, fn)):
            if tgt == "T":
                denv.T[idx] = denv.expr(rhs)
    if got != n:                                               # static residuals = dynamic residuals at the steady state
        static_resid = [sp.Integer(0)] * n
        for tgt, idx, rhs in _statements(os.path.join(jl, "SparseDynamicResid!.jl")):
            if tgt == "residual":
                static_resid[idx - 1] = denv.expr(rhs)
    if missing_ss:
        # no (complete) steady_state_model block: fine for a model whose steady state is zero (`model(linear)`), which is
        # the case when the static residuals vanish at y = 0; anything else needs the reference's numerical solver
        at0 = {s_: 0 for s_ in ysym}
        at0.update({p_: float(v_) for p_, v_ in zip(psym, param_values) if not np.isnan(v_)})
        if all(abs(complex(sp.N(r.subs(at0)))) < 1e-14 for r in static_resid):
            ss_assign += [(ysym[endo.index(v)], sp.Float(0.0)) for v in missing_ss]
        else:
            raise NotImplementedError(f"{name}: the generated files define no steady state for {missing_ss} (models "
                                      "without a steady_state_model block go through the .mod route, which runs Newton's "
                                      "method on the device)")
    colptr = [int(v) for v in mj["dynamic_g1_sparse_colptr"]]
    rowval = [int(v) for v in mj["dynamic_g1_sparse_rowval"]]
    if len(colptr) != 3 * n + nx + 1:
        raise ValueError(f"{name}: dynamic_g1_sparse_colptr has {len(colptr)} entries for a {n} x {3 * n + nx} Jacobian")
    col_of_nz = np.zeros(len(rowval), dtype=int)
    for c in range(3 * n + nx):
        col_of_nz[colptr[c] - 1:colptr[c + 1] - 1] = c
    # full column -> compressed column (src/perturbations.jl:616-635): [lags of i_bkwrd_b | current | leads of i_fwrd_b | shocks]
    k, nf = len(i_bkwrd_b), len(i_fwrd_b)
    comp = {}
    for q, j in enumerate(i_bkwrd_b):
        comp[j] = q
    for j in range(n):
        comp[n + j] = k + j
    for q, j in enumerate(i_fwrd_b):
        comp[2 * n + j] = k + n + q
    for e in range(nx):
        comp[3 * n + e] = k + n + nf + e
    jac: Dict[Tuple[int, int], sp.Expr] = {}
    for tgt, idx, rhs in _statements(os.path.join(jl, "SparseDynamicG1!.jl")):
        if tgt == "T":
            denv.T[idx] = denv.expr(rhs)
        elif tgt == "g1_v":
            c_full = int(col_of_nz[idx - 1])
            if c_full not in comp:
                raise ValueError(f"{name}: Jacobian entry {idx} sits in column {c_full + 1}, which the incidence matrix excludes")
            val = denv.expr(rhs)
            if val != 0:
                jac[(rowval[idx - 1] - 1, comp[c_full])] = val

    # ---- observation / state bookkeeping (estimation.jl:384-390), estimated-parameter map (:517-601) -------------------------
    obs_idx = [endo.index(v) for v in varobs]
    state_ids = sorted(set(obs_idx) | set(i_bkwrd_b))
    obs_idx_state = [state_ids.index(i) for i in obs_idx]
    lagged_state_ids = [p for p, i in enumerate(state_ids) if i in i_bkwrd_b]
    est_type, est_i, est_j = [], [], []
    for e in est:
        if e.kind == "param":
            est_type.append(EST_PARAM); est_i.append(params.index(e.name1)); est_j.append(0)
        elif e.kind == "stderr":
            if e.name1 in exo:
                est_type.append(EST_SD_SHOCK); est_i.append(exo.index(e.name1)); est_j.append(exo.index(e.name1))
            else:
                o = varobs.index(e.name1)
                est_type.append(EST_SD_MEAS); est_i.append(o); est_j.append(o)
        else:
            if e.name1 in exo:
                est_type.append(EST_CORR_SHOCK); est_i.append(exo.index(e.name1)); est_j.append(exo.index(e.name2))
            else:
                est_type.append(EST_CORR_MEAS); est_i.append(varobs.index(e.name1)); est_j.append(varobs.index(e.name2))

    return ModelSpec(
        name=name, endo=endo, n_orig=n_orig, exo=exo, params=params, param_values=param_values,
        linear=False, lli=lli, i_bkwrd_b=i_bkwrd_b, i_fwrd_b=i_fwrd_b, i_both=i_both, i_static=i_static, i_dyn=i_dyn,
        sigma_e0=sigma_e0, sigma_m0=sigma_m0, varobs=varobs, obs_idx=obs_idx, state_ids=state_ids,
        obs_idx_state=obs_idx_state, lagged_state_ids=lagged_state_ids, est=est, est_type=est_type, est_i=est_i,
        est_j=est_j, estimation_options=estimation_options, psym=psym, ysym=ysym, ss_assign=ss_assign,
        static_resid=static_resid, jac=jac, ss_numeric=False, static_jac={},
    )


def build_model_library_from_preprocessor(name: str, modpath: str, out_dir: str) -> str:
    """``build.build_model_library`` for a model given by the preprocessor's output: device model function -> nvcc ->
    ``<out_dir>/libdynare_b200_model_<name>.so`` (+ ``<name>.json``); load it with ``engine.load_model_library``."""
    from . import build
    return build.build_model_library(name, None, out_dir, spec=spec_from_preprocessor(name, modpath))


if __name__ == "__main__":           # python -m dynare_jl_b200.preprocessor <name> <modpath> <out_dir>   (used by the Julia shim)
    import sys
    print(build_model_library_from_preprocessor(sys.argv[1], sys.argv[2], sys.argv[3]))
