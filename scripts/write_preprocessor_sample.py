#!/usr/bin/env python
"""TEST TOOLING: writes a model out in the layout and syntax of the Dynare preprocessor's output
(<dir>/model/json/modfile.json + <dir>/model/julia/*.jl, the files Dynare.jl's `parser` / `load_model_functions` read:
src/DynareParser.jl:10-15, 169-218; src/dynare_functions.jl:10-47), so that dynare_jl_b200.preprocessor can be exercised
where the preprocessor binary itself is not available (it is a JLL artifact, absent from the build image).

  python scripts/write_preprocessor_sample.py fs2000 /root/reference/test/models/fs2000/fs2000.mod tests/golden/preprocessor_fs2000

The model comes from this repo's .mod route (modfile.py + modelgen.py); what is written is the DYNAMIC model -- residuals
and the sparse Jacobian g1_v in CSC order over the n x (3n + nx) columns [y(-1) | y | y(+1) | x], 1-based references
y[j], x[j], params[j], temporary terms T[i] -- not the steady-state-evaluated form the engine consumes, so reading it
back is a real round trip."""
import json
import os
import sys

import numpy as np
import sympy as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dynare_jl_b200  # noqa: E402,F401
from dynare_jl_b200 import modelgen  # noqa: E402
from dynare_jl_b200.modfile import parse_mod  # noqa: E402

SHAPE_CODE = {"beta": 1, "gamma": 2, "normal": 3, "inv_gamma": 4, "uniform": 5, "inv_gamma2": 6, "weibull": 8}


def jl(expr, table):
    """sympy -> Julia source with the preprocessor's indexed references."""
    e = expr.xreplace(table)
    return sp.sstr(e).replace("**", "^")


def fn(name, args, body, asserts=()):
    lines = [f"function {name}({args})"]
    lines += [f"    @assert {a}" for a in asserts]
    lines += ["@inbounds begin"] + body + ["end", "    return nothing", "end", ""]
    return "\n".join(lines)


def main():
    name, mod, out = sys.argv[1:4]
    with open(mod) as f:
        mf = parse_mod(f.read())
    S = modelgen.build_spec(name, mf)
    n, nx = S.n, S.nx
    ym, yp, us = S.dyn_syms["lag"], S.dyn_syms["lead"], S.dyn_syms["exo"]
    tab = {}
    for j in range(n):
        tab[ym[j]] = sp.Symbol(f"y[{j + 1}]")
        tab[S.ysym[j]] = sp.Symbol(f"y[{n + j + 1}]")
        tab[yp[j]] = sp.Symbol(f"y[{2 * n + j + 1}]")
    for e in range(nx):
        tab[us[e]] = sp.Symbol(f"x[{e + 1}]")
    for q, p in enumerate(S.psym):
        tab[p] = sp.Symbol(f"params[{q + 1}]")
    stab = {S.ysym[j]: sp.Symbol(f"y[{j + 1}]") for j in range(n)}
    stab.update({us[e]: sp.Symbol(f"x[{e + 1}]") for e in range(nx)})
    stab.update({p: sp.Symbol(f"params[{q + 1}]") for q, p in enumerate(S.psym)})

    jdir = os.path.join(out, "model", "julia"); os.makedirs(jdir, exist_ok=True)
    os.makedirs(os.path.join(out, "model", "json"), exist_ok=True)
    V = "Vector{<: Real}"
    sig = "T::" + V + ", {0}y::" + V + ", x::" + V + ", params::" + V + ", steady_state::" + V
    sig = type("Sig", (), {"format": staticmethod(lambda extra, _s=sig: _s.replace("{0}", extra))})

    # ---- dynamic residuals and Jacobian (common sub-expressions become temporary terms) --------------------------------
    cols = ym + S.ysym + yp + us
    entries = []                                              # (col, row, expr) in CSC order
    for c, sym in enumerate(cols):
        for i, r in enumerate(S.dyn_resid):
            if sym in r.free_symbols:
                d = sp.diff(r, sym)
                if d != 0:
                    entries.append((c, i, d))
    repl, reduced = sp.cse([r for r in S.dyn_resid] + [d for _, _, d in entries], symbols=sp.numbered_symbols("TT"), order="none")
    ttab = dict(tab)
    tt_lines = []
    for q, (sym, ex) in enumerate(repl):
        tt_lines.append(f"    T[{q + 1}] = {jl(ex, ttab)};")
        ttab[sym] = sp.Symbol(f"T[{q + 1}]")
    nres = len(S.dyn_resid)
    open(os.path.join(jdir, "SparseDynamicResidTT!.jl"), "w").write(fn("SparseDynamicResidTT!", sig.format(""), tt_lines))
    open(os.path.join(jdir, "SparseDynamicG1TT!.jl"), "w").write(fn("SparseDynamicG1TT!", sig.format(""), []))
    body = [f"    residual[{i + 1}] = {jl(reduced[i], ttab)};" for i in range(nres)]
    open(os.path.join(jdir, "SparseDynamicResid!.jl"), "w").write(
        fn("SparseDynamicResid!", sig.format("residual::AbstractVector{<: Real}, "), body,
           [f"length(T) >= {len(repl)}", f"length(residual) == {n}", f"length(y) == {3 * n}", f"length(x) == {nx}",
            f"length(params) == {len(S.psym)}"]))
    body = [f"g1_v[{q + 1}]={jl(reduced[nres + q], ttab)};" for q in range(len(entries))]
    open(os.path.join(jdir, "SparseDynamicG1!.jl"), "w").write(
        fn("SparseDynamicG1!", sig.format("g1_v::Vector{<: Real}, "), body,
           [f"length(T) >= {len(repl)}", f"length(g1_v) == {len(entries)}", f"length(y) == {3 * n}", f"length(x) == {nx}"]))
    colptr = [1]
    for c in range(3 * n + nx):
        colptr.append(colptr[-1] + sum(1 for cc, _, _ in entries if cc == c))
    rowval = [i + 1 for _, i, _ in entries]
    colval = [c + 1 for c, _, _ in entries]

    # ---- static residuals ---------------------------------------------------------------------------------------------------
    ssig = "T::" + V + ", {0}y::" + V + ", x::" + V + ", params::" + V
    ssig = type("Sig", (), {"format": staticmethod(lambda extra, _s=ssig: _s.replace("{0}", extra))})
    open(os.path.join(jdir, "SparseStaticResidTT!.jl"), "w").write(fn("SparseStaticResidTT!", ssig.format(""), []))
    body = [f"    residual[{i + 1}] = {jl(r, stab)};" for i, r in enumerate(S.static_resid)]
    open(os.path.join(jdir, "SparseStaticResid!.jl"), "w").write(
        fn("SparseStaticResid!", ssig.format("residual::AbstractVector{<: Real}, "), body))

    # ---- steady state: ys_[i] assignments, locals of the block keep their names -------------------------------------------------
    ytab = {S.ysym[j]: sp.Symbol(f"ys_[{j + 1}]") for j in range(n)}
    ytab.update({p: sp.Symbol(f"params[{q + 1}]") for q, p in enumerate(S.psym)})
    body = []
    for sym, ex in S.ss_assign:
        if sym in S.ysym:
            if S.endo.index(sym.name) >= S.n_orig:
                continue                                       # auxiliary variables: SetAuxiliaryVariables.jl
            body.append(f"    ys_[{S.ysym.index(sym) + 1}]={jl(ex, ytab)};")
        else:
            local = sym.name[3:] if sym.name.startswith("ss_") else sym.name
            ytab[sym] = sp.Symbol(local)
            body.append(f"    {local}={jl(ex, ytab)};")
    open(os.path.join(jdir, "SteadyState2.jl"), "w").write(
        "function steady_state!(ys_::Vector{<: Real}, exo_::Vector{<: Real}, params::Vector{<: Real})\n@inbounds begin\n"
        + "\n".join(body) + "\nend\nend\n")
    body = []
    for sym, ex in S.ss_assign:
        if sym in S.ysym and S.endo.index(sym.name) >= S.n_orig:
            body.append(f"    y[{S.ysym.index(sym) + 1}]={jl(ex, stab)};")
    open(os.path.join(jdir, "SetAuxiliaryVariables.jl"), "w").write(
        "function set_auxiliary_variables!(y::Vector{<: Real}, x::Vector{<: Real}, params::Vector{<: Real})\n@inbounds begin\n"
        + "\n".join(body) + "\nend\nend\n")

    # ---- modfile.json ------------------------------------------------------------------------------------------------------------
    sym_entry = lambda v: {"name": v, "texName": v, "longName": v}
    statements = [{"statementName": "param_init", "name": p, "value": repr(float(v))}
                  for p, v in zip(S.params, S.param_values) if not np.isnan(v)]
    shocks = {"statementName": "shocks", "overwrite": False, "variance": [], "stderr": [], "covariance": [], "correlation": []}
    for e, u in enumerate(S.exo):
        if S.sigma_e0[e, e] != 0:
            shocks["variance"].append({"name": u, "variance": repr(float(S.sigma_e0[e, e]))})
    for a in range(nx):
        for b in range(a + 1, nx):
            if S.sigma_e0[a, b] != 0:
                shocks["covariance"].append({"name": S.exo[a], "name2": S.exo[b], "covariance": repr(float(S.sigma_e0[a, b]))})
    for i, v in enumerate(S.varobs):
        if S.sigma_m0[i, i] != 0:
            shocks["variance"].append({"name": v, "variance": repr(float(S.sigma_m0[i, i]))})
    statements.append(shocks)
    statements.append({"statementName": "varobs", "vars": list(S.varobs)})
    rows = []
    for e in S.est:
        row = {"init_val": "NaN" if e.init is None else repr(float(e.init)), "lower_bound": "(-Inf)", "upper_bound": "Inf",
               "prior_distribution": SHAPE_CODE[e.shape], "mean": repr(float(e.mean)),
               "std": "Inf" if np.isinf(e.stdev) else repr(float(e.stdev)), "p3": "NaN", "p4": "NaN", "jscale": "NaN"}
        if e.domain is not None:
            row.update(mean="NaN", std="NaN", p3=repr(float(e.domain[0])), p4=repr(float(e.domain[1])))
        if e.kind == "param":
            row = {"param": e.name1, **row}
        elif e.kind == "corr":
            row = {"var1": e.name1, "var2": e.name2, **row}
        else:
            row = {"var": e.name1, **row}
        rows.append(row)
    if rows:
        statements.append({"statementName": "estimated_params", "params": rows})
    if S.estimation_options:
        statements.append({"statementName": "estimation", "options": dict(S.estimation_options)})
    lli = S.lli.astype(int)
    counter = 0
    inc = np.zeros((n, 3), dtype=int)
    for t in range(3):
        for j in range(n):
            if lli[t, j]:
                counter += 1
                inc[j, t] = counter
    mj = {
        "endogenous": [sym_entry(v) for v in S.endo], "exogenous": [sym_entry(v) for v in S.exo],
        "exogenous_deterministic": [], "parameters": [sym_entry(v) for v in S.params],
        "orig_endo_nbr": S.n_orig,
        "aux_vars": [{"endo_index": j + 1, "type": 0 if "LEAD" in S.endo[j] else 1} for j in range(S.n_orig, n)],
        "steady_state_model": not S.ss_numeric, "varobs": list(S.varobs),
        "model_info": {"lead_lag_incidence": inc.tolist(), "nstatic": S.s, "nfwrd": S.nf - len(S.i_both),
                       "npred": S.k - len(S.i_both), "nboth": len(S.i_both), "nsfwrd": S.nf, "nspred": S.k,
                       "ndynamic": n - S.s, "maximum_endo_lag": 1, "maximum_endo_lead": 1, "maximum_exo_lag": 0,
                       "maximum_exo_lead": 0, "maximum_exo_det_lag": 0, "maximum_exo_det_lead": 0, "maximum_lag": 1,
                       "maximum_lead": 1, "NNZDerivatives": [len(entries), -1, -1]},
        "dynamic_g1_sparse_rowval": rowval, "dynamic_g1_sparse_colval": colval, "dynamic_g1_sparse_colptr": colptr,
        "dynamic_tmp_nbr": [len(repl), 0, 0, 0], "static_tmp_nbr": [0, 0, 0, 0],
        "statements": statements,
    }
    with open(os.path.join(out, "model", "json", "modfile.json"), "w") as f:
        json.dump(mj, f, indent=0)
    print(f"{name}: {n} endogenous, {len(entries)} Jacobian non-zeros, {len(repl)} temporary terms -> {out}")


if __name__ == "__main__":
    main()
