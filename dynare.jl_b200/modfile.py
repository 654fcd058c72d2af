"""Reader for the subset of the Dynare ``.mod`` language the likelihood path needs.

The reference hands ``.mod`` files to an external C++ preprocessor binary
(``DynarePreprocessor_jll``; reference ``src/DynarePreprocessor.jl:4-40``) and reads
the JSON it produces (``src/DynareParser.jl:169-218``).  That binary is not part of
the reference tree and is not available here, so this module reads the declarative
blocks directly.  It is a host-side set-up tool (runs once per model, never per draw).

Supported: ``var``, ``varexo``, ``parameters``, parameter calibration statements,
``model;`` / ``model(linear);`` with model-local variables (``# name = expr;``), ``steady_state_model;``,
``shocks;`` (``var x; stderr s;``, ``var x = v;``, ``var x, y = c;``, ``corr a, b = r;``, on shocks and on observed
variables = calibrated measurement errors), ``varobs``, ``estimated_params;``, ``estimated_params_init;``,
``x.prior(shape=..., mean=..., stdev=...)`` statements, the reference's Julia-syntax
``prior!(:x | stdev(:x) | variance(:x) | corr(:a, :b), shape=..., mean=..., stdev=... | variance=..., domain=[a, b])``
(``src/estimation/priors.jl:64-178``; e.g. ``test/models/estimation/nk1.mod:52-75``), ``verbatim`` blocks (skipped) and the
``nobs`` / ``first_obs`` / ``datafile`` options of ``estimation(...)`` (also when the
statement is commented out with ``//``, which is how the reference's fixtures carry it:
``test/models/ls2003/ls2003.mod:85``).
"""
from __future__ import annotations

import ast
import math
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple


@dataclass
class EstimatedParam:
    """One row of ``estimated_params`` (reference ``src/estimation/priors.jl:370-433``)."""
    kind: str                 # 'param' | 'stderr' | 'corr'
    name1: str
    name2: Optional[str]
    shape: str                # 'beta' 'gamma' 'normal' 'inv_gamma' 'uniform' 'inv_gamma2' 'weibull'
    mean: float
    stdev: float
    init: Optional[float] = None
    domain: Optional[Tuple[float, float]] = None


@dataclass
class ModFile:
    endo: List[str] = field(default_factory=list)
    exo: List[str] = field(default_factory=list)
    params: List[str] = field(default_factory=list)
    calibration: List[Tuple[str, str]] = field(default_factory=list)   # (name, expr) in file order
    equations: List[Tuple[str, str]] = field(default_factory=list)     # (lhs, rhs); rhs '0' if none
    model_locals: List[Tuple[str, str]] = field(default_factory=list)  # `# name = expr;` inside the model block
    linear: bool = False
    steady_state_model: List[Tuple[str, str]] = field(default_factory=list)
    initval: List[Tuple[str, str]] = field(default_factory=list)        # (name, expr): starting point of `steady`
    shock_stderr: Dict[str, str] = field(default_factory=dict)
    shock_var: Dict[str, str] = field(default_factory=dict)
    shock_corr: Dict[Tuple[str, str], str] = field(default_factory=dict)
    shock_cov: Dict[Tuple[str, str], str] = field(default_factory=dict)
    varobs: List[str] = field(default_factory=list)
    estimated_params: List[EstimatedParam] = field(default_factory=list)
    estimation_options: Dict[str, str] = field(default_factory=dict)


_SHAPES = {
    "beta_pdf": "beta", "beta": "beta",
    "gamma_pdf": "gamma", "gamma": "gamma",
    "normal_pdf": "normal", "normal": "normal",
    "inv_gamma_pdf": "inv_gamma", "inv_gamma1_pdf": "inv_gamma", "inv_gamma": "inv_gamma",
    "inv_gamma1": "inv_gamma",
    "uniform_pdf": "uniform", "uniform": "uniform",
    "inv_gamma2_pdf": "inv_gamma2", "inv_gamma2": "inv_gamma2",
    "weibull_pdf": "weibull", "weibull": "weibull",
    # Julia-syntax names of prior!(..., shape=...) (src/estimation/priors.jl:147-178)
    "inversegamma1": "inv_gamma", "inversegamma2": "inv_gamma2", "inversegamma": "inv_gamma2",
}


def _split_top(s: str) -> List[str]:
    """Split on commas that are not inside (), [] or {}."""
    out, depth, cur = [], 0, []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append("".join(cur)); cur = []
        else:
            cur.append(ch)
    out.append("".join(cur))
    return [x.strip() for x in out if x.strip()]


def _parse_prior_bang(stmt: str) -> Optional[EstimatedParam]:
    """``prior!(:x, shape=Gamma, mean=0.5, stdev=0.5)`` and its stdev(:x) / variance(:x) / corr(:a, :b) forms."""
    m = re.match(r"prior!\s*\((.*)\)\s*$", stmt, flags=re.S)
    if not m:
        return None
    args = _split_top(m.group(1))
    head = args[0].replace(" ", "")
    mm = re.match(r"^:(\w+)$", head)
    if mm:
        kind, n1, n2 = "param", mm.group(1), None
    else:
        mm = re.match(r"^(stdev|variance)\(:(\w+)\)$", head)
        if mm:
            kind, n1, n2 = ("stderr" if mm.group(1) == "stdev" else "var"), mm.group(2), None
        else:
            mm = re.match(r"^corr\(:(\w+),:(\w+)\)$", head)
            if not mm:
                raise ValueError(f"unsupported prior! target: {args[0]!r}")
            kind, n1, n2 = "corr", mm.group(1), mm.group(2)
    kw = {}
    for item in args[1:]:
        k, v = item.split("=", 1)
        kw[k.strip()] = v.strip()
    shape = _SHAPES[kw["shape"].lower()]
    init = _num(kw["initialvalue"]) if "initialvalue" in kw else None
    dom = None
    if "domain" in kw and kw["domain"].strip() not in ("[]", ""):
        lo, hi = _split_top(kw["domain"].strip()[1:-1])
        dom = (_num(lo), _num(hi))
    # mean / stdev are optional (`missing` in the reference) when a Uniform prior is given by its domain
    # (src/estimation/priors.jl:165-170): the moments of U(a, b) stand in for them
    if "mean" in kw:
        mean = _num(kw["mean"])
    elif shape == "uniform" and dom is not None:
        mean = 0.5 * (dom[0] + dom[1])
    else:
        raise ValueError(f"prior!({args[0]}): `mean` is required for shape {shape}")
    if "stdev" in kw:
        stdev = _num(kw["stdev"])
    elif "variance" in kw:
        stdev = _num(kw["variance"]) ** 0.5
    elif shape == "uniform" and dom is not None:
        stdev = (dom[1] - dom[0]) / 12 ** 0.5
    else:
        raise ValueError(f"prior!({args[0]}): `stdev` or `variance` is required for shape {shape}")
    if dom is not None and shape != "uniform":
        dom = None                                   # the reference reads `domain` for the Uniform shape only
    return EstimatedParam(kind, n1, n2, shape, mean, stdev, init, dom)


def _num(s: str) -> float:
    s = s.strip().lower()
    if s in ("inf", "+inf"):
        return float("inf")
    if s == "-inf":
        return float("-inf")
    return float(_eval_numeric(ast.parse(s.replace("^", "**"), mode="eval").body))


_NUM_FUNCS = {"sqrt": math.sqrt, "exp": math.exp, "log": math.log}
_NUM_BINOPS = {ast.Add: lambda a, b: a + b, ast.Sub: lambda a, b: a - b, ast.Mult: lambda a, b: a * b,
               ast.Div: lambda a, b: a / b, ast.Pow: lambda a, b: a ** b}


def _eval_numeric(node) -> float:
    """Numeric literals of a mod file (e.g. ``(0.20*0.579923)``, ``1.25^2``, ``sqrt(2)``): a whitelist walk over the
    syntax tree -- numbers, + - * / ^, unary sign, sqrt / exp / log, pi -- never ``eval``."""
    if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)) and not isinstance(node.value, bool):
        return float(node.value)
    if isinstance(node, ast.BinOp) and type(node.op) in _NUM_BINOPS:
        return _NUM_BINOPS[type(node.op)](_eval_numeric(node.left), _eval_numeric(node.right))
    if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
        v = _eval_numeric(node.operand)
        return -v if isinstance(node.op, ast.USub) else v
    if isinstance(node, ast.Name) and node.id == "pi":
        return math.pi
    if (isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id in _NUM_FUNCS
            and len(node.args) == 1 and not node.keywords):
        return _NUM_FUNCS[node.func.id](_eval_numeric(node.args[0]))
    raise ValueError("unsupported numeric expression in the mod file")


def _strip_comments(text: str) -> Tuple[str, List[str]]:
    """Return the text without comments and the list of ``//`` comment bodies."""
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    comments = []
    out = []
    for line in text.splitlines():
        m = re.search(r"//|%", line)
        if m:
            comments.append(line[m.end():])
            line = line[: m.start()]
        out.append(line)
    return "\n".join(out), comments


def _names(body: str) -> List[str]:
    body = re.sub(r"\$[^$]*\$", " ", body)
    body = re.sub(r"\([^)]*\)", " ", body)
    return [t for t in re.split(r"[\s,]+", body.strip()) if t]


def _parse_estimation_options(stmt: str, opts: Dict[str, str]) -> None:
    m = re.match(r"\s*estimation\s*\((.*)\)", stmt, flags=re.S)
    if not m:
        return
    for item in re.split(r",(?![^()]*\))", m.group(1)):
        if "=" in item:
            k, v = item.split("=", 1)
            opts[k.strip()] = v.strip().strip("'\"")
        elif item.strip():
            opts[item.strip()] = "true"


def _extract_bang_calls(text: str) -> Tuple[str, List[str]]:
    """Pull the reference's Julia-syntax statements ``name!( ... )`` (balanced parentheses, no terminating ``;``
    required) out of the text; returns (remaining text, calls)."""
    calls, out, pos = [], [], 0
    for m in re.finditer(r"\b\w+!\s*\(", text):
        if m.start() < pos:
            continue
        depth, i = 0, m.end() - 1
        while i < len(text):
            if text[i] == "(":
                depth += 1
            elif text[i] == ")":
                depth -= 1
                if depth == 0:
                    break
            i += 1
        calls.append(" ".join(text[m.start():i + 1].split()))
        out.append(text[pos:m.start()])
        pos = i + 1
        while pos < len(text) and text[pos] in " \t":
            pos += 1
        if pos < len(text) and text[pos] == ";":
            pos += 1
    out.append(text[pos:])
    return "".join(out), calls


def parse_mod(text: str) -> ModFile:
    mf = ModFile()
    clean, comments = _strip_comments(text)
    clean, bang_calls = _extract_bang_calls(clean)
    for c in bang_calls:
        if c.startswith("prior!"):
            ep = _parse_prior_bang(c)
            if ep is not None:
                mf.estimated_params.append(ep)
    # an estimation statement that only exists as a comment still defines the sample
    for c in comments:
        if re.match(r"\s*estimation\s*\(", c):
            _parse_estimation_options(c.strip().rstrip(";"), mf.estimation_options)
    clean = re.sub(r"(?m)^\s*using\s+[\w., ]+$", "", clean)          # Julia statements of the reference front end
    stmts = [s.strip() for s in clean.split(";")]
    block = None
    pending_shock: Optional[str] = None
    for s in stmts:
        if not s:
            continue
        flat = " ".join(s.split())
        if block is not None:
            if flat == "end":
                block = None
                pending_shock = None
                continue
            if block == "verbatim":
                continue
            if block == "model":
                if flat.startswith("#"):
                    lhs, rhs = flat[1:].split("=", 1)
                    mf.model_locals.append((lhs.strip(), rhs.strip()))
                    continue
                if "=" in flat:
                    lhs, rhs = flat.split("=", 1)
                else:
                    lhs, rhs = flat, "0"
                # drop equation tags  [name='...']
                lhs = re.sub(r"^\[[^\]]*\]\s*", "", lhs)
                mf.equations.append((lhs.strip(), rhs.strip()))
            elif block == "steady_state_model":
                lhs, rhs = flat.split("=", 1)
                mf.steady_state_model.append((lhs.strip(), rhs.strip()))
            elif block == "shocks":
                m = re.match(r"var\s+(\w+)\s*,\s*(\w+)\s*=\s*(.+)$", flat)
                if m:
                    mf.shock_cov[(m.group(1), m.group(2))] = m.group(3)
                    continue
                m = re.match(r"var\s+(\w+)\s*=\s*(.+)$", flat)
                if m:
                    mf.shock_var[m.group(1)] = m.group(2)
                    continue
                m = re.match(r"var\s+(\w+)$", flat)
                if m:
                    pending_shock = m.group(1)
                    continue
                m = re.match(r"stderr\b\s*(.+)$", flat)
                if m and pending_shock:
                    mf.shock_stderr[pending_shock] = m.group(1)
                    pending_shock = None
                    continue
                m = re.match(r"corr\s+(\w+)\s*,\s*(\w+)\s*=\s*(.+)$", flat)
                if m:
                    mf.shock_corr[(m.group(1), m.group(2))] = m.group(3)
                    continue
                if re.match(r"(periods|values)\b", flat):
                    pending_shock = None
                    continue
                raise ValueError(f"unsupported shocks statement: {flat!r}")
            elif block == "estimated_params":
                parts = [p.strip() for p in flat.split(",")]
                head = parts[0].split()
                if head[0] == "stderr":
                    kind, n1, n2 = "stderr", head[1], None
                    rest = parts[1:]
                elif head[0] == "corr":
                    kind, n1, n2 = "corr", head[1], parts[1]
                    rest = parts[2:]
                else:
                    kind, n1, n2 = "param", head[0], None
                    rest = parts[1:]
                # forms:  shape, mean, std [, p3, p4 [, scale]]  |  init, shape, ...  |  init, lb, ub, shape, ...
                # lb / ub are optimiser bounds the reference does not read (src/estimation/priors.jl:411-424); a uniform
                # prior with empty mean and std is U(p3, p4) (parse_prior_distribution(::Val{5}), priors.jl:508-517)
                si = next(i for i, p in enumerate(rest) if p.lower() in _SHAPES)
                init = _num(rest[0]) if si >= 1 and rest[0] else None
                shape = _SHAPES[rest[si].lower()]
                tail = rest[si + 1:] + [""] * 4
                dom = None
                if shape == "uniform" and not tail[0] and not tail[1]:
                    if not (tail[2] and tail[3]):
                        raise ValueError(f"uniform prior of {n1}: give mean and std, or p3 and p4")
                    dom = (_num(tail[2]), _num(tail[3]))
                    mean, std = 0.5 * (dom[0] + dom[1]), (dom[1] - dom[0]) / 12 ** 0.5
                else:
                    mean, std = _num(tail[0]), _num(tail[1])
                mf.estimated_params.append(EstimatedParam(kind, n1, n2, shape, mean, std, init, dom))
            elif block == "estimated_params_init":
                parts = [p.strip() for p in flat.split(",")]
                head = parts[0].split()
                if head[0] == "stderr":
                    key = ("stderr", head[1], None)
                elif head[0] == "corr":
                    key = ("corr", head[1], parts[1])
                    parts = parts[1:]
                else:
                    key = ("param", head[0], None)
                for ep in mf.estimated_params:
                    if (ep.kind, ep.name1, ep.name2) == key:
                        ep.init = _num(parts[1])
            elif block == "initval":
                if "=" in flat:
                    lhs, rhs = flat.split("=", 1)
                    mf.initval.append((lhs.strip(), rhs.strip()))
            # other blocks (histval, endval, ...) are ignored
            continue

        m = re.match(r"(model|steady_state_model|shocks|estimated_params|estimated_params_init|"
                     r"initval|endval|histval|observation_trends|estimated_params_bounds|verbatim)"
                     r"\s*(\(([^)]*)\))?$", flat)
        if m:
            block = m.group(1)
            if block == "model" and m.group(3) and "linear" in m.group(3):
                mf.linear = True
            continue
        m = re.match(r"(var|varexo|parameters|varobs)\s+(.*)$", flat, flags=re.S)
        if m:
            target = {"var": mf.endo, "varexo": mf.exo, "parameters": mf.params,
                      "varobs": mf.varobs}[m.group(1)]
            target.extend(_names(m.group(2)))
            continue
        m = re.match(r"(std\s*\(\s*(\w+)\s*\)|corr\s*\(\s*(\w+)\s*,\s*(\w+)\s*\)|(\w+))\.prior\s*\((.*)\)$", flat)
        if m:
            kw = {}
            for item in _split_top(m.group(6)):
                if "=" in item:
                    k, v = item.split("=", 1)
                    kw[k.strip()] = v.strip()
            if m.group(2):
                kind, n1, n2 = "stderr", m.group(2), None
            elif m.group(3):
                kind, n1, n2 = "corr", m.group(3), m.group(4)
            else:
                kind, n1, n2 = "param", m.group(5), None
            stdev = _num(kw["stdev"]) if "stdev" in kw else _num(kw["variance"]) ** 0.5
            mf.estimated_params.append(
                EstimatedParam(kind, n1, n2, _SHAPES[kw["shape"].lower()], _num(kw["mean"]), stdev))
            continue
        if flat.startswith("prior!"):
            ep = _parse_prior_bang(flat)
            if ep is not None:
                mf.estimated_params.append(ep)
            continue
        m = re.match(r"(\w+)\s*=\s*(.+)$", flat)
        if m and m.group(1) in mf.params:
            mf.calibration.append((m.group(1), m.group(2)))
            continue
        if flat.startswith("estimation"):
            _parse_estimation_options(flat, mf.estimation_options)
            continue
        # steady, check, stoch_simul, ... : statements of the reference front end, not needed
    return mf
