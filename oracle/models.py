"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product path.

Hand-written model definitions for the two reference fixtures, transcribed equation by equation
from the model text (reference ``test/models/fs2000/fs2000.mod:14-55`` and
``test/models/ls2003/ls2003.mod:20-32``), *independently* of ``dynare_jl_b200.modelgen``.
The dynamic Jacobian is obtained by complex-step differentiation of these residual functions
(exact to rounding), so it is a genuine second opinion on the symbolic code generator.

Parity: UNPINNED by the reference (it stores no Jacobian, policy matrix or likelihood value;
SURVEY.md section 8c).  The known answers pinned here are mathematical identities.
"""
from __future__ import annotations

import math

import numpy as np


class HandModel:
    """Container mirroring the bookkeeping of the reference ``Model``
    (``src/dynare_containers.jl:94-188``): names, index sets (0-based), calibration."""

    name: str
    endo: list
    exo: list
    params: list
    params0: np.ndarray
    sigma_e0: np.ndarray
    varobs: list
    lagged: list          # names appearing with a lag
    led: list             # names appearing with a lead
    est: list             # (type, name[, name2]) in estimated_params order
    linear = False

    def __init__(self):
        self.n = len(self.endo)
        self.nx = len(self.exo)
        ix = {v: i for i, v in enumerate(self.endo)}
        self.i_bkwrd_b = sorted(ix[v] for v in self.lagged)
        self.i_fwrd_b = sorted(ix[v] for v in self.led)
        self.i_static = [i for i in range(self.n)
                         if i not in self.i_bkwrd_b and i not in self.i_fwrd_b]
        self.i_dyn = [i for i in range(self.n) if i not in self.i_static]
        self.k = len(self.i_bkwrd_b)
        self.nf = len(self.i_fwrd_b)
        self.obs_idx = [ix[v] for v in self.varobs]
        # reference src/estimation/estimation.jl:384-390
        self.state_ids = sorted(set(self.obs_idx) | set(self.i_bkwrd_b))
        self.obs_idx_state = [self.state_ids.index(i) for i in self.obs_idx]
        self.lagged_state_ids = [p for p, i in enumerate(self.state_ids) if i in self.i_bkwrd_b]
        self.ns = len(self.state_ids)
        self.ny = len(self.varobs)

    # --- to be provided by the model -------------------------------------------------------
    def steady_state(self, p):
        raise NotImplementedError

    def dynamic_resid(self, ym, y, yp, u, p):
        raise NotImplementedError

    # --- derived -----------------------------------------------------------------------------
    def static_resid(self, ybar, p):
        return np.real(self.dynamic_resid(ybar, ybar, ybar, np.zeros(self.nx), p))

    def full_jacobian(self, ybar, p):
        """n x (3n + nx), column blocks [y(-1) | y | y(+1) | u]  (reference ``src/model.jl:24-26``)
        at y3 = [ybar; ybar; ybar], u = 0 (``src/perturbations.jl:526-531``), by complex step."""
        n, nx = self.n, self.nx
        h = 1e-30
        J = np.zeros((n, 3 * n + nx))
        base = [np.array(ybar, dtype=complex) for _ in range(3)] + [np.zeros(nx, dtype=complex)]
        col = 0
        for b in range(4):
            for j in range(len(base[b])):
                args = [a.copy() for a in base]
                args[b][j] += 1j * h
                J[:, col] = np.imag(self.dynamic_resid(*args, p)) / h
                col += 1
        return J

    def abcf(self, ybar, p):
        """(A, B, C, Fu) = d/dy(+1), d/dy, d/dy(-1), d/du, each with n columns (resp. nx)."""
        n = self.n
        J = self.full_jacobian(ybar, p)
        return J[:, 2 * n:3 * n], J[:, n:2 * n], J[:, :n], J[:, 3 * n:]

    def compressed_jacobian(self, ybar, p):
        """Reference ``set_compressed_jacobian!`` (``src/perturbations.jl:616-635``)."""
        A, B, C, Fu = self.abcf(ybar, p)
        return np.hstack([C[:, self.i_bkwrd_b], B, A[:, self.i_fwrd_b], Fu])


class FS2000(HandModel):
    name = "fs2000"
    endo = ["m", "P", "c", "e", "W", "R", "k", "d", "n", "l", "gy_obs", "gp_obs", "y", "dA",
            "AUX_ENDO_LEAD_c_1", "AUX_ENDO_LEAD_P_1"]
    exo = ["e_a", "e_m"]
    params = ["alp", "bet", "gam", "mst", "rho", "psi", "del"]
    params0 = np.array([0.33, 0.99, 0.003, 1.011, 0.7, 0.787, 0.02])
    sigma_e0 = np.diag([0.014 ** 2, 0.005 ** 2])
    varobs = ["gp_obs", "gy_obs"]
    lagged = ["m", "P", "k", "y"]
    led = ["m", "P", "c", "e", "n", "AUX_ENDO_LEAD_c_1", "AUX_ENDO_LEAD_P_1"]
    est = [("param", "alp"), ("param", "bet"), ("param", "gam"), ("param", "mst"),
           ("param", "rho"), ("param", "psi"), ("param", "del"),
           ("stderr", "e_a"), ("stderr", "e_m")]
    # test/models/estimation/fs2000.mod:81-91 (estimated_params_init), used as the "mode" stand-in
    theta_mode = np.array([0.4186, 0.9901, 0.0038, 1.0141, 0.8623, 0.6837, 0.0020, 0.0133, 0.0029])
    priors = [("beta", 0.356, 0.02), ("beta", 0.993, 0.002), ("normal", 0.0085, 0.003),
              ("normal", 1.0002, 0.007), ("beta", 0.129, 0.223), ("beta", 0.65, 0.05),
              ("beta", 0.01, 0.005), ("inv_gamma", 0.035449, np.inf), ("inv_gamma", 0.008862, np.inf)]

    def steady_state(self, p):
        alp, bet, gam, mst, rho, psi, dlt = p
        with np.errstate(all="ignore"):
            dA = np.exp(gam)
            gst = 1 / dA
            m = mst
            khst = ((1 - gst * bet * (1 - dlt)) / (alp * gst ** alp * bet)) ** (1 / (alp - 1))
            xist = (((khst * gst) ** alp - (1 - gst * (1 - dlt)) * khst) / mst) ** (-1)
            nust = psi * mst ** 2 / ((1 - alp) * (1 - psi) * bet * gst ** alp * khst ** alp)
            n = xist / (nust + xist)
            P = xist + nust
            k = khst * n
            l = psi * mst * n / ((1 - psi) * (1 - n))
            c = mst / P
            d = l - mst + 1
            y = k ** alp * n ** (1 - alp) * gst ** alp
            R = mst / bet
            W = l / n
            e = 1.0
            gp_obs = m / dA - 1
            gy_obs = dA - 1
        return np.array([m, P, c, e, W, R, k, d, n, l, gy_obs, gp_obs, y, dA, c, P], dtype=float)

    def dynamic_resid(self, ym, y, yp, u, p):
        alp, bet, gam, mst, rho, psi, dlt = p
        m, P, c, e, W, R, k, d, n, l, gy_obs, gp_obs, yy, dA, a1, a2 = y
        m_1, P_1, k_1, y_1 = ym[0], ym[1], ym[6], ym[12]
        m1, P1, c1, e1, n1, a1_1, a2_1 = yp[0], yp[1], yp[2], yp[3], yp[8], yp[14], yp[15]
        e_a, e_m = u
        exp, log = np.exp, np.log
        r = [None] * 16
        r[0] = dA - exp(gam + e_a)
        r[1] = log(m) - ((1 - rho) * log(mst) + rho * log(m_1) + e_m)
        r[2] = (-P / (c1 * P1 * m)
                + bet * P1 * (alp * exp(-alp * (gam + log(e1))) * k ** (alp - 1) * n1 ** (1 - alp)
                              + (1 - dlt) * exp(-(gam + log(e1)))) / (a1_1 * a2_1 * m1))
        r[3] = W - l / n
        r[4] = -(psi / (1 - psi)) * (c * P / (1 - n)) + l / n
        r[5] = R - P * (1 - alp) * exp(-alp * (gam + e_a)) * k_1 ** alp * n ** (-alp) / W
        r[6] = (1 / (c * P)
                - bet * P * (1 - alp) * exp(-alp * (gam + e_a)) * k_1 ** alp * n ** (1 - alp)
                / (m * l * c1 * P1))
        r[7] = c + k - (exp(-alp * (gam + e_a)) * k_1 ** alp * n ** (1 - alp)
                        + (1 - dlt) * exp(-(gam + e_a)) * k_1)
        r[8] = P * c - m
        r[9] = m - 1 + d - l
        r[10] = e - exp(e_a)
        r[11] = yy - k_1 ** alp * n ** (1 - alp) * exp(-alp * (gam + e_a))
        r[12] = gy_obs - (dA * yy / y_1 - 1)
        r[13] = gp_obs - ((P / P_1) * m_1 / dA - 1)
        r[14] = a1 - c1
        r[15] = a2 - P1
        return np.array(r)


class LS2003(HandModel):
    name = "ls2003"
    linear = True
    endo = ["y", "y_s", "R", "pie", "dq", "pie_s", "de", "A", "y_obs", "pie_obs", "R_obs"]
    exo = ["e_R", "e_q", "e_ys", "e_pies", "e_A"]
    params = ["psi1", "psi2", "psi3", "rho_R", "tau", "alpha", "rr", "k", "rho_q", "rho_A",
              "rho_ys", "rho_pies"]
    params0 = np.array([0.54, 0.25, 0.25, 0.5, 0.5, 0.3, 2.51, 0.5, 0.4, 0.2, 0.9, 0.7])
    sigma_e0 = np.diag([1.25 ** 2, 2.5 ** 2, 1.89, 1.89, 1.89])
    varobs = ["y_obs", "R_obs", "pie_obs", "dq", "de"]
    lagged = ["y", "y_s", "R", "dq", "pie_s", "A"]
    led = ["y", "y_s", "pie", "dq", "A"]
    est = [("param", "psi1"), ("param", "psi2"), ("param", "psi3"), ("param", "rho_R"),
           ("param", "alpha"), ("param", "rr"), ("param", "k"), ("param", "tau"),
           ("param", "rho_q"), ("param", "rho_A"), ("param", "rho_ys"), ("param", "rho_pies"),
           ("stderr", "e_R"), ("stderr", "e_q"), ("stderr", "e_A"), ("stderr", "e_ys"),
           ("stderr", "e_pies")]
    priors = [("gamma", 0.5, 0.5), ("gamma", 0.25, 0.125), ("gamma", 0.25, 0.125),
              ("beta", 0.5, 0.2), ("beta", 0.3, 0.1), ("gamma", 2.5, 1.0), ("gamma", 0.5, 0.25),
              ("gamma", 0.5, 0.2), ("beta", 0.4, 0.2), ("beta", 0.5, 0.2), ("beta", 0.8, 0.1),
              ("beta", 0.7, 0.15), ("inv_gamma", 1.2533, 0.6551), ("inv_gamma", 2.5066, 1.3103),
              ("inv_gamma", 1.2533, 0.6551), ("inv_gamma", 1.2533, 0.6551),
              ("inv_gamma", 1.88, 0.9827)]

    @property
    def theta_mode(self):
        return np.array([0.54, 0.25, 0.25, 0.5, 0.3, 2.51, 0.5, 0.5, 0.4, 0.2, 0.9, 0.7,
                         1.25, 2.5, np.sqrt(1.89), np.sqrt(1.89), np.sqrt(1.89)])

    def steady_state(self, p):
        return np.zeros(11)

    def dynamic_resid(self, ym, y_, yp, u, p):
        psi1, psi2, psi3, rho_R, tau, alpha, rr, k, rho_q, rho_A, rho_ys, rho_pies = p
        y, y_s, R, pie, dq, pie_s, de, A, y_obs, pie_obs, R_obs = y_
        y_1, y_s_1, R_1, dq_1, pie_s_1, A_1 = ym[0], ym[1], ym[2], ym[4], ym[5], ym[7]
        y1, y_s1, pie1, dq1, A1 = yp[0], yp[1], yp[3], yp[4], yp[7]
        e_R, e_q, e_ys, e_pies, e_A = u
        exp = np.exp
        om = tau + alpha * (2 - alpha) * (1 - tau)
        r = [None] * 11
        r[0] = y - (y1 - om * (R - pie1) - alpha * om * dq1
                    + alpha * (2 - alpha) * ((1 - tau) / tau) * (y_s - y_s1) - A1)
        r[1] = pie - (exp(-rr / 400) * pie1 + alpha * exp(-rr / 400) * dq1 - alpha * dq
                      + (k / om) * y + alpha * (2 - alpha) * (1 - tau) / (tau * om) * y_s)
        r[2] = pie - (de + (1 - alpha) * dq + pie_s)
        r[3] = R - (rho_R * R_1 + (1 - rho_R) * ((1 + psi1) * pie
                                                 + psi2 * (y + alpha * (2 - alpha) * ((1 - tau) / tau) * y_s)
                                                 + psi3 * de) + e_R)
        r[4] = dq - (rho_q * dq_1 + e_q)
        r[5] = y_s - (rho_ys * y_s_1 + e_ys)
        r[6] = pie_s - (rho_pies * pie_s_1 + e_pies)
        r[7] = A - (rho_A * A_1 + e_A)
        r[8] = y_obs - (y - y_1 + A)
        r[9] = pie_obs - 4 * pie
        r[10] = R_obs - 4 * R
        return np.array(r)


class HSNK(HandModel):
    """Herbst-Schorfheide small New Keynesian model as the reference carries it in
    ``test/models/estimation/test1.mod`` (model :26-36, steady state :38-47, calibrated measurement errors :55-58,
    priors :60-77: parameters, one shock stderr, one measurement stderr, one shock and one measurement correlation)."""
    name = "hs_nk"
    linear = False
    endo = ["y", "R", "pi", "z", "g", "YGR", "INFL", "INT"]
    exo = ["eg", "ep", "ez"]
    params = ["tau", "rho_R", "psi1", "psi2", "rho_g", "rho_z", "kappa", "piA", "gammaQ", "rA", "s_ep", "s_eg", "s_ez"]
    params0 = np.array([2.83, 0.77, 1.8, 0.63, 0.98, 0.88, 0.78, 3.3, 0.52, 0.42, 0.22, 0.72, 0.31])
    sigma_e0 = np.eye(3)
    sigma_m0 = np.diag([(0.20 * 0.579923) ** 2, (0.20 * 1.470832) ** 2, (0.20 * 2.237937) ** 2])
    varobs = ["YGR", "INFL", "INT"]
    lagged = ["y", "R", "z", "g"]
    led = ["y", "pi", "z", "g"]
    est = [("param", "rA"), ("param", "piA"), ("param", "gammaQ"), ("param", "tau"), ("param", "kappa"),
           ("param", "psi1"), ("param", "psi2"), ("param", "rho_R"), ("param", "rho_g"), ("param", "rho_z"),
           ("param", "s_eg"), ("param", "s_ez"), ("stderr", "ep"), ("stderr", "INFL"),
           ("corr", "eg", "ez"), ("corr", "YGR", "INT")]
    priors = [("gamma", 0.5, 0.2), ("gamma", 7.0, 2.0), ("normal", 0.4, 0.2), ("gamma", 2.0, 0.5), ("beta", 0.5, 0.2),
              ("gamma", 1.5, 0.25), ("gamma", 0.5, 0.25), ("beta", 0.5, 0.2), ("beta", 0.5, 0.2), ("beta", 0.5, 0.2),
              ("inv_gamma", 1.2533141373155003, math.sqrt(0.4292036732051032)),
              ("inv_gamma", 0.6266570686577502, math.sqrt(0.1073009183012758)),
              ("inv_gamma", 1.0, 0.1), ("inv_gamma", 0.4, 0.1), ("normal", 0.0, 0.01), ("normal", 0.0, 0.01)]

    @property
    def theta_mode(self):                      # estimated_params_init (:80-93), calibration / prior mean elsewhere
        return np.array([0.333453026636814, 3.427852381637088, 0.624767199308368, 2.263895645223946, 0.9,
                         1.932357745633903, 0.465714640873466, 0.764476492352786, 0.990359718181570,
                         0.917302546854851, 0.632177304244290, 0.192160214260411, 1.0, 0.4, 0.0, 0.0])

    def steady_state(self, p):
        tau, rho_R, psi1, psi2, rho_g, rho_z, kappa, piA, gammaQ, rA, s_ep, s_eg, s_ez = p
        return np.array([0.0, 0.0, 0.0, 0.0, 0.0, gammaQ, piA, piA + rA + 4 * gammaQ])

    def dynamic_resid(self, ym, y_, yp, u, p):
        tau, rho_R, psi1, psi2, rho_g, rho_z, kappa, piA, gammaQ, rA, s_ep, s_eg, s_ez = p
        y, R, pi, z, g, YGR, INFL, INT = y_
        y_1, R_1, z_1, g_1 = ym[0], ym[1], ym[3], ym[4]
        y1, pi1, z1, g1 = yp[0], yp[2], yp[3], yp[4]
        eg, ep, ez = u
        beta = 1 / (1 + rA / 400)
        return np.array([
            y - (y1 - (1 / tau) * (R - pi1 - z1) + g - g1),
            pi - (beta * pi1 + kappa * (y - g)),
            R - (rho_R * R_1 + (1 - rho_R) * (psi1 * pi + psi2 * (y - g)) + s_ep * ep / 100),
            g - (rho_g * g_1 + s_eg * eg / 100),
            z - (rho_z * z_1 + s_ez * ez / 100),
            YGR - (gammaQ + 100 * (y - y_1 + z)),
            INFL - (piA + 400 * pi),
            INT - (piA + rA + 4 * gammaQ + 400 * R),
        ])


class MC4(HandModel):
    """Four linked Lubik-Schorfheide economies (dynare.jl_b200/models/src/mc4.mod, written for this repository as the
    medium-scale fixture: 44 endogenous variables, 20 shocks, 7 observables): the ls2003 block once per country, the foreign
    output of country c loaded on the lagged output of the three others.  Transcribed by hand from the equations."""
    name = "mc4"
    linear = True
    C = 4
    endo = [f"{v}{c}" for c in range(1, 5) for v in ("y", "ys", "R", "pie", "dq", "pies", "de", "A", "yobs", "pieobs", "Robs")]
    exo = [f"{v}{c}" for c in range(1, 5) for v in ("eR", "eq", "eys", "epies", "eA")]
    params = ["psi1", "psi2", "psi3", "rho_R", "tau", "alpha", "rr", "k", "rho_q", "rho_A", "rho_ys", "rho_pies", "spill"]
    params0 = np.array([0.54, 0.25, 0.25, 0.5, 0.5, 0.3, 2.51, 0.5, 0.4, 0.2, 0.8, 0.7, 0.1])
    sigma_e0 = np.diag(np.tile(np.array([1.25, 2.5, 1.37, 1.37, 1.37]) ** 2, 4))
    varobs = ["yobs1", "pieobs1", "Robs1", "yobs2", "pieobs2", "Robs2", "dq1"]
    lagged = [f"{v}{c}" for c in range(1, 5) for v in ("y", "ys", "R", "dq", "pies", "A")]
    led = [f"{v}{c}" for c in range(1, 5) for v in ("y", "ys", "pie", "dq", "A")]
    est = ([("param", p) for p in ("psi1", "psi2", "psi3", "rho_R", "alpha", "rr", "k", "tau", "rho_q", "rho_A", "rho_ys",
                                   "rho_pies", "spill")]
           + [("stderr", f"{v}{c}") for c in range(1, 5) for v in ("eR", "eq", "eys", "epies", "eA")])

    @property
    def theta_mode(self):
        return np.concatenate([[0.54, 0.25, 0.25, 0.5, 0.3, 2.51, 0.5, 0.5, 0.4, 0.2, 0.8, 0.7, 0.1],
                               np.tile([1.25, 2.5, 1.37, 1.37, 1.37], 4)])

    def steady_state(self, p):
        return np.zeros(44)

    def dynamic_resid(self, ym, y_, yp, u, p):
        psi1, psi2, psi3, rho_R, tau, alpha, rr, k, rho_q, rho_A, rho_ys, rho_pies, spill = p
        om = tau + alpha * (2 - alpha) * (1 - tau)
        r = []
        ylag = [ym[11 * c] for c in range(4)]
        for c in range(4):
            b = 11 * c
            y, ys, R, pie, dq, pies, de, A, yobs, pieobs, Robs = y_[b:b + 11]
            y_1, ys_1, R_1, dq_1, pies_1, A_1 = ym[b], ym[b + 1], ym[b + 2], ym[b + 4], ym[b + 5], ym[b + 7]
            y1, ys1, pie1, dq1, A1 = yp[b], yp[b + 1], yp[b + 3], yp[b + 4], yp[b + 7]
            eR, eq, eys, epies, eA = u[5 * c:5 * c + 5]
            others = sum(ylag[o] for o in range(4) if o != c)
            r += [
                y - (y1 - om * (R - pie1) - alpha * om * dq1 + alpha * (2 - alpha) * ((1 - tau) / tau) * (ys - ys1) - A1),
                pie - (np.exp(-rr / 400) * pie1 + alpha * np.exp(-rr / 400) * dq1 - alpha * dq + (k / om) * y
                       + alpha * (2 - alpha) * (1 - tau) / (tau * om) * ys),
                pie - (de + (1 - alpha) * dq + pies),
                R - (rho_R * R_1 + (1 - rho_R) * ((1 + psi1) * pie + psi2 * (y + alpha * (2 - alpha) * ((1 - tau) / tau) * ys)
                                                  + psi3 * de) + eR),
                dq - (rho_q * dq_1 + eq),
                ys - (rho_ys * ys_1 + (spill / 3) * others + eys),
                pies - (rho_pies * pies_1 + epies),
                A - (rho_A * A_1 + eA),
                yobs - (y - y_1 + A),
                pieobs - 4 * pie,
                Robs - 4 * R,
            ]
        return np.array(r)


def get_model(name: str) -> HandModel:
    return {"fs2000": FS2000, "ls2003": LS2003, "hs_nk": HSNK, "mc4": MC4}[name]()
