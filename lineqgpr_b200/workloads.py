"""Synthetic instances of the five BASELINE.json configurations, built with the
host model layer (lineqgpr_b200.model) so that every consumer (GPU path, CPU checker, benchmark)
sees bit-identical (mu, map, Sigma, lb, ub).  Seeded numpy only; no files, no network."""
from __future__ import annotations

import numpy as np

from . import model as M


def _tmvpar_lineqGP(mdl, xtest, sampler):
    pred = M.predict_lineqGP(mdl, xtest)
    a = pred["model"]
    Lam = pred["Lambda"]
    Sigma_eta = Lam @ pred["Sigma"] @ Lam.T + a["kernParam"]["nugget"] * np.eye(Lam.shape[0])   # R/lineqGP.R:613-614
    return dict(mu=Lam @ pred["mu"], map=Lam @ pred["xi.map"], Sigma=Sigma_eta, lb=pred["lb"], ub=pred["ub"],
                **{"class": sampler}), dict(Lambda=Lam, Phi_test=pred["Phi.test"], xi_map=pred["xi.map"], model=a)


def _tmvpar_lineqAGP(mdl, xtest, sampler):
    pred = M.predict_lineqAGP(mdl, xtest)
    a = pred["model"]
    Lam = a["lineqSys"]["LambdaAll"]
    eta_map = Lam @ pred["xiAll.map"]
    Sigma_eta = Lam @ pred["SigmaAll"] @ Lam.T + a["nugget"] * np.eye(Lam.shape[0])             # R/lineqAGP.R:559-560
    return dict(mu=eta_map, Sigma=Sigma_eta, lb=a["lineqSys"]["lb"], ub=a["lineqSys"]["ub"],       # mu := MAP (Q4)
                **{"class": sampler}), dict(Lambda=Lam, Phi_test=pred["PhiAll.test"], xi_map=pred["xiAll.map"], model=a)


def cfg1(m=100, sampler="HMC", n_test=100):
    """1-D monotone lineqGP, m hat knots, 10 noisy points; 1000 HMC samples via simulate."""
    x = np.linspace(0.05, 0.95, 10)
    y = 1.0 / (1.0 + np.exp(-10.0 * (x - 0.5))) + 0.05 * np.random.default_rng(0).standard_normal(10)
    mdl = M.create_lineqGP(x, y, "monotonicity", m=m)
    mdl["varnoise"] = 0.05 ** 2
    mdl["localParam"]["sampler"] = sampler
    xtest = np.linspace(0.0, 1.0, n_test)
    par, aux = _tmvpar_lineqGP(mdl, xtest, sampler)
    aux.update(lineq_model=mdl, xtest=xtest, nsim=1000, burn_in=1, name="cfg1")
    return par, aux


def cfg2(sampler="HMC"):
    """2-D monotone MaxMod lineqGP, final knots 5 x 2 (nb/lineqGPR_MaxModAlgorithm_Monotonicity2D#c5)."""
    rng = np.random.default_rng(2)
    x = rng.uniform(size=(80, 2))
    y = np.arctan(5.0 * x[:, 0]) + 0.5 * x[:, 1]
    mdl = M.create_lineqGP(x, y, "monotonicity", m=[5, 2])
    mdl["ulist"] = [np.array([0.0, 0.2047, 0.4020, 0.6275, 1.0]), np.array([0.0, 1.0])]
    mdl["kernParam"] = dict(par=np.array([1.0, 0.3, 0.3]), type="matern52", nugget=1e-5)
    mdl["varnoise"] = float(np.var(y, ddof=1))
    mdl["localParam"]["sampler"] = sampler
    g = np.linspace(0.0, 1.0, 10)
    xtest = np.array([[a, b] for a in g for b in g])
    par, aux = _tmvpar_lineqGP(mdl, xtest, sampler)
    aux.update(lineq_model=mdl, xtest=xtest, nsim=10 ** 4, burn_in=1, name="cfg2")
    return par, aux


def _additive(D, knots, n, seed, nsim, n_test, seed_test, active=None, sampler="HMC", name="cfgA"):
    rng = np.random.default_rng(seed)
    Dfull = D if active is None else active[1]
    xfull = rng.uniform(size=(n, Dfull))
    k_act = D
    a = 5.0 * (1.0 - (np.arange(1, k_act + 1)) / (3.0 if active else (D + 1.0)))
    x = xfull[:, :k_act]
    y = np.sum(np.arctan(a[None, :] * x), axis=1)
    mdl = M.create_lineqAGP(x, y, ["monotonicity"] * k_act, m=knots)
    for k in range(k_act):
        mdl["kernParam"][k] = dict(par=np.array([1.0, 2.0]), type="matern52")
    mdl["nugget"] = 1e-5
    mdl["varnoise"] = 0.05 * float(np.var(y, ddof=1))
    mdl["localParam"]["sampler"] = sampler
    xtest = np.random.default_rng(seed_test).uniform(size=(n_test, k_act))
    par, aux = _tmvpar_lineqAGP(mdl, xtest, sampler)
    aux.update(lineq_model=mdl, xtest=xtest, nsim=nsim, burn_in=1, name=name)
    return par, aux


def cfg3(sampler="HMC"):
    """5-D additive MaxMod lineqAGP: two active dimensions with 6 and 4 knots."""
    return _additive(2, [6, 4], 100, 8, 10 ** 5, 100, 9, active=(2, 5), sampler=sampler, name="cfg3")


def cfg4(m=1000, sampler="HMC", n_test=200):
    """1-D bounded + monotone lineqGP, m knots: d = 2m, 3m-1 finite bound rows (Q3: mvec := 2m)."""
    x = np.linspace(0.05, 0.95, 10)
    y = 1.0 / (1.0 + np.exp(-10.0 * (x - 0.5))) + 0.05 * np.random.default_rng(0).standard_normal(10)
    mdl = M.create_lineqGP(x, y, ["boundedness", "monotonicity"], m=m)
    mdl["bounds"] = np.array([[-0.2, 1.2], [0.0, np.inf]])      # demos override by hand (Q10)
    mdl["varnoise"] = 0.05 ** 2
    mdl["localParam"]["sampler"] = sampler
    xtest = np.linspace(0.0, 1.0, n_test)
    par, aux = _tmvpar_lineqGP(mdl, xtest, sampler)
    aux.update(lineq_model=mdl, xtest=xtest, nsim=10 ** 6, burn_in=100, name="cfg4")
    return par, aux


def cfg5(D=100, knots=10, n=500, n_test=100000, sampler="HMC"):
    """Additive monotone lineqAGP, D dimensions x `knots` knots (m = D*knots)."""
    return _additive(D, knots, n, 3, 10 ** 6, n_test, 4, sampler=sampler, name="cfg5")


CONFIGS = {"cfg1": cfg1, "cfg2": cfg2, "cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5}
