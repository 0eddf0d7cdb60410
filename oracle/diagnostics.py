"""Rank-normalised split-R-hat and bulk-ESS (Vehtari et al. 2021) in numpy.

TEST INFRASTRUCTURE ONLY: the checker for the engine's on-device diagnostics.  Restates what
`ess(chain; kind=:bulk)` / `rhat(chain)` of MCMCDiagnosticTools [dep, not in /root/reference]
compute for an (iterations x chains) array: split every chain in two, replace draws by normal
scores of pooled fractional ranks, then the BDA3/Stan estimator with Geyer's initial
monotone positive sequence.
"""
import numpy as np
from scipy import special, stats


def split_chains(x):
    """x: [draws, chains] -> [draws//2, 2*chains] (middle draw dropped when odd)."""
    n = x.shape[0]
    h = n // 2
    return np.concatenate([x[:h], x[n - h:]], axis=1)


def rank_normalize(x):
    r = stats.rankdata(x.reshape(-1), method="average").reshape(x.shape)
    return special.ndtri((r - 0.375) / (x.size + 0.25))


def _autocov_fft(x):
    """biased autocovariance per chain, x: [n, m]."""
    n = x.shape[0]
    xc = x - x.mean(axis=0, keepdims=True)
    nfft = 1 << int(np.ceil(np.log2(2 * n)))
    f = np.fft.rfft(xc, n=nfft, axis=0)
    ac = np.fft.irfft(f * np.conj(f), n=nfft, axis=0)[:n]
    return ac / n


def ess_rhat_basic(x, maxlag=None):
    """x: [n, m] already split (and rank-normalised for bulk). Returns (ess, rhat)."""
    n, m = x.shape
    acov = _autocov_fft(x)
    chain_var = acov[0] * n / (n - 1.0)
    mean_var = chain_var.mean()
    var_plus = mean_var * (n - 1.0) / n
    if m > 1:
        var_plus += x.mean(axis=0).var(ddof=1)
    rhat = np.sqrt(var_plus / mean_var)
    rho = np.zeros(n)
    rho[0] = 1.0
    acm = acov.mean(axis=1)
    rho_all = 1.0 - (mean_var - acm) / var_plus
    # Geyer: sum of adjacent pairs while positive, then enforce monotone decrease
    t = 1
    rho_even = 1.0
    rho_odd = rho_all[1]
    rho[1] = rho_odd
    limit = n if maxlag is None else min(n, maxlag)
    while t < limit - 3 and (rho_even + rho_odd) > 0:
        rho_even = rho_all[t + 1]
        rho_odd = rho_all[t + 2]
        if rho_even + rho_odd >= 0:
            rho[t + 1] = rho_even
            rho[t + 2] = rho_odd
        t += 2
    max_t = t
    if rho_even > 0:
        rho[max_t + 1] = rho_even
    t = 1
    while t <= max_t - 2:
        if rho[t + 1] + rho[t + 2] > rho[t - 1] + rho[t]:
            rho[t + 1] = (rho[t - 1] + rho[t]) / 2.0
            rho[t + 2] = rho[t + 1]
        t += 2
    tau = -1.0 + 2.0 * rho[:max_t + 1].sum() + (rho[max_t + 1] if rho_even > 0 else 0.0)
    tau = max(tau, 1.0 / np.log10(n * m))
    return n * m / tau, rhat


def ess_bulk(x):
    """x: [draws, chains] for one parameter."""
    return ess_rhat_basic(rank_normalize(split_chains(x)))[0]


def rhat(x):
    """max of rank-normalised split-R-hat and folded split-R-hat (Vehtari et al. 2021)."""
    s = split_chains(x)
    r1 = ess_rhat_basic(rank_normalize(s))[1]
    r2 = ess_rhat_basic(rank_normalize(np.abs(s - np.median(s))))[1]
    return max(r1, r2)


def rhat_bulk(x):
    return ess_rhat_basic(rank_normalize(split_chains(x)))[1]


def summarize(draws):
    """draws: [chains, draws, params] -> dict(mean, var, ess_bulk, rhat, mcse)."""
    C_, K, P = draws.shape
    out = dict(mean=np.zeros(P), var=np.zeros(P), ess_bulk=np.zeros(P), rhat=np.zeros(P),
               mcse=np.zeros(P))
    for p in range(P):
        x = draws[:, :, p].T
        out["mean"][p] = x.mean()
        out["var"][p] = x.var(ddof=1)
        e, r = ess_rhat_basic(rank_normalize(split_chains(x)))
        out["ess_bulk"][p] = e
        out["rhat"][p] = r
        e_mean = ess_rhat_basic(split_chains(x))[0]
        out["mcse"][p] = np.sqrt(out["var"][p] / e_mean)
    return out
