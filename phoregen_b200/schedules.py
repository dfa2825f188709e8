"""Noise schedules and transition tables, float64 numpy -> float32 frozen tables.

Behaviour follows reference models/common.py:446-544 (beta schedules) and models/transition.py:10-26,179-243
(Gaussian posterior coefficients; D3PM-style K x K one-step / cumulative matrices with an absorbing prior).
Init-time only; a checkpoint overrides every table through load_state_dict.
"""
import numpy as np


def _logistic(v):
    return 1.0 / (1.0 + np.exp(-v))


def sigmoid_alpha_bar(T, scale_start, scale_end, width):
    """alpha_bar(t): a logistic ramp from scale_start (t=0) to scale_end (t=T-1)  (common.py:459-480 'advance')."""
    amp = (scale_end - scale_start) / (_logistic(-width) - _logistic(width))
    shift = 0.5 * (scale_end + scale_start - amp)
    grid = np.linspace(-1.0, 1.0, T)
    return amp * _logistic(-width * grid) + shift


def betas_from_alpha_bar(alpha_bar):
    ratio = np.empty_like(alpha_bar)
    ratio[0] = alpha_bar[0]
    ratio[1:] = alpha_bar[1:] / alpha_bar[:-1]
    return np.clip(1.0 - ratio, 0.0, 1.0)


def beta_schedule(kind, T, **kw):
    """`advance` and `segment` are the two schedules the shipped configs use (configs/train_lig-phore.yml:19-40);
    the simple closed forms of common.py:507-533 are kept for config compatibility."""
    if kind == "advance":
        return betas_from_alpha_bar(sigmoid_alpha_bar(T, kw.get("scale_start", 0.999), kw.get("scale_end", 0.001),
                                                      kw.get("width", 2)))
    if kind == "segment":
        seg, params = kw["time_segment"], kw["segment_diff"]
        if int(np.sum(seg)) != T:
            raise ValueError("time_segment must sum to num_timesteps")
        pieces = [sigmoid_alpha_bar(int(n) + 1, p["scale_start"], p["scale_end"], p["width"])[1:]
                  for n, p in zip(seg, params)]
        return betas_from_alpha_bar(np.concatenate(pieces))
    if kind == "linear":
        return np.linspace(kw["beta_start"], kw["beta_end"], T, dtype=np.float64)
    if kind == "quad":
        return np.linspace(kw["beta_start"] ** 0.5, kw["beta_end"] ** 0.5, T, dtype=np.float64) ** 2
    if kind == "const":
        return kw["beta_end"] * np.ones(T, dtype=np.float64)
    if kind == "jsd":
        return 1.0 / np.linspace(T, 1, T, dtype=np.float64)
    if kind == "sigmoid":
        s = kw.get("s", 6)
        return _logistic(np.linspace(-s, s, T)) * (kw["beta_end"] - kw["beta_start"]) + kw["beta_start"]
    if kind == "cosine":
        s = kw.get("s", 0.008)
        grid = np.linspace(0, T + 1, T + 1)
        ab = np.cos(((grid / (T + 1)) + s) / (1 + s) * np.pi * 0.5) ** 2
        ab = ab / ab[0]
        return np.clip(1 - ab[1:] / ab[:-1], 0, 0.999)
    raise NotImplementedError(kind)


def gaussian_tables(betas):
    """Frozen tables of the position transition (transition.py:14-26)."""
    alphas = 1.0 - betas
    ab = np.cumprod(alphas)
    ab_prev = np.concatenate([[1.0], ab[:-1]])
    return {
        "betas": betas, "alphas": alphas, "alphas_bar": ab, "alphas_bar_prev": ab_prev,
        "coef_x0": np.sqrt(ab_prev) * betas / (1.0 - ab),
        "coef_xt": np.sqrt(alphas) * (1.0 - ab_prev) / (1.0 - ab),
        "std": np.sqrt((1.0 - ab_prev) * betas / (1.0 - ab)),
    }


def prior_probs(kind, K):
    """transition.py:183-196."""
    if kind in (None, "uniform"):
        p = np.ones(K)
    elif kind == "absorb":
        p = 0.01 * np.ones(K)
        p[0] = 1.0
    elif kind == "tomask":
        p = 0.001 * np.ones(K)
        p[-1] = 1.0
    else:
        p = np.asarray(kind, dtype=np.float64)
    return p / p.sum()


def categorical_tables(betas, K, init_prob):
    """Q_t = (1-beta_t) I + beta_t 1 prior^T ; cumulative products ; transposed one-step matrices
    (transition.py:200-243)."""
    prior = prior_probs(init_prob, K)
    T = len(betas)
    one = np.empty((T, K, K))
    for t in range(T):
        one[t] = betas[t] * np.tile(prior[None, :], (K, 1)) + (1.0 - betas[t]) * np.eye(K)
    cum = np.empty_like(one)
    cum[0] = one[0]
    for t in range(1, T):
        cum[t] = cum[t - 1] @ one[t]
    return {"q_mats": cum, "transpopse_q_onestep_mats": np.transpose(one, (0, 2, 1))}, prior
