"""Seeded synthetic problems shared by the parity tests (natural parameterisation of the C ABI)."""
import numpy as np


def make_problem(n, dq, kernel, dz=0, n_combo=0, n_noise=1, n_mean=1, seed=0, zero_mean_group=False, n_pass=1):
    rng = np.random.default_rng(seed)
    xq = rng.standard_normal((n, dq)) if dq > 0 else np.zeros((n, 0))
    p = {"n": n, "dq": dq, "dz": dz, "n_combo": n_combo if dz > 0 else 0, "n_noise": n_noise, "n_mean": n_mean,
         "kernel": kernel, "xq": xq, "n_pass": n_pass}
    f = np.zeros(n)
    if dq > 0:
        f = np.sin(xq[:, 0]) + 0.5 * np.cos(2.0 * xq[:, min(1, dq - 1)]) + 0.1 * xq.sum(1)
    if dz > 0:
        p["level_idx"] = rng.integers(0, n_combo, size=n).astype(np.int32)
        f = f + 0.3 * np.sin(1.0 + p["level_idx"])
    else:
        p["level_idx"] = None
    if n_noise > 1:
        p["noise_idx"] = rng.integers(0, n_noise, size=n).astype(np.int32)
    else:
        p["noise_idx"] = None
    if n_mean > 1:
        mi = rng.integers(0, n_mean, size=n).astype(np.int32)
        if zero_mean_group:
            mi = mi - 1  # group -1 = zero mean (reference source 0 with m_gp_ref='zero')
        p["mean_idx"] = mi
    else:
        p["mean_idx"] = None
    y = f + 0.05 * rng.standard_normal(n)
    y = (y - y.min()) / (y.max() - y.min()) if n > 1 else np.array([0.3])
    p["y"] = y
    return p


def make_hyper(p, seed=1, noise=1e-3, w_scale=0.3):
    rng = np.random.default_rng(seed)
    h = {"w": w_scale * np.exp(0.5 * rng.standard_normal(p["dq"])),
         "sigma_f2": 0.7 + 0.2 * rng.random(),
         "noise": noise * (1.0 + rng.random(p["n_noise"]))}
    if p["dz"] > 0 and p.get("n_pass", 1) > 1:
        base = 0.8 * rng.standard_normal((1, p["n_combo"], p["dz"]))
        h["z"] = base + 0.3 * rng.standard_normal((p["n_pass"], p["n_combo"], p["dz"]))
    elif p["dz"] > 0:
        h["z"] = 0.8 * rng.standard_normal((p["n_combo"], p["dz"]))
    else:
        h["z"] = None
    n_mean = p["n_mean"]
    h["beta"] = 0.3 + 0.1 * rng.standard_normal(n_mean) if n_mean > 0 else None
    return h


def make_candidates(p, m, seed=2):
    rng = np.random.default_rng(seed)
    c = {"m": m, "xq": rng.standard_normal((m, p["dq"])) if p["dq"] > 0 else np.zeros((m, 0))}
    c["level_idx"] = rng.integers(0, p["n_combo"], size=m).astype(np.int32) if p["dz"] > 0 else None
    c["noise_idx"] = rng.integers(0, p["n_noise"], size=m).astype(np.int32) if p["n_noise"] > 1 else None
    if p["n_mean"] > 1:
        lo = -1 if (p["mean_idx"] is not None and p["mean_idx"].min() < 0) else 0
        c["mean_idx"] = rng.integers(lo, p["n_mean"] + lo if lo < 0 else p["n_mean"], size=m).astype(np.int32)
    else:
        c["mean_idx"] = None
    return c


def engine_kwargs(p, device=0):
    return dict(xq=p["xq"], y=p["y"], kernel=p["kernel"], level_idx=p["level_idx"], n_combo=p["n_combo"],
                dz=p["dz"], noise_idx=p["noise_idx"], n_noise=p["n_noise"], mean_idx=p["mean_idx"],
                n_mean=p["n_mean"], device=device, n_pass=p.get("n_pass", 1))
