"""GPU half of the reference pin: the product path (GP_Plus -> MLLObjective.fun / fun_fast -> C ABI -> sm_100a
kernels; GP_Plus.predict -> gpp_factorize + gpp_predict) against fixtures PRODUCED BY THE REFERENCE'S OWN CODE
(tests/golden/ref_*.npz, see tests/golden/make_reference_fixtures.py and tests/test_reference_pin.py).

Tolerances (float64 reference model): data term 1e-9 relative, gradients 1e-8 of the gradient's max-norm,
posterior (with float32 prior constants inside the reference) 2e-8, predictions 1e-7 of the response range.
"""
import warnings

import numpy as np
import pytest
import torch

from test_reference_pin import (MODEL_CASES, TOL_NLL, TOL_POST, build_product_model, grad_tol, load_case,
                                post_scale)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", MODEL_CASES)
def test_engine_objective_matches_the_reference(name):
    from gpplus_b200.optim.mll_scipy import MLLObjective
    z, meta, kw = load_case(name)
    m = build_product_model(z, kw)
    try:
        for add_prior, key, tol in ((False, "nll", TOL_NLL), (True, "post", TOL_POST)):
            obj = MLLObjective(m, add_prior, [0, 0])
            fast = obj.enable_fast_path()
            for k, th in enumerate(z["thetas"]):
                fr, gr = float(z["f_f64_" + key][k]), z["g_f64_" + key][k]
                f, g = obj.fun(th.copy())
                scale = abs(fr) if not add_prior else post_scale(z, k)
                assert abs(f - fr) <= tol * scale, (name, key, k, f, fr)
                assert np.max(np.abs(g - gr)) <= (grad_tol(kw) if not add_prior else max(1e-7, grad_tol(kw))) * np.max(np.abs(gr)), (name, key, k)
                if fast:  # the objective the restart workers use (gpp_objective: transforms + priors in the library)
                    f2, g2 = obj.fun_fast(th.copy())
                    assert abs(f2 - fr) <= tol * scale, (name, key, k, f2, fr)
                    assert np.max(np.abs(g2 - gr)) <= (grad_tol(kw) if not add_prior else max(1e-7, grad_tol(kw))) * np.max(np.abs(gr))
    finally:
        m.release_engine()


@pytest.mark.parametrize("name", MODEL_CASES)
def test_engine_predictions_match_the_reference(name):
    """Mixed batches, batches holding a single level of a categorical column, and a single row: the reference ranks
    the categorical columns of [train, test] together (ExactGP eval mode + setlevels), so a batch that lacks some
    level must still use that level's own latent position / noise / mean."""
    from gpplus_b200.optim.mll_scipy import MLLObjective
    z, meta, kw = load_case(name)
    m = build_product_model(z, kw)
    try:
        MLLObjective(m, True, [0, 0])._load(z["thetas"][0])
        Xte = torch.as_tensor(z["Xte"])
        span = float(z["ytr"].max() - z["ytr"].min())
        for inc, tag in ((True, "noise"), (False, "nonoise")):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                mu, sd = m.predict(Xte.clone(), return_std=True, include_noise=inc)
            assert np.max(np.abs(mu.numpy() - z["pred_mean_" + tag])) <= 1e-7 * span, (name, tag)
            assert np.max(np.abs(sd.numpy() - z["pred_std_" + tag])) <= 1e-6 * span, (name, tag)
        if "single_levels" in z.files:
            col = 0 if name.startswith("c2_") else Xte.shape[1] - 1
            for k, v in enumerate(z["single_levels"]):
                rows = Xte[Xte[:, col] == v]
                mu, sd = m.predict(rows.clone(), return_std=True, include_noise=True)
                assert np.max(np.abs(mu.numpy() - z["single_%d_mean" % k])) <= 1e-7 * span, (name, k)
                assert np.max(np.abs(sd.numpy() - z["single_%d_std" % k])) <= 1e-6 * span, (name, k)
            mu, sd = m.predict(Xte[:1].clone(), return_std=True, include_noise=True)
            assert np.max(np.abs(mu.numpy() - z["one_row_mean"])) <= 1e-7 * span
            assert np.max(np.abs(sd.numpy() - z["one_row_std"])) <= 1e-6 * span
    finally:
        m.release_engine()


@pytest.mark.parametrize("name", [c for c in MODEL_CASES if c.startswith("c1_") or c.startswith("c4_")])
def test_adam_loss_history_matches_the_reference(name):
    """fit_model_torch (optim/mll_torch.py:56-141): the loss is -(log p(y) + log priors) / n; six Adam steps from the
    same start reproduce the reference's loss history and end point."""
    from gpplus_b200.optim import fit_model_torch
    from gpplus_b200.optim.mll_scipy import MLLObjective
    z, meta, kw = load_case(name)
    if "adam_loss_hist" not in z.files:
        pytest.skip("no Adam history stored for this case")
    m = build_product_model(z, kw)
    try:
        obj = MLLObjective(m, True, [0, 0])
        obj._load(z["thetas"][0])
        f_inc, hist = fit_model_torch(m, num_iter=6, num_restarts=0, lr_default=0.01)
        np.testing.assert_allclose(np.asarray(hist[0]), z["adam_loss_hist"], rtol=2e-8)
        assert abs(f_inc - float(z["adam_f_inc"])) <= 2e-8 * abs(float(z["adam_f_inc"]))
        # fit_model_torch restores the best state (the last one here)
        np.testing.assert_allclose(obj.pack_parameters(), z["adam_theta_end"], rtol=0, atol=1e-6)
    finally:
        m.release_engine()
