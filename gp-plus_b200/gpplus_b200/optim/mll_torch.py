"""Adam fitting loop over the engine objective (optim/mll_torch.py:56-141 of the reference).

The loss is the exact marginal log-likelihood *divided by n* plus the log-priors divided by n -- what
gpytorch's ``ExactMarginalLogLikelihood`` computes (mll_torch.py:96,116) -- evaluated by the GPU
engine through ``model.log_marginal()``; restarts re-draw the parameters from their priors
(``reset_parameters``), the best state is restored at the end.
"""
import math
from copy import deepcopy
from typing import List, Optional

import torch


def _exact_mll_per_point(model) -> torch.Tensor:
    n = model.train_targets.shape[0]
    res = model.log_marginal()
    for _, module, prior, closure, _ in model.named_priors():
        res = res + prior.log_prob(closure(module)).sum()
    return res / n


def fit_model_torch(model, model_param_groups: Optional[List] = None, lr_default: float = 0.01, num_iter: int = 100,
                    num_restarts: int = 0, break_steps: int = 50):
    model.train()
    f_inc = math.inf
    best_state = deepcopy(model.state_dict())
    loss_hist_total = []
    for i in range(num_restarts + 1):
        optimizer = torch.optim.Adam(model.parameters() if model_param_groups is None else model_param_groups,
                                     lr=lr_default)
        loss_hist = []
        for j in range(num_iter):
            optimizer.zero_grad()
            loss = -_exact_mll_per_point(model)
            loss.backward()
            optimizer.step()
            loss_hist.append(loss.item())
            if j > break_steps and j % break_steps == 0:
                if (torch.mean(torch.Tensor(loss_hist)[j - break_steps:j]) - loss_hist[j]) <= 0:
                    break
        loss_hist_total.append(loss_hist)
        if loss.item() < f_inc:
            best_state = deepcopy(model.state_dict())
            f_inc = loss.item()
        if i < num_restarts:
            model.reset_parameters()
    model.load_state_dict(best_state)
    return f_inc, loss_hist_total
