"""gpytorch.metrics (>= 1.6): the test-set metrics GP_Plus.evaluation reads (models/gp_plus.py:903-908)."""
import torch


def mean_absolute_error(pred_dist, test_y):
    return torch.abs(pred_dist.mean - test_y).mean(dim=-1)


def mean_squared_error(pred_dist, test_y, squared=True):
    res = torch.square(pred_dist.mean - test_y).mean(dim=-1)
    return res if squared else res ** 0.5


def negative_log_predictive_density(pred_dist, test_y):
    combine_dim = -2 if len(pred_dist.event_shape) > 1 else -1
    return -pred_dist.log_prob(test_y) / test_y.shape[combine_dim]


def mean_standardized_log_loss(pred_dist, test_y):
    combine_dim = -2 if len(pred_dist.event_shape) > 1 else -1
    f_mean = pred_dist.mean
    f_var = pred_dist.variance
    return (0.5 * torch.log(2 * torch.pi * f_var) + torch.square(test_y - f_mean) / (2 * f_var)).mean(dim=combine_dim)
