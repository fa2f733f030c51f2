"""Interval score used as an optional MLL penalty (utils/interval_score.py:5-13)."""
import torch


def interval_score_function(Yu, Yl, Y, alpha=0.05):
    width = Yu - Yl
    out = width + (Y > Yu).to(torch.int64) * 2 / alpha * (Y - Yu) + (Y < Yl).to(torch.int64) * 2 / alpha * (Yl - Y)
    accuracy = torch.sum(out > width).to(torch.float64) / len(out)
    return torch.mean(out), accuracy
