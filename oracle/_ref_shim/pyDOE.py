"""Stub of pyDOE.lhs (test_functions/multi_fidelity.py:5): the classic random Latin hypercube of pyDOE -- one
uniform point per stratum and dimension, strata permuted independently per dimension (numpy global RNG)."""
import numpy as np


def lhs(n, samples=None, criterion=None, iterations=None):
    if samples is None:
        samples = n
    cut = np.linspace(0, 1, samples + 1)
    u = np.random.rand(samples, n)
    a = cut[:samples]
    b = cut[1:samples + 1]
    rdpoints = np.zeros_like(u)
    for j in range(n):
        rdpoints[:, j] = u[:, j] * (b - a) + a
    H = np.zeros_like(rdpoints)
    for j in range(n):
        order = np.random.permutation(range(samples))
        H[:, j] = rdpoints[order, j]
    return H
