from .multifidelity import Multifidelity_likelihood, Multifidelity_noise

__all__ = ["Multifidelity_likelihood", "Multifidelity_noise"]
