class GPyTorchModel:
    """Mixin stub: GPR(ExactGP, GPyTorchModel) only needs the name to exist."""


class BatchedMultiOutputGPyTorchModel(GPyTorchModel):
    pass
