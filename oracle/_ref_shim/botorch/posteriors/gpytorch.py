class GPyTorchPosterior:
    def __init__(self, mvn):
        self.mvn = mvn

    @property
    def mean(self):
        return self.mvn.mean.unsqueeze(-1)

    @property
    def variance(self):
        return self.mvn.variance.unsqueeze(-1)
