class MCSampler:
    pass
