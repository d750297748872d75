class PygLinkPropPredDataset:  # data loading is out of the oracle's scope
    def __init__(self, *a, **k):
        raise NotImplementedError("no datasets / network in this image")


class Evaluator:
    def __init__(self, *a, **k):
        raise NotImplementedError
