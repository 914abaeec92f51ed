"""Names only (TEST INFRASTRUCTURE ONLY): the reference imports these at trainer.py:9-10."""


class DistributedModelParallel:  # never constructed on the single-process oracle path
    def __init__(self, *a, **k):
        raise RuntimeError("DistributedModelParallel is not available in the oracle shim")
