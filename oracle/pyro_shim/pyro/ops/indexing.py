class Vindex:
    """Only imported by the reference's non-multinomial models; not on the SparseMultinomialGDRF path."""

    def __init__(self, tensor):
        self.tensor = tensor

    def __getitem__(self, args):
        raise NotImplementedError("shim: Vindex is outside the accelerated path")
