class DataLoader:  # not used by the oracle
    def __init__(self, *a, **k):
        raise NotImplementedError
