from moleculesde_b200.data import Data, Batch  # noqa: F401
from . import makedirs  # noqa: F401


class InMemoryDataset:  # constructor signature only; datasets are out of scope
    def __init__(self, *a, **k):
        pass


def download_url(*a, **k):
    raise NotImplementedError


def extract_zip(*a, **k):
    raise NotImplementedError
