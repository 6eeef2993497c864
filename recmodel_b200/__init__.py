"""recmodel_b200: a B200-native (sm_100a) implementation of the WMF train + top-N path of
titoeb/RecModel behind the reference's own ``WMF`` class API. See DESIGN.md."""
from .wmf_model import WMF, WMFModel  # noqa: F401
from .ease_model import Ease  # noqa: F401
from .utils import test_coverage, train_test_split_sparse_mat  # noqa: F401

__all__ = ["WMF", "WMFModel", "Ease", "train_test_split_sparse_mat", "test_coverage"]
