"""Drop-in `simple_knn` package (`from simple_knn._C import distCUDA2`)."""
from . import _C  # noqa: F401
