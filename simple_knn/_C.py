"""`simple_knn._C.distCUDA2` (simple-knn/ext.cpp:15-17, spatial.cu:15-26) over the C ABI."""
from youreditableavatar_b200.knn import distCUDA2

__all__ = ["distCUDA2"]
