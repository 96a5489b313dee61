"""nmfgpu_b200 -- B200-native NMF engine behind the nmfgpu C API (libnmfgpu64.so) and its ctypes mirror."""
from .api import (ALGORITHM_BY_NAME, IndexBase, Library, NmfAlgorithm, NmfInitializationMethod, NmfThresholdType,  # noqa: F401
                  ResultType, Session, StorageFormat, Verbosity)
