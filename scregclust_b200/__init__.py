"""scregclust_b200 -- B200-native hot path of scregclust behind a C-ABI CUDA library."""
__version__ = "0.1.0"
