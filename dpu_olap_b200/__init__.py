"""dpu_olap_b200 — B200-native columnar operator path (Filter, Join, Sum, Take, Partition) behind
dpu_olap's host operator API. Hand-written sm_100a CUDA in csrc/, C ABI in include/b200olap.h.

Importing the package does not load the CUDA library; constructing a Context does, and raises if
libb200olap.so is missing (build it with ``python -m dpu_olap_b200.build``). No CPU fallback.
"""
__version__ = "0.1.0"
