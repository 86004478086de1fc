"""B200-native batched RANSAC pose-estimation engine (host-side Python plumbing).

The product is libransac_b200.so (CUDA kernels + C ABI, include/ransac_b200.h); this package
binds it for bench.py and the tests, generates the synthetic workloads of BASELINE.json and
shards candidates across ranks.
"""
from . import capi, synth  # noqa: F401
