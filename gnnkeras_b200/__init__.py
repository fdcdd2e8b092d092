"""gnnkeras_b200 - a B200-native (sm_100a) fixed-point GNN loop behind the GNNkeras API.

The arithmetic lives in ``libgnnfp.so`` (hand-written CUDA, C ABI in ``include/gnnfp.h``); this package is
the host-side mirror of the reference's interface for that path.  Importing the package does not load
the library; any compute entry point does, and fails loudly if it is missing (there is no CPU fallback).
"""
__all__ = ["op", "models", "nets", "graph", "sequencers", "batcher", "synthetic"]
