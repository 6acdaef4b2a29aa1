"""deadtrees_b200 — B200-native (sm_100a) implementation of the cwerner/deadtrees hot path.

Mirrors the reference's module paths for the path in scope:
``network.SemSegment``, ``deployment.tiler.Tiler``, ``deployment.inference.PyTorchInference``,
``utils.data_handling.{make,unmake}_blocks_vectorized``, ``loss.losses`` / ``loss.gdl``,
``data.deadtreedata.val_transform``.  All compute goes through ``libdeadtrees_b200.so``
(hand-written CUDA, C-ABI in ``include/deadtrees_b200.h``); there is no CPU fallback.
"""
__version__ = "0.1.0"
