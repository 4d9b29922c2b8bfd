"""falcon_b200 -- B200-native implementation of falcon's clustering hot path.

Drop-in for the ``falcon.cluster`` entry points named by the north star
(spectrum vectorisation, ``compute_pairwise_distances``, ``generate_clusters``);
see ``falcon_b200.cluster``.  All compute runs in hand-written sm_100a CUDA
kernels behind the C ABI of ``include/falcon_b200.h``; there is no CPU fallback.
"""
__version__ = "0.2.0+b200"
