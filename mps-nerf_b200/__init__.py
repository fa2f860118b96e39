"""mpsnerf-b200: B200-native (sm_100a) implementation of MPS-NeRF's per-ray render hot path.

Import as ``mpsnerf_b200`` (the directory name ``mps-nerf_b200`` is aliased by the
top-level ``mpsnerf_b200`` shim).  Public API mirrors the reference's
``run_nerf_batch`` / ``lib`` surface; see ``DESIGN.md`` and ``INTEGRATION.md``.

    from mpsnerf_b200.run_nerf_batch import config_parser, create_nerf, render
    from mpsnerf_b200.lib.skinnning_batch import SKinningBatch
"""
__version__ = "0.1.0"
