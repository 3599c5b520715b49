"""CPU oracle for the rasteriser hot path.

TEST INFRASTRUCTURE ONLY. Nothing in ``taichi_3d_gaussian_splatting_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it (as the checker / the CPU baseline, never as the product path).
"""
from .gs_oracle import (  # noqa: F401
    OracleRasterisation,
    OracleForwardResult,
    OracleBackwardResult,
    build_oracle,
    lib,
    inverse_SE3_qt,
)
