"""ionsolver_b200 -- B200-native (sm_100a) implementation of IonSolver's extended-LBM MHD time step.

`ionsolver_b200.lbm` mirrors the reference's host surface (LbmConfig / Lbm / LbmDomain / Units); `ionsolver_b200.capi`
is the raw ctypes binding of include/ionsolver_b200.h.  All compute runs in hand-written CUDA kernels inside
libionsolver_b200.so; there is no CPU fallback.
"""
from .lbm import (FloatType, GraphicsConfig, Lbm, LbmConfig, LbmDomain, ModelType, RelaxationTime, TransferField, Units,  # noqa: F401
                  VelocitySet)

__all__ = ["FloatType", "GraphicsConfig", "Lbm", "LbmConfig", "LbmDomain", "ModelType", "RelaxationTime", "TransferField",
           "Units", "VelocitySet"]
