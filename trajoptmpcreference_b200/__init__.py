"""B200-native batched trajectory optimisation behind the TrajoptMPCReference Python API (SQP + Schur complement +
GBD-PCG).  All numerics run in hand-written sm_100a CUDA kernels through the C ABI in include/b2t.h."""
from .api import (SQPSolverMethods, MPCSolverMethods, TrajoptPlant, URDFPlant, TrajoptCost, QuadraticCost, UrdfCost,
                  BoxConstraint, TrajoptConstraint, TrajoptMPCReference, BatchSolver, BatchResult, PCG)
from ._lib import B2TError

__all__ = ["SQPSolverMethods", "MPCSolverMethods", "TrajoptPlant", "URDFPlant", "TrajoptCost", "QuadraticCost", "UrdfCost",
           "BoxConstraint", "TrajoptConstraint", "TrajoptMPCReference", "BatchSolver", "BatchResult", "PCG", "B2TError"]
