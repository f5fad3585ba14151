"""Import the *unmodified* reference (read-only at /root/reference) for golden-vector generation.

Only used in the build container (the GPU box has no /root/reference).  Nothing here is imported by the
product package.  Shims, all external to the reference tree:
  1. stub `bs4` package (the reference imports it but parses with xml.etree),
  2. flat `sys.path` import (the reference's package __init__ is broken),
  3. memoised sympy lambdify of the three Joint.get_*_function methods (bit-identical, ~20x faster),
  4. `QC`: QuadraticCost subclass accepting the iter_1..3 kwargs the solver passes.
"""
import os
import sys

REF = os.environ.get("B2T_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def available() -> bool:
    """REF is the reference tree, or the archive oracle/build_ref.py packs for the GPU box (imported through zipimport)."""
    if REF.endswith(".zip"):
        return os.path.isfile(REF)
    return os.path.isdir(REF) and os.path.isfile(os.path.join(REF, "TrajoptMPCReference.py"))


def load(memoise: bool = True):
    """Returns a namespace with the reference's public classes."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF)
    sys.dont_write_bytecode = True
    for p in (REF, os.path.join(HERE, "stubs")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import types
    import TrajoptMPCReference as T
    import TrajoptPlant as P
    import TrajoptCost as C
    import TrajoptConstraint as K
    import importlib
    PCGmod = importlib.import_module("GBD-PCG-Python")
    from GRiD.URDFParser.Joint import Joint
    from overloading import matrix_

    if memoise and not getattr(Joint, "_b2t_memoised", False):
        def _memo(name):
            orig = getattr(Joint, name)

            def cached(self):
                cache = self.__dict__.setdefault("_lam_cache", {})
                if name not in cache:
                    cache[name] = orig(self)
                return cache[name]
            setattr(Joint, name, cached)
        for name in ("get_transformation_matrix_function", "get_transformation_matrix_hom_function",
                     "get_dtransformation_matrix_hom_function"):
            _memo(name)
        Joint._b2t_memoised = True

    class QC(C.QuadraticCost):
        def value(self, x, u=None, timestep=None, iter_1=0, iter_2=0, iter_3=0):
            return super().value(x, u, timestep)

        def gradient(self, x, u=None, timestep=None, iter_1=0, iter_2=0, iter_3=0):
            return super().gradient(x, u, timestep)

        def hessian(self, x, u=None, timestep=None, iter_1=0, iter_2=0, iter_3=0):
            return super().hessian(x, u, timestep)

    ns = types.SimpleNamespace(
        TrajoptMPCReference=T.TrajoptMPCReference, SQPSolverMethods=T.SQPSolverMethods,
        URDFPlant=P.URDFPlant, TrajoptPlant=P.TrajoptPlant,
        QuadraticCost=C.QuadraticCost, UrdfCost=C.UrdfCost, QC=QC,
        TrajoptConstraint=K.TrajoptConstraint, BoxConstraint=K.BoxConstraint,
        PCG=PCGmod.PCG, matrix_=matrix_)
    return ns
