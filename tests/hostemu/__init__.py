"""TEST INFRASTRUCTURE: host (g++) build of the per-knot math in csrc/b2t_core.cuh, loaded with ctypes."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "trajoptmpcreference_b200", "csrc")


def load(tag):
    import sys
    sys.path.insert(0, ROOT)
    from trajoptmpcreference_b200.model import extract_model, builtin_urdf
    from trajoptmpcreference_b200.codegen import write_header
    write_header(extract_model(builtin_urdf(tag)), tag)
    out = os.path.join(HERE, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libhostemu_%s.so" % tag)
    srcs = [os.path.join(HERE, "hostemu.cpp"), os.path.join(CSRC, "b2t_core.cuh"), os.path.join(CSRC, "gen", "model_%s.h" % tag)]
    if not os.path.isfile(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        cmd = ["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", '-DB2T_MODEL_HEADER="gen/model_%s.h"' % tag,
               "-I", CSRC, srcs[0], "-o", so]
        subprocess.run(cmd, check=True)
    return ctypes.CDLL(so)
