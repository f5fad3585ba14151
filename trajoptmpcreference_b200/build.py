"""Build the per-robot CUDA libraries (libb2t_<robot>.so) in-tree with nvcc for sm_100a.

  python -m trajoptmpcreference_b200.build [robot ...]        (default: every built-in URDF)

Each library = generated model header (codegen.py) + csrc/b2t_lib.cu.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys
import hashlib

from .model import extract_model, builtin_urdf, model_digest
from .codegen import write_header

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
BUILTIN = ["pend", "arm1", "arm2", "arm3", "arm4", "arm6", "cartpole"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--expt-relaxed-constexpr",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def lib_path(tag: str) -> str:
    return os.path.join(LIBDIR, "libb2t_%s.so" % tag)


def _sources_digest(header: str) -> str:
    h = hashlib.sha1()
    for p in (header, os.path.join(CSRC, "b2t_core.cuh"), os.path.join(CSRC, "b2t_kernels.cuh"), os.path.join(CSRC, "b2t_pcg_tm.cuh"), os.path.join(CSRC, "b2t_ilqr.cuh"), os.path.join(CSRC, "b2t_lib.cu"),
              os.path.join(os.path.dirname(HERE), "include", "b2t.h")):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_model(model: dict, tag: str, force: bool = False, verbose: bool = False) -> str:
    """Generate the header for `model`, compile libb2t_<tag>.so if stale, return its path."""
    os.makedirs(LIBDIR, exist_ok=True)
    header = write_header(model, tag)
    out = lib_path(tag)
    stamp = out + ".stamp"
    dig = _sources_digest(header)
    if not force and os.path.isfile(out) and os.path.isfile(stamp) and open(stamp).read().strip() == dig:
        return out
    # two translation units (double solver + C ABI, float solver) compiled in parallel, then linked into one shared library
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(HERE, "_build")
    os.makedirs(objdir, exist_ok=True)
    objs = [os.path.join(objdir, "b2t_%s_part%d.o" % (tag, part)) for part in (1, 2)]
    cmds = [[NVCC] + NVCC_FLAGS + ['-DB2T_MODEL_HEADER="gen/model_%s.h"' % tag, "-DB2T_PART=%d" % part, "-I", CSRC, "-c",
                                   os.path.join(CSRC, "b2t_lib.cu"), "-o", obj] for part, obj in zip((1, 2), objs)]
    with ThreadPoolExecutor(max_workers=2) as ex:
        results = list(ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), cmds))
    log = "".join(" ".join(c) + "\n" + r.stdout + r.stderr for c, r in zip(cmds, results))
    ok = all(r.returncode == 0 for r in results)
    if ok:
        link = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"] + objs + ["-o", out]
        res = subprocess.run(link, capture_output=True, text=True)
        log += " ".join(link) + "\n" + res.stdout + res.stderr
        ok = res.returncode == 0
    with open(out + ".buildlog", "w") as f:
        f.write(log)
    if not ok:
        raise RuntimeError("nvcc failed for %s:\n%s" % (tag, log[-6000:]))
    for obj in objs:
        os.remove(obj)
    if verbose:
        print(log)
    with open(stamp, "w") as f:
        f.write(dig)
    return out


def build_builtin(names=None, force=False, verbose=False, jobs=None):
    """Build the libraries of the built-in robots, `jobs` nvcc processes at a time (default: one per core, at most one per robot)."""
    from concurrent.futures import ThreadPoolExecutor
    names = list(names or BUILTIN)
    jobs = jobs or max(1, min(len(names), os.cpu_count() or 1))
    models = [extract_model(builtin_urdf(name)) for name in names]
    # longest compile first (the 6-joint robot), so that it overlaps all the others
    order = sorted(range(len(names)), key=lambda i: -models[i]["n"])
    outs = [None] * len(names)
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        futs = {i: ex.submit(build_model, models[i], names[i], force, verbose) for i in order}
        for i, f in futs.items():
            outs[i] = f.result()
    return outs


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("-")]
    for p in build_builtin(names or None, force="--force" in sys.argv, verbose="-v" in sys.argv):
        print(p)
