"""URDF -> numeric robot model (topology + constant matrices), no sympy at run time.

Replaces the reference's sympy object model (GRiD/URDFParser/*) by a load-time extractor whose output is what
the CUDA code generator bakes into the kernels.  Conventions follow the reference exactly so that the numbers
are identical (tests/test_model.py compares against tests/golden/models.json, which was dumped from the
reference's own `Robot` objects):

  * joint / link numbering: DFS from the root link in URDF joint order  (URDFParser.py:374-393,420-435)
  * X_j(t) = X_free(t) * X_fixed,  X_fixed = rot(E_rpy) * xlt(skew(xyz))  (SpatialAlgebra.py:20-23,45-46,93; Joint.py:57-88)
    with the transposed-rotation convention rz = [[c,s,0],[-s,c,0],[0,0,1]]   (SpatialAlgebra.py:48-64)
  * constants snapped like `sp.nsimplify(.., tolerance=1e-6, rational=True).evalf()` (Joint.py:90,95), i.e.
    Fraction(v).limit_denominator(10**6)  (sympy _real_to_rational)
  * 4x4 homogeneous transform: rotation (R_free R_fixed)^T, translation t_free + t_fixed  (Joint.py:92-97)
  * spatial inertia from the link's own <origin xyz> (Link.py:48-65, URDFParser.py:270-277)
  * fixed joints folded into the parent (URDFParser.py:330-351)

X_j(t) is affine in (cos t, sin t) for revolute joints and in t for prismatic joints, so the model stores three
constant matrices per joint:  X_j(t) = X0 + f1(t) Xa + f2(t) Xb.
"""
import json
import math
import os
import xml.etree.ElementTree as ET
from fractions import Fraction

import numpy as np

URDF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "urdf")


def _snap(v, den):
    """float(nsimplify(v, tolerance=1/den, rational=True))"""
    v = float(v)
    if v == 0.0 or not math.isfinite(v):
        return v
    fr = Fraction(v).limit_denominator(den)
    return fr.numerator / fr.denominator


_snap6 = np.vectorize(lambda v: _snap(v, 10 ** 6), otypes=[np.float64])
_snap12 = np.vectorize(lambda v: _snap(v, 10 ** 12), otypes=[np.float64])


def _rx(t):
    c, s = math.cos(t), math.sin(t)
    return np.array([[1, 0, 0], [0, c, s], [0, -s, c]], dtype=np.float64)


def _ry(t):
    c, s = math.cos(t), math.sin(t)
    return np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]], dtype=np.float64)


def _rz(t):
    c, s = math.cos(t), math.sin(t)
    return np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]], dtype=np.float64)


def _skew(x, y, z):
    return np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]], dtype=np.float64)


def _rot6(E):
    out = np.zeros((6, 6))
    out[:3, :3] = E
    out[3:, 3:] = E
    return out


def _xlt(rx):
    out = np.eye(6)
    out[3:, :3] = -rx
    return out


# basis matrices of the free 3x3 rotation  E(t) = E0 + cos(t) Ec + sin(t) Es  (same transposed convention)
_REV = {
    2: (np.diag([0.0, 0, 1]), np.diag([1.0, 1, 0]), np.array([[0.0, 1, 0], [-1, 0, 0], [0, 0, 0]])),
    1: (np.diag([0.0, 1, 0]), np.diag([1.0, 0, 1]), np.array([[0.0, 0, -1], [0, 0, 0], [1, 0, 0]])),
    0: (np.diag([1.0, 0, 0]), np.diag([0.0, 1, 1]), np.array([[0.0, 0, 0], [0, 0, 1], [0, -1, 0]])),
}


def _axis_index(axis):
    # Joint.set_type tests axis[2]==1, then axis[1]==1, then axis[0]==1 (Joint.py:55-80)
    for i in (2, 1, 0):
        if float(axis[i]) == 1.0:
            return i
    raise ValueError("joint axis must be a unit basis vector (reference limitation, Joint.py:55-80): %r" % (axis,))


def _floats(s, default="0 0 0"):
    return [float(t) for t in (s if s is not None else default).split()]


def extract_model(urdf_path: str) -> dict:
    """Parse a URDF and return the numeric model dict (JSON-serialisable lists)."""
    root = ET.parse(urdf_path).getroot()
    links = []     # dicts: name, I (6x6)
    for raw in root.findall(".//link"):
        org = raw.find("origin")
        xyz = _floats(org.get("xyz") if org is not None else None)
        inertial = raw.find("inertial")
        if inertial is None:
            mass, I3 = 0.0, np.zeros((3, 3))
        else:
            mass = float(inertial.find("mass").get("value", "0"))
            ri = inertial.find("inertia")
            g = lambda k: float(ri.get(k, "0"))
            I3 = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
        cx = _snap12(_skew(*xyz))
        mc = mass * cx
        I6 = np.zeros((6, 6))
        I6[:3, :3] = I3 + mc @ cx.T
        I6[:3, 3:] = mc
        I6[3:, :3] = mc.T
        I6[3:, 3:] = mass * np.eye(3)
        I6[np.isclose(I6, 0.0, 1e-10, 1e-10)] = 0.0
        links.append({"name": raw.get("name"), "I": I6})
    joints = []
    for raw in root.findall(".//joint"):
        org = raw.find("origin")
        xyz = _floats(org.get("xyz"))
        rpy = _floats(org.get("rpy"))
        jtype = raw.get("type")
        E = _rx(rpy[0]) @ _ry(rpy[1]) @ _rz(rpy[2])
        Xfix = _rot6(E) @ _xlt(_skew(*xyz))
        Hfix_R, Hfix_t = E, np.array(xyz)
        ax = raw.find("axis")
        H0 = np.eye(4); Ha = np.zeros((4, 4)); Hb = np.zeros((4, 4))
        if jtype == "revolute":
            a = _axis_index(_floats(ax.get("xyz")))
            E0, Ec, Es = _REV[a]
            X0, Xa, Xb = _rot6(E0) @ Xfix, _rot6(Ec) @ Xfix, _rot6(Es) @ Xfix
            S = np.zeros(6); S[a] = 1.0
            H0[:3, :3] = (E0 @ Hfix_R).T; Ha[:3, :3] = (Ec @ Hfix_R).T; Hb[:3, :3] = (Es @ Hfix_R).T
            H0[:3, 3] = Hfix_t
        elif jtype == "prismatic":
            a = _axis_index(_floats(ax.get("xyz")))
            e = np.zeros(3); e[a] = 1.0
            X0 = Xfix.copy()
            L = np.zeros((6, 6)); L[3:, :3] = -_skew(*e)
            Xa = L @ Xfix
            Xb = np.zeros((6, 6))
            S = np.zeros(6); S[3 + a] = 1.0
            H0[:3, :3] = Hfix_R.T; H0[:3, 3] = Hfix_t
            Ha[:3, 3] = e
        elif jtype == "fixed":
            X0, Xa, Xb = Xfix.copy(), np.zeros((6, 6)), np.zeros((6, 6))
            S = np.zeros(6)
            H0[:3, :3] = Hfix_R.T; H0[:3, 3] = Hfix_t
        else:
            raise ValueError("only revolute, prismatic and fixed joints are supported (Joint.py:52-85), got %r" % jtype)
        dyn = raw.find("dynamics")
        joints.append({
            "name": raw.get("name"), "jtype": jtype, "parent": raw.find("parent").get("link"),
            "child": raw.find("child").get("link"),
            "X0": _snap6(X0), "Xa": _snap6(Xa), "Xb": _snap6(Xb),
            "H0": _snap6(H0), "Ha": _snap6(Ha), "Hb": _snap6(Hb), "S": S,
            "damping": float(dyn.get("damping")) if dyn is not None and dyn.get("damping") else 0.0})

    # fold fixed joints into their parents (URDFParser.remove_fixed_joints :330-351)
    by_name = {l["name"]: l for l in links}
    for fj in [j for j in joints if j["jtype"] == "fixed"]:
        Xf = fj["X0"]
        for gc in joints:
            if gc["parent"] == fj["child"]:
                gc["parent"] = fj["parent"]
                for key in ("X0", "Xa", "Xb"):
                    gc[key] = gc[key] @ Xf
        by_name[fj["parent"]]["I"] = by_name[fj["parent"]]["I"] + Xf.T @ by_name[fj["child"]]["I"] @ Xf
        joints.remove(fj)
        links.remove(by_name[fj["child"]])

    # DFS renumbering (URDFParser.dfs_order_update :374-393): children of a link in URDF joint order
    children = {j["child"] for j in joints}
    roots = [l["name"] for l in links if l["name"] not in children]
    if len(roots) != 1:
        raise ValueError("URDF must have exactly one root link, found %r" % roots)
    order, parent_of = [], []

    def dfs(link_name, parent_id):
        for j in joints:
            if j["parent"] == link_name:
                jid = len(order)
                order.append(j)
                parent_of.append(parent_id)
                dfs(j["child"], jid)
    dfs(roots[0], -1)
    if len(order) != len(joints):
        raise ValueError("URDF joints do not form a tree rooted at %r (e.g. the reference's malformed models/arm6.urdf)" % roots[0])
    n = len(order)
    model = {
        "name": root.get("name"), "n": n, "parent": parent_of,
        "jtype": [j["jtype"] for j in order],
        "S": [j["S"].tolist() for j in order],
        "damping": [j["damping"] for j in order],
    }
    for key in ("X0", "Xa", "Xb", "H0", "Ha", "Hb"):
        model[key] = [j[key].tolist() for j in order]
    model["I"] = [by_name[j["child"]]["I"].tolist() for j in order]
    return model


def builtin_urdf(name: str) -> str:
    """Path of a URDF shipped with the package ('arm1'..'arm6', 'pend')."""
    p = os.path.join(URDF_DIR, name + ".urdf")
    if not os.path.isfile(p):
        raise FileNotFoundError(p)
    return p


def model_digest(model: dict) -> str:
    import hashlib
    blob = json.dumps({k: model[k] for k in sorted(model) if k != "name"}, sort_keys=True).encode()
    return hashlib.sha1(blob).hexdigest()[:12]
