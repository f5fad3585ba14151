"""The URDF extractor must reproduce the constants of the reference's Robot objects (tests/golden/models.json,
dumped by tests/golden/make_golden.py from the unmodified reference) bit for bit."""
import numpy as np
import pytest

from trajoptmpcreference_b200.model import extract_model, builtin_urdf


@pytest.mark.parametrize("name", ["pend", "arm1", "arm2", "arm3", "arm4", "arm6"])
def test_extractor_matches_reference(name, golden_models):
    ref = golden_models[name]
    m = extract_model(builtin_urdf(name))
    assert m["n"] == ref["n"]
    assert m["parent"] == ref["parent"]
    assert m["jtype"] == ref["jtype"]
    for key in ("S", "X0", "Xa", "Xb", "H0", "Ha", "Hb", "I"):
        assert np.array_equal(np.array(m[key]), np.array(ref[key])), key


def test_malformed_arm6_rejected(tmp_path):
    """The reference's models/arm6.urdf repeats joint5's parent/child for joint6: refuse loudly."""
    from trajoptmpcreference_b200.urdf.make_urdf import arm_urdf
    txt = arm_urdf(6).replace('<parent link="link5"/>\n    <child link="link6"/>', '<parent link="link4"/>\n    <child link="link5"/>')
    p = tmp_path / "bad.urdf"
    p.write_text(txt)
    with pytest.raises(ValueError):
        extract_model(str(p))


def test_prismatic_and_fixed_joint(tmp_path):
    urdf = """<?xml version="1.0" ?><robot name="t">
  <link name="base"/>
  <joint name="j1" type="prismatic"><parent link="base"/><child link="l1"/><origin rpy="0 0 0" xyz="0 0 0.2"/><axis xyz="1 0 0"/></joint>
  <link name="l1"><origin rpy="0 0 0" xyz="0.1 0 0"/><inertial><mass value="1.0"/><inertia ixx="0.1" ixy="0" ixz="0" iyy="0.1" iyz="0" izz="0.1"/></inertial></link>
  <joint name="jf" type="fixed"><parent link="l1"/><child link="l1b"/><origin rpy="0 0 0" xyz="0 0.3 0"/></joint>
  <link name="l1b"><origin rpy="0 0 0" xyz="0 0.1 0"/><inertial><mass value="0.5"/><inertia ixx="0.01" ixy="0" ixz="0" iyy="0.01" iyz="0" izz="0.01"/></inertial></link>
  <joint name="j2" type="revolute"><parent link="l1b"/><child link="l2"/><origin rpy="0 0 0" xyz="0 0.5 0"/><axis xyz="0 1 0"/></joint>
  <link name="l2"><origin rpy="0 0 0" xyz="0 0.25 0"/><inertial><mass value="0.3"/><inertia ixx="0.02" ixy="0" ixz="0" iyy="0.02" iyz="0" izz="0.02"/></inertial></link>
</robot>"""
    p = tmp_path / "t.urdf"
    p.write_text(urdf)
    m = extract_model(str(p))
    assert m["n"] == 2 and m["parent"] == [-1, 0] and m["jtype"] == ["prismatic", "revolute"]
    # the fixed joint's child mass is folded into link 1
    assert abs(np.array(m["I"][0])[3, 3] - 1.5) < 1e-15
    # joint 2's transform includes the folded fixed offset (0.3 + 0.5 along y)
    X = np.array(m["X0"][1]) + np.array(m["Xa"][1])       # t = 0
    assert abs(abs(X[3:, :3]).max() - 0.8) < 1e-12
