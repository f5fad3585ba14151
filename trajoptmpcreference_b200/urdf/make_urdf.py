"""Generate the planar k-link arm URDFs (and the 1-DoF pendulum) used by tests and benchmarks.

Same physical convention as the reference's `models/armK.urdf` (mass 0.1, link length 1, com at 0.5,
revolute joints about z) but written by this script; `arm6` here is the *well-formed* 6-link chain
(the reference's `models/arm6.urdf:75-81` repeats joint5's parent/child for joint6).
`pend` is the 1-link arm with a y-axis joint so that gravity (along z) does work.
"""
import os

LINK = """  <link name="link{i}">
    <origin rpy="1.5707963267948966 0 0" xyz="0 0.5 0"/>
    <inertial>
      <origin rpy="1.5707963267948966 0 0" xyz="0 0.5 0"/>
      <mass value="0.1"/>
      <inertia ixx="0.008395833333333333" ixy="0.0" ixz="0.0" iyy="0.008395833333333333" iyz="0.0" izz="0.00012500000000000003"/>
    </inertial>
  </link>
"""
JOINT = """  <joint name="joint{i}" type="revolute">
    <parent link="{parent}"/>
    <child link="link{i}"/>
    <origin rpy="0 0 0" xyz="{xyz}"/>
    <axis xyz="{axis}"/>
  </joint>
"""


def arm_urdf(k: int, axis: str = "0 0 1", name: str = None) -> str:
    out = ['<?xml version="1.0" ?>\n', '<robot name="%s">\n' % (name or ("%d_link" % k)), '  <link name="base_link"/>\n']
    for i in range(1, k + 1):
        parent = "base_link" if i == 1 else "link%d" % (i - 1)
        xyz = "0 0 0" if i == 1 else "0 1 0"
        out.append(JOINT.format(i=i, parent=parent, xyz=xyz, axis=axis))
        out.append(LINK.format(i=i))
    out.append("</robot>\n")
    return "".join(out)


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for k in range(1, 7):
        with open(os.path.join(here, "arm%d.urdf" % k), "w") as f:
            f.write(arm_urdf(k))
    with open(os.path.join(here, "pend.urdf"), "w") as f:
        f.write(arm_urdf(1, axis="0 1 0", name="pendulum"))


if __name__ == "__main__":
    main()


CARTPOLE = """<?xml version="1.0" ?>
<robot name="cartpole">
  <link name="world"/>
  <joint name="slider" type="prismatic">
    <parent link="world"/>
    <child link="cart"/>
    <origin rpy="0 0 0" xyz="0 0 0"/>
    <axis xyz="1 0 0"/>
  </joint>
  <link name="cart">
    <origin rpy="0 0 0" xyz="0 0 0"/>
    <inertial>
      <origin rpy="0 0 0" xyz="0 0 0"/>
      <mass value="1.0"/>
      <inertia ixx="0.01" ixy="0.0" ixz="0.0" iyy="0.01" iyz="0.0" izz="0.01"/>
    </inertial>
  </link>
  <joint name="hinge" type="revolute">
    <parent link="cart"/>
    <child link="pole"/>
    <origin rpy="0 0 0" xyz="0 0 0"/>
    <axis xyz="0 1 0"/>
  </joint>
  <link name="pole">
    <origin rpy="0 0 0" xyz="0 0 -0.5"/>
    <inertial>
      <origin rpy="0 0 0" xyz="0 0 -0.5"/>
      <mass value="0.2"/>
      <inertia ixx="0.016666666666666666" ixy="0.0" ixz="0.0" iyy="0.016666666666666666" iyz="0.0" izz="0.0005"/>
    </inertial>
  </link>
</robot>
"""


def write_cartpole():
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "cartpole.urdf"), "w") as f:
        f.write(CARTPOLE)


if __name__ == "__main__":
    write_cartpole()
