"""Writes a bp5-style robot description from the 40 numbers of the compact model (include/irrl_b200.h: irrl_parse_urdf) -- the test-side
inverse of csrc/urdf_reader.h.  Only the elements RaiSim / the reader use: inertials, joint origins / axes / dynamics, collision shapes."""
import numpy as np


def F(v):
    return repr(float(v))

NAMES = ("I0", 3), ("I1", 3), ("I2", 4), ("I3", 3), ("rotor", 3), ("off1x", 1), ("off1y", 1), ("off2y", 1), ("toe_z", 1), ("toe_r", 1), ("box_half", 3), \
        ("joint_damping", 1), ("m0", 1), ("com0", 3), ("m1", 1), ("com1", 3), ("m2", 1), ("com2", 3), ("m3", 1), ("com3z", 1), ("knee_z", 1)


def unpack(v):
    out, k = {}, 0
    for n, c in NAMES:
        out[n] = float(v[k]) if c == 1 else np.asarray(v[k:k + c], np.float64)
        k += c
    return out


def write_urdf(path, model, shank_mass=0.064, toe_mass=None, toe_inertia=0.000025, tweak=None):
    """model: dict from unpack().  The shank + toe body of the compact model is split back into a shank link and a toe link with the given
    shank mass (toe = rest); their inertias are chosen so that the merge of the reader reproduces I3 exactly.  tweak(leg, part, d) may edit
    the per-leg numbers (for asymmetry tests)."""
    m = model
    toe_mass = m["m3"] - shank_mass if toe_mass is None else toe_mass
    zt = m["toe_z"]; zs = (m["m3"] * m["com3z"] - toe_mass * zt) / shank_mass
    zc = m["com3z"]; ds, dt = zs - zc, zt - zc
    sh_I = (m["I3"][0] - shank_mass * ds * ds - toe_inertia - toe_mass * dt * dt, m["I3"][1] - shank_mass * ds * ds - toe_inertia - toe_mass * dt * dt, m["I3"][2] - toe_inertia)

    def inertial(mass, com, I):
        return (f'<inertial><mass value="{F(mass)}"/><!-- <mass value="99"/> --><origin xyz="{F(com[0])} {F(com[1])} {F(com[2])}"/>'
                f'<inertia ixx="{F(I[0])}" ixy="{F(I[1])}" ixz="{F(I[2])}" iyy="{F(I[3])}" iyz="{F(I[4])}" izz="{F(I[5])}"/></inertial>')

    x = ['<?xml version="1.0"?>', '<robot name="bp5_test">',
         f'<link name="body">{inertial(m["m0"], m["com0"], (m["I0"][0], 0.0, 0.0, m["I0"][1], 0.0, m["I0"][2]))}'
         f'<collision><origin rpy="0 0 0" xyz="0 0 0"/><geometry><box size="{F(2 * m["box_half"][0])} {F(2 * m["box_half"][1])} {F(2 * m["box_half"][2])}"/></geometry></collision></link>']
    for leg, sx, sy in (("fr", 1, -1), ("fl", 1, 1), ("hr", -1, -1), ("hl", -1, 1)):
        d = dict(off1=(sx * m["off1x"], sy * m["off1y"], 0.0), off2=(0.0, sy * m["off2y"], 0.0), m1=m["m1"], com1=(sx * m["com1"][0], sy * m["com1"][1], m["com1"][2]),
                 m2=m["m2"], com2=(sx * m["com2"][0], sy * m["com2"][1], m["com2"][2]), iyz2=-sy * m["I2"][3])
        if tweak:
            tweak(leg, d)
        dyn = lambda r: f'<limit effort="18" lower="-6.28" upper="6.28" velocity="40"/><dynamics damping="{F(m["joint_damping"])}" rotor_inertia="{F(r)}"/>'
        x += [f'<joint name="torso_to_abduct_{leg}_j" type="continuous"><axis xyz="1 0 0"/><origin rpy="0 0 0" xyz="{F(d["off1"][0])} {F(d["off1"][1])} {F(d["off1"][2])}"/>'
              f'<parent link="body"/><child link="abduct_{leg}"/>{dyn(m["rotor"][0])}</joint>',
              f'<link name="abduct_{leg}">{inertial(d["m1"], d["com1"], (m["I1"][0], 0.0, 0.0, m["I1"][1], 0.0, m["I1"][2]))}</link>',
              f'<joint name="abduct_{leg}_to_thigh_{leg}_j" type="continuous"><axis xyz="0 -1 0"/><origin rpy="0 0 0" xyz="{F(d["off2"][0])} {F(d["off2"][1])} {F(d["off2"][2])}"/>'
              f'<parent link="abduct_{leg}"/><child link="thigh_{leg}"/>{dyn(m["rotor"][1])}</joint>',
              f'<link name="thigh_{leg}">{inertial(d["m2"], d["com2"], (m["I2"][0], 0.0, 0.0, m["I2"][1], d["iyz2"], m["I2"][2]))}</link>',
              f'<joint name="thigh_{leg}_to_knee_{leg}_j" type="continuous"><axis xyz="0 -1 0"/><origin rpy="0 0 0" xyz="0.0 0.0 {F(m["knee_z"])}"/>'
              f'<parent link="thigh_{leg}"/><child link="shank_{leg}"/>{dyn(m["rotor"][2])}</joint>',
              f'<link name="shank_{leg}">{inertial(shank_mass, (0.0, 0.0, zs), (sh_I[0], 0.0, 0.0, sh_I[1], 0.0, sh_I[2]))}</link>',
              f'<link name="toe_{leg}"><collision><origin rpy="0 0 0" xyz="0 0 0"/><geometry><sphere radius="{F(m["toe_r"])}"/></geometry></collision>'
              f'<inertial><mass value="{F(toe_mass)}"/><inertia ixx="{F(toe_inertia)}" ixy="0" ixz="0" iyy="{F(toe_inertia)}" iyz="0" izz="{F(toe_inertia)}"/></inertial></link>',
              f'<joint name="toe_{leg}_joint" type="fixed"><parent link="shank_{leg}"/><child link="toe_{leg}"/><origin xyz="0 0 {F(zt)}"/><dynamics damping="0.0" friction="0.0"/></joint>']
    x.append("</robot>")
    open(path, "w").write("\n".join(x))
