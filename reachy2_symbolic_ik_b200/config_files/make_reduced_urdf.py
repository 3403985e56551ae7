"""Derive ``reachy2_arms.urdf`` -- the two arm chains (torso -> {r,l}_arm_tip) only --
from the reference robot description ``src/config_files/reachy2.urdf``.

The reduced file keeps, for every joint on the two chains, exactly the attributes the
IK-parameter extraction (reference utils.py:661-690) and the host FK sampler read:
name, type, parent, child, origin (xyz, rpy), axis and limits.  Visuals, collisions,
inertials, transmissions, gazebo/ros2_control blocks and every other limb are dropped.

Run once in the build container (the reference checkout is not available elsewhere):
    python reachy2_symbolic_ik_b200/config_files/make_reduced_urdf.py
"""
import os
import xml.etree.ElementTree as ET

SRC = "/root/reference/src/config_files/reachy2.urdf"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reachy2_arms.urdf")


def main() -> None:
    root = ET.parse(SRC).getroot()
    by_child = {j.find("child").attrib["link"]: j for j in root.findall("joint")}
    keep, links = [], ["torso"]
    for tip in ("r_arm_tip", "l_arm_tip"):
        chain, link = [], tip
        while link != "torso":
            j = by_child[link]
            chain.append(j)
            link = j.find("parent").attrib["link"]
        for j in reversed(chain):
            keep.append(j)
            links.append(j.find("child").attrib["link"])
    out = ET.Element("robot", {"name": "reachy2_arms"})
    out.append(ET.Comment(" derived from reachy2.urdf by make_reduced_urdf.py: arm chains only "))
    for name in links:
        ET.SubElement(out, "link", {"name": name})
    for j in keep:
        e = ET.SubElement(out, "joint", {"name": j.attrib["name"], "type": j.attrib["type"]})
        ET.SubElement(e, "parent", {"link": j.find("parent").attrib["link"]})
        ET.SubElement(e, "child", {"link": j.find("child").attrib["link"]})
        o = j.find("origin")
        ET.SubElement(e, "origin", {"xyz": o.attrib["xyz"], "rpy": o.attrib["rpy"]})
        if j.find("axis") is not None:
            ET.SubElement(e, "axis", {"xyz": j.find("axis").attrib["xyz"]})
        if j.find("limit") is not None:
            lim = j.find("limit").attrib
            ET.SubElement(e, "limit", {k: lim[k] for k in ("lower", "upper") if k in lim})
    ET.indent(out)
    ET.ElementTree(out).write(DST, encoding="unicode", xml_declaration=True)
    print(f"wrote {DST}: {len(keep)} joints, {len(links)} links")


if __name__ == "__main__":
    main()
