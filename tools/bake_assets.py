#!/usr/bin/env python
"""Bake robot descriptions into compact JSON model files.

Reads a URDF (+SRDF, + collision meshes) from the read-only reference checkout and writes
``maniskill_b200/assets/robots/<name>.json`` -- *derived data only* (link tree, inertials, joint frames,
limits, collision primitives and <=64-vertex convex hulls, the bound PhysX-GPU cooking also enforces).
The GPU box has no /root/reference, so everything the simulator needs at run time must be in these files.

Usage:  python tools/bake_assets.py            (run once here; outputs are committed)

Reference inputs: mani_skill/assets/robots/panda/panda_v2.urdf, panda_v3.urdf, panda_v2.srdf,
mani_skill/assets/robots/fetch/fetch.urdf, fetch.srdf and their collision meshes.
"""
import json
import os
import struct
import sys
import xml.etree.ElementTree as ET

import numpy as np
from scipy.spatial import ConvexHull

REF = "/root/reference/mani_skill/assets/robots"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "maniskill_b200", "assets", "robots")
MAX_HULL_VERTS = 64


def rpy_to_mat(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def mat_to_quat(R):
    # wxyz, robust branch
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = [(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s]
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = [(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s]
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = [(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s]
    q = np.array(q)
    if q[0] < 0:
        q = -q
    return (q / np.linalg.norm(q)).tolist()


def parse_origin(el):
    xyz = [0.0, 0.0, 0.0]
    rpy = [0.0, 0.0, 0.0]
    if el is not None:
        o = el.find("origin")
        if o is not None:
            if o.get("xyz"):
                xyz = [float(v) for v in o.get("xyz").split()]
            if o.get("rpy"):
                rpy = [float(v) for v in o.get("rpy").split()]
    return xyz, mat_to_quat(rpy_to_mat(*rpy)), rpy_to_mat(*rpy)


def load_stl(path):
    d = open(path, "rb").read()
    head = d[:512].lstrip()
    if head.startswith(b"solid") and b"facet" in d[:2000]:
        verts = []
        for line in d.decode("ascii", "ignore").splitlines():
            line = line.strip()
            if line.startswith("vertex"):
                verts.append([float(v) for v in line.split()[1:4]])
        return np.array(verts, dtype=np.float64)
    n = struct.unpack("<I", d[80:84])[0]
    arr = np.frombuffer(d[84 : 84 + n * 50], dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
    return arr["v"].reshape(-1, 3).astype(np.float64)


def cook_hull(points, max_verts=MAX_HULL_VERTS):
    """Convex hull with at most ``max_verts`` vertices (directional-support decimation like a cooking step)."""
    pts = np.unique(np.round(points, 7), axis=0)
    hull = ConvexHull(pts)
    v = pts[hull.vertices]
    if len(v) > max_verts:
        # pick support points of a Fibonacci direction set, grow until we hit the budget
        chosen = set()
        k = max_verts
        n_dir = 4096
        i = np.arange(n_dir) + 0.5
        phi = np.arccos(1 - 2 * i / n_dir)
        th = np.pi * (1 + 5**0.5) * i
        dirs = np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], 1)
        sup = np.argmax(v @ dirs.T, axis=0)
        # count how often each vertex is a support point; keep the most "extreme" ones
        cnt = np.bincount(sup, minlength=len(v))
        order = np.argsort(-cnt)
        keep = order[:k]
        v = v[keep]
        hull2 = ConvexHull(v)
        v = v[hull2.vertices]
    hull = ConvexHull(v)
    # outward-oriented triangles
    c = v.mean(0)
    tris = []
    for s in hull.simplices:
        a, b, cc = v[s[0]], v[s[1]], v[s[2]]
        nrm = np.cross(b - a, cc - a)
        if np.dot(nrm, a - c) < 0:
            s = [s[0], s[2], s[1]]
        tris.append([int(s[0]), int(s[1]), int(s[2])])
    return v, tris


def bake(name, urdf_path, srdf_path, mesh_root):
    root = ET.parse(urdf_path).getroot()
    links = {}
    for l in root.findall("link"):
        lname = l.get("name")
        inertial = l.find("inertial")
        mass = 0.0
        com = [0.0, 0.0, 0.0]
        I = np.zeros((3, 3))
        if inertial is not None:
            mass = float(inertial.find("mass").get("value"))
            com, _, Rin = parse_origin(inertial)
            ie = inertial.find("inertia")
            g = lambda k: float(ie.get(k, 0.0))
            Il = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
            I = Rin @ Il @ Rin.T
        shapes = []
        for c in l.findall("collision"):
            xyz, quat, _ = parse_origin(c)
            geo = c.find("geometry")
            if geo.find("box") is not None:
                size = [float(v) / 2 for v in geo.find("box").get("size").split()]
                shapes.append(dict(type="box", p=xyz, q=quat, half_size=size))
            elif geo.find("sphere") is not None:
                shapes.append(dict(type="sphere", p=xyz, q=quat, radius=float(geo.find("sphere").get("radius"))))
            elif geo.find("cylinder") is not None:
                cy = geo.find("cylinder")
                shapes.append(dict(type="cylinder", p=xyz, q=quat, radius=float(cy.get("radius")), half_length=float(cy.get("length")) / 2))
            elif geo.find("mesh") is not None:
                m = geo.find("mesh")
                scale = [float(v) for v in m.get("scale", "1 1 1").split()]
                pts = load_stl(os.path.join(mesh_root, m.get("filename"))) * np.array(scale)
                v, tris = cook_hull(pts)
                shapes.append(dict(type="convex", p=xyz, q=quat, vertices=np.round(v, 7).tolist(), triangles=tris))
        links[lname] = dict(name=lname, mass=mass, com=com, inertia=[I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]], collisions=shapes)
    joints = []
    children = {}
    child_set = set()
    for j in root.findall("joint"):
        jt = j.get("type")
        xyz, quat, _ = parse_origin(j)
        axis = [1.0, 0.0, 0.0]
        if j.find("axis") is not None:
            axis = [float(v) for v in j.find("axis").get("xyz").split()]
        lim = j.find("limit")
        lo, hi = -1e30, 1e30
        effort = 0.0
        if lim is not None:
            effort = float(lim.get("effort", 0))
            if jt in ("revolute", "prismatic"):
                lo, hi = float(lim.get("lower", 0)), float(lim.get("upper", 0))
        if jt == "continuous":
            jt = "revolute_unwrapped"
        dyn = j.find("dynamics")
        damping = float(dyn.get("damping", 0)) if dyn is not None else 0.0
        friction = float(dyn.get("friction", 0)) if dyn is not None else 0.0
        mimic = j.find("mimic")
        rec = dict(
            name=j.get("name"), type=jt, parent=j.find("parent").get("link"), child=j.find("child").get("link"),
            p=xyz, q=quat, axis=axis, lower=lo, upper=hi, effort=effort, damping=damping, friction=friction,
        )
        if mimic is not None:
            rec["mimic"] = dict(joint=mimic.get("joint"), multiplier=float(mimic.get("multiplier", 1)), offset=float(mimic.get("offset", 0)))
        joints.append(rec)
        children.setdefault(rec["parent"], []).append(rec)
        child_set.add(rec["child"])
    roots = [n for n in links if n not in child_set]
    assert len(roots) == 1, roots
    # depth-first order, children in URDF joint order (the order sapien's loader walks the tree)
    order = []
    jorder = []

    def dfs(lname, jrec):
        order.append(lname)
        jorder.append(jrec)
        for jr in children.get(lname, []):
            dfs(jr["child"], jr)

    dfs(roots[0], None)
    out_links = []
    idx = {n: i for i, n in enumerate(order)}
    for lname, jr in zip(order, jorder):
        L = dict(links[lname])
        if jr is None:
            L["parent"] = -1
            L["joint"] = dict(name="", type="fixed", p=[0, 0, 0], q=[1, 0, 0, 0], axis=[1, 0, 0], lower=0, upper=0, effort=0, damping=0, friction=0)
        else:
            L["parent"] = idx[jr["parent"]]
            L["joint"] = {k: v for k, v in jr.items() if k not in ("parent", "child")}
        out_links.append(L)
    disabled = []
    if srdf_path and os.path.exists(srdf_path):
        for d in ET.parse(srdf_path).getroot().findall("disable_collisions"):
            a, b = d.get("link1"), d.get("link2")
            if a in idx and b in idx:
                disabled.append([idx[a], idx[b]])
    model = dict(name=name, source=os.path.relpath(urdf_path, "/root/reference"), links=out_links, disabled_collision_pairs=disabled)
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(model, f, separators=(",", ":"))
    nd = sum(1 for L in out_links if L["joint"]["type"] not in ("fixed",))
    nh = sum(1 for L in out_links for s in L["collisions"] if s["type"] == "convex")
    print(f"{name}: {len(out_links)} links, {nd} dof, {nh} hulls, {len(disabled)} disabled pairs")


if __name__ == "__main__":
    bake("panda_v2", f"{REF}/panda/panda_v2.urdf", f"{REF}/panda/panda_v2.srdf", f"{REF}/panda")
    bake("panda_v3", f"{REF}/panda/panda_v3.urdf", f"{REF}/panda/panda_v3.srdf", f"{REF}/panda")
    bake("fetch", f"{REF}/fetch/fetch.urdf", f"{REF}/fetch/fetch.srdf", f"{REF}/fetch")
