"""Writes STAND-IN PartNet-Mobility cabinets so that the reference's OpenCabinetDrawer-v1 (BASELINE.json configs[3]) can run UNCHANGED without the dataset.

The reference builds each sub-scene's cabinet from `$MS_ASSET_DIR/data/partnet_mobility/dataset/<id>/mobility_cvx.urdf` (mani_skill/utils/building/articulations/
partnet_mobility.py:22-39), a network download that is not available here.  This tool writes, for every id of the reference's metadata file
(`assets/partnet_mobility/meta/info_cabinet_{drawer,door}_train.json`: id -> scale), a URDF of the same kind -- a fixed carcass, prismatic drawers opening along -x, box
collisions, a visual named `handle_<k>` on every drawer -- whose dimensions are divided by the id's `scale` (the loader multiplies by it), so that every id yields
the same 0.8 x 0.5 x 0.9 m cabinet (two drawers and a door; the carcass of the mirror task's stand-in, maniskill_b200/envs/open_cabinet_drawer.py), differing per id
only in the handle size.  The origin is the centre of the carcass, like the dataset's models (the task lifts the cabinet by minus its lowest collision point).

usage: python tools/make_standin_partnet.py <reference>/mani_skill/assets/partnet_mobility/meta <out_asset_dir> [door | mirror]      (then MS_ASSET_DIR=<out_asset_dir>)
"""
import json
import os
import sys

W, D, H, T = 0.8, 0.5, 0.9, 0.02   # width (y), depth (x), height (z), panel thickness


def _box(name, p, half, s, visual=True, collision=True):
    size = " ".join(f"{2 * h / s:.9g}" for h in half)
    xyz = " ".join(f"{v / s:.9g}" for v in p)
    out = ""
    if visual:
        out += f'    <visual name="{name}"><origin xyz="{xyz}" rpy="0 0 0"/><geometry><box size="{size}"/></geometry></visual>\n'
    if collision:
        out += f'    <collision name="{name}"><origin xyz="{xyz}" rpy="0 0 0"/><geometry><box size="{size}"/></geometry></collision>\n'
    return out


def _inertial(mass, diag, s):
    """the loader scales lengths by `s`, hence masses by s^3 and inertias by s^5 (uniform density): written divided by that, the loaded link weighs `mass`"""
    return (f'    <inertial><mass value="{mass / s**3:.9g}"/><origin xyz="0 0 0"/><inertia ixx="{diag[0] / s**5:.9g}" iyy="{diag[1] / s**5:.9g}" '
            f'izz="{diag[2] / s**5:.9g}" ixy="0" ixz="0" iyz="0"/></inertial>\n  </link>\n')


def cabinet_urdf(model_id: str, scale: float, variant: int, layout: str = "door") -> str:
    """Three compartments: two drawers (prismatic; OpenCabinetDrawer-v1's handle_types) and, on top, a door hinged on its +y edge (revolute about z;
    OpenCabinetDoor-v1's).  Every id gets the same structure -- ten ids are in both of the reference's lists, and a batched world has one prototype."""
    # layout "mirror": exactly the two-drawer cabinet of maniskill_b200/envs/open_cabinet_drawer.py::standin_cabinet (same handle for every id) -- lets a test
    # compare that mirror task with the reference's module on identical geometry
    s = scale
    z0 = -H / 2
    out = f'<?xml version="1.0"?>\n<robot name="partnet_{model_id}">\n  <link name="base"/>\n'
    out += '  <link name="link_3">\n'
    for name, p, half in (("back", (D / 2 - T / 2, 0, 0), (T / 2, W / 2, H / 2)), ("left", (0, W / 2 - T / 2, 0), (D / 2, T / 2, H / 2)),
                          ("right", (0, -W / 2 + T / 2, 0), (D / 2, T / 2, H / 2)), ("top", (0, 0, H / 2 - T / 2), (D / 2, W / 2, T / 2)),
                          ("bottom", (0, 0, z0 + T / 2), (D / 2, W / 2, T / 2))):
        out += _box(name, p, half, s)
    out += _inertial(20.0, (1.0, 1.0, 1.0), s)
    out += '  <joint name="joint_3" type="fixed"><parent link="base"/><child link="link_3"/><origin xyz="0 0 0" rpy="0 0 0"/></joint>\n'
    n_comp = 2 if layout == "mirror" else 3
    dh = (H - (n_comp + 1) * T) / n_comp
    hw = 0.08 if layout == "mirror" else 0.06 + 0.005 * (variant % 5)     # the handle bar's half width differs between the models
    for i in range(n_comp):
        zc = z0 + T + dh / 2 + i * (dh + T)
        out += f'  <link name="link_{i}">\n'
        if i == 2:   # (only in the three-compartment layout)
            # the link frame sits on the hinge line (front face, +y edge); the panel extends towards -y
            wd = W / 2 - T - 0.005
            out += _box("front", (0, -wd, 0), (T / 2, wd, dh / 2 - 0.005), s)
            out += _box(f"handle_{i}", (-0.03 - T / 2, -2 * wd + 0.06, 0), (0.015, 0.012, hw), s)
            out += _inertial(3.0, (0.06, 0.05, 0.1), s)
            out += (f'  <joint name="joint_{i}" type="revolute"><parent link="link_3"/><child link="link_{i}"/>'
                    f'<origin xyz="{(-D / 2 + T / 2) / s:.9g} {wd / s:.9g} {zc / s:.9g}" rpy="0 0 0"/><axis xyz="0 0 -1"/><limit lower="0" upper="1.57" effort="0" velocity="0"/></joint>\n')
            continue
        out += _box("front", (-D / 2 + T / 2, 0, 0), (T / 2, W / 2 - T - 0.005, dh / 2 - 0.005), s)
        out += _box("tray", (0.0, 0, -dh / 2 + T), (D / 2 - T, W / 2 - 2 * T, T / 2), s)
        out += _box(f"handle_{i}", (-D / 2 - 0.03, 0, 0), (0.015, hw, 0.012), s)
        out += _inertial(3.0, (0.06, 0.05, 0.1), s)
        out += (f'  <joint name="joint_{i}" type="prismatic"><parent link="link_3"/><child link="link_{i}"/><origin xyz="0 0 {zc / s:.9g}" rpy="0 0 0"/>'
                f'<axis xyz="-1 0 0"/><limit lower="0" upper="{0.35 / s:.9g}" effort="0" velocity="0"/></joint>\n')
    return out + "</robot>\n"


def main(meta_dir: str, out_dir: str, layout: str = "door"):
    """`meta_dir`: the reference's assets/partnet_mobility/meta.  The drawer AND the door list are written: the task's asset check wants the whole
    `partnet_mobility_cabinet` group present (mani_skill/utils/assets/data.py:93-95); every id gets the same two-drawer + one-door cabinet."""
    doors = json.load(open(os.path.join(meta_dir, "info_cabinet_door_train.json")))
    drawers = json.load(open(os.path.join(meta_dir, "info_cabinet_drawer_train.json")))
    meta = {**doors, **drawers}
    for k, (model_id, info) in enumerate(sorted(meta.items())):
        d = os.path.join(out_dir, "data", "partnet_mobility", "dataset", str(model_id))
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "mobility_cvx.urdf"), "w") as f:
            f.write(cabinet_urdf(str(model_id), float(info["scale"]), k, layout))
    return len(meta)


if __name__ == "__main__":
    n = main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "door")
    print(f"wrote {n} stand-in cabinets under {sys.argv[2]}/data/partnet_mobility/dataset")
