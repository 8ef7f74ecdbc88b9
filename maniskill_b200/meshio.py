"""Mesh files -> point sets / triangle meshes, and convex-hull cooking (asset ingestion).

What the reference gets from sapien when a builder calls `add_convex_collision_from_file` / `add_multiple_convex_collisions_from_file`
/ `add_visual_from_file` (mani_skill/utils/building/actor_builder.py:104-164) or the URDF loader meets a `<mesh filename=...>`
(mani_skill/utils/building/urdf_loader.py:23-47): STL (binary / ascii), OBJ (with `o` / `g` parts for the multi-convex case) and
binary glTF (`.glb`: positions + indices of every primitive, node transforms applied).  `cook_hull` is the convex cooking step: the hull
of the points, decimated to at most 64 vertices (the limit PhysX's GPU convex meshes have), triangles wound outwards.
"""
from __future__ import annotations

import json
import os
import struct
from typing import List, Tuple

import numpy as np

MAX_HULL_VERTS = 64


# ------------------------------------------------------------------------------------------------ readers
def load_stl(path: str) -> Tuple[np.ndarray, np.ndarray]:
    d = open(path, "rb").read()
    head = d[:512].lstrip()
    if head.startswith(b"solid") and b"facet" in d[:4000]:
        verts = []
        for line in d.decode("ascii", "ignore").splitlines():
            line = line.strip()
            if line.startswith("vertex"):
                verts.append([float(v) for v in line.split()[1:4]])
        v = np.array(verts, dtype=np.float64).reshape(-1, 3)
    else:
        n = struct.unpack("<I", d[80:84])[0]
        arr = np.frombuffer(d[84:84 + n * 50], dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
        v = arr["v"].reshape(-1, 3).astype(np.float64)
    return v, np.arange(len(v), dtype=np.int64).reshape(-1, 3)


def load_obj_parts(path: str) -> List[Tuple[np.ndarray, np.ndarray]]:
    """-> [(vertices, triangles)] per `o` / `g` group (one part when the file has none)."""
    verts: List[List[float]] = []
    parts: List[List[List[int]]] = [[]]
    with open(path, "r", errors="ignore") as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                verts.append([float(t[1]), float(t[2]), float(t[3])])
            elif t[0] in ("o", "g"):
                if parts[-1]:
                    parts.append([])
            elif t[0] == "f":
                idx = [int(s.split("/")[0]) for s in t[1:]]
                idx = [i - 1 if i > 0 else len(verts) + i for i in idx]
                for k in range(1, len(idx) - 1):
                    parts[-1].append([idx[0], idx[k], idx[k + 1]])
    V = np.asarray(verts, dtype=np.float64).reshape(-1, 3)
    out = []
    for faces in parts:
        if not faces:
            continue
        F = np.asarray(faces, dtype=np.int64)
        used, inv = np.unique(F.reshape(-1), return_inverse=True)
        out.append((V[used], inv.reshape(-1, 3)))
    if not out and len(V):
        out.append((V, np.zeros((0, 3), dtype=np.int64)))
    return out


_GLTF_DTYPE = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_GLTF_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


def _quat_xyzw_to_mat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def load_glb_parts(path: str) -> List[Tuple[np.ndarray, np.ndarray, Tuple[float, float, float, float]]]:
    """Binary glTF -> [(vertices, triangles, base colour rgba)] per mesh primitive, node transforms of the default scene applied."""
    d = open(path, "rb").read()
    magic, version, length = struct.unpack("<III", d[:12])
    if magic != 0x46546C67:
        raise RuntimeError(f"{path} is not a binary glTF file")
    off, js, binary = 12, None, b""
    while off < length:
        clen, ctype = struct.unpack("<II", d[off:off + 8])
        chunk = d[off + 8:off + 8 + clen]
        if ctype == 0x4E4F534A:
            js = json.loads(chunk.decode("utf-8"))
        elif ctype == 0x004E4942:
            binary = chunk
        off += 8 + clen

    def accessor(i):
        a = js["accessors"][i]
        bv = js["bufferViews"][a["bufferView"]]
        dt, nc = _GLTF_DTYPE[a["componentType"]], _GLTF_NCOMP[a["type"]]
        start = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
        stride = bv.get("byteStride", 0)
        item = np.dtype(dt).itemsize * nc
        if stride and stride != item:
            raw = np.frombuffer(binary, dtype=np.uint8, count=stride * (a["count"] - 1) + item, offset=start)
            rows = np.lib.stride_tricks.as_strided(raw, shape=(a["count"], item), strides=(stride, 1))
            return np.ascontiguousarray(rows).view(dt).reshape(a["count"], nc)
        return np.frombuffer(binary, dtype=dt, count=a["count"] * nc, offset=start).reshape(a["count"], nc)

    def node_matrix(n):
        if "matrix" in n:
            return np.asarray(n["matrix"], dtype=np.float64).reshape(4, 4).T
        M = np.eye(4)
        if "scale" in n:
            M[:3, :3] = np.diag(n["scale"])
        if "rotation" in n:
            M[:3, :3] = _quat_xyzw_to_mat(n["rotation"]) @ M[:3, :3]
        if "translation" in n:
            M[:3, 3] = n["translation"]
        return M

    out = []

    def visit(ni, parent):
        n = js["nodes"][ni]
        M = parent @ node_matrix(n)
        if "mesh" in n:
            for prim in js["meshes"][n["mesh"]]["primitives"]:
                if prim.get("mode", 4) != 4 or "POSITION" not in prim["attributes"]:
                    continue
                v = accessor(prim["attributes"]["POSITION"]).astype(np.float64)
                f = accessor(prim["indices"]).astype(np.int64).reshape(-1, 3) if "indices" in prim else np.arange(len(v), dtype=np.int64).reshape(-1, 3)
                color = (0.8, 0.8, 0.8, 1.0)
                if "material" in prim:
                    pbr = js["materials"][prim["material"]].get("pbrMetallicRoughness", {})
                    if "baseColorFactor" in pbr:
                        color = tuple(float(c) for c in pbr["baseColorFactor"])
                out.append((v @ M[:3, :3].T + M[:3, 3], f, color))
        for c in n.get("children", []):
            visit(c, M)

    scene = js["scenes"][js.get("scene", 0)] if js.get("scenes") else {"nodes": list(range(len(js.get("nodes", []))))}
    for ni in scene["nodes"]:
        visit(ni, np.eye(4))
    return out


def load_dae_parts(path: str) -> List[Tuple[np.ndarray, np.ndarray, Tuple[float, float, float, float]]]:
    """COLLADA 1.4 (`.dae`; the Fetch robot's visual meshes): `<triangles>` / `<polylist>` / `<polygons>` of every `<geometry>` instantiated by the
    visual scene, node transforms (`<matrix>`, `<translate>`, `<rotate>`, `<scale>`) applied, `<unit meter>` and `<up_axis>` honoured (result is z-up,
    metres); colour = the bound material's phong / lambert diffuse colour where it is a plain colour (textures: grey)."""
    import xml.etree.ElementTree as ET
    root = ET.parse(path).getroot()
    ns = root.tag[:root.tag.index("}") + 1] if root.tag.startswith("{") else ""
    q = lambda tag: ns + tag
    unit = root.find(f"{q('asset')}/{q('unit')}")
    meter = float(unit.get("meter", 1.0)) if unit is not None else 1.0
    up = (root.findtext(f"{q('asset')}/{q('up_axis')}") or "Y_UP").strip()
    floats = lambda text: np.array(text.split(), dtype=np.float64)

    effects = {}
    for e in root.iter(q("effect")):
        col = next((c for d in e.iter(q("diffuse")) for c in d.findall(q("color"))), None)
        if col is not None:
            effects[e.get("id")] = tuple(floats(col.text)[:4])
    materials = {m.get("id"): effects.get(m.find(q("instance_effect")).get("url", "#")[1:]) for m in root.iter(q("material")) if m.find(q("instance_effect")) is not None}

    geoms = {}
    for g in root.iter(q("geometry")):
        mesh = g.find(q("mesh"))
        if mesh is None:
            continue
        sources = {}
        for src in mesh.findall(q("source")):
            fa = src.find(q("float_array"))
            acc = src.find(f"{q('technique_common')}/{q('accessor')}")
            if fa is not None and fa.text:
                stride = int(acc.get("stride", 3)) if acc is not None else 3
                sources[src.get("id")] = floats(fa.text).reshape(-1, stride)
        vert_src = {}
        for vs in mesh.findall(q("vertices")):
            for inp in vs.findall(q("input")):
                if inp.get("semantic") == "POSITION":
                    vert_src[vs.get("id")] = inp.get("source")[1:]
        prims = []
        for kind in ("triangles", "polylist", "polygons"):
            for prim in mesh.findall(q(kind)):
                inputs = prim.findall(q("input"))
                n_off = max(int(i.get("offset", 0)) for i in inputs) + 1
                vin = next((i for i in inputs if i.get("semantic") == "VERTEX"), None)
                if vin is None:
                    continue
                pos = sources.get(vert_src.get(vin.get("source")[1:], ""))
                if pos is None:
                    continue
                off = int(vin.get("offset", 0))
                tris = []
                if kind == "polygons":
                    polys = [np.array(pe.text.split(), dtype=np.int64).reshape(-1, n_off)[:, off] for pe in prim.findall(q("p")) if pe.text]
                else:
                    pe = prim.find(q("p"))
                    if pe is None or not pe.text:
                        continue
                    idx = np.array(pe.text.split(), dtype=np.int64).reshape(-1, n_off)[:, off]
                    if kind == "triangles":
                        polys = list(idx.reshape(-1, 3))
                    else:
                        vc = np.array(prim.find(q("vcount")).text.split(), dtype=np.int64)
                        ends = np.cumsum(vc)
                        polys = [idx[e - c:e] for c, e in zip(vc, ends)]
                for poly in polys:
                    for k in range(1, len(poly) - 1):   # fan
                        tris.append((poly[0], poly[k], poly[k + 1]))
                if tris:
                    prims.append((pos[:, :3], np.array(tris, dtype=np.int64), prim.get("material")))
        geoms[g.get("id")] = prims

    out = []

    def emit(gid, M, bind):
        for pos, tris, mat in geoms.get(gid, []):
            used, inv = np.unique(tris.reshape(-1), return_inverse=True)
            v = pos[used] @ M[:3, :3].T + M[:3, 3]
            v = v * meter
            if up == "Y_UP":
                v = np.stack([v[:, 0], -v[:, 2], v[:, 1]], axis=1)
            elif up == "X_UP":
                v = np.stack([-v[:, 1], v[:, 0], v[:, 2]], axis=1)
            col = materials.get(bind.get(mat, mat)) or (0.8, 0.8, 0.8, 1.0)
            col = tuple(float(c) for c in (list(col) + [1.0])[:4])
            out.append((v.astype(np.float32), inv.reshape(-1, 3).astype(np.int32), col))

    def node_matrix(n):
        M = np.eye(4)
        for ch in n:
            tag = ch.tag[len(ns):]
            if tag == "matrix":
                M = M @ floats(ch.text).reshape(4, 4)
            elif tag == "translate":
                T = np.eye(4)
                T[:3, 3] = floats(ch.text)[:3]
                M = M @ T
            elif tag == "scale":
                M = M @ np.diag(list(floats(ch.text)[:3]) + [1.0])
            elif tag == "rotate":
                x, y, z, deg = floats(ch.text)[:4]
                a = np.array([x, y, z])
                nrm = np.linalg.norm(a)
                if nrm > 0:
                    a = a / nrm
                    t = np.deg2rad(deg)
                    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
                    R = np.eye(4)
                    R[:3, :3] = np.eye(3) + np.sin(t) * K + (1 - np.cos(t)) * (K @ K)
                    M = M @ R
        return M

    def visit(n, parent):
        M = parent @ node_matrix(n)
        for ig in n.findall(q("instance_geometry")):
            bind = {im.get("symbol"): im.get("target", "#")[1:] for im in ig.iter(q("instance_material"))}
            emit(ig.get("url", "#")[1:], M, bind)
        for c in n.findall(q("node")):
            visit(c, M)

    scenes = list(root.iter(q("visual_scene")))
    for vs in scenes:
        for n in vs.findall(q("node")):
            visit(n, np.eye(4))
    if not out:   # no scene graph: every geometry as stored
        for gid in geoms:
            emit(gid, np.eye(4), {})
    return out


_MESH_CACHE: dict = {}
_HULL_CACHE: dict = {}


def load_mesh_parts(path: str) -> List[Tuple[np.ndarray, np.ndarray, tuple]]:
    """-> [(vertices [n,3], triangles [m,3], base colour)] for `.stl`, `.obj`, `.glb`, `.dae` (vertex coordinates as stored, node transforms applied).
    Cached per (file, mtime): the reference builds every sub-scene from the same files (N x the Panda's 20 meshes); the arrays are read-only."""
    if not os.path.exists(path):
        raise RuntimeError(f"mesh file {path} does not exist")
    key = (os.path.abspath(path), os.path.getmtime(path))
    if key not in _MESH_CACHE:
        parts = _load_mesh_parts(path)
        for v, f, _ in parts:
            v.setflags(write=False)
            f.setflags(write=False)
        _MESH_CACHE[key] = parts
    return list(_MESH_CACHE[key])


def _load_mesh_parts(path: str) -> List[Tuple[np.ndarray, np.ndarray, tuple]]:
    ext = os.path.splitext(path)[1].lower()
    if ext == ".stl":
        v, f = load_stl(path)
        return [(v, f, (0.8, 0.8, 0.8, 1.0))]
    if ext == ".obj":
        return [(v, f, (0.8, 0.8, 0.8, 1.0)) for v, f in load_obj_parts(path)]
    if ext in (".glb",):
        return load_glb_parts(path)
    if ext == ".dae":
        return load_dae_parts(path)
    raise RuntimeError(f"unsupported mesh format '{ext}' ({path}): stl, obj, glb and dae are read")


def load_points(path: str) -> np.ndarray:
    parts = load_mesh_parts(path)
    if not parts:
        raise RuntimeError(f"{path} holds no geometry")
    return np.concatenate([p[0] for p in parts])


def load_parts(path: str) -> List[np.ndarray]:
    """Vertex sets of the parts of a multi-convex file (`load_multiple`)."""
    return [p[0] for p in load_mesh_parts(path) if len(p[0]) >= 4]


# ------------------------------------------------------------------------------------------------ convex cooking
def cook_hull(points, max_verts: int = MAX_HULL_VERTS):
    """Convex hull with at most `max_verts` vertices (support points of a Fibonacci direction set, the most extreme ones kept) ->
    (vertices [n,3] float64, triangles [[i,j,k]] wound outwards)."""
    import hashlib
    from scipy.spatial import ConvexHull
    raw = np.ascontiguousarray(points, dtype=np.float64)
    key = (raw.shape, hashlib.blake2b(raw.tobytes(), digest_size=16).digest(), int(max_verts))
    hit = _HULL_CACHE.get(key)     # the same mesh is cooked once per sub-scene by the reference's builders (4096 x 9 Panda hulls)
    if hit is not None:
        return hit[0].copy(), [list(t) for t in hit[1]]
    v, tris = _cook_hull(raw, max_verts)
    _HULL_CACHE[key] = (v, tris)
    return v.copy(), [list(t) for t in tris]


def _cook_hull(points, max_verts):
    from scipy.spatial import ConvexHull
    pts = np.unique(np.round(np.asarray(points, dtype=np.float64), 7), axis=0)
    hull = ConvexHull(pts)
    v = pts[hull.vertices]
    if len(v) > max_verts:
        n_dir = 4096
        i = np.arange(n_dir) + 0.5
        phi = np.arccos(1 - 2 * i / n_dir)
        th = np.pi * (1 + 5**0.5) * i
        dirs = np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], 1)
        sup = np.argmax(v @ dirs.T, axis=0)
        cnt = np.bincount(sup, minlength=len(v))
        v = v[np.argsort(-cnt)[:max_verts]]
        v = v[ConvexHull(v).vertices]
    hull = ConvexHull(v)
    c = v.mean(0)
    tris = []
    for s in hull.simplices:
        a, b, cc = v[s[0]], v[s[1]], v[s[2]]
        if np.dot(np.cross(b - a, cc - a), a - c) < 0:
            s = [s[0], s[2], s[1]]
        tris.append([int(s[0]), int(s[1]), int(s[2])])
    return v, tris
