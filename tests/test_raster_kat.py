"""CPU: the raster ORACLE (oracle/b2s_oracle_raster.cpp -- the checker the CUDA rasteriser must equal bit for bit in segmentation / depth, tests/test_gpu_render.py)
against an independent ray caster written here in numpy: one ray through every pixel centre (pinhole intrinsics of `RenderCameraComponent.set_fovy`, camera frame
x forward / y left / z up like mani_skill/utils/sapien_utils.py:317-366), analytic intersections with a half-space, rotated boxes (slabs), spheres and a convex
hull (half-space clipping with scipy's facets).  Nothing of the rasteriser's triangle setup, edge functions, depth keys or flat-face tests is shared.

Away from silhouettes the two must agree exactly in the segmentation id and to the millimetre quantisation in the position texture; on silhouette pixels a
scan converter and a ray caster legitimately differ by the coverage rule, so a small fraction of disagreeing pixels -- all adjacent to an id change -- is allowed."""
import numpy as np
import pytest

from maniskill_b200 import utils as U
from maniskill_b200.model import SHAPE_BOX, SHAPE_PLANE, SHAPE_SPHERE, ActorRec, SceneDesc, ShapeRec, SimParams, pose7
from maniskill_b200.render import build_visual_table, camera_desc
from oracle import raster
from oracle.oracle import OracleWorld

SHAPE_CONVEX = 4


def _qmat(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _quat(axis, angle):
    a = np.asarray(axis, dtype=float)
    a /= np.linalg.norm(a)
    return np.concatenate([[np.cos(angle / 2)], np.sin(angle / 2) * a])


def _scene():
    rng = np.random.default_rng(3)
    hull_pts = rng.normal(size=(40, 3)) * np.array([0.05, 0.03, 0.04])
    from maniskill_b200.meshio import cook_hull
    hv, ht = cook_hull(hull_pts)
    s = SceneDesc(1, SimParams())
    bodies = [
        ("ground", "static", ShapeRec(SHAPE_PLANE, pose7([0, 0, 0], [0.7071068, 0, -0.7071068, 0]), color=(0.4, 0.5, 0.4, 1)), pose7()),
        ("slab", "kinematic", ShapeRec(SHAPE_BOX, pose7(), np.array([0.25, 0.15, 0.02]), color=(0.8, 0.6, 0.3, 1)), pose7([0.0, 0.0, 0.1], _quat((0, 0, 1), 0.3))),
        ("brick", "kinematic", ShapeRec(SHAPE_BOX, pose7(), np.array([0.04, 0.03, 0.05]), color=(0.9, 0.1, 0.1, 1)), pose7([0.05, -0.06, 0.2], _quat((1, 2, 0.5), 0.9))),
        ("ball", "kinematic", ShapeRec(SHAPE_SPHERE, pose7(), np.array([0.045, 0, 0]), color=(0.1, 0.2, 0.9, 1)), pose7([-0.08, 0.07, 0.17])),
        ("rock", "kinematic", ShapeRec(SHAPE_CONVEX, pose7(), np.zeros(3), vertices=hv, triangles=np.asarray(ht), color=(0.5, 0.5, 0.5, 1)),
         pose7([0.12, 0.1, 0.19], _quat((0.3, -1, 0.2), 1.3))),
    ]
    for name, kind, shape, pose in bodies:
        s.add_actor(ActorRec(name, kind, [shape], pose))
    return s.compile(), hv


def _cast(cm, hv, body, cam):
    """-> (seg [H,W] int, position in the OpenGL camera frame [H,W,3] metres) by ray casting"""
    from scipy.spatial import ConvexHull
    H, W = cam["height"], cam["width"]
    cp = np.asarray(cam["local_pose"], dtype=float)
    Rc, tc = _qmat(cp[3:]), cp[:3]
    v, u = np.meshgrid(np.arange(H) + 0.5, np.arange(W) + 0.5, indexing="ij")
    d_cam = np.stack([np.ones_like(u), -(u - cam["cx"]) / cam["fx"], -(v - cam["cy"]) / cam["fy"]], axis=-1)      # x forward, y left, z up
    d = d_cam @ Rc.T
    best = np.full((H, W), np.inf)          # distance along the optical axis (d_cam.x = 1: the ray parameter IS the depth)
    seg = np.zeros((H, W), dtype=int)
    hull_eq = ConvexHull(hv).equations
    for vis in cm.visuals:
        if vis["hidden"]:
            continue
        row = vis["row"]
        bp = body[row][:7] if row >= 0 else np.array([0, 0, 0, 1, 0, 0, 0], dtype=float)
        Rb, tb = _qmat(bp[3:]), bp[:3]
        lp = np.asarray(vis["pose"], dtype=float)
        R, t = Rb @ _qmat(lp[3:]), tb + Rb @ lp[:3]
        o = (tc - t) @ R                     # ray origin and directions in the shape's frame
        dl = d @ R
        if vis["type"] == SHAPE_PLANE:       # the half-space x <= 0 of the shape frame (normal +x)
            tt = np.where(dl[..., 0] < -1e-12, -o[0] / np.where(dl[..., 0] < -1e-12, dl[..., 0], -1), np.inf)
        elif vis["type"] == SHAPE_SPHERE:
            r = vis["size"][0]
            b = dl @ o
            a = (dl * dl).sum(-1)
            disc = b * b - a * (o @ o - r * r)
            tt = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / a, np.inf)
        elif vis["type"] == SHAPE_BOX:
            h = np.asarray(vis["size"], dtype=float)
            with np.errstate(divide="ignore", invalid="ignore"):
                t1, t2 = (-h - o) / dl, (h - o) / dl
            tn, tf = np.minimum(t1, t2).max(-1), np.maximum(t1, t2).min(-1)
            tt = np.where((tn < tf) & (tn > 0), tn, np.inf)
        else:                                # convex hull: clip the ray against every facet's half-space
            tn, tf = np.zeros((H, W)), np.full((H, W), np.inf)
            for eq in hull_eq:
                den = dl @ eq[:3]
                num = -(o @ eq[:3] + eq[3])
                with np.errstate(divide="ignore", invalid="ignore"):
                    tc_ = num / den
                tn = np.where(den < 0, np.maximum(tn, tc_), tn)
                tf = np.where(den > 0, np.minimum(tf, tc_), tf)
                tf = np.where((den == 0) & (num < 0), -1, tf)
            tt = np.where((tn < tf) & (tn > 0), tn, np.inf)
        tt = np.where((tt > cam["near"]) & (tt < cam["far"]), tt, np.inf)
        closer = tt < best
        best = np.where(closer, tt, best)
        seg = np.where(closer, vis["seg"], seg)
    hit = np.isfinite(best)
    depth = np.where(hit, best, 0.0)
    pos_gl = np.stack([-d_cam[..., 1] * depth, d_cam[..., 2] * depth, -depth], axis=-1)      # OpenGL camera: x right, y up, z backwards
    return seg, pos_gl, hit


@pytest.mark.parametrize("eye,target,fov", [((0.45, 0.1, 0.5), (0.0, 0.0, 0.12), np.pi / 2), ((-0.3, -0.35, 0.35), (0.05, 0.0, 0.15), 1.0), ((0.02, 0.01, 0.9), (0.0, 0.0, 0.0), 0.8)])
def test_raster_oracle_against_ray_casting(eye, target, fov):
    cm, hv = _scene()
    body = OracleWorld(cm, "f32").rigid_body_data()[0].astype(np.float32)
    cam = camera_desc("c", U.look_at(eye, target), 128, 128, fov, 0.01, 100.0)
    (color, posseg), = raster.render(build_visual_table(cm, 1), [cam], body[None])
    seg_r, pos_r = posseg[0, ..., 3].astype(int), posseg[0, ..., :3].astype(float) / 1000.0
    seg_c, pos_c, hit = _cast(cm, hv, body.astype(float), cam)
    ids = set(np.unique(seg_c).tolist())
    assert len(ids) >= 5 and ids == set(np.unique(seg_r).tolist())          # every object (and, seen from the side, the background) is in the picture
    same = seg_r == seg_c
    # disagreements only on silhouettes: each differing pixel has a 4-neighbour with another id in the ray-cast picture
    edge = np.zeros_like(same)
    edge[1:, :] |= seg_c[1:, :] != seg_c[:-1, :]
    edge[:-1, :] |= seg_c[1:, :] != seg_c[:-1, :]
    edge[:, 1:] |= seg_c[:, 1:] != seg_c[:, :-1]
    edge[:, :-1] |= seg_c[:, 1:] != seg_c[:, :-1]
    assert (~same).mean() < 0.01 and (edge | same).all(), ((~same).mean(), int((~same & ~edge).sum()))
    inner = same & hit & ~edge
    err = np.abs(pos_r - pos_c).max(-1)[inner]
    depth = -pos_c[..., 2][inner]
    # int16 millimetres of the texture, plus the 23-bit key of 1/depth over [1/far, 1/near]: one key step (1.2e-5 / m) is d^2 * 1.2e-5 metres of depth
    assert inner.sum() > 8000 and (err < 1.6e-3 + 2.5e-5 * depth ** 2).all(), float((err - 2.5e-5 * depth ** 2).max())
    assert err[depth < 2.0].max() < 1.7e-3
