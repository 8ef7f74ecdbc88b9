"""transforms3d.euler: euler2mat / mat2euler / euler2quat / quat2euler with the axes strings of the original ('sxyz' default)."""
import math

import numpy as np

from .quaternions import mat2quat, quat2mat

_NEXT_AXIS = [1, 2, 0, 1]
_AXES2TUPLE = {
    "sxyz": (0, 0, 0, 0), "sxyx": (0, 0, 1, 0), "sxzy": (0, 1, 0, 0), "sxzx": (0, 1, 1, 0), "syzx": (1, 0, 0, 0), "syzy": (1, 0, 1, 0),
    "syxz": (1, 1, 0, 0), "syxy": (1, 1, 1, 0), "szxy": (2, 0, 0, 0), "szxz": (2, 0, 1, 0), "szyx": (2, 1, 0, 0), "szyz": (2, 1, 1, 0),
    "rzyx": (0, 0, 0, 1), "rxyx": (0, 0, 1, 1), "ryzx": (0, 1, 0, 1), "rxzx": (0, 1, 1, 1), "rxzy": (1, 0, 0, 1), "ryzy": (1, 0, 1, 1),
    "rzxy": (1, 1, 0, 1), "ryxy": (1, 1, 1, 1), "ryxz": (2, 0, 0, 1), "rzxz": (2, 0, 1, 1), "rxyz": (2, 1, 0, 1), "rzyz": (2, 1, 1, 1)}
_EPS4 = np.finfo(float).eps * 4.0


def _axes(axes):
    return _AXES2TUPLE[axes.lower()] if isinstance(axes, str) else axes


def euler2mat(ai, aj, ak, axes="sxyz"):
    firstaxis, parity, repetition, frame = _axes(axes)
    i = firstaxis
    j = _NEXT_AXIS[i + parity]
    k = _NEXT_AXIS[i - parity + 1]
    if frame:
        ai, ak = ak, ai
    if parity:
        ai, aj, ak = -ai, -aj, -ak
    si, sj, sk = math.sin(ai), math.sin(aj), math.sin(ak)
    ci, cj, ck = math.cos(ai), math.cos(aj), math.cos(ak)
    cc, cs = ci * ck, ci * sk
    sc, ss = si * ck, si * sk
    M = np.eye(3)
    if repetition:
        M[i, i] = cj; M[i, j] = sj * si; M[i, k] = sj * ci
        M[j, i] = sj * sk; M[j, j] = -cj * ss + cc; M[j, k] = -cj * cs - sc
        M[k, i] = -sj * ck; M[k, j] = cj * sc + cs; M[k, k] = cj * cc - ss
    else:
        M[i, i] = cj * ck; M[i, j] = sj * sc - cs; M[i, k] = sj * cc + ss
        M[j, i] = cj * sk; M[j, j] = sj * ss + cc; M[j, k] = sj * cs - sc
        M[k, i] = -sj; M[k, j] = cj * si; M[k, k] = cj * ci
    return M


def mat2euler(mat, axes="sxyz"):
    firstaxis, parity, repetition, frame = _axes(axes)
    i = firstaxis
    j = _NEXT_AXIS[i + parity]
    k = _NEXT_AXIS[i - parity + 1]
    M = np.array(mat, dtype=np.float64, copy=False)[:3, :3]
    if repetition:
        sy = math.sqrt(M[i, j] * M[i, j] + M[i, k] * M[i, k])
        if sy > _EPS4:
            ax, ay, az = math.atan2(M[i, j], M[i, k]), math.atan2(sy, M[i, i]), math.atan2(M[j, i], -M[k, i])
        else:
            ax, ay, az = math.atan2(-M[j, k], M[j, j]), math.atan2(sy, M[i, i]), 0.0
    else:
        cy = math.sqrt(M[i, i] * M[i, i] + M[j, i] * M[j, i])
        if cy > _EPS4:
            ax, ay, az = math.atan2(M[k, j], M[k, k]), math.atan2(-M[k, i], cy), math.atan2(M[j, i], M[i, i])
        else:
            ax, ay, az = math.atan2(-M[j, k], M[j, j]), math.atan2(-M[k, i], cy), 0.0
    if parity:
        ax, ay, az = -ax, -ay, -az
    if frame:
        ax, az = az, ax
    return ax, ay, az


def euler2quat(ai, aj, ak, axes="sxyz"):
    firstaxis, parity, repetition, frame = _axes(axes)
    i = firstaxis + 1
    j = _NEXT_AXIS[i + parity - 1] + 1
    k = _NEXT_AXIS[i - parity] + 1
    if frame:
        ai, ak = ak, ai
    if parity:
        aj = -aj
    ai, aj, ak = ai / 2.0, aj / 2.0, ak / 2.0
    ci, si = math.cos(ai), math.sin(ai)
    cj, sj = math.cos(aj), math.sin(aj)
    ck, sk = math.cos(ak), math.sin(ak)
    cc, cs = ci * ck, ci * sk
    sc, ss = si * ck, si * sk
    q = np.empty((4,))
    if repetition:
        q[0] = cj * (cc - ss); q[i] = cj * (cs + sc); q[j] = sj * (cc + ss); q[k] = sj * (cs - sc)
    else:
        q[0] = cj * cc + sj * ss; q[i] = cj * sc - sj * cs; q[j] = cj * ss + sj * cc; q[k] = cj * cs - sj * sc
    if parity:
        q[j] *= -1.0
    return q


def quat2euler(quaternion, axes="sxyz"):
    return mat2euler(quat2mat(quaternion), axes)
