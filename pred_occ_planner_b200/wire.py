"""`traj_utils/BezierTraj` on the ROS1 wire (traj_utils/msg/BezierTraj.msg) <-> `pop_traj_record`, through the
C ABI (pop_traj_to_wire / pop_traj_from_wire): what a ROS bridge on either side of libpop needs."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import TRAJ_DTYPE, PopTrajRecord


def to_wire(rec: np.ndarray, pub_time: float) -> bytes:
    rec = np.ascontiguousarray(rec, TRAJ_DTYPE).reshape(1)
    lib = _lib.load()
    ptr = rec.ctypes.data_as(C.POINTER(PopTrajRecord))
    n = lib.pop_traj_to_wire(ptr, float(pub_time), None, 0)
    if n < 0:
        raise _lib.PopError(n, "pop_traj_to_wire")
    buf = (C.c_uint8 * n)()
    m = lib.pop_traj_to_wire(ptr, float(pub_time), buf, n)
    if m != n:
        raise _lib.PopError(m, "pop_traj_to_wire")
    return bytes(buf)


def from_wire(data: bytes):
    """Returns (record, pub_time, bytes consumed)."""
    rec = np.zeros(1, TRAJ_DTYPE)
    pub = C.c_double(0.0)
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    n = _lib.load().pop_traj_from_wire(buf, len(data), rec.ctypes.data_as(C.POINTER(PopTrajRecord)), C.byref(pub))
    if n < 0:
        raise _lib.PopError(n, "pop_traj_from_wire")
    return rec[0], pub.value, n
