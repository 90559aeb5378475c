"""BezierTraj wire format (traj_utils/msg/BezierTraj.msg, ROS1 serialisation) — host-only, runs without a GPU."""
import struct

import numpy as np
import pytest

import pred_occ_planner_b200 as p
from pred_occ_planner_b200 import wire, worldgen


def _rec(M=3):
    r = worldgen.straight_trajs(np.array([[1.0, -2.0, 0.5]]), np.array([[4.0, 2.0, 1.5]]), 12.25, 1.5, M, 0.3)[0].copy()
    r["drone_id"], r["traj_id"] = 7, 42
    return r


def test_known_bytes():
    r = _rec(2)
    b = wire.to_wire(r, 12.5)
    # int32 id, int32 traj id, time start (sec, nsec), time pub, uint8 order, float32[] duration, Point[] cpts
    assert len(b) == 4 + 4 + 8 + 8 + 1 + 4 + 4 * 2 + 4 + 24 * 10
    assert struct.unpack_from("<iiIIII", b, 0) == (7, 42, 12, 250000000, 12, 500000000)
    assert b[24] == 4
    assert struct.unpack_from("<Iff", b, 25) == (2, np.float32(0.3), np.float32(0.3))
    assert struct.unpack_from("<I", b, 37)[0] == 10
    assert struct.unpack_from("<3d", b, 41) == tuple(r["cpts"][0])
    assert struct.unpack_from("<3d", b, 41 + 24 * 9) == tuple(r["cpts"][9])


@pytest.mark.parametrize("M", [0, 1, 5, p.POP_MAX_PIECES])
def test_round_trip(M):
    r = _rec(max(M, 1))
    if M == 0:
        r["n_pieces"] = 0
    b = wire.to_wire(r, 99.000000001)
    back, pub, n = wire.from_wire(b + b"trailing")
    assert n == len(b) and abs(pub - 99.000000001) < 1e-9
    assert back["drone_id"] == 7 and back["traj_id"] == 42 and back["n_pieces"] == M and back["valid"] == (1 if M else 0)
    assert back["start_time"] == 12.25
    assert np.array_equal(back["duration"][:M], r["duration"][:M]) and np.array_equal(back["cpts"][:5 * M], r["cpts"][:5 * M])
    assert not back["cpts"][5 * M:].any()


def test_malformed_rejected():
    b = wire.to_wire(_rec(3), 1.0)
    with pytest.raises(p.PopError):
        wire.from_wire(b[:-8])                          # truncated control points
    bad = bytearray(b)
    bad[24] = 3                                         # order 3 curves are not what the planner exchanges
    with pytest.raises(p.PopError):
        wire.from_wire(bytes(bad))
    bad = bytearray(b)
    struct.pack_into("<I", bad, 25, 100)                # more pieces than trajectory_piece_max_size
    with pytest.raises(p.PopError):
        wire.from_wire(bytes(bad))
