"""Pins the oracle's Bernstein/Bezier restatement to the reference's own known-answer tests:
traj_utils/test/test_bernstein.cpp:9-19,58-108 (fixture: cpts = (0..4)*1 on t in [2,4]; a
3-piece curve with durations {1,2,3})."""
import numpy as np


def _bp_cpts():
    return np.array([[i, i, i] for i in range(5)], dtype=np.float64)


def test_piece_known_answers(oracle):
    c = _bp_cpts()
    # TestCase1 (:58-66): exact equality
    assert np.array_equal(oracle.piece_eval(c, 2.0, 4.0, 2.0, 0), np.zeros(3))
    assert np.array_equal(oracle.piece_eval(c, 2.0, 4.0, 2.0, 1), np.array([2.0, 2.0, 2.0]))
    assert np.array_equal(oracle.piece_eval(c, 2.0, 4.0, 2.0, 2), np.zeros(3))
    assert np.array_equal(oracle.piece_eval(c, 2.0, 4.0, 4.0, 0), np.array([4.0, 4.0, 4.0]))


def test_coefficient_matrix(oracle):
    # TestControlPoints (:68-72)
    A = np.array([[1, -4, 6, -4, 1], [0, 4, -12, 12, -4], [0, 0, 6, -12, 6], [0, 0, 0, 4, -4], [0, 0, 0, 0, 1]],
                 dtype=np.float64)
    assert np.array_equal(oracle.bernstein_coeff(), A)


def test_derivative_control_points(oracle):
    # TestControlPoints (:74-81): velocity control points all 4, acceleration control points 0.
    # max rates over one piece of duration 2: |(4,4,4)|/2 and 0
    c = _bp_cpts()
    assert np.isclose(oracle.bezier_max_rate([2.0], c, 1), np.linalg.norm([4, 4, 4]) / 2.0, rtol=0, atol=1e-15)
    assert oracle.bezier_max_rate([2.0], c, 2) == 0.0


def _three_piece():
    cp = np.array([[0, 0, 0], [0.5, 0.5, 0.5], [1, 1, 1], [2, 2, 2], [4, 4, 4],
                   [4, 4, 4], [4, 5, 4], [5, 4, 4], [5, 6, 4], [6, 5, 4],
                   [6, 5, 4], [7, 5, 4], [7, 6, 4], [8, 5, 4], [8, 6, 4]], dtype=np.float64)
    return [1.0, 2.0, 3.0], cp


def test_three_piece_curve(oracle):
    # BezierTest (:84-108): duration 6, 3 pieces, max velocity rate > 1.5
    dur, cp = _three_piece()
    assert sum(dur) == 6
    assert oracle.bezier_max_rate(dur, cp, 1) > 1.5
    # end points of each piece are interpolated
    assert np.allclose(oracle.bezier_eval(dur, cp, [0.0])[0], cp[0])
    assert np.allclose(oracle.bezier_eval(dur, cp, [1.0])[0], cp[5])
    assert np.allclose(oracle.bezier_eval(dur, cp, [3.0])[0], cp[10])
    assert np.allclose(oracle.bezier_eval(dur, cp, [6.0])[0], cp[14])
    # velocity is the derivative of position (finite difference)
    t = np.array([0.3, 1.7, 4.2])
    h = 1e-6
    fd = (oracle.bezier_eval(dur, cp, t + h) - oracle.bezier_eval(dur, cp, t - h)) / (2 * h)
    assert np.allclose(oracle.bezier_eval(dur, cp, t, 1), fd, atol=1e-6)
    fd2 = (oracle.bezier_eval(dur, cp, t + h, 1) - oracle.bezier_eval(dur, cp, t - h, 1)) / (2 * h)
    assert np.allclose(oracle.bezier_eval(dur, cp, t, 2), fd2, atol=1e-5)
