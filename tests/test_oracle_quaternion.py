"""The reference's quaternion / ray known-answer tests, applied to BOTH the oracle
(oracle/ref_numpy.py) and the host helpers of the product's rtgs.utils.quaternion.
Vectors and tolerances are those of /root/reference/tests/test_quaternion.py and test_ray.py."""
import numpy as np
import pytest

from oracle import ref_numpy as O
from rtgs.utils import quaternion as Q
from rtgs.utils.types import vec3, vec4

IMPLS = [
    pytest.param(dict(mul=O.quat_mul, conj=O.quat_conj, rot=O.rot_vec3, mat3=O.as_rotation_mat3), id="oracle"),
    pytest.param(dict(mul=Q.mul, conj=Q.conj, rot=Q.rot_vec3, mat3=Q.as_rotation_mat3), id="rtgs.utils"),
]


@pytest.mark.parametrize("f", IMPLS)
def test_mul(f):
    # tests/test_quaternion.py:16-29 — scalar-last (x,y,z,w)
    r = f["mul"]([1, -2, 1, 3], [-1, 2, 3, 2])
    assert np.allclose(np.asarray(r, dtype=np.float64), [-9, -2, 11, 8], atol=0)


@pytest.mark.parametrize("f", IMPLS)
def test_conj(f):
    # tests/test_quaternion.py:32-43
    assert np.array_equal(np.asarray(f["conj"]([1, -2, 1, 3]), dtype=np.float64), [-1, 2, -1, 3])


@pytest.mark.parametrize("f", IMPLS)
def test_rot_vec3(f):
    # tests/test_quaternion.py:93-118 — (1,0,0) about z by pi/2 -> (0,1,0); about y by pi/2 -> (0,0,-1)
    qz = np.array([0, 0, np.sin(np.pi / 4), np.cos(np.pi / 4)])
    qy = np.array([0, np.sin(np.pi / 4), 0, np.cos(np.pi / 4)])
    assert np.allclose(np.asarray(f["rot"](qz, [1, 0, 0]), dtype=np.float64), [0, 1, 0], atol=1e-6)
    assert np.allclose(np.asarray(f["rot"](qy, [1, 0, 0]), dtype=np.float64), [0, 0, -1], atol=1e-6)


@pytest.mark.parametrize("f", IMPLS)
def test_as_rotation_mat3(f):
    # tests/test_quaternion.py:121-153 — M v == q v q*
    rng = np.random.default_rng(42)
    for _ in range(8):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        v = rng.normal(size=3)
        M = np.asarray(f["mat3"](q), dtype=np.float64)
        assert np.allclose(M @ v, np.asarray(f["rot"](q, v), dtype=np.float64), atol=1e-6)
        assert np.allclose(M @ M.T, np.eye(3), atol=1e-6)


def test_axis_angle_roundtrip():
    # tests/test_quaternion.py:65-90 (host helpers only; off the render path)
    v = np.array([0.3, -0.2, 0.5])
    q = Q.from_axis_angle(v)
    assert np.allclose(np.asarray(Q.as_axis_angle(q)), v, atol=1e-6)
    assert np.allclose(np.linalg.norm(np.asarray(q)), 1, atol=1e-6)


def test_inv_follows_reference_definition():
    # utils/quaternion.py:38-47 divides by |q| (not |q|^2); the reference's own test_inv is stale.
    q = np.array([1.0, -2.0, 1.0, 3.0])
    assert np.allclose(np.asarray(Q.inv(q)), O.quat_conj(q) / np.linalg.norm(q), atol=1e-6)
    u = q / np.linalg.norm(q)
    assert np.allclose(np.asarray(Q.mul(u, Q.inv(u))), [0, 0, 0, 1], atol=1e-6)


def test_mat4():
    q = Q.from_axis_angle([0, 0, np.pi / 2])
    m = Q.as_rotation_mat4(q)
    assert m.shape == (4, 4) and np.allclose(m[:3, :3] @ [1, 0, 0], [0, 1, 0], atol=1e-6) and m[3, 3] == 1


def test_from_rotation_matrix_roundtrip():
    rng = np.random.default_rng(0)
    for _ in range(16):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        if q[3] < 0:
            q = -q
        R = O.as_rotation_mat3(q)
        assert np.allclose(np.asarray(Q.from_rotation_matrix(R), dtype=np.float64), q, atol=1e-6)
        assert np.allclose(O._quat_from_rotation_matrix(R), q, atol=1e-12)
