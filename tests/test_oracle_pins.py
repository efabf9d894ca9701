"""Pins oracle/ppf_oracle.c against the reference's OWN host code: kernel.cu compiled for the host
(oracle/_ref/libppf_ref.so, built in place from /root/reference by oracle/Makefile).  Bit-exact."""
import ctypes

import numpy as np
import pytest

from conftest import have_ref

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def libs():
    from oracle import cpu, refgpu
    return refgpu.lib(), cpu.lib()


def test_constants(libs):
    R, O = libs
    assert np.float32(R.refhost_d_angle0()) == np.float32(O.oracle_d_angle0())
    assert np.float32(O.oracle_d_angle0()).view(np.uint32) == 0x3E567750


def test_hash_signed_bytes(libs):
    R, O = libs
    rng = np.random.default_rng(1)
    for n in (4, 12, 16):
        for _ in range(500):
            b = rng.integers(0, 256, n, dtype=np.uint8)
            assert R.refhost_hash(P(b), n) == O.oracle_hash(P(b), n)
    # differs from textbook (unsigned) FNV-1a whenever a byte >= 0x80
    b = np.array([0x80, 0, 0, 0], np.uint8)
    h = 2166136261
    for x in b:
        h = ((h ^ int(x)) * 16777619) & 0xFFFFFFFF
    assert O.oracle_hash(P(b), 4) != h


def test_feature_and_quantiser(libs):
    R, O = libs
    rng = np.random.default_rng(2)
    D = ctypes.c_float(R.refhost_d_angle0())
    for t in range(5000):
        p1, p2 = (rng.normal(size=3) * 50).astype(np.float32), (rng.normal(size=3) * 50).astype(np.float32)
        n1, n2 = rng.normal(size=3).astype(np.float32), rng.normal(size=3).astype(np.float32)
        if t % 5 == 0:
            n2 = n1.copy()                       # parallel normals: acos argument may exceed 1 -> NaN
        if t % 7 == 0:
            n1 = (p2 - p1).astype(np.float32)    # normal along d
        if t % 97 == 0:
            p2 = p1.copy()                       # coincident points: 0/0
        if t % 101 == 0:
            n1 = np.zeros(3, np.float32)         # zero normal
        a, b = np.zeros(4, np.float32), np.zeros(4, np.float32)
        R.refhost_compute_ppf(P(p1), P(n1), P(p2), P(n2), P(a))
        O.oracle_compute_ppf(P(p1), P(n1), P(p2), P(n2), P(b))
        assert (bits(a) == bits(b)).all()
        c, d = np.zeros(4, np.float32), np.zeros(4, np.float32)
        R.refhost_disc_feature(P(a), ctypes.c_float(3.7), D, P(c))
        O.oracle_disc_feature(P(b), ctypes.c_float(3.7), D, P(d))
        assert (bits(c) == bits(d)).all()


def test_quantiser_is_rounded_multiple(libs):
    """x - fmodf(x, d) == RN(k*d) with k = floor(x/d): the identity the CUDA quantiser uses."""
    _, O = libs
    rng = np.random.default_rng(3)
    d = np.float32(0.20943952)
    x = rng.uniform(0, np.pi, 20000).astype(np.float32)
    for xi in x:
        q = np.float32(O.oracle_quant_downf(ctypes.c_float(xi), ctypes.c_float(d)))
        k = int(np.floor(float(xi) / float(d)))
        assert q == np.float32(np.float32(k) * d)


def test_matrix_helpers(libs):
    R, O = libs
    rng = np.random.default_rng(4)
    for _ in range(300):
        th = np.float32(rng.uniform(-np.pi, np.pi))
        for ax in range(3):
            a, b = np.zeros(16, np.float32), np.zeros(16, np.float32)
            R.refhost_rot(ax, ctypes.c_float(th), P(a))
            O.oracle_rot(ax, ctypes.c_float(th), P(b))
            assert (bits(a) == bits(b)).all()
        A, B = rng.normal(size=16).astype(np.float32), rng.normal(size=16).astype(np.float32)
        c, d = np.zeros(16, np.float32), np.zeros(16, np.float32)
        R.refhost_mat4f_mul(P(A), P(B), P(c)); O.oracle_mat4f_mul(P(A), P(B), P(d))
        assert (bits(c) == bits(d)).all()
        v = rng.normal(size=4).astype(np.float32)
        e, f = np.zeros(4, np.float32), np.zeros(4, np.float32)
        R.refhost_mat4f_vmul(P(A), P(v), P(e)); O.oracle_mat4f_vmul(P(A), P(v), P(f))
        assert (bits(e) == bits(f)).all()
        # rigid transform for invht / quaternion
        T = np.zeros(16, np.float32)
        R.refhost_rot(2, ctypes.c_float(th), P(T))
        T2 = np.zeros(16, np.float32)
        R.refhost_rot(1, ctypes.c_float(th * 0.37), P(T2))
        M = np.zeros(16, np.float32)
        R.refhost_mat4f_mul(P(T), P(T2), P(M))
        M[3], M[7], M[11] = rng.normal(size=3).astype(np.float32) * 100
        g, h = np.zeros(16, np.float32), np.zeros(16, np.float32)
        R.refhost_invht(P(M), P(g)); O.oracle_invht(P(M), P(h))
        assert (bits(g) == bits(h)).all()
        q1, q2 = np.zeros(4, np.float32), np.zeros(4, np.float32)
        R.refhost_hrotmat2quat(P(M), P(q1)); O.oracle_hrotmat2quat(P(M), P(q2))
        assert (bits(q1) == bits(q2)).all()
