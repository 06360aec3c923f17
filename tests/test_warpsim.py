"""CPU suite: the warp-/CTA-cooperative truncating pseudo-inverse of the OSC kernels (abr_control_b200/csrc/abrb_coop.cuh)
executed on the CPU by tests/hostsim/warpsim.cpp — every CUDA thread an OS thread, shuffles and barriers emulated — so the
lane mapping (six- or eight-lane groups, spare lanes, the ballot walk of the in-line route, the CTA's queue flush with one
or two records per group) is covered without a GPU.  Reference: numpy.linalg.pinv(A A^T, rcond, hermitian=True), i.e. what
/root/reference/abr_control/controllers/osc.py:138-145 computes for the states below its determinant threshold."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
COMBOS = [(6, 6), (6, 3), (6, 5), (7, 6), (7, 3), (3, 3), (2, 2)]
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def ws():
    d = os.path.join(HERE, "hostsim")
    so, src = os.path.join(d, "_warpsim.so"), os.path.join(d, "warpsim.cpp")
    deps = [src] + [os.path.join(ROOT, "abr_control_b200", "csrc", f) for f in ("abrb_coop.cuh", "abrb_math.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(f) for f in deps):
        subprocess.run(["g++", "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-x", "c++", src,
                        "-o", so], check=True)
    lib = C.CDLL(so)
    for n, kd in COMBOS:
        getattr(lib, f"ws_inline_{n}_{kd}").argtypes = [C.c_uint, dp, dp, dp, C.c_double, C.c_int, dp, dp]
        getattr(lib, f"ws_flush_{n}_{kd}").argtypes = [C.c_int, C.c_int, dp, C.POINTER(C.c_longlong), dp, dp, C.c_double, C.c_int]
        getattr(lib, f"ws_layout_{n}_{kd}").restype = C.c_int
    lib.ws_push_6_6.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_int, dp, dp, dp, dp, C.c_double, C.c_int, dp, dp, dp,
                                C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
    return lib


def _p(a):
    return a.ctypes.data_as(dp)


def _states(rng, count, n, kd):
    """KD x N matrices: a third well conditioned, the rest with one (or two) singular values far below rcond * largest"""
    A = rng.normal(size=(count, kd, n))
    for i in range(count):
        if i % 3 and kd >= 2:
            A[i, -1] = A[i, :-1].T @ rng.normal(size=kd - 1) + 1e-7 * rng.normal(size=n)
        if i % 3 == 2 and kd >= 3:
            A[i, -2] = A[i, :-2].T @ rng.normal(size=kd - 2) + 1e-8 * rng.normal(size=n)
    return A


def _w(A, y, rcond):
    return A.T @ (np.linalg.pinv(A @ A.T, rcond=rcond, hermitian=True) @ y)


def test_group_layout(ws):
    for n, kd in COMBOS:
        lanes, per_warp = getattr(ws, f"ws_layout_{n}_{kd}")(0), getattr(ws, f"ws_layout_{n}_{kd}")(1)
        assert (lanes, per_warp) == ((6, 5) if max(n, kd) <= 6 else (8, 4))
        assert getattr(ws, f"ws_layout_{n}_{kd}")(2) == kd * n + n * (n + 1) // 2 + 2 * kd


@pytest.mark.parametrize("n,kd", COMBOS)
def test_inline_route_over_ballot_masks(ws, n, kd):
    rng = np.random.default_rng(100 * n + kd)
    rcond = 1e-4
    fn = getattr(ws, f"ws_inline_{n}_{kd}")
    masks = [1 << 31, 1 << 30 | 1, 0b10110100101, 0x80000421 | 3 << 14, 0xFFFFFFFF]  # 1, 2, 6, 6, 32 waiting lanes
    for mask in masks if (n, kd) in ((6, 6), (7, 6), (6, 3)) else masks[:3]:
        for two in (1, 0):
            A = np.ascontiguousarray(_states(rng, 32, n, kd))
            y, z = rng.normal(size=(32, kd)), rng.normal(size=(32, kd))
            wy, wz = np.zeros((32, n)), np.zeros((32, n))
            fn(mask, _p(A), _p(y), _p(z), rcond, two, _p(wy), _p(wz))
            for lane in range(32):
                if (mask >> lane) & 1:
                    ey = _w(A[lane], y[lane], rcond)
                    assert np.abs(wy[lane] - ey).max() <= 1e-9 * max(1.0, np.abs(ey).max()), (mask, lane)
                    if two:
                        ez = _w(A[lane], z[lane], rcond)
                        assert np.abs(wz[lane] - ez).max() <= 1e-9 * max(1.0, np.abs(ez).max()), (mask, lane)
                else:  # lanes that did not wait keep what they had in the exchange area
                    assert np.array_equal(wy[lane][:kd], y[lane]) and np.array_equal(wz[lane][:kd], z[lane])


@pytest.mark.parametrize("n,kd", [(6, 6), (6, 3), (7, 6), (3, 3)])
def test_cta_queue_flush(ws, n, kd):
    """du = -L A^T (Mx y + Mx z) added to the rows already written; 128- and 64-thread CTAs; up to one record per group
    (one pass) and beyond (two records per group interleaved); the training signal gets the y part only"""
    rng = np.random.default_rng(7 * n + kd)
    rcond = 1e-4
    rec_len = getattr(ws, f"ws_layout_{n}_{kd}")(2)
    per_warp = getattr(ws, f"ws_layout_{n}_{kd}")(1)
    fn = getattr(ws, f"ws_flush_{n}_{kd}")
    B = 300
    for threads, counts in ((128, (1, per_warp * 4, per_warp * 4 + 1, 32)), (64, (3, per_warp * 2 + 2))):
        for cnt in counts:
            for two in (1, 0):
                A = _states(rng, cnt, n, kd)
                y, z = rng.normal(size=(cnt, kd)), rng.normal(size=(cnt, kd))
                L = np.tril(rng.normal(size=(cnt, n, n))) + 2 * np.eye(n)
                rows = rng.choice(B, size=cnt, replace=False).astype(np.int64)
                rec = np.zeros((cnt, rec_len))
                tri = np.tril_indices(n)
                for i in range(cnt):
                    rec[i, : kd * n] = A[i].ravel()
                    rec[i, kd * n: kd * n + n * (n + 1) // 2] = L[i][tri]
                    rec[i, kd * n + n * (n + 1) // 2: kd * n + n * (n + 1) // 2 + kd] = y[i]
                    rec[i, kd * n + n * (n + 1) // 2 + kd:] = z[i]
                u0, t0 = rng.normal(size=(B, n)), rng.normal(size=(B, n))
                u, tr = u0.copy(), t0.copy()
                fn(threads, cnt, _p(rec), rows.ctypes.data_as(C.POINTER(C.c_longlong)), _p(u), _p(tr), rcond, two)
                eu, et = u0.copy(), t0.copy()
                for i in range(cnt):
                    wy = _w(A[i], y[i], rcond)
                    wz = _w(A[i], z[i], rcond) if two else 0.0
                    eu[rows[i]] -= L[i] @ (wy + wz)
                    et[rows[i]] -= L[i] @ wy
                scale = max(1.0, np.abs(eu).max())
                assert np.abs(u - eu).max() <= 1e-9 * scale, (threads, cnt, two)
                assert np.abs(tr - et).max() <= 1e-9 * scale, (threads, cnt, two)


@pytest.mark.parametrize("shared", [1, 0])
def test_deferral_queue_and_overflow_into_the_inline_route(ws, shared):
    """WarpCoop::pinv as osc_eval calls it (abrb_osc.cuh): waiting lanes leave a record (A, the factor of M, y, z, row) in
    the CTA queue and continue with zeros; lanes that find the queue full are decomposed in line by their warp; padding
    lanes of a ragged warp never wait.  Scratch in shared memory (fp64 kernels) and in registers (fp32 kernels)."""
    rng = np.random.default_rng(5 + shared)
    n = kd = 6
    rcond, rec_len = 1e-4, ws.ws_layout_6_6(2)
    A = np.ascontiguousarray(_states(rng, 32, n, kd))
    L = np.ascontiguousarray(np.tril(rng.normal(size=(32, n, n))))
    y, z = rng.normal(size=(32, kd)), rng.normal(size=(32, kd))
    tri = np.tril_indices(n)
    for slow, valid, qcap in ((0x00F0F00F, 0xFFFFFFFF, 32), (0xFFFFFFFF, 0xFFFFFFFF, 9), (0xFFFF0000, 0x00FFFFFF, 3),
                              (0x0000FFFF, 0xFFFFFFFF, 0), (0, 0xFFFFFFFF, 8)):
        wy, wz = np.zeros((32, n)), np.zeros((32, n))
        qrec, qrow, qcount = np.zeros((max(qcap, 1), rec_len)), np.zeros(max(qcap, 1), dtype=np.int64), C.c_int(0)
        ws.ws_push_6_6(shared, slow, valid, qcap, _p(A), _p(L), _p(y), _p(z), rcond, 1, _p(wy), _p(wz), _p(qrec),
                       qrow.ctypes.data_as(C.POINTER(C.c_longlong)), C.byref(qcount))
        waiting = [l for l in range(32) if (slow >> l) & (valid >> l) & 1]
        assert qcount.value == len(waiting)  # every waiting lane asked for a place
        queued = {int(qrow[i]) - 1000: i for i in range(min(qcap, len(waiting)))}
        assert len(queued) == min(qcap, len(waiting)) and set(queued) <= set(waiting)
        for lane in range(32):
            if lane in queued:
                r = qrec[queued[lane]]
                assert np.array_equal(r[:36], A[lane].ravel()) and np.array_equal(r[36:57], L[lane][tri])
                assert np.array_equal(r[57:63], y[lane]) and np.array_equal(r[63:69], z[lane])
                assert not wy[lane].any() and not wz[lane].any()
            elif lane in waiting:  # no room in the queue: finished in line
                ey, ez = _w(A[lane], y[lane], rcond), _w(A[lane], z[lane], rcond)
                assert np.abs(wy[lane] - ey).max() <= 1e-9 * max(1.0, np.abs(ey).max())
                assert np.abs(wz[lane] - ez).max() <= 1e-9 * max(1.0, np.abs(ez).max())
            else:
                assert (wy[lane] == -7.0).all() and (wz[lane] == -7.0).all()
