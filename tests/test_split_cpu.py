"""Ordered subtree splitting of closest-hit rays (csrc/mr_split.cuh) on the CPU: the task_step / task_replay code the CUDA
tracer runs, driven by a randomised scheduler in the host-check flavour (mirres_test_closest_split), must reproduce the
oracle's bvh_hit_with_normal (helperDi.slang:313-395) bit for bit -- hit flag, t, position, normal and primitive id --
whatever the steal schedule, including negative-t self hits and log overflows (capacity 1 / 2 force restarts)."""
import ctypes

import numpy as np
import pytest

import hostcheck as H
import parity as P


def _run(oracle, sc, w, org, dirs, lanes, cap, seed):
    lib = ctypes.CDLL(H._build.build_hostcheck())
    n = len(org)
    org, dirs = np.ascontiguousarray(org, np.float32), np.ascontiguousarray(dirs, np.float32)
    hit, prim = np.zeros(n, np.int32), np.zeros(n, np.int32)
    t, pos, nrm = np.zeros(n, np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    stats = np.zeros(4, np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.mirres_test_closest_split(ctypes.c_void_p(w.packed[0].data_ptr()), ctypes.c_void_p(w.packed[1].data_ptr()),
                                       p(org), p(dirs), n, lanes, cap, seed, p(hit), p(t), p(pos), p(nrm), p(prim), p(stats))
    assert rc == 0
    oh, ot, op, on, opr = oracle.trace(sc["bvh"], org, dirs)
    assert (hit == oh).all() and (prim == opr).all()
    assert (t == ot).all() and (pos == op).all() and (nrm[oh > 0] == on[oh > 0]).all()
    return stats, ot[oh > 0]


@pytest.mark.parametrize("name", ["T0", "T2", "C1"])
def test_split_walk_equals_reference_walk(oracle, name):
    H.activate()
    sc = P.scene(name)
    w = H.OracleBvhWorker(H.t(sc["vert"]), H.t(sc["tri"]))
    w.update_mesh(H.t(sc["vert"]), H.t(sc["tri"]))
    rng = np.random.default_rng(7)
    hitm = sc["hit"] > 0
    d = rng.standard_normal((int(hitm.sum()), 3)).astype(np.float32)
    o = (sc["pos"][hitm] + 0.01 * d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    steals = restarts = negative = 0
    for lanes, cap in ((32, 8), (8, 2), (32, 1)):
        for org, dirs in ((sc["rays_o"], sc["rays_d"]), (o, d)):
            st, ts = _run(oracle, sc, w, org, dirs, lanes, cap, 11 + lanes)
            steals += int(st[0])
            restarts += int(st[1]) if cap < 8 else 0
            negative += int((ts < 0).sum())
    assert steals > 100 and restarts > 0 and negative > 0  # the schedule split rays, overflowed logs, met negative-t hits
