"""TEST INFRASTRUCTURE ONLY — numpy front end of oracle/libngp_oracle.so (ngp_oracle.c).

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs only.  Function names and argument order mirror the reference's `vren` functions
(models/csrc/binding.cpp:323-342) so parity tests read like calls to the reference.
Every function takes/returns numpy arrays (float32 / int32 / int64 / uint8, C-contiguous).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libngp_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "ngp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libngp_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.ref_raymarching_train.restype = ctypes.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


_i64, _f32 = ctypes.c_int64, ctypes.c_float


def ray_aabb_intersect(rays_o, rays_d, centers, half_sizes, max_hits):
    rays_o, rays_d, centers, half_sizes = _f(rays_o), _f(rays_d), _f(centers), _f(half_sizes)
    R, V = rays_o.shape[0], centers.shape[0]
    cnt = np.empty(R, np.int32); t = np.empty((R, max_hits, 2), np.float32); idx = np.empty((R, max_hits), np.int64)
    lib().ref_ray_aabb_intersect(_p(rays_o), _p(rays_d), _p(centers), _p(half_sizes), _i64(R), _i64(V),
                                 ctypes.c_int(max_hits), _p(cnt), _p(t), _p(idx))
    return cnt, t, idx


def morton3D(coords):
    coords = np.ascontiguousarray(coords, np.int32)
    out = np.empty(coords.shape[0], np.int32)
    lib().ref_morton3D(_p(coords), _i64(coords.shape[0]), _p(out))
    return out


def morton3D_invert(indices):
    indices = np.ascontiguousarray(indices, np.int32)
    out = np.empty((indices.shape[0], 3), np.int32)
    lib().ref_morton3D_invert(_p(indices), _i64(indices.shape[0]), _p(out))
    return out


def packbits(density_grid, thr, n_bytes=None):
    g = _f(density_grid).reshape(-1)
    n = g.size // 8 if n_bytes is None else n_bytes
    out = np.empty(n, np.uint8)
    lib().ref_packbits(_p(g), _i64(n), _f32(thr), _p(out))
    return out


def raymarching_train(rays_o, rays_d, hits_t, bitfield, cascades, scale, esf, noise, grid_size, max_samples):
    rays_o, rays_d, hits_t, noise = _f(rays_o), _f(rays_d), _f(hits_t), _f(noise)
    bitfield = np.ascontiguousarray(bitfield, np.uint8)
    R = rays_o.shape[0]
    rays_a = np.empty((R, 3), np.int64)
    args = (_p(rays_o), _p(rays_d), _p(hits_t), _p(bitfield), ctypes.c_int(cascades), _f32(scale), _f32(esf), _p(noise),
            ctypes.c_int(grid_size), ctypes.c_int(max_samples), _i64(R))
    total = lib().ref_raymarching_train(*args, _p(rays_a), None, None, None, None)
    xyzs = np.empty((total, 3), np.float32); dirs = np.empty((total, 3), np.float32)
    deltas = np.empty(total, np.float32); ts = np.empty(total, np.float32)
    lib().ref_raymarching_train(*args, _p(rays_a), _p(xyzs), _p(dirs), _p(deltas), _p(ts))
    return rays_a, xyzs, dirs, deltas, ts, np.array([total, R], np.int32)


def raymarching_test(rays_o, rays_d, hits_t, alive, bitfield, cascades, scale, esf, grid_size, max_samples, N_samples):
    """hits_t (R,2) is advanced IN PLACE (pass a copy to keep the original)."""
    rays_o, rays_d = _f(rays_o), _f(rays_d)
    assert hits_t.dtype == np.float32 and hits_t.flags.c_contiguous
    alive = np.ascontiguousarray(alive, np.int64); bitfield = np.ascontiguousarray(bitfield, np.uint8)
    A = alive.shape[0]
    xyzs = np.empty((A, N_samples, 3), np.float32); dirs = np.empty((A, N_samples, 3), np.float32)
    deltas = np.empty((A, N_samples), np.float32); ts = np.empty((A, N_samples), np.float32)
    n_eff = np.empty(A, np.int32)
    lib().ref_raymarching_test(_p(rays_o), _p(rays_d), _p(hits_t), _p(alive), _p(bitfield), ctypes.c_int(cascades),
                               _f32(scale), _f32(esf), ctypes.c_int(grid_size), ctypes.c_int(max_samples),
                               ctypes.c_int(N_samples), _i64(A), _p(xyzs), _p(dirs), _p(deltas), _p(ts), _p(n_eff))
    return xyzs, dirs, deltas, ts, n_eff


def composite_alpha_fw(sigmas, deltas, rays_a, T_thr):
    sigmas, deltas = _f(sigmas), _f(deltas); rays_a = np.ascontiguousarray(rays_a, np.int64)
    S, R = sigmas.shape[0], rays_a.shape[0]
    alphas = np.empty(S, np.float32); ws = np.empty(S, np.float32)
    lib().ref_composite_alpha_fw(_p(sigmas), _p(deltas), _p(rays_a), _f32(T_thr), _i64(S), _i64(R), _p(alphas), _p(ws))
    return alphas, ws


def composite_train_fw(sigmas, rgbs, normals_pred, sems, deltas, ts, rays_a, T_thr, classes):
    sigmas, rgbs, normals_pred, sems, deltas, ts = map(_f, (sigmas, rgbs, normals_pred, sems, deltas, ts))
    rays_a = np.ascontiguousarray(rays_a, np.int64)
    S, R = sigmas.shape[0], rays_a.shape[0]
    total = np.empty(R, np.int64); opacity = np.empty(R, np.float32); depth = np.empty(R, np.float32)
    rgb = np.empty((R, 3), np.float32); normal = np.empty((R, 3), np.float32); sem = np.empty((R, classes), np.float32)
    ws = np.empty(S, np.float32)
    lib().ref_composite_train_fw(_p(sigmas), _p(rgbs), _p(normals_pred), _p(sems), _p(deltas), _p(ts), _p(rays_a),
                                 _f32(T_thr), ctypes.c_int(classes), _i64(S), _i64(R), _p(total), _p(opacity), _p(depth),
                                 _p(rgb), _p(normal), _p(sem), _p(ws))
    return total, opacity, depth, rgb, normal, sem, ws


def composite_train_bw(dL_dopacity, dL_ddepth, dL_drgb, dL_dnormal_pred, dL_dsem, dL_dws, sigmas, rgbs, normals_pred,
                       ws, deltas, ts, rays_a, opacity, depth, rgb, normal_pred, T_thr, classes):
    (dL_dopacity, dL_ddepth, dL_drgb, dL_dnormal_pred, dL_dsem, dL_dws, sigmas, rgbs, ws, deltas, ts, opacity, depth,
     rgb) = map(_f, (dL_dopacity, dL_ddepth, dL_drgb, dL_dnormal_pred, dL_dsem, dL_dws, sigmas, rgbs, ws, deltas, ts,
                     opacity, depth, rgb))
    rays_a = np.ascontiguousarray(rays_a, np.int64)
    S, R = sigmas.shape[0], rays_a.shape[0]
    d_sig = np.empty(S, np.float32); d_rgb = np.empty((S, 3), np.float32); d_nrm = np.empty((S, 3), np.float32)
    d_sem = np.empty((S, classes), np.float32)
    lib().ref_composite_train_bw(_p(dL_dopacity), _p(dL_ddepth), _p(dL_drgb), _p(dL_dnormal_pred), _p(dL_dsem),
                                 _p(dL_dws), _p(sigmas), _p(rgbs), _p(ws), _p(deltas), _p(ts), _p(rays_a), _p(opacity),
                                 _p(depth), _p(rgb), _f32(T_thr), ctypes.c_int(classes), _i64(S), _i64(R), _p(d_sig),
                                 _p(d_rgb), _p(d_nrm), _p(d_sem))
    return d_sig, d_rgb, d_nrm, d_sem


def composite_test_fw(sigmas, rgbs, normals, normals_raw, sems, deltas, ts, hits_t, alive, T_thr, classes, n_eff,
                      opacity, depth, rgb, normal, normal_raw, sem):
    """In place on alive / opacity / depth / rgb / normal / normal_raw / sem (float32 C-contiguous arrays)."""
    sigmas, rgbs, normals, normals_raw, sems, deltas, ts = map(_f, (sigmas, rgbs, normals, normals_raw, sems, deltas, ts))
    A, N = sigmas.shape
    lib().ref_composite_test_fw(_p(sigmas), _p(rgbs), _p(normals), _p(normals_raw), _p(sems), _p(deltas), _p(ts),
                                _p(alive), _f32(T_thr), ctypes.c_int(classes), _p(np.ascontiguousarray(n_eff, np.int32)),
                                ctypes.c_int(N), _i64(A), _p(opacity), _p(depth), _p(rgb), _p(normal), _p(normal_raw),
                                _p(sem))


def composite_refloss_fw(sigmas, normals_diff, normals_ori, deltas, ts, rays_a, T_thr):
    sigmas, normals_diff, normals_ori, deltas = map(_f, (sigmas, normals_diff, normals_ori, deltas))
    rays_a = np.ascontiguousarray(rays_a, np.int64)
    S, R = sigmas.shape[0], rays_a.shape[0]
    loss_o = np.empty(R, np.float32); loss_p = np.empty((R, 3), np.float32)
    lib().ref_composite_refloss_fw(_p(sigmas), _p(normals_diff), _p(normals_ori), _p(deltas), _p(rays_a), _f32(T_thr),
                                   _i64(S), _i64(R), _p(loss_o), _p(loss_p))
    return loss_o, loss_p


def composite_refloss_bw(dL_dloss_o, dL_dloss_p, sigmas, normals_diff, normals_ori, deltas, ts, rays_a, loss_o, loss_p,
                         T_thr):
    dL_dloss_o, dL_dloss_p, sigmas, normals_diff, normals_ori, deltas, loss_o, loss_p = map(
        _f, (dL_dloss_o, dL_dloss_p, sigmas, normals_diff, normals_ori, deltas, loss_o, loss_p))
    rays_a = np.ascontiguousarray(rays_a, np.int64)
    S, R = sigmas.shape[0], rays_a.shape[0]
    d_sig = np.empty(S, np.float32); d_diff = np.empty((S, 3), np.float32); d_ori = np.empty(S, np.float32)
    lib().ref_composite_refloss_bw(_p(dL_dloss_o), _p(dL_dloss_p), _p(sigmas), _p(normals_diff), _p(normals_ori),
                                   _p(deltas), _p(rays_a), _p(loss_o), _p(loss_p), _f32(T_thr), _i64(S), _i64(R),
                                   _p(d_sig), _p(d_diff), _p(d_ori))
    return d_sig, d_diff, d_ori


def distortion_loss_fw(ws, deltas, ts, rays_a):
    ws, deltas, ts = map(_f, (ws, deltas, ts)); rays_a = np.ascontiguousarray(rays_a, np.int64)
    S, R = ws.shape[0], rays_a.shape[0]
    loss = np.empty(R, np.float32); wi = np.empty(S, np.float32); wti = np.empty(S, np.float32)
    lib().ref_distortion_loss_fw(_p(ws), _p(deltas), _p(ts), _p(rays_a), _i64(S), _i64(R), _p(loss), _p(wi), _p(wti))
    return loss, wi, wti


def distortion_loss_bw(dL_dloss, ws_incl, wts_incl, ws, deltas, ts, rays_a):
    dL_dloss, ws_incl, wts_incl, ws, deltas, ts = map(_f, (dL_dloss, ws_incl, wts_incl, ws, deltas, ts))
    rays_a = np.ascontiguousarray(rays_a, np.int64)
    S, R = ws.shape[0], rays_a.shape[0]
    out = np.empty(S, np.float32)
    lib().ref_distortion_loss_bw(_p(dL_dloss), _p(ws_incl), _p(wts_incl), _p(ws), _p(deltas), _p(ts), _p(rays_a),
                                 _i64(S), _i64(R), _p(out))
    return out
