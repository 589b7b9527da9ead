"""ctypes/numpy front-end of the CPU oracle (oracle/cnc_oracle.c, cnc_oracle_march.c).

TEST INFRASTRUCTURE ONLY.  The product package `cnc_b200` never imports this module;
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs do, and only as the checker / baseline.

Parity status is documented in the header of cnc_oracle.c (coder and SH: **parity
unpinned** -- third-party torchac / tinycudann sources are not in the reference tree).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcnc_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("cnc_oracle.c", "cnc_oracle_march.c")]
    if (not force) and os.path.exists(_SO) and all(
        os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs
    ):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-B", "libcnc_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.cnc_o_ac_encode.restype = C.c_int64
    return _lib


def _p(a, ctype=None):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle needs contiguous arrays"
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(np.asarray(a), dtype=dt)


# --------------------------------------------------------------------------- layout
def grid_layout(num_dim, resolutions, log2_hashmap_size):
    """offsets_list of GridEncoder.__init__ (ngp.py:197-208): int32 [L+1]."""
    offs, off = [], 0
    maxp = 2 ** log2_hashmap_size
    for r in resolutions:
        n = min(maxp, int(r) ** num_dim)
        n = int(np.ceil(n / 8) * 8)
        offs.append(off)
        off += n
    offs.append(off)
    return np.array(offs, dtype=np.int32)


# --------------------------------------------------------------------------- a1
def grid_rows(pos, T, res):
    pos = _c(pos, np.uint32)
    N, D = pos.shape
    rows = np.empty(N, np.uint32)
    lib().cnc_o_grid_rows(_p(pos), C.c_uint32(N), C.c_uint32(D), C.c_uint32(T), C.c_uint32(res), _p(rows))
    return rows


# --------------------------------------------------------------------------- a2 / a3
def grid_encode_fwd(x, table, offsets, resolutions, n_levels, Rb=128, binary_vxl=None,
                    min_level_id=None, return_rows=False):
    """out [L,N,F] float32 exactly like the kernel (ngp.py:80); offsets/resolutions are the
    (possibly sliced) lists the Python caller hands to the extension (ngp.py:86-109)."""
    x = _c(x, np.float32)
    table = _c(table, np.float32)
    offsets = _c(offsets, np.int32)
    resolutions = _c(resolutions, np.int32)
    N, D = x.shape
    F = table.shape[1]
    out = np.zeros((n_levels, N, F), np.float32)
    vx = None if binary_vxl is None else _c(binary_vxl, np.uint8)
    ml = None if min_level_id is None else _c(min_level_id, np.int32)
    rows = np.empty((n_levels, N, 1 << D), np.int64) if return_rows else None
    rc = lib().cnc_o_grid_encode_fwd(_p(x), _p(table), _p(offsets), _p(resolutions), _p(out),
                                     C.c_uint32(N), C.c_uint32(D), C.c_uint32(F), C.c_uint32(n_levels),
                                     C.c_uint32(Rb), _p(vx), _p(ml), _p(rows))
    assert rc == 0
    return (out, rows) if return_rows else out


def grid_encode_bwd(grad, x, n_rows, offsets, resolutions, n_levels, Rb=128, binary_vxl=None,
                    min_level_id=None, acc64=True):
    grad = _c(grad, np.float32)
    x = _c(x, np.float32)
    offsets = _c(offsets, np.int32)
    resolutions = _c(resolutions, np.int32)
    N, D = x.shape
    F = grad.shape[-1]
    gt = np.zeros((n_rows, F), np.float32)
    vx = None if binary_vxl is None else _c(binary_vxl, np.uint8)
    ml = None if min_level_id is None else _c(min_level_id, np.int32)
    rc = lib().cnc_o_grid_encode_bwd(_p(grad), _p(x), _p(offsets), _p(resolutions), _p(gt),
                                     C.c_uint32(N), C.c_uint32(D), C.c_uint32(F), C.c_uint32(n_levels),
                                     C.c_uint32(Rb), _p(vx), _p(ml), C.c_int(int(acc64)), C.c_uint64(n_rows))
    assert rc == 0
    return gt


def ste_binary(p):
    p = _c(p, np.float32)
    out = np.empty_like(p)
    lib().cnc_o_ste_binary(_p(p), _p(out), C.c_uint64(p.size))
    return out


# --------------------------------------------------------------------------- a10 / a11 / a12
def query_mask(points_i16, binary_vxl, resolution=None, resolution_list=None):
    pts = _c(points_i16, np.int16)
    vx = _c(binary_vxl, np.uint8)
    N, D = pts.shape
    Rb = vx.shape[0]
    mask = np.zeros(N, np.int16)
    ov = np.zeros(N, np.int32)
    rl = None if resolution_list is None else _c(resolution_list, np.int64)
    rc = lib().cnc_o_query_mask(_p(pts), _p(vx), C.c_int32(Rb), _p(mask), _p(ov), _p(rl),
                                C.c_int32(0 if resolution is None else int(resolution)),
                                C.c_int64(N), C.c_int32(D))
    assert rc == 0
    return mask, ov


def align_pack_fwd(feat, cnt, V=0.0):
    feat = _c(feat, np.float32)
    cnt = _c(cnt, np.int64)
    cs = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    N, M, F = cnt.size, int(cnt.max()) if cnt.size else 0, feat.shape[1]
    out = np.zeros((N, M, F), np.float32)
    lib().cnc_o_align_pack_fwd(_p(feat), _p(cnt), _p(cs), _p(out), C.c_int64(N), C.c_int64(M),
                               C.c_int64(F), C.c_float(V))
    return out


def align_pack_bwd(dpacked, cnt):
    dp = _c(dpacked, np.float32)
    cnt = _c(cnt, np.int64)
    cs = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    N, M, F = dp.shape
    out = np.zeros((int(cs[-1]), F), np.float32)
    lib().cnc_o_align_pack_bwd(_p(dp), _p(cnt), _p(cs), _p(out), C.c_int64(N), C.c_int64(M), C.c_int64(F))
    return out


def vote_planes_fwd(pts_i16, table, resolution, hashmap_size, axis):
    pts = _c(pts_i16, np.int16)
    table = _c(table, np.float32)
    F = table.shape[1]
    s = resolution - 2
    out = np.zeros((s, s, F, 2), np.float32)
    lib().cnc_o_vote_planes_fwd(_p(pts), _p(table), _p(out), C.c_uint32(pts.shape[0]),
                                C.c_uint32(resolution), C.c_uint32(F), C.c_uint32(hashmap_size), C.c_uint32(axis))
    return out


def vote_planes_bwd(pts_i16, table, out_sum, grad, resolution, hashmap_size, axis):
    pts = _c(pts_i16, np.int16)
    table = _c(table, np.float32)
    out_sum = _c(out_sum, np.float32)
    grad = _c(grad, np.float32)
    gt = np.zeros_like(table)
    lib().cnc_o_vote_planes_bwd(_p(pts), _p(table), _p(out_sum), _p(grad), _p(gt), C.c_uint32(pts.shape[0]),
                                C.c_uint32(resolution), C.c_uint32(table.shape[1]), C.c_uint32(hashmap_size),
                                C.c_uint32(axis))
    return gt


# --------------------------------------------------------------------------- a14 coder
def cdf_from_p(p):
    p = _c(p, np.float32).reshape(-1)
    c1 = np.empty(p.size, np.uint16)
    lib().cnc_o_cdf_from_p(_p(p), _p(c1), C.c_uint64(p.size))
    return c1


def ac_encode(c1, sym) -> bytes:
    c1 = _c(c1, np.uint16).reshape(-1)
    sym = _c(sym, np.uint8).reshape(-1)
    assert c1.size == sym.size
    cap = max(64, c1.size // 2 + 64)
    while True:
        buf = np.empty(cap, np.uint8)
        n = lib().cnc_o_ac_encode(_p(c1), _p(sym), C.c_uint64(c1.size), _p(buf), C.c_uint64(cap))
        if n <= cap:
            return buf[:n].tobytes()
        cap = int(n) + 64


def ac_decode(c1, data: bytes):
    c1 = _c(c1, np.uint16).reshape(-1)
    buf = np.frombuffer(data, np.uint8).copy() if len(data) else np.zeros(1, np.uint8)
    sym = np.zeros(c1.size, np.uint8)
    lib().cnc_o_ac_decode(_p(c1), C.c_uint64(c1.size), _p(buf), C.c_uint64(len(data)), _p(sym))
    return sym


def encode_float_p(x_pm1, p) -> bytes:
    """reference `encoder(x, p, file)` minus the file write (utils_bpp_acc.py:77-93)."""
    sym = ((np.asarray(x_pm1).reshape(-1) + 1) // 2).astype(np.uint8)
    return ac_encode(cdf_from_p(p), sym)


def decode_float_p(p, data: bytes):
    """reference `decoder(p, file)` (utils_bpp_acc.py:95-110): returns +-1 float32."""
    return ac_decode(cdf_from_p(p), data).astype(np.float32) * 2 - 1


# --------------------------------------------------------------------------- SH / embedder / MLP
def sh16(d01, fp16_round=True):
    d = _c(d01, np.float32)
    out = np.empty((d.shape[0], 16), np.float32)
    lib().cnc_o_sh16(_p(d), _p(out), C.c_uint64(d.shape[0]), C.c_int(int(fp16_round)))
    return out


def freq_embed(x, multires=10):
    """Embedder.embed (ngp.py:569-617): [x, sin(2^0 x), cos(2^0 x), ..., cos(2^9 x)] -> 3+60."""
    x = np.asarray(x, np.float32)
    outs = [x]
    for k in range(multires):
        f = np.float32(2.0 ** k)
        outs += [np.sin(x * f), np.cos(x * f)]
    return np.concatenate(outs, -1).astype(np.float32)


# --------------------------------------------------------------------------- a9 scans / volrend
def pack_info(ray_indices, n_rays):
    ri = _c(ray_indices, np.int64)
    out = np.zeros((n_rays, 2), np.int64)
    lib().cnc_o_pack_info(_p(ri), C.c_int64(ri.size), C.c_int64(n_rays), _p(out))
    return out


def packed_scan(x, packed, op="sum", inclusive=False):
    x = _c(x, np.float32)
    packed = _c(packed, np.int64)
    out = np.zeros_like(x)
    lib().cnc_o_packed_scan(_p(x), _p(packed), C.c_int64(packed.shape[0]), _p(out),
                            C.c_int(0 if op == "sum" else 1), C.c_int(int(inclusive)))
    return out


def render_from_density(t0, t1, sigma, packed, rgb=None):
    t0, t1, sigma = _c(t0, np.float32), _c(t1, np.float32), _c(sigma, np.float32)
    packed = _c(packed, np.int64)
    n, R = t0.size, packed.shape[0]
    w, T, a = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
    rgbc = None if rgb is None else _c(rgb, np.float32)
    col = np.zeros((R, 3), np.float32)
    op, dp = np.zeros(R, np.float32), np.zeros(R, np.float32)
    lib().cnc_o_render_from_density(_p(t0), _p(t1), _p(sigma), _p(rgbc), _p(packed), C.c_int64(R),
                                    _p(w), _p(T), _p(a), _p(col), _p(op), _p(dp))
    return dict(weights=w, trans=T, alphas=a, colors=col, opacities=op, depths=dp)


# --------------------------------------------------------------------------- a8 marching
def ray_aabb_intersect(rays_o, rays_d, aabbs, near=-np.inf, far=np.inf, miss=np.inf):
    o, d, bb = _c(rays_o, np.float32), _c(rays_d, np.float32), _c(aabbs, np.float32).reshape(-1, 6)
    n, m = o.shape[0], bb.shape[0]
    tmin, tmax = np.empty((n, m), np.float32), np.empty((n, m), np.float32)
    hits = np.empty((n, m), np.uint8)
    lib().cnc_o_ray_aabb_intersect(_p(o), _p(d), C.c_int32(n), C.c_float(near), C.c_float(far), _p(bb),
                                   C.c_int32(m), C.c_float(miss), _p(tmin), _p(tmax), _p(hits))
    return tmin, tmax, hits.astype(bool)


def traverse_grids(rays_o, rays_d, binaries, aabbs, near_planes=None, far_planes=None,
                   step_size=1e-3, cone_angle=0.0, traverse_steps_limit=0, rays_mask=None):
    """nerfacc.grid.traverse_grids (grid.py:94-194) -> (t_starts, t_ends, ray_indices, packed_info,
    terminate_planes).  Two passes like the reference host code (grid.cu:441-507)."""
    o, d = _c(rays_o, np.float32), _c(rays_d, np.float32)
    bins = _c(binaries, np.uint8)
    bb = _c(aabbs, np.float32).reshape(-1, 6)
    n, G = o.shape[0], bb.shape[0]
    nearp = np.zeros(n, np.float32) if near_planes is None else _c(near_planes, np.float32)
    farp = np.full(n, np.inf, np.float32) if far_planes is None else _c(far_planes, np.float32)
    tmin, tmax, hits = ray_aabb_intersect(o, d, bb)  # grid.py:157
    t = np.concatenate([tmin, tmax], -1)
    idx = np.argsort(t, axis=-1, kind="stable").astype(np.int64)
    ts = np.take_along_axis(t, idx, -1).astype(np.float32)
    hits8 = _c(hits, np.uint8)
    mask = None if rays_mask is None else _c(rays_mask, np.uint8)
    cnt = np.zeros(n, np.int64)
    term = np.zeros(n, np.float32)
    args = lambda starts, t0, t1, ri: (
        _p(o), _p(d), _p(mask), C.c_int32(n), C.c_int32(G), C.c_int32(bins.shape[-3]),
        C.c_int32(bins.shape[-2]), C.c_int32(bins.shape[-1]), _p(bins), _p(bb), _p(hits8), _p(ts),
        _p(idx), _p(nearp), _p(farp), C.c_float(step_size), C.c_float(cone_angle),
        C.c_int32(traverse_steps_limit), _p(starts), _p(cnt), _p(t0), _p(t1), _p(ri), _p(term))
    lib().cnc_o_traverse_grids(*args(None, None, None, None))
    starts = (np.cumsum(cnt) - cnt).astype(np.int64)
    tot = int(cnt.sum())
    t0, t1 = np.zeros(tot, np.float32), np.zeros(tot, np.float32)
    ri = np.zeros(tot, np.int64)
    if tot:
        lib().cnc_o_traverse_grids(*args(starts, t0, t1, ri))
    return t0, t1, ri, np.stack([starts, cnt], -1), term
