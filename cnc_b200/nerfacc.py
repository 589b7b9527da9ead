"""Host-side mirror of the slice of (vendored, patched) nerfacc 0.5.3 that the CNC hot path calls.

    OccGridEstimator(roi_aabb, resolution, levels)            nerfacc/estimators/occ_grid.py:19-424
        .sampling(rays_o, rays_d, sigma_fn=..., render_step_size=..., stratified=..., ...)   :88-239
        .update_every_n_steps(step, occ_eval_fn, occ_thre, ema_decay, warmup_steps, n)        :242-277
        .binaries  .aabbs  .occs
    ray_aabb_intersect, traverse_grids                        nerfacc/grid.py:20-91, :94-194
    rendering (rgb_sigma_fn returns (rgbs, sigmas, positions): the CNC patch)    nerfacc/volrend.py:14-160
    render_weight_from_density / _transmittance_ / _visibility_               :211-266, :314-364, :423-482
    accumulate_along_rays, accumulate_along_rays_                               :485-575
    pack_info, inclusive_sum, exclusive_sum, inclusive_prod, exclusive_prod     nerfacc/pack.py, scan.py

Same names, arguments and return values.  The CUDA side is csrc/march_render.cu: thread-per-ray DDA
(two passes + one scan, like grid.cu:441-507), and a fused weights + accumulation kernel that replaces
exclusive_sum + exp + 3 x index_add_ (no atomics, fixed summation order).  No CPU path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Tuple, Union

import torch
from torch import Tensor

from ._lib import check, lib, need_cuda, ptr, stream


# ------------------------------------------------------------------------------------------ data specs
@dataclass
class RaySamples:
    """nerfacc/data_specs.py RaySamples (the fields the callers read)."""
    vals: Tensor
    packed_info: Optional[Tensor] = None
    ray_indices: Optional[Tensor] = None
    is_valid: Optional[Tensor] = None


class RayIntervals:
    """nerfacc/data_specs.py RayIntervals.  Edges are stored per sample (left, right), so `vals[is_left]` / `vals[is_right]`
    are t_starts / t_ends like in occ_grid.py:188-189.  The march produces the two edge arrays directly (`t_starts`,
    `t_ends`, not in nerfacc); the interleaved reference layout (`vals`, `is_left`, `is_right`, `ray_indices`,
    `packed_info`) is materialised on first access, so callers inside this package pay for none of it."""

    def __init__(self, vals=None, packed_info=None, ray_indices=None, is_left=None, is_right=None, t_starts=None, t_ends=None,
                 sample_packed_info=None, sample_ray_indices=None):
        self._vals, self._packed_info, self._ray_indices, self._is_left, self._is_right = vals, packed_info, ray_indices, is_left, is_right
        self.t_starts, self.t_ends = t_starts, t_ends
        self._spk, self._sri = sample_packed_info, sample_ray_indices

    @property
    def vals(self):
        if self._vals is None and self.t_starts is not None:
            self._vals = torch.stack([self.t_starts, self.t_ends], dim=-1).reshape(-1)
        return self._vals

    @property
    def is_left(self):
        if self._is_left is None and self.t_starts is not None:
            self._is_left = torch.zeros(2 * self.t_starts.numel(), dtype=torch.bool, device=self.t_starts.device)
            self._is_left[0::2] = True
        return self._is_left

    @property
    def is_right(self):
        if self._is_right is None and self.t_starts is not None:
            self._is_right = ~self.is_left
        return self._is_right

    @property
    def ray_indices(self):
        if self._ray_indices is None and self._sri is not None:
            self._ray_indices = self._sri.repeat_interleave(2)
        return self._ray_indices

    @property
    def packed_info(self):
        if self._packed_info is None and self._spk is not None:
            self._packed_info = self._spk * 2
        return self._packed_info


class _LazySamples(RaySamples):
    """RaySamples of an exact-size march: `vals` (the sample midpoints) and `is_valid` (all true) on first access"""

    def __init__(self, t0, t1, packed_info, ray_indices):
        self._t0, self._t1, self._v, self._ok = t0, t1, None, None
        self.packed_info, self.ray_indices = packed_info, ray_indices

    @property
    def vals(self):
        if self._v is None:
            self._v = (self._t0 + self._t1) * 0.5
        return self._v

    @vals.setter
    def vals(self, v):
        self._v = v

    @property
    def is_valid(self):
        if self._ok is None:
            self._ok = torch.ones(self._t0.numel(), dtype=torch.bool, device=self._t0.device)
        return self._ok

    @is_valid.setter
    def is_valid(self, v):
        self._ok = v


# ------------------------------------------------------------------------------------------ pack / scans
def pack_info(ray_indices: Tensor, n_rays: Optional[int] = None) -> Tensor:
    """(n_rays, 2) [start, count] of every ray's chunk; ray_indices must be sorted (nerfacc/pack.py:10-49)."""
    assert ray_indices.dim() == 1
    if not ray_indices.is_cuda:
        raise NotImplementedError("Only support cuda inputs.")
    if n_rays is None:
        n_rays = int(ray_indices.max()) + 1
    cnt = torch.zeros(n_rays, dtype=torch.long, device=ray_indices.device)
    cnt.index_add_(0, ray_indices, torch.ones_like(ray_indices))
    starts = cnt.cumsum(0) - cnt
    return torch.stack([starts, cnt], dim=-1)


def _scan(inputs: Tensor, packed_info: Optional[Tensor], op: int, inclusive: bool, indices=None) -> Tensor:
    if packed_info is None:
        if indices is not None:
            packed_info = pack_info(indices)
        else:   # batched [n_rays, n_samples] layout
            n, m = inputs.shape
            packed_info = torch.stack([torch.arange(n, device=inputs.device) * m,
                                       torch.full((n,), m, device=inputs.device)], -1)
    x = inputs.contiguous().float()
    need_cuda(inputs=x, packed_info=packed_info)
    out = torch.empty_like(x)
    pk = packed_info.contiguous().to(torch.int64)
    check(lib().cnc_packed_scan(ptr(x), ptr(pk), pk.shape[0], ptr(out), op, int(inclusive), 0, stream()))
    return out


def inclusive_sum(inputs, packed_info=None, indices=None):
    """nerfacc/scan.py:14-53"""
    return _scan(inputs, packed_info, 0, True, indices)


def exclusive_sum(inputs, packed_info=None, indices=None):
    """nerfacc/scan.py:56-100"""
    return _scan(inputs, packed_info, 0, False, indices)


def inclusive_prod(inputs, packed_info=None, indices=None):
    """nerfacc/scan.py:103-146"""
    return _scan(inputs, packed_info, 1, True, indices)


def exclusive_prod(inputs, packed_info=None, indices=None):
    """nerfacc/scan.py:149-194"""
    return _scan(inputs, packed_info, 1, False, indices)


# ------------------------------------------------------------------------------------------ marching
@torch.no_grad()
def ray_aabb_intersect(rays_o: Tensor, rays_d: Tensor, aabbs: Tensor, near_plane: float = -float("inf"),
                       far_plane: float = float("inf"), miss_value: float = float("inf")):
    """(t_mins, t_maxs, hits), each [n_rays, n_aabbs].  nerfacc/grid.py:20-91"""
    assert rays_o.ndim == 2 and rays_o.shape[-1] == 3 and rays_d.shape == rays_o.shape
    assert aabbs.ndim == 2 and aabbs.shape[-1] == 6
    o, d, bb = rays_o.contiguous().float(), rays_d.contiguous().float(), aabbs.contiguous().float()
    need_cuda(rays_o=o, rays_d=d, aabbs=bb)
    n, m = o.shape[0], bb.shape[0]
    t_mins = torch.empty(n, m, device=o.device)
    t_maxs = torch.empty(n, m, device=o.device)
    hits = torch.empty(n, m, dtype=torch.uint8, device=o.device)
    check(lib().cnc_ray_aabb_intersect(ptr(o), ptr(d), n, float(near_plane), float(far_plane), ptr(bb), m, float(miss_value),
                                       ptr(t_mins), ptr(t_maxs), ptr(hits), stream()))
    return t_mins, t_maxs, hits.bool()


_INDEX01 = {}
_SCRATCH = {}
_CAPS = {}


def _ray_capacity(aabbs: Tensor, step_size: float) -> int:
    """upper bound of the samples one ray can produce: the diagonal of the largest box in steps (+ slack per level)"""
    key = (aabbs.data_ptr(), aabbs._version, step_size)
    c = _CAPS.get(key)
    if c is None:
        if len(_CAPS) > 16:
            _CAPS.clear()
        diag = float((aabbs[:, 3:] - aabbs[:, :3]).norm(dim=-1).max())
        c = _CAPS[key] = int(diag / step_size) + 2 * aabbs.shape[0] + 16
    return c


def _index01(n: int, dev) -> Tensor:
    """[[0, 1]] * n (int64): what torch.sort returns as indices for a single box"""
    key = (str(dev), n)
    t = _INDEX01.get(key)
    if t is None:
        if len(_INDEX01) > 16:
            _INDEX01.clear()
        t = _INDEX01[key] = torch.tensor([0, 1], dtype=torch.int64, device=dev).repeat(n, 1)
    return t


@torch.no_grad()
def traverse_grids(rays_o: Tensor, rays_d: Tensor, binaries: Tensor, aabbs: Tensor, near_planes: Optional[Tensor] = None,
                   far_planes: Optional[Tensor] = None, step_size: Optional[float] = 1e-3, cone_angle: Optional[float] = 0.0,
                   traverse_steps_limit: Optional[int] = None, over_allocate: Optional[bool] = False,
                   rays_mask: Optional[Tensor] = None, t_sorted: Optional[Tensor] = None, t_indices: Optional[Tensor] = None,
                   hits: Optional[Tensor] = None) -> Tuple[RayIntervals, RaySamples, Tensor]:
    """March rays through the (multi-level) occupancy grid.  nerfacc/grid.py:94-194.
    `over_allocate` is accepted for compatibility; sizes are always exact (count pass, scan, fill pass)."""
    o, d = rays_o.contiguous().float(), rays_d.contiguous().float()
    need_cuda(rays_o=o, rays_d=d, binaries=binaries, aabbs=aabbs)
    n, dev = o.shape[0], o.device
    if near_planes is None:
        near_planes = torch.zeros_like(o[:, 0])
    if far_planes is None:
        far_planes = torch.full_like(o[:, 0], float("inf"))
    if rays_mask is None:
        rays_mask_u8 = None
    else:
        rays_mask_u8 = rays_mask.contiguous().to(torch.uint8)
    if traverse_steps_limit is None:
        traverse_steps_limit = -1
    if t_sorted is None or t_indices is None or hits is None:
        t_mins, t_maxs, hits = ray_aabb_intersect(o, d, aabbs)
        if aabbs.shape[0] == 1:     # one box: entry before exit (a miss has both at the same value; the stable order is 0, 1)
            t_sorted = torch.cat([t_mins, t_maxs], dim=-1)
            t_indices = _index01(n, dev)
        else:
            t_sorted, t_indices = torch.sort(torch.cat([t_mins, t_maxs], dim=-1), dim=-1)
    G = aabbs.shape[0]
    bins = binaries.contiguous()
    bins_u8 = bins.view(torch.uint8) if bins.dtype == torch.bool else bins.to(torch.uint8)
    hits_u8 = hits.contiguous().to(torch.uint8)
    bb = aabbs.contiguous().float()
    ts, ti = t_sorted.contiguous().float(), t_indices.contiguous().to(torch.int64)
    nearp, farp = near_planes.contiguous().float(), far_planes.contiguous().float()
    cnt = torch.empty(n, dtype=torch.int64, device=dev)
    term = torch.empty(n, device=dev)
    L = lib()

    def run(starts, t0, t1, ri, c, tm):
        check(L.cnc_traverse_grids(ptr(o), ptr(d), ptr(rays_mask_u8), n, G, bins.shape[-3], bins.shape[-2], bins.shape[-1],
                                   ptr(bins_u8), ptr(bb), ptr(hits_u8), ptr(ts), ptr(ti), ptr(nearp), ptr(farp),
                                   float(step_size), float(cone_angle), int(traverse_steps_limit), ptr(starts), ptr(c),
                                   ptr(t0), ptr(t1), ptr(ri), ptr(tm), stream()))

    cap = _ray_capacity(bb, float(step_size)) if (traverse_steps_limit <= 0 and step_size and step_size > 0) else 0
    one_walk = 0 < cap and 0 < n <= 148 * 64 * 2 and n * cap <= (1 << 26)
    if one_walk:
        # ONE walk: every ray writes its samples into its own slice of a scratch buffer (cap = an upper bound of the samples
        # of a ray: box diagonal / step) and its count; a copy kernel packs them.  The walk is the cost (a serial chain per
        # ray), so this is half of count pass + fill pass.  A ray that fills its slice (cannot happen within the bound) sends
        # the call down the two-pass road.
        key = (str(dev), n, cap, stream())       # (per stream: the slices are live until the pack kernel has run)
        sc = _SCRATCH.get(key)
        if sc is None:
            if len(_SCRATCH) > 4:
                _SCRATCH.clear()
            sc = _SCRATCH[key] = (torch.empty(n * cap, device=dev), torch.empty(n * cap, device=dev),
                                  torch.arange(n, dtype=torch.int64, device=dev) * cap)
        check(L.cnc_traverse_grids(ptr(o), ptr(d), ptr(rays_mask_u8), n, G, bins.shape[-3], bins.shape[-2], bins.shape[-1],
                                   ptr(bins_u8), ptr(bb), ptr(hits_u8), ptr(ts), ptr(ti), ptr(nearp), ptr(farp),
                                   float(step_size), float(cone_angle), cap, ptr(sc[2]), ptr(cnt), ptr(sc[0]), ptr(sc[1]), None,
                                   ptr(term), stream()))
        ends = cnt.cumsum(0)
        starts = ends - cnt
        total, most = torch.stack([ends[-1], cnt.max()]).tolist()      # the one host sync of the reference too (data_spec.hpp:91)
        one_walk = most < cap
    if one_walk:
        t0 = torch.empty(total, device=dev)
        t1 = torch.empty(total, device=dev)
        ri = torch.empty(total, dtype=torch.int64, device=dev)
        packed = torch.stack([starts, cnt], dim=-1)
        if total:
            check(L.cnc_pack_ray_chunks(ptr(sc[0]), ptr(sc[1]), cap, ptr(packed), n, ptr(t0), ptr(t1), ptr(ri), stream()))
        if rays_mask_u8 is not None:
            term = torch.where(rays_mask_u8.bool(), term, nearp)     # (masked rays: the kernel leaves their plane untouched)
    else:
        run(None, None, None, None, cnt, None)
        starts = cnt.cumsum(0) - cnt
        total = int(cnt.sum())                       # the one host sync of the reference too (data_spec.hpp:91)
        t0 = torch.empty(total, device=dev)
        t1 = torch.empty(total, device=dev)
        ri = torch.empty(total, dtype=torch.int64, device=dev)
        if total:
            run(starts, t0, t1, ri, None, term)
        else:
            term.copy_(nearp)
        packed = torch.stack([starts, cnt], dim=-1)
    intervals = RayIntervals(t_starts=t0, t_ends=t1, sample_packed_info=packed, sample_ray_indices=ri)
    samples = _LazySamples(t0, t1, packed, ri)
    return intervals, samples, term


def _compact(masks: Tensor, t_starts: Tensor, t_ends: Tensor, ray_indices: Tensor, packed_info: Tensor):
    """(ray_indices[masks], t_starts[masks], t_ends[masks]) (occ_grid.py:192-197) as one prefix sum, one host read of the
    count and one kernel; the (start, count) table of what is kept rides along on the returned `ray_indices`
    (`_packed` below picks it up, so `rendering` does not rebuild it with an index_add over all samples)."""
    n = masks.shape[0]
    if n == 0:
        return ray_indices, t_starts, t_ends
    keep = masks.contiguous().view(torch.uint8) if masks.dtype == torch.bool else masks.contiguous().to(torch.uint8)
    rank = keep.cumsum(0, dtype=torch.int64)
    total = int(rank[-1])                          # the host sync `nonzero` / boolean indexing has as well
    dev = masks.device
    t0, t1 = torch.empty(total, device=dev), torch.empty(total, device=dev)
    ri = torch.empty(total, dtype=torch.int64, device=dev)
    R = packed_info.shape[0]
    if total == 0:
        ri._cnc_packed = torch.zeros(R, 2, dtype=torch.int64, device=dev)
        return ri, t0, t1
    pk = torch.empty(R, 2, dtype=torch.int64, device=dev)
    check(lib().cnc_compact_samples(ptr(keep), ptr(rank), n, ptr(t_starts.contiguous()), ptr(t_ends.contiguous()),
                                    ptr(ray_indices.contiguous()), ptr(t0), ptr(t1), ptr(ri), ptr(packed_info.contiguous()), R,
                                    ptr(pk), stream()))
    ri._cnc_packed = pk
    return ri, t0, t1


# ------------------------------------------------------------------------------------------ volume rendering
def _packed(packed_info, ray_indices, n_rays, like):
    if packed_info is None:
        if ray_indices is None:
            raise ValueError("flattened samples need packed_info or ray_indices")
        packed_info = getattr(ray_indices, "_cnc_packed", None)      # left by OccGridEstimator.sampling for exactly this tensor
        if packed_info is None or packed_info.shape[0] != n_rays:
            packed_info = pack_info(ray_indices, n_rays)
    return packed_info.contiguous().to(torch.int64)


class _RenderDensity(torch.autograd.Function):
    """weights / transmittance / alphas from densities with the analytic backward of
    w_i = exp(-sum_{j<i} s_j d_j) (1 - exp(-s_i d_i))  (volrend.py:211-266,314-364; scan backward scan.cu:100-110)."""

    @staticmethod
    def forward(ctx, t_starts, t_ends, sigmas, packed_info, prefix_trans):
        t0, t1, sg = t_starts.contiguous().float(), t_ends.contiguous().float(), sigmas.contiguous().float()
        need_cuda(t_starts=t0, t_ends=t1, sigmas=sg)
        w, T, al = torch.empty_like(sg), torch.empty_like(sg), torch.empty_like(sg)
        pt = None if prefix_trans is None else prefix_trans.contiguous().float()
        check(lib().cnc_render_from_density(ptr(t0), ptr(t1), ptr(sg), None, ptr(packed_info), packed_info.shape[0], ptr(pt),
                                            ptr(w), ptr(T), ptr(al), None, None, None, stream()))
        ctx.save_for_backward(t0, t1, T, al, packed_info)
        return w, T, al

    @staticmethod
    def backward(ctx, gw, gT, ga):
        t0, t1, T, al, packed_info = ctx.saved_tensors
        dt = t1 - t0
        gw = torch.zeros_like(T) if gw is None else gw
        gT = torch.zeros_like(T) if gT is None else gT
        ga = torch.zeros_like(T) if ga is None else ga
        # d/d(sd_k): alpha term at k, transmittance term for every later sample of the ray
        g_alpha = (gw * T + ga) * (1 - al)
        g_trans = (gw * al + gT) * T                      # dL/dT_i * T_i  (T_i = exp(-excl_sum))
        out = torch.empty_like(T)
        check(lib().cnc_packed_scan(ptr(g_trans.contiguous()), ptr(packed_info), packed_info.shape[0], ptr(out), 0, 0, 1, stream()))
        return None, None, (g_alpha - out) * dt, None, None


class _RenderAll(torch.autograd.Function):
    """weights / transmittance / alphas AND the three per-ray accumulations (colour, opacity, weighted depth) of
    `rendering` (volrend.py:14-160) in one warp-per-ray pass -- no `index_add_` atomics, a fixed summation tree -- with the
    analytic backward: d/dw_i = gC_r . c_i + gO_r + gD_r m_i, d/dc_i = w_i gC_r, then the backward of `_RenderDensity`."""

    @staticmethod
    def forward(ctx, t_starts, t_ends, sigmas, rgbs, packed_info, ray_indices):
        t0, t1, sg, c = (x.contiguous().float() for x in (t_starts, t_ends, sigmas, rgbs))
        need_cuda(t_starts=t0, t_ends=t1, sigmas=sg, rgbs=c)
        R = packed_info.shape[0]
        w, T, al = torch.empty_like(sg), torch.empty_like(sg), torch.empty_like(sg)
        colors = torch.empty(R, 3, device=sg.device)
        opac, depth = torch.empty(R, device=sg.device), torch.empty(R, device=sg.device)
        check(lib().cnc_render_from_density(ptr(t0), ptr(t1), ptr(sg), ptr(c), ptr(packed_info), R, None, ptr(w), ptr(T), ptr(al),
                                            ptr(colors), ptr(opac), ptr(depth), stream()))
        ctx.save_for_backward(t0, t1, T, al, w, c, packed_info, ray_indices)
        ctx.mark_non_differentiable(T, al)
        ctx.set_materialize_grads(False)     # outputs nobody differentiates arrive as None, not as zero-filled tensors
        return w, T, al, colors, opac, depth

    @staticmethod
    def backward(ctx, gw, gT, ga, gC, gO, gD):
        t0, t1, T, al, w, c, packed_info, ri = ctx.saved_tensors
        f = lambda g: None if g is None else g.contiguous().float()
        gw, gC, gO, gD = f(gw), f(gC), f(gO), f(gD)
        if gw is None and gC is None and gO is None and gD is None:
            return None, None, None, None, None, None
        g_sigma, g_rgb = torch.empty_like(T), torch.empty_like(c)
        check(lib().cnc_render_bwd(ptr(t0), ptr(t1), ptr(T), ptr(al), ptr(w), ptr(c), ptr(packed_info), packed_info.shape[0],
                                   ptr(gC), ptr(gO), ptr(gD), ptr(gw), ptr(g_sigma), ptr(g_rgb), stream()))
        return None, None, g_sigma, g_rgb, None, None


def render_weight_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None, n_rays=None, prefix_trans=None):
    """(weights, trans, alphas).  nerfacc/volrend.py:314-364"""
    if t_starts.dim() != 1:
        raise NotImplementedError("cnc_b200.nerfacc handles the flattened (packed) sample layout CNC uses")
    pk = _packed(packed_info, ray_indices, n_rays, sigmas)
    return _RenderDensity.apply(t_starts, t_ends, sigmas, pk, prefix_trans)


def render_transmittance_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None, n_rays=None, prefix_trans=None):
    """(trans, alphas).  nerfacc/volrend.py:211-266"""
    _, T, al = render_weight_from_density(t_starts, t_ends, sigmas, packed_info, ray_indices, n_rays, prefix_trans)
    return T, al


@torch.no_grad()
def render_visibility_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None, n_rays=None,
                                   early_stop_eps: float = 1e-4, alpha_thre: float = 0.0, prefix_trans=None):
    """nerfacc/volrend.py:423-482"""
    T, al = render_transmittance_from_density(t_starts, t_ends, sigmas, packed_info, ray_indices, n_rays, prefix_trans)
    vis = T >= early_stop_eps
    if alpha_thre > 0:
        vis = vis & (al >= alpha_thre)
    return vis


def accumulate_along_rays(weights, values=None, ray_indices=None, n_rays=None):
    """sum_i w_i * v_i per ray -> [n_rays, D].  nerfacc/volrend.py:485-549"""
    src = weights[..., None] if values is None else weights[..., None] * values
    if ray_indices is None:
        return src.sum(dim=-2)
    assert n_rays is not None
    out = torch.zeros(n_rays, src.shape[-1], device=src.device, dtype=src.dtype)
    out.index_add_(0, ray_indices, src)
    return out


def accumulate_along_rays_(weights, values=None, ray_indices=None, outputs=None) -> None:
    """in-place variant used by the test-time renderer.  nerfacc/volrend.py:552-575"""
    src = weights[..., None] if values is None else weights[..., None] * values
    if ray_indices is None:
        outputs.add_(src.sum(dim=-2))
    else:
        outputs.index_add_(0, ray_indices, src)


def rendering(t_starts, t_ends, ray_indices=None, n_rays=None, rgb_sigma_fn: Optional[Callable] = None,
              rgb_alpha_fn: Optional[Callable] = None, render_bkgd=None):
    """(colors, opacities, depths, extras).  nerfacc/volrend.py:14-160 with the CNC patch: `rgb_sigma_fn`
    returns (rgbs, sigmas, positions) and `extras` carries sigmas / rgbs / positions (volrend.py:89,108-115)."""
    if ray_indices is not None:
        assert t_starts.shape == t_ends.shape == ray_indices.shape
    if rgb_sigma_fn is None:
        raise ValueError("cnc_b200.nerfacc.rendering needs `rgb_sigma_fn` (the path CNC uses)")
    if t_starts.shape[0] != 0:
        rgbs, sigmas, positions = rgb_sigma_fn(t_starts, t_ends, ray_indices)
    else:
        positions = None
        rgbs = torch.empty((0, 3), device=t_starts.device)
        sigmas = torch.empty((0,), device=t_starts.device)
    assert rgbs.shape[-1] == 3 and sigmas.shape == t_starts.shape
    if t_starts.dim() == 1 and ray_indices is not None and t_starts.is_cuda and t_starts.shape[0] != 0:
        pk = _packed(None, ray_indices, n_rays, sigmas)
        weights, trans, alphas, colors, opac, dsum = _RenderAll.apply(t_starts, t_ends, sigmas, rgbs, pk, ray_indices)
        extras = {"weights": weights, "alphas": alphas, "trans": trans, "sigmas": sigmas, "rgbs": rgbs, "positions": positions}
        opacities, depths = opac.unsqueeze(-1), dsum.unsqueeze(-1)
    else:
        weights, trans, alphas = render_weight_from_density(t_starts, t_ends, sigmas, ray_indices=ray_indices, n_rays=n_rays)
        extras = {"weights": weights, "alphas": alphas, "trans": trans, "sigmas": sigmas, "rgbs": rgbs, "positions": positions}
        colors = accumulate_along_rays(weights, values=rgbs, ray_indices=ray_indices, n_rays=n_rays)
        opacities = accumulate_along_rays(weights, values=None, ray_indices=ray_indices, n_rays=n_rays)
        depths = accumulate_along_rays(weights, values=(t_starts + t_ends)[..., None] / 2.0, ray_indices=ray_indices, n_rays=n_rays)
    depths = depths / opacities.clamp_min(torch.finfo(rgbs.dtype).eps)
    if render_bkgd is not None:
        colors = colors + render_bkgd * (1.0 - opacities)
    return colors, opacities, depths, extras


@torch.no_grad()
def render_fused(t_starts, t_ends, sigmas, rgbs, packed_info, render_bkgd=None):
    """Inference-time tail in one kernel: (colors, opacities, depths) == rendering(...)[:3] for given sigmas / rgbs."""
    t0, t1, sg, c = (x.contiguous().float() for x in (t_starts, t_ends, sigmas, rgbs))
    need_cuda(t_starts=t0, sigmas=sg, rgbs=c)
    pk = packed_info.contiguous().to(torch.int64)
    R = pk.shape[0]
    colors = torch.empty(R, 3, device=t0.device)
    opac = torch.empty(R, device=t0.device)
    depth = torch.empty(R, device=t0.device)
    check(lib().cnc_render_from_density(ptr(t0), ptr(t1), ptr(sg), ptr(c), ptr(pk), R, None, None, None, None, ptr(colors),
                                        ptr(opac), ptr(depth), stream()))
    opac = opac[:, None]
    depth = depth[:, None] / opac.clamp_min(torch.finfo(torch.float32).eps)
    if render_bkgd is not None:
        colors = colors + render_bkgd * (1.0 - opac)
    return colors, opac, depth


# ------------------------------------------------------------------------------------------ estimator
def _enlarge_aabb(aabb, factor: float) -> Tensor:
    center = (aabb[:3] + aabb[3:]) / 2
    extent = (aabb[3:] - aabb[:3]) / 2
    return torch.cat([center - extent * factor, center + extent * factor])


class Premarch:
    """The occupancy march of a ray batch, issued before the batch is needed (not in nerfacc).

    The march depends on the rays and on the occupancy grid only -- not on the field -- so a training loop that knows its
    next batch can run it on a side stream while the current step's forward / backward occupy the device: count pass,
    prefix sum, and the fill pass into a buffer kept from step to step (no size is needed on the host in between; a batch
    that outgrows the buffer is filled again at its exact size when it is taken).  `take` hands the result to
    `OccGridEstimator.sampling`; the samples are the ones the in-line march produces, bit for bit: same kernels, same random
    jitter draw (the caller keeps the order of draws, see trainer.TrainStep)."""

    def __init__(self, device, capacity: int = 1 << 20):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.capacity = capacity
        self.sets = [None, None]          # two buffer sets: the consumer of one step may still hold views of the other
        self.turn = 0
        self.pending = None
        self.taken = 0                    # marches handed to a step so far
        self.total_host = torch.zeros(1, dtype=torch.int64).pin_memory()

    def _buffers(self, k):
        b = self.sets[k]
        if b is None or b[0].numel() < self.capacity:
            b = self.sets[k] = (torch.empty(self.capacity, device=self.device), torch.empty(self.capacity, device=self.device),
                                torch.empty(self.capacity, dtype=torch.int64, device=self.device))
        return b

    @torch.no_grad()
    def issue(self, est: "OccGridEstimator", rays_o: Tensor, rays_d: Tensor, near_plane: float = 0.0, far_plane: float = 1e10,
              t_min=None, t_max=None, render_step_size: float = 1e-3, stratified: bool = False, cone_angle: float = 0.0,
              after: Optional[torch.cuda.Event] = None) -> None:
        """start the march of (rays_o, rays_d) against the estimator's CURRENT grid.  It is ordered behind `after` (an event
        the caller recorded once the rays and the grid were final) or, without one, behind everything queued on the current
        stream so far."""
        entry = after
        if entry is None:
            entry = torch.cuda.Event()
            entry.record()
        o, d = rays_o.contiguous().float(), rays_d.contiguous().float()
        n, dev = o.shape[0], o.device
        k = self.turn
        self.turn ^= 1
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(entry)
            near_planes, far_planes = est._planes(o, near_plane, far_plane, t_min, t_max)
            if stratified:
                near_planes += torch.rand_like(near_planes) * render_step_size
            t_mins, t_maxs, hits = ray_aabb_intersect(o, d, est.aabbs)
            if est.aabbs.shape[0] == 1:
                t_sorted, t_indices = torch.cat([t_mins, t_maxs], dim=-1), _index01(n, dev)
            else:
                t_sorted, t_indices = torch.sort(torch.cat([t_mins, t_maxs], dim=-1), dim=-1)
            bins = est.binaries.contiguous()
            args = dict(o=o, d=d, bins=bins, bins_u8=bins.view(torch.uint8) if bins.dtype == torch.bool else bins.to(torch.uint8),
                        bb=est.aabbs.contiguous().float(), hits=hits.contiguous().to(torch.uint8), ts=t_sorted.contiguous().float(),
                        ti=t_indices.contiguous(), near=near_planes.contiguous(), far=far_planes.contiguous(),
                        step=float(render_step_size), cone=float(cone_angle))
            cnt = torch.empty(n, dtype=torch.int64, device=dev)
            _march_pass(args, None, None, None, None, None, cnt)
            ends = cnt.cumsum(0)
            starts = ends - cnt
            self.total_host.copy_(ends[-1:], non_blocking=True)
            counted = torch.cuda.Event()
            counted.record()
            t0, t1, ri = self._buffers(k)
            fits = (ends <= t0.numel()).to(torch.uint8)          # rays that would run past the buffer are left out (see take)
            _march_pass(args, fits, starts, t0, t1, ri, None)
            filled = torch.cuda.Event()
            filled.record()
        self.pending = dict(key=self._key(est, rays_o, rays_d, near_plane, far_plane, render_step_size, stratified, cone_angle),
                            args=args, starts=starts, cnt=cnt, bufs=(t0, t1, ri), counted=counted, filled=filled,
                            keep=(rays_o, rays_d))

    @staticmethod
    def _key(est, rays_o, rays_d, near_plane, far_plane, render_step_size, stratified, cone_angle):
        """what a pending march was issued for: the ray tensors (storage address AND version counter: a batch written in
        place into the same buffer is a different batch), their number, the march parameters, the grid object"""
        return (rays_o.data_ptr(), rays_o._version, rays_d.data_ptr(), rays_d._version, int(rays_o.shape[0]), float(near_plane),
                float(far_plane), float(render_step_size), bool(stratified), float(cone_angle), id(est.binaries), est.binaries._version)

    def matches(self, est, rays_o, rays_d, near_plane, far_plane, render_step_size, stratified, cone_angle) -> bool:
        p = self.pending
        return p is not None and p["key"] == self._key(est, rays_o, rays_d, near_plane, far_plane, render_step_size, stratified, cone_angle)

    def drop(self) -> None:
        self.pending = None

    @torch.no_grad()
    def take(self, est, rays_o, rays_d, near_plane, far_plane, t_min, t_max, render_step_size, stratified, cone_angle):
        """(t_starts, t_ends, ray_indices, packed_info) of the pending march; the current stream waits for it"""
        if t_min is not None or t_max is not None or not self.matches(est, rays_o, rays_d, near_plane, far_plane, render_step_size,
                                                                      stratified, cone_angle):
            raise RuntimeError("Premarch.take: no march pending for these rays and this grid (check with matches() first)")
        p, self.pending = self.pending, None
        self.taken += 1
        p["counted"].synchronize()
        total = int(self.total_host[0])
        cur = torch.cuda.current_stream()
        cur.wait_event(p["filled"])
        t0, t1, ri = p["bufs"]
        if total > t0.numel():                # outgrown: fill at the exact size now, and start the next one larger
            self.capacity = max(2 * self.capacity, int(total * 1.5))
            t0, t1 = torch.empty(total, device=self.device), torch.empty(total, device=self.device)
            ri = torch.empty(total, dtype=torch.int64, device=self.device)
            for t in p["args"].values():
                if isinstance(t, Tensor):
                    t.record_stream(cur)
            _march_pass(p["args"], None, p["starts"], t0, t1, ri, None)
        for t in (t0, t1, ri, p["starts"], p["cnt"]):
            t.record_stream(cur)
        packed = torch.stack([p["starts"], p["cnt"]], dim=-1)
        return t0[:total], t1[:total], ri[:total], packed


def _march_pass(a, rays_mask, starts, t0, t1, ri, cnt) -> None:
    G, bins = a["bb"].shape[0], a["bins"]
    check(lib().cnc_traverse_grids(ptr(a["o"]), ptr(a["d"]), ptr(rays_mask), a["o"].shape[0], G, bins.shape[-3], bins.shape[-2],
                                   bins.shape[-1], ptr(a["bins_u8"]), ptr(a["bb"]), ptr(a["hits"]), ptr(a["ts"]), ptr(a["ti"]),
                                   ptr(a["near"]), ptr(a["far"]), a["step"], a["cone"], -1, ptr(starts), ptr(cnt), ptr(t0), ptr(t1),
                                   ptr(ri), None, stream()))


class OccGridEstimator(torch.nn.Module):
    """Occupancy-grid transmittance estimator.  nerfacc/estimators/occ_grid.py:19-424"""

    DIM: int = 3

    def __init__(self, roi_aabb: Union[List[int], Tensor], resolution: Union[int, List[int], Tensor] = 128, levels: int = 1, **kwargs):
        super().__init__()
        if "contraction_type" in kwargs:
            raise ValueError("`contraction_type` is not supported anymore for nerfacc >= 0.4.0.")
        if isinstance(resolution, int):
            resolution = [resolution] * self.DIM
        if isinstance(resolution, (list, tuple)):
            resolution = torch.tensor(resolution, dtype=torch.int32)
        assert isinstance(resolution, Tensor) and resolution.shape[0] == self.DIM
        if isinstance(roi_aabb, (list, tuple)):
            roi_aabb = torch.tensor(roi_aabb, dtype=torch.float32)
        assert isinstance(roi_aabb, Tensor) and roi_aabb.shape[0] == self.DIM * 2
        aabbs = torch.stack([_enlarge_aabb(roi_aabb, 2 ** i) for i in range(levels)], dim=0)
        self.cells_per_lvl = int(resolution.prod().item())
        self.levels = levels
        self.register_buffer("resolution", resolution)
        self.register_buffer("aabbs", aabbs)
        self.register_buffer("occs", torch.zeros(self.levels * self.cells_per_lvl))
        self.register_buffer("binaries", torch.zeros([levels] + resolution.tolist(), dtype=torch.bool))
        coords = torch.stack(torch.meshgrid([torch.arange(int(r)) for r in resolution.tolist()], indexing="ij"), dim=-1).long()
        self.register_buffer("grid_coords", coords.reshape(self.cells_per_lvl, self.DIM), persistent=False)
        self.register_buffer("grid_indices", torch.arange(self.cells_per_lvl), persistent=False)

    @property
    def device(self):
        return self.occs.device

    @torch.no_grad()
    def sampling(self, rays_o: Tensor, rays_d: Tensor, sigma_fn: Optional[Callable] = None, alpha_fn: Optional[Callable] = None,
                 near_plane: float = 0.0, far_plane: float = 1e10, t_min: Optional[Tensor] = None, t_max: Optional[Tensor] = None,
                 render_step_size: float = 1e-3, early_stop_eps: float = 1e-4, alpha_thre: float = 0.0, stratified: bool = False,
                 cone_angle: float = 0.0, premarched: Optional["Premarch"] = None) -> Tuple[Tensor, Tensor, Tensor]:
        """(ray_indices, t_starts, t_ends) of the samples that survive occupancy + visibility skipping.  :88-239
        `premarched` (not in nerfacc): the occupancy march of exactly these rays, issued ahead of time (`Premarch`)."""
        if premarched is None:
            near_planes, far_planes = self._planes(rays_o, near_plane, far_plane, t_min, t_max)
        if premarched is not None:
            t_starts, t_ends, ray_indices, packed_info = premarched.take(self, rays_o, rays_d, near_plane, far_plane, t_min, t_max,
                                                                         render_step_size, stratified, cone_angle)
        else:
            if stratified:
                near_planes += torch.rand_like(near_planes) * render_step_size
            intervals, samples, _ = traverse_grids(rays_o, rays_d, self.binaries, self.aabbs, near_planes=near_planes,
                                                   far_planes=far_planes, step_size=render_step_size, cone_angle=cone_angle)
            t_starts, t_ends = intervals.t_starts, intervals.t_ends      # == vals[is_left], vals[is_right] (occ_grid.py:188-189)
            ray_indices, packed_info = samples.ray_indices, samples.packed_info
        if (alpha_thre > 0.0 or early_stop_eps > 0.0) and (sigma_fn is not None or alpha_fn is not None):
            if alpha_thre > 0.0:   # (min(alpha_thre <= 0, mean) cannot make the test below true: no device read needed)
                alpha_thre = min(alpha_thre, self.occs.mean().item())
            if sigma_fn is None:
                raise NotImplementedError("alpha_fn is not on the CNC path; pass sigma_fn")
            sigmas = sigma_fn(t_starts, t_ends, ray_indices) if t_starts.shape[0] != 0 else torch.empty((0,), device=t_starts.device)
            assert sigmas.shape == t_starts.shape, "sigmas must have shape of (N,)! Got {}".format(sigmas.shape)
            masks = render_visibility_from_density(t_starts=t_starts, t_ends=t_ends, sigmas=sigmas, packed_info=packed_info,
                                                   early_stop_eps=early_stop_eps, alpha_thre=alpha_thre)
            ray_indices, t_starts, t_ends = _compact(masks, t_starts, t_ends, ray_indices, packed_info)
        return ray_indices, t_starts, t_ends

    @staticmethod
    def _planes(rays_o, near_plane, far_plane, t_min, t_max):
        near_planes = torch.full_like(rays_o[..., 0], fill_value=near_plane)
        far_planes = torch.full_like(rays_o[..., 0], fill_value=far_plane)
        if t_min is not None:
            near_planes = torch.clamp(near_planes, min=t_min)
        if t_max is not None:
            far_planes = torch.clamp(far_planes, max=t_max)
        return near_planes, far_planes

    @torch.no_grad()
    def update_every_n_steps(self, step: int, occ_eval_fn: Callable, occ_thre: float = 1e-2, ema_decay: float = 0.95,
                             warmup_steps: int = 256, n: int = 16) -> None:
        """:242-277"""
        if not self.training:
            raise RuntimeError("You should only call this function only during training. "
                               "Please call _update() directly if you want to update the field during inference.")
        if step % n == 0 and self.training:
            self._update(step=step, occ_eval_fn=occ_eval_fn, occ_thre=occ_thre, ema_decay=ema_decay, warmup_steps=warmup_steps)

    @torch.no_grad()
    def _get_all_cells(self) -> List[Tensor]:
        """:349-361"""
        return [self.grid_indices[self.occs[lvl * self.cells_per_lvl + self.grid_indices] >= 0.0] for lvl in range(self.levels)]

    @torch.no_grad()
    def _sample_uniform_and_occupied_cells(self, n: int) -> List[Tensor]:
        """:363-385"""
        out = []
        for lvl in range(self.levels):
            uniform = torch.randint(self.cells_per_lvl, (n,), device=self.device)
            uniform = uniform[self.occs[lvl * self.cells_per_lvl + uniform] >= 0.0]
            occupied = torch.nonzero(self.binaries[lvl].flatten())[:, 0]
            if n < len(occupied):
                occupied = occupied[torch.randint(len(occupied), (n,), device=self.device)]
            out.append(torch.cat([uniform, occupied], dim=0))
        return out

    @torch.no_grad()
    def _update(self, step: int, occ_eval_fn: Callable, occ_thre: float = 0.01, ema_decay: float = 0.95, warmup_steps: int = 256) -> None:
        """EMA update of the occupancy field and re-binarisation.  :387-424"""
        lvl_indices = self._get_all_cells() if step < warmup_steps else self._sample_uniform_and_occupied_cells(self.cells_per_lvl // 4)
        for lvl, indices in enumerate(lvl_indices):
            grid_coords = self.grid_coords[indices]
            x = (grid_coords + torch.rand_like(grid_coords, dtype=torch.float32)) / self.resolution
            x = self.aabbs[lvl, :3] + x * (self.aabbs[lvl, 3:] - self.aabbs[lvl, :3])
            occ = occ_eval_fn(x).squeeze(-1)
            cell_ids = lvl * self.cells_per_lvl + indices
            self.occs[cell_ids] = torch.maximum(self.occs[cell_ids] * ema_decay, occ)
        thre = torch.clamp(self.occs[self.occs >= 0].mean(), max=occ_thre)
        self.binaries = (self.occs > thre).view(self.binaries.shape)
