"""Host-side mirror of the render glue: examples/utils.py:83-216 `render_image_with_occgrid` and
examples/utils.py:316-489 `render_image_with_occgrid_test` (first version: the reference's wavefront
loop on the fused kernels).

Same signature and return values; `Rays` is the reference's namedtuple (origins, viewdirs).  During training the
whole batch is one chunk, at test time `test_chunk_size` rays per chunk (utils.py:169-174).  The closures go through
the radiance field, i.e. under `torch.no_grad()` (the `sigma_fn` visibility pass, evaluation) they hit the fused
`cnc_field_fwd` kernel, with autograd they take the differentiable path.
"""
from __future__ import annotations

import collections
from typing import Optional

import torch

import ctypes

from ._lib import check, lib, ptr, stream
from .nerfacc import (OccGridEstimator, accumulate_along_rays_, ray_aabb_intersect, render_fused, render_weight_from_density,
                      rendering, traverse_grids)

Rays = collections.namedtuple("Rays", ("origins", "viewdirs"))


def namedtuple_map(fn, tup):
    return type(tup)(*(None if x is None else fn(x) for x in tup))


def sample_points(origins: torch.Tensor, viewdirs: torch.Tensor, ray_indices: torch.Tensor, t_starts: torch.Tensor,
                  t_ends: torch.Tensor):
    """(origins[ri] + viewdirs[ri] * (t_starts + t_ends)[:, None] / 2, viewdirs[ri]) -- the query points of
    examples/utils.py:250-262 -- in one pass (csrc/march_render.cu, `cnc_sample_points`); no gradient"""
    n = t_starts.shape[0]
    pos = torch.empty(n, 3, device=origins.device)
    dirs = torch.empty(n, 3, device=origins.device)
    check(lib().cnc_sample_points(ptr(origins.contiguous()), ptr(viewdirs.contiguous()), ptr(ray_indices.contiguous()),
                                  ptr(t_starts.contiguous()), ptr(t_ends.contiguous()), n, ptr(pos), ptr(dirs), stream()))
    return pos, dirs


def render_image_with_occgrid(radiance_field: torch.nn.Module, estimator: OccGridEstimator, rays: Rays, near_plane: float = 0.0,
                              far_plane: float = 1e10, render_step_size: float = 1e-3, render_bkgd: Optional[torch.Tensor] = None,
                              cone_angle: float = 0.0, alpha_thre: float = 0.0, test_chunk_size: int = 8192, timestamps=None,
                              return_extra=False, tmp=None, premarch=None):
    """(colors, opacities, depths, n_rendering_samples[, extras]); `premarch` (not in the reference): a nerfacc.Premarch that
    may hold the occupancy march of exactly these rays, issued ahead of time"""
    if timestamps is not None:
        raise NotImplementedError("timestamps belong to the dynamic-scene fields, which the CNC scripts do not use")
    rays_shape = rays.origins.shape
    if len(rays_shape) == 3:
        num_rays = rays_shape[0] * rays_shape[1]
        rays = namedtuple_map(lambda r: r.reshape([num_rays] + list(r.shape[2:])), rays)
    else:
        num_rays = rays_shape[0]

    def positions_of(chunk_rays, t_starts, t_ends, ray_indices):
        o, d = chunk_rays.origins, chunk_rays.viewdirs
        if o.is_cuda and o.dtype == torch.float32 and not (o.requires_grad or d.requires_grad or t_starts.requires_grad):
            return sample_points(o, d, ray_indices, t_starts, t_ends)      # the two lines below as one kernel
        t_dirs = d[ray_indices]
        return o[ray_indices] + t_dirs * (t_starts + t_ends)[:, None] / 2.0, t_dirs

    results, extras = [], None
    chunk = torch.iinfo(torch.int32).max if radiance_field.training else test_chunk_size
    for i in range(0, num_rays, chunk):
        chunk_rays = namedtuple_map(lambda r: r[i:i + chunk], rays)

        def sigma_fn(t_starts, t_ends, ray_indices):
            positions, _ = positions_of(chunk_rays, t_starts, t_ends, ray_indices)
            return radiance_field.query_density(positions).squeeze(-1)

        def rgb_sigma_fn(t_starts, t_ends, ray_indices):
            positions, t_dirs = positions_of(chunk_rays, t_starts, t_ends, ray_indices)
            rgbs, sigmas = radiance_field(positions, t_dirs)
            return rgbs, sigmas.squeeze(-1), positions

        ahead = premarch if (premarch is not None and premarch.matches(estimator, chunk_rays.origins, chunk_rays.viewdirs, near_plane,
                                                                       far_plane, render_step_size, radiance_field.training,
                                                                       cone_angle)) else None
        ray_indices, t_starts, t_ends = estimator.sampling(chunk_rays.origins, chunk_rays.viewdirs, sigma_fn=sigma_fn,
                                                           near_plane=near_plane, far_plane=far_plane,
                                                           render_step_size=render_step_size,
                                                           stratified=radiance_field.training, cone_angle=cone_angle,
                                                           alpha_thre=alpha_thre, premarched=ahead)
        rgb, opacity, depth, extras = rendering(t_starts, t_ends, ray_indices, n_rays=chunk_rays.origins.shape[0],
                                                rgb_sigma_fn=rgb_sigma_fn, render_bkgd=render_bkgd)
        results.append([rgb, opacity, depth, len(t_starts)])
    colors, opacities, depths, n_samples = [torch.cat(r, dim=0) if isinstance(r[0], torch.Tensor) else r for r in zip(*results)]
    out = (colors.view((*rays_shape[:-1], -1)), opacities.view((*rays_shape[:-1], -1)), depths.view((*rays_shape[:-1], -1)),
           sum(n_samples))
    return out + (extras,) if return_extra else out


@torch.no_grad()
def render_image_with_occgrid_test(max_samples: int, radiance_field: torch.nn.Module, estimator: OccGridEstimator, rays: Rays,
                                   near_plane: float = 0.0, far_plane: float = 1e10, render_step_size: float = 1e-3,
                                   render_bkgd: Optional[torch.Tensor] = None, cone_angle: float = 0.0, alpha_thre: float = 0.0,
                                   early_stop_eps: float = 1e-4, timestamps=None, samples_per_round: Optional[int] = None,
                                   device_loop: Optional[bool] = None, rounds_per_check: int = 16):
    """Test-time renderer, examples/utils.py:316-489: all rays of the image advance together, a few samples per ray and
    round (more as rays die), a ray leaves the wavefront once its opacity exceeds 1 - early_stop_eps or it reaches the far
    plane.  -> (rgb, opacity, depth, total_samples).  Every round is march (`cnc_traverse_grids` with a step limit, the
    ray mask and the previous termination planes) -> fused field forward -> `cnc_render_from_density` with the rays'
    running transmittance as prefix -> index_add of colour / opacity / depth.

    `samples_per_round` (not in the reference): a fixed number of samples per live ray and round instead of the reference's
    `max(min(num_rays // n_alive, 64), min_samples)`.  The reference's schedule starts at ONE sample per ray and round, i.e.
    hundreds of rounds of a few launches and two host syncs each; a round is cheap on this hardware only when it is large.
    The image is the same within `early_stop_eps` (a ray is retired at the end of the round in which it crosses the
    threshold, so it may take up to samples_per_round - 1 samples more); `total_samples` grows accordingly.

    `device_loop` (default: on whenever the field is the fused product field on CUDA and the reference schedule is asked
    for): the whole loop runs on the device -- `_render_test_device` below -- with the reference's round schedule, sample
    placement and stopping rules; the host only looks at a `done` flag every `rounds_per_check` rounds."""
    if device_loop is None:
        device_loop = (samples_per_round is None and timestamps is None and rays.origins.is_cuda
                       and getattr(radiance_field, "fused_available", lambda: False)() and not torch.is_grad_enabled()
                       and getattr(radiance_field, "fused", True))
    if device_loop:
        return _render_test_device(max_samples, radiance_field, estimator, rays, near_plane, far_plane, render_step_size, render_bkgd,
                                   cone_angle, alpha_thre, early_stop_eps, rounds_per_check)
    if timestamps is not None:
        raise NotImplementedError("timestamps belong to the dynamic-scene fields, which the CNC scripts do not use")
    rays_shape = rays.origins.shape
    if len(rays_shape) == 3:
        num_rays = rays_shape[0] * rays_shape[1]
        rays = namedtuple_map(lambda r: r.reshape([num_rays] + list(r.shape[2:])), rays)
    else:
        num_rays = rays_shape[0]
    device = rays.origins.device
    opacity = torch.zeros(num_rays, 1, device=device)
    depth = torch.zeros(num_rays, 1, device=device)
    rgb = torch.zeros(num_rays, 3, device=device)
    ray_mask = torch.ones(num_rays, device=device, dtype=torch.bool)
    min_samples = 1 if cone_angle == 0 else 4   # 1 for synthetic scenes, 4 for real scenes
    iter_samples = total_samples = 0
    rays_o, rays_d = rays.origins.contiguous(), rays.viewdirs.contiguous()
    near_planes = torch.full_like(rays_o[..., 0], fill_value=near_plane)
    far_planes = torch.full_like(rays_o[..., 0], fill_value=far_plane)
    t_mins, t_maxs, hits = ray_aabb_intersect(rays_o, rays_d, estimator.aabbs)
    n_grids = estimator.binaries.size(0)
    if n_grids > 1:
        t_sorted, t_indices = torch.sort(torch.cat([t_mins, t_maxs], -1), -1)
    else:
        t_sorted = torch.cat([t_mins, t_maxs], -1)
        t_indices = torch.arange(0, n_grids * 2, device=device, dtype=torch.int64).expand(num_rays, n_grids * 2)
    opc_thre = 1 - early_stop_eps
    while iter_samples < max_samples:
        n_alive = int(ray_mask.sum())
        if n_alive == 0:
            break
        if samples_per_round is None:
            n_samples = max(min(num_rays // n_alive, 64), min_samples)   # the number of samples to add on each ray
        else:
            n_samples = max(int(samples_per_round), min_samples)
        iter_samples += n_samples
        intervals, samples, termination_planes = traverse_grids(
            rays_o, rays_d, estimator.binaries, estimator.aabbs, near_planes, far_planes, render_step_size, cone_angle,
            n_samples, True, ray_mask, t_sorted, t_indices, hits)
        t_starts, t_ends = intervals.t_starts, intervals.t_ends   # == vals[is_left], vals[is_right] (utils.py:431-432)
        ray_indices = samples.ray_indices                         # every returned sample is valid (exact-size march)
        packed_info = samples.packed_info
        if ray_indices.numel():
            t_dirs = rays_d[ray_indices]
            positions = rays_o[ray_indices] + t_dirs * (t_starts[:, None] + t_ends[:, None]) / 2.0
            rgbs, sigmas = radiance_field(positions, t_dirs)
            weights, _, alphas = render_weight_from_density(t_starts, t_ends, sigmas.squeeze(-1), ray_indices=ray_indices,
                                                            n_rays=num_rays, prefix_trans=1 - opacity[ray_indices].squeeze(-1))
            if alpha_thre > 0:
                vis = alphas >= alpha_thre
                ray_indices, rgbs, weights, t_starts, t_ends = ray_indices[vis], rgbs[vis], weights[vis], t_starts[vis], t_ends[vis]
            accumulate_along_rays_(weights, values=rgbs, ray_indices=ray_indices, outputs=rgb)
            accumulate_along_rays_(weights, values=None, ray_indices=ray_indices, outputs=opacity)
            accumulate_along_rays_(weights, values=(t_starts + t_ends)[..., None] / 2.0, ray_indices=ray_indices, outputs=depth)
        near_planes = torch.where(ray_mask, termination_planes, near_planes)   # dead rays keep their last plane
        # early stopping, and rays that have reached the far plane (fewer samples than asked for) leave the wavefront
        ray_mask = torch.logical_and(opacity.view(-1) <= opc_thre, packed_info[:, 1] == n_samples)
        total_samples += int(ray_indices.shape[0])
    if render_bkgd is not None:
        rgb = rgb + render_bkgd * (1.0 - opacity)
    depth = depth / opacity.clamp_min(torch.finfo(rgb.dtype).eps)
    return (rgb.view((*rays_shape[:-1], -1)), opacity.view((*rays_shape[:-1], -1)), depth.view((*rays_shape[:-1], -1)),
            total_samples)


def _render_test_device(max_samples, field, estimator, rays, near_plane, far_plane, render_step_size, render_bkgd, cone_angle,
                        alpha_thre, early_stop_eps, rounds_per_check):
    """examples/utils.py:316-489 with the per-round decisions moved to the device (csrc/march_render.cu, wf_* kernels):

        cnc_wavefront_begin      live rays -> samples per ray of the round, iter_samples, stop conditions      (:395-403)
        cnc_wavefront_march      traverse_grids(n_samples, over_allocate, ray_mask, near_planes) -> packed samples, their
                                 positions and directions in fixed-capacity buffers (n_live * n <= n_rays)       (:405-432)
        cnc_field_fwd_n          the fused field kernel; its sample count is read from the device               (:434)
        cnc_wavefront_composite  weights with the running transmittance, 3 accumulations, near planes, ray mask (:436-478)

    No `.item()`, no boolean-mask compaction, no allocation inside the loop: rounds are queued `rounds_per_check` at a time
    and the host reads one `done` word between batches (kernels of rounds after `done` return immediately)."""
    rays_shape = rays.origins.shape
    if len(rays_shape) == 3:
        num_rays = rays_shape[0] * rays_shape[1]
        rays = namedtuple_map(lambda r: r.reshape([num_rays] + list(r.shape[2:])), rays)
    else:
        num_rays = rays_shape[0]
    dev = rays.origins.device
    rays_o, rays_d = rays.origins.contiguous().float(), rays.viewdirs.contiguous().float()
    min_samples = 1 if cone_angle == 0 else 4
    cap = num_rays * min_samples
    f32 = dict(device=dev, dtype=torch.float32)
    rgb, opacity, depth = torch.zeros(num_rays, 3, **f32), torch.zeros(num_rays, **f32), torch.zeros(num_rays, **f32)
    ray_mask = torch.ones(num_rays, dtype=torch.uint8, device=dev)
    near_planes = torch.full((num_rays,), float(near_plane), **f32)
    far_planes = torch.full((num_rays,), float(far_plane), **f32)
    t_mins, t_maxs, hits = ray_aabb_intersect(rays_o, rays_d, estimator.aabbs)
    n_grids = estimator.binaries.size(0)
    if n_grids > 1:
        t_sorted, t_indices = torch.sort(torch.cat([t_mins, t_maxs], -1), -1)
    else:
        t_sorted = torch.cat([t_mins, t_maxs], -1)
        t_indices = torch.arange(0, n_grids * 2, device=dev, dtype=torch.int64).expand(num_rays, n_grids * 2)
    t_sorted, t_indices = t_sorted.contiguous().float(), t_indices.contiguous()
    hits_u8 = hits.contiguous().to(torch.uint8)
    bins = estimator.binaries.contiguous()
    bins_u8 = bins.view(torch.uint8) if bins.dtype == torch.bool else bins.to(torch.uint8)
    aabbs = estimator.aabbs.contiguous().float()
    state = torch.zeros(16, dtype=torch.int32, device=dev)
    state[5] = num_rays
    ray_base = torch.empty(num_rays, dtype=torch.int32, device=dev)
    ray_cnt = torch.empty(num_rays, dtype=torch.int32, device=dev)
    ray_term = torch.empty(num_rays, **f32)
    t0, t1 = torch.empty(cap, **f32), torch.empty(cap, **f32)
    pos, dirs = torch.empty(cap, 3, **f32), torch.empty(cap, 3, **f32)
    sigma, rgbs = torch.empty(cap, **f32), torch.empty(cap, 3, **f32)
    mb = field.mlp_base
    encs = (mb.encoding_xyz, mb.encoding_xy, mb.encoding_xz, mb.encoding_yz)
    bits = [e.sign_bits() for e in encs]
    blob = field._fused_blob()
    aabb_c = ctypes.addressof(field._aabb_c())
    L, st = lib(), stream()
    n_cnt = state[4:].data_ptr()
    done_host = torch.zeros(16, dtype=torch.int32).pin_memory()
    rounds = 0
    while rounds < max_samples:          # every round adds >= 1 to iter_samples
        for _ in range(rounds_per_check):
            check(L.cnc_wavefront_begin(ptr(state), num_rays, min_samples, int(max_samples), st))
            check(L.cnc_wavefront_march(ptr(rays_o), ptr(rays_d), num_rays, n_grids, bins.shape[-3], bins.shape[-2], bins.shape[-1],
                                        ptr(bins_u8), ptr(aabbs), ptr(hits_u8), ptr(t_sorted), ptr(t_indices), ptr(far_planes),
                                        float(render_step_size), float(cone_angle), ptr(state), ptr(ray_mask), ptr(near_planes), cap,
                                        ptr(ray_base), ptr(ray_cnt), ptr(ray_term), ptr(t0), ptr(t1), ptr(pos), ptr(dirs), st))
            check(L.cnc_field_fwd_n(ptr(pos), ptr(dirs), aabb_c, *[ptr(b) for b in bits], ptr(mb.encoding_xyz.offsets_list),
                                    ptr(mb.encoding_xyz.resolutions_list), ptr(mb.encoding_xy.offsets_list),
                                    ptr(mb.encoding_xy.resolutions_list), ptr(blob), ptr(sigma), ptr(rgbs), n_cnt, cap, st))
            check(L.cnc_wavefront_composite(ptr(state), ptr(ray_mask), ptr(near_planes), ptr(ray_base), ptr(ray_cnt), ptr(ray_term),
                                            ptr(t0), ptr(t1), ptr(sigma), ptr(rgbs), ptr(rgb), ptr(opacity), ptr(depth), num_rays, cap,
                                            float(alpha_thre), float(1 - early_stop_eps), st))
        rounds += rounds_per_check
        done_host.copy_(state, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        # done: the flag, or nothing left alive / the sample budget used (the next begin would set it)
        if int(done_host[3]) or int(done_host[5]) == 0 or int(done_host[2]) >= max_samples:
            break
    total_samples = int(done_host[8:10].view(torch.int64)[0])
    opacity = opacity[:, None]
    if render_bkgd is not None:
        rgb = rgb + render_bkgd * (1.0 - opacity)
    depth = depth[:, None] / opacity.clamp_min(torch.finfo(rgb.dtype).eps)
    return (rgb.view((*rays_shape[:-1], -1)), opacity.view((*rays_shape[:-1], -1)), depth.view((*rays_shape[:-1], -1)),
            total_samples)
