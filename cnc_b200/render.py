"""Host-side mirror of the render glue: examples/utils.py:83-216 `render_image_with_occgrid`.

Same signature and return values; `Rays` is the reference's namedtuple (origins, viewdirs).  During training the
whole batch is one chunk, at test time `test_chunk_size` rays per chunk (utils.py:169-174).  The closures go through
the radiance field, i.e. under `torch.no_grad()` (the `sigma_fn` visibility pass, evaluation) they hit the fused
`cnc_field_fwd` kernel, with autograd they take the differentiable path.
"""
from __future__ import annotations

import collections
from typing import Optional

import torch

from .nerfacc import OccGridEstimator, render_fused, rendering

Rays = collections.namedtuple("Rays", ("origins", "viewdirs"))


def namedtuple_map(fn, tup):
    return type(tup)(*(None if x is None else fn(x) for x in tup))


def render_image_with_occgrid(radiance_field: torch.nn.Module, estimator: OccGridEstimator, rays: Rays, near_plane: float = 0.0,
                              far_plane: float = 1e10, render_step_size: float = 1e-3, render_bkgd: Optional[torch.Tensor] = None,
                              cone_angle: float = 0.0, alpha_thre: float = 0.0, test_chunk_size: int = 8192, timestamps=None,
                              return_extra=False, tmp=None):
    """(colors, opacities, depths, n_rendering_samples[, extras])"""
    if timestamps is not None:
        raise NotImplementedError("timestamps belong to the dynamic-scene fields, which the CNC scripts do not use")
    rays_shape = rays.origins.shape
    if len(rays_shape) == 3:
        num_rays = rays_shape[0] * rays_shape[1]
        rays = namedtuple_map(lambda r: r.reshape([num_rays] + list(r.shape[2:])), rays)
    else:
        num_rays = rays_shape[0]

    def positions_of(chunk_rays, t_starts, t_ends, ray_indices):
        t_dirs = chunk_rays.viewdirs[ray_indices]
        return chunk_rays.origins[ray_indices] + t_dirs * (t_starts + t_ends)[:, None] / 2.0, t_dirs

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

        ray_indices, t_starts, t_ends = estimator.sampling(chunk_rays.origins, chunk_rays.viewdirs, sigma_fn=sigma_fn,
                                                           near_plane=near_plane, far_plane=far_plane,
                                                           render_step_size=render_step_size,
                                                           stratified=radiance_field.training, cone_angle=cone_angle,
                                                           alpha_thre=alpha_thre)
        rgb, opacity, depth, extras = rendering(t_starts, t_ends, ray_indices, n_rays=chunk_rays.origins.shape[0],
                                                rgb_sigma_fn=rgb_sigma_fn, render_bkgd=render_bkgd)
        results.append([rgb, opacity, depth, len(t_starts)])
    colors, opacities, depths, n_samples = [torch.cat(r, dim=0) if isinstance(r[0], torch.Tensor) else r for r in zip(*results)]
    out = (colors.view((*rays_shape[:-1], -1)), opacities.view((*rays_shape[:-1], -1)), depths.view((*rays_shape[:-1], -1)),
           sum(n_samples))
    return out + (extras,) if return_extra else out
