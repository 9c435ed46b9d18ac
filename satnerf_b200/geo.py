"""GPU geometry either side of the render path, behind the reference's dataset helpers:

    get_rays(cols, rows, rpc, min_alt, max_alt)              datasets/satellite.py:18-65
    SatelliteGeometry.rays_for_image(...)                    :185-205 (get_rays + normalize_rays + get_sun_dirs + hstack)
    SatelliteGeometry.get_latlonalt_from_nerf_prediction()   :246-274
    SatelliteGeometry.get_dsm_from_nerf_prediction()         :276-338 (plyflatten raster; no GeoTIFF writing here)

`rpc` is anything with rpcm.RPCModel's attributes (row_num, row_den, col_num, col_den, *_offset, *_scale) or a dict of
them.  Everything runs on the current CUDA device through the C ABI; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import capi

_RPC_SCALARS = ("row_offset", "row_scale", "col_offset", "col_scale", "lat_offset", "lat_scale", "lon_offset", "lon_scale",
                "alt_offset", "alt_scale")


def rpc_struct(rpc) -> capi.RpcModel:
    get = (lambda k: rpc[k]) if isinstance(rpc, dict) else (lambda k: getattr(rpc, k))
    m = capi.RpcModel()
    for k in ("row_num", "row_den", "col_num", "col_den"):
        v = [float(x) for x in get(k)]
        if len(v) != 20:
            raise ValueError(f"rpc.{k} must hold 20 coefficients")
        setattr(m, k, (C.c_double * 20)(*v))
    for k in _RPC_SCALARS:
        setattr(m, k, float(get(k)))
    return m


def rescale_rpc(rpc, alpha):
    """sat_utils.py:44-57: the RPC of the image resampled by `alpha` (row / col scales and offsets times alpha)."""
    get = (lambda k: rpc[k]) if isinstance(rpc, dict) else (lambda k: getattr(rpc, k))
    out = {k: (list(get(k)) if k.endswith(("num", "den")) else float(get(k))) for k in ("row_num", "row_den", "col_num", "col_den") + _RPC_SCALARS}
    for k in ("row_scale", "col_scale", "row_offset", "col_offset"):
        out[k] *= float(alpha)
    return out


def _dev(device):
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("satnerf_b200.geo runs on CUDA devices only (no CPU fallback)")
    return dev


def _rpc_rays(rpc, cols, rows, width, n, min_alt, max_alt, center, rng, sun_d, device):
    dev = _dev(device)
    ray_cols = 11 if sun_d is not None else 8
    rays = torch.empty(n, ray_cols, device=dev, dtype=torch.float32)
    iters = torch.zeros(1, device=dev, dtype=torch.int32)
    m = rpc_struct(rpc)
    c3 = (C.c_double * 3)(*[float(x) for x in center]) if center is not None else None
    s3 = (C.c_float * 3)(*[float(x) for x in sun_d]) if sun_d is not None else None
    cp = rp = None
    if cols is not None:
        cols_d = torch.as_tensor(np.asarray(cols, dtype=np.float64)).to(dev).contiguous()
        rows_d = torch.as_tensor(np.asarray(rows, dtype=np.float64)).to(dev).contiguous()
        cp, rp = C.c_void_p(cols_d.data_ptr()), C.c_void_p(rows_d.data_ptr())
    with torch.cuda.device(dev):
        capi._check(capi.lib().snb_rpc_rays(C.byref(m), cp, rp, int(width), n, float(min_alt), float(max_alt), c3, float(rng if rng else 0.0), s3,
                                            C.c_void_p(rays.data_ptr()), ray_cols, C.c_void_p(iters.data_ptr()), capi._stream(dev)), "snb_rpc_rays")
    if int(iters.item()) > 100:
        raise RuntimeError("Max localization iterations (100) exceeded")      # rpcm.MaxLocalizationIterationsError
    return rays


def get_rays(cols, rows, rpc, min_alt, max_alt, device="cuda"):
    """datasets/satellite.py:18-65: (h*w, 8) float32 rays [origin (ECEF), unit direction, near = 0, far] for the pixels (cols, rows)."""
    cols = np.asarray(cols).reshape(-1)
    return _rpc_rays(rpc, cols, np.asarray(rows).reshape(-1), 1, cols.shape[0], min_alt, max_alt, None, None, None, device)


def get_sun_dirs(sun_elevation_deg, sun_azimuth_deg):
    """datasets/satellite.py:229-244 (one row of it)."""
    el, az = np.radians(sun_elevation_deg), np.radians(sun_azimuth_deg)
    return np.array([np.sin(az) * np.cos(el), np.cos(az) * np.cos(el), np.sin(el)])


def utm_zone_number(lat, lon):
    """utm.latlon_to_zone_number (utm 0.7.0), incl. the Norway / Svalbard exceptions."""
    if 56 <= lat < 64 and 3 <= lon < 12:
        return 32
    if 72 <= lat <= 84 and lon >= 0:
        if lon < 9:
            return 31
        if lon < 21:
            return 33
        if lon < 33:
            return 35
        if lon < 42:
            return 37
    return int((lon + 180) / 6) % 60 + 1


class SatelliteGeometry:
    """The geometric state of datasets/satellite.py::SatelliteDataset (scene `center`, `range` from scene.loc, :123-135)."""

    def __init__(self, center, range_, device="cuda"):
        self.center = [float(c) for c in center]
        self.range = float(range_)
        self.device = _dev(device)

    def rays_for_image(self, rpc, width, height, min_alt, max_alt, sun_elevation_deg=None, sun_azimuth_deg=None):
        """load_data's per-image ray block (:185-205, :215): all pixels of the w x h grid, normalised, sun direction appended."""
        sun = get_sun_dirs(sun_elevation_deg, sun_azimuth_deg) if sun_elevation_deg is not None else None
        return _rpc_rays(rpc, None, None, int(width), int(width) * int(height), min_alt, max_alt, self.center, self.range, sun, self.device)

    def _cloud(self, rays, depth, zone):
        rays = rays.detach().to(self.device, torch.float32).contiguous()
        depth = depth.detach().to(self.device, torch.float32).reshape(-1).contiguous()
        n = rays.shape[0]
        if depth.numel() != n:
            raise ValueError("one depth per ray expected")
        cloud = torch.empty(n, 3, device=self.device, dtype=torch.float64)
        latlon = torch.empty(n, 2, device=self.device, dtype=torch.float64)
        c3 = (C.c_double * 3)(*self.center)
        with torch.cuda.device(self.device):
            capi._check(capi.lib().snb_dsm_points(C.c_void_p(rays.data_ptr()), rays.shape[1], C.c_void_p(depth.data_ptr()), n, c3, self.range, int(zone),
                                                  C.c_void_p(cloud.data_ptr()), C.c_void_p(latlon.data_ptr()), capi._stream(self.device)), "snb_dsm_points")
        return cloud, latlon

    def get_latlonalt_from_nerf_prediction(self, rays, depth):
        """:246-274 -> (lats, lons, alts) float64 device tensors."""
        cloud, latlon = self._cloud(rays, depth, 0)
        return latlon[:, 0], latlon[:, 1], cloud[:, 2]

    def get_dsm_from_nerf_prediction(self, rays, depth, roi=None, resolution=0.5, radius=1):
        """:276-338: point cloud in UTM (zone of the first point, :290 + sat_utils.utm_from_latlon) and its plyflatten raster.
        roi = (xoff, yoff, size, resolution) of a ground-truth region (the numbers of the reference's roi_txt, :294-299).
        Returns (dsm (ysize, xsize) float32 device tensor, (xoff, yoff, resolution))."""
        _, latlon0 = self._cloud(rays[:1], depth.reshape(-1)[:1], 0)
        lat0, lon0 = [float(v) for v in latlon0[0].tolist()]
        cloud, _ = self._cloud(rays, depth, utm_zone_number(lat0, lon0))
        if roi is not None:
            xoff, yoff, xsize, resolution = float(roi[0]), float(roi[1]), int(roi[2]), float(roi[3])
            ysize = xsize
            yoff += ysize * resolution
        else:
            xmin, xmax = float(cloud[:, 0].min()), float(cloud[:, 0].max())
            ymin, ymax = float(cloud[:, 1].min()), float(cloud[:, 1].max())
            xoff = math.floor(xmin / resolution) * resolution
            xsize = int(1 + math.floor((xmax - xoff) / resolution))
            yoff = math.ceil(ymax / resolution) * resolution
            ysize = int(1 - math.floor((ymin - yoff) / resolution))
        dsm = torch.empty(ysize, xsize, device=self.device, dtype=torch.float32)
        ws = torch.empty(xsize * ysize * 12 + 512, device=self.device, dtype=torch.uint8)
        with torch.cuda.device(self.device):
            capi._check(capi.lib().snb_dsm_rasterize(C.c_void_p(cloud.data_ptr()), cloud.shape[0], xoff, yoff, resolution, xsize, ysize, int(radius),
                                                     C.c_void_p(dsm.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(), capi._stream(self.device)),
                        "snb_dsm_rasterize")
        return dsm, (xoff, yoff, resolution)
