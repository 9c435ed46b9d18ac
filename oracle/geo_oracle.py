"""CPU restatement (numpy, float64) of the geometry either side of the render path -- TEST INFRASTRUCTURE ONLY: imported by
tests/ (and nothing in the product path).

Follows the reference where the reference holds the arithmetic:
    get_rays                 datasets/satellite.py:18-65          normalize_rays           :218-227
    latlon_to_ecef_custom    sat_utils.py:59-74                   ecef_to_latlon_custom    sat_utils.py:76-97
    get_latlonalt_from_nerf_prediction  datasets/satellite.py:246-274        DSM grid bounds  :300-306
and restates the published algorithms of the three dependencies that are NOT vendored in /root/reference and not installed
here (requirements.txt: `rpcm` unpinned, `pyproj==3.0.1`, `plyflatten==0.2.0`):
    rpcm.RPCModel.projection / localization_iterative (RPC00B rational polynomials; secant iteration, eps 2 then 0.1,
        stop when the squared normalised residual < 1e-18, at most 100 iterations);
    pyproj "+proj=utm" forward = PROJ's extended transverse Mercator = Krueger series of order 6 (Karney 2011), WGS84;
    plyflatten(cloud, xoff, yoff, resolution, xsize, ysize, radius, sigma=inf): mean of the altitudes of the points within
        `radius` cells of each cell (distance from the point to the cell centre, in cells), NaN where none.
PARITY UNPINNED for those three: no golden vector of rpcm / pyproj / plyflatten exists in the reference and none can be
generated here; they are anchored on properties instead (projection(localization(.)) round trip, UTM scale / false easting
/ inverse series, rasterised planes).
"""
import math

import numpy as np


def apply_poly(c, x, y, z):
    out = c[0]
    out = out + c[1] * y + c[2] * x + c[3] * z
    out = out + c[4] * y * x + c[5] * y * z + c[6] * x * z
    out = out + c[7] * y * y + c[8] * x * x + c[9] * z * z
    out = out + c[10] * x * y * z
    out = out + c[11] * y * y * y
    out = out + c[12] * y * x * x + c[13] * y * z * z + c[14] * y * y * x
    out = out + c[15] * x * x * x
    out = out + c[16] * x * z * z + c[17] * y * y * z + c[18] * x * x * z
    out = out + c[19] * z * z * z
    return out


def apply_rfm(num, den, x, y, z):
    return apply_poly(num, x, y, z) / apply_poly(den, x, y, z)


def projection(rpc, lon, lat, alt):
    nlon = (np.asarray(lon, float) - rpc["lon_offset"]) / rpc["lon_scale"]
    nlat = (np.asarray(lat, float) - rpc["lat_offset"]) / rpc["lat_scale"]
    nalt = (np.asarray(alt, float) - rpc["alt_offset"]) / rpc["alt_scale"]
    col = apply_rfm(rpc["col_num"], rpc["col_den"], nlat, nlon, nalt) * rpc["col_scale"] + rpc["col_offset"]
    row = apply_rfm(rpc["row_num"], rpc["row_den"], nlat, nlon, nalt) * rpc["row_scale"] + rpc["row_offset"]
    return col, row


def localization(rpc, col, row, alt):
    """Per-pixel stopping rule (rpcm iterates until ALL pixels have converged; the extra steps move a converged pixel by
    less than the tolerance)."""
    col, row, alt = [np.asarray(v, float).reshape(-1) for v in (col, row, alt)]
    ncol = (col - rpc["col_offset"]) / rpc["col_scale"]
    nrow = (row - rpc["row_offset"]) / rpc["row_scale"]
    nalt = (alt - rpc["alt_offset"]) / rpc["alt_scale"]
    lon = -np.ones_like(ncol); lat = -np.ones_like(ncol)
    eps = 2.0
    f = lambda la, lo: (apply_rfm(rpc["col_num"], rpc["col_den"], la, lo, nalt), apply_rfm(rpc["row_num"], rpc["row_den"], la, lo, nalt))
    x0, y0 = f(lat, lon); x1, y1 = f(lat, lon + eps); x2, y2 = f(lat + eps, lon)
    n = 0
    while True:
        todo = ~((x0 - ncol) ** 2 + (y0 - nrow) ** 2 < 1e-18)
        if not todo.any():
            break
        if n > 100:
            raise RuntimeError("Max localization iterations (100) exceeded")
        e1x, e1y, e2x, e2y, ux, uy = x1 - x0, y1 - y0, x2 - x0, y2 - y0, ncol - x0, nrow - y0
        a1 = (ux * e1x + uy * e1y) / (e1x * e1x + e1y * e1y)
        a2 = (ux * e2x + uy * e2y) / (e2x * e2x + e2y * e2y)
        lon = np.where(todo, lon + a1 * eps, lon); lat = np.where(todo, lat + a2 * eps, lat)
        eps = 0.1
        nx0, ny0 = f(lat, lon); nx1, ny1 = f(lat, lon + eps); nx2, ny2 = f(lat + eps, lon)
        x0, y0, x1, y1, x2, y2 = [np.where(todo, a, b) for a, b in ((nx0, x0), (ny0, y0), (nx1, x1), (ny1, y1), (nx2, x2), (ny2, y2))]
        n += 1
    return lon * rpc["lon_scale"] + rpc["lon_offset"], lat * rpc["lat_scale"] + rpc["lat_offset"]


def latlon_to_ecef_custom(lat, lon, alt):                       # sat_utils.py:59-74
    rad_lat = lat * (np.pi / 180.0); rad_lon = lon * (np.pi / 180.0)
    a = 6378137.0; finv = 298.257223563; f = 1 / finv; e2 = 1 - (1 - f) * (1 - f)
    v = a / np.sqrt(1 - e2 * np.sin(rad_lat) * np.sin(rad_lat))
    return (v + alt) * np.cos(rad_lat) * np.cos(rad_lon), (v + alt) * np.cos(rad_lat) * np.sin(rad_lon), (v * (1 - e2) + alt) * np.sin(rad_lat)


def ecef_to_latlon_custom(x, y, z):                             # sat_utils.py:76-97
    a = 6378137.0; e = 8.1819190842622e-2
    asq = a ** 2; esq = e ** 2
    b = np.sqrt(asq * (1 - esq)); bsq = b ** 2
    ep = np.sqrt((asq - bsq) / bsq)
    p = np.sqrt((x ** 2) + (y ** 2))
    th = np.arctan2(a * z, b * p)
    lon = np.arctan2(y, x)
    lat = np.arctan2((z + (ep ** 2) * b * (np.sin(th) ** 3)), (p - esq * a * (np.cos(th) ** 3)))
    N = a / (np.sqrt(1 - esq * (np.sin(lat) ** 2)))
    alt = p / np.cos(lat) - N
    return lat * 180 / np.pi, lon * 180 / np.pi, alt


def get_rays(cols, rows, rpc, min_alt, max_alt):                # datasets/satellite.py:18-65 -> (n, 8) float32
    cols = np.asarray(cols, float).reshape(-1); rows = np.asarray(rows, float).reshape(-1)
    lons, lats = localization(rpc, cols, rows, max_alt * np.ones_like(cols))
    xyz_near = np.vstack(latlon_to_ecef_custom(lats, lons, max_alt * np.ones_like(cols))).T
    lons, lats = localization(rpc, cols, rows, min_alt * np.ones_like(cols))
    xyz_far = np.vstack(latlon_to_ecef_custom(lats, lons, min_alt * np.ones_like(cols))).T
    d = xyz_far - xyz_near
    fars = np.linalg.norm(d, axis=1)
    rays_d = d / fars[:, np.newaxis]
    return np.hstack([xyz_near, rays_d, np.zeros_like(fars)[:, None], fars[:, None]]).astype(np.float32)


def normalize_rays(rays, center, rng):                          # datasets/satellite.py:218-227 (float32 arithmetic, as on the FloatTensor)
    rays = rays.astype(np.float32).copy()
    for k in range(3):
        rays[:, k] = (rays[:, k] - np.float32(center[k])) / np.float32(rng)
    rays[:, 6] = rays[:, 6] / np.float32(rng); rays[:, 7] = rays[:, 7] / np.float32(rng)
    return rays


def latlonalt_from_prediction(rays, depth, center, rng):        # datasets/satellite.py:246-274
    rays = rays.astype(np.float64); depth = depth.astype(np.float64).reshape(-1, 1)
    xyz = (rays[:, 0:3] + rays[:, 3:6] * depth) * rng
    xyz = xyz + np.asarray(center, float)[None]
    return ecef_to_latlon_custom(xyz[:, 0], xyz[:, 1], xyz[:, 2])


_A, _F, _K0 = 6378137.0, 1.0 / 298.257223563, 0.9996


def _kruger():
    n = _F / (2.0 - _F)
    A = _A / (1 + n) * (1 + n ** 2 / 4 + n ** 4 / 64 + n ** 6 / 256)
    al = [n / 2 - 2 * n ** 2 / 3 + 5 * n ** 3 / 16 + 41 * n ** 4 / 180 - 127 * n ** 5 / 288 + 7891 * n ** 6 / 37800,
          13 * n ** 2 / 48 - 3 * n ** 3 / 5 + 557 * n ** 4 / 1440 + 281 * n ** 5 / 630 - 1983433 * n ** 6 / 1935360,
          61 * n ** 3 / 240 - 103 * n ** 4 / 140 + 15061 * n ** 5 / 26880 + 167603 * n ** 6 / 181440,
          49561 * n ** 4 / 161280 - 179 * n ** 5 / 168 + 6601661 * n ** 6 / 7257600,
          34729 * n ** 5 / 80640 - 3418889 * n ** 6 / 1995840,
          212378941 * n ** 6 / 319334400]
    be = [n / 2 - 2 * n ** 2 / 3 + 37 * n ** 3 / 96 - n ** 4 / 360 - 81 * n ** 5 / 512 + 96199 * n ** 6 / 604800,
          n ** 2 / 48 + n ** 3 / 15 - 437 * n ** 4 / 1440 + 46 * n ** 5 / 105 - 1118711 * n ** 6 / 3870720,
          17 * n ** 3 / 480 - 37 * n ** 4 / 840 - 209 * n ** 5 / 4480 + 5569 * n ** 6 / 90720,
          4397 * n ** 4 / 161280 - 11 * n ** 5 / 504 - 830251 * n ** 6 / 7257600,
          4583 * n ** 5 / 161280 - 108847 * n ** 6 / 3991680,
          20648693 * n ** 6 / 638668800]
    return A, al, be


def utm_forward(lat, lon, zone):
    """(+proj=utm +zone=<zone>, no +south: the reference's proj string carries the zone letter where PROJ reads an integer)"""
    A, al, _ = _kruger()
    e = math.sqrt(_F * (2 - _F))
    phi = np.radians(np.asarray(lat, float)); lam = np.radians(np.asarray(lon, float) - (zone * 6 - 183))
    sp = np.sin(phi)
    t = np.sinh(np.arctanh(sp) - e * np.arctanh(e * sp))
    xi0 = np.arctan2(t, np.cos(lam)); eta0 = np.arcsinh(np.sin(lam) / np.sqrt(t * t + np.cos(lam) ** 2))
    xi, eta = xi0.copy(), eta0.copy()
    for j in range(1, 7):
        xi = xi + al[j - 1] * np.sin(2 * j * xi0) * np.cosh(2 * j * eta0)
        eta = eta + al[j - 1] * np.cos(2 * j * xi0) * np.sinh(2 * j * eta0)
    return 500000.0 + _K0 * A * eta, _K0 * A * xi


def utm_inverse(east, north, zone):
    """Inverse Krueger series (only the tests use it: forward / inverse round trip)."""
    A, _, be = _kruger()
    e = math.sqrt(_F * (2 - _F))
    xi = np.asarray(north, float) / (_K0 * A); eta = (np.asarray(east, float) - 500000.0) / (_K0 * A)
    xi0, eta0 = xi.copy(), eta.copy()
    for j in range(1, 7):
        xi0 = xi0 - be[j - 1] * np.sin(2 * j * xi) * np.cosh(2 * j * eta)
        eta0 = eta0 - be[j - 1] * np.cos(2 * j * xi) * np.sinh(2 * j * eta)
    chi = np.arcsin(np.sin(xi0) / np.cosh(eta0))
    lam = np.arctan2(np.sinh(eta0), np.cos(xi0))
    phi = chi.copy()
    for _ in range(8):                                           # conformal -> geodetic latitude (fixed point)
        phi = np.arcsin(np.tanh(np.arctanh(np.sin(chi)) + e * np.arctanh(e * np.sin(phi))))
    return np.degrees(phi), np.degrees(lam) + (zone * 6 - 183)


def dsm_bounds(cloud, resolution=0.5):                          # datasets/satellite.py:300-306
    xmin, xmax = cloud[:, 0].min(), cloud[:, 0].max()
    ymin, ymax = cloud[:, 1].min(), cloud[:, 1].max()
    xoff = np.floor(xmin / resolution) * resolution
    xsize = int(1 + np.floor((xmax - xoff) / resolution))
    yoff = np.ceil(ymax / resolution) * resolution
    ysize = int(1 - np.floor((ymin - yoff) / resolution))
    return xoff, yoff, xsize, ysize


def plyflatten(cloud, xoff, yoff, resolution, xsize, ysize, radius=1):
    s = np.zeros((ysize, xsize)); c = np.zeros((ysize, xsize), dtype=np.int64)
    for x, y, z in cloud:
        fx, fy = (x - xoff) / resolution, (yoff - y) / resolution
        ci, cj = int(math.floor(fx)), int(math.floor(fy))
        for dj in range(-radius, radius + 1):
            for di in range(-radius, radius + 1):
                ii, jj = ci + di, cj + dj
                if not (0 <= ii < xsize and 0 <= jj < ysize):
                    continue
                if radius > 0 and (fx - (ii + 0.5)) ** 2 + (fy - (jj + 0.5)) ** 2 > radius * radius:
                    continue
                s[jj, ii] += z; c[jj, ii] += 1
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(c > 0, s / np.maximum(c, 1), np.nan).astype(np.float32)


def synthetic_rpc(seed=0, width=512, height=512, lon0=-81.66, lat0=30.33, alt0=20.0):
    """An RPC shaped like a WorldView-3 crop over Jacksonville (the reference's JAX scenes): nearly affine in (lon, lat) with an
    altitude parallax and small higher-order terms, denominators close to 1."""
    r = np.random.default_rng(seed)
    def poly(lead):
        c = 1e-3 * r.standard_normal(20)
        c[4:] *= 0.2
        for k, v in lead.items():
            c[k] = v
        return c.tolist()
    return {"col_num": poly({0: 0.002, 1: 1.0, 2: 0.03, 3: -0.06}), "col_den": poly({0: 1.0}),
            "row_num": poly({0: -0.001, 1: 0.02, 2: -1.0, 3: 0.04}), "row_den": poly({0: 1.0}),
            "col_offset": width / 2.0, "col_scale": width / 2.0, "row_offset": height / 2.0, "row_scale": height / 2.0,
            "lon_offset": lon0, "lon_scale": 0.0015 * width / 512, "lat_offset": lat0, "lat_scale": 0.0013 * height / 512,
            "alt_offset": alt0, "alt_scale": 60.0}
