"""CPU checks of the geometry oracle (oracle/geo_oracle.py).  rpcm, pyproj and plyflatten are not vendored in the reference
and not installed here (parity unpinned for them, see the oracle's header): the restatements are anchored on properties and
on one public known answer (the UTM coordinates of the CN Tower, zone 17T: 630 084 m E, 4 833 438 m N)."""
import numpy as np

from oracle import geo_oracle as g


def test_rpc_localization_inverts_projection():
    rpc = g.synthetic_rpc(seed=3)
    cols, rows = np.meshgrid(np.arange(0, 512, 37.0), np.arange(0, 512, 41.0))
    for alt in (-20.0, 20.0, 75.0):
        lon, lat = g.localization(rpc, cols.ravel(), rows.ravel(), np.full(cols.size, alt))
        c, r = g.projection(rpc, lon, lat, alt)
        assert np.abs(c - cols.ravel()).max() < 1e-6 and np.abs(r - rows.ravel()).max() < 1e-6      # pixels


def test_rays_are_unit_and_bounded_by_the_altitude_planes():
    rpc = g.synthetic_rpc(seed=4)
    cols, rows = np.meshgrid(np.arange(0, 512, 64.0), np.arange(0, 512, 64.0))
    rays = g.get_rays(cols.ravel(), rows.ravel(), rpc, -10.0, 60.0).astype(np.float64)
    assert np.abs(np.linalg.norm(rays[:, 3:6], axis=1) - 1).max() < 1e-6
    assert (rays[:, 6] == 0).all() and (rays[:, 7] > 70.0).all() and (rays[:, 7] < 120.0).all()      # 70 m of altitude seen obliquely
    far = rays[:, 0:3] + rays[:, 3:6] * rays[:, 7:8]
    _, _, alt_near = g.ecef_to_latlon_custom(rays[:, 0], rays[:, 1], rays[:, 2])
    _, _, alt_far = g.ecef_to_latlon_custom(far[:, 0], far[:, 1], far[:, 2])
    assert np.abs(alt_near - 60.0).max() < 1.0 and np.abs(alt_far + 10.0).max() < 1.0               # float32 ECEF coordinates: ~0.5 m


def test_ecef_round_trip():
    lat, lon, alt = np.array([30.33, -12.5, 67.0]), np.array([-81.66, 140.2, 3.0]), np.array([12.0, 850.0, -30.0])
    la, lo, al = g.ecef_to_latlon_custom(*g.latlon_to_ecef_custom(lat, lon, alt))
    assert np.abs(la - lat).max() < 1e-9 and np.abs(lo - lon).max() < 1e-9 and np.abs(al - alt).max() < 1e-3      # (Bowring, one step)


def test_utm_known_answer_and_round_trip():
    e, n = g.utm_forward(np.array([43.642566]), np.array([-79.387139]), 17)
    assert abs(e[0] - 630084) < 1.0 and abs(n[0] - 4833438) < 1.0
    lat = np.array([30.33, 30.5, 29.9, 0.0]); lon = np.array([-81.66, -80.1, -83.9, -81.0])
    e, n = g.utm_forward(lat, lon, 17)
    la, lo = g.utm_inverse(e, n, 17)
    assert np.abs(la - lat).max() < 1e-10 and np.abs(lo - lon).max() < 1e-10
    e0, n0 = g.utm_forward(np.array([0.0]), np.array([-81.0]), 17)              # central meridian, equator
    assert abs(e0[0] - 500000.0) < 1e-6 and abs(n0[0]) < 1e-6
    e1, n1 = g.utm_forward(np.array([1e-4]), np.array([-81.0]), 17)             # scale factor 0.9996 on the central meridian
    assert abs(n1[0] / (1e-4 * np.pi / 180 * 6335439.327) - 0.9996) < 1e-6      # meridional radius of curvature at the equator


def test_plyflatten_of_a_plane():
    xs, ys = np.meshgrid(np.arange(100.25, 110, 0.5), np.arange(200.25, 206, 0.5))
    cloud = np.stack([xs.ravel(), ys.ravel(), 0.1 * xs.ravel() + 3.0], 1)
    xoff, yoff, xsize, ysize = g.dsm_bounds(cloud, 0.5)
    dsm = g.plyflatten(cloud, xoff, yoff, 0.5, xsize, ysize, radius=0)
    assert dsm.shape == (ysize, xsize) and np.isfinite(dsm[:12]).all()      # (the reference's bounds leave one empty row below the lowest point)
    assert np.abs(dsm[0] - (0.1 * (xoff + 0.5 * (np.arange(xsize) + 0.5)) + 3.0)).max() < 1e-5      # one point per cell, at its centre
    sparse = g.plyflatten(cloud[::7], xoff, yoff, 0.5, xsize, ysize, radius=0)
    assert np.isnan(sparse[:12]).any()
    filled = g.plyflatten(cloud[::7], xoff, yoff, 0.5, xsize, ysize, radius=1)
    assert np.isnan(filled).sum() < np.isnan(sparse).sum()
