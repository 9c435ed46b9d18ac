// Geometry either side of the render path: rays from an RPC camera model before it, the digital
// surface model (DSM) from the rendered depths after it.  Element-wise float64 work, HBM-bound: one thread per pixel /
// point, coalesced (row-major) reads and writes, grids sized in multiples of the SM count.
//
//   rpc_rays_kernel       datasets/satellite.py:18-65 get_rays + :218-227 normalize_rays + :229-244 get_sun_dirs:
//                         localise every pixel at the maximum and at the minimum altitude through the RPC model (rpcm's
//                         iterative inverse, restated -- rpcm is a dependency that is not vendored), geodetic -> ECEF
//                         (sat_utils.py:59-74), origin / unit direction / bounds, the reference's float32 cast and scene
//                         normalisation, sun direction appended.
//   dsm_points_kernel     datasets/satellite.py:246-274: point = o + d * depth in float64, de-normalised to ECEF,
//                         ECEF -> geodetic (sat_utils.py:76-97), geodetic -> UTM (sat_utils.py:99-113 calls pyproj;
//                         restated here as the Krueger series of order 6, which is what PROJ's transverse Mercator evaluates).
//   dsm_splat / dsm_finish  plyflatten(cloud, xoff, yoff, resolution, xsize, ysize, radius, sigma=inf)
//                         (datasets/satellite.py:308): every point adds its altitude to the cells within `radius` cells of
//                         its own; a cell's value is the mean of what it received, NaN when nothing did.  Sums are kept in
//                         64-bit fixed point (2^-20 m), so the result does not depend on the order of the atomics.
#include "common.cuh"
#include <cmath>

namespace snb {

__device__ __forceinline__ double rpc_poly(const double* __restrict__ c, double x, double y, double z) {
    // RPC00B term order with x = latitude, y = longitude, z = altitude (all normalised)
    double out = c[0];
    out += c[1] * y + c[2] * x + c[3] * z;
    out += c[4] * y * x + c[5] * y * z + c[6] * x * z;
    out += c[7] * y * y + c[8] * x * x + c[9] * z * z;
    out += c[10] * x * y * z;
    out += c[11] * y * y * y;
    out += c[12] * y * x * x + c[13] * y * z * z + c[14] * y * y * x;
    out += c[15] * x * x * x;
    out += c[16] * x * z * z + c[17] * y * y * z + c[18] * x * x * z;
    out += c[19] * z * z * z;
    return out;
}
__device__ __forceinline__ void rpc_project_n(const snb_rpc_model& m, double nlat, double nlon, double nalt, double* col, double* row) {
    *col = rpc_poly(m.col_num, nlat, nlon, nalt) / rpc_poly(m.col_den, nlat, nlon, nalt);
    *row = rpc_poly(m.row_num, nlat, nlon, nalt) / rpc_poly(m.row_den, nlat, nlon, nalt);
}

// rpcm RPCModel.localization_iterative for one pixel: secant steps on the normalised (lon, lat) with the image-space basis
// re-estimated by finite differences (first step 2, then 0.1), until the squared normalised residual is below 1e-18.
__device__ void rpc_localize(const snb_rpc_model& m, double col, double row, double alt, double* lon_out, double* lat_out, int* iters) {
    const double ncol = (col - m.col_offset) / m.col_scale, nrow = (row - m.row_offset) / m.row_scale, nalt = (alt - m.alt_offset) / m.alt_scale;
    double lon = -1.0, lat = -1.0, eps = 2.0;
    double x0, y0, x1, y1, x2, y2;
    rpc_project_n(m, lat, lon, nalt, &x0, &y0);
    rpc_project_n(m, lat, lon + eps, nalt, &x1, &y1);
    rpc_project_n(m, lat + eps, lon, nalt, &x2, &y2);
    int n = 0;
    while (!((x0 - ncol) * (x0 - ncol) + (y0 - nrow) * (y0 - nrow) < 1e-18) && n <= 100) {
        const double e1x = x1 - x0, e1y = y1 - y0, e2x = x2 - x0, e2y = y2 - y0, ux = ncol - x0, uy = nrow - y0;
        const double a1 = (ux * e1x + uy * e1y) / (e1x * e1x + e1y * e1y);      // (the basis is assumed orthogonal, as rpcm does)
        const double a2 = (ux * e2x + uy * e2y) / (e2x * e2x + e2y * e2y);
        lon += a1 * eps; lat += a2 * eps;
        eps = 0.1;
        rpc_project_n(m, lat, lon, nalt, &x0, &y0);
        rpc_project_n(m, lat, lon + eps, nalt, &x1, &y1);
        rpc_project_n(m, lat + eps, lon, nalt, &x2, &y2);
        ++n;
    }
    *lon_out = lon * m.lon_scale + m.lon_offset; *lat_out = lat * m.lat_scale + m.lat_offset; *iters = n;
}

// sat_utils.py:59-74 latlon_to_ecef_custom
__device__ __forceinline__ void latlon_to_ecef(double lat, double lon, double alt, double* x, double* y, double* z) {
    const double rad_lat = lat * (M_PI / 180.0), rad_lon = lon * (M_PI / 180.0);
    const double a = 6378137.0, finv = 298.257223563, f = 1.0 / finv, e2 = 1.0 - (1.0 - f) * (1.0 - f);
    const double sl = sin(rad_lat), cl = cos(rad_lat);
    const double v = a / sqrt(1.0 - e2 * sl * sl);
    *x = (v + alt) * cl * cos(rad_lon); *y = (v + alt) * cl * sin(rad_lon); *z = (v * (1.0 - e2) + alt) * sl;
}

struct RpcRaysArgs {
    snb_rpc_model rpc;
    const double *cols, *rows;       // explicit pixel coordinates, or null: the w x h grid (np.meshgrid(arange(w), arange(h)), row-major)
    int w; long long n;
    double min_alt, max_alt, center[3], range;
    int normalize, has_sun; float sun[3];
    float* rays; int ray_cols;
    int* max_iters;                  // optional: atomicMax of the iteration count (> 100: rpcm raises MaxLocalizationIterationsError)
};

__global__ void rpc_rays_kernel(const __grid_constant__ RpcRaysArgs A) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += (long long)gridDim.x * blockDim.x) {
        const double col = A.cols ? A.cols[i] : (double)(i % A.w), row = A.rows ? A.rows[i] : (double)(i / A.w);
        double lon, lat, xn, yn, zn, xf, yf, zf; int it0, it1;
        rpc_localize(A.rpc, col, row, A.max_alt, &lon, &lat, &it0);         // the points of maximum altitude are the closest to the camera
        latlon_to_ecef(lat, lon, A.max_alt, &xn, &yn, &zn);
        rpc_localize(A.rpc, col, row, A.min_alt, &lon, &lat, &it1);
        latlon_to_ecef(lat, lon, A.min_alt, &xf, &yf, &zf);
        if (A.max_iters) { int m = it0 > it1 ? it0 : it1; if (m > 8) atomicMax(A.max_iters, m); }
        const double dx = xf - xn, dy = yf - yn, dz = zf - zn, len = sqrt(dx * dx + dy * dy + dz * dz);
        // float64 -> float32 (satellite.py:63), then the scene normalisation on the float32 tensor (:218-227)
        float o[3] = {(float)xn, (float)yn, (float)zn}, d[3] = {(float)(dx / len), (float)(dy / len), (float)(dz / len)};
        float nearv = 0.f, farv = (float)len;
        if (A.normalize) {
            const float r = (float)A.range;
#pragma unroll
            for (int k = 0; k < 3; ++k) o[k] = __fdiv_rn(__fsub_rn(o[k], (float)A.center[k]), r);
            nearv = __fdiv_rn(nearv, r); farv = __fdiv_rn(farv, r);
        }
        float* out = A.rays + i * A.ray_cols;
        out[0] = o[0]; out[1] = o[1]; out[2] = o[2]; out[3] = d[0]; out[4] = d[1]; out[5] = d[2]; out[6] = nearv; out[7] = farv;
        if (A.has_sun && A.ray_cols >= 11) { out[8] = A.sun[0]; out[9] = A.sun[1]; out[10] = A.sun[2]; }
    }
}

// ---- DSM ---------------------------------------------------------------------------------------------------------
// sat_utils.py:76-97 ecef_to_latlon_custom
__device__ __forceinline__ void ecef_to_latlon(double x, double y, double z, double* lat, double* lon, double* alt) {
    const double a = 6378137.0, e = 8.1819190842622e-2, asq = a * a, esq = e * e;
    const double b = sqrt(asq * (1.0 - esq)), bsq = b * b, ep = sqrt((asq - bsq) / bsq);
    const double p = sqrt(x * x + y * y), th = atan2(a * z, b * p);
    const double sth = sin(th), cth = cos(th);
    const double lo = atan2(y, x);
    const double la = atan2(z + ep * ep * b * sth * sth * sth, p - esq * a * cth * cth * cth);
    const double sla = sin(la), N = a / sqrt(1.0 - esq * sla * sla);
    *alt = p / cos(la) - N; *lon = lo * 180.0 / M_PI; *lat = la * 180.0 / M_PI;
}
// geodetic (degrees) -> UTM easting / northing of `zone` on WGS84, northern-hemisphere false northing (the reference builds
// "+proj=utm +zone=<number><letter>": PROJ reads the number and no +south flag).  Krueger series, order 6.
__device__ __forceinline__ void latlon_to_utm(double lat, double lon, int zone, double* east, double* north) {
    const double a = 6378137.0, f = 1.0 / 298.257223563, n = f / (2.0 - f), k0 = 0.9996;
    const double n2 = n * n, n3 = n2 * n, n4 = n3 * n, n5 = n4 * n, n6 = n5 * n;
    const double A_ = a / (1.0 + n) * (1.0 + n2 / 4.0 + n4 / 64.0 + n6 / 256.0);
    const double al[6] = {n / 2.0 - 2.0 * n2 / 3.0 + 5.0 * n3 / 16.0 + 41.0 * n4 / 180.0 - 127.0 * n5 / 288.0 + 7891.0 * n6 / 37800.0,
                          13.0 * n2 / 48.0 - 3.0 * n3 / 5.0 + 557.0 * n4 / 1440.0 + 281.0 * n5 / 630.0 - 1983433.0 * n6 / 1935360.0,
                          61.0 * n3 / 240.0 - 103.0 * n4 / 140.0 + 15061.0 * n5 / 26880.0 + 167603.0 * n6 / 181440.0,
                          49561.0 * n4 / 161280.0 - 179.0 * n5 / 168.0 + 6601661.0 * n6 / 7257600.0,
                          34729.0 * n5 / 80640.0 - 3418889.0 * n6 / 1995840.0,
                          212378941.0 * n6 / 319334400.0};
    const double e = sqrt(f * (2.0 - f));
    const double phi = lat * (M_PI / 180.0), lam = (lon - (double)(zone * 6 - 183)) * (M_PI / 180.0);
    const double sp = sin(phi);
    const double t = sinh(atanh(sp) - e * atanh(e * sp));
    const double cl = cos(lam), sl = sin(lam);
    const double xi0 = atan2(t, cl), eta0 = asinh(sl / sqrt(t * t + cl * cl));
    double xi = xi0, eta = eta0;
#pragma unroll
    for (int j = 1; j <= 6; ++j) {
        xi += al[j - 1] * sin(2.0 * j * xi0) * cosh(2.0 * j * eta0);
        eta += al[j - 1] * cos(2.0 * j * xi0) * sinh(2.0 * j * eta0);
    }
    *east = 500000.0 + k0 * A_ * eta; *north = k0 * A_ * xi;
}

struct DsmPointsArgs {
    const float *rays, *depth; int ray_cols; long long n;
    double center[3], range; int zone;            // zone <= 0: no projection, cloud = (lon, lat, alt)
    double* cloud;                                // (n, 3): east, north, alt
    double* latlon;                               // optional (n, 2): lat, lon in degrees
};
__global__ void dsm_points_kernel(const __grid_constant__ DsmPointsArgs A) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += (long long)gridDim.x * blockDim.x) {
        const float* r = A.rays + i * A.ray_cols;
        const double dep = (double)A.depth[i];
        double p[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) p[k] = ((double)r[k] + (double)r[3 + k] * dep) * A.range + A.center[k];      // satellite.py:262-268
        double lat, lon, alt; ecef_to_latlon(p[0], p[1], p[2], &lat, &lon, &alt);
        double e = lon, nn = lat;
        if (A.zone > 0) latlon_to_utm(lat, lon, A.zone, &e, &nn);
        A.cloud[i * 3] = e; A.cloud[i * 3 + 1] = nn; A.cloud[i * 3 + 2] = alt;
        if (A.latlon) { A.latlon[i * 2] = lat; A.latlon[i * 2 + 1] = lon; }
    }
}

struct DsmSplatArgs {
    const double* cloud; long long n;
    double xoff, yoff, resolution; int xsize, ysize, radius;
    long long* sum; int* cnt;                     // (ysize, xsize) fixed-point sums (2^-20 m) and counts, zeroed by the caller
};
constexpr double kDsmFix = 1048576.0;
__global__ void dsm_splat_kernel(const __grid_constant__ DsmSplatArgs A) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += (long long)gridDim.x * blockDim.x) {
        const double x = A.cloud[i * 3], y = A.cloud[i * 3 + 1], z = A.cloud[i * 3 + 2];
        if (!(isfinite(x) && isfinite(y) && isfinite(z))) continue;
        // column / row of the cell that contains the point: x grows to the right from xoff, y DOWN from yoff (the raster's
        // geotransform is (resolution, 0, xoff, 0, -resolution, yoff), satellite.py:329)
        const double fx = (x - A.xoff) / A.resolution, fy = (A.yoff - y) / A.resolution;
        const long long ci = (long long)floor(fx), cj = (long long)floor(fy);
        const long long q = llrint(z * kDsmFix);
        for (int dj = -A.radius; dj <= A.radius; ++dj)
            for (int di = -A.radius; di <= A.radius; ++di) {
                const long long ii = ci + di, jj = cj + dj;
                if (ii < 0 || jj < 0 || ii >= A.xsize || jj >= A.ysize) continue;
                const double ddx = fx - ((double)ii + 0.5), ddy = fy - ((double)jj + 0.5);          // distance to the cell centre, in cells
                if (A.radius > 0 && ddx * ddx + ddy * ddy > (double)A.radius * (double)A.radius) continue;
                atomicAdd(reinterpret_cast<unsigned long long*>(A.sum + jj * A.xsize + ii), (unsigned long long)q);
                atomicAdd(A.cnt + jj * A.xsize + ii, 1);
            }
    }
}
__global__ void dsm_finish_kernel(const long long* __restrict__ sum, const int* __restrict__ cnt, float* __restrict__ dsm, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dsm[i] = cnt[i] > 0 ? (float)(((double)sum[i] / kDsmFix) / (double)cnt[i]) : nanf("");
}

static int grid_for(long long n) { long long b = (n + 255) / 256; if (b > 148 * 16) b = 148 * 16; if (b < 1) b = 1; return (int)b; }

}  // namespace snb

using namespace snb;

extern "C" SNB_API int snb_rpc_rays(const snb_rpc_model* rpc, const double* cols, const double* rows, int width, long long n_pixels,
                                    double min_alt, double max_alt, const double* center3, double range, const float* sun_dir3,
                                    float* rays, int ray_cols, int* max_iters, void* stream) {
    if (!rpc || !rays || n_pixels < 0) SNB_FAIL(-1, "snb_rpc_rays: bad argument");
    if ((cols == nullptr) != (rows == nullptr)) SNB_FAIL(-1, "snb_rpc_rays: cols and rows are given together or not at all");
    if (!cols && width < 1) SNB_FAIL(-1, "snb_rpc_rays: the pixel grid needs its width");
    if (ray_cols != 8 && ray_cols != 11) SNB_FAIL(-1, "snb_rpc_rays: ray_cols must be 8 or 11");
    if (sun_dir3 && ray_cols != 11) SNB_FAIL(-1, "snb_rpc_rays: a sun direction needs 11 ray columns");
    if (center3 && !(range > 0.0)) SNB_FAIL(-1, "snb_rpc_rays: scene range must be positive");
    if (n_pixels == 0) return 0;
    RpcRaysArgs A; memset(&A, 0, sizeof(A));
    A.rpc = *rpc; A.cols = cols; A.rows = rows; A.w = width; A.n = n_pixels; A.min_alt = min_alt; A.max_alt = max_alt;
    A.normalize = center3 != nullptr; A.range = range;
    if (center3) { A.center[0] = center3[0]; A.center[1] = center3[1]; A.center[2] = center3[2]; }
    A.has_sun = sun_dir3 != nullptr;
    if (sun_dir3) { A.sun[0] = sun_dir3[0]; A.sun[1] = sun_dir3[1]; A.sun[2] = sun_dir3[2]; }
    A.rays = rays; A.ray_cols = ray_cols; A.max_iters = max_iters;
    rpc_rays_kernel<<<grid_for(n_pixels), 256, 0, (cudaStream_t)stream>>>(A);
    SNB_CHECK_LAUNCH();
    return 0;
}

extern "C" SNB_API int snb_dsm_points(const float* rays, int ray_cols, const float* depth, long long n_rays, const double* center3, double range,
                                      int utm_zone, double* cloud, double* latlon, void* stream) {
    if (!rays || !depth || !cloud || !center3 || n_rays < 0 || ray_cols < 6) SNB_FAIL(-1, "snb_dsm_points: bad argument");
    if (n_rays == 0) return 0;
    DsmPointsArgs A; memset(&A, 0, sizeof(A));
    A.rays = rays; A.depth = depth; A.ray_cols = ray_cols; A.n = n_rays; A.range = range; A.zone = utm_zone; A.cloud = cloud; A.latlon = latlon;
    A.center[0] = center3[0]; A.center[1] = center3[1]; A.center[2] = center3[2];
    dsm_points_kernel<<<grid_for(n_rays), 256, 0, (cudaStream_t)stream>>>(A);
    SNB_CHECK_LAUNCH();
    return 0;
}

extern "C" SNB_API int snb_dsm_rasterize(const double* cloud, long long n_points, double xoff, double yoff, double resolution, int xsize, int ysize,
                                         int radius, float* dsm, void* workspace, size_t workspace_bytes, void* stream) {
    if (!cloud || !dsm || n_points < 0 || xsize < 1 || ysize < 1 || !(resolution > 0.0) || radius < 0) SNB_FAIL(-1, "snb_dsm_rasterize: bad argument");
    const size_t cells = (size_t)xsize * ysize, need = cells * 12 + 256;
    if (!workspace || workspace_bytes < need) SNB_FAIL(-4, "snb_dsm_rasterize: workspace too small (%zu < %zu)", workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    long long* sum = (long long*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    int* cnt = (int*)(sum + cells);
    SNB_CUDA(cudaMemsetAsync(sum, 0, cells * 12, st));
    if (n_points > 0) {
        DsmSplatArgs A; memset(&A, 0, sizeof(A));
        A.cloud = cloud; A.n = n_points; A.xoff = xoff; A.yoff = yoff; A.resolution = resolution; A.xsize = xsize; A.ysize = ysize; A.radius = radius;
        A.sum = sum; A.cnt = cnt;
        dsm_splat_kernel<<<grid_for(n_points), 256, 0, st>>>(A);
        SNB_CHECK_LAUNCH();
    }
    dsm_finish_kernel<<<grid_for((long long)cells), 256, 0, st>>>(sum, cnt, dsm, (long long)cells);
    SNB_CHECK_LAUNCH();
    return 0;
}
