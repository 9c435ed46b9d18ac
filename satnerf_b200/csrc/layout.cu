// Flat parameter layout of the three fields + error plumbing.
// Order = state_dict() order of the reference modules: models/satnerf.py:104-153 (fc_net,
// sigma_from_xyz, feats_from_xyz, rgb_from_xyzdir, sun_v_net, sky_color, beta_from_xyz),
// models/snerf.py:100-146 (same without beta), models/nerf.py:157-177 (no sun/sky/beta).
#include "common.cuh"

namespace snb {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}

int build_layout(const snb_field_desc* f, FieldLayout* L) {
    if (!f || !L) SNB_FAIL(-1, "null field descriptor");
    if (f->variant < SNB_NERF || f->variant > SNB_SATNERF) SNB_FAIL(-1, "unknown variant %d", f->variant);
    if (f->n_layers < 2 || f->n_layers > kMaxTrunk) SNB_FAIL(-1, "fc_layers=%d unsupported (2..%d)", f->n_layers, kMaxTrunk);
    if (f->width < 8 || f->width % 2) SNB_FAIL(-1, "fc_units=%d unsupported (even, >=8)", f->width);
    if (f->skip_layer == 0 || f->skip_layer >= f->n_layers) SNB_FAIL(-1, "skip layer %d out of range", f->skip_layer);
    if (f->variant == SNB_SATNERF && (f->t_dims < 1 || f->t_dims > 64)) SNB_FAIL(-1, "t_embbeding_tau=%d unsupported", f->t_dims);
    memset(L, 0, sizeof(*L));
    L->variant = f->variant; L->n_layers = f->n_layers; L->width = f->width; L->skip = f->skip_layer;
    L->t_dims = f->variant == SNB_SATNERF ? f->t_dims : 0;
    L->in_xyz = f->pe_xyz > 0 ? 6 * f->pe_xyz : 3;
    L->in_dir = f->variant == SNB_NERF ? (f->pe_dir > 0 ? 6 * f->pe_dir : 3) : 0;
    const int h = f->width, h2 = h / 2;
    int64_t off = 0; int n = 0;
    auto lin = [&](int n_out, int n_in) {
        Lin l; l.w = off; off += (int64_t)n_out * n_in; l.b = off; off += n_out; l.n_out = n_out; l.n_in = n_in; ++n; return l;
    };
    for (int i = 0; i < f->n_layers; ++i)
        L->trunk[i] = lin(h, i == 0 ? L->in_xyz : (i == f->skip_layer ? h + L->in_xyz : h));
    L->sigma = lin(1, h);
    L->feats = lin(h, h);
    L->rgb0 = lin(h2, h + L->in_dir);
    L->rgb2 = lin(3, h2);
    if (f->variant != SNB_NERF) {
        L->sun[0] = lin(h2, h + 3); L->sun[1] = lin(h2, h2); L->sun[2] = lin(h2, h2); L->sun[3] = lin(1, h2);
        L->sky0 = lin(h2, 3); L->sky2 = lin(3, h2);
    }
    if (f->variant == SNB_SATNERF) { L->beta0 = lin(h2, h + L->t_dims); L->beta2 = lin(1, h2); }
    L->n_lin = n; L->n_params = off;
    L->n_channels = f->variant == SNB_SATNERF ? 9 : (f->variant == SNB_SNERF ? 8 : 4);
    return 0;
}

}  // namespace snb

using namespace snb;

extern "C" SNB_API int snb_abi_version(void) { return SNB_ABI_VERSION; }
extern "C" SNB_API const char* snb_last_error(void) { return g_err; }

extern "C" SNB_API int64_t snb_launch_count(int reset) {
    int64_t n = (int64_t)g_launches; if (reset) g_launches = 0; return n;
}

extern "C" SNB_API int snb_device_supports_tc(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}

extern "C" SNB_API int64_t snb_param_count(const snb_field_desc* f) {
    FieldLayout L; int r = build_layout(f, &L); return r ? r : L.n_params;
}

extern "C" SNB_API int snb_param_layout(const snb_field_desc* f, int64_t* w_off, int64_t* b_off, int32_t* n_out, int32_t* n_in, int cap) {
    FieldLayout L; SNB_TRY(build_layout(f, &L));
    const Lin* all[64]; int n = 0;
    for (int i = 0; i < L.n_layers; ++i) all[n++] = &L.trunk[i];
    all[n++] = &L.sigma; all[n++] = &L.feats; all[n++] = &L.rgb0; all[n++] = &L.rgb2;
    if (L.variant != SNB_NERF) { for (int i = 0; i < 4; ++i) all[n++] = &L.sun[i]; all[n++] = &L.sky0; all[n++] = &L.sky2; }
    if (L.variant == SNB_SATNERF) { all[n++] = &L.beta0; all[n++] = &L.beta2; }
    for (int i = 0; i < n && i < cap; ++i) {
        if (w_off) w_off[i] = all[i]->w;
        if (b_off) b_off[i] = all[i]->b;
        if (n_out) n_out[i] = all[i]->n_out;
        if (n_in) n_in[i] = all[i]->n_in;
    }
    return n;
}
