// Internal helpers shared by the translation units of libsatnerf_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "satnerf_b200.h"

namespace snb {

// ---- error reporting (thread-local message behind snb_last_error) ----
void set_error(const char* fmt, ...);
#define SNB_FAIL(code, ...) do { ::snb::set_error(__VA_ARGS__); return (code); } while (0)
#define SNB_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) \
    SNB_FAIL(-2, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
extern unsigned long long g_launches;      // kernels launched by this library (snb_launch_count)
#define SNB_CHECK_LAUNCH() do { ++::snb::g_launches; SNB_CUDA(cudaGetLastError()); } while (0)
#define SNB_TRY(expr) do { int r_ = (expr); if (r_ != 0) return r_; } while (0)

// ---- flat parameter layout -------------------------------------------------------------------
struct Lin { int64_t w, b; int n_out, n_in; };   // offsets in floats into the flat buffer

constexpr int kMaxTrunk = 16;
struct FieldLayout {
    int variant, n_layers, width, skip, t_dims, in_xyz, in_dir;   // in_xyz = 3 or 6*pe_xyz
    Lin trunk[kMaxTrunk];
    Lin sigma, feats, rgb0, rgb2;
    Lin sun[4];          // sun_v_net.{0,2,4,6}
    Lin sky0, sky2;
    Lin beta0, beta2;
    int n_lin;
    int64_t n_params;
    int n_channels;      // 9 / 8 / 4
};
int build_layout(const snb_field_desc* f, FieldLayout* out);   // validates the descriptor

// ---- bump allocator over the caller-provided workspace -----------------------------------------
struct Arena {
    char* base; size_t cap, off; bool overflow;
    Arena(void* p, size_t n) : base((char*)p), cap(n), off(0), overflow(false) {}
    template <class T> T* take(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
        if (base == nullptr) { off += bytes; return nullptr; }       // sizing pass
        if (off + bytes > cap) { overflow = true; return nullptr; }
        T* r = (T*)(base + off); off += bytes; return r;
    }
};

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- activations (shared by forward epilogues and backward) ------------------------------------
enum Act { ACT_NONE = 0, ACT_SIN = 1, ACT_SIN30 = 2, ACT_RELU = 3, ACT_SIGMOID = 4, ACT_SOFTPLUS = 5, ACT_SIGMOID_PAD = 6 };

}  // namespace snb
