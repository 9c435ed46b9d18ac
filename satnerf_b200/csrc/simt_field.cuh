// Interfaces of the fp32 CUDA-core field evaluation (simt_field.cu).
#pragma once
#include "common.cuh"

namespace snb {

// A (virtual) matrix operand: element(m, k) = p[(m / div) * ld + k].  div = S broadcasts a per-ray
// row to the S samples of that ray (the reference's repeat_interleave, models/satnerf.py:25-27).
struct Src { const float* p; int ld, k, div; };

struct LinFwd  { Src a0, a1; const float* W; int ldw; const float* b; float* pre; float* out; int ldo; int M, N; };
struct LinBwdIn { const float* dY; int ldy, N; const float* W; int ldw, k_off, K; const float* pre; float* dX; int ldx, accumulate, M; };
struct LinBwdW { const float* dY; int ldy, N; Src a0, a1; float* partial; int M, m_per_z; };

struct FieldInputs { Src xyz, aux, temb; int n_points; int x3 = 0; };      // x3: forward contractions on the tensor cores with fp16 hi+lo operands (SNB_FP16X3_TC):
                                                                           // 2 = build the packed weights in the chunk's x3_w first (first chunk of a pass), 1 = reuse them

// Buffers for one chunk of points, carved from the caller's workspace.
struct FieldChunk {
    float *enc, *enc_dir;
    float *pre[kMaxTrunk], *act[kMaxTrunk];
    float *feat, *rgb1_pre, *rgb1, *sun_pre[3], *sun_act[3], *sky1_pre, *sky1, *beta1_pre, *beta1;
    float *d_feat, *d_a, *d_b, *d_t, *partial;
    unsigned char* x3_w;       // SNB_FP16X3_TC: packed hi / lo weight tiles (built by the first chunk of a pass)
    // keep=true allocates one buffer per layer (pre- and post-activation) for the backward pass
    size_t plan(Arena& ar, const FieldLayout& L, int Pc, int Rc, bool keep);
};

int field_forward_chunk(const FieldLayout& L, const float* params, const FieldChunk& c, const FieldInputs& in,
                        float* raw, bool sigma_only, cudaStream_t st);
int field_backward_chunk(const FieldLayout& L, const float* params, float* g_params, const FieldChunk& c,
                         const FieldInputs& in, const float* d_head, float* g_t_ray, int S, cudaStream_t st);
int launch_points(const float* rays, int ray_cols, int dir_col, const float* z, float* xyz, int r0, int n_rays, int S, cudaStream_t st);

}  // namespace snb
