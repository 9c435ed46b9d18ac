// Interfaces of the tcgen05 (sm_100a tensor-core) render path (tc_field.cu).
#pragma once
#include "common.cuh"

namespace snb {

int tc_workspace(const FieldLayout& L, const snb_pass_desc* p, bool backward, size_t* bytes);
int tc_render_forward(const FieldLayout& L, const snb_pass_desc* p, const snb_render_io* io,
                      void* workspace, size_t workspace_bytes, cudaStream_t st);
// returns 1 when no tensor-core backward exists for this configuration (caller falls back to the
// fp32 CUDA-core backward chain, which is still a GPU path), 0 on success, <0 on error.
int tc_render_backward(const FieldLayout& L, const snb_pass_desc* p, const snb_render_io* io,
                       const snb_render_grads* g, void* workspace, size_t workspace_bytes, cudaStream_t st);

int tc_debug_read(void* dst, size_t bytes);
int tc_debug_hang_info(unsigned int* out4);
bool tc_bwd_supported(const FieldLayout& L, const snb_pass_desc* p);
int tc_bwd_workspace(const FieldLayout& L, const snb_pass_desc* p, size_t* bytes);
int tc_stash_bytes(const FieldLayout& L, const snb_pass_desc* p, size_t* bytes);

}  // namespace snb
