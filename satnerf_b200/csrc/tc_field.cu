#include "tc_field.cuh"
namespace snb {
int tc_workspace(const FieldLayout&, const snb_pass_desc*, bool, size_t* bytes) { *bytes = 0; return 0; }
int tc_render_forward(const FieldLayout&, const snb_pass_desc*, const snb_render_io*, void*, size_t, cudaStream_t) {
    SNB_FAIL(-5, "tensor-core path not built yet");
}
int tc_render_backward(const FieldLayout&, const snb_pass_desc*, const snb_render_io*, const snb_render_grads*, void*, size_t, cudaStream_t) { return 1; }
}
