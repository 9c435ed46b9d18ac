// Interfaces of the tensor-core backward building blocks (tc_backward.cu).
#pragma once
#include "common.cuh"

namespace snb {

// One output tile of a weight-gradient GEMM: rows = 128 features of the M operand (feature groups a_fg0, a_fg0+1),
// columns = 64*b_nfg features of the N operand, reduced over 128-point tiles [k_tile0, k_tile0 + k_tiles).
struct DwItem {
    long long a_off, b_off;      // byte offsets (from `base`) of the two operands in atom layout
    int a_fgs, b_fgs;            // feature groups (of 64) per point tile in each operand array
    int a_fg0, b_fg0, b_nfg;     // b_nfg in 1..4
    int k_tile0, k_tiles;
    long long out_off;           // float offset of the [128][64*b_nfg] fp32 partial in the partial buffer
};

int launch_atoms_pack(const float* src, int P, int F, int ld, void* dst, cudaStream_t st);
int launch_dw(const DwItem* d_items, int n_items, const void* base, float* partial, cudaStream_t st);

}  // namespace snb
