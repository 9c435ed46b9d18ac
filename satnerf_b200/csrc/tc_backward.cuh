// Interfaces of the tensor-core weight-gradient kernel (tc_backward.cu).
#pragma once
#include "common.cuh"

namespace snb {

constexpr int kDwPairs = 74;             // CTA pairs of the launch (148 SMs); piece i runs on pair i % kDwPairs
constexpr int kDwMaxAcc = 3;

// One piece of weight-gradient work, run by one CTA pair over the 128-point tiles [k_tile0, k_tile0 + k_tiles):
//  n_acc == 1  "main": rows = up to 256 features of the M operand (feature groups [a_fg0, a_fg0 + a_nfg), a_nfg <= 4; rows of
//              missing groups hold copies and are ignored), columns = 64 * b_nfg features of the N operand (b_nfg <= 8: all 512
//              TMEM columns);
//  n_acc  > 1  "small": up to three M operands (256-feature blocks of different arrays) against ONE 64-feature N operand
//              (the extra-input block [x sun t 1] or the head-gradient block), one accumulator each -- so that a small piece
//              streams about as many bytes per stage as a main piece and sweeps the points in step with the main pieces that
//              read the same arrays (they are adjacent in the list and run at the same time: the second reader hits L2).
struct DwItem {
    long long a_off[kDwMaxAcc]; int a_fgs[kDwMaxAcc], a_fg0[kDwMaxAcc], a_nfg[kDwMaxAcc]; int n_acc;
    long long b_off; int b_fgs, b_fg0, b_nfg;      // byte offsets are from `base`; *_fgs = feature groups (of 64) per point tile in the array
    int k_tile0, k_tiles;
    long long out_off[kDwMaxAcc];                  // float offsets of the [256][64 * b_nfg] fp32 partials in the partial buffer
    int sync_off, sync_n;                          // pacing group (see kDwSyncEvery): first counter of the group in `sync`, members; sync_n <= 1: none
};

// Pieces that read the same arrays only share them through L2 while they stay within a few stages of each other (the whole
// machine streams ~5 MB per microsecond through a 126 MB cache).  Every kDwSyncEvery stages the leader CTA's producer of each
// member of a group arrives on a global counter and waits (bounded: pacing only, never a correctness condition -- pieces of a
// group that are not co-resident simply time out) until all members have passed the previous checkpoint.
constexpr int kDwSyncEvery = 8;

int launch_atoms_pack(const float* src, int P, int F, int ld, void* dst, cudaStream_t st);
int launch_dw(const DwItem* d_items, int n_items, const void* base, float* partial, int* sync, cudaStream_t st);

}  // namespace snb
