// Producer / issuer halves of the warp-specialised tile pipeline, shared by the fused kernels that follow the protocol of
// tc_field.cu (see the header comment there and DESIGN.md 4.1):
//
//   warp 0   pipe_producer   streams this CTA's half (CG = 2) of every pre-swizzled weight tile into the shared-memory ring
//   warp 1   pipe_issuer     leader CTA: tcgen05.mma issue loop, two ring stages (8 MMAs) per elected region;
//            pipe_relay      peer CTA: reports "my half of the stage has landed" to the leader's stage barrier
//
// Barriers (one set per CTA, see Smem): full[] / empty[] per ring stage; a_ready (first ready signal of a GEMM: K-slabs
// [0, k_early) of the A tile and the accumulator columns of chunk 0 are usable), a_ready2 (the rest); acc_full / acc_full2
// (accumulator of N-chunk 0 / 1 complete); slab_free[s] (the last chunk's MMAs are done with K-slab s of the A tile).
// The weight stream is the concatenation, GEMM by GEMM and chunk by chunk, of [chunk_n rows][64 fp16] 128B-swizzled tiles
// (k_slabs per chunk); GEMMs with TcGemm::aux append one 32B-swizzled [chunk_n][16] tile per chunk (forward only).
#pragma once
#include "tc_common.cuh"

namespace snb {

template <int CG>
__device__ __forceinline__ void pipe_producer(const TcProgram& P, const Smem& sm, const unsigned char* packed, int stage_bytes,
                                              int unit, int n_units, int n_work, int tiles_per_group, uint32_t cta_rank, int lane) {
    int st = 0; uint32_t ph = 0;
    for (int wk = unit; wk < n_work; wk += n_units) {
        for (int t = 0; t < tiles_per_group; ++t) {
            const unsigned char* src = packed;
            for (int gi = 0; gi < P.n_gemms; ++gi) {
                const int per_chunk = P.g[gi].k_slabs + (P.g[gi].aux ? 1 : 0);
                const int n = P.g[gi].n_chunks * per_chunk;
                for (int i = 0, sc = 0; i < n; ++i) {
                    const uint32_t tile_bytes = (uint32_t)P.g[gi].chunk_n * (sc < P.g[gi].k_slabs ? 128u : 32u), bytes = tile_bytes / CG;
                    if (lane == 0) {
                        mbar_wait(&sm.empty[st], ph ^ 1, 1);
                        mbar_arrive_expect_tx(&sm.full[st], bytes);
                        bulk_g2s(sm.b + (size_t)st * stage_bytes, src + cta_rank * bytes, bytes, &sm.full[st]);   // this CTA's rows of the tile
                    }
                    __syncwarp();
                    src += tile_bytes;
                    if (++sc == per_chunk) sc = 0;
                    if (++st == P.n_stages) { st = 0; ph ^= 1; }
                }
            }
        }
    }
}

// peer CTA of a pair: when a stage has landed here, arrive on the LEADER's full barrier of that stage (count 2 there).
// Relaxed: a cluster-scope release costs ~1000 cycles per arrival and the data were delivered by the async proxy before this
// CTA's barrier completed.
__device__ __forceinline__ void pipe_relay(const TcProgram& P, const Smem& sm, int unit, int n_units, int n_work, int tiles_per_group, int lane) {
    int st = 0; uint32_t ph = 0;
    const uint32_t leader_full = mapa_u32(smem_u32(sm.full), 0);
    for (int wk = unit; wk < n_work; wk += n_units)
        for (int t = 0; t < tiles_per_group; ++t)
            for (int gi = 0; gi < P.n_gemms; ++gi) {
                const int n = P.g[gi].n_chunks * (P.g[gi].k_slabs + (P.g[gi].aux ? 1 : 0));
                for (int i = 0; i < n; ++i) {
                    mbar_wait(&sm.full[st], ph, 5);
                    if (lane == 0) mbar_arrive_cluster_relaxed(leader_full + (uint32_t)st * 8u);
                    __syncwarp();
                    if (++st == P.n_stages) { st = 0; ph ^= 1; }
                }
            }
}

template <int CG>
__device__ __forceinline__ void pipe_mma(uint32_t d_tm, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    if (CG == 2) umma_f16_ss_2cta(d_tm, da, db, idesc, acc); else umma_f16_ss(d_tm, da, db, idesc, acc);
}
template <int CG>
__device__ __forceinline__ void pipe_commit(uint64_t* bar) { if (CG == 2) umma_commit_2cta(bar, 3); else umma_commit(bar); }

// Single-thread issue.  The loop is latency-critical: one elected region issues up to two weight stages (8 MMAs, 1024 tensor
// cycles) and must cost less than that -- every lane polls the barriers, the GEMM's fields live in registers, descriptors are
// 64-bit adds on precomputed bases.  TcGemm::accumulate = 1: the GEMM adds onto what the previous one left in TMEM.
template <int CG>
__device__ __forceinline__ void pipe_issuer(const TcProgram& P, const Smem& sm, int stage_bytes, int unit, int n_units, int n_work,
                                            int tiles_per_group, uint32_t tmem) {
    int st = 0; uint32_t ph = 0, ready_ph = 0;
    const uint32_t a_base = smem_u32(sm.a), b_base = smem_u32(sm.b);
    const uint64_t a_desc0 = umma_desc_k_sw128(a_base), b_desc0 = umma_desc_k_sw128(b_base);
    const uint32_t stage_desc = (uint32_t)(stage_bytes >> 4);
    const uint64_t aux_a_desc = umma_desc_k_sw32(smem_u32(sm.a_aux)), aux_b_desc0 = umma_desc_k_sw32(b_base);
    const bool deep = P.n_stages >= 4;
    for (int wk = unit; wk < n_work; wk += n_units) {
        for (int t = 0; t < tiles_per_group; ++t) {
            for (int gi = 0; gi < P.n_gemms; ++gi) {
                const int k_slabs = P.g[gi].k_slabs, n_chunks = P.g[gi].n_chunks, chunk_n = P.g[gi].chunk_n, K = P.g[gi].K;
                const int k_early = P.g[gi].k_early, free_slabs = P.g[gi].free_slabs, aux = P.g[gi].aux;
                const uint32_t acc0 = P.g[gi].accumulate ? 1u : 0u;
                const uint32_t idesc = CG == 2 ? umma_idesc_f16_m256((uint32_t)chunk_n) : umma_idesc_f16((uint32_t)chunk_n);
                mbar_wait(sm.a_ready, ready_ph, 2);          // (both ready barriers flip once per GEMM: one parity)
                for (int j = 0; j < n_chunks; ++j) {
                    const uint32_t d_tm = tmem + (uint32_t)(j * chunk_n);
                    for (int s = 0; s < k_slabs;) {
                        if (j == 0 && s == k_early) mbar_wait(sm.a_ready2, ready_ph, 7);     // the rest of the input tile / accumulator columns of the later chunks
                        const bool pair = deep && s + 1 < k_slabs && !(j == 0 && s + 1 == k_early);
                        int st1 = st + 1; uint32_t ph1 = ph; if (st1 == P.n_stages) { st1 = 0; ph1 ^= 1; }
                        const uint64_t da = a_desc0 + (uint64_t)((uint32_t)s * (kSlabBytes >> 4));
                        const uint64_t db = b_desc0 + (uint64_t)((uint32_t)st * stage_desc);
                        const uint64_t db1 = b_desc0 + (uint64_t)((uint32_t)st1 * stage_desc);
                        int ksteps = K - s * 64; ksteps = (ksteps > 64 ? 64 : ksteps) >> 4;
                        int ksteps1 = K - (s + 1) * 64; ksteps1 = (ksteps1 > 64 ? 64 : ksteps1) >> 4;
                        mbar_wait(&sm.full[st], ph, 3);           // (pair mode: this CTA's copy and the peer's relay arrival)
                        if (pair) mbar_wait(&sm.full[st1], ph1, 3);
                        tc_fence_after();
                        if (elect_one()) {        // warp-uniform operands + elect: UTCHMMA takes uniform registers directly
                            if (ksteps == 4) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) pipe_mma<CG>(d_tm, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, ((s | k) != 0) ? 1u : acc0);
                            } else {
                                for (int k = 0; k < ksteps; ++k) pipe_mma<CG>(d_tm, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, ((s | k) != 0) ? 1u : acc0);
                            }
                            pipe_commit<CG>(&sm.empty[st]);
                            if (j >= 1 && s < free_slabs) pipe_commit<CG>(&sm.slab_free[s]);
                            if (pair) {
                                const uint64_t da1 = da + (uint64_t)(kSlabBytes >> 4);
                                if (ksteps1 == 4) {
#pragma unroll
                                    for (int k = 0; k < 4; ++k) pipe_mma<CG>(d_tm, da1 + (uint64_t)(2 * k), db1 + (uint64_t)(2 * k), idesc, 1u);
                                } else {
                                    for (int k = 0; k < ksteps1; ++k) pipe_mma<CG>(d_tm, da1 + (uint64_t)(2 * k), db1 + (uint64_t)(2 * k), idesc, 1u);
                                }
                                pipe_commit<CG>(&sm.empty[st1]);
                                if (j >= 1 && s + 1 < free_slabs) pipe_commit<CG>(&sm.slab_free[s + 1]);
                            }
                            // one barrier per N-chunk index: two commits of one GEMM on a single barrier could both land before the
                            // epilogue looks at it, and the parity wait would miss a phase
                            if (!aux && s + (pair ? 2 : 1) == k_slabs) pipe_commit<CG>(j == 0 ? sm.acc_full : sm.acc_full2);
                        }
                        __syncwarp();
                        if (pair) { s += 2; st = st1 + 1; ph = ph1; if (st == P.n_stages) { st = 0; ph ^= 1; } }
                        else { s += 1; st = st1; ph = ph1; }
                    }
                    if (aux) {
                        // the chunk's last K-step: [1 1 | xyz..] x [bias hi lo | W_xyz..] adds the bias (and the skip layer's xyz term)
                        mbar_wait(&sm.full[st], ph, 3);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t dbx = aux_b_desc0 + (uint64_t)((uint32_t)st * stage_desc);
                            pipe_mma<CG>(d_tm, aux_a_desc, dbx, idesc, 1u);
                            pipe_commit<CG>(&sm.empty[st]);
                            pipe_commit<CG>(j == 0 ? sm.acc_full : sm.acc_full2);
                        }
                        __syncwarp();
                        if (++st == P.n_stages) { st = 0; ph ^= 1; }
                    }
                }
                ready_ph ^= 1;
            }
        }
    }
}

// Fills the pipeline fields of a GEMM list: two_idx / n_two (acc_full2 phases), free_slabs / store2_idx / n_store2 (slab_free
// phases) and k_early (early start of the next GEMM on the K-slabs chunk 0's epilogue has already rewritten).
// stores(g): the GEMM's epilogue rewrites the A tile from its accumulators, chunk by chunk (=> chunk 0 may go straight into the
// K-slabs the last chunk's MMAs have released, and the next GEMM may start on them early).
template <class StoresFn>
inline void pipe_finish_program(TcProgram* P, StoresFn stores, int first_k_early) {
    const int ng = P->n_gemms;
    P->n_two = 0; P->n_store2 = 0;
    for (int i = 0; i < ng; ++i) {
        TcGemm& g = P->g[i];
        g.two_idx = P->n_two; if (g.n_chunks == 2) ++P->n_two;
        g.store2_idx = P->n_store2; g.free_slabs = 0;
        if (stores(g) && g.n_chunks == 2) {
            int fs = (g.chunk_n + 63) / 64; if (fs > g.k_slabs) fs = g.k_slabs;     // only K-slabs this GEMM's MMAs actually walk are released one by one
            g.free_slabs = fs; ++P->n_store2;
        }
    }
    P->g[0].k_early = first_k_early;
    for (int i = 1; i < ng; ++i) {
        const TcGemm& pr = P->g[i - 1];
        int ke = 0;
        if (pr.n_chunks == 2 && stores(pr) && pr.chunk_n % 64 == 0) ke = pr.chunk_n / 64;
        // at least one stage of every GEMM waits for the second signal of BOTH CTAs of a pair (barrier phase aliasing otherwise)
        if (ke > P->g[i].k_slabs - 1) ke = P->g[i].k_slabs - 1;
        P->g[i].k_early = ke;
    }
}

}  // namespace snb
