// kc_vote.cuh — K1: vote consensus (mode with first-seen ties) over dictionary-coded groups.
//
// Replaces voting_consensus (reference consensus_utils.py:936-982) on pre-sanitised input.
// One thread owns one group (the n candidate cells of one field of one record) in registers.
// HBM-bound streaming op: 4n bytes in, 8 bytes out per group, O(1) integer ops per byte, no reuse,
// no tensor cores.  Two front-ends feed the same register core:
//   * vote_tma_kernel    — n in {8,16,32,64}: persistent CTAs, TMA 2-D tiled loads into hardware-swizzled
//                          shared memory through an mbarrier ring, conflict-free LDS.128 per thread.
//   * vote_direct_kernel — any n <= 64 (and n <= 4 where a thread's cells are one coalesced vector load).
#pragma once

#include "kc_common.cuh"

namespace kc {

template <int N>
struct MaskOf {
    using type = uint32_t;
};
template <>
struct MaskOf<64> {
    using type = uint64_t;
};
__device__ __forceinline__ int popc_m(uint32_t m) { return __popc(m); }
__device__ __forceinline__ int popc_m(uint64_t m) { return __popcll(m); }

// Mode of the voting cells of one group.  raw[i]: code >= 0, KC_CODE_NONE (-1) or absent (< -1).
// none_code >= 0 makes None cells vote as that code (bool fields: None -> False, cu:956).
// Classes are visited in first-seen order and a later class must be STRICTLY larger to win, which is
// exactly Counter.most_common(1) (cu:958,969).  The scan stops once the unvisited cells cannot reach
// the best count, so agreeing data costs ~2 passes instead of n.
template <int N>
__device__ __forceinline__ void vote_core(const int32_t (&raw)[N], int32_t none_code, int32_t &win_code, uint32_t &meta) {
    using M = typename MaskOf<N>::type;
    int32_t v[N];
    M live = 0;
    int present = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        int32_t c = raw[i];
        const bool absent = c < KC_CODE_NONE;
        present += absent ? 0 : 1;
        c = (c == KC_CODE_NONE) ? none_code : c;
        c = absent ? KC_CODE_NONE : c;
        v[i] = c;
        live |= (c >= 0) ? (M(1) << i) : M(0);
    }
    const int voters = popc_m(live);
    int best_cnt = 0, best_idx = 0;
    int32_t best_code = KC_CODE_NONE;
    bool tie = false;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (live == 0 || popc_m(live) < best_cnt) break;
        if ((live >> i) & 1) {
            const int32_t c = v[i];
            M eq = 0;
#pragma unroll
            for (int j = i; j < N; ++j) eq |= (v[j] == c) ? (M(1) << j) : M(0);
            const int cnt = popc_m(eq);
            if (cnt > best_cnt) {
                best_cnt = cnt;
                best_idx = i;
                best_code = c;
                tie = false;
            } else if (cnt == best_cnt) {
                tie = true;
            }
            live &= ~eq;
        }
    }
    win_code = best_code;
    meta = pack_meta(best_idx, best_cnt, voters, present,
                     best_cnt > 0 ? (KC_FLAG_HAS_VALUE | (tie ? KC_FLAG_TIE : 0u)) : 0u);
}

// ---------------------------------------------------------------- direct front-end

// NP = n rounded up to a power of two (compile-time register array); cells beyond n are padded absent.
template <int NP, bool VEC>
__global__ void __launch_bounds__(256) vote_direct_kernel(const int32_t *__restrict__ codes, int64_t n_groups, int n,
                                                          const int32_t *__restrict__ none_code, int n_fields,
                                                          int32_t *__restrict__ win, uint32_t *__restrict__ meta) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += stride) {
        int32_t raw[NP];
        if constexpr (VEC) {  // n == NP, rows are 16-byte aligned multiples of 16 bytes
            if constexpr (NP >= 4) {
                const int4 *p = reinterpret_cast<const int4 *>(codes + g * NP);
#pragma unroll
                for (int q = 0; q < NP / 4; ++q) {
                    const int4 t = ldg_stream_v4(p + q);
                    raw[4 * q + 0] = t.x;
                    raw[4 * q + 1] = t.y;
                    raw[4 * q + 2] = t.z;
                    raw[4 * q + 3] = t.w;
                }
            } else if constexpr (NP == 2) {
                const int2 t = __ldg(reinterpret_cast<const int2 *>(codes + g * 2));
                raw[0] = t.x;
                raw[1] = t.y;
            } else {
                raw[0] = __ldg(codes + g);
            }
        } else {
            const int32_t *p = codes + g * n;
#pragma unroll
            for (int i = 0; i < NP; ++i) raw[i] = (i < n) ? __ldg(p + i) : KC_CODE_ABSENT;
        }
        const int32_t nc = none_code ? __ldg(none_code + (g % n_fields)) : KC_CODE_NONE;
        int32_t w;
        uint32_t m;
        vote_core<NP>(raw, nc, w, m);
        stg_stream_u32(win + g, (uint32_t)w);
        stg_stream_u32(meta + g, m);
    }
}

// ---------------------------------------------------------------- TMA front-end

template <int ROW_BYTES>
struct Swizzle {  // TMA swizzle mode for a row of ROW_BYTES (rows wider than 128 B are split into 128 B box rows)
    static constexpr uint32_t kMask = ROW_BYTES >= 128 ? 7u : (ROW_BYTES == 64 ? 3u : 1u);
    __device__ static __forceinline__ uint32_t apply(uint32_t off) { return off ^ (((off >> 7) & kMask) << 4); }
};

// Persistent kernel: CTA b owns tiles b, b+grid, ...; tile = TILE consecutive groups; thread t owns row t.
// Ring of STAGES smem buffers, each filled by ONE cp.async.bulk.tensor.2d (hardware swizzle) that
// completes on the stage's mbarrier.  After the LDS of a stage a __syncthreads() frees it and thread 0
// immediately re-arms it with the tile STAGES ahead, so up to STAGES tiles per CTA are in flight while the
// register core runs.  Out-of-range rows of the last tile are zero-filled by TMA and never stored.
template <int N, int TILE, int STAGES>
__global__ void __launch_bounds__(TILE) vote_tma_kernel(const __grid_constant__ CUtensorMap tmap, int64_t n_groups,
                                                        const int32_t *__restrict__ none_code, int n_fields,
                                                        int32_t *__restrict__ win, uint32_t *__restrict__ meta) {
    constexpr int ROW_BYTES = N * 4;
    constexpr int BOX_ROWS_PER_GROUP = ROW_BYTES > 128 ? ROW_BYTES / 128 : 1;
    constexpr uint32_t STAGE_BYTES = TILE * ROW_BYTES;
    static_assert(STAGE_BYTES % 1024 == 0, "stage must keep the 1024-byte swizzle alignment");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[STAGES];

    const int tid = threadIdx.x;
    const int64_t n_tiles = (n_groups + TILE - 1) / TILE;
    const int64_t first = blockIdx.x;
    const int64_t step = gridDim.x;
    uint64_t policy = 0;

    if (tid == 0) {
        tma_prefetch_desc(&tmap);
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
        fence_barrier_init();
        policy = policy_evict_first();
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            const int64_t t = first + (int64_t)s * step;
            if (t < n_tiles) {
                mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                tma_load_2d(smem + (size_t)s * STAGE_BYTES, &tmap, 0, (int32_t)(t * TILE * BOX_ROWS_PER_GROUP), &full_bar[s],
                            policy);
            }
        }
    }
    __syncthreads();

    int stage = 0;
    uint32_t parity = 0;
    for (int64_t t = first; t < n_tiles; t += step) {
        mbar_wait(&full_bar[stage], parity);
        int32_t raw[N];
        const uint32_t base = smem_u32(smem + (size_t)stage * STAGE_BYTES);
        const uint32_t row_off = (uint32_t)tid * ROW_BYTES;
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
            const int4 v4 = lds_v4(base + Swizzle<ROW_BYTES>::apply(row_off + q * 16));
            raw[4 * q + 0] = v4.x;
            raw[4 * q + 1] = v4.y;
            raw[4 * q + 2] = v4.z;
            raw[4 * q + 3] = v4.w;
        }
        __syncthreads();  // every row of this stage is in registers: the buffer can be refilled
        if (tid == 0) {
            const int64_t tn = t + (int64_t)STAGES * step;
            if (tn < n_tiles) {
                mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                tma_load_2d(smem + (size_t)stage * STAGE_BYTES, &tmap, 0, (int32_t)(tn * TILE * BOX_ROWS_PER_GROUP),
                            &full_bar[stage], policy);
            }
        }
        const int64_t g = t * TILE + tid;
        if (g < n_groups) {
            const int32_t nc = none_code ? __ldg(none_code + (g % n_fields)) : KC_CODE_NONE;
            int32_t w;
            uint32_t m;
            vote_core<N>(raw, nc, w, m);
            stg_stream_u32(win + g, (uint32_t)w);
            stg_stream_u32(meta + g, m);
        }
        if (++stage == STAGES) {
            stage = 0;
            parity ^= 1;
        }
    }
}

}  // namespace kc
