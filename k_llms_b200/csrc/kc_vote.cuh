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
__device__ __forceinline__ int ffs_mask(uint32_t m) { return __ffs((int)m); }
__device__ __forceinline__ int ffs_mask(uint64_t m) { return __ffsll((long long)m); }

template <typename M, int N>
__host__ __device__ constexpr M full_mask() {
    M m = 0;
    for (int i = 0; i < N; ++i) m |= M(1) << i;
    return m;
}

__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) {  // bitwise majority, one LOP3
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// Exact first-seen-order scan (the general case).  Classes are visited in the order of their first cell and a
// later class must be STRICTLY larger to win, which is Counter.most_common(1) (cu:958,969).  `live` = voting
// cells, `nones` = None cells that vote as `nc` (nc >= 0) and are part of `live`.  The scan stops as soon as
// the unvisited cells cannot reach the best count.
template <int N, typename M>
__device__ __forceinline__ uint32_t vote_scan(const int32_t (&v)[N], M live, M nones, int32_t nc, int present, int32_t &win_code) {
    const int voters = popc_m(live);
    int best_cnt = 0, best_idx = 0;
    int32_t best_code = KC_CODE_NONE;
    bool tie = false;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if ((live >> i) & 1) {
            const int32_t c = (v[i] == KC_CODE_NONE) ? nc : v[i];
            M eq = (c == nc) ? nones : M(0);
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
            if (popc_m(live) < best_cnt) live = 0;  // nothing left can win or tie
        }
    }
    win_code = best_code;
    return pack_meta(best_idx, best_cnt, voters, present,
                     best_cnt > 0 ? (KC_FLAG_HAS_VALUE | (tie ? KC_FLAG_TIE : 0u)) : 0u);
}

// Mode of the voting cells of one group.  raw[i]: code >= 0, KC_CODE_NONE (-1) or absent (< -1).
// none_code >= 0 makes None cells vote as that code (bool fields: None -> False, cu:956).
//
// Fast path: a bitwise majority-of-majorities over nine cells (4 LOP3) guesses the mode; ONE equality pass
// counts it.  If the guess holds a strict majority of the voting cells it is the unique mode (no tie is
// possible) and its first cell is the first-seen original (cu:971) — done in ~1 pass.  Any other outcome
// (no strict majority, garbage guess) falls through to the exact scan, so correctness never depends on the guess.
template <int N>
__device__ __forceinline__ void vote_core(const int32_t (&raw)[N], int32_t none_code, int32_t &win_code, uint32_t &meta) {
    using M = typename MaskOf<N>::type;
    constexpr M kFull = full_mask<M, N>();
    int32_t lo = raw[0];
    M neg = 0;  // cells < 0: None or absent
#pragma unroll
    for (int i = 0; i < N; ++i) {
        lo = min(lo, raw[i]);
        neg |= (raw[i] < 0) ? (M(1) << i) : M(0);
    }
    M absent = 0;
    if (lo < KC_CODE_NONE) {  // rare: some candidate is not part of this node (nested payloads)
#pragma unroll
        for (int i = 0; i < N; ++i) absent |= (raw[i] < KC_CODE_NONE) ? (M(1) << i) : M(0);
    }
    const M nones_all = neg & ~absent;
    const M nones = none_code >= 0 ? nones_all : M(0);  // None cells that vote
    const M live = (kFull & ~neg) | nones;
    const int present = N - popc_m(absent);
    const int voters = popc_m(live);

    if constexpr (N >= 9) {
        const uint32_t guess = maj3(maj3(raw[0], raw[1], raw[2]), maj3(raw[3], raw[4], raw[5]), maj3(raw[6], raw[7], raw[8]));
        const int32_t c = (int32_t)guess;
        if (c >= 0) {
            M eq = (c == none_code) ? nones : M(0);
#pragma unroll
            for (int j = 0; j < N; ++j) eq |= (raw[j] == c) ? (M(1) << j) : M(0);
            const int cnt = popc_m(eq);
            if (2 * cnt > voters) {
                win_code = c;
                meta = pack_meta(ffs_mask(eq) - 1, cnt, voters, present, KC_FLAG_HAS_VALUE);
                return;
            }
        }
    }
    // absent cells (< -1) are never live and never equal a code >= 0, so the scan can read raw[] as is
    meta = vote_scan<N, M>(raw, live, nones, none_code, present, win_code);
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
