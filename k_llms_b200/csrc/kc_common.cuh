// kc_common.cuh — sm_100a PTX helpers shared by the consensus kernels: mbarrier, TMA (cp.async.bulk.tensor),
// streaming stores, the packed result word.  No libraries; inline PTX only.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kllms_b200.h"

namespace kc {

constexpr int kMaxN = KC_MAX_CANDIDATES;

__host__ __device__ __forceinline__ uint32_t pack_meta(uint32_t idx, uint32_t support, uint32_t nn, uint32_t present,
                                                       uint32_t flags) {
    return (idx & 0x3Fu) | ((support & 0x7Fu) << 6) | ((nn & 0x7Fu) << 13) | ((present & 0x7Fu) << 20) |
           ((flags & 0x1Fu) << 27);
}

// ---------------------------------------------------------------- shared-memory addresses, mbarrier

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// make mbarrier.init visible to the async (TMA) proxy
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "KC_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra KC_DONE_%=;\n\t"
        "bra KC_WAIT_%=;\n\t"
        "KC_DONE_%=:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Variants taking 32-bit shared-space addresses (no generic->shared conversion in the loop).
__device__ __forceinline__ void mbar_init_a(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_a(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "KC_WAITA_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra KC_DONEA_%=;\n\t"
        "bra KC_WAITA_%=;\n\t"
        "KC_DONEA_%=:\n\t"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// ---------------------------------------------------------------- TMA: 2-D tiled tensor load, global -> swizzled smem

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// evict-first L2 policy for read-once streams
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// Input tiles of a kernel that is NOT bandwidth-bound: normal priority, so that its result lines (also normal) are
// evicted — written back — while it runs.  With evict-first inputs the results outlive the kernel as up to an L2's worth
// of dirty lines and are written back under the next kernel's input stream (measured: K1 after K2 lost 7 %).
__device__ __forceinline__ uint64_t policy_evict_normal() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int32_t c0, int32_t c1, uint64_t *bar,
                                            uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], "
        "[%4], %5;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// 32-bit shared-space addresses + dependency register.
__device__ __forceinline__ void tma_load_2d_a(uint32_t smem_dst, const CUtensorMap *map, int32_t c0, int32_t c1, uint32_t bar,
                                              uint64_t policy, uint32_t dep) {
    asm volatile(
        "{\n\t"
        ".reg .b32 kc_dep;\n\t"
        "mov.b32 kc_dep, %6;\n\t"
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], "
        "[%4], %5;\n\t"
        "}\n" ::"r"(smem_dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar), "l"(policy), "r"(dep)
        : "memory");
}

// ---------------------------------------------------------------- global memory: streaming loads / stores

// read-only path, default L1 allocation: a thread's neighbouring 16-byte pieces share 32-byte sectors, so the
// second piece should hit L1 instead of going back to L2
__device__ __forceinline__ int4 ldg_nc_v4(const void *p) {
    int4 r;
    asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// Result stores.  mc == true: `p` is an NVSwitch MULTICAST address (CUDA multicast object mapped on every GPU of
// the group); multimem.st makes the switch replicate the store into each GPU's copy of the buffer — the all-gather
// of the outputs happens inside the producing kernel, tile by tile, instead of in a separate collective.
// Where a kernel's results go.
//   LOCAL     plain stores (L1 no-allocate) to the given addresses;
//   MULTIMEM  the addresses are NVSwitch multicast addresses: multimem.st, the switch replicates every store into all
//             GPUs' copies — also back into the sender's own, so every GPU RECEIVES world x its share;
//   PEERS     the addresses are local and lie in a buffer that n_peers other GPUs map as well (symmetric memory): one
//             local store plus one store per peer at address + delta[k] (P2P over NVLink).  Each GPU receives only
//             (world - 1) shares and sends as many: less ingress than multicast, the better trade on full-duplex links.
//   PEERS_PACKED  (votes) the full result (winning code, result word) stays in LOCAL arrays — the owning rank's decoder
//             needs the first-seen index in it — and ONE packed word code:18 | support:7 | present:7, everything a remote
//             consumer needs for the value and its confidence, goes to `packed` (local + peers): 4 instead of 8 bytes per
//             vote field over NVLink.  A winning code >= 2^18 sets *overflow (the caller then falls back to PEERS).
struct OutRoute {
    uint32_t mode;       // KC_OUT_LOCAL / KC_OUT_MULTIMEM / KC_OUT_PEERS / 3 = PEERS_PACKED / 4 = WIRE (full result + local wire word, kc_push.cuh)
    int32_t n_peers;     // PEERS*, only
    long long delta[7];  // byte offsets local address -> the same address in peer k's mapping
    uint32_t *packed;    // PEERS_PACKED: local address (inside the shared buffer) of the packed vote words
    uint32_t *overflow;  // PEERS_PACKED / WIRE: device flag
    uint32_t wire_wide;  // WIRE (mode 4): 0 = u16 words code:6|support:5|present:5, 1 = u32 words code:18|support:7|present:7
    __host__ __device__ bool local() const { return mode == 0; }
};

// the P2P copies, out of line: the local path pays one uniform compare per store for them
__device__ __noinline__ void store_peers_u32(void *p, uint32_t v, const OutRoute &r) {
    for (int k = 0; k < r.n_peers; ++k)
        asm volatile("st.global.u32 [%0], %1;" ::"l"(reinterpret_cast<char *>(p) + r.delta[k]), "r"(v) : "memory");
}
__device__ __noinline__ void store_peers_f64(void *p, double v, const OutRoute &r) {
    for (int k = 0; k < r.n_peers; ++k)
        asm volatile("st.global.f64 [%0], %1;" ::"l"(reinterpret_cast<char *>(p) + r.delta[k]), "d"(v) : "memory");
}

__device__ __forceinline__ void store_local_u32(void *p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void store_local_f64(void *p, double v) {
    asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

__device__ __forceinline__ void store_out_u32(void *p, uint32_t v, const OutRoute &r) {
    if (r.mode == 1u) {
        asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    } else {
        asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
        if (r.mode >= 2u) store_peers_u32(p, v, r);
    }
}
// K1's two result words of group g, routed
__device__ __forceinline__ void store_vote_result(int32_t *win, uint32_t *meta, int64_t g, int32_t w, uint32_t m, const OutRoute &r) {
    if (r.mode == 4u) {  // full result locally + the wire word in this rank's slot (a push kernel replicates the slot)
        store_local_u32(win + g, (uint32_t)w);
        store_local_u32(meta + g, m);
        const uint32_t support = (m >> 6) & 0x7Fu, present = (m >> 20) & 0x7Fu;
        if (r.wire_wide) {
            if (support != 0 && (uint32_t)w >= (1u << 18)) atomicOr(r.overflow, 1u);
            const uint32_t word = ((uint32_t)w & 0x3FFFFu) | (support << 18) | (present << 25);
            store_local_u32(r.packed + g, word);
            if (r.n_peers) store_peers_u32(r.packed + g, word, r);
        } else {
            if (support > 31u || present > 31u || (support != 0 && (uint32_t)w > 63u)) atomicOr(r.overflow, 1u);
            const uint16_t word = (uint16_t)(((uint32_t)w & 63u) | ((support & 31u) << 6) | ((present & 31u) << 11));
            uint16_t *p = reinterpret_cast<uint16_t *>(r.packed) + g;
            asm volatile("st.global.L1::no_allocate.u16 [%0], %1;" ::"l"(p), "h"(word) : "memory");
            for (int k = 0; k < r.n_peers; ++k)  // fused reassembly: a warp's 32 words are one 64-byte store per peer
                asm volatile("st.global.u16 [%0], %1;" ::"l"(reinterpret_cast<char *>(p) + r.delta[k]), "h"(word) : "memory");
        }
        return;
    }
    if (r.mode != 3u) {
        store_out_u32(win + g, (uint32_t)w, r);
        store_out_u32(meta + g, m, r);
        return;
    }
    store_local_u32(win + g, (uint32_t)w);
    store_local_u32(meta + g, m);
    const uint32_t support = (m >> 6) & 0x7Fu, present = (m >> 20) & 0x7Fu;
    if (support != 0 && (uint32_t)w >= (1u << 18)) atomicOr(r.overflow, 1u);
    const uint32_t word = ((uint32_t)w & 0x3FFFFu) | (support << 18) | (present << 25);
    store_local_u32(r.packed + g, word);
    store_peers_u32(r.packed + g, word, r);
}

__device__ __forceinline__ void store_out_f64(void *p, double v, const OutRoute &r) {
    if (r.mode == 1u) {
        asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
    } else {
        asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
        if (r.mode >= 2u) store_peers_f64(p, v, r);
    }
}

__device__ __forceinline__ int4 lds_v4(uint32_t addr) {
    int4 r;
    asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}

}  // namespace kc
