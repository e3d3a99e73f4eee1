// kc_push.cuh — reassembly of a sharded batch over NVLink: pack this rank's results into the WIRE format and replicate them
// into every peer's copy of the gathered buffer with 16-byte stores (P2P over NVLink / NVSwitch, symmetric memory).
//
// Round 1 pushed every result from inside K1 / K2 as a 4- or 8-byte scalar store per thread and peer (192 B per record):
// at 8 GPUs the step was bound by NVLink ingress (613 GB/s of a ~770 GB/s link), at 2 GPUs by store issue, and the routed
// launches had to use the slower general K2 kernel.  Here the compute kernels keep their fast local-store paths; this kernel
// runs on a second stream for chunk c while chunk c + 1 is computed, reads chunk c's results (L2-resident) and writes
//     vote     u16  code:6 | support:5 | present:5        (n <= 31, codes < 64; else u32 code:18 | support:7 | present:7)
//     numeric  f64 value + u16 kind:2 | payload:10         (n <= 31; else the u32 result word)
// = 128 B per S32 record instead of 192 / 288, as whole 16-byte vectors: one st.global.v4 per lane, peer and 8 results.
// A remote consumer gets the winning code / value and everything the confidence needs (support / present, support / nn, ...);
// the first-seen index and the tie flag stay with the owning rank (its decoder needs them, nobody else does).
#pragma once

#include "kc_common.cuh"

namespace kc {

struct PushArgs {
    const int32_t *win;     // K1 results of the chunk (local, full); NULL: K1 already wrote the wire words (kc_vote_i32_wire)
    const uint32_t *vmeta;
    int64_t gv;
    const double *value;    // K2 results of the chunk (local, full); == wire_value: K2 wrote its values straight into the slot
    const uint32_t *nmeta;
    int64_t gx;
    uint8_t *wire_votes;    // local addresses of the chunk's part of this rank's slot in the gathered buffer
    uint8_t *wire_value;
    uint8_t *wire_nmeta;
    int32_t n_peers;
    long long delta[7];     // byte distance from the local mapping of the buffer to peer k's
    int32_t wide;           // 0: 16-bit words, 1: 32-bit words
    uint32_t *overflow;     // set when a result does not fit the 16-bit words (caller repeats with wide = 1)
};

__device__ __forceinline__ uint4 ld_v4(const void *p) {
    uint4 r;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_v4(void *p, uint4 v) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_peers(uint8_t *p, uint4 v, const PushArgs &a) {
    for (int k = 0; k < a.n_peers; ++k) st_v4(p + a.delta[k], v);
}
__device__ __forceinline__ void st_all(uint8_t *p, uint4 v, const PushArgs &a) {
    st_v4(p, v);
    st_peers(p, v, a);
}

__host__ __device__ __forceinline__ uint32_t wire_vote16(int32_t win, uint32_t meta, uint32_t &bad) {
    const uint32_t support = (meta >> 6) & 0x7Fu, present = (meta >> 20) & 0x7Fu;
    bad |= (support > 31u || present > 31u || (support != 0 && (uint32_t)win > 63u)) ? 1u : 0u;
    return ((uint32_t)win & 63u) | ((support & 31u) << 6) | ((present & 31u) << 11);
}
__host__ __device__ __forceinline__ uint32_t wire_vote32(int32_t win, uint32_t meta, uint32_t &bad) {
    const uint32_t support = (meta >> 6) & 0x7Fu, present = (meta >> 20) & 0x7Fu;
    bad |= (support != 0 && (uint32_t)win >= (1u << 18)) ? 1u : 0u;
    return ((uint32_t)win & 0x3FFFFu) | (support << 18) | (present << 25);
}
// kind 0: value, confidence round(support / nn, 5)          payload support | nn << 5
// kind 1: the single non-None cell, confidence 1 / present  payload present
// kind 2: no finite value, confidence nn / present          payload nn | present << 5
// kind 3: no value: present == 0 -> empty, else all None    payload present
__host__ __device__ __forceinline__ uint32_t wire_num16(uint32_t meta, uint32_t &bad) {
    const uint32_t support = (meta >> 6) & 0x7Fu, nn = (meta >> 13) & 0x7Fu, present = (meta >> 20) & 0x7Fu, flags = meta >> 27;
    bad |= (support > 31u || nn > 31u || present > 31u) ? 1u : 0u;
    if (flags & KC_FLAG_HAS_VALUE) return (flags & KC_FLAG_SINGLE) ? ((1u << 10) | (present & 31u)) : ((support & 31u) | ((nn & 31u) << 5));
    if (flags & KC_FLAG_NO_FINITE) return (2u << 10) | (nn & 31u) | ((present & 31u) << 5);
    return (3u << 10) | (present & 31u);
}

// One 16-byte output vector per thread and round; the three segments of the chunk are walked with one flat index.
__device__ __forceinline__ void push_chunk(const PushArgs &a, int64_t first, int64_t stride, uint32_t &bad) {
    const int64_t per_v = a.wide ? 4 : 8;                 // vote results per vector
    // segments with nothing to do (already in the slot and no peer to copy to) are skipped
    const int64_t uv = (!a.win && a.n_peers == 0) ? 0 : a.gv / per_v;
    const int64_t ux = (reinterpret_cast<const uint8_t *>(a.value) == a.wire_value && a.n_peers == 0) ? 0 : a.gx / 2;
    const int64_t um = a.wide ? a.gx / 4 : a.gx / 8;
    const int64_t total = uv + ux + um;
    for (int64_t u = first; u < total; u += stride) {
        if (u < uv) {
            uint4 o;
            if (!a.win) {
                st_peers(a.wire_votes + u * 16, ld_v4(a.wire_votes + u * 16), a);
                continue;
            }
            if (a.wide) {
                const uint4 w = ld_v4(a.win + u * 4), m = ld_v4(a.vmeta + u * 4);
                o.x = wire_vote32((int32_t)w.x, m.x, bad);
                o.y = wire_vote32((int32_t)w.y, m.y, bad);
                o.z = wire_vote32((int32_t)w.z, m.z, bad);
                o.w = wire_vote32((int32_t)w.w, m.w, bad);
            } else {
                const uint4 w0 = ld_v4(a.win + u * 8), w1 = ld_v4(a.win + u * 8 + 4);
                const uint4 m0 = ld_v4(a.vmeta + u * 8), m1 = ld_v4(a.vmeta + u * 8 + 4);
                o.x = wire_vote16((int32_t)w0.x, m0.x, bad) | (wire_vote16((int32_t)w0.y, m0.y, bad) << 16);
                o.y = wire_vote16((int32_t)w0.z, m0.z, bad) | (wire_vote16((int32_t)w0.w, m0.w, bad) << 16);
                o.z = wire_vote16((int32_t)w1.x, m1.x, bad) | (wire_vote16((int32_t)w1.y, m1.y, bad) << 16);
                o.w = wire_vote16((int32_t)w1.z, m1.z, bad) | (wire_vote16((int32_t)w1.w, m1.w, bad) << 16);
            }
            st_all(a.wire_votes + u * 16, o, a);
        } else if (u < uv + ux) {
            const int64_t k = u - uv;
            const uint4 v = ld_v4(a.value + k * 2);
            if (reinterpret_cast<const uint8_t *>(a.value) == a.wire_value) st_peers(a.wire_value + k * 16, v, a);
            else st_all(a.wire_value + k * 16, v, a);
        } else {
            const int64_t k = u - uv - ux;
            uint4 o;
            if (a.wide) {
                o = ld_v4(a.nmeta + k * 4);
            } else {
                const uint4 m0 = ld_v4(a.nmeta + k * 8), m1 = ld_v4(a.nmeta + k * 8 + 4);
                o.x = wire_num16(m0.x, bad) | (wire_num16(m0.y, bad) << 16);
                o.y = wire_num16(m0.z, bad) | (wire_num16(m0.w, bad) << 16);
                o.z = wire_num16(m1.x, bad) | (wire_num16(m1.y, bad) << 16);
                o.w = wire_num16(m1.z, bad) | (wire_num16(m1.w, bad) << 16);
            }
            st_all(a.wire_nmeta + k * 16, o, a);
        }
    }
}

__global__ void __launch_bounds__(256) push_kernel(const __grid_constant__ PushArgs a) {
    uint32_t bad = 0;
    push_chunk(a, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x, bad);
    if (bad && a.overflow) atomicOr(a.overflow, 1u);
}

// (A resident variant of this kernel — one small CTA per SM launched before the step, fed by device flags — was built and
// measured in round 2 and removed: CTAs of kernels with different L1 / shared-memory carveouts cannot share an SM, K1 wants the
// L1 and K2 the shared memory, and forcing one carveout on all three cost K1 more (+0.06 ms) than the overlap returned.)

}  // namespace kc
