// kc_extra.cuh — confidences (exact Python round(x, 5)) and K3, the per-candidate logprob sum.
#pragma once

#include "kc_common.cuh"
#include "kc_vote.cuh"

namespace kc {

// CPython float.__round__(x, 5) for finite x >= 0 (Objects/floatobject.c double_round: dtoa mode 3, i.e. the
// EXACT binary value rounded half-even to 5 decimals, then strtod).  x = M * 2^-sh exactly; q = x*1e5 rounded
// half-even in 128-bit integer arithmetic; q / 1e5 in IEEE double is the double nearest to the decimal
// q*10^-5, which is what strtod returns.  Used by cu:982,1178,1187,1219.
__device__ __forceinline__ double py_round5(double x) {
    const uint64_t bits = (uint64_t)__double_as_longlong(x);
    const int biased = (int)((bits >> 52) & 0x7FF);
    uint64_t mant = bits & 0xFFFFFFFFFFFFFull;
    int exp2;  // x = mant * 2^exp2
    if (biased == 0) {
        exp2 = -1074;
    } else {
        mant |= 1ull << 52;
        exp2 = biased - 1075;
    }
    if (mant == 0) return x;
    if (exp2 >= 0) return x;  // integer-valued: nothing to round
    const int sh = -exp2;     // >= 1
    if (sh >= 128) return 0.0;
    const uint64_t lo = mant * 100000ull;  // low 64 bits of the 70-bit product
    const uint64_t hi = __umul64hi(mant, 100000ull);
    uint64_t q, rem_hi, rem_lo, half_hi, half_lo;
    if (sh >= 64) {
        const int s = sh - 64;  // 0..63
        q = s == 0 ? hi : (hi >> s);
        rem_hi = s == 0 ? 0 : (hi & ((1ull << s) - 1));
        rem_lo = lo;
        half_hi = s == 0 ? 0 : (1ull << (s - 1));
        half_lo = s == 0 ? (1ull << 63) : 0;
    } else {
        q = (hi << (64 - sh)) | (lo >> sh);  // sh in 1..63; hi < 2^6 so no bits are lost for sh >= 6,
                                             // and for sh < 6 the product fits 64 bits only if hi == 0 (x >= 2^46: not a confidence)
        rem_hi = 0;
        rem_lo = lo & ((1ull << sh) - 1);
        half_hi = 0;
        half_lo = 1ull << (sh - 1);
    }
    const bool gt = rem_hi > half_hi || (rem_hi == half_hi && rem_lo > half_lo);
    const bool eq = rem_hi == half_hi && rem_lo == half_lo;
    if (gt || (eq && (q & 1))) ++q;
    return __ddiv_rn((double)q, 100000.0);
}

// One thread per group: the confidence the reference attaches to the group's consensus value.
__global__ void __launch_bounds__(256) confidence_kernel(const uint32_t *__restrict__ meta, int64_t n_groups, bool numeric,
                                                         const double *__restrict__ pvf_in, double *__restrict__ conf) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += stride) {
        const uint32_t m = __ldg(meta + g);
        const double pvf = pvf_in ? __ldg(pvf_in + g) : 1.0;
        const double support = (double)KC_META_SUPPORT(m), nn = (double)KC_META_NN(m), present = (double)KC_META_PRESENT(m);
        const uint32_t flags = KC_META_FLAGS(m);
        double c;
        if (flags & KC_FLAG_HAS_VALUE) {
            if (!numeric)
                c = py_round5(__dmul_rn(pvf, __ddiv_rn(support, present)));  // cu:973,982
            else if (flags & KC_FLAG_SINGLE)
                c = __dmul_rn(__dmul_rn(pvf, __ddiv_rn(1.0, present)), 1.0);  // cu:1444,1086 (unrounded)
            else
                c = py_round5(__ddiv_rn(support, nn));  // cu:1177-1178,1186-1187,1218-1219
        } else if (flags & KC_FLAG_NO_FINITE) {
            c = __dmul_rn(pvf, __ddiv_rn(nn, present));  // cu:1444,1116
        } else {
            c = present == 0.0 ? pvf : 0.0;  // cu:1396 / cu:1402
        }
        conf[g] = c;
    }
}

// K3: one warp per sequence.  Lane l adds elements l, l+32, ... (coalesced 128-byte warp loads), then a
// xor butterfly 16,8,4,2,1 — the fixed order include/kllms_b200.h documents and oracle/consensus_oracle.c restates.
__global__ void __launch_bounds__(256) logprob_sum_kernel(const float *__restrict__ lp, const int64_t *__restrict__ offsets,
                                                          int64_t n_seq, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s = warp; s < n_seq; s += n_warps) {
        const int64_t b = __ldg(offsets + s), e = __ldg(offsets + s + 1);
        float acc = 0.0f;
        for (int64_t i = b + lane; i < e; i += 32) acc = __fadd_rn(acc, __ldg(lp + i));
#pragma unroll
        for (int st = 16; st >= 1; st >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xFFFFFFFFu, acc, st));
        if (lane == 0) out[s] = acc;
    }
}

// ---------------------------------------------------------------- K3b: likelihood-weighted vote (self-defined spec)

// exp(x) for x <= 0, built only from single IEEE fp32 operations in a fixed order so that the C oracle
// (oracle/consensus_oracle.c: ko_exp_f32) reproduces it bit for bit: t = x*log2(e); k = floor(t + 0.5); f = t - k;
// 2^f by a degree-5 polynomial (Horner, separate multiply and add); scale by 2^k through the exponent field.
__device__ __forceinline__ float kexp(float x) {
    x = x < -87.0f ? -87.0f : x;
    const float t = __fmul_rn(x, 1.44269504f);
    const float k = floorf(__fadd_rn(t, 0.5f));
    const float f = __fadd_rn(t, -k);
    float p = 0.00133336f;
    p = __fadd_rn(__fmul_rn(p, f), 0.00961813f);
    p = __fadd_rn(__fmul_rn(p, f), 0.05550411f);
    p = __fadd_rn(__fmul_rn(p, f), 0.24022651f);
    p = __fadd_rn(__fmul_rn(p, f), 0.69314718f);
    p = __fadd_rn(__fmul_rn(p, f), 1.0f);
    return __fmul_rn(p, __int_as_float(((int)k + 127) << 23));
}

// One thread per group.  Candidate weight w_c = kexp(s_c - max_k s_k) with s = seq_logprob[record]; class weight =
// sum of its cells' weights in ascending candidate order (fp32); the heaviest class wins, ties -> first seen.
template <int NP>
__global__ void __launch_bounds__(128) weighted_vote_kernel(const int32_t *__restrict__ codes, const float *__restrict__ seq_lp,
                                                            int64_t n_groups, int n, FieldMap fm, bool has_nc,
                                                            int32_t *__restrict__ win, uint32_t *__restrict__ meta,
                                                            float *__restrict__ weight) {
    using M = typename MaskOf<NP>::type;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += stride) {
        const int64_t rec = g / fm.n_fields;
        const uint32_t field = (uint32_t)(g - rec * fm.n_fields);
        const int32_t nc = has_nc ? __ldg(fm.none_code + field) : KC_CODE_NONE;
        int32_t x[NP];
        float w[NP];
        float smax = -3.0e38f;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            int32_t c = (i < n) ? __ldg(codes + g * n + i) : KC_CODE_ABSENT;
            const float s = (i < n) ? __ldg(seq_lp + rec * n + i) : -3.0e38f;
            w[i] = s;
            smax = (i < n && s > smax) ? s : smax;
            c = (c == KC_CODE_NONE) ? nc : c;           // None votes as none_code where it is >= 0
            x[i] = c < KC_CODE_NONE ? KC_CODE_NONE : c;  // absent cells never vote
        }
        M live = 0;
        int present = 0;
        float total = 0.0f;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            w[i] = kexp(__fadd_rn(w[i], -smax));
            const bool absent = (i >= n) || __ldg(codes + g * n + (i < n ? i : 0)) < KC_CODE_NONE;
            present += absent ? 0 : 1;
            if (x[i] >= 0) {
                live |= M(1) << i;
                total = __fadd_rn(total, w[i]);
            }
        }
        const int voters = popc_m(live);
        float best_w = -1.0f;
        int best_idx = 0, best_cnt = 0;
        int32_t best_code = KC_CODE_NONE;
        bool tie = false;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            if ((live >> i) & 1) {
                const int32_t c = x[i];
                M eq = 0;
                float cw = 0.0f;
#pragma unroll
                for (int j = i; j < NP; ++j) {
                    const bool e = x[j] == c;
                    eq |= e ? (M(1) << j) : M(0);
                    cw = e ? __fadd_rn(cw, w[j]) : cw;
                }
                if (cw > best_w) {
                    best_w = cw;
                    best_idx = i;
                    best_cnt = popc_m(eq);
                    best_code = c;
                    tie = false;
                } else if (cw == best_w) {
                    tie = true;
                }
                live &= ~eq;
            }
        }
        win[g] = best_code;
        meta[g] = pack_meta(best_idx, best_cnt, voters, present,
                            voters > 0 ? (KC_FLAG_HAS_VALUE | (tie ? KC_FLAG_TIE : 0u)) : 0u);
        weight[g] = voters > 0 ? __fdiv_rn(best_w, total) : 0.0f;
    }
}

}  // namespace kc
