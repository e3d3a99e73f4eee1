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

// K3, staged: sequences of a few dozen tokens leave a warp-per-sequence kernel with one or two elements per lane and
// ~25 instructions of bookkeeping per sequence (measured 0.36 of HBM peak on config 4).  Here a CTA copies the CONTIGUOUS
// token range of its T sequences into shared memory with 16-byte cp.async (coalesced, no register staging), then every
// thread sums ONE sequence in exactly the documented order: partial l = elements l, l+32, ... added left to right from
// +0.0f, then the tree the xor butterfly 16,8,4,2,1 forms (lane 0's view of it).  A tile whose tokens do not fit CAP
// floats (long sequences) falls back to the warp-per-sequence loop.
template <int T, int CAP>
__global__ void __launch_bounds__(T) logprob_sum_tile_kernel(const float *__restrict__ lp, const int64_t *__restrict__ offsets,
                                                             int64_t n_seq, float *__restrict__ out) {
    extern __shared__ __align__(16) float tok[];  // CAP + 4 floats
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n_tiles = (n_seq + T - 1) / T;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t s0 = tile * T;
        const int cnt = (int)min((int64_t)T, n_seq - s0);
        // all four offset loads in flight together: the tile's range and this thread's sequence
        const int64_t t0 = __ldg(offsets + s0), t1 = __ldg(offsets + s0 + cnt);
        const int64_t my_b = __ldg(offsets + s0 + (tid < cnt ? tid : 0)), my_e = __ldg(offsets + s0 + (tid < cnt ? tid + 1 : 0));
        const int64_t t0a = t0 & ~int64_t(3);  // the copy starts at the 16-byte boundary at or below t0
        const int64_t span = t1 - t0a;
        if (span <= CAP) {
            const int n_vec = (int)(span >> 2);  // whole 16-byte pieces inside [t0a, t1)
            const uint32_t dst = smem_u32(tok);
            for (int v = tid; v < n_vec; v += T)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)v * 16u), "l"(lp + t0a + (int64_t)v * 4) : "memory");
            for (int i = (n_vec << 2) + tid; i < (int)span; i += T) tok[i] = __ldg(lp + t0a + i);  // < 4 stragglers
            asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            if (tid < cnt) {
                const int64_t b = my_b, e = my_e;
                const float *x = tok + (b - t0a);
                const int len = (int)(e - b);
                float p[32];
#pragma unroll
                for (int l = 0; l < 32; ++l) p[l] = 0.0f;
                for (int k = 0; k < len; k += 32) {
#pragma unroll
                    for (int l = 0; l < 32; ++l)
                        if (k + l < len) p[l] = __fadd_rn(p[l], x[k + l]);
                }
#pragma unroll
                for (int st = 16; st >= 1; st >>= 1) {
#pragma unroll
                    for (int l = 0; l < st; ++l) p[l] = __fadd_rn(p[l], p[l + st]);
                }
                out[s0 + tid] = p[0];
            }
            __syncthreads();  // the tile buffer is reused
        } else {
            for (int q = warp; q < cnt; q += T / 32) {
                const int64_t b = __ldg(offsets + s0 + q), e = __ldg(offsets + s0 + q + 1);
                float acc = 0.0f;
                for (int64_t i = b + lane; i < e; i += 32) acc = __fadd_rn(acc, __ldg(lp + i));
#pragma unroll
                for (int st = 16; st >= 1; st >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xFFFFFFFFu, acc, st));
                if (lane == 0) out[s0 + q] = acc;
            }
        }
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

// K3b, weights once per record: the candidate weights depend on the record only, not on the field.  A CTA of T threads
// covers T consecutive groups (fields of a handful of records); its warps first compute w_c = kexp(s_c - max s) for
// those records into shared memory (lane = candidate, warp max by shuffles), then every thread votes its group with the
// weights read from there (same class sums, same order as weighted_vote_kernel).
template <int NP, int T>
__global__ void __launch_bounds__(T) weighted_vote_rec_kernel(const int32_t *__restrict__ codes, const float *__restrict__ seq_lp,
                                                              int64_t n_groups, int n, FieldMap fm, bool has_nc,
                                                              int32_t *__restrict__ win, uint32_t *__restrict__ meta,
                                                              float *__restrict__ weight) {
    using M = typename MaskOf<NP>::type;
    extern __shared__ float wts[];  // [records of the tile][NP]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n_tiles = (n_groups + T - 1) / T;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t g0 = tile * T, g1 = min(g0 + T, n_groups);
        const int64_t r0 = g0 / fm.n_fields, r1 = (g1 - 1) / fm.n_fields;
        for (int64_t r = r0 + warp; r <= r1; r += T / 32) {
            float s_lo = (lane < n) ? __ldg(seq_lp + r * n + lane) : -3.0e38f;
            float s_hi = (NP > 32 && lane + 32 < n) ? __ldg(seq_lp + r * n + lane + 32) : -3.0e38f;
            float smax = fmaxf(s_lo, s_hi);
#pragma unroll
            for (int st = 16; st >= 1; st >>= 1) smax = fmaxf(smax, __shfl_xor_sync(0xFFFFFFFFu, smax, st));
            float *w = wts + (r - r0) * NP;
            if (lane < NP) w[lane] = kexp(__fadd_rn(s_lo, -smax));
            if (NP > 32) w[lane + 32] = kexp(__fadd_rn(s_hi, -smax));
        }
        __syncthreads();
        const int64_t g = g0 + tid;
        if (g < g1) {
            // record and field of this thread without a 64-bit division: offset inside the tile's first record
            const uint32_t fpos = (uint32_t)(g0 - r0 * fm.n_fields) + (uint32_t)tid;
            const uint32_t rec_local = fm.div_small(fpos);
            const uint32_t field = fpos - rec_local * fm.n_fields;
            const int32_t nc = has_nc ? __ldg(fm.none_code + field) : KC_CODE_NONE;
            const float *w = wts + rec_local * NP;
            // (requesting the row BEFORE the weights phase was tried: it keeps 32 more registers live across the barrier,
            // 67 -> 94, and the lost occupancy cost more than the overlap gained: 0.357 -> 0.379 ms)
            int32_t x[NP], rawrow[NP];
            if (n == NP) load_row<NP, true>(codes, g, n, rawrow);
            else load_row<NP, false>(codes, g, n, rawrow);
            M live = 0;
            int present = 0;
            float total = 0.0f, wmax = -1.0f;
            int32_t cmax = KC_CODE_NONE;  // the code of this group's heaviest VOTING cell (first of equals)
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                const int32_t raw = rawrow[i];
                int32_t c = (raw == KC_CODE_NONE) ? nc : raw;  // None votes as none_code where it is >= 0
                c = c < KC_CODE_NONE ? KC_CODE_NONE : c;        // absent cells never vote
                x[i] = c;
                present += raw < KC_CODE_NONE ? 0 : 1;
                if (c >= 0) {
                    live |= M(1) << i;
                    const float wi = w[i];
                    total = __fadd_rn(total, wi);
                    if (wi > wmax) {
                        wmax = wi;
                        cmax = c;
                    }
                }
            }
            const int voters = popc_m(live);
            float best_w = -1.0f;
            int best_idx = 0, best_cnt = 0;
            int32_t best_code = KC_CODE_NONE;
            bool tie = false;
            float consumed = 0.0f;
            // Order of the walk: always the class of the heaviest cell still waiting (the first one: this group's heaviest
            // voting cell).  The outcome does not depend on the order (the heaviest class wins, equal weights go to the class
            // seen first = smaller first index, `tie` says whether another class equals the winner), but the COST does: a warp
            // loops until its slowest lane is done.  The heaviest voter's class usually holds more than half of the weight and
            // ends the walk at once; where it does not, going by weight makes the remainder shrink fastest.  (Round 1 walked
            // in first-seen order after the record's heaviest candidate: 4.3 class passes per warp, and it parked every
            // group's codes in a shared-memory plane to fetch the next class's code by a data-dependent index; here the pass
            // over the cells finds the next class itself — no shared memory for the codes at all.)
            int32_t c = cmax;
            while (live) {
                // Every class still waiting sums a subset of the unconsumed weights, so its fp32 sum is at most
                // (total - consumed) up to rounding (< 32 * 2^-23 relative on each side); 5e-5 * total is a safe slack.
                // Below best_w it can neither win nor tie: stop.
                if (__fadd_rn(__fadd_rn(total, -consumed), __fmul_rn(total, 5e-5f)) < best_w) break;
                M eq = 0;
                float cw = 0.0f, nw = -1.0f;
                int32_t nc2 = KC_CODE_NONE;
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const bool e = x[j] == c;  // cells with this code that came earlier were consumed with their class
                    const float wj = w[j];
                    eq |= e ? (M(1) << j) : M(0);
                    cw = e ? __fadd_rn(cw, wj) : cw;
                    const bool waiting = !e && ((live >> j) & 1);
                    if (waiting && wj > nw) {  // heaviest cell of the classes still waiting (first of equals)
                        nw = wj;
                        nc2 = x[j];
                    }
                }
                const int i = ffs_mask(eq) - 1;  // the class's first cell
                if (cw > best_w) {
                    best_w = cw;
                    best_idx = i;
                    best_cnt = popc_m(eq);
                    best_code = c;
                    tie = false;
                } else if (cw == best_w) {
                    tie = true;
                    if (i < best_idx) {  // an equally heavy class that was seen earlier
                        best_idx = i;
                        best_cnt = popc_m(eq);
                        best_code = c;
                    }
                }
                consumed = __fadd_rn(consumed, cw);
                live &= ~eq;
                c = nc2;
            }
            win[g] = best_code;
            meta[g] = pack_meta(best_idx, best_cnt, voters, present, voters > 0 ? (KC_FLAG_HAS_VALUE | (tie ? KC_FLAG_TIE : 0u)) : 0u);
            weight[g] = voters > 0 ? __fdiv_rn(best_w, total) : 0.0f;
        }
        __syncthreads();
    }
}

}  // namespace kc
