// kc_numeric.cuh — K2: numeric consensus.  One thread owns one group of n float64 cells.
//
// Replaces the numeric branch of consensus_as_primitive (reference consensus_utils.py:1098-1219):
//   sort -> chain adjacent values closer than max(abs_eps, rel_eps*max(|a|,|b|,1)) into clusters (cu:1127-1144)
//   -> unique largest cluster: float(np.mean(cluster)), support = size            (cu:1171-1187)
//   -> tie: neighbours' support, then (spread, -|center|) ordering                (cu:1189-1219)
// Bit-exactness: np.mean / np.std are numpy pairwise sums (8 accumulators, (r0+r1)+(r2+r3)+..., tail
// left-to-right, reduction seeded with +0.0); every add/mul/div/sqrt below is a single IEEE-754
// round-to-nearest operation in that order (the file is compiled with -fmad=false as well).
//
// The sort is a Batcher merge-exchange network on registers (static indices).  The sorted row is then
// parked in shared memory as scratch[i*T + tid]: for a fixed i a warp touches 32 consecutive doubles, and
// because T*8 is a multiple of 128 B the bank of a word depends on tid only, so the data-dependent indices
// of the cluster walk never cause a bank conflict.
#pragma once

#include <utility>

#include "kc_common.cuh"
#include "kc_vote.cuh"  // MaskOf, Swizzle

namespace kc {

constexpr uint32_t kNoneHi = (uint32_t)(KC_F64_NONE_BITS >> 32);
constexpr uint32_t kNoneLo = (uint32_t)(KC_F64_NONE_BITS & 0xFFFFFFFFu);
constexpr uint32_t kAbsentLo = (uint32_t)(KC_F64_ABSENT_BITS & 0xFFFFFFFFu);

__device__ __forceinline__ int ffs_m(uint32_t m) { return __ffs((int)m); }
__device__ __forceinline__ int ffs_m(uint64_t m) { return __ffsll((long long)m); }

// 10.0**k for k = -6..6 exactly as CPython computes it (correctly rounded decimal literals; checked in
// oracle/gen_golden.py's environment): consensus_utils.py:1156-1157.
__device__ __constant__ double kPow10[13] = {1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 1.0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6};

struct Scratch {  // per-thread view of the [i][tid] scratch plane
    const double *base;
    int stride;
    __device__ __forceinline__ double operator[](int i) const { return base[(size_t)i * stride]; }
};

// numpy DOUBLE_pairwise_sum (n <= 128) over f(0..n-1), then the +0.0 seed of add.reduce.
template <typename F>
__device__ __forceinline__ double np_sum(F f, int n) {
    double res;
    if (n < 8) {
        res = -0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, f(i));
    } else {
        double r0 = f(0), r1 = f(1), r2 = f(2), r3 = f(3), r4 = f(4), r5 = f(5), r6 = f(6), r7 = f(7);
        int i = 8;
        const int n8 = n - (n & 7);
        for (; i < n8; i += 8) {
            r0 = __dadd_rn(r0, f(i + 0));
            r1 = __dadd_rn(r1, f(i + 1));
            r2 = __dadd_rn(r2, f(i + 2));
            r3 = __dadd_rn(r3, f(i + 3));
            r4 = __dadd_rn(r4, f(i + 4));
            r5 = __dadd_rn(r5, f(i + 5));
            r6 = __dadd_rn(r6, f(i + 6));
            r7 = __dadd_rn(r7, f(i + 7));
        }
        res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)), __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
        for (; i < n; ++i) res = __dadd_rn(res, f(i));
    }
    return __dadd_rn(0.0, res);
}

__device__ __forceinline__ double np_mean(const Scratch &xs, int s, int len) {
    return __ddiv_rn(np_sum([&](int i) { return xs[s + i]; }, len), (double)len);
}

__device__ __forceinline__ double np_median(const Scratch &xs, int s, int len) {  // ascending input
    if (len & 1) return __dadd_rn(0.0, __dadd_rn(-0.0, xs[s + len / 2]));
    return __ddiv_rn(__dadd_rn(0.0, __dadd_rn(__dadd_rn(-0.0, xs[s + len / 2 - 1]), xs[s + len / 2])), 2.0);
}

__device__ __forceinline__ double np_std(const Scratch &xs, int s, int len) {
    const double mean = np_mean(xs, s, len);
    const double ss = np_sum(
        [&](int i) {
            const double d = __dadd_rn(xs[s + i], -mean);
            return __dmul_rn(d, d);
        },
        len);
    return __dsqrt_rn(__ddiv_rn(ss, (double)len));
}

__device__ __forceinline__ bool is_close(double a, double b, double rel_eps, double abs_eps) {  // cu:1146-1148
    const double denom = fmax(fmax(fabs(a), fabs(b)), 1.0);
    return fabs(__dadd_rn(a, -b)) <= fmax(abs_eps, __dmul_rn(rel_eps, denom));
}

__device__ __forceinline__ bool is_close_pow10(double a, double b, double rel_eps, double abs_eps) {  // cu:1153-1160
    if (a == 0.0 || b == 0.0) return is_close(a, b, rel_eps, abs_eps);
    for (int k = 0; k < 13; ++k)
        if (is_close(a, __dmul_rn(b, kPow10[k]), rel_eps, abs_eps)) return true;
    return false;
}

// Tie between equally large clusters (cu:1189-1219).  `starts` has one bit per cluster start (< m).
struct TieResult {
    double value;
    int support;
};

template <typename M>
__device__ __noinline__ TieResult numeric_tie(const Scratch xs, M starts, int m, int top, double rel_eps, double abs_eps) {
    int best_s = -1, best_support = 0;
    double best_spread = 0.0, best_center = 0.0;
    M sk = starts;
    while (sk) {
        const int s = ffs_m(sk) - 1;
        sk &= sk - 1;
        const int e = sk ? ffs_m(sk) - 1 : m;
        const int len = e - s;
        if (len != top) continue;
        const double center = np_median(xs, s, len);
        int support = top;
        M so = starts;
        while (so) {
            const int os = ffs_m(so) - 1;
            so &= so - 1;
            const int oe = so ? ffs_m(so) - 1 : m;
            const int olen = oe - os;
            if (olen >= top) continue;  // only strictly smaller clusters lend support (cu:1200)
            const double oc = np_median(xs, os, olen);
            if (is_close(center, oc, rel_eps, abs_eps) || is_close(fabs(center), fabs(oc), rel_eps, abs_eps) ||
                is_close_pow10(center, oc, rel_eps, abs_eps))
                support += olen;
        }
        const double spread = len > 1 ? np_std(xs, s, len) : 0.0;
        const bool better = best_s < 0 || support > best_support ||
                            (support == best_support &&
                             (spread < best_spread || (spread == best_spread && fabs(center) > fabs(best_center))));
        if (better) {
            best_s = s;
            best_support = support;
            best_spread = spread;
            best_center = center;
        }
    }
    return TieResult{np_mean(xs, best_s, top), best_support};
}

// Batcher merge-exchange sorting network, ascending, N a power of two.  The comparator list is built at
// compile time and expanded as a fold over an index_sequence so that every register index is a constant
// (a loop nest with data-free but irregular bounds is not reliably unrolled and would push x[] to local memory).
template <int N>
struct BatcherNet {
    int a[N * N / 2 + 1];
    int b[N * N / 2 + 1];
    int count;
    constexpr BatcherNet() : a{}, b{}, count(0) {
        for (int p = 1; p < N; p *= 2)
            for (int k = p; k >= 1; k /= 2)
                for (int j = k % p; j <= N - 1 - k; j += 2 * k)
                    for (int i = 0; i < k; ++i)
                        if (i + j + k < N && (i + j) / (2 * p) == (i + j + k) / (2 * p)) {
                            a[count] = i + j;
                            b[count] = i + j + k;
                            ++count;
                        }
    }
};

template <int A, int B, int N>
__device__ __forceinline__ void compare_exchange(double (&x)[N]) {
    const double lo = x[A], hi = x[B];
    const bool sw = hi < lo;
    x[A] = sw ? hi : lo;
    x[B] = sw ? lo : hi;
}

template <int N, size_t... I>
__device__ __forceinline__ void sort_network_impl(double (&x)[N], std::index_sequence<I...>) {
    constexpr BatcherNet<N> net{};
    (compare_exchange<net.a[I], net.b[I], N>(x), ...);
}

template <int N>
__device__ __forceinline__ void sort_network(double (&x)[N]) {
    if constexpr (N > 1) {
        constexpr BatcherNet<N> net{};
        sort_network_impl<N>(x, std::make_index_sequence<net.count>{});
    }
}

// x[]: raw cells of one group.  scratch: this thread's column of the [N][T] plane (stride T doubles).
template <int N>
__device__ __forceinline__ void numeric_core(double (&x)[N], double rel_eps, double abs_eps, double *scratch, int stride,
                                             double &value, uint32_t &meta) {
    using M = typename MaskOf<N>::type;
    int present = 0, nn = 0, m = 0, first_nn = 0;
    double single = 0.0;
    const double inf = __longlong_as_double(0x7FF0000000000000LL);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const uint32_t hi = (uint32_t)__double2hiint(x[i]);
        const uint32_t lo = (uint32_t)__double2loint(x[i]);
        const bool tagged = hi == kNoneHi;
        const bool absent = tagged && lo == kAbsentLo;
        const bool none = tagged && lo == kNoneLo;
        const bool counted = !absent && !none;
        const bool finite = counted && ((hi & 0x7FF00000u) != 0x7FF00000u);  // cu:1105-1114
        present += absent ? 0 : 1;
        const bool first_counted = counted && nn == 0;
        first_nn = first_counted ? i : first_nn;
        single = first_counted ? x[i] : single;  // the cell the single-value rule returns
        nn += counted ? 1 : 0;
        m += finite ? 1 : 0;
        x[i] = finite ? x[i] : inf;  // non-numbers sort to the end
    }
    if (nn == 0) {
        value = __longlong_as_double(0x7FF8000000000000LL);
        meta = pack_meta(0, 0, 0, present, 0);
        return;
    }
    if (nn == 1) {  // cu:1085-1086: the original object
        value = single;
        meta = pack_meta(first_nn, 1, 1, present, KC_FLAG_HAS_VALUE | KC_FLAG_SINGLE);
        return;
    }
    if (m == 0) {  // cu:1115-1116
        value = __longlong_as_double(0x7FF8000000000000LL);
        meta = pack_meta(0, 0, nn, present, KC_FLAG_NO_FINITE);
        return;
    }
    sort_network<N>(x);

    // cluster starts: bit i set <=> x[i] opens a cluster (not close to x[i-1]); cu:1130-1143.
    // |b-a| <= max(abs_eps, rel_eps*max(|a|,|b|,1))  <=>  (b-a) <= tol(a) or (b-a) <= tol(b),
    // tol(v) = max(abs_eps, rel_eps*max(|v|,1)): exact because fl(rel_eps * .) is monotone for rel_eps >= 0.
    M starts = 1;
    {
        double tol_prev = fmax(abs_eps, __dmul_rn(rel_eps, fmax(fabs(x[0]), 1.0)));
#pragma unroll
        for (int i = 1; i < N; ++i) {
            const double tol = fmax(abs_eps, __dmul_rn(rel_eps, fmax(fabs(x[i]), 1.0)));
            const double d = __dadd_rn(x[i], -x[i - 1]);
            const bool close = (d <= tol_prev) || (d <= tol);
            starts |= close ? M(0) : (M(1) << i);
            tol_prev = tol;
        }
    }
    if (m < N) starts &= (M(1) << m) - 1;  // drop the +inf tail (m >= 1 here)

#pragma unroll
    for (int i = 0; i < N; ++i) scratch[(size_t)i * stride] = x[i];
    const Scratch xs{scratch, stride};

    int top = 0, n_top = 0, top_s = 0;
    {
        M sk = starts;
        while (sk) {
            const int s = ffs_m(sk) - 1;
            sk &= sk - 1;
            const int e = sk ? ffs_m(sk) - 1 : m;
            const int len = e - s;
            if (len > top) {
                top = len;
                n_top = 1;
                top_s = s;
            } else if (len == top) {
                ++n_top;
            }
        }
    }
    uint32_t flags = KC_FLAG_HAS_VALUE;
    int support = top;
    if (n_top == 1) {
        value = np_mean(xs, top_s, top);  // cu:1174-1178 / 1183-1187
    } else {
        const TieResult tr = numeric_tie<M>(xs, starts, m, top, rel_eps, abs_eps);
        value = tr.value;
        support = tr.support;
        flags |= KC_FLAG_TIE;
    }
    meta = pack_meta(0, support, nn, present, flags);
}

// ---------------------------------------------------------------- direct front-end (any n <= NP)

template <int NP, int T>
__global__ void __launch_bounds__(T) numeric_direct_kernel(const double *__restrict__ vals, int64_t n_groups, int n,
                                                           double rel_eps, double abs_eps, double *__restrict__ out_value,
                                                           uint32_t *__restrict__ out_meta) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    double *scratch = reinterpret_cast<double *>(smem_raw) + threadIdx.x;
    const double absent = __longlong_as_double((long long)KC_F64_ABSENT_BITS);
    const int64_t stride = (int64_t)gridDim.x * T;
    for (int64_t g = (int64_t)blockIdx.x * T + threadIdx.x; g < n_groups; g += stride) {
        double x[NP];
        const double *p = vals + g * n;
        if (n == NP && NP >= 2) {
            const int4 *p4 = reinterpret_cast<const int4 *>(p);
#pragma unroll
            for (int q = 0; q < NP / 2; ++q) {
                const int4 t = ldg_stream_v4(p4 + q);
                x[2 * q + 0] = __hiloint2double(t.y, t.x);
                x[2 * q + 1] = __hiloint2double(t.w, t.z);
            }
        } else {
#pragma unroll
            for (int i = 0; i < NP; ++i) x[i] = (i < n) ? __ldg(p + i) : absent;
        }
        double v;
        uint32_t m;
        numeric_core<NP>(x, rel_eps, abs_eps, scratch, T, v, m);
        stg_stream_f64(out_value + g, v);
        stg_stream_u32(out_meta + g, m);
    }
}

// ---------------------------------------------------------------- TMA front-end (n in {4,8,16,32})

// Same ring as vote_tma_kernel; rows are n*8 bytes.  Box rows are at most 128 B wide (the widest TMA
// swizzle span), so a 256 B row (n = 32) is two box rows.
template <int N, int TILE, int STAGES>
__global__ void __launch_bounds__(TILE) numeric_tma_kernel(const __grid_constant__ CUtensorMap tmap, int64_t n_groups,
                                                           double rel_eps, double abs_eps, double *__restrict__ out_value,
                                                           uint32_t *__restrict__ out_meta) {
    constexpr int ROW_BYTES = N * 8;
    constexpr int BOX_ROWS_PER_GROUP = ROW_BYTES > 128 ? ROW_BYTES / 128 : 1;
    constexpr uint32_t STAGE_BYTES = TILE * ROW_BYTES;
    static_assert(STAGE_BYTES % 1024 == 0, "stage must keep the 1024-byte swizzle alignment");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    double *scratch = reinterpret_cast<double *>(smem + (size_t)STAGES * STAGE_BYTES) + threadIdx.x;
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
        double x[N];
        const uint32_t base = smem_u32(smem + (size_t)stage * STAGE_BYTES);
        const uint32_t row_off = (uint32_t)tid * ROW_BYTES;
#pragma unroll
        for (int q = 0; q < N / 2; ++q) {
            const int4 v4 = lds_v4(base + Swizzle<ROW_BYTES>::apply(row_off + q * 16));
            x[2 * q + 0] = __hiloint2double(v4.y, v4.x);
            x[2 * q + 1] = __hiloint2double(v4.w, v4.z);
        }
        __syncthreads();
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
            double v;
            uint32_t m;
            numeric_core<N>(x, rel_eps, abs_eps, scratch, TILE, v, m);
            stg_stream_f64(out_value + g, v);
            stg_stream_u32(out_meta + g, m);
        }
        if (++stage == STAGES) {
            stage = 0;
            parity ^= 1;
        }
    }
}

}  // namespace kc
