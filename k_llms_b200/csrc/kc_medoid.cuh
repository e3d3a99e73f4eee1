// kc_medoid.cuh — K4: similarity medoid of multi-word strings.
//
// Replaces, for groups of strings, the fallback of consensus_as_primitive (reference consensus_utils.py:1221-1237):
// pairwise levenshtein_similarity (cu:745-761: 1 - dist/max_len on normalize_string()ed text, floored at 1e-8),
// np.nanmean of every row of the n x n matrix (diagonal NaN), first argmax.  One WARP per group:
//   * lanes split the n(n-1)/2 pairs; each distance is Myers' bit-parallel algorithm (one 64-bit word, O(text) word
//     operations) with the SHORTER string as the pattern.  The caller guarantees min(len_i, len_j) <= 64 for every pair
//     (always true under the reference's default method, which sends pairs of two long strings to the embeddings
//     service instead, cu:813); groups that violate it stay on the host;
//   * lane i then sums row i in numpy's pairwise order (diagonal contributes +0.0, as nanmean's NaN->0 copy does) and
//     divides by the n-1 valid entries; a shuffle reduction picks the first maximum.
// Strings arrive normalised (lower-case [a-z0-9]); the host does normalize_string() and the final round(pvf*avg, 5).
#pragma once

#include "kc_common.cuh"
#include "kc_numeric.cuh"  // np_sum

namespace kc {

constexpr int kMedoidMaxN = 64;       // strings per group
constexpr int kMedoidMaxPattern = 64;  // the shorter string of every pair must fit one 64-bit word
constexpr int kAlphabet = 36;

__device__ __forceinline__ int alnum_index(uint8_t c) { return c <= '9' ? c - '0' : c - 'a' + 10; }

// Edit distance of pattern p (m <= 64 chars, Peq table given) against text t (Myers 1999 / Hyyro 2003).
__device__ __forceinline__ int myers64(const uint64_t *peq, int m, const uint8_t *t, int tn) {
    uint64_t pv = ~0ull, mv = 0;
    int score = m;
    const uint64_t top = 1ull << (m - 1);
    for (int k = 0; k < tn; ++k) {
        const uint64_t eq = peq[alnum_index(t[k])];
        const uint64_t xv = eq | mv;
        const uint64_t xh = (((eq & pv) + pv) ^ pv) | eq;
        uint64_t ph = mv | ~(xh | pv);
        uint64_t mh = pv & xh;
        score += (ph & top) ? 1 : 0;
        score -= (mh & top) ? 1 : 0;
        ph = (ph << 1) | 1ull;
        mh <<= 1;
        pv = mh | ~(xv | ph);
        mv = ph & xv;
    }
    return score;
}

// chars: all strings back to back; str_off[s]..str_off[s+1] string s; grp_off[g]..grp_off[g+1] the strings of group g.
// Dynamic shared memory per warp: a Peq table per LANE (36 u64 = 288 B -> 9 KB per warp, rebuilt for every pair by the
// lane that owns it) and the distance matrix (u16 [64][64] = 8 KB).
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) medoid_kernel(const uint8_t *__restrict__ chars, const int32_t *__restrict__ str_off,
                                                            const int32_t *__restrict__ grp_off, int64_t n_groups,
                                                            int32_t *__restrict__ best_idx, double *__restrict__ best_avg) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr size_t PEQ_BYTES = 32 * kAlphabet * 8;               // per warp
    constexpr size_t DIST_BYTES = kMedoidMaxN * kMedoidMaxN * 2;   // per warp
    uint8_t *wbase = smem_raw + (size_t)warp * (PEQ_BYTES + DIST_BYTES);
    uint64_t *peq = reinterpret_cast<uint64_t *>(wbase) + lane * kAlphabet;
    uint16_t *dist = reinterpret_cast<uint16_t *>(wbase + PEQ_BYTES);

    const int64_t gw = (int64_t)blockIdx.x * WARPS + warp;
    const int64_t gstep = (int64_t)gridDim.x * WARPS;
    for (int64_t g = gw; g < n_groups; g += gstep) {
        const int s0 = __ldg(grp_off + g), k = __ldg(grp_off + g + 1) - s0;
        const int n_pairs = k * (k - 1) / 2;
        for (int p = lane; p < n_pairs; p += 32) {
            // (i, j), i < j, from the linear index over the upper triangle
            int i = 0, rem = p;
            while (rem >= k - 1 - i) {
                rem -= k - 1 - i;
                ++i;
            }
            const int j = i + 1 + rem;
            int ao = __ldg(str_off + s0 + i), al = __ldg(str_off + s0 + i + 1) - ao;
            int bo = __ldg(str_off + s0 + j), bl = __ldg(str_off + s0 + j + 1) - bo;
            if (al > bl) {  // pattern = the shorter string
                const int to = ao, tl = al;
                ao = bo; al = bl; bo = to; bl = tl;
            }
            int d;
            if (al == 0) {
                d = bl;
            } else {  // al <= kMedoidMaxPattern by contract
                for (int c = 0; c < kAlphabet; ++c) peq[c] = 0;
                for (int q = 0; q < al; ++q) peq[alnum_index(__ldg(chars + ao + q))] |= 1ull << q;
                d = myers64(peq, al, chars + bo, bl);
            }
            dist[i * kMedoidMaxN + j] = (uint16_t)d;
            dist[j * kMedoidMaxN + i] = (uint16_t)d;
        }
        __syncwarp();
        double my_avg = -1.0;
        int my_idx = 0x7FFFFFFF;
        for (int i = lane; i < k; i += 32) {
            const int li = __ldg(str_off + s0 + i + 1) - __ldg(str_off + s0 + i);
            const double tot = np_sum(
                [&](int j) {
                    if (j == i) return 0.0;  // nanmean works on a copy with NaN -> 0
                    const int lj = __ldg(str_off + s0 + j + 1) - __ldg(str_off + s0 + j);
                    const int mx = max(li, lj);
                    if (mx == 0) return 1.0;  // cu:756-757
                    const double s = __dadd_rn(1.0, -__ddiv_rn((double)dist[i * kMedoidMaxN + j], (double)mx));
                    return s > 1e-8 ? s : 1e-8;  // cu:761
                },
                k);
            const double avg = __ddiv_rn(tot, (double)(k - 1));
            if (avg > my_avg) {  // first maximum within this lane's rows (ascending i)
                my_avg = avg;
                my_idx = i;
            }
        }
#pragma unroll
        for (int st = 16; st >= 1; st >>= 1) {
            const double oa = __shfl_xor_sync(0xFFFFFFFFu, my_avg, st);
            const int oi = __shfl_xor_sync(0xFFFFFFFFu, my_idx, st);
            if (oa > my_avg || (oa == my_avg && oi < my_idx)) {
                my_avg = oa;
                my_idx = oi;
            }
        }
        if (lane == 0) {
            best_idx[g] = my_idx;
            best_avg[g] = my_avg;
        }
        __syncwarp();
    }
}

}  // namespace kc
