/*
 * consensus_oracle.c — columnar CPU oracle for the consensus hot path.
 *
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Built by oracle/Makefile into oracle/_build/libkllms_oracle.so
 * and loaded only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, as the checker.
 *
 * Plain-C restatement of what /root/reference/k_llms/utils/consensus_utils.py ("cu") computes for one
 * scalar field, on the SAME columnar encoding the CUDA kernels read (include/kllms_b200.h):
 *   ko_vote_i32      <- voting_consensus, cu:936-982 (Counter mode, first-seen ties cu:958,969)
 *   ko_numeric_f64   <- consensus_as_primitive numeric branch, cu:1098-1219
 *                       (_cluster_1d cu:1127-1144, tie rules cu:1189-1219)
 *   ko_np_*          <- numpy 2.3 reductions the reference calls at cu:1176,1185,1191,1192,1217:
 *                       pairwise summation with 8 accumulators (numpy/_core/src/umath/loops_utils.h.src,
 *                       third-party, not under /root/reference; checked against the installed numpy
 *                       in tests/test_oracle_columnar.py)
 *
 * Parity pin: no golden vectors exist upstream.  This file is pinned (a) against the
 * reference-generated vectors in tests/golden/ through the object-level oracle
 * (oracle/consensus_py.py, itself fuzzed against the running reference), and (b) against numpy directly.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../include/kllms_b200.h"

#define MAXN KC_MAX_CANDIDATES

/* ------------------------------------------------------------------ numpy reductions (float64) */

/* DOUBLE_pairwise_sum for n <= 128 (PW_BLOCKSIZE); clusters have at most 64 members. */
static double np_pairwise(const double *a, int n) {
    if (n < 8) {
        double res = -0.0; /* numpy starts the small-block loop from -0.0 */
        for (int i = 0; i < n; i++) res += a[i];
        return res;
    }
    double r[8];
    for (int k = 0; k < 8; k++) r[k] = a[k];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int k = 0; k < 8; k++) r[k] += a[i + k];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
}

double ko_np_sum(const double *a, int n) { return 0.0 + np_pairwise(a, n); } /* add.reduce starts at +0.0 */
double ko_np_mean(const double *a, int n) { return ko_np_sum(a, n) / (double)n; }

/* np.median of an ASCENDING array: mean of the middle one or two elements. */
double ko_np_median_sorted(const double *a, int n) {
    if (n & 1) return ko_np_mean(a + n / 2, 1);
    return ko_np_mean(a + n / 2 - 1, 2);
}

/* np.std (ddof=0): sqrt(sum((x - mean)^2) / n), both sums pairwise. */
double ko_np_std(const double *a, int n) {
    double mean = ko_np_mean(a, n);
    double d[MAXN];
    for (int i = 0; i < n; i++) {
        double t = a[i] - mean;
        d[i] = t * t;
    }
    return sqrt(ko_np_sum(d, n) / (double)n);
}

/* ------------------------------------------------------------------ K1 oracle: vote */

void ko_vote_i32(const int32_t *codes, int64_t n_groups, int32_t n, const int32_t *none_code, int32_t n_fields,
                 int32_t *win_code, uint32_t *meta) {
    for (int64_t g = 0; g < n_groups; g++) {
        int32_t nc = (none_code && n_fields > 0) ? none_code[g % n_fields] : -1;
        int32_t cls_code[MAXN];
        int cls_count[MAXN], cls_first[MAXN], n_cls = 0, present = 0, voters = 0;
        for (int c = 0; c < n; c++) {
            int32_t v = codes[g * n + c];
            if (v < KC_CODE_NONE) continue; /* absent */
            present++;
            if (v == KC_CODE_NONE) {
                if (nc < 0) continue; /* None does not vote in string groups (cu:964) */
                v = nc;               /* bool: None -> False (cu:956) */
            }
            voters++;
            int k = 0;
            while (k < n_cls && cls_code[k] != v) k++;
            if (k == n_cls) {
                cls_code[k] = v;
                cls_count[k] = 0;
                cls_first[k] = c;
                n_cls++;
            }
            cls_count[k]++;
        }
        if (voters == 0) {
            win_code[g] = KC_CODE_NONE;
            meta[g] = KC_META_PACK(0, 0, 0, present, 0);
            continue;
        }
        int best = 0, ties = 0;
        for (int k = 1; k < n_cls; k++)
            if (cls_count[k] > cls_count[best]) best = k; /* strict: first seen keeps ties */
        for (int k = 0; k < n_cls; k++)
            if (k != best && cls_count[k] == cls_count[best]) ties = 1;
        win_code[g] = cls_code[best];
        meta[g] = KC_META_PACK(cls_first[best], cls_count[best], voters, present,
                               KC_FLAG_HAS_VALUE | (ties ? KC_FLAG_TIE : 0));
    }
}

/* ------------------------------------------------------------------ K2 oracle: numeric */

static int is_close(double a, double b, double rel_eps, double abs_eps) { /* cu:1146-1148 */
    double denom = fmax(fmax(fabs(a), fabs(b)), 1.0);
    return fabs(a - b) <= fmax(abs_eps, rel_eps * denom);
}

static const double POW10[13] = {1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 1.0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6}; /* 10.0**k */

static int is_close_pow10(double a, double b, double rel_eps, double abs_eps) { /* cu:1153-1160 */
    if (a == 0.0 || b == 0.0) return is_close(a, b, rel_eps, abs_eps);
    for (int k = 0; k < 13; k++)
        if (is_close(a, b * POW10[k], rel_eps, abs_eps)) return 1;
    return 0;
}

void ko_numeric_f64(const double *vals, int64_t n_groups, int32_t n, double rel_eps, double abs_eps, double *value,
                    uint32_t *meta) {
    const double qnan = NAN;
    for (int64_t g = 0; g < n_groups; g++) {
        double xs[MAXN];
        int m = 0, present = 0, nn = 0, first_nn = 0;
        for (int c = 0; c < n; c++) {
            double v = vals[g * n + c];
            uint64_t bits;
            memcpy(&bits, &v, 8);
            if ((bits >> 32) == (KC_F64_ABSENT_BITS >> 32)) continue; /* the high word tags a cell */
            present++;
            if ((bits >> 32) == (KC_F64_NONE_BITS >> 32)) continue;
            if (nn == 0) first_nn = c;
            nn++;
            if (isfinite(v)) xs[m++] = v; /* cu:1105-1114 */
        }
        if (nn == 0) {
            value[g] = qnan;
            meta[g] = KC_META_PACK(0, 0, 0, present, 0);
            continue;
        }
        if (nn == 1) { /* cu:1085-1086: the original object, untouched */
            value[g] = vals[g * n + first_nn];
            meta[g] = KC_META_PACK(first_nn, 1, 1, present, KC_FLAG_HAS_VALUE | KC_FLAG_SINGLE);
            continue;
        }
        if (m == 0) { /* cu:1115-1116 */
            value[g] = qnan;
            meta[g] = KC_META_PACK(0, 0, nn, present, KC_FLAG_NO_FINITE);
            continue;
        }
        for (int i = 1; i < m; i++) { /* xs.sort() */
            double t = xs[i];
            int j = i - 1;
            while (j >= 0 && xs[j] > t) {
                xs[j + 1] = xs[j];
                j--;
            }
            xs[j + 1] = t;
        }
        int start[MAXN], size[MAXN], n_cl = 0; /* _cluster_1d, cu:1127-1144 */
        start[0] = 0;
        for (int i = 0; i + 1 < m; i++)
            if (!is_close(xs[i], xs[i + 1], rel_eps, abs_eps)) {
                size[n_cl] = i + 1 - start[n_cl];
                n_cl++;
                start[n_cl] = i + 1;
            }
        size[n_cl] = m - start[n_cl];
        n_cl++;
        int top = 0, n_top = 0, first_top = 0;
        for (int k = 0; k < n_cl; k++)
            if (size[k] > top) top = size[k];
        for (int k = n_cl - 1; k >= 0; k--)
            if (size[k] == top) {
                n_top++;
                first_top = k;
            }
        if (n_top == 1) { /* cu:1171-1187 */
            value[g] = ko_np_mean(xs + start[first_top], top);
            meta[g] = KC_META_PACK(0, top, nn, present, KC_FLAG_HAS_VALUE);
            continue;
        }
        /* tie between equally large clusters, cu:1189-1219 */
        double center[MAXN], spread[MAXN];
        for (int k = 0; k < n_cl; k++) {
            center[k] = ko_np_median_sorted(xs + start[k], size[k]);
            spread[k] = size[k] > 1 ? ko_np_std(xs + start[k], size[k]) : 0.0;
        }
        int best = -1, best_support = 0;
        for (int k = 0; k < n_cl; k++) {
            if (size[k] != top) continue;
            int support = top;
            for (int o = 0; o < n_cl; o++) {
                if (o == k || size[o] >= top) continue;
                if (is_close(center[k], center[o], rel_eps, abs_eps) ||
                    is_close(fabs(center[k]), fabs(center[o]), rel_eps, abs_eps) ||
                    is_close_pow10(center[k], center[o], rel_eps, abs_eps))
                    support += size[o];
            }
            /* sort key (-support, spread, -|center|), stable: strict improvement only */
            int better = best < 0 || support > best_support ||
                         (support == best_support &&
                          (spread[k] < spread[best] ||
                           (spread[k] == spread[best] && fabs(center[k]) > fabs(center[best]))));
            if (better) {
                best = k;
                best_support = support;
            }
        }
        value[g] = ko_np_mean(xs + start[best], size[best]);
        meta[g] = KC_META_PACK(0, best_support, nn, present, KC_FLAG_HAS_VALUE | KC_FLAG_TIE);
    }
}

/* ------------------------------------------------------------------ K3 oracle: logprob sums (self-defined spec) */

/* fp32, fixed order: lane l of 32 adds elements l, l+32, ... left to right from +0.0f; then the
 * partials are combined by a xor butterfly with strides 16, 8, 4, 2, 1 (lane i += lane i^s). */
void ko_logprob_sum_f32(const float *lp, const int64_t *offsets, int64_t n_seq, float *out) {
    for (int64_t s = 0; s < n_seq; s++) {
        volatile float lane[32];
        for (int l = 0; l < 32; l++) {
            float acc = 0.0f;
            for (int64_t i = offsets[s] + l; i < offsets[s + 1]; i += 32) acc += lp[i];
            lane[l] = acc;
        }
        for (int stride = 16; stride >= 1; stride >>= 1) {
            float nxt[32];
            for (int l = 0; l < 32; l++) nxt[l] = lane[l] + lane[l ^ stride];
            for (int l = 0; l < 32; l++) lane[l] = nxt[l];
        }
        out[s] = lane[0];
    }
}

/* ------------------------------------------------------------------ K3b oracle: likelihood-weighted vote (self-defined) */

/* exp(x), x <= 0, from single fp32 operations in a fixed order (mirrors kc::kexp; compiled with -ffp-contract=off). */
float ko_exp_f32(float x) {
    if (x < -87.0f) x = -87.0f;
    volatile float t = x * 1.44269504f;
    volatile float th = t + 0.5f;
    float k = floorf(th);
    volatile float f = t - k;
    volatile float p = 0.00133336f;
    p = p * f; p = p + 0.00961813f;
    p = p * f; p = p + 0.05550411f;
    p = p * f; p = p + 0.24022651f;
    p = p * f; p = p + 0.69314718f;
    p = p * f; p = p + 1.0f;
    union { int32_t i; float f; } sc;
    sc.i = ((int32_t)k + 127) << 23;
    volatile float r = p * sc.f;
    return r;
}

void ko_weighted_vote_i32(const int32_t *codes, const float *seq_lp, int64_t n_records, int32_t n_fields, int32_t n,
                          const int32_t *none_code, int32_t *win_code, uint32_t *meta, float *weight) {
    for (int64_t r = 0; r < n_records; r++) {
        float smax = -3.0e38f, w[MAXN];
        for (int c = 0; c < n; c++)
            if (seq_lp[r * n + c] > smax) smax = seq_lp[r * n + c];
        for (int c = 0; c < n; c++) {
            volatile float d = seq_lp[r * n + c] - smax;
            w[c] = ko_exp_f32(d);
        }
        for (int32_t f = 0; f < n_fields; f++) {
            int64_t g = r * n_fields + f;
            int32_t nc = none_code ? none_code[f] : -1;
            int32_t cls_code[MAXN];
            int cls_count[MAXN], cls_first[MAXN], n_cls = 0, present = 0, voters = 0;
            float cls_w[MAXN];
            volatile float total = 0.0f;
            for (int c = 0; c < n; c++) {
                int32_t v = codes[g * n + c];
                if (v < KC_CODE_NONE) continue;
                present++;
                if (v == KC_CODE_NONE) {
                    if (nc < 0) continue;
                    v = nc;
                }
                voters++;
                total = total + w[c];
                int k = 0;
                while (k < n_cls && cls_code[k] != v) k++;
                if (k == n_cls) {
                    cls_code[k] = v;
                    cls_count[k] = 0;
                    cls_first[k] = c;
                    cls_w[k] = 0.0f;
                    n_cls++;
                }
                cls_count[k]++;
                volatile float acc = cls_w[k] + w[c];
                cls_w[k] = acc;
            }
            if (voters == 0) {
                win_code[g] = KC_CODE_NONE;
                meta[g] = KC_META_PACK(0, 0, 0, present, 0);
                weight[g] = 0.0f;
                continue;
            }
            int best = 0, ties = 0;
            for (int k = 1; k < n_cls; k++)
                if (cls_w[k] > cls_w[best]) best = k;
            for (int k = 0; k < n_cls; k++)
                if (k != best && cls_w[k] == cls_w[best]) ties = 1;
            win_code[g] = cls_code[best];
            meta[g] = KC_META_PACK(cls_first[best], cls_count[best], voters, present,
                                   KC_FLAG_HAS_VALUE | (ties ? KC_FLAG_TIE : 0));
            volatile float q = cls_w[best] / total;
            weight[g] = q;
        }
    }
}

/* ------------------------------------------------------------------ K4 oracle: similarity medoid of string groups */

static int edit_distance(const uint8_t *a, int al, const uint8_t *b, int bl) {
    int prev[2048], cur[2048];
    if (al > 2047 || bl > 2047) return -1;
    for (int j = 0; j <= bl; j++) prev[j] = j;
    for (int i = 1; i <= al; i++) {
        cur[0] = i;
        for (int j = 1; j <= bl; j++) {
            int v = prev[j] + 1;
            if (cur[j - 1] + 1 < v) v = cur[j - 1] + 1;
            if (prev[j - 1] + (a[i - 1] != b[j - 1]) < v) v = prev[j - 1] + (a[i - 1] != b[j - 1]);
            cur[j] = v;
        }
        memcpy(prev, cur, sizeof(int) * (size_t)(bl + 1));
    }
    return prev[bl];
}

/* cu:1221-1237 with levenshtein_similarity (cu:745-761) on already normalised strings: sim matrix, np.nanmean per row
 * (a copy with NaN -> 0 summed in numpy's pairwise order, divided by the count of non-NaN), first argmax. */
void ko_medoid_str(const uint8_t *chars, const int32_t *str_off, const int32_t *grp_off, int64_t n_groups, int32_t *best_idx,
                   double *best_avg) {
    for (int64_t g = 0; g < n_groups; g++) {
        int s0 = grp_off[g], k = grp_off[g + 1] - s0;
        double best = -1.0;
        int bi = 0;
        for (int i = 0; i < k; i++) {
            double row[MAXN];
            int li = str_off[s0 + i + 1] - str_off[s0 + i];
            for (int j = 0; j < k; j++) {
                if (j == i) {
                    row[j] = 0.0;
                    continue;
                }
                int lj = str_off[s0 + j + 1] - str_off[s0 + j];
                int mx = li > lj ? li : lj;
                if (mx == 0) {
                    row[j] = 1.0;
                    continue;
                }
                int d = edit_distance(chars + str_off[s0 + i], li, chars + str_off[s0 + j], lj);
                double sim = 1.0 - ((double)d / (double)mx);
                row[j] = sim > 1e-8 ? sim : 1e-8;
            }
            double avg = ko_np_sum(row, k) / (double)(k - 1);
            if (avg > best) {
                best = avg;
                bi = i;
            }
        }
        best_idx[g] = bi;
        best_avg[g] = best;
    }
}
