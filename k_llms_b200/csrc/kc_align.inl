// kc_align.inl — H2, included by kc_json.cpp inside its anonymous namespace (it uses the scanner, the token helpers and the
// json.dumps emitters defined there).

// =====================================================================================================================
// H2 — the alignment pre-pass of the client path, natively: recursive_list_alignments
// (consensus_utils.py:458-613) with lists_alignment (:383-430: dynamic threshold :185-252, reference list :255-333,
// assignment :336-380, pruning :109-149) and majority ordering (majority_sorting.py:8-112).  Works on a value TREE parsed
// from the candidate texts; key_mappings (a by-product the consolidation never reads) are not produced.
// =====================================================================================================================

enum AType : uint8_t { A_NONE = 0, A_BOOL, A_INT, A_FLOAT, A_STR, A_LIST, A_DICT };

struct AVal {
    AType t = A_NONE;
    bool b = false;
    double num = 0.0;                              // A_INT (as float(v)), A_FLOAT
    std::string s;                                 // A_STR: the value; A_INT: its decimal text
    std::vector<int32_t> items;                    // A_LIST: children
    std::vector<std::pair<std::string, int32_t>> kv;  // A_DICT: insertion order (a repeated key keeps its first position)
};

struct ATree {
    std::vector<AVal> v;
    int32_t add(AType t) {
        v.emplace_back();
        v.back().t = t;
        return (int32_t)v.size() - 1;
    }
    int32_t none() { return add(A_NONE); }
};

// json.loads of one value into the tree; false: not valid JSON / non-ASCII / nesting too deep
bool aparse(Scanner &sc, ATree &tr, int32_t &out, int depth) {
    if (depth > 200) return false;  // the same limit as Scanner::skip_nested: both parsers accept the same texts
    sc.ws();
    if (sc.p >= sc.end) return false;
    const char c = *sc.p;
    if (c == '{') {
        ++sc.p;
        out = tr.add(A_DICT);
        tr.v[(size_t)out].kv.reserve(8);  // one allocation for the usual object instead of four doublings
        sc.ws();
        if (sc.p < sc.end && *sc.p == '}') {
            ++sc.p;
            return true;
        }
        for (;;) {
            sc.ws();
            if (sc.p >= sc.end || *sc.p != '"') return false;
            const char *kp;
            uint32_t kl;
            bool kesc;
            if (!sc.string(kp, kl, kesc)) return false;
            std::string key;
            if (kesc) unescape(kp, kl, key);
            else key.assign(kp, kl);
            sc.ws();
            if (sc.p >= sc.end || *sc.p != ':') return false;
            ++sc.p;
            int32_t child;
            if (!aparse(sc, tr, child, depth + 1)) return false;
            bool found = false;
            for (auto &e : tr.v[(size_t)out].kv)
                if (e.first == key) {
                    e.second = child;
                    found = true;
                    break;
                }
            if (!found) tr.v[(size_t)out].kv.emplace_back(std::move(key), child);
            sc.ws();
            if (sc.p < sc.end && *sc.p == ',') {
                ++sc.p;
                continue;
            }
            if (sc.p < sc.end && *sc.p == '}') {
                ++sc.p;
                return true;
            }
            return false;
        }
    }
    if (c == '[') {
        ++sc.p;
        out = tr.add(A_LIST);
        tr.v[(size_t)out].items.reserve(8);
        sc.ws();
        if (sc.p < sc.end && *sc.p == ']') {
            ++sc.p;
            return true;
        }
        for (;;) {
            int32_t child;
            if (!aparse(sc, tr, child, depth + 1)) return false;
            tr.v[(size_t)out].items.push_back(child);
            sc.ws();
            if (sc.p < sc.end && *sc.p == ',') {
                ++sc.p;
                continue;
            }
            if (sc.p < sc.end && *sc.p == ']') {
                ++sc.p;
                return true;
            }
            return false;
        }
    }
    Tok t;
    if (!sc.value(t)) return false;
    switch (t.type) {
        case T_NULL: out = tr.add(A_NONE); break;
        case T_TRUE: out = tr.add(A_BOOL); tr.v[(size_t)out].b = true; break;
        case T_FALSE: out = tr.add(A_BOOL); break;
        case T_INT: out = tr.add(A_INT); tr.v[(size_t)out].num = t.num; int_text(t, tr.v[(size_t)out].s); break;
        case T_FLOAT: out = tr.add(A_FLOAT); tr.v[(size_t)out].num = t.num; break;
        case T_STR: out = tr.add(A_STR); tok_string(t, tr.v[(size_t)out].s); break;
        default: return false;
    }
    return true;
}

// json.dumps(value) with Python's default separators
void adump(const ATree &tr, int32_t id, std::string &out) {
    const AVal &v = tr.v[(size_t)id];
    switch (v.t) {
        case A_NONE: out += "null"; break;
        case A_BOOL: out += v.b ? "true" : "false"; break;
        case A_INT: out += v.s; break;
        case A_FLOAT: json_float(v.num, out); break;
        case A_STR: json_string(v.s, out); break;
        case A_LIST: {
            out += "[";
            for (size_t i = 0; i < v.items.size(); ++i) {
                if (i) out += ", ";
                adump(tr, v.items[i], out);
            }
            out += "]";
            break;
        }
        case A_DICT: {
            out += "{";
            for (size_t i = 0; i < v.kv.size(); ++i) {
                if (i) out += ", ";
                json_string(v.kv[i].first, out);
                out += ": ";
                adump(tr, v.kv[i].second, out);
            }
            out += "}";
            break;
        }
    }
}


// ---------------------------------------------------------------- similarities (consensus_utils.py:626-917)

constexpr double kSimFloor = 1e-8;  // SIMILARITY_SCORE_LOWER_BOUND, cu:78

struct AlignCtx {
    ATree tr;
    bool decline = false;  // not decidable here: a pair of strings both longer than 50 characters (cu:813 asks the embeddings
                           // service), or nesting beyond the recursion budget -> the record takes the Python path
    std::vector<int32_t> dp;        // edit-distance row
};

int32_t edit_distance(AlignCtx &cx, const std::string &a0, const std::string &b0) {
    const std::string *a = &a0, *b = &b0;
    if (a->size() < b->size()) std::swap(a, b);
    const int32_t al = (int32_t)a->size(), bl = (int32_t)b->size();
    if (bl == 0) return al;
    cx.dp.resize((size_t)bl + 1);
    for (int32_t j = 0; j <= bl; ++j) cx.dp[(size_t)j] = j;
    for (int32_t i = 1; i <= al; ++i) {
        int32_t diag = cx.dp[0];
        cx.dp[0] = i;
        const char ca = (*a)[(size_t)i - 1];
        for (int32_t j = 1; j <= bl; ++j) {
            const int32_t up = cx.dp[(size_t)j];
            const int32_t v = std::min(std::min(up + 1, cx.dp[(size_t)j - 1] + 1), diag + (ca != (*b)[(size_t)j - 1] ? 1 : 0));
            diag = up;
            cx.dp[(size_t)j] = v;
        }
    }
    return cx.dp[(size_t)bl];
}

bool a_falsy(const AVal &v) {  // `not bool(v)`
    switch (v.t) {
        case A_NONE: return true;
        case A_BOOL: return !v.b;
        case A_INT: return v.s == "0";
        case A_FLOAT: return v.num == 0.0;
        case A_STR: return v.s.empty();
        case A_LIST: return v.items.empty();
        case A_DICT: return v.kv.empty();
    }
    return false;
}

inline bool a_numeric(const AVal &v) { return v.t == A_BOOL || v.t == A_INT || v.t == A_FLOAT; }  // isinstance(v, (int, float))
inline double a_number(const AVal &v) { return v.t == A_BOOL ? (v.b ? 1.0 : 0.0) : v.num; }

bool py_isclose(double a, double b, double rel_tol) {  // math.isclose(a, b, rel_tol=rel_tol)
    if (a == b) return true;
    if (std::isinf(a) || std::isinf(b)) return false;
    const double diff = std::fabs(b - a);
    return diff <= std::fabs(rel_tol * b) || diff <= std::fabs(rel_tol * a);
}

double string_similarity(AlignCtx &cx, const std::string &s1, const std::string &s2) {  // cu:797-824, method "embeddings"
    if (s1.size() > 50 && s2.size() > 50) cx.decline = true;  // the caller gives the record to the Python path
    if (s1 == s2) return 1.0;  // equal texts normalise alike: distance 0 (or two empty forms) — the common pair among candidates
    thread_local std::string a, b;
    sanitize(s1, a);  // == normalize_string (cu:660-673) on ASCII
    sanitize(s2, b);
    const size_t longest = std::max(a.size(), b.size());
    if (longest == 0) return 1.0;
    const double sim = 1.0 - ((double)edit_distance(cx, a, b) / (double)longest);
    return sim > kSimFloor ? sim : kSimFloor;
}

double numerical_similarity(const AVal &v1, const AVal &v2) {  // cu:827-841
    if (v1.t == A_BOOL && v2.t == A_BOOL) return v1.b == v2.b ? 1.0 : kSimFloor;
    const double a = a_number(v1), b = a_number(v2);
    if (py_isclose(a, b, 0.01)) return 1.0;
    const bool eq = (v1.t == A_INT && v2.t == A_INT) ? v1.s == v2.s : a == b;
    return eq ? 1.0 : kSimFloor;
}

const AVal *a_get(const ATree &tr, const AVal &d, const std::string &key) {
    for (auto &e : d.kv)
        if (e.first == key) return &tr.v[(size_t)e.second];
    return nullptr;
}

double generic_similarity(AlignCtx &cx, const AVal *p1, const AVal *p2) {  // cu:892-917; nullptr == None
    static const AVal kNone;
    const AVal &v1 = p1 ? *p1 : kNone, &v2 = p2 ? *p2 : kNone;
    if (a_falsy(v1) && a_falsy(v2)) return 1.0;
    if (v1.t == A_NONE || v2.t == A_NONE) return kSimFloor;
    if (v1.t == A_STR && v2.t == A_STR) return string_similarity(cx, v1.s, v2.s);
    if (a_numeric(v1) && a_numeric(v2)) return numerical_similarity(v1, v2);
    if (v1.t == A_DICT && v2.t == A_DICT) {  // cu:844-869.  The reference adds the fields in set order (hash-seed dependent);
        std::vector<const std::string *> keys;  // here: sorted keys — equal up to the rounding of the sum
        auto want = [](const std::string &k) { return k.rfind("reasoning___", 0) != 0 && k.rfind("source___", 0) != 0; };
        for (auto &e : v1.kv)
            if (want(e.first)) keys.push_back(&e.first);
        for (auto &e : v2.kv)
            if (want(e.first) && !a_get(cx.tr, v1, e.first)) keys.push_back(&e.first);
        if (keys.empty()) return 1.0;
        std::sort(keys.begin(), keys.end(), [](const std::string *x, const std::string *y) { return *x < *y; });
        double total = 0.0;
        for (const std::string *k : keys) total += generic_similarity(cx, a_get(cx.tr, v1, *k), a_get(cx.tr, v2, *k));
        return total / (double)keys.size();
    }
    if (v1.t == A_LIST && v2.t == A_LIST) {  // cu:872-889
        const size_t longest = std::max(v1.items.size(), v2.items.size());
        if (longest == 0) return 1.0;
        double total = 0.0;
        for (size_t i = 0; i < longest; ++i)
            total += generic_similarity(cx, i < v1.items.size() ? &cx.tr.v[(size_t)v1.items[i]] : nullptr,
                                        i < v2.items.size() ? &cx.tr.v[(size_t)v2.items[i]] : nullptr);
        return total / (double)longest;
    }
    return kSimFloor;
}


// ---------------------------------------------------------------- min-cost assignment
//
// scipy.optimize.linear_sum_assignment (what cu:362 calls) is Crouse's shortest-augmenting-path algorithm
// ("On implementing 2D rectangular assignment algorithms", 2016).  Ties are resolved by its exact scan order (columns
// visited from the last to the first, a free column preferred among equals), so the algorithm is restated here step for
// step and checked against scipy on tie-heavy matrices (tests/test_align_native.py).  Returns pairs sorted by row.
bool lsap(int nr, int nc, const double *cost_in, std::vector<int> &rows, std::vector<int> &cols) {
    rows.clear();
    cols.clear();
    if (nr == 0 || nc == 0) return true;
    const bool transpose = nc < nr;
    std::vector<double> temp;
    const double *cost = cost_in;
    if (transpose) {
        temp.resize((size_t)nr * nc);
        for (int i = 0; i < nr; ++i)
            for (int j = 0; j < nc; ++j) temp[(size_t)j * nr + i] = cost_in[(size_t)i * nc + j];
        std::swap(nr, nc);
        cost = temp.data();
    }
    for (size_t k = 0; k < (size_t)nr * nc; ++k)
        if (cost[k] != cost[k] || cost[k] == -INFINITY) return false;
    std::vector<double> u((size_t)nr, 0.0), v((size_t)nc, 0.0), shortest((size_t)nc);
    std::vector<int> path((size_t)nc, -1), col4row((size_t)nr, -1), row4col((size_t)nc, -1), remaining((size_t)nc);
    std::vector<char> SR((size_t)nr), SC((size_t)nc);
    for (int cur = 0; cur < nr; ++cur) {
        double min_val = 0.0;
        int num_remaining = nc, i = cur, sink = -1;
        for (int it = 0; it < nc; ++it) remaining[(size_t)it] = nc - it - 1;
        std::fill(SR.begin(), SR.end(), 0);
        std::fill(SC.begin(), SC.end(), 0);
        std::fill(shortest.begin(), shortest.end(), INFINITY);
        while (sink == -1) {
            int index = -1;
            double lowest = INFINITY;
            SR[(size_t)i] = 1;
            for (int it = 0; it < num_remaining; ++it) {
                const int j = remaining[(size_t)it];
                const double r = min_val + cost[(size_t)i * nc + j] - u[(size_t)i] - v[(size_t)j];
                if (r < shortest[(size_t)j]) {
                    path[(size_t)j] = i;
                    shortest[(size_t)j] = r;
                }
                if (shortest[(size_t)j] < lowest || (shortest[(size_t)j] == lowest && row4col[(size_t)j] == -1)) {
                    lowest = shortest[(size_t)j];
                    index = it;
                }
            }
            min_val = lowest;
            if (min_val == INFINITY) return false;
            const int j = remaining[(size_t)index];
            if (row4col[(size_t)j] == -1) sink = j;
            else i = row4col[(size_t)j];
            SC[(size_t)j] = 1;
            remaining[(size_t)index] = remaining[(size_t)--num_remaining];
        }
        u[(size_t)cur] += min_val;
        for (int r = 0; r < nr; ++r)
            if (SR[(size_t)r] && r != cur) u[(size_t)r] += min_val - shortest[(size_t)col4row[(size_t)r]];
        for (int j = 0; j < nc; ++j)
            if (SC[(size_t)j]) v[(size_t)j] -= min_val - shortest[(size_t)j];
        int j = sink;
        for (;;) {
            const int r = path[(size_t)j];
            row4col[(size_t)j] = r;
            std::swap(col4row[(size_t)r], j);
            if (r == cur) break;
        }
    }
    if (transpose) {  // rows of the caller are the columns here: report them in ascending order
        std::vector<int> order((size_t)nr);
        for (int r = 0; r < nr; ++r) order[(size_t)r] = r;
        std::sort(order.begin(), order.end(), [&](int x, int y) { return col4row[(size_t)x] < col4row[(size_t)y]; });
        for (int r : order) {
            rows.push_back(col4row[(size_t)r]);
            cols.push_back(r);
        }
    } else {
        for (int r = 0; r < nr; ++r) {
            rows.push_back(r);
            cols.push_back(col4row[(size_t)r]);
        }
    }
    return true;
}


// ---------------------------------------------------------------- lists_alignment (consensus_utils.py:81-430, majority_sorting.py)

// numpy's float64 add.reduce over a contiguous row: pairwise with eight accumulators (the whole row is one block below
// 128 elements), exactly as the K2 / K4 kernels restate it
template <typename F>
double host_np_sum(F f, int n) {
    double res;
    if (n < 8) {
        res = -0.0;
        for (int i = 0; i < n; ++i) res += f(i);
    } else {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = f(j);
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += f(i + j);
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += f(i);
    }
    return 0.0 + res;
}

struct Idx {
    int li, pos;
    bool operator==(const Idx &o) const { return li == o.li && pos == o.pos; }
    bool operator<(const Idx &o) const { return li != o.li ? li < o.li : pos < o.pos; }
};

struct ListAligner {
    AlignCtx &cx;
    const std::vector<std::vector<int32_t>> &lists;  // per candidate: node ids of its elements
    std::vector<double> dense;                       // [flat a * total + flat b], NaN = not computed (small inputs)
    std::unordered_map<uint64_t, double> sparse;     // the same for big inputs (only the pairs that are asked for)
    std::vector<int> base;                           // flat index of (li, 0)
    size_t total = 0;

    ListAligner(AlignCtx &c, const std::vector<std::vector<int32_t>> &l) : cx(c), lists(l) {
        for (auto &x : lists) {
            base.push_back((int)total);
            total += x.size();
        }
        if (total <= 512) dense.assign(total * total, NAN);
    }
    double sim(Idx a, Idx b) {  // _PairSims.get (cu:81-106): a symmetric memo
        size_t fa = (size_t)(base[(size_t)a.li] + a.pos), fb = (size_t)(base[(size_t)b.li] + b.pos);
        auto compute = [&] {
            return generic_similarity(cx, &cx.tr.v[(size_t)lists[(size_t)a.li][(size_t)a.pos]], &cx.tr.v[(size_t)lists[(size_t)b.li][(size_t)b.pos]]);
        };
        if (!dense.empty()) {
            double &m = dense[fa * total + fb];
            if (m != m) {
                m = compute();
                dense[fb * total + fa] = m;
            }
            return m;
        }
        if (fa > fb) std::swap(fa, fb);
        const uint64_t key = ((uint64_t)fa << 32) | (uint64_t)fb;
        auto it = sparse.find(key);
        if (it != sparse.end()) return it->second;
        const double m = compute();
        sparse.emplace(key, m);
        return m;
    }

    static double low_cutoff(std::vector<double> s) {  // cu:152-174
        if (s.empty()) return 0.0;
        std::sort(s.begin(), s.end());
        double cut = s[0];
        const int head = (int)(0.2 * (double)s.size());
        std::vector<double> gaps;
        for (int i = 1; i < head; ++i) gaps.push_back(s[(size_t)i] - s[(size_t)i - 1]);
        if (!gaps.empty()) {
            std::vector<double> g = gaps;
            std::sort(g.begin(), g.end());
            const size_t n = g.size();
            const double median = (n & 1) ? g[n / 2] : (g[n / 2 - 1] + g[n / 2]) / 2.0;
            const double jump = median * 3;
            size_t at = 0;
            for (size_t i = 0; i < gaps.size(); ++i)
                if (gaps[i] > jump) {
                    at = i;
                    break;
                }
            if (gaps[at] > jump) cut = s[at + 1] + 0.0001;
        }
        return cut;
    }

    double dynamic_threshold() {  // cu:185-252
        const int L = (int)lists.size();
        if (L < 2) return 0.5;
        std::vector<double> best_scores;
        for (int i = 0; i < L; ++i) {
            if (lists[(size_t)i].empty()) continue;
            std::vector<std::vector<char>> taken((size_t)L);
            for (int j = 0; j < L; ++j) taken[(size_t)j].assign(lists[(size_t)j].size(), 0);
            for (int ki = 0; ki < (int)lists[(size_t)i].size(); ++ki) {
                double top = 0.5;
                Idx partner{-1, -1};
                for (int j = i + 1; j < L; ++j)
                    for (int kj = 0; kj < (int)lists[(size_t)j].size(); ++kj) {
                        if (taken[(size_t)j][(size_t)kj]) continue;
                        const double s = sim({i, ki}, {j, kj});
                        if (s > top) {
                            top = s;
                            partner = {j, kj};
                        }
                    }
                if (partner.li >= 0 && top > 0) {
                    best_scores.push_back(top);
                    taken[(size_t)partner.li][(size_t)partner.pos] = 1;
                }
            }
        }
        std::sort(best_scores.begin(), best_scores.end());
        const double floor_ = low_cutoff(best_scores);
        for (double x : best_scores)
            if (x >= floor_) return std::max(0.5, 0.95 * x);
        return 0.5;
    }

    // cu:306-311: consensus_as_primitive over the (list, position) TUPLES, i.e. their similarity medoid — positions and list
    // numbers compared as numbers: equal, both zero, or within 1 % count as 1.0, anything else as 1e-8
    static double index_part(int a, int b) {
        if (a == 0 && b == 0) return 1.0;
        if (py_isclose((double)a, (double)b, 0.01)) return 1.0;
        return a == b ? 1.0 : kSimFloor;
    }
    static Idx elect(const std::vector<Idx> &members) {
        const int n = (int)members.size();
        if (n == 1) return members[0];
        std::vector<double> sims((size_t)n * n, 0.0);
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j) {
                double total = 0.0;
                total += index_part(members[(size_t)i].li, members[(size_t)j].li);
                total += index_part(members[(size_t)i].pos, members[(size_t)j].pos);
                sims[(size_t)i * n + j] = sims[(size_t)j * n + i] = total / 2.0;
            }
        int best = 0;
        double best_avg = -INFINITY;
        for (int i = 0; i < n; ++i) {
            const double avg = host_np_sum([&](int j) { return j == i ? 0.0 : sims[(size_t)i * n + j]; }, n) / (double)(n - 1);
            if (avg > best_avg) {
                best_avg = avg;
                best = i;
            }
        }
        return members[(size_t)best];
    }

    std::vector<Idx> build_reference(double min_support_ratio, double threshold) {  // cu:255-333
        struct Group {
            Idx rep;
            std::vector<Idx> members;
            std::vector<char> used;  // lists that already have a member
        };
        std::vector<Group> groups;  // insertion order; a re-keyed group moves to the end (dict pop + insert)
        const int L = (int)lists.size();
        for (int li = 0; li < L; ++li)
            for (int pos = 0; pos < (int)lists[(size_t)li].size(); ++pos) {
                const Idx cand{li, pos};
                double best = -1;
                int home = -1;
                for (int gi = 0; gi < (int)groups.size(); ++gi) {
                    if (groups[(size_t)gi].used[(size_t)li]) continue;
                    const double s = sim(cand, groups[(size_t)gi].rep);
                    if (s >= threshold && s > best) {
                        best = s;
                        home = gi;
                    }
                }
                if (home < 0) {
                    Group g;
                    g.rep = cand;
                    g.members.push_back(cand);
                    g.used.assign((size_t)L, 0);
                    g.used[(size_t)li] = 1;
                    groups.push_back(std::move(g));
                    continue;
                }
                Group &g = groups[(size_t)home];
                g.members.push_back(cand);
                g.used[(size_t)li] = 1;
                const Idx new_rep = elect(g.members);
                if (!(new_rep == g.rep)) {
                    Group moved = std::move(g);
                    moved.rep = new_rep;
                    groups.erase(groups.begin() + home);
                    groups.push_back(std::move(moved));
                }
            }
        struct Kept {
            double ratio;
            Idx rep;
        };
        std::vector<Kept> kept;
        for (auto &g : groups) {
            const double ratio = (double)g.members.size() / (double)L;
            if (ratio >= min_support_ratio) kept.push_back({ratio, g.rep});
        }
        std::stable_sort(kept.begin(), kept.end(), [](const Kept &a, const Kept &b) {
            if (-a.ratio != -b.ratio) return -a.ratio < -b.ratio;
            return a.rep < b.rep;
        });
        std::vector<Idx> reference;
        for (auto &k : kept) reference.push_back(k.rep);
        return reference;
    }

    // aligned[li][slot] = position in lists[li] or -1 (cu:336-380)
    std::vector<std::vector<int>> assign(const std::vector<Idx> &reference, double threshold) {
        const int L = (int)lists.size(), R = (int)reference.size();
        std::vector<std::vector<int>> aligned((size_t)L, std::vector<int>((size_t)R, -1));
        if (R == 0) return aligned;
        std::vector<double> simm, cost;
        std::vector<int> rows, cols;
        for (int li = 0; li < L; ++li) {
            const int n = (int)lists[(size_t)li].size();
            if (n == 0) continue;
            simm.assign((size_t)R * n, 0.0);
            cost.assign((size_t)R * n, 0.0);
            for (int r = 0; r < R; ++r)
                for (int pos = 0; pos < n; ++pos) {
                    const Idx me{li, pos};
                    const double s = (me == reference[(size_t)r]) ? 1.0 : sim(me, reference[(size_t)r]);
                    simm[(size_t)r * n + pos] = s;
                    cost[(size_t)r * n + pos] = 1.0 - s;
                }
            lsap(R, n, cost.data(), rows, cols);
            for (size_t k = 0; k < rows.size(); ++k) {
                const int r = rows[k], pos = cols[k];
                // a None ELEMENT assigned to a slot leaves the slot None (cu:377-378 stores lst[pos] itself)
                if (simm[(size_t)r * n + pos] >= threshold && aligned[(size_t)li][(size_t)r] < 0 &&
                    cx.tr.v[(size_t)lists[(size_t)li][(size_t)pos]].t != A_NONE)
                    aligned[(size_t)li][(size_t)r] = pos;
            }
        }
        return aligned;
    }

    static void prune(std::vector<std::vector<int>> &aligned, double min_support_ratio) {  // cu:109-149
        if (aligned.empty()) return;
        const int width = (int)aligned[0].size();
        if (width == 0) return;
        std::vector<double> support((size_t)width);
        double top = -INFINITY;
        for (int c = 0; c < width; ++c) {
            int filled = 0;
            for (auto &row : aligned) filled += row[(size_t)c] >= 0;
            support[(size_t)c] = (double)filled / (double)aligned.size();
            top = std::max(top, support[(size_t)c]);
        }
        const double bar = top < min_support_ratio ? std::min(min_support_ratio, top) : min_support_ratio;
        for (auto &row : aligned) {
            std::vector<int> keep;
            for (int c = 0; c < width; ++c)
                if (support[(size_t)c] >= bar) keep.push_back(row[(size_t)c]);
            row.swap(keep);
        }
    }

    // ms:8-23 matches aligned cells to their source list by OBJECT IDENTITY.  Elements parsed from JSON are distinct objects,
    // except the values CPython shares: True / False, the integers -5 .. 256, strings of at most one character.  For those
    // the identity dictionary keeps the LAST position holding an equal value.
    bool shared_object(const AVal &a, const AVal &b) const {
        if (a.t != b.t) return false;
        if (a.t == A_BOOL) return a.b == b.b;
        if (a.t == A_INT) return a.s == b.s && a.num >= -5 && a.num <= 256;
        if (a.t == A_STR) return a.s.size() <= 1 && a.s == b.s;
        return false;
    }
    int identity_position(int li, int pos) const {
        const auto &lst = lists[(size_t)li];
        const AVal &me = cx.tr.v[(size_t)lst[(size_t)pos]];
        int where = pos;
        for (int k = pos + 1; k < (int)lst.size(); ++k)
            if (shared_object(me, cx.tr.v[(size_t)lst[(size_t)k]])) where = k;
        return where;
    }

    void order_by_majority(std::vector<std::vector<int>> &aligned) {  // ms:78-112
        if (aligned.empty()) return;
        const int width = (int)aligned[0].size(), L = (int)aligned.size();
        std::vector<std::vector<int>> pos((size_t)L, std::vector<int>((size_t)width, -1));
        for (int r = 0; r < L; ++r)
            for (int c = 0; c < width; ++c)
                if (aligned[(size_t)r][(size_t)c] >= 0) pos[(size_t)r][(size_t)c] = identity_position(r, aligned[(size_t)r][(size_t)c]);
        std::vector<std::vector<int>> wins((size_t)width, std::vector<int>((size_t)width, 0));
        for (auto &row : pos)
            for (int a = 0; a < width; ++a)
                for (int b = 0; b < width; ++b)
                    if (row[(size_t)a] >= 0 && row[(size_t)b] >= 0 && row[(size_t)a] < row[(size_t)b]) ++wins[(size_t)a][(size_t)b];
        std::vector<std::vector<int>> after((size_t)width);
        std::vector<int> indeg((size_t)width, 0);
        for (int a = 0; a < width; ++a)
            for (int b = 0; b < width; ++b)
                if (a != b && wins[(size_t)a][(size_t)b] > wins[(size_t)b][(size_t)a]) {
                    after[(size_t)a].push_back(b);
                    ++indeg[(size_t)b];
                }
        std::vector<double> mean_pos((size_t)width);
        for (int c = 0; c < width; ++c) {
            double total = 0.0;
            int count = 0;
            for (auto &row : pos)
                if (row[(size_t)c] >= 0) {
                    total += row[(size_t)c];
                    ++count;
                }
            mean_pos[(size_t)c] = count ? total / count : INFINITY;
        }
        std::vector<char> done((size_t)width, 0), ready((size_t)width, 0);
        for (int c = 0; c < width; ++c) ready[(size_t)c] = indeg[(size_t)c] == 0;
        std::vector<int> order;
        for (;;) {  // heap of (mean position, column): the smallest pair pops first
            int u = -1;
            for (int c = 0; c < width; ++c)
                if (ready[(size_t)c] && !done[(size_t)c] &&
                    (u < 0 || mean_pos[(size_t)c] < mean_pos[(size_t)u] || (mean_pos[(size_t)c] == mean_pos[(size_t)u] && c < u)))
                    u = c;
            if (u < 0) break;
            done[(size_t)u] = 1;
            order.push_back(u);
            for (int v : after[(size_t)u])
                if (--indeg[(size_t)v] == 0) ready[(size_t)v] = 1;
        }
        if ((int)order.size() < width) {  // columns caught in a cycle: by mean position (stable)
            std::vector<int> rest;
            for (int c = 0; c < width; ++c)
                if (!done[(size_t)c]) rest.push_back(c);
            std::stable_sort(rest.begin(), rest.end(), [&](int a, int b) { return mean_pos[(size_t)a] < mean_pos[(size_t)b]; });
            order.insert(order.end(), rest.begin(), rest.end());
        }
        for (auto &row : aligned) {
            std::vector<int> re;
            for (int c : order) re.push_back(row[(size_t)c]);
            row.swap(re);
        }
    }

    // cu:383-430 without a reference list: rows of equal width, cells = positions in the candidate's own list or -1
    std::vector<std::vector<int>> run(double min_support_ratio) {
        const double thr = dynamic_threshold();
        const std::vector<Idx> reference = build_reference(min_support_ratio, thr);
        std::vector<std::vector<int>> aligned = assign(reference, 0.95 * thr);
        prune(aligned, min_support_ratio);
        order_by_majority(aligned);
        return aligned;
    }
};

// recursive_list_alignments (cu:458-613) on the tree: returns the aligned value of every candidate (node id, -1 == None)
void align_values(AlignCtx &cx, std::vector<int32_t> &values, double min_support_ratio, int depth) {
    if (values.empty()) return;
    if (depth > 100) {  // deeper than any real payload: leave it to the Python pre-pass
        cx.decline = true;
        return;
    }
    int first = -1;
    for (int32_t id : values)
        if (id >= 0 && cx.tr.v[(size_t)id].t != A_NONE) {
            first = id;
            break;
        }
    if (first < 0) return;  // all None
    const AType ft = cx.tr.v[(size_t)first].t;
    if (ft != A_DICT && ft != A_LIST) return;  // scalars are left alone (cu:507-512)
    for (int32_t id : values)  // mixed types are left alone as well
        if (id >= 0 && cx.tr.v[(size_t)id].t != A_NONE && cx.tr.v[(size_t)id].t != ft) return;
    const size_t n = values.size();
    if (ft == A_DICT) {  // cu:516-548: every candidate gets every key, keys sorted, missing -> None, recursively
        std::vector<std::string> keys;
        for (int32_t id : values)
            if (id >= 0 && cx.tr.v[(size_t)id].t == A_DICT)
                for (auto &e : cx.tr.v[(size_t)id].kv) keys.push_back(e.first);
        std::sort(keys.begin(), keys.end());
        keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
        std::vector<int32_t> out((size_t)n);
        for (size_t c = 0; c < n; ++c) out[c] = cx.tr.add(A_DICT);
        std::vector<int32_t> column((size_t)n);
        for (auto &key : keys) {
            for (size_t c = 0; c < n; ++c) {
                column[c] = -1;
                const int32_t id = values[c];
                if (id >= 0 && cx.tr.v[(size_t)id].t == A_DICT)
                    for (auto &e : cx.tr.v[(size_t)id].kv)
                        if (e.first == key) column[c] = e.second;
            }
            align_values(cx, column, min_support_ratio, depth + 1);
            for (size_t c = 0; c < n; ++c) {
                const int32_t child = column[c] >= 0 ? column[c] : cx.tr.none();
                cx.tr.v[(size_t)out[c]].kv.emplace_back(key, child);
            }
        }
        values = out;
        return;
    }
    // lists (cu:550-613)
    std::vector<std::vector<int32_t>> lists((size_t)n);
    bool any = false;
    for (size_t c = 0; c < n; ++c) {
        const int32_t id = values[c];
        if (id >= 0 && cx.tr.v[(size_t)id].t == A_LIST) lists[c] = cx.tr.v[(size_t)id].items;
        any |= !lists[c].empty();
    }
    std::vector<std::vector<int32_t>> rows((size_t)n);
    if (any) {
        ListAligner la(cx, lists);
        const std::vector<std::vector<int>> aligned = la.run(min_support_ratio);
        for (size_t c = 0; c < n; ++c)
            for (int pos : aligned[c]) rows[c].push_back(pos >= 0 ? lists[c][(size_t)pos] : -1);
    }
    const size_t width = rows.empty() ? 0 : rows[0].size();
    std::vector<int32_t> column((size_t)n);
    for (size_t col = 0; col < width; ++col) {
        for (size_t c = 0; c < n; ++c) column[c] = rows[c][col];
        align_values(cx, column, min_support_ratio, depth + 1);
        for (size_t c = 0; c < n; ++c) rows[c][col] = column[c];
    }
    for (size_t c = 0; c < n; ++c) {
        const int32_t out = cx.tr.add(A_LIST);
        for (int32_t cell : rows[c]) {
            const int32_t child = cell >= 0 ? cell : cx.tr.none();  // evaluated before the reference below is taken
            cx.tr.v[(size_t)out].items.push_back(child);
        }
        values[c] = out;
    }
}


