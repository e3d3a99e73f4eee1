// kc_json.cpp — H1: native host columnariser / decoder.
//
// kc_consolidate_json(): for every record, n candidate JSON texts in -> consensus JSON text + likelihoods JSON text out,
// with only the CUDA kernels in between.  It does, in C++ and multi-threaded, what the reference does per request in
// Python around the hot path:
//     _safe_parse_content            consolidation.py:25-38   (json.loads, or {"text": content})
//     recursive_list_alignments      consensus_utils.py:516-548 (dict part: every candidate gets every key, keys SORTED)
//     consensus_values dispatcher    consensus_utils.py:1376-1454 (scalar fields, nested objects, lists element-wise)
//     lists_alignment                consensus_utils.py:185-430 + majority_sorting.py (H2 below; records with list fields)
//     sanitize_value / `v or False`  consensus_utils.py:925-933, 956   -> local dictionary codes (int8 cells)
//     _format_consensus_content      consolidation.py:41-60   (json.dumps of the consensus; {"text": s} -> s)
//     similarity medoid              consensus_utils.py:1221-1237 for multi-word string fields (batched into one K4 launch)
// Records it cannot express (a key mixing objects with other types, string groups outside K4's contract, non-ASCII
// text, mixed-type bool groups) are NOT guessed at: they get status 1 and the Python path handles them.
//
// Text formats follow CPython exactly: float -> float.__repr__ (shortest round-trip digits, exponent form outside
// 1e-4 <= |x| < 1e16, always a fractional part), json.dumps separators ", " / ": " and ensure_ascii escaping.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/kllms_b200.h"
#include "kc_jsoncore.cuh"  // to_double: exact decimal -> float64 without strtod (host-callable)

namespace {

enum TokType : uint8_t { T_MISSING = 0, T_NULL, T_TRUE, T_FALSE, T_INT, T_FLOAT, T_STR, T_NESTED };

// A token is a VIEW into the candidate text (no copies): strings keep their raw inner span and an "has escapes" flag
// and are unescaped only where the value is needed; numbers keep their text span and the strtod value.
struct Tok {
    TokType type = T_MISSING;
    bool esc = false;
    const char *p = nullptr;
    uint32_t len = 0;
    double num = 0.0;  // T_INT / T_FLOAT (strtod: correctly rounded, like float(int) / float(text))
};

struct Item {
    std::string_view key;
    Tok tok;
};

inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
// str.split() / str.strip() whitespace within ASCII: \t \n \v \f \r, \x1c-\x1f and space
inline bool is_py_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f); }

// raw inner span of a JSON string -> value (ASCII only; \uXXXX above 0x7F never gets here)
void unescape(const char *p, uint32_t len, std::string &out) {
    out.clear();
    const char *e = p + len;
    while (p < e) {
        if (*p != '\\') {
            out.push_back(*p++);
            continue;
        }
        ++p;
        switch (*p++) {
            case 'b': out.push_back('\b'); break;
            case 'f': out.push_back('\f'); break;
            case 'n': out.push_back('\n'); break;
            case 'r': out.push_back('\r'); break;
            case 't': out.push_back('\t'); break;
            case 'u': {
                unsigned v = 0;
                for (int i = 0; i < 4; ++i) {
                    const char h = p[i];
                    v = (v << 4) | (unsigned)(h <= '9' ? h - '0' : (h | 0x20) - 'a' + 10);
                }
                p += 4;
                out.push_back((char)(v & 0x7F));
                break;
            }
            default: out.push_back(p[-1]); break;  // \" \\ \/
        }
    }
}

inline void tok_string(const Tok &t, std::string &out) {  // value of a T_STR token
    if (t.esc) unescape(t.p, t.len, out);
    else out.assign(t.p, t.len);
}

struct Scanner {
    const char *p, *end;
    bool non_ascii = false;
    void ws() {
        while (p < end && is_ws(*p)) ++p;
    }
    bool lit(const char *s, size_t n) {
        if ((size_t)(end - p) >= n && memcmp(p, s, n) == 0) {
            p += n;
            return true;
        }
        return false;
    }
    // at the opening quote: validates, reports the raw inner span
    bool string(const char *&sp, uint32_t &slen, bool &esc) {
        ++p;
        sp = p;
        esc = false;
        while (p < end) {
            const unsigned char c = (unsigned char)*p;
            if (c == '"') {
                slen = (uint32_t)(p - sp);
                ++p;
                return true;
            }
            if (c < 0x20) return false;  // json.loads(strict=True) rejects raw control characters
            if (c >= 0x80) non_ascii = true;
            if (c != '\\') {
                ++p;
                continue;
            }
            esc = true;
            if (++p >= end) return false;
            const char k = *p++;
            if (k == 'u') {
                if (end - p < 4) return false;
                unsigned v = 0;
                for (int i = 0; i < 4; ++i) {
                    const char h = p[i];
                    v <<= 4;
                    if (h >= '0' && h <= '9') v |= (unsigned)(h - '0');
                    else if (h >= 'a' && h <= 'f') v |= (unsigned)(h - 'a' + 10);
                    else if (h >= 'A' && h <= 'F') v |= (unsigned)(h - 'A' + 10);
                    else return false;
                }
                p += 4;
                if (v >= 0x80) non_ascii = true;  // parity for non-ASCII text is unpinned: hand the record to Python
            } else if (!(k == '"' || k == '\\' || k == '/' || k == 'b' || k == 'f' || k == 'n' || k == 'r' || k == 't')) {
                return false;
            }
        }
        return false;
    }
    // at '{' or '[': validate the nested value exactly as json.loads would (a malformed inner value makes the WHOLE candidate
    // text invalid JSON, which the caller then treats as free text, consolidation.py:25-38) and skip it
    bool skip_nested(int depth = 0) {
        if (depth > 200) return false;
        const char open = *p++;
        const char close = open == '{' ? '}' : ']';
        ws();
        if (p < end && *p == close) {
            ++p;
            return true;
        }
        for (;;) {
            ws();
            if (open == '{') {
                if (p >= end || *p != '"') return false;
                const char *sp;
                uint32_t sl;
                bool esc;
                if (!string(sp, sl, esc)) return false;
                ws();
                if (p >= end || *p != ':') return false;
                ++p;
            }
            Tok inner;
            if (!value(inner, depth + 1)) return false;
            ws();
            if (p < end && *p == ',') {
                ++p;
                continue;
            }
            if (p < end && *p == close) {
                ++p;
                return true;
            }
            return false;
        }
    }
    bool number(Tok &t) {
        const char *s = p;
        if (p < end && *p == '-') ++p;
        if (p >= end) return false;
        if (*p == '0') {
            ++p;
        } else if (*p >= '1' && *p <= '9') {
            while (p < end && *p >= '0' && *p <= '9') ++p;
        } else {
            return false;
        }
        bool is_float = false;
        if (p < end && *p == '.') {
            ++p;
            if (p >= end || *p < '0' || *p > '9') return false;
            while (p < end && *p >= '0' && *p <= '9') ++p;
            is_float = true;
        }
        if (p < end && (*p == 'e' || *p == 'E')) {
            const char *q = p + 1;
            if (q < end && (*q == '+' || *q == '-')) ++q;
            if (q < end && *q >= '0' && *q <= '9') {
                while (q < end && *q >= '0' && *q <= '9') ++q;
                p = q;
                is_float = true;
            }
        }
        t.p = s;
        t.len = (uint32_t)(p - s);
        t.type = is_float ? T_FLOAT : T_INT;
        if (kc::js::to_double((const uint8_t *)s, t.len, t.num)) {
            // the exact integer-arithmetic conversion of the device path (kc_jsoncore.cuh: <= 19 significant digits, checked against
            // CPython on 280 k texts), several times faster than strtod; texts beyond its range fall through to strtod
        } else if (t.len < 40) {  // strtod needs a terminated buffer
            char buf[40];
            memcpy(buf, s, t.len);
            buf[t.len] = 0;
            t.num = strtod(buf, nullptr);
        } else {
            t.num = strtod(std::string(s, t.len).c_str(), nullptr);
        }
        return true;
    }
    bool value(Tok &t, int depth = 0) {
        ws();
        if (p >= end) return false;
        const char c = *p;
        if (c == '"') {
            t.type = T_STR;
            return string(t.p, t.len, t.esc);
        }
        if (c == '{' || c == '[') {
            t.type = T_NESTED;
            t.p = p;
            const bool ok = skip_nested(depth);
            t.len = (uint32_t)(p - t.p);
            return ok;
        }
        if (c == 't' && lit("true", 4)) { t.type = T_TRUE; return true; }
        if (c == 'f' && lit("false", 5)) { t.type = T_FALSE; return true; }
        if (c == 'n' && lit("null", 4)) { t.type = T_NULL; return true; }
        if (c == 'N' && lit("NaN", 3)) { t.type = T_FLOAT; t.num = NAN; return true; }
        if (c == 'I' && lit("Infinity", 8)) { t.type = T_FLOAT; t.num = INFINITY; return true; }
        if (c == '-' && lit("-Infinity", 9)) { t.type = T_FLOAT; t.num = -INFINITY; return true; }
        return number(t);
    }
};

// json.loads(text) for a flat object; anything else that json.loads would ACCEPT (top-level list, number, ...) is
// reported through `not_object`; a parse failure means the reference wraps the text (consolidation.py:37-38).
// `odd` is set for things this fast path does not model (escaped keys): the record goes to the Python path.
bool parse_object(const char *s, size_t len, std::vector<Item> &out, bool &not_object, bool &non_ascii, bool &odd) {
    Scanner sc{s, s + len};
    not_object = false;
    out.clear();
    sc.ws();
    if (sc.p >= sc.end) return false;
    if (*sc.p != '{') {
        Tok t;
        const bool ok = sc.value(t);
        sc.ws();
        if (ok && sc.p == sc.end) not_object = true;
        non_ascii = sc.non_ascii;
        return ok && sc.p == sc.end;
    }
    ++sc.p;
    sc.ws();
    if (sc.p < sc.end && *sc.p == '}') {
        ++sc.p;
    } else {
        for (;;) {
            sc.ws();
            if (sc.p >= sc.end || *sc.p != '"') return false;
            const char *kp;
            uint32_t kl;
            bool kesc;
            if (!sc.string(kp, kl, kesc)) return false;
            if (kesc) odd = true;
            sc.ws();
            if (sc.p >= sc.end || *sc.p != ':') return false;
            ++sc.p;
            Item it;
            it.key = std::string_view(kp, kl);
            if (!sc.value(it.tok)) return false;
            out.push_back(it);
            sc.ws();
            if (sc.p < sc.end && *sc.p == ',') {
                ++sc.p;
                continue;
            }
            if (sc.p < sc.end && *sc.p == '}') {
                ++sc.p;
                break;
            }
            return false;
        }
    }
    sc.ws();
    non_ascii = sc.non_ascii;
    return sc.p == sc.end;
}

// ---------------------------------------------------------------- CPython text formats

// float.__repr__: shortest round-trip digits; fixed notation for -4 <= exponent10 < 16, else d[.ddd]e+XX.
void py_float_repr(double x, std::string &out) {
    if (std::isnan(x)) { out += "nan"; return; }
    if (std::isinf(x)) { out += x < 0 ? "-inf" : "inf"; return; }
    char buf[40];
    auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::scientific);  // shortest digits: d[.ddd]e+XX
    std::string_view sv(buf, (size_t)(r.ptr - buf));
    size_t i = 0;
    if (sv[i] == '-') { out.push_back('-'); ++i; }
    const size_t epos = sv.find('e');
    std::string digits;
    for (size_t k = i; k < epos; ++k)
        if (sv[k] != '.') digits.push_back(sv[k]);
    const int exp10 = atoi(std::string(sv.substr(epos + 1)).c_str());
    const int decpt = exp10 + 1;  // position of the decimal point relative to the digit string
    const int nd = (int)digits.size();
    if (decpt > 16 || decpt < -3) {
        out.push_back(digits[0]);
        if (nd > 1) {
            out.push_back('.');
            out.append(digits, 1, std::string::npos);
        }
        out.push_back('e');
        const int e = decpt - 1;
        out.push_back(e < 0 ? '-' : '+');
        const int ae = e < 0 ? -e : e;
        if (ae < 10) out.push_back('0');
        out += std::to_string(ae);
    } else if (decpt <= 0) {
        out += "0.";
        out.append((size_t)(-decpt), '0');
        out += digits;
    } else if (decpt >= nd) {
        out += digits;
        out.append((size_t)(decpt - nd), '0');
        out += ".0";
    } else {
        out.append(digits, 0, (size_t)decpt);
        out.push_back('.');
        out.append(digits, (size_t)decpt, std::string::npos);
    }
}

// json.dumps float: repr, but NaN / Infinity / -Infinity spelled the JSON way
void json_float(double x, std::string &out) {
    if (std::isnan(x)) out += "NaN";
    else if (std::isinf(x)) out += x < 0 ? "-Infinity" : "Infinity";
    else py_float_repr(x, out);
}

// json.dumps(str) with ensure_ascii=True (input is ASCII)
void json_string(std::string_view s, std::string &out) {
    static const char *hex = "0123456789abcdef";
    out.push_back('"');
    for (unsigned char c : s) {
        switch (c) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            case '\b': out += "\\b"; break;
            case '\f': out += "\\f"; break;
            default:
                if (c < 0x20) {
                    out += "\\u00";
                    out.push_back(hex[c >> 4]);
                    out.push_back(hex[c & 15]);
                } else {
                    out.push_back((char)c);
                }
        }
    }
    out.push_back('"');
}

// str(int) of a JSON integer token: the digits as written (JSON forbids leading zeros), except "-0" -> "0"
void int_text(const Tok &t, std::string &out) {
    if (t.len == 2 && t.p[0] == '-' && t.p[1] == '0') out += "0";
    else out.append(t.p, t.len);
}

// str(v) as Python prints the value (for the enum-likeness test and for sanitising)
void py_str(const Tok &t, std::string &out) {
    switch (t.type) {
        case T_TRUE: out += "True"; break;
        case T_FALSE: out += "False"; break;
        case T_INT: int_text(t, out); break;
        case T_FLOAT: py_float_repr(t.num, out); break;
        case T_STR: {
            if (t.esc) {
                std::string tmp;
                unescape(t.p, t.len, tmp);
                out += tmp;
            } else {
                out.append(t.p, t.len);
            }
            break;
        }
        default: break;
    }
}

int word_count(const std::string &s) {  // len(s.strip().split())
    int n = 0;
    bool in = false;
    for (unsigned char c : s) {
        const bool sp = is_py_space(c);
        if (!sp && !in) ++n;
        in = !sp;
    }
    return n;
}

void sanitize(const std::string &s, std::string &out) {  // consensus_utils.py:925-933 on ASCII
    out.resize(s.size());  // at most as long: written in place, trimmed once (no per-character capacity checks)
    char *o = out.data();
    size_t k = 0;
    for (unsigned char c : s) {
        if (c >= 'A' && c <= 'Z') c = (unsigned char)(c + 32);
        o[k] = (char)c;
        k += ((c >= 'a' && c <= 'z') || (c >= '0' && c <= '9')) ? 1 : 0;
    }
    out.resize(k);
}

void json_value(const Tok &t, std::string &out) {
    switch (t.type) {
        case T_TRUE: out += "true"; break;
        case T_FALSE: out += "false"; break;
        case T_INT: int_text(t, out); break;
        case T_FLOAT: json_float(t.num, out); break;
        case T_STR: {
            std::string tmp;
            tok_string(t, tmp);
            json_string(tmp, out);
            break;
        }
        default: out += "null"; break;
    }
}

// ---------------------------------------------------------------- per-record plan

enum GroupKind : uint8_t { G_ALLNULL = 0, G_VOTE_STR, G_VOTE_BOOL, G_NUMERIC, G_MEDOID };

struct Group {
    GroupKind kind;
    std::string_view key;
    int64_t row = -1;       // row in the vote / numeric cell matrix, or the medoid group index
    uint32_t m_first = 0;   // G_MEDOID: first string of the group in Record::mlen, and how many (the non-None cells)
    uint32_t m_count = 0;
};

// Shape of the consensus object: a dict node lists its children (sorted keys), a leaf points at its group.
struct Node {
    std::string_view key;
    int32_t group = -1;        // >= 0: leaf
    bool is_list = false;      // a list node: children are its columns, in order (no keys)
    std::vector<int32_t> kids; // dict / list: indices into Record::nodes
};

struct Record {
    uint8_t status = 0;        // 0 native, 1 needs the Python path
    std::vector<Node> nodes;   // nodes[0] is the root dict
    std::vector<Group> groups;
    std::vector<Tok> cells;    // groups.size() * n tokens, group-major (T_MISSING / T_NULL count as None)
    std::string mchars;        // normalize_string() of the cells of the medoid groups, back to back
    std::vector<int32_t> mlen; // their lengths
    std::shared_ptr<void> tree;  // records with lists: the aligned value tree the cells and keys point into
};

const double kF64None = [] { const uint64_t b = KC_F64_NONE_BITS; double d; memcpy(&d, &b, 8); return d; }();

// Scalar field: which kernel decides it (cu:1405-1411 vote, cu:1443-1453 numeric / medoid).  False: Python path.
bool plan_leaf(Record &rec, Group &g, const Tok *cells, int n, std::string &tmp) {
    const Tok *first = nullptr;
    for (int c = 0; c < n && !first; ++c)
        if (cells[c].type > T_NULL) first = &cells[c];
    if (!first) {
        g.kind = G_ALLNULL;
    } else if (first->type == T_NESTED) {
        rec.status = 1;
        return false;
    } else if (first->type == T_STR || first->type == T_TRUE || first->type == T_FALSE) {
        bool multi_word = false, all_str = true;
        int live = 0;
        for (int c = 0; c < n; ++c) {
            const Tok &t = cells[c];
            if (t.type <= T_NULL) continue;
            ++live;
            if (t.type == T_NESTED) {  // str(dict) is almost never enum-like: leave it to Python
                rec.status = 1;
                return false;
            }
            all_str &= t.type == T_STR;
            if (t.type == T_STR) {  // numbers and bools print as one word
                tmp.clear();
                py_str(t, tmp);
                multi_word |= word_count(tmp) >= 3;
            }
        }
        if (multi_word) {
            // Not enum-like (cu:1405): the similarity medoid of consensus_as_primitive (cu:1221-1237).  K4 takes it when
            // every pair is a Levenshtein pair inside its contract (same rule as columnar.Plan._medoid_on_device under
            // the default string_similarity_method "embeddings"); anything else goes to the Python path.
            if (!all_str) {
                rec.status = 1;
                return false;
            }
            g.kind = G_MEDOID;
            g.m_first = (uint32_t)rec.mlen.size();
            g.m_count = (uint32_t)live;
            if (live >= 2) {
                int long_raw = 0, long_norm = 0;
                thread_local std::string norm;
                for (int c = 0; c < n; ++c) {
                    const Tok &t = cells[c];
                    if (t.type != T_STR) continue;
                    tmp.clear();
                    py_str(t, tmp);
                    sanitize(tmp, norm);  // == normalize_string (cu:660-673) on ASCII text
                    long_raw += tmp.size() > 50;
                    long_norm += norm.size() > 64;
                    if (norm.size() > 2000 || long_raw > 1 || long_norm > 1) {
                        rec.status = 1;
                        return false;
                    }
                    rec.mchars += norm;
                    rec.mlen.push_back((int32_t)norm.size());
                }
            }
            return true;
        }
        if (first->type == T_STR) {
            g.kind = G_VOTE_STR;
        } else {
            for (int c = 0; c < n; ++c)  // `v or False` on non-bool values compares Python objects: Python path
                if (cells[c].type > T_FALSE) {
                    rec.status = 1;
                    return false;
                }
            g.kind = G_VOTE_BOOL;
        }
    } else {
        g.kind = G_NUMERIC;
    }
    return true;
}

constexpr int kMaxDepth = 16;

// One dict level of the alignment pre-pass + dispatcher (cu:516-548 then cu:1414-1426): `items[c]` is candidate c's
// (key-sorted) dict at this node, or nullptr where it has none — the pre-pass turns None into a dict of Nones, so after
// it EVERY candidate is a dict here and parent_valid_frac stays 1.  Keys are visited in sorted order; a key whose values
// are all objects recurses, any list (or a mix of objects and scalars) sends the record to the Python path.
struct LevelScratch {  // per recursion depth, reused across records (no allocation in the steady state)
    std::vector<std::string_view> keys;
    std::vector<size_t> cursor;
    std::vector<Tok> cells;
    std::vector<std::vector<Item>> child_items;
    std::vector<const std::vector<Item> *> child;
};

void plan_dict(Record &rec, int32_t node, const std::vector<Item> *const *items, int n, int depth) {
    thread_local std::vector<LevelScratch> levels(kMaxDepth + 1);  // sized once: references stay valid through the recursion
    thread_local std::string tmp;
    LevelScratch &L = levels[(size_t)depth];
    std::vector<std::string_view> &keys = L.keys;
    keys.clear();
    for (int c = 0; c < n; ++c)
        if (items[c])
            for (auto &it : *items[c]) keys.push_back(it.key);
    std::sort(keys.begin(), keys.end());  // code-point order == byte order for ASCII (consensus_utils.py:521-522)
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    std::vector<size_t> &cursor = L.cursor;
    cursor.assign((size_t)n, 0);
    std::vector<Tok> &cells = L.cells;
    cells.resize((size_t)n);
    for (size_t ki = 0; ki < keys.size(); ++ki) {
        const std::string_view key = keys[ki];
        for (int c = 0; c < n; ++c) {  // merge: both sides are sorted by key; of duplicates the last one wins
            cells[(size_t)c] = Tok();
            if (!items[c]) continue;
            const auto &its = *items[c];
            size_t &k = cursor[(size_t)c];
            while (k < its.size() && its[k].key == key) cells[(size_t)c] = its[k++].tok;
        }
        if (key.find("reasoning___") != std::string_view::npos || key.find("source___") != std::string_view::npos) continue;  // cu:1292
        const Tok *first = nullptr;
        for (int c = 0; c < n && !first; ++c)
            if (cells[(size_t)c].type > T_NULL) first = &cells[(size_t)c];
        if (first && first->type == T_NESTED) {
            if (*first->p != '{') {  // a list: the whole record goes through the alignment pre-pass (tree path below)
                rec.status = 2;
                return;
            }
            if (depth + 1 >= kMaxDepth) {
                rec.status = 1;
                return;
            }
            auto &child_items = L.child_items;
            if ((int)child_items.size() < n) child_items.resize((size_t)n);
            auto &child = L.child;
            child.assign((size_t)n, nullptr);
            for (int c = 0; c < n; ++c) {
                const Tok &t = cells[(size_t)c];
                if (t.type <= T_NULL) continue;
                if (t.type != T_NESTED || *t.p != '{') {  // not all of one type: the pre-pass leaves the values alone (cu:507-512)
                    rec.status = 1;
                    return;
                }
                bool not_object = false, non_ascii = false, odd = false;
                if (!parse_object(t.p, t.len, child_items[(size_t)c], not_object, non_ascii, odd) || not_object || non_ascii || odd) {
                    rec.status = 1;
                    return;
                }
                std::stable_sort(child_items[(size_t)c].begin(), child_items[(size_t)c].end(),
                                 [](const Item &a, const Item &b) { return a.key < b.key; });
                child[(size_t)c] = &child_items[(size_t)c];
            }
            const int32_t kid = (int32_t)rec.nodes.size();
            rec.nodes.emplace_back();
            rec.nodes[(size_t)kid].key = key;
            rec.nodes[(size_t)node].kids.push_back(kid);
            plan_dict(rec, kid, child.data(), n, depth + 1);
            if (rec.status) return;
            continue;
        }
        const size_t base = rec.cells.size();
        rec.cells.insert(rec.cells.end(), cells.begin(), cells.end());
        const Tok *gcells = &rec.cells[base];
        Group g;
        g.key = key;
        if (!plan_leaf(rec, g, gcells, n, tmp)) return;
        const int32_t kid = (int32_t)rec.nodes.size();
        rec.nodes.emplace_back();
        rec.nodes[(size_t)kid].key = key;
        rec.nodes[(size_t)kid].group = (int32_t)rec.groups.size();
        rec.nodes[(size_t)node].kids.push_back(kid);
        rec.groups.push_back(g);
    }
}

void plan_record_tree(const char *const *texts, const int64_t *lens, int n, Record &rec);

void plan_record(const char *const *texts, const int64_t *lens, int n, Record &rec) {
    thread_local std::vector<std::vector<Item>> cands;
    if ((int)cands.size() < n) cands.resize((size_t)n);
    static const char kTextKey[] = "text";
    for (int c = 0; c < n; ++c) {
        const char *s = texts[c];
        const size_t len = lens ? (size_t)lens[c] : strlen(s);
        if (len == 0) {  // `if choice.message.content:` drops empty contents, changing n (consolidation.py:92): Python path
            rec.status = 1;
            return;
        }
        bool not_object = false, non_ascii = false, odd = false;
        std::vector<Item> &items = cands[(size_t)c];
        const bool ok = parse_object(s, len, items, not_object, non_ascii, odd);
        if (!ok || !non_ascii)  // a failed parse may have stopped early: look at every byte
            for (size_t i = 0; i < len && !non_ascii; ++i)
                if ((unsigned char)s[i] >= 0x80) non_ascii = true;
        if (non_ascii || odd || (ok && not_object)) {
            rec.status = 1;
            return;
        }
        if (!ok) {  // {"text": content}: the whole text is the (already unescaped) value
            items.clear();
            Item it;
            it.key = std::string_view(kTextKey, 4);
            it.tok.type = T_STR;
            it.tok.p = s;
            it.tok.len = (uint32_t)len;
            items.push_back(it);
        }
        // sort by key; of duplicates the LAST one in the text wins (dict construction)
        std::stable_sort(items.begin(), items.end(), [](const Item &a, const Item &b) { return a.key < b.key; });
    }
    rec.groups.clear();
    rec.cells.clear();
    rec.mchars.clear();
    rec.mlen.clear();
    rec.nodes.clear();
    rec.nodes.emplace_back();
    thread_local std::vector<const std::vector<Item> *> top;
    top.assign((size_t)n, nullptr);
    for (int c = 0; c < n; ++c) top[(size_t)c] = &cands[(size_t)c];
    plan_dict(rec, 0, top.data(), n, 0);
    if (rec.status == 2) plan_record_tree(texts, lens, n, rec);  // list fields: align first (H2), then plan on the aligned tree
}

void encode_vote(GroupKind kind, const Tok *toks, int n, int8_t *cells) {
    thread_local std::vector<std::string> seen;
    thread_local std::string tmp, san;
    size_t n_seen = 0;
    for (int c = 0; c < n; ++c) {
        const Tok &t = toks[c];
        if (kind == G_VOTE_BOOL) {
            cells[c] = (t.type == T_TRUE) ? 1 : 0;  // None and False -> False (cu:956)
            continue;
        }
        if (t.type <= T_NULL) {
            cells[c] = KC_CODE_NONE;
            continue;
        }
        tmp.clear();
        py_str(t, tmp);
        sanitize(tmp, san);
        size_t k = 0;
        while (k < n_seen && seen[k] != san) ++k;
        if (k == n_seen) {
            if (seen.size() <= n_seen) seen.emplace_back();
            seen[n_seen++] = san;
        }
        cells[c] = (int8_t)k;
    }
}

void encode_numeric(const Tok *toks, int n, double *cells) {
    for (int c = 0; c < n; ++c) {
        const Tok &t = toks[c];
        if (t.type <= T_NULL) cells[c] = kF64None;
        else if (t.type == T_INT || t.type == T_FLOAT) cells[c] = t.num;  // non-finite values are dropped by the kernel
        else cells[c] = NAN;                                               // bool / str / nested: counted, never clustered
    }
}

// CPython round(x, 5) on the host (same algorithm as kc::py_round5): exact value * 1e5, half-even, one division
double py_round5(double x) {
    if (!(x > 0.0) || !std::isfinite(x)) return x;
    uint64_t bits;
    memcpy(&bits, &x, 8);
    const int biased = (int)((bits >> 52) & 0x7FF);
    uint64_t mant = bits & 0xFFFFFFFFFFFFFull;
    int exp2;
    if (biased == 0) exp2 = -1074;
    else { mant |= 1ull << 52; exp2 = biased - 1075; }
    if (exp2 >= 0) return x;
    const int sh = -exp2;
    if (sh >= 128) return 0.0;
    const unsigned __int128 prod = (unsigned __int128)mant * 100000u;
    unsigned __int128 q = prod >> sh;
    const unsigned __int128 rem = prod - (q << sh);
    const unsigned __int128 half = (unsigned __int128)1 << (sh - 1);
    if (rem > half || (rem == half && (q & 1))) ++q;
    return (double)(uint64_t)q / 100000.0;
}

struct EmitCtx {
    const Record &rec;
    int n;
    const uint32_t *vmeta;
    const double *nvalue;
    const uint32_t *nmeta;
    const int32_t *midx;
    const double *mavg;
};

// value and confidence of one leaf group (the epilogue of cu:971-982, cu:1085-1086, cu:1116, cu:1177-1219, cu:1233-1237)
void emit_leaf(const EmitCtx &cx, size_t gi, std::string &content, std::string &lik) {
    const Record &rec = cx.rec;
    const int n = cx.n;
    const uint32_t *vmeta = cx.vmeta, *nmeta = cx.nmeta;
    const double *nvalue = cx.nvalue, *mavg = cx.mavg;
    const int32_t *midx = cx.midx;
    const Group &g = rec.groups[gi];
    const Tok *cells = &rec.cells[gi * (size_t)n];
        double conf = 0.0;
        Tok value;  // T_MISSING == None
        if (g.kind == G_VOTE_STR || g.kind == G_VOTE_BOOL) {
            const uint32_t m = vmeta[g.row];
            const uint32_t idx = KC_META_IDX(m), support = KC_META_SUPPORT(m), present = KC_META_PRESENT(m);
            if (g.kind == G_VOTE_BOOL) {
                value.type = (cells[idx].type == T_TRUE) ? T_TRUE : T_FALSE;  // the processed key (cu:958)
            } else {
                value = cells[idx];  // first original whose sanitised form wins (cu:971)
            }
            conf = py_round5(1.0 * ((double)support / (double)present));
        } else if (g.kind == G_NUMERIC) {
            const uint32_t m = nmeta[g.row];
            const uint32_t idx = KC_META_IDX(m), support = KC_META_SUPPORT(m), nn = KC_META_NN(m), present = KC_META_PRESENT(m);
            const uint32_t flags = KC_META_FLAGS(m);
            if (flags & KC_FLAG_HAS_VALUE) {
                if (flags & KC_FLAG_SINGLE) {
                    value = cells[idx];
                    conf = 1.0 * (1.0 / (double)present) * (1.0 / 1.0);
                } else {
                    value.type = T_FLOAT;
                    value.num = nvalue[g.row];
                    conf = py_round5((double)support / (double)nn);
                }
            } else if (flags & KC_FLAG_NO_FINITE) {
                conf = 1.0 * ((double)nn / (double)present);
            } else {
                conf = present == 0 ? 1.0 : 0.0;
            }
        } else if (g.kind == G_MEDOID) {
            // cu:1444 then cu:1085-1086 (one non-None cell, unrounded) or cu:1233-1237 (the medoid, rounded)
            const double sub = 1.0 * ((double)g.m_count / (double)n);
            int want = g.m_count >= 2 ? midx[g.row] : 0;
            for (int c = 0; c < n; ++c) {
                if (cells[c].type <= T_NULL) continue;
                if (want-- == 0) {
                    value = cells[c];
                    break;
                }
            }
            conf = g.m_count >= 2 ? py_round5(sub * mavg[g.row]) : sub * (1.0 / 1.0);
        }  // G_ALLNULL: None, 0.0 (cu:1401-1402)
        json_value(value, content);
        json_float(conf, lik);
}

void emit_node(const EmitCtx &cx, int32_t ni, std::string &content, std::string &lik) {
    const Node &node = cx.rec.nodes[(size_t)ni];
    if (node.group >= 0) {
        emit_leaf(cx, (size_t)node.group, content, lik);
        return;
    }
    content += node.is_list ? "[" : "{";
    lik += node.is_list ? "[" : "{";
    bool first = true;
    for (int32_t kid : node.kids) {
        if (!first) {
            content += ", ";
            lik += ", ";
        }
        first = false;
        if (!node.is_list) {
            const std::string_view key = cx.rec.nodes[(size_t)kid].key;
            json_string(key, content);
            json_string(key, lik);
            content += ": ";
            lik += ": ";
        }
        emit_node(cx, kid, content, lik);
    }
    content += node.is_list ? "]" : "}";
    lik += node.is_list ? "]" : "}";
}

void emit_record(const Record &rec, int n, const uint32_t *vmeta, const double *nvalue, const uint32_t *nmeta, const int32_t *midx,
                 const double *mavg, std::string &content, std::string &lik) {
    const EmitCtx cx{rec, n, vmeta, nvalue, nmeta, midx, mavg};
    content.clear();
    lik.clear();
    emit_node(cx, 0, content, lik);
    // {"text": s} -> s (consolidation.py:55-57): a single top-level string field called "text"
    const Node &root = rec.nodes[0];
    if (root.kids.size() == 1) {
        const Node &only = rec.nodes[(size_t)root.kids[0]];
        if (only.group >= 0 && only.key == "text") {
            const Group &g = rec.groups[(size_t)only.group];
            const Tok *cells = &rec.cells[(size_t)only.group * (size_t)n];
            const Tok *picked = nullptr;
            if (g.kind == G_VOTE_STR) {
                picked = &cells[KC_META_IDX(vmeta[g.row])];
            } else if (g.kind == G_MEDOID) {
                int want = g.m_count >= 2 ? midx[g.row] : 0;
                for (int c = 0; c < n && !picked; ++c)
                    if (cells[c].type > T_NULL && want-- == 0) picked = &cells[c];
            }
            if (picked && picked->type == T_STR) {
                content.clear();
                tok_string(*picked, content);
            }
        }
    }
}

char *dup_string(const std::string &s) {
    char *p = (char *)malloc(s.size() + 1);
    if (p) memcpy(p, s.c_str(), s.size() + 1);
    return p;
}

template <typename F>
void parallel_for(int64_t n, int threads, F fn) {
    std::atomic<int64_t> next{0};
    const int64_t grain = 256;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&] {
            for (;;) {
                const int64_t b = next.fetch_add(grain);
                if (b >= n) return;
                const int64_t e = std::min(n, b + grain);
                for (int64_t i = b; i < e; ++i) fn(i);
            }
        });
    for (auto &th : pool) th.join();
}


#include "kc_align.inl"  // H2: value tree, similarities, assignment, lists_alignment, recursive_list_alignments

// ---------------------------------------------------------------- H1 for records with list fields: plan on the aligned tree

Tok tok_of(const AVal &v) {  // a scalar of the tree as the token the leaf planner / encoders / emitters work on
    static const char kBrace[] = "{";
    Tok t;
    switch (v.t) {
        case A_NONE: t.type = T_NULL; break;
        case A_BOOL: t.type = v.b ? T_TRUE : T_FALSE; break;
        case A_INT: t.type = T_INT; t.p = v.s.data(); t.len = (uint32_t)v.s.size(); t.num = v.num; break;
        case A_FLOAT: t.type = T_FLOAT; t.num = v.num; break;
        case A_STR: t.type = T_STR; t.p = v.s.data(); t.len = (uint32_t)v.s.size(); break;  // already unescaped
        default: t.type = T_NESTED; t.p = kBrace; t.len = 1; break;
    }
    return t;
}

// The dispatcher (cu:1376-1454) over ALIGNED candidates: after the pre-pass every candidate is a dict with the same sorted
// keys at a dict node and a list of the same width at a list node (cu:516-548, 550-613), so parent_valid_frac stays 1 and a
// node is a dict, a list or a scalar field.  Anything the pre-pass left unaligned (mixed types) goes to the Python path.
void plan_tree_value(Record &rec, AlignCtx &cx, int32_t node, const std::vector<int32_t> &ids, int n, int depth, std::string &tmp) {
    if (depth > 48) {
        rec.status = 1;
        return;
    }
    const AVal *first = nullptr;
    for (int c = 0; c < n && !first; ++c)
        if (cx.tr.v[(size_t)ids[(size_t)c]].t != A_NONE) first = &cx.tr.v[(size_t)ids[(size_t)c]];
    if (first && (first->t == A_DICT || first->t == A_LIST)) {
        const AType ft = first->t;
        for (int c = 0; c < n; ++c)
            if (cx.tr.v[(size_t)ids[(size_t)c]].t != ft) {  // not what the pre-pass produces from uniform input
                rec.status = 1;
                return;
            }
        std::vector<int32_t> child((size_t)n);
        if (ft == A_DICT) {
            const size_t width = first->kv.size();
            for (int c = 0; c < n; ++c)
                if (cx.tr.v[(size_t)ids[(size_t)c]].kv.size() != width) {
                    rec.status = 1;
                    return;
                }
            for (size_t k = 0; k < width; ++k) {
                const std::string &key = cx.tr.v[(size_t)ids[0]].kv[k].first;
                for (int c = 0; c < n; ++c) {
                    const auto &e = cx.tr.v[(size_t)ids[(size_t)c]].kv[k];
                    if (e.first != key) {
                        rec.status = 1;
                        return;
                    }
                    child[(size_t)c] = e.second;
                }
                if (key.find("reasoning___") != std::string::npos || key.find("source___") != std::string::npos) continue;  // cu:1292
                const int32_t kid = (int32_t)rec.nodes.size();
                rec.nodes.emplace_back();
                rec.nodes[(size_t)kid].key = key;
                rec.nodes[(size_t)node].kids.push_back(kid);
                plan_tree_value(rec, cx, kid, child, n, depth + 1, tmp);
                if (rec.status) return;
            }
        } else {
            rec.nodes[(size_t)node].is_list = true;
            const size_t width = first->items.size();
            for (int c = 0; c < n; ++c)
                if (cx.tr.v[(size_t)ids[(size_t)c]].items.size() != width) {
                    rec.status = 1;
                    return;
                }
            for (size_t k = 0; k < width; ++k) {
                for (int c = 0; c < n; ++c) child[(size_t)c] = cx.tr.v[(size_t)ids[(size_t)c]].items[k];
                const int32_t kid = (int32_t)rec.nodes.size();
                rec.nodes.emplace_back();
                rec.nodes[(size_t)node].kids.push_back(kid);
                plan_tree_value(rec, cx, kid, child, n, depth + 1, tmp);
                if (rec.status) return;
            }
        }
        return;
    }
    // a scalar field (or all None)
    const size_t base = rec.cells.size();
    for (int c = 0; c < n; ++c) rec.cells.push_back(tok_of(cx.tr.v[(size_t)ids[(size_t)c]]));
    Group g;
    g.key = rec.nodes[(size_t)node].key;
    if (!plan_leaf(rec, g, &rec.cells[base], n, tmp)) return;
    rec.nodes[(size_t)node].group = (int32_t)rec.groups.size();
    rec.groups.push_back(g);
}

void plan_record_tree(const char *const *texts, const int64_t *lens, int n, Record &rec) {
    rec.status = 0;
    rec.groups.clear();
    rec.cells.clear();
    rec.mchars.clear();
    rec.mlen.clear();
    rec.nodes.clear();
    auto cxp = std::make_shared<AlignCtx>();
    AlignCtx &cx = *cxp;
    std::vector<int32_t> values((size_t)n);
    {
        size_t bytes = 0;
        for (int c = 0; c < n; ++c) bytes += lens ? (size_t)lens[c] : strlen(texts[c]);
        cx.tr.v.reserve(bytes / 6 + 16);  // a value per ~6 bytes of JSON, plus what the alignment adds: no regrowth (moves of every node) while parsing
    }
    for (int c = 0; c < n; ++c) {  // _safe_parse_content (consolidation.py:25-38); the caller already ruled out empty / non-ASCII text
        const size_t len = lens ? (size_t)lens[c] : strlen(texts[c]);
        Scanner sc{texts[c], texts[c] + len};
        const size_t mark = cx.tr.v.size();
        bool ok = aparse(sc, cx.tr, values[(size_t)c], 0);
        if (ok) {
            sc.ws();
            ok = sc.p == sc.end;
        }
        if (ok && cx.tr.v[(size_t)values[(size_t)c]].t != A_DICT) {  // valid JSON but not an object: Python path (as the flat planner does)
            rec.status = 1;
            return;
        }
        if (!ok) {  // {"text": content}
            cx.tr.v.resize(mark);
            const int32_t str = cx.tr.add(A_STR);
            cx.tr.v[(size_t)str].s.assign(texts[c], len);
            const int32_t d = cx.tr.add(A_DICT);
            cx.tr.v[(size_t)d].kv.emplace_back("text", str);
            values[(size_t)c] = d;
        }
    }
    align_values(cx, values, /*min_support_ratio=*/0.51, 0);  // ConsensusSettings default (cu:41); other settings: Python path
    if (cx.decline) {
        rec.status = 1;
        return;
    }
    thread_local std::string tmp;
    rec.nodes.emplace_back();
    plan_tree_value(rec, cx, 0, values, n, 0, tmp);
    if (rec.status == 0) rec.tree = cxp;  // from here on nothing is added to the tree: cells and keys point into it
}

// ---------------------------------------------------------------- a batch of records: plan, encode, emit

struct Batch {
    int n = 0;
    std::vector<Record> recs;
    int64_t gv = 0, gx = 0, gm = 0;       // vote / numeric / medoid groups of the records with status 0
    std::vector<uint8_t> m_chars;         // medoid groups of the whole batch in CSR form for ONE K4 launch
    std::vector<int32_t> m_str_off{0}, m_grp_off{0};
    int32_t m_max_group = 2;
};

int default_threads() { return (int)std::min(32u, std::max(1u, std::thread::hardware_concurrency())); }  // parsing saturates memory / malloc beyond ~32

void plan_batch(Batch &b, const char *const *texts, const int64_t *lens, int64_t n_records, int n, int threads) {
    b.n = n;
    b.recs.assign((size_t)n_records, Record());
    parallel_for(n_records, threads, [&](int64_t r) { plan_record(texts + r * n, lens ? lens + r * n : nullptr, n, b.recs[(size_t)r]); });
    for (auto &rec : b.recs) {
        if (rec.status) continue;
        for (auto &g : rec.groups) {
            if (g.kind == G_VOTE_STR || g.kind == G_VOTE_BOOL) g.row = b.gv++;
            else if (g.kind == G_NUMERIC) g.row = b.gx++;
            else if (g.kind == G_MEDOID && g.m_count >= 2) {
                g.row = b.gm++;
                for (uint32_t k = 0; k < g.m_count; ++k) b.m_str_off.push_back(b.m_str_off.back() + rec.mlen[g.m_first + k]);
                b.m_grp_off.push_back(b.m_grp_off.back() + (int32_t)g.m_count);
                b.m_max_group = std::max(b.m_max_group, (int32_t)g.m_count);
            }
        }
        if (!rec.mchars.empty()) b.m_chars.insert(b.m_chars.end(), rec.mchars.begin(), rec.mchars.end());
    }
}

void encode_batch(const Batch &b, int8_t *codes, double *vals, int threads) {
    const int n = b.n;
    parallel_for((int64_t)b.recs.size(), threads, [&](int64_t r) {
        const Record &rec = b.recs[(size_t)r];
        if (rec.status) return;
        for (size_t gi = 0; gi < rec.groups.size(); ++gi) {
            const Group &g = rec.groups[gi];
            const Tok *toks = &rec.cells[gi * (size_t)n];
            if (g.kind == G_VOTE_STR || g.kind == G_VOTE_BOOL) encode_vote(g.kind, toks, n, codes + g.row * n);
            else if (g.kind == G_NUMERIC) encode_numeric(toks, n, vals + g.row * n);
        }
    });
}

void emit_batch(const Batch &b, const uint32_t *vmeta, const double *nvalue, const uint32_t *nmeta, const int32_t *midx, const double *mavg,
                int threads, char **out_content, char **out_likelihoods, uint8_t *out_status) {
    parallel_for((int64_t)b.recs.size(), threads, [&](int64_t r) {
        const Record &rec = b.recs[(size_t)r];
        out_status[r] = rec.status;
        out_content[r] = nullptr;
        out_likelihoods[r] = nullptr;
        if (rec.status) return;
        std::string content, lik;
        emit_record(rec, b.n, vmeta, nvalue, nmeta, midx, mavg, content, lik);
        out_content[r] = dup_string(content);
        out_likelihoods[r] = dup_string(lik);
    });
}

}  // namespace

extern "C" {

int kc_consolidate_json(const char *const *texts, const int64_t *lens, int64_t n_records, int32_t n, double rel_eps,
                        double abs_eps, int device, int32_t threads, char **out_content, char **out_likelihoods,
                        uint8_t *out_status) {
    if (n < 2 || n > KC_MAX_CANDIDATES || n_records < 0 || !texts || !out_content || !out_likelihoods || !out_status)
        return KC_EINVAL;
    if (threads <= 0) threads = default_threads();
    const bool timing = getenv("KC_JSON_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto t0 = now();
    Batch batch;
    plan_batch(batch, texts, lens, n_records, n, threads);
    const auto t1 = now();
    const int64_t gv = batch.gv, gx = batch.gx, gm = batch.gm;
    std::vector<int32_t> m_idx((size_t)gm);
    std::vector<double> m_avg((size_t)gm);
    // page-locked staging buffers are expensive to create: keep them (grow-only) across calls
    static std::mutex pool_mu;
    static struct { void *p = nullptr; size_t cap = 0; } pool[6];
    std::lock_guard<std::mutex> pool_lock(pool_mu);  // also serialises callers (one staging pool per process)
    auto pinned = [&](int slot, size_t bytes) -> void * {
        if (bytes == 0) return nullptr;
        if (pool[slot].cap < bytes) {
            kc_host_free(pool[slot].p);
            pool[slot].cap = bytes + bytes / 4;
            pool[slot].p = kc_host_alloc(pool[slot].cap);
            if (!pool[slot].p) pool[slot].cap = 0;
        }
        return pool[slot].p;
    };
    int8_t *h_codes = (int8_t *)pinned(0, (size_t)gv * n);
    double *h_vals = (double *)pinned(1, (size_t)gx * n * 8);
    int32_t *h_win = (int32_t *)pinned(2, (size_t)gv * 4);
    uint32_t *h_vmeta = (uint32_t *)pinned(3, (size_t)gv * 4);
    double *h_value = (double *)pinned(4, (size_t)gx * 8);
    uint32_t *h_nmeta = (uint32_t *)pinned(5, (size_t)gx * 4);
    const auto t2 = now();
    auto t3 = t2, t4 = t2;
    int rc = KC_OK;
    if ((gv && (!h_codes || !h_win || !h_vmeta)) || (gx && (!h_vals || !h_value || !h_nmeta))) rc = KC_ENOMEM;
    if (!rc) {
        encode_batch(batch, h_codes, h_vals, threads);
        t3 = now();
        // one "field" per group: the two halves are independent calls of the host-buffer entry
        if (gv) rc = kc_consensus_host_i8(h_codes, 1, nullptr, nullptr, 0, gv, n, rel_eps, abs_eps, h_win, h_vmeta, nullptr, nullptr, device, nullptr);
        if (!rc && gx) rc = kc_consensus_host_i8(nullptr, 0, nullptr, h_vals, 1, gx, n, rel_eps, abs_eps, nullptr, nullptr, h_value, h_nmeta, device, nullptr);
        if (!rc && gm)
            rc = kc_medoid_str_host(batch.m_chars.data(), (int64_t)batch.m_chars.size(), batch.m_str_off.data(), batch.m_grp_off.data(), gm,
                                    batch.m_max_group, m_idx.data(), m_avg.data(), device);
    }
    t4 = now();
    if (!rc) emit_batch(batch, h_vmeta, h_value, h_nmeta, m_idx.data(), m_avg.data(), threads, out_content, out_likelihoods, out_status);
    const auto t5 = now();
    if (timing)
        fprintf(stderr, "kc_consolidate_json: parse+plan %.1f ms, rows+alloc %.1f ms, encode %.1f ms, gpu %.1f ms, emit %.1f ms (%d threads)\n",
                ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4), ms(t4, t5), threads);
    return rc;
}

// The same in two phases, for callers that run K1 / K2 / K4 themselves (their own streams, another device, a test that puts
// a checker in their place): kc_json_plan parses, plans and encodes a batch into host arrays the handle owns; the caller
// computes the result columns for them; kc_json_emit turns those into the consensus texts.  The candidate texts must stay
// alive until kc_json_emit returns (cells are views into them).
struct kc_json_batch {
    Batch batch;
    std::vector<int8_t> codes;
    std::vector<double> vals;
    int threads;
};

int kc_json_plan(const char *const *texts, const int64_t *lens, int64_t n_records, int32_t n, int32_t threads, kc_json_batch **out) {
    if (n < 2 || n > KC_MAX_CANDIDATES || n_records < 0 || !texts || !out) return KC_EINVAL;
    if (threads <= 0) threads = default_threads();
    kc_json_batch *h = new (std::nothrow) kc_json_batch;
    if (!h) return KC_ENOMEM;
    h->threads = threads;
    plan_batch(h->batch, texts, lens, n_records, n, threads);
    h->codes.resize((size_t)h->batch.gv * n);
    h->vals.resize((size_t)h->batch.gx * n);
    encode_batch(h->batch, h->codes.data(), h->vals.data(), threads);
    *out = h;
    return KC_OK;
}

int kc_json_inputs(const kc_json_batch *h, const int8_t **vote_cells, int64_t *n_vote_groups, const double **num_cells,
                   int64_t *n_num_groups, const uint8_t **medoid_chars, const int32_t **medoid_str_off, const int32_t **medoid_grp_off,
                   int64_t *n_medoid_groups, int32_t *max_medoid_group) {
    if (!h) return KC_EINVAL;
    if (vote_cells) *vote_cells = h->codes.data();
    if (n_vote_groups) *n_vote_groups = h->batch.gv;
    if (num_cells) *num_cells = h->vals.data();
    if (n_num_groups) *n_num_groups = h->batch.gx;
    if (medoid_chars) *medoid_chars = h->batch.m_chars.data();
    if (medoid_str_off) *medoid_str_off = h->batch.m_str_off.data();
    if (medoid_grp_off) *medoid_grp_off = h->batch.m_grp_off.data();
    if (n_medoid_groups) *n_medoid_groups = h->batch.gm;
    if (max_medoid_group) *max_medoid_group = h->batch.m_max_group;
    return KC_OK;
}

int kc_json_emit(kc_json_batch *h, const uint32_t *vote_meta, const double *num_value, const uint32_t *num_meta, const int32_t *medoid_idx,
                 const double *medoid_avg, char **out_content, char **out_likelihoods, uint8_t *out_status) {
    if (!h || !out_content || !out_likelihoods || !out_status) return KC_EINVAL;
    if ((h->batch.gv && !vote_meta) || (h->batch.gx && (!num_value || !num_meta)) || (h->batch.gm && (!medoid_idx || !medoid_avg))) return KC_EINVAL;
    emit_batch(h->batch, vote_meta, num_value, num_meta, medoid_idx, medoid_avg, h->threads, out_content, out_likelihoods, out_status);
    return KC_OK;
}

void kc_json_free(kc_json_batch *h) { delete h; }

// H2: recursive_list_alignments(values, "embeddings", embed, client, min_support_ratio)[0] for ONE record of n candidate
// values given as JSON texts (null = None): out_texts[c] = json.dumps of candidate c's aligned value (free with
// kc_free_strings).  Returns 0; 1 if the record needs the Python path (a pair of strings both longer than 50 characters
// would be compared through embeddings, cu:813; non-ASCII text); negative for invalid JSON.
int kc_align_json(const char *const *texts, const int64_t *lens, int32_t n, double min_support_ratio, char **out_texts) {
    if (!texts || !out_texts || n < 1) return KC_EINVAL;
    AlignCtx cx;
    std::vector<int32_t> values((size_t)n);
    for (int32_t c = 0; c < n; ++c) {
        out_texts[c] = nullptr;
        const size_t len = lens ? (size_t)lens[c] : strlen(texts[c]);
        Scanner sc{texts[c], texts[c] + len};
        if (!aparse(sc, cx.tr, values[(size_t)c], 0)) return KC_EINVAL;
        sc.ws();
        if (sc.p != sc.end) return KC_EINVAL;
        if (sc.non_ascii) return 1;
        for (size_t i = 0; i < len; ++i)
            if ((unsigned char)texts[c][i] >= 0x80) return 1;
    }
    align_values(cx, values, min_support_ratio, 0);
    if (cx.decline) return 1;
    std::string out;
    for (int32_t c = 0; c < n; ++c) {
        out.clear();
        adump(cx.tr, values[(size_t)c], out);
        out_texts[c] = dup_string(out);
    }
    return 0;
}

// generic_similarity (consensus_utils.py:892-917, default string method) of two JSON values; test hook of the native alignment.
// Returns 0 and *out, 1 if a pair of long strings would need embeddings, negative on invalid JSON.
int kc_debug_similarity_json(const char *a, const char *b, double *out) {
    AlignCtx cx;
    int32_t ia, ib;
    Scanner sa{a, a + strlen(a)}, sb{b, b + strlen(b)};
    if (!aparse(sa, cx.tr, ia, 0) || !aparse(sb, cx.tr, ib, 0) || sa.non_ascii || sb.non_ascii) return KC_EINVAL;
    sa.ws();
    sb.ws();
    if (sa.p != sa.end || sb.p != sb.end) return KC_EINVAL;
    // pointers into the vector: no insertion happens below
    *out = generic_similarity(cx, &cx.tr.v[(size_t)ia], &cx.tr.v[(size_t)ib]);
    return cx.decline ? 1 : 0;
}

// scipy.optimize.linear_sum_assignment(cost) restated (test hook): pairs sorted by row; returns their count or -1.
int kc_debug_lsap(int32_t nr, int32_t nc, const double *cost, int32_t *row_ind, int32_t *col_ind) {
    std::vector<int> rows, cols;
    if (!lsap(nr, nc, cost, rows, cols)) return -1;
    for (size_t k = 0; k < rows.size(); ++k) {
        row_ind[k] = rows[k];
        col_ind[k] = cols[k];
    }
    return (int)rows.size();
}

// Unit-cost edit distance of two byte strings (what python-Levenshtein's `distance` returns for the ASCII strings
// normalize_string() produces, consensus_utils.py:745-761).  Host helper for the similarity medoid / list alignment.
int32_t kc_levenshtein(const char *a, int32_t alen, const char *b, int32_t blen) {
    if (alen < 0 || blen < 0 || (!a && alen) || (!b && blen)) return -1;
    if (alen < blen) {
        std::swap(a, b);
        std::swap(alen, blen);
    }
    if (blen == 0) return alen;
    thread_local std::vector<int32_t> row;
    row.resize((size_t)blen + 1);
    for (int32_t j = 0; j <= blen; ++j) row[(size_t)j] = j;
    for (int32_t i = 1; i <= alen; ++i) {
        int32_t diag = row[0];
        row[0] = i;
        const char ca = a[i - 1];
        for (int32_t j = 1; j <= blen; ++j) {
            const int32_t up = row[(size_t)j];
            const int32_t v = std::min(std::min(up + 1, row[(size_t)j - 1] + 1), diag + (ca != b[j - 1] ? 1 : 0));
            diag = up;
            row[(size_t)j] = v;
        }
    }
    return row[(size_t)blen];
}

void kc_free_strings(char **arr, int64_t count) {
    if (!arr) return;
    for (int64_t i = 0; i < count; ++i) {
        free(arr[i]);
        arr[i] = nullptr;
    }
}

}  // extern "C"
