/*
 * json_rows.h -- native request-body parser (host only; compiled into libb200forest.so).
 *
 * The reference turns a request into a DataFrame with FastAPI + pydantic + pandas: json.loads of the body, one
 * LoanApplicant object per row, pd.DataFrame(list_of_rows) (reference app/main.py:42-54, app/model.py:8-34) -- O(N x 23)
 * Python objects before any arithmetic.  This parser goes from the request BYTES to 23 columns in one pass:
 * float64 arrays for the numeric features and Arrow string buffers (int32 offsets + UTF-8 bytes) for the categorical
 * ones -- exactly what the row encoder (row_encoder.h) and the drift detector read.
 *
 * It is a FAST PATH with a deliberately narrow grammar, not a validator: it accepts only the regular shape of a
 * request -- an array of objects whose keys are feature names, categorical values plain JSON strings (printable ASCII,
 * no escapes), numeric values plain JSON numbers -- and answers B2F_EIRREGULAR for anything else (unknown or repeated
 * keys, escapes, non-ASCII bytes, null / true / false, strings where numbers are expected, malformed JSON, ...).
 * The caller then hands the same bytes to the general validator (pydantic, server.parse_request), which produces the
 * reference's coercions and 422 responses.  For every body the fast path accepts, the two paths give identical columns
 * (tests/test_server_cpu.py checks that on a corpus and on generated bodies): strings are copied verbatim, numbers go
 * through strtod, which -- like Python's float() -- is correctly rounded.
 */
#pragma once
#include <locale.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/b2f.h"

struct b2f_json_parser {
    int n_cat = 0, n_num = 0;
    std::vector<std::string> names;       /* n_cat categorical names, then n_num numeric names */
    std::vector<std::string> default_str; /* per categorical feature */
    std::vector<double> default_num;      /* per numeric feature */
    /* result of the last parse */
    int64_t n_rows = 0;
    std::vector<std::vector<double>> num;       /* [n_num][n_rows] */
    std::vector<std::vector<int32_t>> str_off;  /* [n_cat][n_rows + 1] */
    std::vector<std::vector<uint8_t>> str_data; /* [n_cat] */
};

extern "C" b2f_json_parser *b2f_json_parser_create(int n_cat, int n_num, const char *names, const int32_t *name_offsets, const char *default_strs,
                                                   const int32_t *default_str_offsets, const double *default_nums) {
    if (n_cat < 0 || n_num < 0 || n_cat + n_num < 1 || n_cat + n_num > 64 || !names || !name_offsets || (n_cat > 0 && (!default_strs || !default_str_offsets)) ||
        (n_num > 0 && !default_nums))
        return nullptr;
    b2f_json_parser *p = new b2f_json_parser();
    p->n_cat = n_cat;
    p->n_num = n_num;
    for (int f = 0; f < n_cat + n_num; ++f) p->names.emplace_back(names + name_offsets[f], (size_t)(name_offsets[f + 1] - name_offsets[f]));
    for (int j = 0; j < n_cat; ++j) p->default_str.emplace_back(default_strs + default_str_offsets[j], (size_t)(default_str_offsets[j + 1] - default_str_offsets[j]));
    p->default_num.assign(default_nums, default_nums + n_num);
    p->num.resize(n_num);
    p->str_off.resize(n_cat);
    p->str_data.resize(n_cat);
    return p;
}

extern "C" void b2f_json_parser_destroy(b2f_json_parser *p) { delete p; }

namespace jsonrows {
static inline const char *skip_ws(const char *s, const char *e) {
    while (s < e && (*s == ' ' || *s == '\n' || *s == '\r' || *s == '\t')) ++s;
    return s;
}
/* a plain string: opening quote at s; printable ASCII without '"' and '\\' inside.  Returns the closing quote or NULL. */
static inline const char *plain_string_end(const char *s, const char *e) {
    for (++s; s < e; ++s) {
        const unsigned char c = (unsigned char)*s;
        if (c == '"') return s;
        if (c < 0x20 || c >= 0x7f || c == '\\') return nullptr;
    }
    return nullptr;
}
/* a JSON number (RFC 8259 grammar, nothing else) at [s, e); returns its end or NULL */
static inline const char *number_end(const char *s, const char *e) {
    if (s < e && *s == '-') ++s;
    if (s >= e) return nullptr;
    if (*s == '0') {
        ++s;
    } else if (*s >= '1' && *s <= '9') {
        while (s < e && *s >= '0' && *s <= '9') ++s;
    } else {
        return nullptr;
    }
    if (s < e && *s == '.') {
        ++s;
        if (s >= e || *s < '0' || *s > '9') return nullptr;
        while (s < e && *s >= '0' && *s <= '9') ++s;
    }
    if (s < e && (*s == 'e' || *s == 'E')) {
        ++s;
        if (s < e && (*s == '+' || *s == '-')) ++s;
        if (s >= e || *s < '0' || *s > '9') return nullptr;
        while (s < e && *s >= '0' && *s <= '9') ++s;
    }
    return s;
}
/* Clinger's fast path: a decimal with at most 15 significant digits and a decimal exponent within +-22 is
 * (integer < 2^53) x or / (an exactly representable power of ten): ONE correctly rounded operation, so the result is
 * the correctly rounded value -- the same number strtod / Python's float() return.  False = not applicable. */
static inline bool exact_decimal(const char *s, const char *e, double *out) {
    static const double p10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    bool neg = false;
    if (*s == '-') {
        neg = true;
        ++s;
    }
    uint64_t m = 0;
    int digits = 0, exp10 = 0;
    bool frac = false;
    for (; s < e; ++s) {
        const char c = *s;
        if (c >= '0' && c <= '9') {
            if (m != 0 || c != '0') ++digits; /* leading zeros carry no significance */
            if (digits > 15) return false;
            m = m * 10 + (uint64_t)(c - '0');
            if (frac) --exp10;
        } else if (c == '.') {
            frac = true;
        } else { /* e / E */
            ++s;
            bool eneg = false;
            if (s < e && (*s == '+' || *s == '-')) eneg = *s++ == '-';
            int x = 0;
            for (; s < e; ++s) {
                x = x * 10 + (*s - '0');
                if (x > 1000) return false;
            }
            exp10 += eneg ? -x : x;
            break;
        }
    }
    if (exp10 < -22 || exp10 > 22) return false;
    double v = (double)m;
    v = exp10 < 0 ? v / p10[-exp10] : v * p10[exp10];
    *out = neg ? -v : v;
    return true;
}
}  // namespace jsonrows

/* Returns the number of rows (>= 0), or B2F_EIRREGULAR when the body is not in the fast path's grammar (the result
 * buffers are then unspecified), or B2F_EINVAL on a NULL argument. */
extern "C" int64_t b2f_json_parser_parse(b2f_json_parser *p, const char *body, int64_t len) {
    using namespace jsonrows;
    if (!p || !body || len < 0) return B2F_EINVAL;
    const int nc = p->n_cat, nf = p->n_cat + p->n_num;
    /* clear() keeps capacity: after one huge body the column vectors would pin their peak size for the life of the
     * service, so buffers far larger than this body can need (a row is >= 2 bytes of body) are released first */
    const size_t keep_rows = (size_t)len / 2 + 1024;
    for (auto &v : p->num) {
        if (v.capacity() > 8 * keep_rows) std::vector<double>().swap(v);
        v.clear();
    }
    for (int j = 0; j < nc; ++j) {
        if (p->str_off[j].capacity() > 8 * keep_rows) std::vector<int32_t>().swap(p->str_off[j]);
        if (p->str_data[j].capacity() > 8 * (size_t)len + 65536) std::vector<uint8_t>().swap(p->str_data[j]);
        p->str_off[j].clear();
        p->str_off[j].push_back(0);
        p->str_data[j].clear();
    }
    p->n_rows = 0;
    const char *s = body, *e = body + len;
    s = skip_ws(s, e);
    if (s >= e || *s != '[') return B2F_EIRREGULAR;
    s = skip_ws(s + 1, e);
    int64_t rows = 0;
    char numbuf[64];
    if (s < e && *s == ']') {
        ++s;
    } else {
        for (;;) {
            if (s >= e || *s != '{') return B2F_EIRREGULAR;
            s = skip_ws(s + 1, e);
            uint64_t seen = 0;
            int expect = 0;
            if (s < e && *s == '}') {
                ++s;
            } else {
                for (;;) {
                    if (s >= e || *s != '"') return B2F_EIRREGULAR;
                    const char *kq = plain_string_end(s, e);
                    if (!kq) return B2F_EIRREGULAR;
                    const size_t klen = (size_t)(kq - s - 1);
                    int f = -1;
                    for (int t = 0; t < nf; ++t) { /* keys usually come in schema order: start at the expected one */
                        const int k = expect + t < nf ? expect + t : expect + t - nf;
                        if (p->names[k].size() == klen && memcmp(p->names[k].data(), s + 1, klen) == 0) {
                            f = k;
                            break;
                        }
                    }
                    if (f < 0 || (seen >> f) & 1) return B2F_EIRREGULAR; /* unknown or repeated key */
                    seen |= 1ull << f;
                    expect = f + 1 < nf ? f + 1 : 0;
                    s = skip_ws(kq + 1, e);
                    if (s >= e || *s != ':') return B2F_EIRREGULAR;
                    s = skip_ws(s + 1, e);
                    if (f < nc) {
                        if (s >= e || *s != '"') return B2F_EIRREGULAR;
                        const char *vq = plain_string_end(s, e);
                        if (!vq) return B2F_EIRREGULAR;
                        p->str_data[f].insert(p->str_data[f].end(), (const uint8_t *)s + 1, (const uint8_t *)vq);
                        if (p->str_data[f].size() > 0x7fffffffu) return B2F_EIRREGULAR;
                        p->str_off[f].push_back((int32_t)p->str_data[f].size());
                        s = vq + 1;
                    } else {
                        const char *ne = number_end(s, e);
                        if (!ne || (size_t)(ne - s) >= sizeof(numbuf)) return B2F_EIRREGULAR;
                        double v;
                        if (!exact_decimal(s, ne, &v)) {
                            memcpy(numbuf, s, (size_t)(ne - s));
                            numbuf[ne - s] = 0;
                            static const locale_t c_locale = newlocale(LC_ALL_MASK, "C", (locale_t)0); /* '.' whatever the process locale */
                            v = strtod_l(numbuf, nullptr, c_locale); /* correctly rounded, as Python's float() */
                        }
                        if (!(v - v == 0.0)) return B2F_EIRREGULAR; /* overflow to infinity: let the general path decide */
                        bool integer_literal = true;
                        for (const char *q = s; q < ne; ++q)
                            if (*q == '.' || *q == 'e' || *q == 'E') integer_literal = false;
                        if (integer_literal) {
                            /* the general path reads these as integers first: "-0" is the integer 0 (-> +0.0), and only
                             * up to 15 digits is the integer -> float64 step trivially exact */
                            if ((ne - s) - (*s == '-') > 15) return B2F_EIRREGULAR;
                            if (v == 0.0) v = 0.0;
                        }
                        p->num[f - nc].push_back(v);
                        s = ne;
                    }
                    s = skip_ws(s, e);
                    if (s < e && *s == ',') {
                        s = skip_ws(s + 1, e);
                        continue;
                    }
                    if (s < e && *s == '}') {
                        ++s;
                        break;
                    }
                    return B2F_EIRREGULAR;
                }
            }
            /* absent fields take the schema defaults (reference app/model.py:12-34) */
            for (int j = 0; j < nc; ++j)
                if (!((seen >> j) & 1)) {
                    const std::string &d = p->default_str[j];
                    p->str_data[j].insert(p->str_data[j].end(), (const uint8_t *)d.data(), (const uint8_t *)d.data() + d.size());
                    p->str_off[j].push_back((int32_t)p->str_data[j].size());
                }
            for (int k = 0; k < p->n_num; ++k)
                if (!((seen >> (nc + k)) & 1)) p->num[k].push_back(p->default_num[k]);
            ++rows;
            s = skip_ws(s, e);
            if (s < e && *s == ',') {
                s = skip_ws(s + 1, e);
                continue;
            }
            if (s < e && *s == ']') {
                ++s;
                break;
            }
            return B2F_EIRREGULAR;
        }
    }
    s = skip_ws(s, e);
    if (s != e) return B2F_EIRREGULAR; /* trailing bytes */
    p->n_rows = rows;
    return rows;
}

extern "C" const double *b2f_json_parser_numeric(const b2f_json_parser *p, int k) {
    return (p && k >= 0 && k < p->n_num) ? p->num[k].data() : nullptr;
}
extern "C" const int32_t *b2f_json_parser_str_offsets(const b2f_json_parser *p, int j) {
    return (p && j >= 0 && j < p->n_cat) ? p->str_off[j].data() : nullptr;
}
extern "C" const uint8_t *b2f_json_parser_str_data(const b2f_json_parser *p, int j, int64_t *nbytes) {
    if (!p || j < 0 || j >= p->n_cat) return nullptr;
    if (nbytes) *nbytes = (int64_t)p->str_data[j].size();
    return p->str_data[j].data();
}
