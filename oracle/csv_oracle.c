/* CPU ORACLE for the CSV source (K6) — TEST INFRASTRUCTURE ONLY, never linked into the product library.
 *
 * Sequential restatement, pointer style, of the reference's default CSV reader
 * (paths relative to /root/reference/tuplex/):
 *   csvmonkey::CsvReader::try_parse / read_row      core/include/physical/csvmonkey.h:523-672,727-760
 *   csvmonkey::CsvCell::as_str                      core/include/physical/csvmonkey.h:320-335
 *   VFCSVStreamCursor: '\n' appended at EOF         core/src/physical/CSVReader.cc:66-100,167-177
 *   CSVReader::read (cell count check, exceptions)  core/src/physical/CSVReader.cc:388-634
 *   decodeCells (null values, typed parse)          codegen/src/FlattenedTuple.cc:1215-1330
 *   fast_atoi64 / fast_atod / fast_atob             utils/src/StringUtils.cc:22-63,71-163,180-255
 *   runtime wrappers that trim whitespace           runtime/src/Runtime.cc:319-385
 *
 * Pinned by: the row-parser known-answer tests (test/core/CSVRowParseGeneratorTests.cc:256-980) restated in
 * tests/test_csv_oracle.py, csvmonkey itself compiled from the reference tree into oracle/_ref/csv_ref
 * (cell-for-cell comparison on fuzzed inputs), and the Zillow end-to-end md5 computed from the raw CSV fixture.
 */
#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { T_I64 = 0, T_F64 = 1, T_BOOL = 2, T_STR = 3, T_SKIP = 0xFF };
enum { EC_NULLERROR = 50, EC_BADPARSE_STRING_INPUT = 70 };

typedef struct {
    uint32_t row, code, line_start, line_end;
} csv_bad;

typedef struct {
    uint32_t n_out_cols;
    uint8_t out_types[256];
    uint64_t n_rows, n_normal, n_bad;
    /* per output column */
    int64_t *fixed[256];
    uint32_t *offsets[256];
    uint8_t *bytes[256];
    uint64_t bytes_len[256], bytes_cap[256];
    uint32_t *rowmap;
    csv_bad *bad;
    uint64_t cap_rows, cap_bad;
    /* cell dump (for comparison with csvmonkey) */
    uint8_t *dump;
    uint64_t dump_len, dump_cap;
} csv_result;

static int is_ws(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\x0b' || c == '\x0c'; }

/* utils/src/StringUtils.cc:22-63 */
static int core_atoi64(const char *start, const char *end, int64_t *out) {
    if (start == end) return 1;
    uint64_t x = 0;
    const char *p = start;
    int neg = 0;
    if (*p == '-') {
        neg = 1;
        ++p;
    }
    while (*p >= '0' && *p <= '9') {
        x = x * 10 + (uint64_t)(*p - '0');
        ++p;
    }
    if (p != end) return 1;
    *out = (int64_t)(neg ? (uint64_t)0 - x : x);
    return 0;
}
/* utils/src/StringUtils.cc:71-163 */
static int core_atod(const char *start, const char *end, double *out) {
    if (start == end) return 1;
    int frac;
    double sign, value, scale;
    const char *p = start;
    sign = 1.0;
    if ('-' == *p) {
        sign = -1.0;
        ++p;
    } else if ('+' == *p)
        ++p;
    for (value = 0.0; *p >= '0' && *p <= '9'; p++) value = 10.0 * value + (*p - '0');
    if (*p == '.') {
        double pow10 = 10.0;
        ++p;
        while (*p >= '0' && *p <= '9') {
            value += (*p - '0') / pow10;
            pow10 *= 10.0;
            ++p;
        }
    }
    frac = 0;
    scale = 1.0;
    if (('e' == *p) || ('E' == *p)) {
        unsigned int exponent;
        ++p;
        if ('-' == *p) {
            frac = 1;
            ++p;
        } else if ('+' == *p)
            ++p;
        for (exponent = 0; *p >= '0' && *p <= '9'; p++) exponent = exponent * 10 + (unsigned)(*p - '0');
        if (exponent > 308) exponent = 308;
        while (exponent >= 50) {
            scale *= 1E50;
            exponent -= 50;
        }
        while (exponent >= 8) {
            scale *= 1E8;
            exponent -= 8;
        }
        while (exponent > 0) {
            scale *= 10.0;
            exponent -= 1;
        }
    }
    const char *nanstr = "nan";
    int nanmatch = 0;
    if (p == start)
        while (((*p == nanstr[nanmatch]) || (*p == toupper(nanstr[nanmatch]))) && nanmatch < 3) {
            p++;
            nanmatch++;
        }
    const char *infstr = "infinity";
    int infmatch = 0;
    if (p == start)
        while (((*p == infstr[infmatch]) || (*p == toupper(infstr[infmatch]))) && infmatch < 8) {
            p++;
            infmatch++;
        }
    if (p != end) return 1;
    if (nanmatch == 3)
        *out = NAN;
    else if (infmatch == 3 || infmatch == 8)
        *out = INFINITY;
    else
        *out = sign * (frac ? (value / scale) : (value * scale));
    return 0;
}
/* utils/src/StringUtils.cc:180-255 */
static int core_atob(const char *start, const char *end, int *out) {
    if (start == end) return 1;
    long length = end - start;
    char b[8] = {0};
    if (length > 5) return 1;
    for (long i = 0; i < length; ++i) b[i] = (char)tolower((unsigned char)start[i]);
    switch (length) {
        case 1:
            if (b[0] == 'y' || b[0] == 't') return *out = 1, 0;
            if (b[0] == 'n' || b[0] == 'f') return *out = 0, 0;
            return 1;
        case 2: return strcmp(b, "no") == 0 ? (*out = 0, 0) : 1;
        case 3: return strcmp(b, "yes") == 0 ? (*out = 1, 0) : 1;
        case 4: return strcmp(b, "true") == 0 ? (*out = 1, 0) : 1;
        case 5: return strcmp(b, "false") == 0 ? (*out = 0, 0) : 1;
    }
    return 1;
}
/* runtime/src/Runtime.cc:319-365: trim, then convert */
static void rt_trim(const char **ps, const char **pe) {
    const char *start = *ps, *end = *pe;
    while (start < end && is_ws(*start)) start++;
    end--;
    while (end > start && is_ws(*end)) end--;
    end++;
    *ps = start;
    *pe = end;
}
int csv_oracle_atoi64(const char *s, int64_t *out) {
    const char *a = s, *e = s + strlen(s);
    rt_trim(&a, &e);
    return core_atoi64(a, e, out);
}
int csv_oracle_atod(const char *s, double *out) {
    const char *a = s, *e = s + strlen(s);
    rt_trim(&a, &e);
    return core_atod(a, e, out);
}
int csv_oracle_atob(const char *s, int *out) { return core_atob(s, s + strlen(s), out); }

static void dump_put(csv_result *r, const void *p, uint64_t n) {
    if (r->dump_len + n > r->dump_cap) {
        r->dump_cap = (r->dump_len + n) * 2 + 64;
        r->dump = (uint8_t *)realloc(r->dump, r->dump_cap);
    }
    memcpy(r->dump + r->dump_len, p, n);
    r->dump_len += n;
}
static void bytes_put(csv_result *r, uint32_t c, const char *p, uint64_t n) {
    if (r->bytes_len[c] + n > r->bytes_cap[c]) {
        r->bytes_cap[c] = (r->bytes_len[c] + n) * 2 + 64;
        r->bytes[c] = (uint8_t *)realloc(r->bytes[c], r->bytes_cap[c]);
    }
    memcpy(r->bytes[c] + r->bytes_len[c], p, n);
    r->bytes_len[c] += n;
}

typedef struct {
    const char *ptr;
    size_t size;
    int escaped;
} cell_t;

/* as_str (csvmonkey.h:320-335) into a NUL-terminated scratch string; returns length */
static size_t cell_as_str(const cell_t *c, char quote, char *out) {
    if (!c->escaped) {
        memcpy(out, c->ptr, c->size);
        out[c->size] = 0;
        return c->size;
    }
    size_t o = 0;
    for (size_t i = 0; i < c->size;) {
        if (c->ptr[i] == quote) i++;
        if (i >= c->size) break;
        out[o++] = c->ptr[i++];
    }
    out[o] = 0;
    return o;
}

/* Parses `n` bytes. col_types[n_file_cols]: T_* or T_SKIP. dump_cells != 0 additionally records every row's
 * dequoted cells as [u32 count]([u32 len] bytes)* for comparison with csvmonkey. */
csv_result *csv_oracle_parse(const uint8_t *bytes, uint64_t n, char delim, char quote, int skip_header, uint32_t n_file_cols,
                             const uint8_t *col_types, uint32_t n_nulls, const char *const *nulls, int dump_cells) {
    csv_result *r = (csv_result *)calloc(1, sizeof(csv_result));
    char *buf = (char *)malloc(n + 2);
    memcpy(buf, bytes, n);
    buf[n] = '\n'; /* CSVReader.cc:94-100 */
    buf[n + 1] = 0;
    const char *endp = buf + n + 1;
    for (uint32_t c = 0; c < n_file_cols; ++c)
        if (col_types[c] != T_SKIP) r->out_types[r->n_out_cols++] = col_types[c];
    size_t cells_cap = 64;
    cell_t *cells = (cell_t *)malloc(cells_cap * sizeof(cell_t));
    char *scratch = (char *)malloc(n + 2);
    const char *p = buf;
    uint64_t rowno = 0;
    int first = 1;
    for (;;) {
        /* ---- try_parse ---- */
        size_t count = 0;
        while (p < endp && (*p == '\r' || *p == '\n')) ++p; /* newline_skip */
        if (p >= endp) break;
        const char *line_start = p;
        int underrun = 0;
        for (;;) { /* cell_start */
            if (count == cells_cap) {
                cells_cap *= 2;
                cells = (cell_t *)realloc(cells, cells_cap * sizeof(cell_t));
            }
            cell_t *cell = &cells[count];
            cell->escaped = 0;
            if (p >= endp) {
                underrun = 1;
                break;
            }
            if (*p == '\r' || *p == '\n') {
                cell->ptr = p;
                cell->size = 0;
                ++count;
                break;
            }
            if (*p == quote) {
                const char *cs = ++p;
                int done = 0, rowdone = 0;
                while (!done) {
                    while (p < endp && *p != quote) ++p;
                    if (p >= endp) {
                        underrun = 1;
                        break;
                    }
                    ++p; /* past the quote */
                    if (p >= endp) {
                        underrun = 1;
                        break;
                    }
                    if (*p == delim) {
                        cell->ptr = cs;
                        cell->size = (size_t)(p - cs - 1);
                        ++count;
                        ++p;
                        done = 1;
                    } else if (*p == '\r' || *p == '\n') {
                        cell->ptr = cs;
                        cell->size = (size_t)(p - cs - 1);
                        ++count;
                        done = rowdone = 1;
                    } else {
                        cell->escaped = 1;
                        ++p;
                    }
                }
                if (underrun || rowdone) break;
            } else {
                const char *cs = p;
                while (*p != delim && *p != '\r' && *p != '\n') ++p; /* buf[n] == '\n' stops this */
                cell->ptr = cs;
                cell->size = (size_t)(p - cs);
                ++count;
                if (*p == delim)
                    ++p;
                else
                    break;
            }
        }
        if (underrun) break; /* read_row(): no more rows (yield_incomplete_row = false) */
        const char *line_end = p; /* the terminating newline */
        ++p;
        if (first && skip_header) {
            first = 0;
            continue;
        }
        first = 0;
        /* ---- CSVReader::read row handling ---- */
        if (dump_cells) {
            uint32_t cnt = (uint32_t)count;
            dump_put(r, &cnt, 4);
            for (size_t i = 0; i < count; ++i) {
                uint32_t len = (uint32_t)cell_as_str(&cells[i], quote, scratch);
                dump_put(r, &len, 4);
                dump_put(r, scratch, len);
            }
        }
        if (rowno + 1 > r->cap_rows) {
            uint64_t cap = r->cap_rows ? r->cap_rows * 2 : 1024;
            for (uint32_t c = 0; c < r->n_out_cols; ++c) {
                if (r->out_types[c] == T_STR)
                    r->offsets[c] = (uint32_t *)realloc(r->offsets[c], (cap + 1) * 4);
                else
                    r->fixed[c] = (int64_t *)realloc(r->fixed[c], cap * 8);
            }
            r->rowmap = (uint32_t *)realloc(r->rowmap, cap * 4);
            r->cap_rows = cap;
        }
        uint32_t code = 0;
        uint64_t slot_save[256];
        for (uint32_t c = 0; c < r->n_out_cols; ++c) slot_save[c] = r->bytes_len[c];
        if (count != n_file_cols)
            code = EC_BADPARSE_STRING_INPUT; /* CSVReader.cc:470-479 */
        else {
            uint32_t oc = 0;
            for (uint32_t c = 0; c < n_file_cols && !code; ++c) {
                if (col_types[c] == T_SKIP) continue;
                size_t len = cell_as_str(&cells[c], quote, scratch);
                int isnull = 0;
                for (uint32_t k = 0; k < n_nulls; ++k)
                    if (strcmp(scratch, nulls[k]) == 0) isnull = 1; /* compareToNullValues on the 0-terminated cell */
                if (isnull) {
                    code = EC_NULLERROR; /* non-Option column: FlattenedTuple.cc:1266-1279 */
                    break;
                }
                const char *a = scratch, *e = scratch + len;
                switch (col_types[c]) {
                    case T_I64: {
                        int64_t v;
                        rt_trim(&a, &e);
                        if (core_atoi64(a, e, &v))
                            code = EC_BADPARSE_STRING_INPUT;
                        else
                            r->fixed[oc][r->n_normal] = v;
                    } break;
                    case T_F64: {
                        double d;
                        rt_trim(&a, &e);
                        if (core_atod(a, e, &d))
                            code = EC_BADPARSE_STRING_INPUT;
                        else
                            memcpy(&r->fixed[oc][r->n_normal], &d, 8);
                    } break;
                    case T_BOOL: {
                        int bv;
                        if (core_atob(a, e, &bv))
                            code = EC_BADPARSE_STRING_INPUT;
                        else
                            r->fixed[oc][r->n_normal] = bv;
                    } break;
                    default:
                        r->offsets[oc][r->n_normal] = (uint32_t)r->bytes_len[oc];
                        bytes_put(r, oc, scratch, len);
                }
                ++oc;
            }
        }
        if (code) {
            for (uint32_t c = 0; c < r->n_out_cols; ++c) r->bytes_len[c] = slot_save[c];
            if (r->n_bad + 1 > r->cap_bad) {
                r->cap_bad = r->cap_bad ? r->cap_bad * 2 : 64;
                r->bad = (csv_bad *)realloc(r->bad, r->cap_bad * sizeof(csv_bad));
            }
            csv_bad b = {(uint32_t)rowno, code, (uint32_t)(line_start - buf), (uint32_t)(line_end - buf)};
            r->bad[r->n_bad++] = b;
        } else {
            r->rowmap[r->n_normal] = (uint32_t)rowno;
            r->n_normal++;
        }
        rowno++;
    }
    r->n_rows = rowno;
    for (uint32_t c = 0; c < r->n_out_cols; ++c)
        if (r->out_types[c] == T_STR) {
            if (!r->offsets[c]) r->offsets[c] = (uint32_t *)malloc(4);
            r->offsets[c][r->n_normal] = (uint32_t)r->bytes_len[c];
        }
    free(cells);
    free(scratch);
    free(buf);
    return r;
}

void csv_oracle_free(csv_result *r) {
    if (!r) return;
    for (uint32_t c = 0; c < 256; ++c) {
        free(r->fixed[c]);
        free(r->offsets[c]);
        free(r->bytes[c]);
    }
    free(r->rowmap);
    free(r->bad);
    free(r->dump);
    free(r);
}

/* ---- CSV sink oracle: fast_csvwriter (core/src/physical/PipelineBuilder.cc:1550-1722) + quoteForCSV
 * (runtime/src/Runtime.cc:682-738), one row at a time with snprintf. Known answers: test/runtime/RuntimeTest.cc:207-213. */
#include <stdio.h>
typedef struct {
    uint8_t type;
    uint8_t pad[7];
    const void *data;
    const uint32_t *offsets;
    uint64_t data_bytes;
    const uint32_t *valid; /* same layout as tplx_ocol (tplx_oracle.c); the CSV sink writes non-Option columns only */
} sink_col;

/* returns bytes written; out may be NULL to size */
uint64_t csv_oracle_write(const sink_col *cols, uint32_t n_cols, uint64_t n_rows, char delim, char quote, uint8_t *out) {
    uint64_t pos = 0;
    char tmp[32];
    for (uint64_t r = 0; r < n_rows; ++r) {
        for (uint32_t c = 0; c < n_cols; ++c) {
            if (cols[c].type == T_STR) {
                const char *s = (const char *)cols[c].data + cols[c].offsets[r];
                uint32_t len = cols[c].offsets[r + 1] - cols[c].offsets[r];
                size_t num_quotes = 0;
                int need_to_quote = 0;
                for (uint32_t i = 0; i < len; ++i) {
                    if (s[i] == quote) num_quotes++;
                    if (s[i] == delim || s[i] == '\n' || s[i] == '\r') need_to_quote = 1;
                }
                if (num_quotes > 0 || need_to_quote) {
                    if (out) out[pos] = (uint8_t)quote;
                    pos++;
                    for (uint32_t i = 0; i < len; ++i) {
                        if (s[i] == quote) {
                            if (out) out[pos] = (uint8_t)quote;
                            pos++;
                        }
                        if (out) out[pos] = (uint8_t)s[i];
                        pos++;
                    }
                    if (out) out[pos] = (uint8_t)quote;
                    pos++;
                } else {
                    if (out) memcpy(out + pos, s, len);
                    pos += len;
                }
            } else if (cols[c].type == T_F64) {
                /* ryu d2fixed_buffered_n(d, 8, buf): printf("%.8f") digits (glibc is exact), specials "nan" / "[-]Infinity" */
                double d = ((const double *)cols[c].data)[r];
                char fb[400];
                int k = isnan(d) ? snprintf(fb, sizeof fb, "nan") : isinf(d) ? snprintf(fb, sizeof fb, d < 0 ? "-Infinity" : "Infinity")
                                                                           : snprintf(fb, sizeof fb, "%.8f", d);
                if (out) memcpy(out + pos, fb, (size_t)k);
                pos += (uint64_t)k;
            } else {
                int64_t v = ((const int64_t *)cols[c].data)[r];
                int k = cols[c].type == T_BOOL ? snprintf(tmp, sizeof tmp, "%s", v ? "true" : "false") : snprintf(tmp, sizeof tmp, "%lld", (long long)v);
                if (out) memcpy(out + pos, tmp, (size_t)k);
                pos += (uint64_t)k;
            }
            if (out) out[pos] = (uint8_t)(c + 1 == n_cols ? '\n' : delim);
            pos++;
        }
    }
    return pos;
}
