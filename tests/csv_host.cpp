// Host shim over tuplex_b200/csrc/csvops.cuh: the very functions the CUDA CSV kernels call (span walk, state
// composition, row machine, cell decoders, per-row parse with verification, sequential repair), driven here by plain
// loops in place of the CUDA grid / block scans, so that tests/test_csv_host.py can fuzz them on the CPU against the
// oracle (oracle/csv_oracle.c) and the reference's csvmonkey (oracle/_ref/csv_ref). Test infrastructure only.
#include "../tuplex_b200/csrc/csvops.cuh"
#include <stdlib.h>
#include <string.h>
#include <vector>
using namespace tplx;

struct HostCsv {
    uint32_t n_out = 0, n_rows = 0, n_good = 0, n_bad = 0, sequential = 0;
    uint8_t out_types[TPLX_MAX_COLS];
    std::vector<uint64_t> fixed[TPLX_MAX_COLS];
    std::vector<uint32_t> offsets[TPLX_MAX_COLS];
    std::vector<uint8_t> bytes[TPLX_MAX_COLS];
    std::vector<uint32_t> rowmap;
    std::vector<uint32_t> bad;  // 4 per row: row, code, line_start, line_end
};

extern "C" {
HostCsv *hcsv_parse(const uint8_t *data, uint32_t n, uint8_t delim, uint8_t quote, int skip_header, uint32_t n_file_cols,
                    const uint8_t *col_types, uint32_t n_nulls, const char *const *nulls) {
    HostCsv *H = new HostCsv();
    const uint64_t padded = ((uint64_t)n + 1 + CSV_SPAN - 1) / CSV_SPAN * CSV_SPAN;
    std::vector<uint8_t> buf(padded + 64, 0);
    memcpy(buf.data(), data, n);
    buf[n] = '\n';
    const uint32_t n_spans = (uint32_t)(padded / CSV_SPAN);
    // pass 1 + scan (sequential composition stands in for the block / tile scans)
    std::vector<CsvState> st(n_spans);
    for (uint32_t s = 0; s < n_spans; ++s) st[s] = csv_walk_span(buf.data(), (uint64_t)s * CSV_SPAN, quote, 1u << 31, [](uint64_t) {});
    std::vector<uint32_t> row_end;
    CsvState pre{0, 0, 0};
    for (uint32_t s = 0; s < n_spans; ++s) {
        const uint32_t par0 = pre.par;   // file starts outside quotes
        uint32_t row = pre.c0;
        if (row_end.size() < (size_t)row + st[s].c0 + st[s].c1 + 1) row_end.resize((size_t)row + st[s].c0 + st[s].c1 + 1);
        csv_walk_span(buf.data(), (uint64_t)s * CSV_SPAN, quote, par0, [&](uint64_t pos) { row_end[row++] = (uint32_t)pos; });
        pre = csv_compose(pre, st[s]);
    }
    uint32_t total = pre.c0;
    bool sequential = pre.par != 0;
    CsvParseParams P{};
    std::vector<uint8_t> kind(col_types, col_types + n_file_cols), slot(n_file_cols, 0);
    for (uint32_t c = 0; c < n_file_cols; ++c) {
        if (kind[c] == CSV_SKIP) continue;
        slot[c] = (uint8_t)P.n_out;
        P.out_types[P.n_out] = kind[c];
        P.strk[P.n_out] = kind[c] == TPLX_T_STR ? (int8_t)P.n_str++ : (int8_t)-1;
        ++P.n_out;
    }
    P.nulls.n = (uint8_t)n_nulls;
    uint32_t nvb = 0;
    for (uint32_t k = 0; k < n_nulls; ++k) {
        P.nulls.off[k] = (uint8_t)nvb;
        memcpy(P.nulls.bytes + nvb, nulls[k], strlen(nulls[k]));
        nvb += (uint32_t)strlen(nulls[k]);
    }
    P.nulls.off[n_nulls] = (uint8_t)nvb;
    std::vector<uint64_t> tmp[TPLX_MAX_COLS], lens, good;
    std::vector<uint32_t> code;
    uint32_t flags[4] = {0, 0, 0, 0};
    uint32_t nd = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (sequential) {
            row_end.assign((size_t)n / 2 + 2, 0);
            total = csv_find_rows_sequential(buf.data(), n, delim, quote, row_end.data());
        }
        P.r0 = skip_header ? 1 : 0;
        nd = total > P.r0 ? total - P.r0 : 0;
        P.buf = buf.data();
        P.n = n;
        P.row_end = row_end.data();
        P.nd = nd;
        P.delim = delim;
        P.quote = quote;
        P.n_file_cols = n_file_cols;
        P.col_kind = kind.data();
        P.col_slot = slot.data();
        P.flags = flags;
        for (uint32_t c = 0; c < P.n_out; ++c) {
            tmp[c].assign(nd + 1, 0);
            P.tmp[c] = tmp[c].data();
        }
        lens.assign((size_t)(P.n_str ? P.n_str : 1) * (nd + 1), 0);
        good.assign(nd + 1, 0);
        code.assign(nd + 1, 0);
        P.lens = lens.data();
        P.good = good.data();
        P.code = code.data();
        flags[0] = 0;
        for (uint32_t i = 0; i < nd; ++i) csv_parse_one_row(P, i);
        if (!flags[0]) break;
        if (sequential) abort();
        sequential = true;
    }
    H->sequential = sequential;
    H->n_out = P.n_out;
    H->n_rows = nd;
    memcpy(H->out_types, P.out_types, sizeof H->out_types);
    for (uint32_t c = 0; c < P.n_out; ++c)
        if (P.out_types[c] == TPLX_T_STR) H->offsets[c].push_back(0);
    for (uint32_t i = 0; i < nd; ++i) {
        if (code[i]) {
            const uint32_t r = P.r0 + i;
            H->bad.push_back(i);
            H->bad.push_back(code[i]);
            H->bad.push_back(csv_row_start(buf.data(), row_end.data(), r));
            H->bad.push_back(row_end[r]);
            ++H->n_bad;
            continue;
        }
        H->rowmap.push_back(i);
        ++H->n_good;
        for (uint32_t c = 0; c < P.n_out; ++c) {
            if (P.out_types[c] != TPLX_T_STR) {
                H->fixed[c].push_back(tmp[c][i]);
                continue;
            }
            const uint64_t info = tmp[c][i];
            const uint32_t b = (uint32_t)info, raw = (uint32_t)(info >> 32) & 0x7FFFFFFFu;
            if (!(info >> 63))
                H->bytes[c].insert(H->bytes[c].end(), buf.begin() + b, buf.begin() + b + raw);
            else {
                std::vector<uint8_t> t(raw + 1);
                uint32_t l = csv_dequote(buf.data(), b, b + raw, quote, t.data(), raw + 1);
                H->bytes[c].insert(H->bytes[c].end(), t.begin(), t.begin() + l);
                if (l != lens[(size_t)P.strk[c] * (nd + 1) + i]) abort();
            }
            H->offsets[c].push_back((uint32_t)H->bytes[c].size());
        }
    }
    return H;
}
void hcsv_counts(HostCsv *H, uint32_t *out) { out[0] = H->n_out; out[1] = H->n_rows; out[2] = H->n_good; out[3] = H->n_bad; out[4] = H->sequential; }
uint32_t hcsv_type(HostCsv *H, uint32_t c) { return H->out_types[c]; }
const uint64_t *hcsv_fixed(HostCsv *H, uint32_t c) { return H->fixed[c].data(); }
const uint32_t *hcsv_offsets(HostCsv *H, uint32_t c) { return H->offsets[c].data(); }
const uint8_t *hcsv_bytes(HostCsv *H, uint32_t c, uint64_t *n) { *n = H->bytes[c].size(); return H->bytes[c].data(); }
const uint32_t *hcsv_rowmap(HostCsv *H) { return H->rowmap.data(); }
const uint32_t *hcsv_bad(HostCsv *H) { return H->bad.data(); }
void hcsv_free(HostCsv *H) { delete H; }
// CSV sink (K7): the device row writer over host column arrays
unsigned long long hcsv_write(unsigned n_cols, const uint8_t *types, const uint64_t *const *data, const uint32_t *const *offsets,
                              const uint8_t *const *bytes, unsigned long long n_rows, uint8_t delim, uint8_t quote, uint8_t *out) {
    CsvSinkCols C{};
    C.n_cols = n_cols;
    C.delim = delim;
    C.quote = quote;
    for (unsigned c = 0; c < n_cols; ++c) {
        C.types[c] = types[c];
        C.data[c] = data[c];
        C.offsets[c] = offsets[c];
        C.bytes[c] = bytes[c];
    }
    unsigned long long pos = 0;
    for (unsigned long long r = 0; r < n_rows; ++r) {
        const uint64_t l = csv_sink_row_len(C, r);
        if (out) csv_sink_row_write(C, r, out + pos);
        pos += l;
    }
    return pos;
}
int hcsv_atod(const uint8_t *s, uint32_t len, double *out) { return csv_atod(s, len, out); }
int hcsv_atob(const uint8_t *s, uint32_t len, long long *out) { int64_t v = 0; bool ok = csv_atob(s, len, &v); *out = v; return ok; }
}
