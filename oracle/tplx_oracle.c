/*
 * tplx_oracle.c — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, row-at-a-time restatement of the reference's normal-case row pipeline
 * (paths relative to /root/reference/tuplex/). It must never be linked, imported or called by the
 * product path (tuplex_b200/): only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it.
 *
 * What is restated, and from where:
 *   block loop + row numbering     core/src/physical/TuplexSourceTaskBuilder.cc:104-215,
 *                                  core/src/physical/TransformTask.cc:763-889
 *   per-row pipeline / filter      core/src/physical/PipelineBuilder.cc:565-1025 (filter :615-700)
 *   i64/f64 arithmetic             codegen/src/BlockGeneratorVisitor.cc:152-836,
 *                                  codegen/src/LLVMEnvironment.cc:1377-1430
 *   string builtins                codegen/src/FunctionRegistry.cc:83-148,2119-2334,
 *                                  runtime/src/Runtime.cc:319-341,387-540, runtime/src/StringFunctions.cc:71-110,
 *                                  utils/src/StringUtils.cc:22-63, codegen/src/BlockGeneratorVisitor.cc:675-775,
 *                                  :3869-3903, :4469-4690
 *   aggregates                     core/src/physical/AggregateFunctions.cc, TransformTask.cc:208-375
 *   row / partition / exception    utils/src/Serializer.cc:1016-1117, core/include/Partition.h:130-139,
 *   byte formats                   core/include/physical/TransformTask.h:47-92, IExceptionableTask.h:22-36
 *
 * Strings are handled the way the reference handles them: NUL-terminated heap buffers, libc strstr /
 * snprintf / tolower — deliberately NOT the (pointer,length,flag) views the CUDA VM uses, so that the two
 * implementations are independent.
 *
 * Parity pinned: yes — see oracle/README.md (Zillow md5 via the reference's own zillow.cpp built into
 * oracle/_ref, TPC-H Q6 golden of test/core/TPCH.cc:85-97, README and python/tests goldens).
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/tplx_ir.h"

/* ------------------------------------------------------------------------------------------- */
/* per-row arena (the reference's rtmalloc / rtfree_all, runtime/src/Runtime.cc:186-303)        */
/* ------------------------------------------------------------------------------------------- */
typedef struct arena {
    char *buf;
    size_t used, cap;
} arena;

static char *rtmalloc(arena *a, size_t n) {
    if (a->used + n > a->cap) {
        /* keep old blocks alive until rtfree_all: allocate a fresh bigger block, leak the old one into a chain */
        size_t ncap = (a->cap ? a->cap * 2 : 4096);
        while (ncap < n + 16) ncap *= 2;
        char *nb = (char *)malloc(ncap + sizeof(char *));
        *(char **)nb = a->buf ? a->buf - sizeof(char *) : NULL;
        a->buf = nb + sizeof(char *);
        a->cap = ncap;
        a->used = 0;
    }
    char *p = a->buf + a->used;
    a->used += (n + 7) & ~(size_t)7;
    return p;
}
static void rtfree_all(arena *a) {
    /* free chained older blocks, keep the newest */
    if (!a->buf) return;
    char *blk = *(char **)(a->buf - sizeof(char *));
    while (blk) {
        char *next = *(char **)blk;
        free(blk);
        blk = next;
    }
    *(char **)(a->buf - sizeof(char *)) = NULL;
    a->used = 0;
}
static void arena_destroy(arena *a) {
    rtfree_all(a);
    if (a->buf) free(a->buf - sizeof(char *));
    a->buf = NULL;
    a->cap = a->used = 0;
}

/* ------------------------------------------------------------------------------------------- */
/* builtins                                                                                     */
/* ------------------------------------------------------------------------------------------- */
/* LLVMEnvironment::floorDivision / floorModulo (codegen/src/LLVMEnvironment.cc:1377-1430) */
int64_t tplx_o_floordiv(int64_t x, int64_t y) {
    int64_t q = x / y, r = x % y;
    if ((r != 0) && ((r < 0) != (y < 0))) --q;
    return q;
}
int64_t tplx_o_floormod(int64_t x, int64_t y) {
    int64_t r = x % y;
    if ((r != 0) && ((r < 0) != (y < 0))) r += y;
    return r;
}
double tplx_o_fmod(double x, double y) {
    double r = fmod(x, y);
    if (r != 0.0 && ((r < 0.0) != (y < 0.0))) r = r + y;
    return r;
}

/* fast_atoi64: inner parser utils/src/StringUtils.cc:22-63, whitespace wrapper runtime/src/Runtime.cc:319-341.
 * returns 0 on success, TPLX_EC_VALUEERROR otherwise. `s` is NUL-terminated with length len. */
int32_t tplx_o_atoi64(const char *s, int64_t len, int64_t *out) {
    const char *start = s, *end = s + len;
    while (start < end && (*start == ' ' || *start == '\t' || *start == '\n' || *start == '\r' || *start == '\x0b' || *start == '\x0c'))
        start++;
    end--;
    while (end > start && (*end == ' ' || *end == '\t' || *end == '\n' || *end == '\r' || *end == '\x0b' || *end == '\x0c')) end--;
    end++;
    if (start == end) return TPLX_EC_VALUEERROR; /* NULLERROR mapped to ValueError */
    uint64_t x = 0; /* unsigned: the reference's signed overflow wraps in practice */
    const char *p = start;
    int neg = 0;
    if (*p == '-') {
        neg = 1;
        ++p;
    }
    while (*p >= '0' && *p <= '9') {
        x = (x * 10) + (uint64_t)(*p - '0');
        ++p;
    }
    if (p != end) return TPLX_EC_VALUEERROR;
    *out = neg ? (int64_t)(0 - x) : (int64_t)x;
    return 0;
}

/* strRfind (runtime/src/Runtime.cc:387-397): std::string::rfind */
int64_t tplx_o_rfind(const char *s, const char *needle) {
    size_t n = strlen(s), m = strlen(needle);
    if (m > n) return -1;
    for (int64_t i = (int64_t)(n - m); i >= 0; --i)
        if (memcmp(s + i, needle, m) == 0) return i;
    return -1;
}

/* strReplace (runtime/src/Runtime.cc:401-540); result size excludes the NUL here */
static char *o_replace(arena *a, const char *str, const char *from, const char *to, size_t *res_len) {
    if (str[0] == '\0') {
        *res_len = 0;
        return (char *)str;
    }
    if (from[0] == '\0' && to[0] == '\0') {
        *res_len = strlen(str);
        return (char *)str;
    }
    size_t len = strlen(str), tolen = strlen(to), fromlen = strlen(from);
    if (from[0] == '\0') {
        size_t retlen = (tolen + 1) * (1 + len);
        char *res = rtmalloc(a, retlen + 1);
        size_t pos = 0;
        for (size_t i = 0; i < len; ++i) {
            for (size_t j = 0; j < tolen; ++j) res[pos++] = to[j];
            res[pos++] = str[i];
        }
        for (size_t j = 0; j < tolen; ++j) res[pos++] = to[j];
        res[pos] = '\0';
        *res_len = strlen(res);
        return res;
    }
    size_t count = 0;
    const char *p = str, *q;
    while ((q = strstr(p, from)) != NULL) {
        count++;
        p = q + fromlen;
    }
    size_t retlen = len + count * tolen - count * fromlen;
    char *ret = rtmalloc(a, retlen + 1);
    char *w = ret;
    p = str;
    while ((q = strstr(p, from)) != NULL) {
        memcpy(w, p, (size_t)(q - p));
        w += q - p;
        memcpy(w, to, tolen);
        w += tolen;
        p = q + fromlen;
    }
    strcpy(w, p);
    *res_len = retlen;
    return ret;
}

/* processSliceIndex, positive stride (codegen/src/BlockGeneratorVisitor.cc:4618-4690) */
static int64_t o_slice_index(int64_t index, int64_t len) {
    if (index < -len) return 0;
    if (index <= -1) return index + len;
    if (index < len) return index;
    return len;
}

/* ------------------------------------------------------------------------------------------- */
/* column blocks and results                                                                    */
/* ------------------------------------------------------------------------------------------- */
typedef struct tplx_ocol {
    uint8_t type;
    uint8_t pad[7];
    const void *data;        /* fixed: 8-byte values; str: bytes */
    const uint32_t *offsets; /* str: n+1 */
    uint64_t data_bytes;
    const uint32_t *valid;   /* Option[T] column: bit (r & 31) of word r >> 5 set = value present; NULL = not an Option column */
} tplx_ocol;

typedef struct tplx_oexc {
    int64_t row, row_no, code, op_id;
} tplx_oexc;

typedef struct tplx_oresult {
    uint64_t n_out, n_exc, n_accs, n_cols;
    uint8_t col_types[TPLX_MAX_COLS];
    int64_t *fixed[TPLX_MAX_COLS];     /* fixed width out columns */
    uint32_t *offsets[TPLX_MAX_COLS];  /* str out columns */
    char *bytes[TPLX_MAX_COLS];
    uint64_t bytes_len[TPLX_MAX_COLS], bytes_cap[TPLX_MAX_COLS];
    tplx_oexc *exc;
    uint64_t exc_cap, out_cap;
    int64_t acc_seq[TPLX_MAX_ACCS];  /* sequential fold in row order = reference order within a task */
    int64_t acc_tree[TPLX_MAX_ACCS]; /* same fixed reduction tree as the CUDA kernels */
} tplx_oresult;

typedef struct oval {
    int64_t i;      /* i64 / bool / f64 bits */
    const char *s;  /* NUL-terminated */
    int64_t len;    /* strlen(s) */
} oval;

typedef struct ostage {
    tplx_stage_header h;
    const uint8_t *in_types;
    const tplx_outcol *out_cols;
    const tplx_acc *accs;
    const int64_t *opids;
    const tplx_instr *instrs;
    const uint8_t *cpool;
} ostage;

static size_t pad8(size_t v) { return (v + 7) & ~(size_t)7; }

static int parse_stage(const void *desc, uint64_t bytes, ostage *s) {
    if (bytes < sizeof(tplx_stage_header)) return -1;
    const uint8_t *p = (const uint8_t *)desc;
    memcpy(&s->h, p, sizeof(s->h));
    if (s->h.magic != TPLX_IR_MAGIC || s->h.version != TPLX_IR_VERSION || s->h.total_bytes != bytes) return -1;
    size_t off = sizeof(s->h);
    s->in_types = p + off;
    off += pad8(s->h.n_in_cols);
    s->out_cols = (const tplx_outcol *)(p + off);
    off += pad8(s->h.n_out_cols * sizeof(tplx_outcol));
    s->accs = (const tplx_acc *)(p + off);
    off += s->h.n_accs * sizeof(tplx_acc);
    s->opids = (const int64_t *)(p + off);
    off += s->h.n_ops * sizeof(int64_t);
    s->instrs = (const tplx_instr *)(p + off);
    off += (size_t)s->h.n_instr * sizeof(tplx_instr);
    s->cpool = p + off;
    /* a nested prefilter stage is an execution hint only: the oracle runs the full stage over every row */
    return off + pad8(s->h.const_bytes) + s->h.prefilter_bytes + s->h.fused_bytes == bytes ? 0 : -1;
}

static double as_f(int64_t bits) {
    double d;
    memcpy(&d, &bits, 8);
    return d;
}
static int64_t as_i(double d) {
    int64_t b;
    memcpy(&b, &d, 8);
    return b;
}

static char *dup_n(arena *a, const char *p, size_t n) {
    char *r = rtmalloc(a, n + 1);
    memcpy(r, p, n);
    r[n] = 0;
    return r;
}

/* Evaluate one row. regs has n_slots entries (string values live in slot s; slot s+1 unused).
 * Returns 0 = row kept, 1 = filtered, 2 = exception (ec/opidx set). */
static int eval_row(const ostage *S, const tplx_ocol *cols, uint64_t row, oval *regs, arena *A, int64_t *ec, uint32_t *opidx) {
    const uint32_t n = S->h.n_instr;
    for (uint32_t pc = 0; pc < n; ++pc) {
        const tplx_instr *in = &S->instrs[pc];
        if (in->guard != TPLX_NOSLOT && regs[in->guard].i == 0) continue;
        oval *d = in->dst != TPLX_NOSLOT ? &regs[in->dst] : NULL;
        const oval *a = in->a != TPLX_NOSLOT ? &regs[in->a] : NULL;
        const oval *b = in->b != TPLX_NOSLOT ? &regs[in->b] : NULL;
        const oval *c = in->c != TPLX_NOSLOT ? &regs[in->c] : NULL;
        /* constant operands (TPLX_F_*_CONST): scalars are the immediate, strings are constant-pool views
         * (offset | length << 32) that the reference would hold as NUL-terminated global strings */
        oval ka, kb, kc;
        if (in->flags & (TPLX_F_A_CONST | TPLX_F_B_CONST | TPLX_F_C_CONST)) {
            const int op = in->op;
            const int strsel = (op == TPLX_OP_SEL || op == TPLX_OP_MOV) && (in->flags & 3) == 2;
            const int a_str = strsel || (op >= TPLX_OP_SLEN && op <= TPLX_OP_SSTRIP && op != TPLX_OP_SFMTD && op != TPLX_OP_I2S) || op == TPLX_OP_SFINDE;
            const int b_str = strsel || op == TPLX_OP_SFIND || op == TPLX_OP_SRFIND || op == TPLX_OP_SIN || op == TPLX_OP_SEQ || op == TPLX_OP_SFINDE || op == TPLX_OP_SRFINDK ||
                              op == TPLX_OP_SSTARTS || op == TPLX_OP_SENDS || op == TPLX_OP_SCONCAT || op == TPLX_OP_SREPLACE;
            const int c_str = op == TPLX_OP_SREPLACE;
#define KONST(dst, enc, is_str) do { if (is_str) { (dst).len = (int64_t)((uint64_t)(enc) >> 32); \
                (dst).s = dup_n(A, (const char *)S->cpool + ((uint64_t)(enc) & 0xFFFFFFFFull), (size_t)(dst).len); (dst).i = 0; } \
            else { (dst).i = (enc); (dst).s = NULL; (dst).len = 0; } } while (0)
            if (in->flags & TPLX_F_A_CONST) { KONST(ka, in->imm2, a_str); a = &ka; }
            if (in->flags & TPLX_F_B_CONST) { KONST(kb, in->imm, b_str); b = &kb; }
            if (in->flags & TPLX_F_C_CONST) { KONST(kc, in->imm2, c_str); c = &kc; }
#undef KONST
        }
#define RAISE(code) do { *ec = (code); *opidx = in->opidx; return 2; } while (0)
        switch (in->op) {
            case TPLX_OP_NOP: break;
            case TPLX_OP_LDCOL: {
                const tplx_ocol *ci = &cols[in->imm];
                if (in->flags == TPLX_T_STR) {
                    uint32_t o0 = ci->offsets[row], o1 = ci->offsets[row + 1];
                    /* the reference deserialises a NUL-terminated copy out of the row buffer */
                    d->s = dup_n(A, (const char *)ci->data + o0, o1 - o0);
                    d->len = o1 - o0;
                } else d->i = ((const int64_t *)ci->data)[row];
                break;
            }
            case TPLX_OP_LDI: d->i = in->imm; break;
            case TPLX_OP_LDROW: d->i = (int64_t)row; break;
            case TPLX_OP_LDS: d->len = (int64_t)((uint64_t)in->imm >> 32); d->s = dup_n(A, (const char *)S->cpool + ((uint64_t)in->imm & 0xFFFFFFFFull), (size_t)d->len); break;
            case TPLX_OP_MOV: *d = *a; break;
            case TPLX_OP_SEL: { oval v = c->i ? *a : *b; *d = v; break; }
            case TPLX_OP_IADD: d->i = (int64_t)((uint64_t)a->i + (uint64_t)b->i); break;
            case TPLX_OP_ISUB: d->i = (int64_t)((uint64_t)a->i - (uint64_t)b->i); break;
            case TPLX_OP_IMUL: d->i = (int64_t)((uint64_t)a->i * (uint64_t)b->i); break;
            case TPLX_OP_IFLOORDIV: if (b->i == 0) RAISE(TPLX_EC_ZERODIVISIONERROR); d->i = tplx_o_floordiv(a->i, b->i); break;
            case TPLX_OP_IMOD: if (b->i == 0) RAISE(TPLX_EC_ZERODIVISIONERROR); d->i = tplx_o_floormod(a->i, b->i); break;
            case TPLX_OP_INEG: d->i = (int64_t)(0 - (uint64_t)a->i); break;
            case TPLX_OP_IAND: d->i = a->i & b->i; break;
            case TPLX_OP_IOR: d->i = a->i | b->i; break;
            case TPLX_OP_IXOR: d->i = a->i ^ b->i; break;
            case TPLX_OP_ISHL: d->i = (int64_t)((uint64_t)a->i << (b->i & 63)); break;
            case TPLX_OP_ISHR: d->i = a->i >> (b->i & 63); break;
            case TPLX_OP_IABS: d->i = a->i < 0 ? (int64_t)(0 - (uint64_t)a->i) : a->i; break;
            case TPLX_OP_FADD: d->i = as_i(as_f(a->i) + as_f(b->i)); break;
            case TPLX_OP_FSUB: d->i = as_i(as_f(a->i) - as_f(b->i)); break;
            case TPLX_OP_FMUL: d->i = as_i(as_f(a->i) * as_f(b->i)); break;
            case TPLX_OP_FDIV: if (as_f(b->i) == 0.0) RAISE(TPLX_EC_ZERODIVISIONERROR); d->i = as_i(as_f(a->i) / as_f(b->i)); break;
            case TPLX_OP_FMOD: if (as_f(b->i) == 0.0) RAISE(TPLX_EC_ZERODIVISIONERROR); d->i = as_i(tplx_o_fmod(as_f(a->i), as_f(b->i))); break;
            case TPLX_OP_FNEG: d->i = as_i(-as_f(a->i)); break;
            case TPLX_OP_FABS: d->i = as_i(fabs(as_f(a->i))); break;
            case TPLX_OP_FFLOORDIV: {
                if (as_f(b->i) == 0.0) RAISE(TPLX_EC_ZERODIVISIONERROR);
                int64_t xi = (int64_t)as_f(a->i), yi = (int64_t)as_f(b->i);
                if (yi == 0) RAISE(TPLX_EC_ZERODIVISIONERROR);
                d->i = as_i((double)tplx_o_floordiv(xi, yi));
                break;
            }
            case TPLX_OP_I2F: d->i = as_i((double)a->i); break;
            case TPLX_OP_F2I: d->i = (int64_t)as_f(a->i); break;
            case TPLX_OP_ICMP: {
                int64_t x = a->i, y = b->i;
                int r = 0;
                switch (in->flags & 7) {
                    case TPLX_CMP_EQ: r = x == y; break;
                    case TPLX_CMP_NE: r = x != y; break;
                    case TPLX_CMP_LT: r = x < y; break;
                    case TPLX_CMP_LE: r = x <= y; break;
                    case TPLX_CMP_GT: r = x > y; break;
                    default: r = x >= y; break;
                }
                d->i = r;
                break;
            }
            case TPLX_OP_FCMP: {
                double x = as_f(a->i), y = as_f(b->i);
                int r = 0;
                switch (in->flags & 7) {
                    case TPLX_CMP_EQ: r = x == y; break;
                    case TPLX_CMP_NE: r = islessgreater(x, y); break; /* FCMP_ONE */
                    case TPLX_CMP_LT: r = x < y; break;
                    case TPLX_CMP_LE: r = x <= y; break;
                    case TPLX_CMP_GT: r = x > y; break;
                    default: r = x >= y; break;
                }
                d->i = r;
                break;
            }
            case TPLX_OP_BAND: d->i = (a->i != 0) && (b->i != 0); break;
            case TPLX_OP_BOR: d->i = (a->i != 0) || (b->i != 0); break;
            case TPLX_OP_BNOT: d->i = a->i == 0; break;
            case TPLX_OP_SLEN: d->i = a->len; break;
            case TPLX_OP_SFIND: { const char *r = strstr(a->s, b->s); d->i = r ? (int64_t)(r - a->s) : -1; break; }
            case TPLX_OP_SRFIND: d->i = tplx_o_rfind(a->s, b->s); break;
            /* fused idioms, evaluated as the primitive sequences they stand for (include/tplx_ir.h) */
            case TPLX_OP_SFINDE: { const char *r = strstr(a->s, b->s); d->i = r ? (int64_t)(r - a->s) : a->len; break; }
            case TPLX_OP_SRFINDK: { int64_t r = tplx_o_rfind(a->s, b->s); d->i = r < 0 ? 0 : (int64_t)((uint64_t)r + (uint64_t)in->imm2); break; }
            case TPLX_OP_SIN: d->i = strstr(b->s, a->s) != NULL; break;
            case TPLX_OP_SEQ: d->i = (strcmp(a->s, b->s) == 0) != (in->flags & 1); break;
            case TPLX_OP_STRUTH: d->i = a->len > 0; break;
            case TPLX_OP_SSTARTS: d->i = b->len <= a->len && memcmp(a->s, b->s, (size_t)b->len) == 0; break;
            case TPLX_OP_SENDS: d->i = b->len <= a->len && memcmp(a->s + a->len - b->len, b->s, (size_t)b->len) == 0; break;
            case TPLX_OP_SSLICE: {
                int64_t len = a->len;
                int64_t st = (in->flags & TPLX_SL_HAS_START) ? o_slice_index(b->i, len) : 0;
                int64_t en = (in->flags & TPLX_SL_HAS_END) ? o_slice_index(c->i, len) : len;
                if (st < en) { d->s = dup_n(A, a->s + st, (size_t)(en - st)); d->len = en - st; }
                else { d->s = dup_n(A, "", 0); d->len = 0; }
                break;
            }
            case TPLX_OP_SINDEX: {
                int64_t idx = b->i;
                if (idx < 0) idx += a->len;
                if (idx < 0 || idx >= a->len) RAISE(TPLX_EC_INDEXERROR);
                d->s = dup_n(A, a->s + idx, 1);
                d->len = 1;
                break;
            }
            case TPLX_OP_SLOWER: case TPLX_OP_SUPPER: {
                char *r = rtmalloc(A, (size_t)a->len + 1);
                for (int64_t i = 0; i < a->len; ++i)
                    r[i] = (char)(in->op == TPLX_OP_SLOWER ? tolower((unsigned char)a->s[i]) : toupper((unsigned char)a->s[i]));
                r[a->len] = 0;
                d->s = r;
                d->len = a->len;
                break;
            }
            case TPLX_OP_SSTRIP: {
                int64_t i = 0, e = a->len;
                const char *ws = " \t\n\r\x0b\x0c";
                if (in->flags & 1) while (i < e && strchr(ws, a->s[i]) && a->s[i]) ++i;
                if (in->flags & 2) while (e > i && strchr(ws, a->s[e - 1]) && a->s[e - 1]) --e;
                d->s = dup_n(A, a->s + i, (size_t)(e - i));
                d->len = e - i;
                break;
            }
            case TPLX_OP_SCONCAT: {
                if (a->len == 0) { *d = *b; break; }
                if (b->len == 0) { *d = *a; break; }
                char *r = rtmalloc(A, (size_t)(a->len + b->len) + 1);
                memcpy(r, a->s, (size_t)a->len);
                memcpy(r + a->len, b->s, (size_t)b->len + 1);
                d->s = r;
                d->len = a->len + b->len;
                break;
            }
            case TPLX_OP_SREPLACE: {
                size_t rl = 0;
                char *r = o_replace(A, a->s, b->s, c->s, &rl);
                d->s = r;
                d->len = (int64_t)rl;
                break;
            }
            case TPLX_OP_SFMTD: {
                /* snprintf with a C %d conversion, which consumes an int (formatStr, BlockGeneratorVisitor.cc:675-775) */
                char fmt[32], out[64];
                int nlen;
                if (in->flags & 2) { /* str.format / f-string: fmt keeps the whole 64-bit value (Runtime.cc:544-607) */
                    if (in->flags & 1) snprintf(fmt, sizeof fmt, "%%0%dlld", (int)in->imm);
                    else if (in->imm) snprintf(fmt, sizeof fmt, "%%%dlld", (int)in->imm);
                    else snprintf(fmt, sizeof fmt, "%%lld");
                    nlen = snprintf(out, sizeof out, fmt, (long long)a->i);
                } else {
                    if (in->flags & 1) snprintf(fmt, sizeof fmt, "%%0%dd", (int)in->imm);
                    else if (in->imm) snprintf(fmt, sizeof fmt, "%%%dd", (int)in->imm);
                    else snprintf(fmt, sizeof fmt, "%%d");
                    nlen = snprintf(out, sizeof out, fmt, (int)a->i);
                }
                d->s = dup_n(A, out, (size_t)nlen);
                d->len = nlen;
                break;
            }
            case TPLX_OP_I2S: {
                char out[32];
                int nlen = snprintf(out, sizeof out, "%lld", (long long)a->i);
                d->s = dup_n(A, out, (size_t)nlen);
                d->len = nlen;
                break;
            }
            case TPLX_OP_S2I: {
                int64_t v = 0;
                if (tplx_o_atoi64(a->s, a->len, &v)) RAISE(TPLX_EC_VALUEERROR);
                d->i = v;
                break;
            }
            case TPLX_OP_S2F: { /* float(str): the runtime's fast_atod wrapper (csv_oracle.c restates it) */
                extern int csv_oracle_atod(const char *s, double *out);
                double dv = 0;
                if (csv_oracle_atod(a->s, &dv)) RAISE(TPLX_EC_VALUEERROR);
                d->i = as_i(dv);
                break;
            }
            case TPLX_OP_FILTER: if (a->i == 0) return 1; break;
            case TPLX_OP_RAISE: RAISE(in->imm);
            default: *ec = -1; *opidx = in->opidx; return 2;
        }
#undef RAISE
    }
    return 0;
}

static void res_grow_rows(tplx_oresult *r) {
    uint64_t ncap = r->out_cap ? r->out_cap * 2 : 1024;
    for (uint64_t c = 0; c < r->n_cols; ++c) {
        if (r->col_types[c] == TPLX_T_STR) r->offsets[c] = (uint32_t *)realloc(r->offsets[c], (ncap + 1) * 4);
        else r->fixed[c] = (int64_t *)realloc(r->fixed[c], ncap * 8);
    }
    r->out_cap = ncap;
}

static uint64_t acc_identity(uint8_t kind) {
    switch (kind) {
        case TPLX_ACC_SUM_I64: case TPLX_ACC_SUM_F64: return 0;
        case TPLX_ACC_MIN_I64: return 0x7FFFFFFFFFFFFFFFull;
        case TPLX_ACC_MAX_I64: return 0x8000000000000000ull;
        case TPLX_ACC_MIN_F64: return 0x7FF0000000000000ull;
        default: return 0xFFF0000000000000ull;
    }
}
static uint64_t acc_combine(uint8_t kind, uint64_t a, uint64_t b) {
    switch (kind) {
        case TPLX_ACC_SUM_I64: return a + b;
        case TPLX_ACC_SUM_F64: return (uint64_t)as_i(as_f((int64_t)a) + as_f((int64_t)b));
        case TPLX_ACC_MIN_I64: return (int64_t)b < (int64_t)a ? b : a;
        case TPLX_ACC_MAX_I64: return (int64_t)b > (int64_t)a ? b : a;
        case TPLX_ACC_MIN_F64: return as_f((int64_t)b) < as_f((int64_t)a) ? b : a;
        default: return as_f((int64_t)b) > as_f((int64_t)a) ? b : a;
    }
}

/* ---- hash endpoint (simple chained map; only the result SET matters) --------------------------------- */
typedef struct hent {
    struct hent *next;
    uint64_t hash;
    uint32_t blob_len;
    uint64_t first_row;
    uint64_t acc[TPLX_MAX_ACCS];
    char blob[];
} hent;
typedef struct hmap {
    hent **buckets;
    uint64_t nb, n;
    hent **order; /* first-seen order */
    uint64_t order_cap;
} hmap;

static uint64_t fnv(const char *p, size_t n) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; ++i) h = (h ^ (unsigned char)p[i]) * 0x100000001b3ull;
    return h;
}
static hent *hmap_get(hmap *m, const char *blob, uint32_t len, const ostage *S) {
    if (!m->buckets) {
        m->nb = 1 << 16;
        m->buckets = (hent **)calloc(m->nb, sizeof(hent *));
    }
    if (m->n > m->nb) { /* grow */
        uint64_t nnb = m->nb * 4;
        hent **nbk = (hent **)calloc(nnb, sizeof(hent *));
        for (uint64_t i = 0; i < m->nb; ++i)
            for (hent *e = m->buckets[i]; e;) {
                hent *nx = e->next;
                e->next = nbk[e->hash & (nnb - 1)];
                nbk[e->hash & (nnb - 1)] = e;
                e = nx;
            }
        free(m->buckets);
        m->buckets = nbk;
        m->nb = nnb;
    }
    uint64_t h = fnv(blob, len);
    for (hent *e = m->buckets[h & (m->nb - 1)]; e; e = e->next)
        if (e->hash == h && e->blob_len == len && memcmp(e->blob, blob, len) == 0) return e;
    hent *e = (hent *)malloc(sizeof(hent) + len);
    e->hash = h;
    e->blob_len = len;
    memcpy(e->blob, blob, len);
    for (uint32_t k = 0; k < S->h.n_accs; ++k) e->acc[k] = acc_identity(S->accs[k].kind);
    e->next = m->buckets[h & (m->nb - 1)];
    m->buckets[h & (m->nb - 1)] = e;
    if (m->n == m->order_cap) {
        m->order_cap = m->order_cap ? m->order_cap * 2 : 1024;
        m->order = (hent **)realloc(m->order, m->order_cap * sizeof(hent *));
    }
    m->order[m->n++] = e;
    return e;
}

static void out_append_str(tplx_oresult *r, uint64_t c, const char *s, uint64_t len) {
    if (r->bytes_len[c] + len > r->bytes_cap[c]) {
        uint64_t nc = r->bytes_cap[c] ? r->bytes_cap[c] * 2 : 4096;
        while (nc < r->bytes_len[c] + len) nc *= 2;
        r->bytes[c] = (char *)realloc(r->bytes[c], nc);
        r->bytes_cap[c] = nc;
    }
    memcpy(r->bytes[c] + r->bytes_len[c], s, len);
    r->bytes_len[c] += len;
}

/*
 * Run a stage over a column block the way one reference TransformTask would:
 * rows in order; per row: pipeline; kept rows appended to the output, exceptions get
 * rowNo = outputRowCounter++ shared with normal rows (TransformTask.cc:764,885).
 * tile_rows / nt / fin_nt parameterise the documented reduction tree for acc_tree.
 */
int tplx_oracle_run(const void *desc, uint64_t desc_bytes, const tplx_ocol *cols, uint64_t n_rows, int64_t first_row_no,
                    uint32_t tile_R, uint32_t tile_NT, uint32_t fin_NT, tplx_oresult *res) {
    ostage S;
    if (parse_stage(desc, desc_bytes, &S)) return -1;
    memset(res, 0, sizeof(*res));
    const uint32_t na = S.h.n_accs;
    res->n_accs = na;
    const int is_hash = S.h.endpoint == TPLX_EP_HASH;
    res->n_cols = is_hash ? S.h.n_keys + na : (uint64_t)(S.h.n_out_cols - S.h.hidden_out_cols);
    /* Option[T] outputs: the hidden `is None` companions (tplx_outcol.null_of) sit right behind the visible columns and are
     * reported too, so that a test can fold them back into None values */
    if (!is_hash)
        for (uint64_t c = res->n_cols; c < S.h.n_out_cols && S.out_cols[c].null_of; ++c) res->n_cols = c + 1;
    for (uint64_t c = 0; c < S.h.n_out_cols && c < res->n_cols; ++c) res->col_types[c] = S.out_cols[c].type;
    if (is_hash)
        for (uint32_t k = 0; k < na; ++k) {
            uint8_t kind = S.accs[k].kind;
            res->col_types[S.h.n_keys + k] = (kind == TPLX_ACC_SUM_F64 || kind == TPLX_ACC_MIN_F64 || kind == TPLX_ACC_MAX_F64) ? TPLX_T_F64 : TPLX_T_I64;
        }
    oval *regs = (oval *)calloc((size_t)S.h.n_slots + 2, sizeof(oval));
    arena A = {0};
    hmap H = {0};
    int64_t row_counter = first_row_no;
    uint64_t seq[TPLX_MAX_ACCS];
    for (uint32_t k = 0; k < na; ++k) seq[k] = (uint64_t)S.accs[k].init; /* intermediate starts from the initial value */
    /* reduction tree state */
    const uint64_t T = (uint64_t)tile_R * tile_NT;
    const uint64_t n_tiles = T ? (n_rows + T - 1) / T : 0;
    uint64_t *thr = NULL, *tile_part = NULL;
    if (na && S.h.endpoint == TPLX_EP_AGGREGATE && T) {
        thr = (uint64_t *)malloc((size_t)tile_NT * na * 8);
        tile_part = (uint64_t *)malloc((size_t)(n_tiles ? n_tiles : 1) * na * 8);
    }
    char *keyblob = (char *)malloc(1 << 16);
    for (uint64_t row = 0; row < n_rows; ++row) {
        if (thr && row % T == 0)
            for (uint32_t t = 0; t < tile_NT; ++t)
                for (uint32_t k = 0; k < na; ++k) thr[(size_t)t * na + k] = acc_identity(S.accs[k].kind);
        int64_t ec = 0;
        uint32_t opidx = 0;
        int st = eval_row(&S, cols, row, regs, &A, &ec, &opidx);
        if (st == 0) {
            if (S.h.endpoint == TPLX_EP_MEMORY) {
                if (res->n_out == res->out_cap) res_grow_rows(res);
                for (uint64_t c = 0; c < res->n_cols; ++c) {
                    const oval *v = &regs[S.out_cols[c].slot];
                    if (res->col_types[c] == TPLX_T_STR) {
                        res->offsets[c][res->n_out] = (uint32_t)res->bytes_len[c];
                        out_append_str(res, c, v->s, (uint64_t)v->len);
                    } else res->fixed[c][res->n_out] = v->i;
                }
                res->n_out++;
                row_counter++;
            } else if (S.h.endpoint == TPLX_EP_AGGREGATE) {
                for (uint32_t k = 0; k < na; ++k) {
                    uint64_t v = (uint64_t)regs[S.accs[k].slot].i;
                    seq[k] = acc_combine(S.accs[k].kind, seq[k], v);
                    if (thr) {
                        uint64_t lr = row % T;
                        uint64_t *p = &thr[(size_t)(lr % tile_NT) * na + k];
                        *p = acc_combine(S.accs[k].kind, *p, v);
                    }
                }
            } else {
                uint32_t bl = 0;
                for (uint32_t k = 0; k < S.h.n_keys; ++k) {
                    const oval *v = &regs[S.out_cols[k].slot];
                    if (S.out_cols[k].type == TPLX_T_STR) {
                        uint32_t l = (uint32_t)v->len;
                        memcpy(keyblob + bl, &l, 4);
                        memcpy(keyblob + bl + 4, v->s, l);
                        bl += 4 + l;
                    } else {
                        memcpy(keyblob + bl, &v->i, 8);
                        bl += 8;
                    }
                }
                hent *e = hmap_get(&H, keyblob, bl, &S);
                for (uint32_t k = 0; k < na; ++k) e->acc[k] = acc_combine(S.accs[k].kind, e->acc[k], (uint64_t)regs[S.accs[k].slot].i);
            }
        } else if (st == 2) {
            if (res->n_exc == res->exc_cap) {
                res->exc_cap = res->exc_cap ? res->exc_cap * 2 : 256;
                res->exc = (tplx_oexc *)realloc(res->exc, res->exc_cap * sizeof(tplx_oexc));
            }
            tplx_oexc *x = &res->exc[res->n_exc++];
            x->row = (int64_t)row;
            x->row_no = S.h.endpoint == TPLX_EP_MEMORY ? row_counter++ : (int64_t)(res->n_exc - 1);
            x->code = ec;
            x->op_id = S.h.n_ops ? S.opids[opidx] : 0;
        }
        rtfree_all(&A);
        /* end of tile: warp tree (shfl_down 16..1 on 32-lane groups), warps sequential */
        if (thr && ((row + 1) % T == 0 || row + 1 == n_rows)) {
            uint64_t tile = row / T;
            for (uint32_t k = 0; k < na; ++k) {
                uint8_t kind = S.accs[k].kind;
                uint64_t tilev = 0;
                for (uint32_t w = 0; w < tile_NT / 32; ++w) {
                    uint64_t lane[32];
                    for (uint32_t l = 0; l < 32; ++l) lane[l] = thr[(size_t)(w * 32 + l) * na + k];
                    for (uint32_t o = 16; o; o >>= 1)
                        for (uint32_t l = 0; l + o < 32 && l < o; ++l) lane[l] = acc_combine(kind, lane[l], lane[l + o]);
                    tilev = w == 0 ? lane[0] : acc_combine(kind, tilev, lane[0]);
                }
                tile_part[(size_t)tile * na + k] = tilev;
            }
        }
    }
    if (S.h.endpoint == TPLX_EP_AGGREGATE) {
        for (uint32_t k = 0; k < na; ++k) {
            res->acc_seq[k] = (int64_t)seq[k];
            if (thr) {
                uint8_t kind = S.accs[k].kind;
                uint64_t *f = (uint64_t *)malloc((size_t)fin_NT * 8);
                for (uint32_t t = 0; t < fin_NT; ++t) {
                    uint64_t v = acc_identity(kind);
                    for (uint64_t tile = t; tile < n_tiles; tile += fin_NT) v = acc_combine(kind, v, tile_part[(size_t)tile * na + k]);
                    f[t] = v;
                }
                uint64_t tot = 0;
                for (uint32_t w = 0; w < fin_NT / 32; ++w) {
                    uint64_t lane[32];
                    for (uint32_t l = 0; l < 32; ++l) lane[l] = f[w * 32 + l];
                    for (uint32_t o = 16; o; o >>= 1)
                        for (uint32_t l = 0; l + o < 32 && l < o; ++l) lane[l] = acc_combine(kind, lane[l], lane[l + o]);
                    tot = w == 0 ? lane[0] : acc_combine(kind, tot, lane[0]);
                }
                res->acc_tree[k] = (int64_t)acc_combine(kind, (uint64_t)S.accs[k].init, tot);
                free(f);
            }
        }
        res->n_out = 1;
    }
    if (is_hash) {
        /* table -> rows, first-seen order; bucket started from init and combine(init, v) runs once per group
         * (TransformTask.cc:358-375, LocalBackend.cc:2148-2217) */
        for (uint64_t i = 0; i < H.n; ++i) {
            hent *e = H.order[i];
            if (res->n_out == res->out_cap) res_grow_rows(res);
            const char *p = e->blob;
            for (uint32_t k = 0; k < S.h.n_keys; ++k) {
                if (S.out_cols[k].type == TPLX_T_STR) {
                    uint32_t l;
                    memcpy(&l, p, 4);
                    res->offsets[k][res->n_out] = (uint32_t)res->bytes_len[k];
                    out_append_str(res, k, p + 4, l);
                    p += 4 + l;
                } else {
                    memcpy(&res->fixed[k][res->n_out], p, 8);
                    p += 8;
                }
            }
            for (uint32_t k = 0; k < na; ++k) {
                uint64_t v = acc_combine(S.accs[k].kind, (uint64_t)S.accs[k].init, e->acc[k]);
                v = acc_combine(S.accs[k].kind, (uint64_t)S.accs[k].init, v);
                res->fixed[S.h.n_keys + k][res->n_out] = (int64_t)v;
            }
            res->n_out++;
        }
        for (uint64_t i = 0; i < H.n; ++i) free(H.order[i]);
        free(H.order);
        free(H.buckets);
    }
    /* terminal offsets */
    if (res->out_cap == 0) res_grow_rows(res);
    for (uint64_t c = 0; c < res->n_cols; ++c)
        if (res->col_types[c] == TPLX_T_STR) res->offsets[c][res->n_out] = (uint32_t)res->bytes_len[c];
    free(keyblob);
    free(thr);
    free(tile_part);
    free(regs);
    arena_destroy(&A);
    return 0;
}

void tplx_oracle_free(tplx_oresult *r) {
    for (uint64_t c = 0; c < TPLX_MAX_COLS; ++c) {
        free(r->fixed[c]);
        free(r->offsets[c]);
        free(r->bytes[c]);
    }
    free(r->exc);
    memset(r, 0, sizeof(*r));
}

/* ------------------------------------------------------------------------------------------- */
/* row / partition / exception byte formats                                                     */
/* ------------------------------------------------------------------------------------------- */
static int o_present(const tplx_ocol *c, uint64_t i) { return !c->valid || ((c->valid[i >> 5] >> (i & 31)) & 1u); }
static uint64_t o_bitmap_bytes(const tplx_ocol *cols, uint32_t n_cols) { /* calcBitmapSize, Serializer.cc:29-41 */
    uint32_t n_opt = 0;
    for (uint32_t c = 0; c < n_cols; ++c) n_opt += cols[c].valid != NULL;
    return (uint64_t)((n_opt + 63) / 64) * 8;
}
static uint64_t o_row_size(const tplx_ocol *cols, uint32_t n_cols, uint64_t i) {
    uint64_t sz = 8ull * n_cols + o_bitmap_bytes(cols, n_cols), var = 0;
    int has = 0;
    for (uint32_t c = 0; c < n_cols; ++c)
        if (cols[c].type == TPLX_T_STR) {
            has = 1;
            if (o_present(&cols[c], i)) var += (uint64_t)(cols[c].offsets[i + 1] - cols[c].offsets[i]) + 1;
        }
    return sz + (has ? 8 + var : 0);
}
/* Serializer::serialize (utils/src/Serializer.cc:1016-1117): bitmap of the Option fields (bit k = the k-th Option field is None,
 * :1041-1059), one 8-byte slot per field, total var-len bytes, var-len payload. A None string keeps its slot (offset of the next
 * var field, size 0) and contributes no bytes (appendWithoutInference(option<string>), :313-338); a None number has a zero slot. */
static uint64_t o_write_row(const tplx_ocol *cols, uint32_t n_cols, uint64_t i, uint8_t *dst) {
    const uint64_t bm = o_bitmap_bytes(cols, n_cols);
    if (bm) {
        memset(dst, 0, bm);
        uint32_t k = 0;
        for (uint32_t c = 0; c < n_cols; ++c) {
            if (!cols[c].valid) continue;
            if (!o_present(&cols[c], i)) dst[k / 8] |= (uint8_t)(1u << (k % 8)); /* little endian int64 words */
            ++k;
        }
        dst += bm;
    }
    uint64_t var_off = 8ull * n_cols + 8, total = 0;
    int has = 0;
    for (uint32_t c = 0; c < n_cols; ++c) {
        if (cols[c].type == TPLX_T_STR) {
            has = 1;
            if (!o_present(&cols[c], i)) {
                int64_t info = (int64_t)((var_off - 8ull * c) & 0xFFFFFFFFull);
                memcpy(dst + 8ull * c, &info, 8);
                continue;
            }
            uint32_t o0 = cols[c].offsets[i], len = cols[c].offsets[i + 1] - o0;
            int64_t info = (int64_t)((var_off - 8ull * c) & 0xFFFFFFFFull) | ((int64_t)(len + 1) << 32);
            memcpy(dst + 8ull * c, &info, 8);
            memcpy(dst + var_off, (const char *)cols[c].data + o0, len);
            dst[var_off + len] = 0;
            var_off += len + 1;
            total += len + 1;
        } else if (!o_present(&cols[c], i)) {
            memset(dst + 8ull * c, 0, 8);
        } else memcpy(dst + 8ull * c, (const int64_t *)cols[c].data + i, 8);
    }
    if (has) memcpy(dst + 8ull * n_cols, &total, 8);
    return bm + 8ull * n_cols + (has ? 8 + total : 0);
}

/* rows [0,n) of a column block -> partitions, split like rowToMemorySink (TransformTask.h:47-92).
 * buf NULL: only compute sizes. part_offsets gets n_parts+1 entries. Returns total bytes. */
uint64_t tplx_oracle_to_partitions(const tplx_ocol *cols, uint32_t n_cols, uint64_t n, uint64_t partition_bytes, uint8_t *buf,
                                   uint64_t *part_offsets, uint32_t max_parts, uint32_t *n_parts) {
    const uint64_t capacity = partition_bytes - 8;
    uint64_t pos = 0, written = 0, rows_in_part = 0, part_start = 0;
    uint32_t np = 0;
    int open = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t sz = o_row_size(cols, n_cols, i);
        if (!open || written + sz > capacity) {
            if (open && buf) memcpy(buf + part_start, &rows_in_part, 8);
            if (part_offsets && np < max_parts) part_offsets[np] = pos;
            np++;
            part_start = pos;
            pos += 8;
            written = 0;
            rows_in_part = 0;
            open = 1;
        }
        if (buf) o_write_row(cols, n_cols, i, buf + pos);
        pos += sz;
        written += sz;
        rows_in_part++;
    }
    if (!open) { /* one empty partition */
        if (part_offsets && np < max_parts) part_offsets[np] = pos;
        np++;
        if (buf) memset(buf + pos, 0, 8);
        pos += 8;
    } else if (buf) memcpy(buf + part_start, &rows_in_part, 8);
    if (part_offsets && np <= max_parts) part_offsets[np] = pos;
    *n_parts = np;
    return pos;
}

/* exception partition: int64 numRows, then [rowNo, ecCode, opID, size, input row] (IExceptionableTask.h:22-36) */
uint64_t tplx_oracle_exception_partition(const tplx_ocol *in_cols, uint32_t n_cols, const tplx_oexc *exc, uint64_t n_exc, uint8_t *buf) {
    uint64_t pos = 8;
    if (buf) memcpy(buf, &n_exc, 8);
    for (uint64_t i = 0; i < n_exc; ++i) {
        uint64_t sz = o_row_size(in_cols, n_cols, (uint64_t)exc[i].row);
        if (buf) {
            int64_t h[4] = {exc[i].row_no, exc[i].code, exc[i].op_id, (int64_t)sz};
            memcpy(buf + pos, h, 32);
            o_write_row(in_cols, n_cols, (uint64_t)exc[i].row, buf + pos + 32);
        }
        pos += 32 + sz;
    }
    return pos;
}
