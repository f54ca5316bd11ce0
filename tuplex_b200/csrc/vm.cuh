// vm.cuh — per-row evaluator of the tplx op program on sm_100a.
//
// Replaces the body of the LLVM-generated processRow pipeline
// (reference tuplex/core/src/physical/PipelineBuilder.cc:90-170,565-1025) and the UDF code that
// BlockGeneratorVisitor/FunctionRegistry emit. One thread evaluates one row at a time; the
// program is warp-uniform, so decode is uniform and only the data path diverges.
// The register file lives in shared memory as regs[slot][thread] (8-byte slots; bank-conflict
// free because consecutive threads touch consecutive 8-byte words).
//
// Semantics cited per op in include/tplx_ir.h; the CPU restatement that this file must match
// bit-for-bit is oracle/tplx_oracle.c.
#pragma once
#include <stdint.h>
#include "../../include/tplx_ir.h"

namespace tplx {

struct StrV {
    const uint8_t *p;
    uint32_t len;
    uint32_t flags;
};

struct VMThread {
    bool alive;          // row still in the pipeline (valid, not filtered, no exception)
    uint32_t exc_code;   // 0 = none
    uint32_t exc_op;     // operator index of the raising op
    uint8_t *scratch;    // per-thread arena for materialised strings
    uint32_t scr_used;
    uint32_t scr_cap;
};

struct ColIn {
    const void *data;
    const uint32_t *offsets;
    uint64_t type;
};

__device__ __forceinline__ uint8_t sch(const StrV &s, uint32_t i) {
    uint8_t c = s.p[i];
    // ASCII-only case mapping, tolower/toupper in the C locale (runtime/src/StringFunctions.cc:71-110)
    if (s.flags == TPLX_SF_LOWER) {
        if ((uint8_t)(c - 'A') < 26u) c += 32;
    } else if (s.flags == TPLX_SF_UPPER) {
        if ((uint8_t)(c - 'a') < 26u) c -= 32;
    }
    return c;
}

__device__ __forceinline__ bool is_pyspace(uint8_t c) {
    // string.whitespace = ' \t\n\r\x0b\x0c' (runtime/src/Runtime.cc:322-333)
    return c == ' ' || (c >= 9 && c <= 13);
}

// strstr semantics (FunctionRegistry.cc:2165-2188): first occurrence or -1; empty needle -> 0
__device__ __noinline__ int64_t str_find(const StrV &h, const StrV &n) {
    if (n.len == 0) return 0;
    if (n.len > h.len) return -1;
    uint8_t n0 = sch(n, 0);
    uint32_t last = h.len - n.len;
    for (uint32_t i = 0; i <= last; ++i) {
        if (sch(h, i) != n0) continue;
        uint32_t j = 1;
        while (j < n.len && sch(h, i + j) == sch(n, j)) ++j;
        if (j == n.len) return (int64_t)i;
    }
    return -1;
}

// std::string::rfind (runtime/src/Runtime.cc:387-397): last occurrence or -1; empty needle -> len
__device__ __noinline__ int64_t str_rfind(const StrV &h, const StrV &n) {
    if (n.len > h.len) return -1;
    if (n.len == 0) return (int64_t)h.len;
    uint8_t n0 = sch(n, 0);
    for (int64_t i = (int64_t)(h.len - n.len); i >= 0; --i) {
        if (sch(h, (uint32_t)i) != n0) continue;
        uint32_t j = 1;
        while (j < n.len && sch(h, (uint32_t)i + j) == sch(n, j)) ++j;
        if (j == n.len) return i;
    }
    return -1;
}

__device__ __forceinline__ bool str_eq(const StrV &a, const StrV &b) {
    if (a.len != b.len) return false;
    for (uint32_t i = 0; i < a.len; ++i)
        if (sch(a, i) != sch(b, i)) return false;
    return true;
}

// Python slice index normalisation for stride +1 (BlockGeneratorVisitor.cc:4618-4690)
__device__ __forceinline__ int64_t slice_index(int64_t idx, int64_t len) {
    if (idx < -len) return 0;
    if (idx <= -1) return idx + len;
    if (idx < len) return idx;
    return len;
}

__device__ __forceinline__ uint8_t *scratch_alloc(VMThread &t, uint32_t n, uint32_t opidx) {
    if (t.scr_used + n > t.scr_cap) {
        // internal limit, not a Python error: hand the row to the resolve path
        t.exc_code = TPLX_EC_NORMALCASEVIOLATION;
        t.exc_op = opidx;
        t.alive = false;
        return nullptr;
    }
    uint8_t *p = t.scratch + t.scr_used;
    t.scr_used += n;
    return p;
}

__device__ __forceinline__ void raise_exc(VMThread &t, uint32_t code, uint32_t opidx) {
    t.exc_code = code;
    t.exc_op = opidx;
    t.alive = false;
}

// floor division / modulo with sign fix-up (codegen/src/LLVMEnvironment.cc:1377-1430)
__device__ __forceinline__ int64_t floordiv_i64(int64_t x, int64_t y) {
    int64_t q = x / y, r = x % y;
    if (r != 0 && ((r < 0) != (y < 0))) --q;
    return q;
}
__device__ __forceinline__ int64_t floormod_i64(int64_t x, int64_t y) {
    int64_t r = x % y;
    if (r != 0 && ((r < 0) != (y < 0))) r += y;
    return r;
}

template <int NT>
struct VM {
    // regs points at regs[0][tid]; slot s lives at regs[s*NT]
    static __device__ __forceinline__ uint64_t &R(uint64_t *regs, uint32_t s) { return regs[s * NT]; }
    static __device__ __forceinline__ double RF(uint64_t *regs, uint32_t s) {
        return __longlong_as_double((long long)regs[s * NT]);
    }
    static __device__ __forceinline__ StrV RS(uint64_t *regs, uint32_t s) {
        StrV v;
        v.p = (const uint8_t *)regs[s * NT];
        uint64_t m = regs[(s + 1) * NT];
        v.len = (uint32_t)m;
        v.flags = (uint32_t)(m >> 32);
        return v;
    }
    static __device__ __forceinline__ void WS(uint64_t *regs, uint32_t s, const uint8_t *p, uint32_t len,
                                              uint32_t flags) {
        regs[s * NT] = (uint64_t)p;
        regs[(s + 1) * NT] = (uint64_t)len | ((uint64_t)flags << 32);
    }

    // Run instructions [pc0, pc1) for this thread's row.
    static __device__ void run(const tplx_instr *__restrict__ prog, uint32_t pc0, uint32_t pc1,
                               uint64_t *__restrict__ regs, const ColIn *__restrict__ cols, uint64_t row,
                               const uint8_t *__restrict__ cpool, VMThread &t) {
        for (uint32_t pc = pc0; pc < pc1; ++pc) {
            // uniform fetch (broadcast from shared memory)
            const uint4 w0 = *reinterpret_cast<const uint4 *>(&prog[pc]);
            const uint32_t op = w0.x & 0xFF;
            const uint32_t flags = (w0.x >> 8) & 0xFF;
            const uint32_t dst = w0.x >> 16;
            const uint32_t a = w0.y & 0xFFFF, b = w0.y >> 16;
            const uint32_t c = w0.z & 0xFFFF, guard = w0.z >> 16;
            const uint32_t opidx = w0.w & 0xFFFF;
            bool act = t.alive;
            if (guard != TPLX_NOSLOT) act = act && (R(regs, guard) != 0);
            if (!__any_sync(0xFFFFFFFFu, act)) continue;  // whole warp idle for this op
            const int64_t imm = prog[pc].imm;
            if (act) {
                switch (op) {
                    case TPLX_OP_LDCOL: {
                        const ColIn &ci = cols[imm];
                        if (flags == TPLX_T_STR) {
                            uint32_t o0 = ci.offsets[row], o1 = ci.offsets[row + 1];
                            WS(regs, dst, (const uint8_t *)ci.data + o0, o1 - o0, 0);
                        } else {
                            R(regs, dst) = ((const uint64_t *)ci.data)[row];
                        }
                        break;
                    }
                    case TPLX_OP_LDI: R(regs, dst) = (uint64_t)imm; break;
                    case TPLX_OP_LDROW: R(regs, dst) = row; break;
                    case TPLX_OP_LDS: WS(regs, dst, cpool + imm, (uint32_t)prog[pc].imm2, 0); break;
                    case TPLX_OP_MOV:
                        R(regs, dst) = R(regs, a);
                        if (flags == 2) R(regs, dst + 1) = R(regs, a + 1);
                        break;
                    case TPLX_OP_SEL: {
                        uint32_t s = R(regs, c) ? a : b;
                        uint64_t v0 = R(regs, s);
                        if (flags == 2) {
                            uint64_t v1 = R(regs, s + 1);
                            R(regs, dst + 1) = v1;
                        }
                        R(regs, dst) = v0;
                        break;
                    }
                    case TPLX_OP_IADD: R(regs, dst) = R(regs, a) + R(regs, b); break;
                    case TPLX_OP_ISUB: R(regs, dst) = R(regs, a) - R(regs, b); break;
                    case TPLX_OP_IMUL: R(regs, dst) = R(regs, a) * R(regs, b); break;
                    case TPLX_OP_IFLOORDIV: {
                        int64_t y = (int64_t)R(regs, b);
                        if (y == 0) { raise_exc(t, TPLX_EC_ZERODIVISIONERROR, opidx); break; }
                        R(regs, dst) = (uint64_t)floordiv_i64((int64_t)R(regs, a), y);
                        break;
                    }
                    case TPLX_OP_IMOD: {
                        int64_t y = (int64_t)R(regs, b);
                        if (y == 0) { raise_exc(t, TPLX_EC_ZERODIVISIONERROR, opidx); break; }
                        R(regs, dst) = (uint64_t)floormod_i64((int64_t)R(regs, a), y);
                        break;
                    }
                    case TPLX_OP_INEG: R(regs, dst) = (uint64_t)0 - R(regs, a); break;
                    case TPLX_OP_IAND: R(regs, dst) = R(regs, a) & R(regs, b); break;
                    case TPLX_OP_IOR: R(regs, dst) = R(regs, a) | R(regs, b); break;
                    case TPLX_OP_IXOR: R(regs, dst) = R(regs, a) ^ R(regs, b); break;
                    case TPLX_OP_ISHL: R(regs, dst) = R(regs, a) << (R(regs, b) & 63); break;
                    case TPLX_OP_ISHR: R(regs, dst) = (uint64_t)((int64_t)R(regs, a) >> (R(regs, b) & 63)); break;
                    case TPLX_OP_IABS: {
                        int64_t x = (int64_t)R(regs, a);
                        R(regs, dst) = (uint64_t)(x < 0 ? (uint64_t)0 - (uint64_t)x : (uint64_t)x);
                        break;
                    }
                    // single-rounded IEEE ops; the _rn intrinsics are never contracted into FMA
                    case TPLX_OP_FADD: R(regs, dst) = (uint64_t)__double_as_longlong(__dadd_rn(RF(regs, a), RF(regs, b))); break;
                    case TPLX_OP_FSUB: R(regs, dst) = (uint64_t)__double_as_longlong(__dsub_rn(RF(regs, a), RF(regs, b))); break;
                    case TPLX_OP_FMUL: R(regs, dst) = (uint64_t)__double_as_longlong(__dmul_rn(RF(regs, a), RF(regs, b))); break;
                    case TPLX_OP_FDIV: {
                        double y = RF(regs, b);
                        if (y == 0.0) { raise_exc(t, TPLX_EC_ZERODIVISIONERROR, opidx); break; }
                        R(regs, dst) = (uint64_t)__double_as_longlong(__ddiv_rn(RF(regs, a), y));
                        break;
                    }
                    case TPLX_OP_FMOD: {
                        double x = RF(regs, a), y = RF(regs, b);
                        if (y == 0.0) { raise_exc(t, TPLX_EC_ZERODIVISIONERROR, opidx); break; }
                        double r = fmod(x, y);  // == LLVM frem, exact
                        if (r != 0.0 && ((r < 0.0) != (y < 0.0))) r = __dadd_rn(r, y);
                        R(regs, dst) = (uint64_t)__double_as_longlong(r);
                        break;
                    }
                    case TPLX_OP_FNEG: R(regs, dst) = R(regs, a) ^ 0x8000000000000000ull; break;
                    case TPLX_OP_FABS: R(regs, dst) = R(regs, a) & 0x7FFFFFFFFFFFFFFFull; break;
                    case TPLX_OP_FFLOORDIV: {
                        double y = RF(regs, b);
                        if (y == 0.0) { raise_exc(t, TPLX_EC_ZERODIVISIONERROR, opidx); break; }
                        int64_t xi = (int64_t)RF(regs, a), yi = (int64_t)y;
                        if (yi == 0) { raise_exc(t, TPLX_EC_ZERODIVISIONERROR, opidx); break; }
                        R(regs, dst) = (uint64_t)__double_as_longlong((double)floordiv_i64(xi, yi));
                        break;
                    }
                    case TPLX_OP_I2F: R(regs, dst) = (uint64_t)__double_as_longlong((double)(int64_t)R(regs, a)); break;
                    case TPLX_OP_F2I: R(regs, dst) = (uint64_t)(int64_t)RF(regs, a); break;
                    case TPLX_OP_ICMP: {
                        int64_t x = (int64_t)R(regs, a), y = (int64_t)R(regs, b);
                        bool r;
                        switch (flags) {
                            case TPLX_CMP_EQ: r = x == y; break;
                            case TPLX_CMP_NE: r = x != y; break;
                            case TPLX_CMP_LT: r = x < y; break;
                            case TPLX_CMP_LE: r = x <= y; break;
                            case TPLX_CMP_GT: r = x > y; break;
                            default: r = x >= y; break;
                        }
                        R(regs, dst) = r;
                        break;
                    }
                    case TPLX_OP_FCMP: {
                        double x = RF(regs, a), y = RF(regs, b);
                        bool r;
                        switch (flags) {  // ordered predicates: false if either is NaN
                            case TPLX_CMP_EQ: r = x == y; break;
                            case TPLX_CMP_NE: r = (x < y) || (x > y); break;  // FCMP_ONE
                            case TPLX_CMP_LT: r = x < y; break;
                            case TPLX_CMP_LE: r = x <= y; break;
                            case TPLX_CMP_GT: r = x > y; break;
                            default: r = x >= y; break;
                        }
                        R(regs, dst) = r;
                        break;
                    }
                    case TPLX_OP_BAND: R(regs, dst) = (R(regs, a) != 0) & (R(regs, b) != 0); break;
                    case TPLX_OP_BOR: R(regs, dst) = (R(regs, a) != 0) | (R(regs, b) != 0); break;
                    case TPLX_OP_BNOT: R(regs, dst) = (R(regs, a) == 0); break;
                    case TPLX_OP_SLEN: R(regs, dst) = (uint64_t)(uint32_t)R(regs, a + 1); break;
                    case TPLX_OP_SFIND: R(regs, dst) = (uint64_t)str_find(RS(regs, a), RS(regs, b)); break;
                    case TPLX_OP_SRFIND: R(regs, dst) = (uint64_t)str_rfind(RS(regs, a), RS(regs, b)); break;
                    case TPLX_OP_SIN: R(regs, dst) = str_find(RS(regs, b), RS(regs, a)) >= 0; break;
                    case TPLX_OP_SEQ: R(regs, dst) = (uint64_t)(str_eq(RS(regs, a), RS(regs, b)) != (bool)(flags & 1)); break;
                    case TPLX_OP_STRUTH: R(regs, dst) = ((uint32_t)R(regs, a + 1)) != 0; break;
                    case TPLX_OP_SSTARTS: {
                        StrV s = RS(regs, a), p = RS(regs, b);
                        bool r = p.len <= s.len;
                        for (uint32_t i = 0; r && i < p.len; ++i) r = sch(s, i) == sch(p, i);
                        R(regs, dst) = r;
                        break;
                    }
                    case TPLX_OP_SENDS: {
                        StrV s = RS(regs, a), p = RS(regs, b);
                        bool r = p.len <= s.len;
                        for (uint32_t i = 0; r && i < p.len; ++i) r = sch(s, s.len - p.len + i) == sch(p, i);
                        R(regs, dst) = r;
                        break;
                    }
                    case TPLX_OP_SSLICE: {
                        StrV s = RS(regs, a);
                        int64_t len = s.len;
                        int64_t st = (flags & TPLX_SL_HAS_START) ? slice_index((int64_t)R(regs, b), len) : 0;
                        int64_t en = (flags & TPLX_SL_HAS_END) ? slice_index((int64_t)R(regs, c), len) : len;
                        if (st < en) WS(regs, dst, s.p + st, (uint32_t)(en - st), s.flags);
                        else WS(regs, dst, s.p, 0, 0);
                        break;
                    }
                    case TPLX_OP_SINDEX: {
                        StrV s = RS(regs, a);
                        int64_t idx = (int64_t)R(regs, b);
                        if (idx < 0) idx += s.len;
                        if (idx < 0 || idx >= (int64_t)s.len) { raise_exc(t, TPLX_EC_INDEXERROR, opidx); break; }
                        WS(regs, dst, s.p + idx, 1, s.flags);
                        break;
                    }
                    case TPLX_OP_SLOWER: {
                        StrV s = RS(regs, a);
                        WS(regs, dst, s.p, s.len, TPLX_SF_LOWER);
                        break;
                    }
                    case TPLX_OP_SUPPER: {
                        StrV s = RS(regs, a);
                        WS(regs, dst, s.p, s.len, TPLX_SF_UPPER);
                        break;
                    }
                    case TPLX_OP_SSTRIP: {
                        StrV s = RS(regs, a);
                        uint32_t i = 0, e = s.len;
                        if (flags & 1) while (i < e && is_pyspace(s.p[i])) ++i;
                        if (flags & 2) while (e > i && is_pyspace(s.p[e - 1])) --e;
                        WS(regs, dst, s.p + i, e - i, s.flags);
                        break;
                    }
                    case TPLX_OP_SCONCAT: {
                        // empty operand returns the other side unchanged (BlockGeneratorVisitor.cc:381-436)
                        StrV x = RS(regs, a), y = RS(regs, b);
                        if (x.len == 0) { WS(regs, dst, y.p, y.len, y.flags); break; }
                        if (y.len == 0) { WS(regs, dst, x.p, x.len, x.flags); break; }
                        uint8_t *o = scratch_alloc(t, x.len + y.len, opidx);
                        if (!o) break;
                        for (uint32_t i = 0; i < x.len; ++i) o[i] = sch(x, i);
                        for (uint32_t i = 0; i < y.len; ++i) o[x.len + i] = sch(y, i);
                        WS(regs, dst, o, x.len + y.len, 0);
                        break;
                    }
                    case TPLX_OP_SREPLACE: {
                        // runtime/src/Runtime.cc:401-540
                        StrV s = RS(regs, a), f = RS(regs, b), r = RS(regs, c);
                        if (s.len == 0 || (f.len == 0 && r.len == 0)) { WS(regs, dst, s.p, s.len, s.flags); break; }
                        if (f.len == 0) {
                            uint32_t n = (r.len + 1) * s.len + r.len;
                            uint8_t *o = scratch_alloc(t, n, opidx);
                            if (!o) break;
                            uint32_t pos = 0;
                            for (uint32_t i = 0; i < s.len; ++i) {
                                for (uint32_t j = 0; j < r.len; ++j) o[pos++] = sch(r, j);
                                o[pos++] = sch(s, i);
                            }
                            for (uint32_t j = 0; j < r.len; ++j) o[pos++] = sch(r, j);
                            WS(regs, dst, o, pos, 0);
                            break;
                        }
                        // count pass
                        uint32_t count = 0, i = 0;
                        while (i + f.len <= s.len) {
                            uint32_t j = 0;
                            while (j < f.len && sch(s, i + j) == sch(f, j)) ++j;
                            if (j == f.len) { ++count; i += f.len; } else ++i;
                        }
                        uint32_t n = s.len + count * r.len - count * f.len;
                        uint8_t *o = scratch_alloc(t, n ? n : 1, opidx);
                        if (!o) break;
                        uint32_t pos = 0;
                        i = 0;
                        while (i < s.len) {
                            bool m = false;
                            if (i + f.len <= s.len) {
                                uint32_t j = 0;
                                while (j < f.len && sch(s, i + j) == sch(f, j)) ++j;
                                m = (j == f.len);
                            }
                            if (m) {
                                for (uint32_t j = 0; j < r.len; ++j) o[pos++] = sch(r, j);
                                i += f.len;
                            } else o[pos++] = sch(s, i++);
                        }
                        WS(regs, dst, o, pos, 0);
                        break;
                    }
                    case TPLX_OP_SFMTD: {
                        // snprintf("%[0]<w>d", (int)v): C %d consumes an int (BlockGeneratorVisitor.cc:675-775)
                        int32_t v = (int32_t)(int64_t)R(regs, a);
                        uint32_t width = (uint32_t)imm;
                        uint32_t mag = v < 0 ? (uint32_t)0 - (uint32_t)v : (uint32_t)v;
                        uint8_t dig[12];
                        uint32_t nd = 0;
                        do { dig[nd++] = (uint8_t)('0' + mag % 10); mag /= 10; } while (mag);
                        uint32_t body = nd + (v < 0 ? 1 : 0);
                        uint32_t total = body > width ? body : width;
                        uint8_t *o = scratch_alloc(t, total, opidx);
                        if (!o) break;
                        uint32_t pos = 0, padn = total - body;
                        if (flags & 1) {
                            if (v < 0) o[pos++] = '-';
                            for (uint32_t i = 0; i < padn; ++i) o[pos++] = '0';
                        } else {
                            for (uint32_t i = 0; i < padn; ++i) o[pos++] = ' ';
                            if (v < 0) o[pos++] = '-';
                        }
                        while (nd) o[pos++] = dig[--nd];
                        WS(regs, dst, o, total, 0);
                        break;
                    }
                    case TPLX_OP_I2S: {
                        int64_t v = (int64_t)R(regs, a);
                        uint64_t mag = v < 0 ? (uint64_t)0 - (uint64_t)v : (uint64_t)v;
                        uint8_t dig[20];
                        uint32_t nd = 0;
                        do { dig[nd++] = (uint8_t)('0' + mag % 10); mag /= 10; } while (mag);
                        uint32_t total = nd + (v < 0 ? 1 : 0);
                        uint8_t *o = scratch_alloc(t, total, opidx);
                        if (!o) break;
                        uint32_t pos = 0;
                        if (v < 0) o[pos++] = '-';
                        while (nd) o[pos++] = dig[--nd];
                        WS(regs, dst, o, total, 0);
                        break;
                    }
                    case TPLX_OP_S2I: {
                        // fast_atoi64 (runtime/src/Runtime.cc:319-341 + utils/src/StringUtils.cc:22-63)
                        StrV s = RS(regs, a);
                        uint32_t i = 0, e = s.len;
                        while (i < e && is_pyspace(s.p[i])) ++i;
                        if (e > i) {
                            uint32_t e2 = e - 1;
                            while (e2 > i && is_pyspace(s.p[e2])) --e2;
                            e = e2 + 1;
                        }
                        if (i == e) { raise_exc(t, TPLX_EC_VALUEERROR, opidx); break; }
                        bool neg = false;
                        if (s.p[i] == '-') { neg = true; ++i; }
                        uint64_t x = 0;
                        while (i < s.len) {
                            uint8_t d = (uint8_t)(s.p[i] - '0');
                            if (d > 9) break;
                            x = x * 10 + d;
                            ++i;
                        }
                        if (i != e) { raise_exc(t, TPLX_EC_VALUEERROR, opidx); break; }
                        R(regs, dst) = neg ? (uint64_t)0 - x : x;
                        break;
                    }
                    case TPLX_OP_FILTER:
                        if (R(regs, a) == 0) t.alive = false;
                        break;
                    case TPLX_OP_RAISE: raise_exc(t, (uint32_t)imm, opidx); break;
                    default: break;
                }
            }
        }
    }
};

}  // namespace tplx
