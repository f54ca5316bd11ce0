// vecvm.cuh — K1v: the row kernel for stages made of fixed-width columns only (BASELINE config 0 / C1 and every numeric
// map / filter / project pipeline): the closed-form fast path the reference gets from specialising its block loop per stage
// (tuplex/core/src/physical/TuplexSourceTaskBuilder.cc:104-215 + PipelineBuilder.cc:565-700 for fixed-width rows).
//
// Same op program, same per-row semantics as vm.cuh (every op is the same single integer / IEEE operation), evaluated
// VECTOR-AT-A-TIME: a thread owns V = 2J rows of a tile; one instruction dispatch (uniform fetch + decode) is followed by 2J
// row operations, so the interpretive overhead per row shrinks by 2J. B200 mapping:
//   * inputs: coalesced 128-bit global loads (LDG.E.128: thread t reads rows 2t, 2t+1 of every 512-row slab of the tile);
//   * register file in shared memory as regs[slot][local row] — one conflict-free LDS.128 / STS.128 moves two rows;
//     because a slot is indexed by the local row it IS the staging array of an output column (no copy before the write);
//   * filter mask -> warp ballots -> bitmaps -> rows_tile_finish (decoupled look-back scan + dense coalesced write), shared
//     with the scalar kernel.
// Strings never enter this kernel (the host selects it only for programs without string values).
#pragma once
#include <stdint.h>
#include "kernels.cuh"

namespace tplx {

// internal micro-ops produced by the pre-decoder (never part of the IR): strength-reduced forms with a constant power-of-two
// divisor; floored semantics make both exact for every dividend:  x // 2^k == x >> k (arithmetic),  x % 2^k == x & (2^k - 1)
constexpr uint32_t UOP_ISHR_FLOORDIV = 200;  // imm = k
constexpr uint32_t UOP_IAND_MOD = 201;       // imm = 2^k - 1

__device__ __forceinline__ uint32_t spread16(uint32_t x) {  // bit i -> bit 2i
    x &= 0xFFFFu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}

template <int J>
struct VecVM {
    static constexpr uint32_t V = 2 * J;                 // rows per thread
    static constexpr uint32_t T = V * NT;                // rows per tile
    static constexpr uint32_t SLAB = NT * 16;            // bytes between the two-row groups j and j+1 of one slot
    static constexpr uint32_t SLOT_BYTES = T * 8;        // one slot = T values indexed by local row

    struct State {
        uint32_t alive;  // bit v: row v of this thread is still in the pipeline
        uint32_t exc;    // bit v: row v raised
    };
    // local row of (j, b) for this thread: lr = j * 2 NT + 2 tid + b
    static __device__ __forceinline__ uint32_t lrow(uint32_t j, uint32_t b) { return j * 2 * NT + 2 * threadIdx.x + b; }

    static __device__ __forceinline__ ulonglong2 ld2(const uint8_t *rb, uint32_t off, uint32_t j) {
        return *reinterpret_cast<const ulonglong2 *>(rb + off + j * SLAB);
    }
    static __device__ __forceinline__ void st2(uint8_t *rb, uint32_t off, uint32_t j, ulonglong2 v) {
        *reinterpret_cast<ulonglong2 *>(rb + off + j * SLAB) = v;
    }
    static __device__ __forceinline__ void raise_row(State &st, uint32_t v, uint32_t code, uint32_t opidx, uint32_t *exc_stage, uint32_t lr) {
        st.alive &= ~(1u << v);
        st.exc |= 1u << v;
        exc_stage[lr] = code | (opidx << 16);
    }

    // rb = regs base + tid * 16. tile_row0 = first input row of the tile. Rows >= n_rows are inactive from the start.
    static __device__ void run(const DInstr *__restrict__ prog, uint32_t n_instr, uint8_t *__restrict__ rb, const ColIn *__restrict__ cols,
                               uint64_t tile_row0, uint64_t n_rows, State &st, uint32_t *__restrict__ exc_stage) {
        for (uint32_t pc = 0; pc < n_instr; ++pc) {
            const uint4 w0 = *reinterpret_cast<const uint4 *>(&prog[pc]);
            const uint2 w1 = *reinterpret_cast<const uint2 *>(&prog[pc].c);
            const uint32_t op = w0.x & 0xFF, flags = (w0.x >> 8) & 0xFF, opidx = w0.x >> 16;
            const uint32_t dst = w0.y, a = w0.z, b = w0.w, c = w1.x, guard = w1.y;
            const uint64_t imm = (uint64_t)prog[pc].imm, imm2 = (uint64_t)prog[pc].imm2;
            uint32_t act = st.alive;  // rows this instruction executes for
            if (guard != NOOFF) {
#pragma unroll
                for (uint32_t j = 0; j < J; ++j) {
                    const ulonglong2 g = ld2(rb, guard, j);
                    if (g.x == 0) act &= ~(1u << (2 * j));
                    if (g.y == 0) act &= ~(2u << (2 * j));
                }
            }
            if (!__any_sync(0xFFFFFFFFu, act != 0)) continue;
            const bool guarded = guard != NOOFF;
            // operands: constants ride in the immediates (a <- imm2, b <- imm, c <- imm2)
#define LDA(j) ((flags & TPLX_F_A_CONST) ? make_ulonglong2(imm2, imm2) : ld2(rb, a, j))
#define LDB(j) ((flags & TPLX_F_B_CONST) ? make_ulonglong2(imm, imm) : ld2(rb, b, j))
            // store: unguarded ops write both rows unconditionally (a row that is not alive never reaches an output);
            // guarded ops must leave the destination of rows outside the guard untouched (phi of an if-converted branch)
#define STD(j, R0, R1)                                                                     \
    do {                                                                                   \
        ulonglong2 _r = make_ulonglong2((R0), (R1));                                       \
        if (guarded) {                                                                     \
            const ulonglong2 _o = ld2(rb, dst, j);                                         \
            if (!((act >> (2 * (j))) & 1u)) _r.x = _o.x;                                   \
            if (!((act >> (2 * (j) + 1)) & 1u)) _r.y = _o.y;                               \
        }                                                                                  \
        st2(rb, dst, j, _r);                                                               \
    } while (0)
#define BIN(EXPR)                                                                          \
    _Pragma("unroll") for (uint32_t j = 0; j < J; ++j) {                                   \
        const ulonglong2 A = LDA(j), B = LDB(j);                                           \
        uint64_t r0, r1;                                                                   \
        { const uint64_t x = A.x, y = B.x; r0 = (EXPR); }                                  \
        { const uint64_t x = A.y, y = B.y; r1 = (EXPR); }                                  \
        STD(j, r0, r1);                                                                    \
    }                                                                                      \
    break
#define UNA(EXPR)                                                                          \
    _Pragma("unroll") for (uint32_t j = 0; j < J; ++j) {                                   \
        const ulonglong2 A = LDA(j);                                                       \
        uint64_t r0, r1;                                                                   \
        { const uint64_t x = A.x; r0 = (EXPR); }                                           \
        { const uint64_t x = A.y; r1 = (EXPR); }                                           \
        STD(j, r0, r1);                                                                    \
    }                                                                                      \
    break
#define F(x) __longlong_as_double((long long)(x))
#define U(d) ((uint64_t)__double_as_longlong(d))
            // ops that can raise: evaluated row by row for the active rows only
#define RAISING(...)                                                                       \
    _Pragma("unroll") for (uint32_t v = 0; v < V; ++v) {                                   \
        if (!((act >> v) & 1u)) continue;                                                  \
        const uint32_t off8 = (v >> 1) * SLAB + (v & 1) * 8;                               \
        const uint64_t x = (flags & TPLX_F_A_CONST) ? imm2 : *reinterpret_cast<const uint64_t *>(rb + a + off8); \
        const uint64_t y = (flags & TPLX_F_B_CONST) ? imm : *reinterpret_cast<const uint64_t *>(rb + b + off8);  \
        uint64_t r;                                                                        \
        bool bad = false;                                                                  \
        __VA_ARGS__;                                                                       \
        if (bad) raise_row(st, v, TPLX_EC_ZERODIVISIONERROR, opidx, exc_stage, lrow(v >> 1, v & 1)); \
        else *reinterpret_cast<uint64_t *>(rb + dst + off8) = r;                           \
    }                                                                                      \
    break
            switch (op) {
                case TPLX_OP_LDCOL: {
                    const ColIn &ci = cols[imm];
                    const uint64_t *src = reinterpret_cast<const uint64_t *>(ci.data);
#pragma unroll
                    for (uint32_t j = 0; j < J; ++j) {
                        const uint64_t row = tile_row0 + lrow(j, 0);
                        ulonglong2 v = make_ulonglong2(0, 0);
                        if (row + 1 < n_rows) v = *reinterpret_cast<const ulonglong2 *>(src + row);  // 16-byte aligned: row is even, base is
                        else if (row < n_rows) v.x = src[row];
                        STD(j, v.x, v.y);
                    }
                    break;
                }
                case TPLX_OP_LDI:
#pragma unroll
                    for (uint32_t j = 0; j < J; ++j) STD(j, imm, imm);
                    break;
                case TPLX_OP_LDROW:
#pragma unroll
                    for (uint32_t j = 0; j < J; ++j) {
                        const uint64_t row = tile_row0 + lrow(j, 0);
                        STD(j, row, row + 1);
                    }
                    break;
                case TPLX_OP_MOV: UNA(x);
                case TPLX_OP_SEL:
#pragma unroll
                    for (uint32_t j = 0; j < J; ++j) {
                        const ulonglong2 A = LDA(j), B = LDB(j), Cn = ld2(rb, c, j);
                        STD(j, Cn.x ? A.x : B.x, Cn.y ? A.y : B.y);
                    }
                    break;
                case TPLX_OP_IADD: BIN(x + y);
                case TPLX_OP_ISUB: BIN(x - y);
                case TPLX_OP_IMUL: BIN(x * y);
                case TPLX_OP_INEG: UNA((uint64_t)0 - x);
                case TPLX_OP_IAND: BIN(x & y);
                case TPLX_OP_IOR: BIN(x | y);
                case TPLX_OP_IXOR: BIN(x ^ y);
                case TPLX_OP_ISHL: BIN(x << (y & 63));
                case TPLX_OP_ISHR: BIN((uint64_t)((int64_t)x >> (y & 63)));
                case UOP_ISHR_FLOORDIV: UNA((uint64_t)((int64_t)x >> (imm & 63)));
                case UOP_IAND_MOD: UNA(x & imm);
                case TPLX_OP_IABS: UNA((int64_t)x < 0 ? (uint64_t)0 - x : x);
                case TPLX_OP_FADD: BIN(U(__dadd_rn(F(x), F(y))));
                case TPLX_OP_FSUB: BIN(U(__dsub_rn(F(x), F(y))));
                case TPLX_OP_FMUL: BIN(U(__dmul_rn(F(x), F(y))));
                case TPLX_OP_FNEG: UNA(x ^ 0x8000000000000000ull);
                case TPLX_OP_FABS: UNA(x & 0x7FFFFFFFFFFFFFFFull);
                case TPLX_OP_I2F: UNA(U((double)(int64_t)x));
                case TPLX_OP_F2I: UNA((uint64_t)(int64_t)F(x));
                case TPLX_OP_BAND: BIN((uint64_t)((x != 0) & (y != 0)));
                case TPLX_OP_BOR: BIN((uint64_t)((x != 0) | (y != 0)));
                case TPLX_OP_BNOT: UNA((uint64_t)(x == 0));
                case TPLX_OP_ICMP:
                    switch (flags & 7) {
                        case TPLX_CMP_EQ: BIN((uint64_t)(x == y));
                        case TPLX_CMP_NE: BIN((uint64_t)(x != y));
                        case TPLX_CMP_LT: BIN((uint64_t)((int64_t)x < (int64_t)y));
                        case TPLX_CMP_LE: BIN((uint64_t)((int64_t)x <= (int64_t)y));
                        case TPLX_CMP_GT: BIN((uint64_t)((int64_t)x > (int64_t)y));
                        default: BIN((uint64_t)((int64_t)x >= (int64_t)y));
                    }
                    break;
                case TPLX_OP_FCMP:
                    switch (flags & 7) {  // ordered predicates: false when either side is NaN
                        case TPLX_CMP_EQ: BIN((uint64_t)(F(x) == F(y)));
                        case TPLX_CMP_NE: BIN((uint64_t)((F(x) < F(y)) || (F(x) > F(y))));
                        case TPLX_CMP_LT: BIN((uint64_t)(F(x) < F(y)));
                        case TPLX_CMP_LE: BIN((uint64_t)(F(x) <= F(y)));
                        case TPLX_CMP_GT: BIN((uint64_t)(F(x) > F(y)));
                        default: BIN((uint64_t)(F(x) >= F(y)));
                    }
                    break;
                case TPLX_OP_IFLOORDIV: RAISING({ if ((int64_t)y == 0) bad = true; else r = (uint64_t)floordiv_i64((int64_t)x, (int64_t)y); });
                case TPLX_OP_IMOD: RAISING({ if ((int64_t)y == 0) bad = true; else r = (uint64_t)floormod_i64((int64_t)x, (int64_t)y); });
                case TPLX_OP_FDIV: RAISING({ if (F(y) == 0.0) bad = true; else r = U(__ddiv_rn(F(x), F(y))); });
                case TPLX_OP_FMOD: RAISING({
                    if (F(y) == 0.0) bad = true;
                    else {
                        double m = fmod(F(x), F(y));  // == LLVM frem, exact
                        if (m != 0.0 && ((m < 0.0) != (F(y) < 0.0))) m = __dadd_rn(m, F(y));
                        r = U(m);
                    }
                });
                case TPLX_OP_FFLOORDIV: RAISING({
                    const int64_t xi = (int64_t)F(x), yi = (int64_t)F(y);
                    if (F(y) == 0.0 || yi == 0) bad = true;
                    else r = U((double)floordiv_i64(xi, yi));
                });
                case TPLX_OP_FILTER:
#pragma unroll
                    for (uint32_t j = 0; j < J; ++j) {
                        const ulonglong2 A = ld2(rb, a, j);
                        if (A.x == 0) st.alive &= ~(act & (1u << (2 * j)));
                        if (A.y == 0) st.alive &= ~(act & (2u << (2 * j)));
                    }
                    break;
                case TPLX_OP_RAISE:
#pragma unroll
                    for (uint32_t v = 0; v < V; ++v)
                        if ((act >> v) & 1u) raise_row(st, v, (uint32_t)imm, opidx, exc_stage, lrow(v >> 1, v & 1));
                    break;
                default: break;
            }
#undef LDA
#undef LDB
#undef STD
#undef BIN
#undef UNA
#undef F
#undef U
#undef RAISING
        }
    }
};

// ---- tile tail of the vector kernel: block-wide decoupled look-back + dense write (fixed-width outputs only) -------------------
// The scalar kernel's look-back (rows_tile_finish: one warp, a 32-tile window per round, K values behind a status flag and a fence)
// is fine for heavy tiles; a fixed-width tile is evaluated in ~1-2 us, and then the speed at which inclusive prefixes propagate
// (32 tiles per L2 round trip) bounds the whole kernel (measured: 59 G rows/s on C1). Here a tile's scan state is ONE 64-bit word
//   [status:2 | rows kept:31 | exception rows:31]   status 1 = this tile's counts, 2 = inclusive prefix
// written with a single relaxed store (no fence: status and values travel together), and ALL 256 threads look back at once: thread t
// polls tile - 1 - t, the block adds the counts up to the nearest inclusive word — 256 tiles per round trip.
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void vec_tile_finish(const KParams &P, uint32_t tile, uint64_t base, uint32_t T, uint32_t W, uint8_t *s_regs,
                                                uint32_t *keep_bits, uint32_t *exc_bits, uint32_t *keep_pre, uint32_t *exc_pre,
                                                uint32_t *exc_stage, uint64_t *s_vals, uint64_t *s_scr) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // word prefixes of the two bitmaps (warp 0)
    if (warp == 0) {
        uint32_t ck = 0, ce = 0;
        for (uint32_t w0 = 0; w0 < W; w0 += 32) {
            const uint32_t w = w0 + lane;
            const uint32_t pk = w < W ? __popc(keep_bits[w]) : 0, pe = w < W ? __popc(exc_bits[w]) : 0;
            uint32_t ik = pk, ie = pe;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, ik, o), b = __shfl_up_sync(0xFFFFFFFFu, ie, o);
                if (lane >= (uint32_t)o) { ik += a; ie += b; }
            }
            if (w < W) { keep_pre[w] = ck + ik - pk; exc_pre[w] = ce + ie - pe; }
            ck += __shfl_sync(0xFFFFFFFFu, ik, 31);
            ce += __shfl_sync(0xFFFFFFFFu, ie, 31);
        }
        if (lane == 0) {
            s_vals[0] = ck;
            s_vals[1] = ce;
            // publish this tile's counts (tile 0: they are its inclusive prefix)
            st_cg_u64(P.tile_state + tile, ((uint64_t)(tile == 0 ? 2u : 1u) << 62) | ((uint64_t)ck << 31) | (uint64_t)ce);
        }
    }
    __syncthreads();
    const uint32_t n_keep = (uint32_t)s_vals[0], n_exc = (uint32_t)s_vals[1];
    // ---- look-back: 256 predecessors per round ----
    uint64_t pre_keep = 0, pre_exc = 0;
    if (tile > 0) {
        int64_t p = (int64_t)tile - 1;
        while (true) {
            const int64_t q = p - (int64_t)tid;
            uint64_t wd = (uint64_t)2 << 62;  // tiles before the first count as an inclusive prefix of zero
            if (q >= 0) do { wd = ld_relaxed_u64(P.tile_state + q); } while ((wd >> 62) == 0);
            const bool incl = (wd >> 62) == 2;
            const uint32_t im = __ballot_sync(0xFFFFFFFFu, incl);
            const uint32_t first = im ? (uint32_t)(__ffs(im) - 1) : 32u;  // nearest inclusive word inside this warp's 32 tiles
            uint64_t k = lane <= first ? (wd >> 31) & 0x7FFFFFFFull : 0, e = lane <= first ? wd & 0x7FFFFFFFull : 0;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                k += __shfl_xor_sync(0xFFFFFFFFu, k, o);
                e += __shfl_xor_sync(0xFFFFFFFFu, e, o);
            }
            if (lane == 0) {
                s_scr[warp * 2] = (k << 32) | e;      // both < 2^31 * 32: fit 32 bits each? counts are < 2^31 in total, so yes
                s_scr[warp * 2 + 1] = im ? 1 : 0;
            }
            __syncthreads();
            bool done = false;
            for (uint32_t w = 0; w < NT / 32 && !done; ++w) {  // warps cover tile-1-32w .. : nearest first
                pre_keep += s_scr[w * 2] >> 32;
                pre_exc += s_scr[w * 2] & 0xFFFFFFFFull;
                done = s_scr[w * 2 + 1] != 0;
            }
            __syncthreads();
            if (done) break;
            p -= NT;
        }
        if (tid == 0) st_cg_u64(P.tile_state + tile, ((uint64_t)2 << 62) | ((pre_keep + n_keep) << 31) | (pre_exc + n_exc));
    }
    if (tid == 0 && tile == P.n_tiles - 1) {
        P.totals[0] = pre_keep + n_keep;
        P.totals[1] = pre_exc + n_exc;
    }
    // ---- write: fixed-width outputs straight from the register file (a slot is indexed by the local row) ----
    if (n_keep) {
        for (uint32_t c = 0; c < P.n_out; ++c) {
            const OutCol &oc = P.out[c];
            const uint64_t *st = reinterpret_cast<const uint64_t *>(s_regs + oc.stage_off);
            for (uint32_t lr = tid; lr < T; lr += NT)
                if (bit_test(keep_bits, lr)) oc.data[pre_keep + bit_rank(keep_bits, keep_pre, lr)] = st[lr];
        }
    }
    if (n_exc) {
        if (pre_exc + n_exc <= P.cap_exc) {
            for (uint32_t lr = tid; lr < T; lr += NT) {
                if (!bit_test(exc_bits, lr)) continue;
                const uint32_t ke = bit_rank(exc_bits, exc_pre, lr), kk = bit_rank(keep_bits, keep_pre, lr);
                tplx_exception_rec rec;
                rec.row = (int64_t)(base + lr);
                rec.row_no = P.first_row_no + (int64_t)(pre_keep + pre_exc + kk + ke);  // rows written + exceptions so far (TransformTask.cc:764,885)
                const uint32_t es = exc_stage[lr];
                rec.code = es & 0xFFFF;
                rec.op_id = P.opids[es >> 16];
                P.exc[pre_exc + ke] = rec;
            }
        } else if (tid == 0) atomicOr(&P.counters[1], 4u);
    }
}

// K1v kernel: persistent CTAs, ticketed tiles of T = 2J * 256 rows, VecVM evaluation, shared tail (rows_tile_finish).
// Shared memory: prog | cols | regs (n_slots x T x 8 B; output columns are staged in place) | misc (bitmaps, scan scratch).
template <int J>
__global__ void __launch_bounds__(NT) stage_rows_vec_kernel(const __grid_constant__ KParams P) {  // parameters in the constant bank
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr uint32_t T = VecVM<J>::T, W = T / 32;
    const uint32_t K = P.K;

    DInstr *s_prog = reinterpret_cast<DInstr *>(smem);
    ColIn *s_cols = reinterpret_cast<ColIn *>(smem + P.smem_cols_off);
    uint8_t *s_regs = smem + P.smem_regs_off;
    uint32_t *keep_bits = reinterpret_cast<uint32_t *>(smem + P.smem_misc_off);
    uint32_t *exc_bits = keep_bits + W;
    uint32_t *keep_pre = exc_bits + W;
    uint32_t *exc_pre = keep_pre + W + 1;
    uint32_t *exc_stage = exc_pre + W + 1;
    uint64_t *s_vals = reinterpret_cast<uint64_t *>(exc_stage + T);
    uint64_t *s_excl = s_vals + MAX_SCAN;
    uint64_t *s_warp = s_excl + MAX_SCAN;
    uint32_t *s_ctl = reinterpret_cast<uint32_t *>(s_warp + NT / 32 + 1);

    for (uint32_t i = tid; i < P.n_instr * (sizeof(DInstr) / 16); i += NT)
        reinterpret_cast<uint4 *>(s_prog)[i] = reinterpret_cast<const uint4 *>(P.prog)[i];
    for (uint32_t i = tid; i < P.n_in * (sizeof(ColIn) / 8); i += NT)
        reinterpret_cast<uint64_t *>(s_cols)[i] = reinterpret_cast<const uint64_t *>(P.in)[i];
    const uint32_t state_stride = 1 + 2 * K;

    while (true) {
        __syncthreads();
        if (tid == 0) s_ctl[0] = atomicAdd(&P.counters[0], 1u);
        __syncthreads();
        const uint32_t tile = s_ctl[0];
        if (tile >= P.n_tiles) break;
        const uint64_t base = (uint64_t)tile * T;

        typename VecVM<J>::State st;
        st.exc = 0;
        if (base + T <= P.n_rows) st.alive = (1u << (2 * J)) - 1u;  // full tile (uniform test): every row starts alive
        else {
            st.alive = 0;
#pragma unroll
            for (uint32_t v = 0; v < 2 * J; ++v)
                if (base + VecVM<J>::lrow(v >> 1, v & 1) < P.n_rows) st.alive |= 1u << v;
        }
        VecVM<J>::run(s_prog, P.n_instr, s_regs + tid * 16, s_cols, base, P.n_rows, st, exc_stage);
        // bitmaps indexed by local row: thread t holds rows (2t, 2t+1) of slab j, so bit i of a bitmap word comes from lane i / 2
        // (+ 16 for the word's upper half of the warp), even / odd row by the parity of i: one extra ballot interleaves the two masks
        const bool any_exc = __any_sync(0xFFFFFFFFu, st.exc != 0);
#pragma unroll
        for (uint32_t j = 0; j < J; ++j) {
            const uint32_t ke = __ballot_sync(0xFFFFFFFFu, (st.alive >> (2 * j)) & 1u), ko = __ballot_sync(0xFFFFFFFFu, (st.alive >> (2 * j + 1)) & 1u);
            const uint32_t pick = (lane & 1) ? ko : ke, sh = lane >> 1;
            const uint32_t w0 = __ballot_sync(0xFFFFFFFFu, (pick >> sh) & 1u), w1 = __ballot_sync(0xFFFFFFFFu, (pick >> (16 + sh)) & 1u);
            uint32_t e0 = 0, e1 = 0;
            if (any_exc) {
                const uint32_t ee = __ballot_sync(0xFFFFFFFFu, (st.exc >> (2 * j)) & 1u), eo = __ballot_sync(0xFFFFFFFFu, (st.exc >> (2 * j + 1)) & 1u);
                const uint32_t pe = (lane & 1) ? eo : ee;
                e0 = __ballot_sync(0xFFFFFFFFu, (pe >> sh) & 1u);
                e1 = __ballot_sync(0xFFFFFFFFu, (pe >> (16 + sh)) & 1u);
            }
            if (lane == 0) {
                const uint32_t w = j * (2 * NT / 32) + 2 * warp;
                keep_bits[w] = w0;
                keep_bits[w + 1] = w1;
                exc_bits[w] = e0;
                exc_bits[w + 1] = e1;
            }
        }
        __syncthreads();
        vec_tile_finish(P, tile, base, T, W, s_regs, keep_bits, exc_bits, keep_pre, exc_pre, exc_stage, s_vals, s_excl);
    }
}

}  // namespace tplx
