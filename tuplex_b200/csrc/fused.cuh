// fused.cuh — K3f: streaming scan-aggregate for the closed form described by tplx_fused_header
// (range predicates on fixed-width columns -> sums of const | col | col*col; the TPC-H Q6 class).
//
// Same work as K3 through the VM (reference: the JIT'd filter/aggregate loop,
// tuplex/core/src/physical/PipelineBuilder.cc:615-700,2525-2608), without interpretation: one thread folds
// the rows t, t+256, ... of a 4096-row tile exactly like stage_agg_kernel does, with the same per-row IEEE
// operations (__dmul_rn, __dadd_rn) and the same warp / tile / finalize tree — so its result is bit-identical
// to the VM path and to the oracle's acc_tree. Loads are plain coalesced 8-byte column reads, all predicates are
// evaluated branch-free so that every load of a row is independent (maximum bytes in flight per thread).
#pragma once
#include <stdint.h>
#include "kernels.cuh"

namespace tplx {

struct FusedParams {
    uint32_t n_preds, n_terms;
    tplx_fused_pred preds[TPLX_MAX_FUSED_PREDS];
    tplx_fused_term terms[TPLX_MAX_ACCS];
};

__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// Predicates are normalised on the host (launch_fused) to an inclusive range test on an ordered 64-bit key:
//   i64 column : key = value
//   f64 column : key = order-preserving transform of the IEEE bits (-0.0 canonicalised to +0.0; NaN keys fall
//                outside [key(-inf), key(+inf)], so ordered comparisons with NaN are false like FCMP_O*)
//   cast       : i64 column converted with sitofp, then as f64
// strict bounds become inclusive by stepping one key, missing bounds become the extreme keys.
__host__ __device__ __forceinline__ int64_t f64_key(uint64_t bits) {
    if (bits == 0x8000000000000000ull) bits = 0;  // -0.0 == +0.0
    return (int64_t)(bits ^ ((uint64_t)((int64_t)bits >> 63) & 0x7FFFFFFFFFFFFFFFull));
}
__device__ __forceinline__ bool pred_pass(const tplx_fused_pred &p, uint64_t raw) {
    // p.flags here: 0 = i64 key, 1 = f64 key, 2 = cast then f64 key; p.lo / p.hi = inclusive key bounds
    int64_t k = (int64_t)raw;
    if (p.flags == 2) k = f64_key((uint64_t)__double_as_longlong((double)(int64_t)raw));
    else if (p.flags == 1) k = f64_key(raw);
    return (k >= p.lo) & (k <= p.hi);
}

template <int NP, int NTM>
__global__ void __launch_bounds__(NT) fused_scan_agg_kernel(const KParams *__restrict__ Pg, const FusedParams *__restrict__ Fg) {
    __shared__ uint64_t s_wacc[(NT / 32) * NTM];
    __shared__ FusedParams F;
    __shared__ const uint64_t *s_col[TPLX_MAX_COLS];
    const KParams &P = *Pg;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t i = tid; i < sizeof(FusedParams) / 4; i += NT) reinterpret_cast<uint32_t *>(&F)[i] = reinterpret_cast<const uint32_t *>(Fg)[i];
    for (uint32_t i = tid; i < P.n_in; i += NT) s_col[i] = reinterpret_cast<const uint64_t *>(P.in[i].data);
    __syncthreads();
    constexpr uint32_t R = 16, T = R * NT;

    for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        const uint64_t base = (uint64_t)tile * T;
        uint64_t acc[NTM];
#pragma unroll
        for (int k = 0; k < NTM; ++k) acc[k] = acc_identity(F.terms[k].kind);
#pragma unroll 4
        for (uint32_t s = 0; s < R; ++s) {
            const uint64_t row = base + (uint64_t)s * NT + tid;
            const bool valid = row < P.n_rows;
            const uint64_t r = valid ? row : 0;  // clamp: loads stay in bounds, result discarded
            bool pass = valid;
            uint64_t pv[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) pv[p] = ld_stream_u64(s_col[F.preds[p].col] + r);
            uint64_t ta[NTM], tb[NTM];
#pragma unroll
            for (int k = 0; k < NTM; ++k) {
                ta[k] = F.terms[k].op != TPLX_FT_CONST ? ld_stream_u64(s_col[F.terms[k].col_a] + r) : 0;
                tb[k] = F.terms[k].op == TPLX_FT_MUL ? ld_stream_u64(s_col[F.terms[k].col_b] + r) : 0;
            }
#pragma unroll
            for (int p = 0; p < NP; ++p) pass = pass & pred_pass(F.preds[p], pv[p]);
            if (pass) {
#pragma unroll
                for (int k = 0; k < NTM; ++k) {
                    const tplx_fused_term &tm = F.terms[k];
                    uint64_t v;
                    if (tm.kind == TPLX_ACC_SUM_F64) {
                        const double a = tm.cast_a ? (double)(int64_t)ta[k] : __longlong_as_double((long long)ta[k]);
                        const double b = tm.cast_b ? (double)(int64_t)tb[k] : __longlong_as_double((long long)tb[k]);
                        const double g = tm.op == TPLX_FT_CONST ? __longlong_as_double((long long)tm.imm) : (tm.op == TPLX_FT_COL ? a : __dmul_rn(a, b));
                        v = (uint64_t)__double_as_longlong(g);
                    } else {
                        v = tm.op == TPLX_FT_CONST ? (uint64_t)tm.imm : (tm.op == TPLX_FT_COL ? ta[k] : ta[k] * tb[k]);
                    }
                    acc[k] = acc_combine(tm.kind, acc[k], v);
                }
            }
        }
        // identical tree to stage_agg_kernel: shfl_down 16..1, warps sequential
#pragma unroll
        for (int k = 0; k < NTM; ++k) {
            uint64_t v = acc[k];
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                uint64_t other = __shfl_down_sync(0xFFFFFFFFu, v, o);
                v = acc_combine(F.terms[k].kind, v, other);
            }
            if (lane == 0) s_wacc[warp * NTM + k] = v;
        }
        __syncthreads();
        if (tid < NTM && tid < P.n_accs) {
            uint64_t v = s_wacc[tid];
            for (uint32_t w = 1; w < NT / 32; ++w) v = acc_combine(F.terms[tid].kind, v, s_wacc[w * NTM + tid]);
            P.tile_partials[(size_t)tile * P.n_accs + tid] = v;
        }
        __syncthreads();
    }
}

}  // namespace tplx
