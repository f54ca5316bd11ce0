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

// ---------------------------------------------------------------------------------------------------------
// TMA-staged variant: the distinct input columns of a 1024-row chunk are brought into a shared-memory ring by
// bulk async copies (cp.async.bulk global->shared, completion on an mbarrier; SASS: UBLKCP + SYNCS) issued by
// one elected thread, TMA_STAGES chunks deep, so bytes in flight per SM no longer depend on registers or
// occupancy. Consumers read conflict-free 8-byte words from shared memory. Same row->thread mapping and the same
// reduction tree as the LDG variant (bit-identical results).
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t TMA_CHUNK = 1024;   // rows per stage (4 sub-iterations of 256 threads)
constexpr uint32_t TMA_STAGES = 3;    // x 2 resident CTAs per SM
constexpr uint32_t TMA_MAX_UCOLS = 6;  // distinct columns staged per chunk

struct FusedTmaParams {
    FusedParams f;                       // preds[i].col / terms[k].col_a,col_b index into ucols[]
    uint32_t n_ucols;
    uint32_t ucol[TMA_MAX_UCOLS];        // input column number of each staged column
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

template <int NP, int NTM>
__global__ void __launch_bounds__(NT) fused_scan_agg_tma_kernel(const KParams *__restrict__ Pg, const FusedTmaParams *__restrict__ Fg) {
    extern __shared__ __align__(128) uint8_t dsm[];  // ring: [stage][ucol][TMA_CHUNK] uint64
    __shared__ uint64_t s_wacc[(NT / 32) * NTM];
    __shared__ FusedTmaParams FT;
    __shared__ const uint64_t *s_col[TMA_MAX_UCOLS];
    __shared__ __align__(8) uint64_t full_bar[TMA_STAGES];
    const KParams &P = *Pg;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t i = tid; i < sizeof(FusedTmaParams) / 4; i += NT) reinterpret_cast<uint32_t *>(&FT)[i] = reinterpret_cast<const uint32_t *>(Fg)[i];
    __syncthreads();
    const FusedParams &F = FT.f;
    const uint32_t nuc = FT.n_ucols;
    if (tid < nuc) s_col[tid] = reinterpret_cast<const uint64_t *>(P.in[FT.ucol[tid]].data);
    if (tid == 0) {
        for (uint32_t s = 0; s < TMA_STAGES; ++s) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint64_t *ring = reinterpret_cast<uint64_t *>(dsm);
    constexpr uint32_t R = 16, T = R * NT, CPT = T / TMA_CHUNK;  // chunks per tile
    const uint32_t my_tiles = P.n_tiles > blockIdx.x ? (P.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t n_chunks = my_tiles * CPT;
    auto chunk_base = [&](uint32_t q) { return ((uint64_t)blockIdx.x + (uint64_t)(q / CPT) * gridDim.x) * T + (uint64_t)(q % CPT) * TMA_CHUNK; };
    auto issue = [&](uint32_t q) {  // thread 0 only; only full chunks go through TMA
        const uint64_t b0 = chunk_base(q);
        if (b0 + TMA_CHUNK > P.n_rows) return;
        const uint32_t st = q % TMA_STAGES;
        mbar_expect_tx(&full_bar[st], nuc * TMA_CHUNK * 8);
        for (uint32_t c = 0; c < nuc; ++c) bulk_g2s(ring + ((size_t)st * nuc + c) * TMA_CHUNK, s_col[c] + b0, TMA_CHUNK * 8, &full_bar[st]);
    };
    if (tid == 0)
        for (uint32_t q = 0; q < TMA_STAGES - 1 && q < n_chunks; ++q) issue(q);

    uint64_t acc[NTM];
    for (uint32_t q = 0; q < n_chunks; ++q) {
        const uint32_t st = q % TMA_STAGES;
        const uint64_t b0 = chunk_base(q);
        if ((q % CPT) == 0) {
#pragma unroll
            for (int k = 0; k < NTM; ++k) acc[k] = acc_identity(F.terms[k].kind);
        }
        // keep the ring full: the stage freed by the previous iteration's barrier is refilled now
        if (tid == 0 && q + TMA_STAGES - 1 < n_chunks) issue(q + TMA_STAGES - 1);
        const bool staged = b0 + TMA_CHUNK <= P.n_rows;
        if (staged) mbar_wait(&full_bar[st], (q / TMA_STAGES) & 1);
        const uint64_t *sbase = ring + (size_t)st * nuc * TMA_CHUNK;
#pragma unroll
        for (uint32_t s = 0; s < TMA_CHUNK / NT; ++s) {
            const uint32_t lr = s * NT + tid;
            const uint64_t row = b0 + lr;
            const bool valid = row < P.n_rows;
            bool pass = valid;
            uint64_t pv[NP], ta[NTM], tb[NTM];
            if (staged) {
#pragma unroll
                for (int p = 0; p < NP; ++p) pv[p] = sbase[F.preds[p].col * TMA_CHUNK + lr];
#pragma unroll
                for (int k = 0; k < NTM; ++k) {
                    ta[k] = F.terms[k].op != TPLX_FT_CONST ? sbase[F.terms[k].col_a * TMA_CHUNK + lr] : 0;
                    tb[k] = F.terms[k].op == TPLX_FT_MUL ? sbase[F.terms[k].col_b * TMA_CHUNK + lr] : 0;
                }
            } else {  // ragged tail chunk: plain loads
                const uint64_t r = valid ? row : 0;
#pragma unroll
                for (int p = 0; p < NP; ++p) pv[p] = ld_stream_u64(s_col[F.preds[p].col] + r);
#pragma unroll
                for (int k = 0; k < NTM; ++k) {
                    ta[k] = F.terms[k].op != TPLX_FT_CONST ? ld_stream_u64(s_col[F.terms[k].col_a] + r) : 0;
                    tb[k] = F.terms[k].op == TPLX_FT_MUL ? ld_stream_u64(s_col[F.terms[k].col_b] + r) : 0;
                }
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
        __syncthreads();  // every thread has finished reading this stage: it may be refilled
        if ((q % CPT) == CPT - 1) {
            const uint32_t tile = blockIdx.x + (q / CPT) * gridDim.x;
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
}

}  // namespace tplx
