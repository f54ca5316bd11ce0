// option.cuh — Option[T] columns around the row kernels (the NULL-aware normal case).
//
// The reference keeps, for a schema with Option fields, a per-row bitmap in front of the row's slots
// (Serializer.cc:1041-1059, calcBitmapSize :29-41) and the generated code tests a field's bit before it uses the value
// (FlattenedTuple::getIsNull; `x is None` / `x == None`: BlockGeneratorVisitor.cc:1030-1150). Here a column block carries one
// validity bitmap per Option column (tplx_column.valid). The op program reads it through an "is None" COMPANION input column of
// type bool (stage descriptor: in_types entry TPLX_T_NULLOF | column) that these kernels expand from the bitmap right before the
// stage runs — the row kernels themselves (K1 / K1v / K1r / K1m and their specialised builds) stay free of bitmap addressing and
// vectorise the flag like any other fixed-width column. Option outputs come back the same way: value column + hidden companion,
// packed into the result's validity bitmap (tplx_gpu_result_fetch_validity) by valid_pack_kernel.
#pragma once
#include <stdint.h>

namespace tplx {

// companion[r] = 1 when row r holds None (validity bit clear), else 0
__global__ void __launch_bounds__(256) valid_expand_kernel(const uint32_t *__restrict__ valid, uint64_t n, uint64_t *__restrict__ companion) {
    const uint64_t r = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= n) return;
    companion[r] = ((valid[r >> 5] >> (r & 31)) & 1u) ? 0ull : 1ull;
}

// words[w] bit l = row 32 w + l holds a value (companion == 0); one ballot per warp
__global__ void __launch_bounds__(256) valid_pack_kernel(const uint64_t *__restrict__ companion, uint64_t n, uint32_t *__restrict__ words) {
    const uint64_t r = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    const bool present = r < n && companion[r] == 0;
    const uint32_t w = __ballot_sync(0xFFFFFFFFu, present);
    if ((threadIdx.x & 31) == 0 && (r >> 5) < ((n + 31) >> 5)) words[r >> 5] = w;
}

}  // namespace tplx
