// gather.cuh — late materialisation of "late" columns for the rows a prefilter stage let through.
//
// When the inputs of tplx_gpu_stage_run_host lie in page-locked host memory, only the columns the prefilter
// reads are copied to HBM up front. The remaining columns are needed for the surviving rows only: these kernels
// read exactly those rows through the mapped host address (many independent, coalesced reads in flight, so the
// transfer is PCIe-bandwidth bound rather than latency bound) and build compact device columns indexed by the
// position in the survivor list. The dense launch then addresses them with the work index instead of the row.
// (The reference has no counterpart: it always deserialises whole rows, TuplexSourceTaskBuilder.cc:184-189.)
#pragma once
#include <stdint.h>
#include "vm.cuh"
#include "csvops.cuh"

namespace tplx {

constexpr uint64_t COL_COMPACT = 0x100;  // ColIn.type flag: column is indexed by work position, not by input row

struct GatherCols {
    uint32_t n_cols;
    uint32_t pad;
    uint8_t type[TPLX_MAX_COLS];
    const void *src_data[TPLX_MAX_COLS];
    const uint32_t *src_off[TPLX_MAX_COLS];
    uint64_t *dst_data[TPLX_MAX_COLS];   // fixed width
    uint32_t *dst_off[TPLX_MAX_COLS];    // strings: n+1 offsets
    uint8_t *dst_bytes[TPLX_MAX_COLS];
    uint64_t *lens[TPLX_MAX_COLS];       // strings: length per survivor (then exclusive scan), n+1 entries
    uint32_t *srcpos[TPLX_MAX_COLS];     // strings: source byte offset per survivor
    // lazy columns of a CSV block (K6): src_data = the CSV text, src_ref[c][row] = start | raw_len << 32 | escaped << 63
    const uint64_t *src_ref[TPLX_MAX_COLS];
    uint8_t quote;
};

// one thread per (survivor, column): fixed-width values and string lengths
__global__ void gather_pass1(const uint64_t *__restrict__ rowlist, uint64_t n, const GatherCols *__restrict__ G) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t row = rowlist[i];
    for (uint32_t c = 0; c < G->n_cols; ++c) {
        if (G->type[c] == TPLX_T_STR && G->src_ref[c]) {
            const uint64_t info = G->src_ref[c][row];
            const uint32_t b = (uint32_t)info, raw = (uint32_t)(info >> 32) & 0x7FFFFFFFu;
            G->lens[c][i] = (info >> 63) ? csv_dequoted_len(reinterpret_cast<const uint8_t *>(G->src_data[c]), b, b + raw, G->quote) : raw;
            G->srcpos[c][i] = b;
        } else if (G->type[c] == TPLX_T_STR) {
            const uint2 o = make_uint2(G->src_off[c][row], G->src_off[c][row + 1]);
            G->lens[c][i] = o.y - o.x;
            G->srcpos[c][i] = o.x;
        } else {
            G->dst_data[c][i] = reinterpret_cast<const uint64_t *>(G->src_data[c])[row];
        }
    }
}

// one warp per survivor: copy the bytes of every string column (lanes read consecutive bytes -> coalesced)
__global__ void gather_pass2(uint64_t n, const GatherCols *__restrict__ G, const uint64_t *__restrict__ rowlist) {
    const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (i > n) return;
    for (uint32_t c = 0; c < G->n_cols; ++c) {
        if (G->type[c] != TPLX_T_STR) continue;
        const uint64_t d0 = G->lens[c][i];  // exclusive scan: destination offset (entry n = total)
        if (lane == 0) G->dst_off[c][i] = (uint32_t)d0;
        if (i == n) continue;
        const uint32_t len = (uint32_t)(G->lens[c][i + 1] - d0);
        const uint8_t *src = reinterpret_cast<const uint8_t *>(G->src_data[c]) + G->srcpos[c][i];
        uint8_t *dst = G->dst_bytes[c] + d0;
        if (G->src_ref[c] && (G->src_ref[c][rowlist[i]] >> 63)) {  // cell with doubled quotes: dequoted by one lane
            if (lane == 0) {
                const uint32_t raw = (uint32_t)(G->src_ref[c][rowlist[i]] >> 32) & 0x7FFFFFFFu;
                csv_dequote(src, 0, raw, G->quote, dst, len);
            }
            continue;
        }
        for (uint32_t k = lane; k < len; k += 32) dst[k] = src[k];
    }
}

}  // namespace tplx
