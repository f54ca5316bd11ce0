// tplx_gpu_rowfmt.inl — host side of K5/K2 (included by tplx_gpu.cu).
// Partition / row / exception byte formats: see rowfmt.cuh header comment for the reference citations.

static int32_t device_scan(Device *d, const uint64_t *in, uint64_t *out, uint64_t n, bool write_total) {
    const uint32_t nb = (uint32_t)std::max<uint64_t>(1, (n + SCAN_ITEMS - 1) / SCAN_ITEMS);
    uint64_t *sums = nullptr;
    CU(cudaMallocAsync(&sums, (nb + 1) * 8, d->stream));
    scan_block_sums<<<nb, RF_NT, 0, d->stream>>>(in, sums, n);
    scan_of_sums<<<1, 1024, 0, d->stream>>>(sums, nb);
    scan_downsweep<<<nb, RF_NT, 0, d->stream>>>(in, out, sums, n, write_total ? 1 : 0);
    CU(cudaGetLastError());
    CU(cudaFreeAsync(sums, d->stream));
    return TPLX_OK;
}

// K in-place exclusive scans (totals written at [n]) of arrays that lie `stride` elements apart: 3 launches in all
static int32_t device_scan_batched(Device *d, uint64_t *arrays, uint64_t stride, uint32_t K, uint64_t n) {
    if (K == 0) return TPLX_OK;
    const uint32_t nb = (uint32_t)std::max<uint64_t>(1, (n + SCAN_ITEMS - 1) / SCAN_ITEMS);
    uint64_t *sums = nullptr;
    const uint64_t ss = nb + 1;
    CU(cudaMallocAsync(&sums, ss * K * 8, d->stream));
    scan_block_sums<<<dim3(nb, K), RF_NT, 0, d->stream>>>(arrays, sums, n, stride, ss);
    scan_of_sums<<<dim3(1, K), 1024, 0, d->stream>>>(sums, nb, ss);
    scan_downsweep<<<dim3(nb, K), RF_NT, 0, d->stream>>>(arrays, arrays, sums, n, 1, stride, ss);
    CU(cudaGetLastError());
    CU(cudaFreeAsync(sums, d->stream));
    return TPLX_OK;
}

// Walk the rows of one partition on the host to find row starts. Row length is only known from the row
// itself (fixed slots + optional var-len total), so this is inherently sequential per partition
// (Deserializer::inferLength, Serializer.cc:1227-1284); it touches 8-16 bytes per row.
static bool walk_partition(const uint8_t *part, uint64_t bytes, uint32_t n_cols, bool has_var, uint64_t base,
                           std::vector<uint64_t> &row_off, uint32_t bitmap_bytes = 0) {
    if (bytes < 8) return false;
    int64_t n_rows;
    memcpy(&n_rows, part, 8);
    if (n_rows < 0) return false;
    uint64_t pos = 8;
    for (int64_t r = 0; r < n_rows; ++r) {
        uint64_t fixed = bitmap_bytes + 8ull * n_cols + (has_var ? 8 : 0);
        if (pos + fixed > bytes) return false;
        row_off.push_back(base + pos);
        uint64_t var = 0;
        if (has_var) memcpy(&var, part + pos + bitmap_bytes + 8ull * n_cols, 8);
        pos += fixed + var;
        if (pos > bytes) return false;
    }
    return true;
}

extern "C" int32_t tplx_gpu_block_from_partitions(int32_t device, const uint8_t *const *partitions,
                                                  const uint64_t *partition_bytes, uint32_t n_partitions,
                                                  const uint8_t *col_types, uint32_t n_cols, tplx_block **out) {
    Device *d = get_device(device);
    if (!d) return fail(TPLX_E_NODEVICE, "block_from_partitions: device not initialised (no CPU fallback)");
    if (!out || !col_types || n_cols == 0 || n_cols > TPLX_MAX_COLS || (n_partitions && (!partitions || !partition_bytes)))
        return fail(TPLX_E_BADARG, "block_from_partitions: bad arguments");
    CU(cudaSetDevice(d->id));
    RowFmtCols C{};
    C.n_cols = n_cols;
    for (uint32_t c = 0; c < n_cols; ++c) {
        C.types[c] = col_types[c] & 0x3F;
        C.strk[c] = C.types[c] == TPLX_T_STR ? (int8_t)C.n_str++ : (int8_t)-1;
        C.optk[c] = (col_types[c] & TPLX_T_OPTION) ? (int8_t)C.n_opt++ : (int8_t)-1;  // Option[T] field: bit optk of the row bitmap
    }
    C.bitmap_bytes = (C.n_opt + 63) / 64 * 8;
    // 1. upload raw partition bytes (async) while the host walks row starts
    uint64_t total = 0;
    std::vector<uint64_t> pbase(n_partitions);
    for (uint32_t p = 0; p < n_partitions; ++p) {
        pbase[p] = total;
        total += partition_bytes[p];
    }
    uint8_t *raw = nullptr;
    CU(cudaMallocAsync(&raw, std::max<uint64_t>(total, 16), d->stream));
    for (uint32_t p = 0; p < n_partitions; ++p)
        CU(cudaMemcpyAsync(raw + pbase[p], partitions[p], partition_bytes[p], cudaMemcpyHostToDevice, d->stream));
    std::vector<uint64_t> row_off;
    for (uint32_t p = 0; p < n_partitions; ++p)
        if (!walk_partition(partitions[p], partition_bytes[p], n_cols, C.n_str > 0, pbase[p], row_off, C.bitmap_bytes)) {
            cudaFreeAsync(raw, d->stream);
            return fail(TPLX_E_BADARG, "block_from_partitions: malformed partition");
        }
    const uint64_t n = row_off.size();
    tplx_block *b = new tplx_block();
    b->dev = d;
    b->n_rows = n;
    uint64_t *d_row_off = nullptr, *lens = nullptr;
    CU(cudaMallocAsync(&d_row_off, std::max<uint64_t>(n, 1) * 8, d->stream));
    if (n) CU(cudaMemcpyAsync(d_row_off, row_off.data(), n * 8, cudaMemcpyHostToDevice, d->stream));
    if (C.n_str) {
        CU(cudaMallocAsync(&lens, (size_t)C.n_str * (n + 1) * 8, d->stream));
        CU(cudaMemsetAsync(lens, 0, (size_t)C.n_str * (n + 1) * 8, d->stream));
    }
    for (uint32_t c = 0; c < n_cols; ++c) {
        if (C.types[c] == TPLX_T_STR) {
            void *o = nullptr;
            CU(cudaMallocAsync(&o, (n + 1) * 4, d->stream));
            b->owned.push_back(o);
            C.offsets[c] = static_cast<uint32_t *>(o);
        } else {
            void *v = nullptr;
            CU(cudaMallocAsync(&v, std::max<uint64_t>(n, 2) * 8, d->stream));
            b->owned.push_back(v);
            C.data[c] = static_cast<uint64_t *>(v);
        }
    }
    for (uint32_t c = 0; c < n_cols; ++c) {
        if (C.optk[c] < 0) continue;
        void *v = nullptr;
        CU(cudaMallocAsync(&v, (n + 31) / 32 * 4 + 16, d->stream));
        b->owned.push_back(v);
        C.valid[c] = static_cast<uint32_t *>(v);
    }
    const uint32_t nb = (uint32_t)((n + RF_NT) / RF_NT);
    if (n) rows_to_cols_pass1<<<nb, RF_NT, 0, d->stream>>>(raw, d_row_off, n, C, lens);
    std::vector<uint64_t> str_total(C.n_str, 0);
    for (uint32_t k = 0; k < C.n_str; ++k) {
        int32_t rc = device_scan(d, lens + (size_t)k * (n + 1), lens + (size_t)k * (n + 1), n, true);
        if (rc) return rc;
        CU(cudaMemcpyAsync(&str_total[k], lens + (size_t)k * (n + 1) + n, 8, cudaMemcpyDeviceToHost, d->stream));
    }
    CU(cudaStreamSynchronize(d->stream));
    for (uint32_t c = 0; c < n_cols; ++c) {
        if (C.types[c] != TPLX_T_STR) continue;
        if (str_total[C.strk[c]] > 0xFFFFFFFFull) {
            tplx_gpu_block_free(b);
            return fail(TPLX_E_OVERFLOW, "block_from_partitions: string column exceeds 4 GiB; pass fewer partitions per block");
        }
        void *by = nullptr;
        CU(cudaMallocAsync(&by, std::max<uint64_t>(str_total[C.strk[c]], 16), d->stream));
        b->owned.push_back(by);
        C.bytes[c] = static_cast<uint8_t *>(by);
    }
    if (C.n_str) rows_to_cols_pass2<<<nb, RF_NT, 0, d->stream>>>(raw, d_row_off, n, C, lens);
    CU(cudaGetLastError());
    for (uint32_t c = 0; c < n_cols; ++c) {
        ColIn ci{};
        ci.type = C.types[c];
        if (C.types[c] == TPLX_T_STR) {
            ci.data = C.bytes[c];
            ci.offsets = C.offsets[c];
            b->data_bytes.push_back(str_total[C.strk[c]]);
        } else {
            ci.data = C.data[c];
            b->data_bytes.push_back(n * 8);
        }
        b->cols.push_back(ci);
        b->valid.push_back(C.optk[c] >= 0 ? C.valid[c] : nullptr);
    }
    CU(cudaFreeAsync(raw, d->stream));
    CU(cudaFreeAsync(d_row_off, d->stream));
    if (lens) CU(cudaFreeAsync(lens, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    *out = b;
    return TPLX_OK;
}

static void result_rowfmt(const tplx_result *r, RowFmtCols &C) {
    memset(&C, 0, sizeof(C));
    C.n_cols = (uint32_t)(r->out.size() - r->hidden);
    for (uint32_t c = 0; c < C.n_cols; ++c) {
        C.types[c] = r->out_types[c];
        if (C.types[c] == TPLX_T_STR) {
            C.strk[c] = (int8_t)C.n_str++;
            C.offsets[c] = r->out[c].offsets;
            C.bytes[c] = r->out[c].bytes;
        } else {
            C.strk[c] = -1;
            C.data[c] = r->out[c].data;
        }
        C.optk[c] = -1;
        if (c < r->out_valid.size() && r->out_valid[c]) {  // Option[T] output column: takes part in the row bitmap
            C.optk[c] = (int8_t)C.n_opt++;
            C.valid[c] = r->out_valid[c];
        }
    }
    C.bitmap_bytes = (C.n_opt + 63) / 64 * 8;
}

extern "C" int32_t tplx_gpu_result_partitions(tplx_result *r, uint64_t partition_bytes, uint8_t *buf, uint64_t buf_bytes,
                                              uint64_t *bytes_needed, uint64_t *part_offsets, uint32_t max_parts,
                                              uint32_t *n_parts) {
    if (!r || !bytes_needed || !n_parts || partition_bytes <= 8) return fail(TPLX_E_BADARG, "result_partitions: bad arguments");
    if (r->agg_out) {
        // aggregate result: one row with one slot per accumulator (LocalBackend.cc:1180-1207)
        uint64_t need = 8 + 8ull * r->n_accs;
        *bytes_needed = need;
        *n_parts = 1;
        if (!buf) return TPLX_OK;
        if (buf_bytes < need || max_parts < 1) return fail(TPLX_E_BADARG, "result_partitions: buffer too small");
        int64_t one = 1;
        memcpy(buf, &one, 8);
        int32_t rc = tplx_gpu_result_fetch_aggregate(r, reinterpret_cast<int64_t *>(buf + 8));
        if (rc) return rc;
        if (part_offsets) { part_offsets[0] = 0; part_offsets[1] = need; }
        return TPLX_OK;
    }
    Device *d = r->dev;
    std::lock_guard<std::mutex> lk(d->mu);
    CU(cudaSetDevice(d->id));
    RowFmtCols C;
    result_rowfmt(r, C);
    const uint64_t n = r->n_out;
    const uint32_t parts_cap = 1u << 16;
    uint64_t *sizes = nullptr, *first = nullptr;
    uint32_t *d_np = nullptr;
    CU(cudaMallocAsync(&sizes, (n + 1) * 8, d->stream));
    CU(cudaMallocAsync(&first, (parts_cap + 1) * 8, d->stream));
    CU(cudaMallocAsync(&d_np, 4, d->stream));
    const uint32_t nb = (uint32_t)((n + RF_NT) / RF_NT);
    cols_row_sizes<<<nb, RF_NT, 0, d->stream>>>(C, n, sizes);
    int32_t rc = device_scan(d, sizes, sizes, n, true);
    if (rc) return rc;
    split_partitions<<<1, 1, 0, d->stream>>>(sizes, n, partition_bytes - 8, first, parts_cap, d_np);
    uint32_t np = 0;
    uint64_t total_rows_bytes = 0;
    CU(cudaMemcpyAsync(&np, d_np, 4, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaMemcpyAsync(&total_rows_bytes, sizes + n, 8, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    auto cleanup = [&]() {
        cudaFreeAsync(sizes, d->stream);
        cudaFreeAsync(first, d->stream);
        cudaFreeAsync(d_np, d->stream);
    };
    if (np == 0xFFFFFFFFu) { cleanup(); return fail(TPLX_E_OVERFLOW, "result_partitions: too many partitions"); }
    const uint64_t need = total_rows_bytes + 8ull * np;
    *bytes_needed = need;
    *n_parts = np;
    if (!buf) { cleanup(); return TPLX_OK; }
    if (buf_bytes < need || max_parts < np) { cleanup(); return fail(TPLX_E_BADARG, "result_partitions: buffer too small"); }
    uint8_t *dout = nullptr;
    CU(cudaMallocAsync(&dout, std::max<uint64_t>(need, 16), d->stream));
    cols_to_rows<<<std::max<uint32_t>(nb, (np + RF_NT - 1) / RF_NT), RF_NT, 0, d->stream>>>(C, n, sizes, first, np, dout);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(buf, dout, need, cudaMemcpyDeviceToHost, d->stream));
    std::vector<uint64_t> h_first(np + 1), h_off(n ? 0 : 0);
    CU(cudaMemcpyAsync(h_first.data(), first, (np + 1) * 8, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    if (part_offsets) {
        // partition p starts at 8*p + row_off[first_p]; fetch the few row offsets needed
        for (uint32_t p = 0; p <= np; ++p) {
            uint64_t ro = 0;
            CU(cudaMemcpy(&ro, sizes + h_first[p], 8, cudaMemcpyDeviceToHost));
            part_offsets[p] = 8ull * p + ro;
        }
    }
    CU(cudaFreeAsync(dout, d->stream));
    cleanup();
    return TPLX_OK;
}

extern "C" int32_t tplx_gpu_result_exception_partition(tplx_result *r, uint8_t *buf, uint64_t buf_bytes, uint64_t *bytes_needed) {
    if (!r || !bytes_needed) return fail(TPLX_E_BADARG, "result_exception_partition: bad arguments");
    if (!r->block) return fail(TPLX_E_BADARG, "result_exception_partition: input block no longer available");
    for (uint8_t m : r->block->mapped)
        if (m == 2) return fail(TPLX_E_UNSUPPORTED, "result_exception_partition: block has lazy CSV columns (use the row text, tplx_gpu_csv_result_fetch_row_ends)");
    Device *d = r->dev;
    std::lock_guard<std::mutex> lk(d->mu);
    CU(cudaSetDevice(d->id));
    const tplx_block *b = r->block;
    RowFmtCols C;
    memset(&C, 0, sizeof(C));
    // the ORIGINAL input row in the normal-case input schema: the physical columns (not the `is None` companions the executor
    // appended for the op program); Option[T] input columns carry their bit in the row bitmap
    C.n_cols = (uint32_t)b->cols.size() - (r->stage ? r->stage->n_companions : 0);
    for (uint32_t c = 0; c < C.n_cols; ++c) {
        C.types[c] = (uint8_t)b->cols[c].type;
        if (C.types[c] == TPLX_T_STR) {
            C.strk[c] = (int8_t)C.n_str++;
            C.offsets[c] = const_cast<uint32_t *>(b->cols[c].offsets);
            C.bytes[c] = static_cast<uint8_t *>(const_cast<void *>(b->cols[c].data));
        } else {
            C.strk[c] = -1;
            C.data[c] = static_cast<uint64_t *>(const_cast<void *>(b->cols[c].data));
        }
        C.optk[c] = -1;
        if (c < b->valid.size() && b->valid[c]) {
            C.optk[c] = (int8_t)C.n_opt++;
            C.valid[c] = const_cast<uint32_t *>(b->valid[c]);
        }
    }
    C.bitmap_bytes = (C.n_opt + 63) / 64 * 8;
    const uint64_t ne = r->n_exc;
    if (ne == 0) {
        *bytes_needed = 8;
        if (buf) {
            if (buf_bytes < 8) return fail(TPLX_E_BADARG, "result_exception_partition: buffer too small");
            memset(buf, 0, 8);
        }
        return TPLX_OK;
    }
    uint64_t *sizes = nullptr;
    CU(cudaMallocAsync(&sizes, (ne + 1) * 8, d->stream));
    const uint32_t nb = (uint32_t)((ne + RF_NT) / RF_NT);
    exc_sizes<<<nb, RF_NT, 0, d->stream>>>(C, r->exc, ne, sizes);
    int32_t rc = device_scan(d, sizes, sizes, ne, true);
    if (rc) return rc;
    uint64_t total = 0;
    CU(cudaMemcpyAsync(&total, sizes + ne, 8, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    *bytes_needed = 8 + total;
    if (!buf) { cudaFreeAsync(sizes, d->stream); return TPLX_OK; }
    if (buf_bytes < 8 + total) { cudaFreeAsync(sizes, d->stream); return fail(TPLX_E_BADARG, "result_exception_partition: buffer too small"); }
    uint8_t *dout = nullptr;
    CU(cudaMallocAsync(&dout, 8 + total, d->stream));
    exc_write<<<nb, RF_NT, 0, d->stream>>>(C, r->exc, ne, sizes, dout);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(buf, dout, 8 + total, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    CU(cudaFreeAsync(dout, d->stream));
    CU(cudaFreeAsync(sizes, d->stream));
    return TPLX_OK;
}
