// tplx_gpu_csv.inl — host side of K6 (CSV bytes -> column block); included by tplx_gpu.cu.
// Kernel pipeline and reference citations: csv.cuh / csvops.cuh.

struct tplx_csv_buffer {
    Device *dev = nullptr;
    uint8_t *d = nullptr;  // n data bytes, '\n' at [n], zero padded to a CSV_TILE multiple
    uint64_t n = 0, padded = 0;
    cudaEvent_t ready = nullptr;
};

struct tplx_csv_result {
    Device *dev = nullptr;
    tplx_csv_info info{};
    tplx_csv_bad_row *bad = nullptr;  // device
    uint32_t *rowmap = nullptr;       // device
    uint32_t *row_end = nullptr;      // device: newline position of every row found (header included)
    uint32_t n_rows_total = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

extern "C" int32_t tplx_gpu_csv_upload(int32_t device, const void *bytes, uint64_t n_bytes, tplx_csv_buffer **out) {
    Device *d = get_device(device);
    if (!d) return fail(TPLX_E_NODEVICE, "csv_upload: device not initialised (no CPU fallback)");
    if (!out || (!bytes && n_bytes)) return fail(TPLX_E_BADARG, "csv_upload: bad arguments");
    if (n_bytes > 0xFFFFFFFFull - 4 * (uint64_t)CSV_TILE) return fail(TPLX_E_OVERFLOW, "csv_upload: pass less than 4 GiB per buffer (32-bit positions)");
    CU(cudaSetDevice(d->id));
    tplx_csv_buffer *b = new tplx_csv_buffer();
    b->dev = d;
    b->n = n_bytes;
    b->padded = align_up(n_bytes + 1, CSV_TILE);
    CU(cudaMallocAsync((void **)&b->d, b->padded + 64, d->copy_stream));  // slack: cells are read as whole 8-byte words
    if (n_bytes) CU(cudaMemcpyAsync(b->d, bytes, n_bytes, cudaMemcpyHostToDevice, d->copy_stream));
    CU(cudaMemsetAsync(b->d + n_bytes, 0, b->padded + 64 - n_bytes, d->copy_stream));
    CU(cudaMemsetAsync(b->d + n_bytes, '\n', 1, d->copy_stream));  // the newline VFCSVStreamCursor appends (CSVReader.cc:94-100)
    CU(cudaEventCreateWithFlags(&b->ready, cudaEventDisableTiming));
    CU(cudaEventRecord(b->ready, d->copy_stream));
    *out = b;
    return TPLX_OK;
}

extern "C" int32_t tplx_gpu_csv_buffer_free(tplx_csv_buffer *b) {
    if (!b) return TPLX_OK;
    cudaSetDevice(b->dev->id);
    if (b->d) cudaFreeAsync(b->d, b->dev->stream);
    if (b->ready) cudaEventDestroy(b->ready);
    delete b;
    return TPLX_OK;
}

namespace {
struct CsvTemps {  // device allocations that live for one parse call
    cudaStream_t st;
    std::vector<void *> ptrs;
    ~CsvTemps() {
        for (void *p : ptrs) cudaFreeAsync(p, st);
    }
    template <class T>
    cudaError_t alloc(T **p, size_t count) {
        void *v = nullptr;
        cudaError_t e = cudaMallocAsync(&v, std::max<size_t>(count * sizeof(T), 16), st);
        if (e == cudaSuccess) ptrs.push_back(v);
        *p = static_cast<T *>(v);
        return e;
    }
};
}  // namespace

extern "C" int32_t tplx_gpu_csv_parse(tplx_csv_buffer *cb, const tplx_csv_desc *desc, tplx_block **out_block,
                                      tplx_csv_result **out_res) {
    if (!cb || !desc || !out_block || !out_res || !desc->col_types || desc->n_file_cols == 0)
        return fail(TPLX_E_BADARG, "csv_parse: bad arguments");
    if (desc->delimiter == desc->quotechar || desc->delimiter == '\n' || desc->delimiter == '\r' || desc->quotechar == '\n' ||
        desc->quotechar == '\r' || desc->delimiter == 0 || desc->quotechar == 0)
        return fail(TPLX_E_BADARG, "csv_parse: delimiter / quotechar must be distinct non-newline, non-NUL bytes");
    Device *d = cb->dev;
    CU(cudaSetDevice(d->id));
    std::lock_guard<std::mutex> lk(d->mu);
    cudaStream_t st = d->stream;
    CU(cudaStreamWaitEvent(st, cb->ready, 0));

    // column map
    CsvParseParams P{};
    std::vector<uint8_t> kind(desc->n_file_cols), slot(desc->n_file_cols, 0);
    for (uint32_t c = 0; c < desc->n_file_cols; ++c) {
        const uint8_t t = desc->col_types[c];
        kind[c] = t;
        if (t == TPLX_CSV_SKIP) continue;
        if (t > TPLX_T_STR) return fail(TPLX_E_BADARG, "csv_parse: unknown column type");
        if (P.n_out == TPLX_MAX_COLS) return fail(TPLX_E_UNSUPPORTED, "csv_parse: more than TPLX_MAX_COLS columns selected");
        slot[c] = (uint8_t)P.n_out;
        P.out_types[P.n_out] = t;
        const bool lazy = t == TPLX_T_STR && desc->col_lazy && desc->col_lazy[c];
        P.strk[P.n_out] = (t == TPLX_T_STR && !lazy) ? (int8_t)P.n_str++ : (int8_t)-1;
        ++P.n_out;
    }
    if (desc->n_null_values > 8) return fail(TPLX_E_UNSUPPORTED, "csv_parse: at most 8 null values");
    P.nulls.n = desc->n_null_values;
    uint32_t nvb = 0;
    for (uint32_t k = 0; k < desc->n_null_values; ++k) {
        const size_t l = strlen(desc->null_values[k]);
        if (nvb + l > sizeof(P.nulls.bytes)) return fail(TPLX_E_UNSUPPORTED, "csv_parse: null values exceed 64 bytes");
        P.nulls.off[k] = (uint8_t)nvb;
        memcpy(P.nulls.bytes + nvb, desc->null_values[k], l);
        nvb += (uint32_t)l;
    }
    P.nulls.off[desc->n_null_values] = (uint8_t)nvb;

    tplx_csv_result *res = new tplx_csv_result();
    res->dev = d;
    std::unique_ptr<tplx_csv_result, int32_t (*)(tplx_csv_result *)> res_guard(res, tplx_gpu_csv_result_free);
    CU(cudaEventCreate(&res->ev0));
    CU(cudaEventCreate(&res->ev1));
    CU(cudaEventRecord(res->ev0, st));
    uint32_t launches = 0;

    CsvTemps T{st, {}};
    const uint32_t n_tiles = (uint32_t)(cb->padded / CSV_TILE);
    CsvState *tiles = nullptr;
    uint2 *tile_start = nullptr;
    uint32_t *totals = nullptr, *row_end = nullptr, *flags = nullptr, *span_state = nullptr;
    CU(T.alloc(&span_state, (size_t)n_tiles * CSV_NT));
    CU(T.alloc(&tiles, n_tiles));
    CU(T.alloc(&tile_start, n_tiles));
    CU(T.alloc(&totals, 2));
    CU(T.alloc(&flags, 4));
    CU(cudaMemsetAsync(flags, 0, 16, st));

    // ---- rows by quote parity ---------------------------------------------------------------------
    csv_tile_states<<<n_tiles, CSV_NT, 0, st>>>(cb->d, desc->quotechar, tiles, span_state);
    csv_scan_tiles<<<1, 1024, 0, st>>>(tiles, n_tiles, tile_start, totals);
    launches += 2;
    CU(cudaGetLastError());
    uint32_t h_tot[2] = {0, 0};
    CU(cudaMemcpyAsync(h_tot, totals, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    bool sequential = h_tot[0] != 0;  // file ends inside quotes by parity: let the exact machine decide
    uint8_t *d_kind = nullptr, *d_slot = nullptr;
    CU(T.alloc(&d_kind, desc->n_file_cols));
    CU(T.alloc(&d_slot, desc->n_file_cols));
    CU(cudaMemcpyAsync(d_kind, kind.data(), desc->n_file_cols, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_slot, slot.data(), desc->n_file_cols, cudaMemcpyHostToDevice, st));

    uint32_t n_rows_total = 0, nd = 0;
    uint64_t n_good = 0;
    std::vector<uint64_t> str_total(P.n_str, 0);
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (!sequential) {
            n_rows_total = h_tot[1];
            CU(cudaMallocAsync((void **)&row_end, ((size_t)n_rows_total + 1) * 4, st));
            res->row_end = row_end;
            if (n_rows_total) {
                csv_row_ends<<<n_tiles, CSV_NT, 0, st>>>(cb->d, desc->quotechar, tile_start, span_state, row_end);
                ++launches;
            }
        } else {
            if (res->row_end) CU(cudaFreeAsync(res->row_end, st));
            CU(cudaMallocAsync((void **)&row_end, ((size_t)(cb->n / 2) + 2) * 4, st));  // a row has at least one byte and a newline
            res->row_end = row_end;
            csv_rows_sequential<<<1, 32, 0, st>>>(cb->d, (uint32_t)cb->n, desc->delimiter, desc->quotechar, row_end, totals);
            ++launches;
            CU(cudaMemcpyAsync(h_tot, totals, 8, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            n_rows_total = h_tot[1];
        }
        CU(cudaGetLastError());
        P.r0 = desc->skip_header ? 1u : 0u;
        nd = n_rows_total > P.r0 ? n_rows_total - P.r0 : 0;
        P.buf = cb->d;
        P.n = (uint32_t)cb->n;
        P.row_end = row_end;
        P.nd = nd;
        P.delim = desc->delimiter;
        P.quote = desc->quotechar;
        P.n_file_cols = desc->n_file_cols;
        P.col_kind = d_kind;
        P.col_slot = d_slot;
        P.flags = flags;
        for (uint32_t c = 0; c < P.n_out; ++c) CU(T.alloc(&P.tmp[c], nd));
        CU(T.alloc(&P.lens, ((size_t)P.n_str + 1) * ((size_t)nd + 1)));  // string lengths, then the good flags: one batched scan
        P.good = P.lens + (size_t)P.n_str * ((size_t)nd + 1);
        CU(T.alloc(&P.code, nd));
        if (nd) {
            csv_parse_rows<<<(nd + CSV_NT - 1) / CSV_NT, CSV_NT, 0, st>>>(P);
            ++launches;
            CU(cudaGetLastError());
        }
        // output positions of the good rows and of their string bytes
        int32_t rc = device_scan_batched(d, P.lens, (uint64_t)nd + 1, P.n_str + 1, nd);
        if (rc) return rc;
        launches += 3;
        uint32_t h_flag = 0;
        CU(cudaMemcpyAsync(&h_flag, flags, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(&n_good, P.good + nd, 8, cudaMemcpyDeviceToHost, st));
        for (uint32_t k = 0; k < P.n_str; ++k)
            CU(cudaMemcpyAsync(&str_total[k], P.lens + (size_t)k * (nd + 1) + nd, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (!h_flag) break;
        if (sequential) return fail(TPLX_E_CUDA, "csv_parse: internal error (sequential rows failed verification)");
        sequential = true;  // irregular quoting: redo with the exact sequential row finder
        CU(cudaMemsetAsync(flags, 0, 16, st));
    }

    // ---- compaction into the block ----------------------------------------------------------------
    tplx_block *b = new tplx_block();
    b->dev = d;
    b->n_rows = n_good;
    std::unique_ptr<tplx_block, int32_t (*)(tplx_block *)> blk_guard(b, tplx_gpu_block_free);
    CsvCompactParams C{};
    C.nd = nd;
    C.r0 = P.r0;
    C.n_out = P.n_out;
    C.n_str = P.n_str;
    C.quote = desc->quotechar;
    C.lens = P.lens;
    C.good = P.good;
    C.code = P.code;
    C.row_end = row_end;
    C.buf = cb->d;
    const uint64_t n_bad = nd - n_good;
    for (uint32_t c = 0; c < P.n_out; ++c) {
        C.out_types[c] = P.out_types[c];
        C.strk[c] = P.strk[c];
        C.tmp[c] = P.tmp[c];
        ColIn ci{};
        ci.type = P.out_types[c];
        void *p = nullptr;
        if (P.out_types[c] == TPLX_T_STR && P.strk[c] < 0) {  // lazy: cell references into the CSV buffer
            CU(cudaMallocAsync(&p, std::max<uint64_t>(n_good * 8, 16), st));
            b->owned.push_back(p);
            C.refs[c] = static_cast<uint64_t *>(p);
            ci.data = cb->d;
            ci.offsets = reinterpret_cast<const uint32_t *>(p);
            b->data_bytes.push_back(n_good * 8);
            b->mapped.resize(P.n_out, 0);
            b->mapped[c] = 2;
            b->csv_quote = desc->quotechar;
        } else if (P.out_types[c] == TPLX_T_STR) {
            const uint64_t tot = str_total[P.strk[c]];
            if (tot > 0xFFFFFFFFull) return fail(TPLX_E_OVERFLOW, "csv_parse: string column exceeds 4 GiB");
            CU(cudaMallocAsync(&p, align_up(tot, 16) + 16, st));
            b->owned.push_back(p);
            C.bytes[c] = static_cast<uint8_t *>(p);
            void *o = nullptr;
            CU(cudaMallocAsync(&o, (n_good + 1) * 4, st));
            b->owned.push_back(o);
            CU(cudaMemsetAsync(o, 0, 4, st));
            C.offsets[c] = static_cast<uint32_t *>(o);
            ci.data = p;
            ci.offsets = C.offsets[c];
            b->data_bytes.push_back(tot);
        } else {
            CU(cudaMallocAsync(&p, std::max<uint64_t>(n_good * 8, 16), st));
            b->owned.push_back(p);
            C.data[c] = static_cast<uint64_t *>(p);
            ci.data = p;
            b->data_bytes.push_back(n_good * 8);
        }
        b->cols.push_back(ci);
    }
    CU(cudaMallocAsync((void **)&res->rowmap, std::max<uint64_t>(n_good * 4, 16), st));
    CU(cudaMallocAsync((void **)&res->bad, std::max<uint64_t>(n_bad * sizeof(tplx_csv_bad_row), 16), st));
    C.rowmap = res->rowmap;
    C.bad = res->bad;
    if (nd) {
        csv_compact<<<(nd + CSV_NT - 1) / CSV_NT, CSV_NT, 0, st>>>(C);
        ++launches;
        if (P.n_str) {
            csv_copy_strings<<<(nd + CSV_NT - 1) / CSV_NT, CSV_NT, 0, st>>>(C);  // one warp per 32 rows
            ++launches;
        }
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(res->ev1, st));
    CU(cudaEventCreateWithFlags(&b->ready, cudaEventDisableTiming));
    CU(cudaEventRecord(b->ready, st));
    res->n_rows_total = n_rows_total;
    res->info.n_rows = nd;
    res->info.n_normal = n_good;
    res->info.n_bad = n_bad;
    res->info.sequential_rows = sequential ? 1 : 0;
    res->info.kernel_launches = launches;
    *out_block = blk_guard.release();
    *out_res = res_guard.release();
    return TPLX_OK;
}

extern "C" int32_t tplx_gpu_csv_result_info(tplx_csv_result *r, tplx_csv_info *info) {
    if (!r || !info) return fail(TPLX_E_BADARG, "csv_result_info: bad arguments");
    CU(cudaSetDevice(r->dev->id));
    CU(cudaEventSynchronize(r->ev1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, r->ev0, r->ev1));
    r->info.parse_ms = ms;
    *info = r->info;
    return TPLX_OK;
}

extern "C" int32_t tplx_gpu_csv_result_fetch_bad_rows(tplx_csv_result *r, tplx_csv_bad_row *rows) {
    if (!r || (!rows && r->info.n_bad)) return fail(TPLX_E_BADARG, "csv_result_fetch_bad_rows: bad arguments");
    if (!r->info.n_bad) return TPLX_OK;
    CU(cudaSetDevice(r->dev->id));
    CU(cudaEventSynchronize(r->ev1));
    CU(cudaMemcpyAsync(rows, r->bad, r->info.n_bad * sizeof(tplx_csv_bad_row), cudaMemcpyDeviceToHost, r->dev->d2h_stream));
    CU(cudaStreamSynchronize(r->dev->d2h_stream));
    return TPLX_OK;
}

extern "C" int32_t tplx_gpu_csv_result_fetch_rowmap(tplx_csv_result *r, uint32_t *rowmap) {
    if (!r || (!rowmap && r->info.n_normal)) return fail(TPLX_E_BADARG, "csv_result_fetch_rowmap: bad arguments");
    if (!r->info.n_normal) return TPLX_OK;
    CU(cudaSetDevice(r->dev->id));
    CU(cudaEventSynchronize(r->ev1));
    CU(cudaMemcpyAsync(rowmap, r->rowmap, r->info.n_normal * 4, cudaMemcpyDeviceToHost, r->dev->d2h_stream));
    CU(cudaStreamSynchronize(r->dev->d2h_stream));
    return TPLX_OK;
}

extern "C" int32_t tplx_gpu_csv_result_fetch_row_ends(tplx_csv_result *r, uint32_t *ends) {
    if (!r || (!ends && r->info.n_rows)) return fail(TPLX_E_BADARG, "csv_result_fetch_row_ends: bad arguments");
    if (!r->info.n_rows) return TPLX_OK;
    CU(cudaSetDevice(r->dev->id));
    CU(cudaEventSynchronize(r->ev1));
    const uint32_t r0 = r->n_rows_total - (uint32_t)r->info.n_rows;  // 1 when a header row was skipped
    if (r0) {  // ends[0] = end of the header row, so that ends[i] / ends[i + 1] bracket data row i
        CU(cudaMemcpyAsync(ends, r->row_end, ((size_t)r->info.n_rows + 1) * 4, cudaMemcpyDeviceToHost, r->dev->d2h_stream));
    } else {
        ends[0] = 0xFFFFFFFFu;  // no row before data row 0: it starts at byte 0
        CU(cudaMemcpyAsync(ends + 1, r->row_end, (size_t)r->info.n_rows * 4, cudaMemcpyDeviceToHost, r->dev->d2h_stream));
    }
    CU(cudaStreamSynchronize(r->dev->d2h_stream));
    return TPLX_OK;
}

extern "C" int32_t tplx_gpu_csv_result_free(tplx_csv_result *r) {
    if (!r) return TPLX_OK;
    cudaSetDevice(r->dev->id);
    if (r->row_end) cudaFreeAsync(r->row_end, r->dev->stream);
    if (r->bad) cudaFreeAsync(r->bad, r->dev->stream);
    if (r->rowmap) cudaFreeAsync(r->rowmap, r->dev->stream);
    if (r->ev0) cudaEventDestroy(r->ev0);
    if (r->ev1) cudaEventDestroy(r->ev1);
    delete r;
    return TPLX_OK;
}

// ---- K7: CSV sink ----------------------------------------------------------------------------------------------------
extern "C" int32_t tplx_gpu_result_csv(tplx_result *r, uint32_t n_cols, uint8_t delimiter, uint8_t quotechar, uint8_t *buf,
                                       uint64_t buf_bytes, uint64_t *bytes_needed) {
    if (!r || !bytes_needed || r->agg_out) return fail(TPLX_E_BADARG, "result_csv: needs a row result");
    for (size_t c = 0; c + r->hidden < r->out.size() && c < r->out_valid.size() && (n_cols == 0 || c < n_cols); ++c)
        if (r->out_valid[c]) return fail(TPLX_E_UNSUPPORTED, "result_csv: Option[T] column (None cells): use the host row writer");
    Device *d = r->dev;
    CsvSinkCols C{};
    C.n_cols = (uint32_t)(r->out.size() - r->hidden);
    if (n_cols && n_cols < C.n_cols) C.n_cols = n_cols;
    C.delim = delimiter;
    C.quote = quotechar;
    for (uint32_t c = 0; c < C.n_cols; ++c) {
        C.types[c] = r->out_types[c];
        C.data[c] = r->out[c].data;
        C.offsets[c] = r->out[c].offsets;
        C.bytes[c] = r->out[c].bytes;
    }
    tplx_result_info info;
    int32_t rc = tplx_gpu_result_info(r, &info);  // synchronises the stage
    if (rc) return rc;
    const uint64_t n = r->n_out;
    if (n == 0 || C.n_cols == 0) {
        *bytes_needed = 0;
        return TPLX_OK;
    }
    std::lock_guard<std::mutex> lk(d->mu);
    CU(cudaSetDevice(d->id));
    CsvTemps T{d->stream, {}};
    uint64_t *sizes = nullptr;
    CU(T.alloc(&sizes, n + 1));
    const uint32_t nb = (uint32_t)((n + CSV_NT - 1) / CSV_NT);
    uint32_t *d_unsup = nullptr, h_unsup = 0;
    CU(T.alloc(&d_unsup, 4));
    CU(cudaMemsetAsync(d_unsup, 0, 16, d->stream));
    csv_sink_sizes<<<nb, CSV_NT, 0, d->stream>>>(C, n, sizes, d_unsup);
    rc = device_scan(d, sizes, sizes, n, true);
    if (rc) return rc;
    uint64_t total = 0;
    CU(cudaMemcpyAsync(&total, sizes + n, 8, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaMemcpyAsync(&h_unsup, d_unsup, 4, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    if (h_unsup) return fail(TPLX_E_UNSUPPORTED, "result_csv: an f64 value of magnitude >= 2^63 needs the host formatter");
    *bytes_needed = total;
    if (!buf) return TPLX_OK;
    if (buf_bytes < total) return fail(TPLX_E_BADARG, "result_csv: buffer too small");
    uint8_t *out = nullptr;
    CU(T.alloc(&out, total));
    csv_sink_write<<<nb, CSV_NT, 0, d->stream>>>(C, n, sizes, out);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(buf, out, total, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    r->launches += 5;
    return TPLX_OK;
}
