// tplx_gpu_join.inl — host side of K8 (join.cuh), included by tplx_gpu.cu.
// Reference: HashJoinStage (tuplex/core/src/physical/HashJoinStage.cc), the hash-table endpoint of the build stage
// (TransformTask.cc:769-842) and the probe inside the row pipeline (PipelineBuilder.cc:2110-2523).

struct tplx_join {
    Device *dev = nullptr;
    const tplx_block *build = nullptr;  // borrowed: must outlive the table
    uint32_t key_col = 0;
    JoinTableDev T{};
    std::vector<void *> owned;
    uint64_t n_rows = 0, n_keys = 0, n_null = 0;
    double build_ms = 0;
    uint32_t launches = 0;
};

static JoinKey join_key_of(const tplx_block *b, uint32_t col) {
    JoinKey k{};
    k.data = b->cols[col].data;
    k.offsets = b->cols[col].offsets;
    k.valid = col < b->valid.size() ? b->valid[col] : nullptr;
    k.type = (uint32_t)(b->cols[col].type & 0xFF);
    return k;
}

extern "C" int32_t tplx_gpu_join_destroy(tplx_join *j) {
    if (!j) return TPLX_OK;
    cudaSetDevice(j->dev->id);
    for (void *p : j->owned) cudaFreeAsync(p, j->dev->stream);
    delete j;
    return TPLX_OK;
}

extern "C" int32_t tplx_gpu_join_build(const tplx_block *b, uint32_t key_col, tplx_join **out) {
    if (!b || !out || key_col >= b->cols.size()) return fail(TPLX_E_BADARG, "join_build: bad arguments");
    const uint32_t kt = (uint32_t)(b->cols[key_col].type & 0xFF);
    if (kt == TPLX_T_F64) return fail(TPLX_E_UNSUPPORTED, "join_build: f64 keys are not hashable on this path (PipelineBuilder.cc:1175-1177)");
    if (b->n_rows >= 0xFFFFFFFEull) return fail(TPLX_E_OVERFLOW, "join_build: build side exceeds 2^32 - 2 rows");
    for (uint8_t m : b->mapped)
        if (m) return fail(TPLX_E_UNSUPPORTED, "join_build: block columns must be device resident");
    Device *d = b->dev;
    std::lock_guard<std::mutex> lk(d->mu);
    CU(cudaSetDevice(d->id));
    if (b->ready) CU(cudaStreamWaitEvent(d->stream, b->ready, 0));
    std::unique_ptr<tplx_join, int32_t (*)(tplx_join *)> j(new tplx_join(), tplx_gpu_join_destroy);
    j->dev = d;
    j->build = b;
    j->key_col = key_col;
    j->n_rows = b->n_rows;
    const uint64_t n = b->n_rows;
    uint64_t cap = 1024;
    while (cap < 2 * n) cap <<= 1;  // load <= 0.5
    JoinTableDev &T = j->T;
    T.cap = cap;
    T.mask = cap - 1;
    T.key = join_key_of(b, key_col);
    auto alloc = [&](void **p, size_t bytes) -> int32_t {
        CU(cudaMallocAsync(p, std::max<size_t>(bytes, 16), d->stream));
        j->owned.push_back(*p);
        return TPLX_OK;
    };
    int32_t rc;
    uint64_t *cnt = nullptr;
    uint32_t *row_group = nullptr, *cursor = nullptr, *big_list = nullptr, *n_big = nullptr, *tmp = nullptr;
    if ((rc = alloc((void **)&T.slots, cap * 4))) return rc;
    if ((rc = alloc((void **)&T.start, (cap + 2) * 8))) return rc;
    if ((rc = alloc((void **)&T.rows, n * 4))) return rc;
    // temporaries of the build (freed below, stream ordered)
    CU(cudaMallocAsync((void **)&cnt, (cap + 2) * 8, d->stream));
    CU(cudaMallocAsync((void **)&row_group, std::max<uint64_t>(n, 1) * 4, d->stream));
    CU(cudaMallocAsync((void **)&cursor, (cap + 1) * 4, d->stream));
    CU(cudaMallocAsync((void **)&big_list, (cap + 1) * 4, d->stream));
    CU(cudaMallocAsync((void **)&n_big, 16, d->stream));
    auto drop_tmp = [&]() {
        cudaFreeAsync(cnt, d->stream);
        cudaFreeAsync(row_group, d->stream);
        cudaFreeAsync(cursor, d->stream);
        cudaFreeAsync(big_list, d->stream);
        cudaFreeAsync(n_big, d->stream);
        if (tmp) cudaFreeAsync(tmp, d->stream);
    };
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CU(cudaEventRecord(e0, d->stream));
    CU(cudaMemsetAsync(T.slots, 0, cap * 4, d->stream));
    CU(cudaMemsetAsync(cnt, 0, (cap + 2) * 8, d->stream));
    CU(cudaMemsetAsync(cursor, 0, (cap + 1) * 4, d->stream));
    CU(cudaMemsetAsync(n_big, 0, 16, d->stream));
    const uint32_t nb = (uint32_t)((n + JOIN_NT - 1) / JOIN_NT);
    if (n) join_insert_kernel<<<nb, JOIN_NT, 0, d->stream>>>(T, n, row_group, cnt);
    rc = device_scan(d, cnt, T.start, cap + 1, true);  // start[cap + 1] = n
    if (rc) { drop_tmp(); return rc; }
    if (n) join_fill_kernel<<<nb, JOIN_NT, 0, d->stream>>>(T, n, row_group, cursor);
    const uint32_t ng = (uint32_t)((cap + 1 + JOIN_NT - 1) / JOIN_NT);
    join_order_small_kernel<<<ng, JOIN_NT, 0, d->stream>>>(T, big_list, n_big);
    CU(cudaGetLastError());
    // three counts for the host: large groups, rows in the null bucket, distinct keys (= non-empty groups, counted from cnt)
    uint64_t h_null[2] = {0, 0};
    uint32_t h_big = 0;
    CU(cudaMemcpyAsync(&h_big, n_big, 4, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaMemcpyAsync(h_null, T.start + cap, 16, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    j->n_null = h_null[1] - h_null[0];
    j->launches = 6;
    if (h_big) {
        CU(cudaMallocAsync((void **)&tmp, std::max<uint64_t>(n, 1) * 4, d->stream));
        join_order_big_kernel<<<h_big, JOIN_NT, 0, d->stream>>>(T, big_list, tmp);
        join_copy_big_kernel<<<h_big, JOIN_NT, 0, d->stream>>>(T, big_list, tmp);
        CU(cudaGetLastError());
        j->launches += 2;
    }
    CU(cudaEventRecord(e1, d->stream));
    drop_tmp();
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    j->build_ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *out = j.release();
    return TPLX_OK;
}

extern "C" int32_t tplx_gpu_join_info(const tplx_join *j, uint64_t *n_rows, uint64_t *n_null_rows, double *build_ms, uint32_t *kernel_launches) {
    if (!j) return fail(TPLX_E_BADARG, "join_info: bad arguments");
    if (n_rows) *n_rows = j->n_rows;
    if (n_null_rows) *n_null_rows = j->n_null;
    if (build_ms) *build_ms = j->build_ms;
    if (kernel_launches) *kernel_launches = j->launches;
    return TPLX_OK;
}

// gathers column `c` of `src` through idx[0 .. n_out) into output column `oc` of r; nullable: also a validity bitmap
static int32_t join_gather_column(tplx_result *r, size_t oc, const tplx_block *src, uint32_t c, const uint32_t *idx, uint64_t n_out, bool nullable) {
    Device *d = r->dev;
    const uint32_t t = (uint32_t)(src->cols[c].type & 0xFF);
    const uint32_t *sv = c < src->valid.size() ? src->valid[c] : nullptr;
    const uint32_t nb = (uint32_t)((n_out + JOIN_NT - 1) / JOIN_NT);
    int32_t rc;
    r->out_types[oc] = (uint8_t)t;
    if (t == TPLX_T_STR) {
        uint64_t *lens = nullptr;
        CU(cudaMallocAsync((void **)&lens, (n_out + 1) * 8, d->stream));
        if (n_out) join_str_len_kernel<<<nb, JOIN_NT, 0, d->stream>>>(src->cols[c].offsets, idx, n_out, lens);
        rc = device_scan(d, lens, lens, n_out, true);
        if (rc) { cudaFreeAsync(lens, d->stream); return rc; }
        uint64_t tot = 0;
        CU(cudaMemcpyAsync(&tot, lens + n_out, 8, cudaMemcpyDeviceToHost, d->stream));
        CU(cudaStreamSynchronize(d->stream));
        if (tot > 0xFFFFFFFFull) { cudaFreeAsync(lens, d->stream); return fail(TPLX_E_OVERFLOW, "join_probe: an output string column exceeds 4 GiB; probe smaller blocks"); }
        if ((rc = dalloc(r, &r->out[oc].offsets, n_out + 1))) return rc;
        if ((rc = dalloc(r, &r->out[oc].bytes, (size_t)align_up(tot, 16) + 16))) return rc;
        const uint32_t nw = (uint32_t)(((n_out + 1) * JOIN_STR_LANES + JOIN_NT - 1) / JOIN_NT);
        join_str_copy_kernel<<<nw, JOIN_NT, 0, d->stream>>>(reinterpret_cast<const uint8_t *>(src->cols[c].data), src->cols[c].offsets, idx, n_out, lens,
                                                            r->out[oc].offsets, r->out[oc].bytes);
        CU(cudaFreeAsync(lens, d->stream));
        r->out[oc].cap_bytes = tot;
        r->str_bytes[oc] = tot;
        r->launches += 5;
    } else {
        if ((rc = dalloc(r, &r->out[oc].data, n_out))) return rc;
        if (n_out) join_gather_fixed_kernel<<<nb, JOIN_NT, 0, d->stream>>>(reinterpret_cast<const uint64_t *>(src->cols[c].data), idx, n_out, r->out[oc].data);
        r->launches += 1;
    }
    if (nullable || sv) {
        uint32_t *w = nullptr;
        if ((rc = dalloc(r, &w, (n_out + 31) / 32 + 1))) return rc;
        if (n_out) join_gather_valid_kernel<<<nb, JOIN_NT, 0, d->stream>>>(sv, idx, n_out, w);
        r->out_valid[oc] = w;
        r->launches += 1;
    }
    CU(cudaGetLastError());
    return TPLX_OK;
}

extern "C" int32_t tplx_gpu_join_probe(tplx_join *j, const tplx_block *p, uint32_t key_col, uint32_t flags, tplx_result **out) {
    if (!j || !p || !out || key_col >= p->cols.size()) return fail(TPLX_E_BADARG, "join_probe: bad arguments");
    if (p->dev->id != j->dev->id) return fail(TPLX_E_BADARG, "join_probe: probe block and table live on different devices");
    Device *d = p->dev;  // the probe block's execution lane
    const tplx_block *b = j->build;
    const JoinKey pk = join_key_of(p, key_col);
    if (pk.type != j->T.key.type) return fail(TPLX_E_BADARG, "join_probe: key types differ (JoinOperator.cc:121-131)");
    if (p->n_rows >= 0xFFFFFFFFull) return fail(TPLX_E_OVERFLOW, "join_probe: probe block exceeds 2^32 - 1 rows");
    for (uint8_t m : p->mapped)
        if (m) return fail(TPLX_E_UNSUPPORTED, "join_probe: block columns must be device resident");
    const size_t n_outcols = p->cols.size() + b->cols.size() - 1;
    if (n_outcols > TPLX_MAX_COLS) return fail(TPLX_E_UNSUPPORTED, "join_probe: more than TPLX_MAX_COLS output columns");
    const bool left_outer = (flags & TPLX_JOIN_LEFT_OUTER) != 0, build_first = (flags & TPLX_JOIN_BUILD_FIRST) != 0;
    if (left_outer && build_first) return fail(TPLX_E_UNSUPPORTED, "join_probe: a left join with the build side on the left is a right join (PipelineBuilder.cc:2309-2313)");
    std::lock_guard<std::mutex> lk(d->mu);
    CU(cudaSetDevice(d->id));
    if (p->ready) CU(cudaStreamWaitEvent(d->stream, p->ready, 0));
    std::unique_ptr<tplx_result, int32_t (*)(tplx_result *)> rg(new tplx_result(), tplx_gpu_result_free);
    tplx_result *r = rg.get();
    r->dev = d;
    r->block = p;
    r->n_in = p->n_rows;
    CU(cudaEventCreate(&r->ev0));
    CU(cudaEventCreate(&r->ev1));
    CU(cudaEventCreate(&r->evk0));
    CU(cudaEventCreate(&r->evk1));
    CU(cudaEventRecord(r->ev0, d->stream));
    CU(cudaEventRecord(r->evk0, d->stream));
    const uint64_t n = p->n_rows;
    const uint32_t nb = (uint32_t)((n + JOIN_NT - 1) / JOIN_NT);
    uint32_t *grp = nullptr;
    uint64_t *cnt = nullptr;
    int32_t rc;
    if ((rc = dalloc(r, &grp, n))) return rc;
    if ((rc = dalloc(r, &cnt, n + 1))) return rc;
    if (n) join_probe_count_kernel<<<nb, JOIN_NT, 0, d->stream>>>(j->T, pk, n, left_outer ? 1u : 0u, grp, cnt);
    rc = device_scan(d, cnt, cnt, n, true);
    if (rc) return rc;
    uint64_t n_out = 0;
    CU(cudaMemcpyAsync(&n_out, cnt + n, 8, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    uint32_t *out_probe = nullptr, *out_build = nullptr;
    if ((rc = dalloc(r, &out_probe, n_out))) return rc;
    if ((rc = dalloc(r, &out_build, n_out))) return rc;
    if (n) join_probe_emit_kernel<<<nb, JOIN_NT, 0, d->stream>>>(j->T, n, grp, cnt, out_probe, out_build);
    CU(cudaGetLastError());
    r->launches = 5;
    r->n_out = n_out;
    r->out.assign(n_outcols, OutCol{});
    r->out_types.assign(n_outcols, TPLX_T_I64);
    r->str_bytes.assign(n_outcols, 0);
    r->out_valid.assign(n_outcols, nullptr);
    // | non-key columns of the first side | key (from the probe row) | non-key columns of the second side |  (JoinOperator.cc:163-184,
    // createInnerJoinBucketLoop PipelineBuilder.cc:2160-2198)
    size_t oc = 0;
    auto side = [&](const tplx_block *blk, uint32_t kc, const uint32_t *idx, bool nullable) -> int32_t {
        for (uint32_t c = 0; c < blk->cols.size(); ++c) {
            if (c == kc) continue;
            int32_t rc2 = join_gather_column(r, oc++, blk, c, idx, n_out, nullable);
            if (rc2) return rc2;
        }
        return TPLX_OK;
    };
    if (build_first) {
        if ((rc = side(b, j->key_col, out_build, false))) return rc;
        if ((rc = join_gather_column(r, oc++, p, key_col, out_probe, n_out, false))) return rc;
        if ((rc = side(p, key_col, out_probe, false))) return rc;
    } else {
        if ((rc = side(p, key_col, out_probe, false))) return rc;
        if ((rc = join_gather_column(r, oc++, p, key_col, out_probe, n_out, false))) return rc;
        if ((rc = side(b, j->key_col, out_build, left_outer))) return rc;
    }
    CU(cudaEventRecord(r->evk1, d->stream));
    CU(cudaEventRecord(r->ev1, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    *out = rg.release();
    return TPLX_OK;
}

extern "C" int32_t tplx_gpu_result_fetch_validity(tplx_result *r, uint32_t col, uint32_t *words, uint32_t *nullable) {
    if (!r || col + r->hidden >= r->out.size()) return fail(TPLX_E_BADARG, "result_fetch_validity: bad arguments");
    const uint32_t *w = col < r->out_valid.size() ? r->out_valid[col] : nullptr;
    if (nullable) *nullable = w ? 1 : 0;
    if (!w || !words || !r->n_out) return TPLX_OK;
    CU(cudaSetDevice(r->dev->id));
    CU(cudaMemcpyAsync(words, w, (r->n_out + 31) / 32 * 4, cudaMemcpyDeviceToHost, r->dev->d2h_stream));
    CU(cudaStreamSynchronize(r->dev->d2h_stream));
    return TPLX_OK;
}

extern "C" int32_t tplx_gpu_result_device_validity(tplx_result *r, uint32_t col, const uint32_t **words) {
    if (!r || !words || col + r->hidden >= r->out.size()) return fail(TPLX_E_BADARG, "result_device_validity: bad arguments");
    *words = col < r->out_valid.size() ? r->out_valid[col] : nullptr;
    return TPLX_OK;
}
