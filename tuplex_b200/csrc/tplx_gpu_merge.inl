// tplx_gpu_merge.inl — host side of K9 (merge.cuh), included by tplx_gpu.cu.
// Reference: ResolveTask::executeInOrder / emitNormalRows (tuplex/core/src/physical/ResolveTask.cc:300-375,878-1258).

extern "C" int32_t tplx_gpu_result_merge_resolved(tplx_result *res, const tplx_block *resolved, const int64_t *resolved_row_nos, int64_t first_row_no,
                                                  tplx_result **out) {
    if (!res || !resolved || !out || (resolved->n_rows && !resolved_row_nos)) return fail(TPLX_E_BADARG, "result_merge_resolved: bad arguments");
    Device *d = res->dev;
    if (resolved->dev->id != d->id) return fail(TPLX_E_BADARG, "result_merge_resolved: the resolved rows live on another device");
    // visible output columns = all minus the trailing internal ones (`is None` companions were folded into out_valid already)
    const size_t n_vis = res->out.size() - res->hidden;
    if (resolved->cols.size() != n_vis) return fail(TPLX_E_BADARG, "result_merge_resolved: resolved block must have the stage's output columns");
    for (size_t c = 0; c < n_vis; ++c)
        if ((uint8_t)resolved->cols[c].type != res->out_types[c]) return fail(TPLX_E_BADARG, "result_merge_resolved: column type differs from the stage's output schema");
    for (uint8_t m : resolved->mapped)
        if (m) return fail(TPLX_E_UNSUPPORTED, "result_merge_resolved: block columns must be device resident");
    const uint64_t n_norm = res->n_out, n_res = resolved->n_rows, n_excs = res->n_exc;
    if (n_res > n_excs) return fail(TPLX_E_BADARG, "result_merge_resolved: more resolved rows than exception records");
    if (n_norm + n_res >= 0x7FFFFFFFull) return fail(TPLX_E_OVERFLOW, "result_merge_resolved: more than 2^31 - 1 rows");
    std::lock_guard<std::mutex> lk(d->mu);
    CU(cudaSetDevice(d->id));
    // a_m = rows written before the exception that resolved row m replaces: its row number minus the exceptions before it
    std::vector<tplx_exception_rec> excs(n_excs);
    if (n_excs) {
        CU(cudaMemcpyAsync(excs.data(), res->exc, n_excs * sizeof(tplx_exception_rec), cudaMemcpyDeviceToHost, d->stream));
        CU(cudaStreamSynchronize(d->stream));
    }
    std::vector<int64_t> all(n_excs);
    for (uint64_t k = 0; k < n_excs; ++k) all[k] = excs[k].row_no;
    std::sort(all.begin(), all.end());
    std::vector<uint64_t> a_res(n_res);
    for (uint64_t m = 0; m < n_res; ++m) {
        if (m && resolved_row_nos[m] <= resolved_row_nos[m - 1]) return fail(TPLX_E_BADARG, "result_merge_resolved: row numbers must be strictly ascending");
        auto it = std::lower_bound(all.begin(), all.end(), resolved_row_nos[m]);
        if (it == all.end() || *it != resolved_row_nos[m]) return fail(TPLX_E_BADARG, "result_merge_resolved: a row number is not one of the result's exception records");
        const int64_t a = resolved_row_nos[m] - first_row_no - (int64_t)(it - all.begin());
        if (a < 0 || (uint64_t)a > n_norm) return fail(TPLX_E_BADARG, "result_merge_resolved: row number outside the block's output stream (first_row_no?)");
        a_res[m] = (uint64_t)a;
    }
    std::unique_ptr<tplx_result, int32_t (*)(tplx_result *)> rg(new tplx_result(), tplx_gpu_result_free);
    tplx_result *r = rg.get();
    r->dev = d;
    r->n_in = res->n_in;
    CU(cudaEventCreate(&r->ev0));
    CU(cudaEventCreate(&r->ev1));
    CU(cudaEventCreate(&r->evk0));
    CU(cudaEventCreate(&r->evk1));
    if (resolved->ready) CU(cudaStreamWaitEvent(d->stream, resolved->ready, 0));
    CU(cudaEventRecord(r->ev0, d->stream));
    CU(cudaEventRecord(r->evk0, d->stream));
    const uint64_t n = n_norm + n_res;
    uint64_t *d_a = nullptr;
    uint32_t *sel = nullptr;
    int32_t rc;
    if ((rc = dalloc(r, &d_a, n_res))) return rc;
    if ((rc = dalloc(r, &sel, n))) return rc;
    if (n_res) CU(cudaMemcpyAsync(d_a, a_res.data(), n_res * 8, cudaMemcpyHostToDevice, d->stream));
    const uint32_t nb = (uint32_t)((n + 255) / 256);
    if (n) merge_select_kernel<<<nb, 256, 0, d->stream>>>(d_a, n_res, n_norm, sel);
    r->launches = 1;
    r->n_out = n;
    r->out.assign(n_vis, OutCol{});
    r->out_types.assign(res->out_types.begin(), res->out_types.begin() + n_vis);
    r->str_bytes.assign(n_vis, 0);
    r->out_valid.assign(n_vis, nullptr);
    for (size_t c = 0; c < n_vis; ++c) {
        const uint32_t *va = c < res->out_valid.size() ? res->out_valid[c] : nullptr;
        const uint32_t *vb = c < resolved->valid.size() ? resolved->valid[c] : nullptr;
        if (r->out_types[c] == TPLX_T_STR) {
            uint64_t *lens = nullptr;
            CU(cudaMallocAsync((void **)&lens, (n + 1) * 8, d->stream));
            if (n) merge_str_len_kernel<<<nb, 256, 0, d->stream>>>(res->out[c].offsets, resolved->cols[c].offsets, sel, n, lens);
            rc = device_scan(d, lens, lens, n, true);
            if (rc) { cudaFreeAsync(lens, d->stream); return rc; }
            uint64_t tot = 0;
            CU(cudaMemcpyAsync(&tot, lens + n, 8, cudaMemcpyDeviceToHost, d->stream));
            CU(cudaStreamSynchronize(d->stream));  // a_res (host vector) has been consumed by now as well
            if (tot > 0xFFFFFFFFull) { cudaFreeAsync(lens, d->stream); return fail(TPLX_E_OVERFLOW, "result_merge_resolved: a string column exceeds 4 GiB"); }
            if ((rc = dalloc(r, &r->out[c].offsets, n + 1))) return rc;
            if ((rc = dalloc(r, &r->out[c].bytes, (size_t)align_up(tot, 16) + 16))) return rc;
            const uint32_t nw = (uint32_t)(((n + 1) * MERGE_STR_LANES + 255) / 256);
            merge_str_copy_kernel<<<nw, 256, 0, d->stream>>>(res->out[c].bytes, res->out[c].offsets, reinterpret_cast<const uint8_t *>(resolved->cols[c].data),
                                                            resolved->cols[c].offsets, sel, n, lens, r->out[c].offsets, r->out[c].bytes);
            CU(cudaFreeAsync(lens, d->stream));
            r->out[c].cap_bytes = tot;
            r->str_bytes[c] = tot;
            r->launches += 5;
        } else {
            if ((rc = dalloc(r, &r->out[c].data, n))) return rc;
            if (n) merge_fixed_kernel<<<nb, 256, 0, d->stream>>>(res->out[c].data, reinterpret_cast<const uint64_t *>(resolved->cols[c].data), sel, n, r->out[c].data);
            r->launches += 1;
        }
        if (va || vb) {
            uint32_t *w = nullptr;
            if ((rc = dalloc(r, &w, (n + 31) / 32 + 1))) return rc;
            if (n) merge_valid_kernel<<<nb, 256, 0, d->stream>>>(va, vb, sel, n, w);
            r->out_valid[c] = w;
            r->launches += 1;
        }
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(r->evk1, d->stream));
    CU(cudaEventRecord(r->ev1, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    *out = rg.release();
    return TPLX_OK;
}
