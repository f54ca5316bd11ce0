// tplx_gpu_hash.inl — host side of K4 (included by tplx_gpu.cu).

static uint64_t pow2_at_least(uint64_t v) {
    uint64_t p = 1024;
    while (p < v) p <<= 1;
    return p;
}

static int32_t hash_alloc(Device *d, tplx_stage *s, uint64_t cap, uint64_t heap_cap, HashTableDev &T) {
    memset(&T, 0, sizeof(T));
    T.cap = cap;
    T.mask = cap - 1;
    T.max_keys = cap / 2;  // load factor <= 0.5, like the reference's maps (utils/src/hashmap.cc:227-242)
    T.heap_cap = heap_cap;
    const uint32_t na = (uint32_t)s->accs.size();
    CU(cudaMalloc(&T.state, cap * 8));
    CU(cudaMalloc(&T.keyoff, cap * 8));
    CU(cudaMalloc(&T.accs, std::max<uint64_t>(1, (uint64_t)na) * cap * 8));
    CU(cudaMalloc(&T.heap, std::max<uint64_t>(heap_cap, 16)));
    CU(cudaMalloc(&T.counters, 8 * 8));
    CU(cudaMemsetAsync(T.state, 0, cap * 8, d->stream));
    CU(cudaMemsetAsync(T.counters, 0, 64, d->stream));
    for (uint32_t k = 0; k < na; ++k)
        hash_init_accs<<<(uint32_t)((cap + 255) / 256), 256, 0, d->stream>>>(T, na, nullptr, k, s->accs[k].kind);
    CU(cudaGetLastError());
    return TPLX_OK;
}

static uint32_t key_blob_estimate(const tplx_stage *s) {
    uint32_t b = 0;
    for (uint32_t k = 0; k < s->hdr.n_keys; ++k) b += s->out_cols[k].type == TPLX_T_STR ? 4 + 28 : 8;
    return b;
}

// grow table (and heap) so that at least want_keys keys / want_heap bytes fit; existing entries are rehashed
static int32_t hash_grow(Device *d, tplx_stage *s, StageDev *sd, uint64_t want_keys, uint64_t want_heap) {
    HashTable *old = sd->ht;
    uint64_t cap = pow2_at_least(want_keys * 2 + 1024);
    uint64_t heap = std::max<uint64_t>(want_heap, 1 << 20);
    if (old && old->d.cap >= cap && old->d.heap_cap >= heap) return TPLX_OK;
    if (old) { cap = std::max(cap, old->d.cap); heap = std::max(heap, old->d.heap_cap); }
    HashTable *nt = new HashTable();
    nt->device = d->id;
    nt->n_accs = (uint32_t)s->accs.size();
    int32_t rc = hash_alloc(d, s, cap, heap, nt->d);
    if (rc) return rc;
    if (old) {
        uint64_t cnt[4];
        CU(cudaMemcpyAsync(cnt, old->d.counters, 32, cudaMemcpyDeviceToHost, d->stream));
        CU(cudaStreamSynchronize(d->stream));
        CU(cudaMemcpyAsync(nt->d.heap, old->d.heap, std::min<uint64_t>(cnt[1], old->d.heap_cap), cudaMemcpyDeviceToDevice, d->stream));
        uint32_t *kinds = nullptr;
        hash_rehash<<<(uint32_t)((old->d.cap + 255) / 256), 256, 0, d->stream>>>(old->d, nt->d, nt->n_accs, kinds);
        CU(cudaGetLastError());
        uint64_t ncnt[8] = {cnt[0], std::min<uint64_t>(cnt[1], old->d.heap_cap), 0, 0, 0, 0, 0, 0};
        CU(cudaMemcpyAsync(nt->d.counters, ncnt, 64, cudaMemcpyHostToDevice, d->stream));
        CU(cudaStreamSynchronize(d->stream));
        hash_table_destroy(old);
    }
    sd->ht = nt;
    return TPLX_OK;
}

extern "C" int32_t tplx_gpu_stage_hash_reserve(tplx_stage *s, int32_t device, uint64_t expected_keys) {
    Device *d = get_device(device);
    if (!d) return fail(TPLX_E_NODEVICE, "hash_reserve: device not initialised (no CPU fallback)");
    if (!s || s->hdr.endpoint != TPLX_EP_HASH) return fail(TPLX_E_BADARG, "hash_reserve: not a hash stage");
    std::lock_guard<std::mutex> lk(d->mu);
    CU(cudaSetDevice(d->id));
    StageDev *sd = nullptr;
    int32_t rc = stage_dev(s, d, &sd);
    if (rc) return rc;
    return hash_grow(d, s, sd, expected_keys, expected_keys * key_blob_estimate(s));
}

// shared by run_hash and hash_merge: launch K4 over a block with a given program
static int32_t launch_hash(tplx_stage *s, StageDev *sd, const tplx_block *b, KParams &P, const Layout &L, HashParams &H,
                           tplx_result *r, uint64_t *n_exc_out) {
    Device *d = sd->dev;
    const uint64_t n = b->n_rows;
    const uint32_t na = P.n_accs;
    const uint32_t smem = L.misc_off + (uint32_t)((size_t)na * SM_BUCKETS * 8 + SM_BUCKETS * 4 + 32);
    if (smem > (uint32_t)d->smem_optin) return fail(TPLX_E_UNSUPPORTED, "hash stage needs more shared memory than one SM has");
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, stage_hash_kernel, NT, smem));
    if (occ < 1) return fail(TPLX_E_UNSUPPORTED, "hash stage kernel cannot be resident");
    KParams *dP = nullptr;
    HashParams *dH = nullptr;
    uint32_t *counters = nullptr, *ovf = nullptr, *ovf2 = nullptr;
    int32_t rc = dalloc(r, &dP, 1);
    if (rc) return rc;
    rc = dalloc(r, &dH, 1);
    if (rc) return rc;
    rc = dalloc(r, &counters, 4);
    if (rc) return rc;
    P.counters = counters;
    P.scratch_per_thread = s->materialises ? std::max<uint32_t>(s->hdr.scratch_bytes, 64) : 0;
    uint64_t cap_exc = std::max<uint64_t>(4096, n / 32);
    rc = dalloc(r, &P.exc, cap_exc);
    if (rc) return rc;
    P.cap_exc = cap_exc;
    if (!sd->ht) {
        // no reservation: size for the worst case of this block, bounded; grows on demand below
        rc = hash_grow(d, s, sd, std::min<uint64_t>(n, 1ull << 22), std::min<uint64_t>(n, 1ull << 22) * key_blob_estimate(s));
        if (rc) return rc;
    }
    const uint64_t cap_ovf = std::max<uint64_t>(n, 1);
    rc = dalloc(r, &ovf, cap_ovf);
    if (rc) return rc;
    rc = dalloc(r, &ovf2, cap_ovf);
    if (rc) return rc;
    const uint32_t *rowlist = nullptr;
    uint64_t n_list = 0;
    CU(cudaEventRecord(r->evk0, d->stream));
    for (int round = 0; round < 40; ++round) {
        H.ht = sd->ht->d;
        H.rowlist = rowlist;
        H.n_list = n_list;
        H.overflow_rows = (round & 1) ? ovf2 : ovf;
        H.cap_overflow = cap_ovf;
        const uint64_t n_work = rowlist ? n_list : n;
        const uint32_t n_tiles = (uint32_t)((n_work + (uint64_t)P.R * NT - 1) / ((uint64_t)P.R * NT));
        const uint32_t grid = std::max<uint32_t>(1, std::min<uint32_t>(n_tiles, (uint32_t)(occ * d->prop.multiProcessorCount)));
        rc = ensure_scratch(d, (size_t)grid * NT * P.scratch_per_thread);
        if (rc) return rc;
        P.scratch = d->scratch;
        // counters[2] (overflow rows) restarts every round; [3] (exceptions) accumulates
        CU(cudaMemsetAsync(sd->ht->d.counters + 2, 0, 8, d->stream));
        if (round == 0) {
            CU(cudaMemsetAsync(sd->ht->d.counters + 3, 0, 8, d->stream));
            CU(cudaMemsetAsync(counters, 0, 16, d->stream));
        }
        CU(cudaMemcpyAsync(dP, &P, sizeof(P), cudaMemcpyHostToDevice, d->stream));
        CU(cudaMemcpyAsync(dH, &H, sizeof(H), cudaMemcpyHostToDevice, d->stream));
        if (n_work) {
            stage_hash_kernel<<<grid, NT, smem, d->stream>>>(dP, dH);
            CU(cudaGetLastError());
            r->launches += 1;
        }
        uint64_t cnt[4];
        CU(cudaMemcpyAsync(cnt, sd->ht->d.counters, 32, cudaMemcpyDeviceToHost, d->stream));
        CU(cudaStreamSynchronize(d->stream));
        *n_exc_out = cnt[3];
        if (cnt[2] == 0) break;
        // some rows found the table or heap full: grow x4 and retry exactly those rows
        rowlist = H.overflow_rows;
        n_list = cnt[2];
        rc = hash_grow(d, s, sd, std::max<uint64_t>(sd->ht->d.max_keys * 4, cnt[0] + n_list), std::max<uint64_t>(sd->ht->d.heap_cap * 4, cnt[1] * 2));
        if (rc) return rc;
        if (round == 39) return fail(TPLX_E_OVERFLOW, "hash table growth did not converge");
    }
    CU(cudaEventRecord(r->evk1, d->stream));
    return TPLX_OK;
}

static int32_t run_hash(tplx_stage *s, StageDev *sd, const tplx_block *b, tplx_result *r) {
    const uint32_t R = 4;
    Layout L = make_layout(s, R, false);
    KParams P;
    fill_common(P, s, sd, b, L, R);
    HashParams H;
    memset(&H, 0, sizeof(H));
    H.n_keys = s->hdr.n_keys;
    for (uint32_t k = 0; k < H.n_keys; ++k) {
        H.key_slot[k] = s->out_cols[k].slot;
        H.key_type[k] = s->out_cols[k].type;
    }
    uint64_t n_exc = 0;
    int32_t rc = launch_hash(s, sd, b, P, L, H, r, &n_exc);
    if (rc) return rc;
    r->n_exc = std::min<uint64_t>(n_exc, P.cap_exc);
    r->exc = P.exc;
    r->n_out = 0;
    if (r->n_exc) {
        std::vector<tplx_exception_rec> recs(r->n_exc);
        CU(cudaMemcpy(recs.data(), r->exc, r->n_exc * sizeof(tplx_exception_rec), cudaMemcpyDeviceToHost));
        std::sort(recs.begin(), recs.end(), [](const tplx_exception_rec &a, const tplx_exception_rec &b) { return a.row < b.row; });
        for (size_t i = 0; i < recs.size(); ++i) recs[i].row_no = (int64_t)i;
        CU(cudaMemcpy(r->exc, recs.data(), r->n_exc * sizeof(tplx_exception_rec), cudaMemcpyHostToDevice));
    }
    return TPLX_OK;
}

// packed block = key columns followed by one 8-byte column per accumulator (raw partials, no init)
static int32_t hash_merge_locked(tplx_stage *s, const tplx_block *packed);
extern "C" int32_t tplx_gpu_stage_hash_merge(tplx_stage *s, const tplx_block *packed) {
    if (!s || !packed || s->hdr.endpoint != TPLX_EP_HASH) return fail(TPLX_E_BADARG, "hash_merge: bad arguments");
    std::lock_guard<std::mutex> lk(packed->dev->mu);
    return hash_merge_locked(s, packed);
}
static int32_t hash_merge_locked(tplx_stage *s, const tplx_block *packed) {
    const uint32_t nk = s->hdr.n_keys, na = (uint32_t)s->accs.size();
    if (packed->cols.size() != nk + na) return fail(TPLX_E_BADARG, "hash_merge: packed block must hold key + accumulator columns");
    if (packed->n_rows == 0) return TPLX_OK;
    Device *d = packed->dev;
    CU(cudaSetDevice(d->id));
    if (packed->ready) CU(cudaStreamWaitEvent(d->stream, packed->ready, 0));
    StageDev *sd = nullptr;
    int32_t rc = stage_dev(s, d, &sd);
    if (rc) return rc;
    // identity program: load every packed column into its own slot(s)
    tplx_stage tmp;
    tmp.hdr = s->hdr;
    tmp.accs = s->accs;
    uint32_t slot = 0;
    std::vector<uint32_t> col_slot;
    for (uint32_t c = 0; c < nk + na; ++c) {
        tplx_instr in{};
        in.op = TPLX_OP_LDCOL;
        in.flags = (uint8_t)packed->cols[c].type;
        in.dst = (uint16_t)slot;
        in.a = in.b = in.c = in.guard = TPLX_NOSLOT;
        in.imm = c;
        tmp.instrs.push_back(in);
        tmp.in_types.push_back((uint8_t)packed->cols[c].type);
        col_slot.push_back(slot);
        slot += packed->cols[c].type == TPLX_T_STR ? 2 : 1;
    }
    tmp.hdr.n_slots = (uint16_t)slot;
    tmp.hdr.n_instr = nk + na;
    for (uint32_t k = 0; k < na; ++k) tmp.accs[k].slot = (uint16_t)col_slot[nk + k];
    DInstr *dprog = nullptr;
    std::vector<DInstr> dec = predecode(tmp.instrs);
    CU(cudaMallocAsync(&dprog, dec.size() * sizeof(DInstr), d->stream));
    CU(cudaMemcpyAsync(dprog, dec.data(), dec.size() * sizeof(DInstr), cudaMemcpyHostToDevice, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    StageDev tsd = *sd;
    tsd.prog = dprog;
    const uint32_t R = 4;
    Layout L = make_layout(&tmp, R, false);
    KParams P;
    fill_common(P, &tmp, &tsd, packed, L, R);
    HashParams H;
    memset(&H, 0, sizeof(H));
    H.n_keys = nk;
    for (uint32_t k = 0; k < nk; ++k) {
        H.key_slot[k] = col_slot[k];
        H.key_type[k] = (uint32_t)packed->cols[k].type;
    }
    tplx_result tr;
    tr.dev = d;
    CU(cudaEventCreate(&tr.evk0));
    CU(cudaEventCreate(&tr.evk1));
    uint64_t n_exc = 0;
    rc = launch_hash(s, &tsd, packed, P, L, H, &tr, &n_exc);
    sd->ht = tsd.ht;  // growth may have replaced the table
    for (void *p : tr.owned) cudaFreeAsync(p, d->stream);
    cudaFreeAsync(dprog, d->stream);
    cudaEventDestroy(tr.evk0);
    cudaEventDestroy(tr.evk1);
    CU(cudaStreamSynchronize(d->stream));
    return rc;
}

// raw: accumulators without the initial value — the form exchanged between GPUs before the owner finishes.
// sel: which slots to emit (all, or the keys one rank owns).
static int32_t hash_export_locked(tplx_stage *s, Device *d, bool raw, HashSel sel, tplx_result **out);
static int32_t hash_finish_impl(tplx_stage *s, int32_t device, bool raw, tplx_result **out) {
    Device *d = get_device(device);
    if (!d) return fail(TPLX_E_NODEVICE, "hash_finish: device not initialised (no CPU fallback)");
    if (!s || !out || s->hdr.endpoint != TPLX_EP_HASH) return fail(TPLX_E_BADARG, "hash_finish: not a hash stage");
    std::lock_guard<std::mutex> lk(d->mu);
    return hash_export_locked(s, d, raw, HashSel{-1, 1}, out);
}
static int32_t hash_export_locked(tplx_stage *s, Device *d, bool raw, HashSel sel, tplx_result **out) {
    CU(cudaSetDevice(d->id));
    StageDev *sd = nullptr;
    int32_t rc = stage_dev(s, d, &sd);
    if (rc) return rc;
    if (!sd->ht) {
        rc = hash_grow(d, s, sd, 1024, 1 << 20);
        if (rc) return rc;
    }
    HashTableDev &T = sd->ht->d;
    const uint32_t nk = s->hdr.n_keys, na = (uint32_t)s->accs.size();
    tplx_result *r = new tplx_result();
    r->dev = d;
    r->stage = s;
    CU(cudaEventCreate(&r->ev0));
    CU(cudaEventCreate(&r->ev1));
    CU(cudaEventCreate(&r->evk0));
    CU(cudaEventCreate(&r->evk1));
    CU(cudaEventRecord(r->ev0, d->stream));
    CU(cudaEventRecord(r->evk0, d->stream));
    uint64_t *pos = nullptr;
    rc = dalloc(r, &pos, T.cap + 1);
    if (rc) return rc;
    const uint32_t nb = (uint32_t)((T.cap + 255) / 256);
    hash_flag_slots<<<nb, 256, 0, d->stream>>>(T, pos, sel);
    rc = device_scan(d, pos, pos, T.cap, true);
    if (rc) return rc;
    uint64_t n_out = 0;
    CU(cudaMemcpyAsync(&n_out, pos + T.cap, 8, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    HashEmit E;
    memset(&E, 0, sizeof(E));
    E.n_keys = nk;
    E.n_accs = na;
    uint32_t *d_types = nullptr;
    rc = dalloc(r, &d_types, TPLX_MAX_KEYS);
    if (rc) return rc;
    uint32_t h_types[TPLX_MAX_KEYS] = {0};
    for (uint32_t k = 0; k < nk; ++k) h_types[k] = E.key_types[k] = s->out_cols[k].type;
    CU(cudaMemcpyAsync(d_types, h_types, sizeof(h_types), cudaMemcpyHostToDevice, d->stream));
    r->out.assign(nk + na, OutCol{});
    r->out_types.assign(nk + na, TPLX_T_I64);
    r->str_bytes.assign(nk + na, 0);
    for (uint32_t k = 0; k < nk; ++k) {
        r->out_types[k] = (uint8_t)E.key_types[k];
        if (E.key_types[k] == TPLX_T_STR) {
            uint64_t *lens = nullptr;
            rc = dalloc(r, &lens, n_out + 1);
            if (rc) return rc;
            CU(cudaMemsetAsync(lens, 0, (n_out + 1) * 8, d->stream));
            hash_key_lens<<<nb, 256, 0, d->stream>>>(T, pos, nk, d_types, k, lens, sel);
            rc = device_scan(d, lens, lens, n_out, true);
            if (rc) return rc;
            uint64_t tot = 0;
            CU(cudaMemcpyAsync(&tot, lens + n_out, 8, cudaMemcpyDeviceToHost, d->stream));
            CU(cudaStreamSynchronize(d->stream));
            if (tot > 0xFFFFFFFFull) return fail(TPLX_E_OVERFLOW, "hash_finish: key column exceeds 4 GiB");
            E.key_lens_scan[k] = lens;
            rc = dalloc(r, &E.key_offsets[k], n_out + 1);
            if (rc) return rc;
            rc = dalloc(r, &E.key_bytes[k], tot);
            if (rc) return rc;
            r->out[k].offsets = E.key_offsets[k];
            r->out[k].bytes = E.key_bytes[k];
            r->str_bytes[k] = tot;
        } else {
            rc = dalloc(r, &E.key_data[k], n_out);
            if (rc) return rc;
            r->out[k].data = E.key_data[k];
        }
    }
    for (uint32_t k = 0; k < na; ++k) {
        E.acc_kinds[k] = s->accs[k].kind;
        // raw: combining with the identity leaves the partial unchanged
        E.acc_init[k] = raw ? 0 : s->accs[k].init;
        rc = dalloc(r, &E.acc_data[k], n_out);
        if (rc) return rc;
        r->out[nk + k].data = E.acc_data[k];
        const uint8_t kind = s->accs[k].kind;
        r->out_types[nk + k] = (kind == TPLX_ACC_SUM_F64 || kind == TPLX_ACC_MIN_F64 || kind == TPLX_ACC_MAX_F64) ? TPLX_T_F64 : TPLX_T_I64;
    }
    if (raw)
        for (uint32_t k = 0; k < na; ++k) {
            // identity per kind so that acc_combine(init, v) == v
            switch (s->accs[k].kind) {
                case TPLX_ACC_SUM_I64: case TPLX_ACC_SUM_F64: E.acc_init[k] = 0; break;
                case TPLX_ACC_MIN_I64: E.acc_init[k] = INT64_MAX; break;
                case TPLX_ACC_MAX_I64: E.acc_init[k] = INT64_MIN; break;
                case TPLX_ACC_MIN_F64: E.acc_init[k] = 0x7FF0000000000000ll; break;
                default: E.acc_init[k] = (int64_t)0xFFF0000000000000ull; break;
            }
        }
    hash_emit<<<nb, 256, 0, d->stream>>>(T, pos, n_out, E, sel);
    CU(cudaGetLastError());
    CU(cudaEventRecord(r->evk1, d->stream));
    CU(cudaEventRecord(r->ev1, d->stream));
    r->launches = 4;
    r->n_out = n_out;
    r->n_in = 0;
    CU(cudaStreamSynchronize(d->stream));
    *out = r;
    return TPLX_OK;
}

extern "C" int32_t tplx_gpu_stage_hash_finish(tplx_stage *s, int32_t device, tplx_result **out) {
    return hash_finish_impl(s, device, false, out);
}
extern "C" int32_t tplx_gpu_stage_hash_export_raw(tplx_stage *s, int32_t device, tplx_result **out) {
    return hash_finish_impl(s, device, true, out);
}
extern "C" int32_t tplx_gpu_stage_hash_reset(tplx_stage *s, int32_t device) {
    Device *d = get_device(device);
    if (!d || !s) return fail(TPLX_E_BADARG, "hash_reset: bad arguments");
    std::lock_guard<std::mutex> lk(d->mu);
    for (auto &sd : s->devs)
        if (sd.dev == d && sd.ht) {
            CU(cudaSetDevice(d->id));
            CU(cudaStreamSynchronize(d->stream));
            hash_table_destroy(sd.ht);
            sd.ht = nullptr;
        }
    return TPLX_OK;
}
