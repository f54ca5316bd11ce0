// tplx_gpu_comm.inl — the one exchange step of the path, inside the C ABI (included by tplx_gpu.cu).
//
// Reference counterparts: the driver-side combine of per-task partial aggregates (TransformTask::combineAggregate /
// fetchAggregate, core/src/physical/TransformTask.cc:218-299, called from LocalBackend.cc:917-960) and the merge of the per-task
// hash tables into the final map (LocalBackend::createFinalHashmap, core/src/ee/local/LocalBackend.cc:2219-2376).
// Here every GPU is one rank of an NCCL communicator (NVLink 5 / NVSwitch):
//   aggregate       ncclAllGather of the per-rank partial (n_accs x 8 bytes), folded in RANK ORDER by a one-warp kernel
//                   -> every rank holds the same bits, f64 sums are reproducible (fixed association)
//   aggregateByKey  every key has an owner rank = hash(key) mod world. Each rank splits its table into `world` packed
//                   (key columns, raw partial columns) blocks on the device, sizes are exchanged with one ncclAllGather, the
//                   payload with grouped ncclSend/ncclRecv straight between device buffers (an all-to-all over NVSwitch),
//                   and every rank merges what it owns into a fresh table with the accumulators' combine operation.
// NCCL is resolved at run time (dlopen of the copy the process already holds, e.g. PyTorch's, else libnccl.so.2), so the
// library itself has no link-time dependency; without NCCL these entry points return TPLX_E_UNSUPPORTED.
#include <dlfcn.h>

extern "C" {
typedef struct ncclComm *tplx_ncclComm_t;
typedef struct { char internal[128]; } tplx_ncclUniqueId;
}
static_assert(TPLX_COMM_ID_BYTES == sizeof(tplx_ncclUniqueId), "unique id size");

namespace {
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(tplx_ncclUniqueId *) = nullptr;
    int (*CommInitRank)(tplx_ncclComm_t *, int, tplx_ncclUniqueId, int) = nullptr;
    int (*CommInitAll)(tplx_ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(tplx_ncclComm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, tplx_ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, tplx_ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, tplx_ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
constexpr int NCCL_UINT8 = 1, NCCL_UINT64 = 5;  // ncclDataType_t (nccl.h)

NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {  // the copy this process already loaded (PyTorch bundles one) wins
            api.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
            if (api.lib) break;
        }
        for (const char *n : names) {
            if (api.lib) break;
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        }
        if (!api.lib) return;
        auto sym = [&](const char *n) { return dlsym(api.lib, n); };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.AllGather && api.Send && api.Recv &&
                 api.GroupStart && api.GroupEnd && api.GetErrorString;
    });
    return api.ok ? &api : nullptr;
}
}  // namespace

struct DeviceComm {
    tplx_ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    uint64_t *buf = nullptr;  // world * TPLX_MAX_ACCS values (all-gather of partial aggregates) + world * world * 64 sizes
};
static std::vector<std::pair<int, DeviceComm *>> g_comms;  // device id -> communicator

static DeviceComm *comm_of(int device) {
    for (auto &p : g_comms)
        if (p.first == device) return p.second;
    return nullptr;
}

#define NC(call)                                                                                                   \
    do {                                                                                                           \
        int _e = (call);                                                                                           \
        if (_e != 0) return fail(TPLX_E_CUDA, std::string(#call) + ": NCCL " + (N->GetErrorString ? N->GetErrorString(_e) : "error")); \
    } while (0)

extern "C" int32_t tplx_gpu_comm_unique_id(uint8_t *id) {
    NcclApi *N = nccl_api();
    if (!N) return fail(TPLX_E_UNSUPPORTED, "comm_unique_id: NCCL (libnccl.so.2) is not available in this process");
    if (!id) return fail(TPLX_E_BADARG, "comm_unique_id: bad arguments");
    tplx_ncclUniqueId u;
    NC(N->GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return TPLX_OK;
}

static int32_t comm_attach(int device, tplx_ncclComm_t c, int rank, int world) {
    DeviceComm *dc = new DeviceComm();
    dc->comm = c;
    dc->rank = rank;
    dc->world = world;
    CU(cudaSetDevice(device));
    CU(cudaMalloc(&dc->buf, ((size_t)world * TPLX_MAX_ACCS + (size_t)world * world * 64) * 8));
    g_comms.emplace_back(device, dc);
    return TPLX_OK;
}

extern "C" int32_t tplx_gpu_comm_init(int32_t device, int32_t rank, int32_t world, const uint8_t *id) {
    NcclApi *N = nccl_api();
    if (!N) return fail(TPLX_E_UNSUPPORTED, "comm_init: NCCL (libnccl.so.2) is not available in this process");
    Device *d = get_device(device);
    if (!d) return fail(TPLX_E_NODEVICE, "comm_init: device not initialised");
    if (!id || world < 1 || rank < 0 || rank >= world) return fail(TPLX_E_BADARG, "comm_init: bad arguments");
    std::lock_guard<std::mutex> lk(g_mu);
    if (comm_of(device)) return fail(TPLX_E_BADARG, "comm_init: device already has a communicator");
    CU(cudaSetDevice(d->id));
    tplx_ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    tplx_ncclComm_t c = nullptr;
    NC(N->CommInitRank(&c, world, u, rank));
    return comm_attach(device, c, rank, world);
}

extern "C" int32_t tplx_gpu_comm_init_local(const int32_t *devices, int32_t n) {
    NcclApi *N = nccl_api();
    if (!N) return fail(TPLX_E_UNSUPPORTED, "comm_init_local: NCCL (libnccl.so.2) is not available in this process");
    if (!devices || n < 1) return fail(TPLX_E_BADARG, "comm_init_local: bad arguments");
    std::lock_guard<std::mutex> lk(g_mu);
    std::vector<int> devs(devices, devices + n);
    for (int dv : devs) {
        if (!get_device(dv)) return fail(TPLX_E_NODEVICE, "comm_init_local: device not initialised");
        if (comm_of(dv)) return fail(TPLX_E_BADARG, "comm_init_local: device already has a communicator");
    }
    std::vector<tplx_ncclComm_t> comms(n, nullptr);
    NC(N->CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) {
        int32_t rc = comm_attach(devs[i], comms[i], i, n);
        if (rc) return rc;
    }
    return TPLX_OK;
}

extern "C" int32_t tplx_gpu_comm_info(int32_t device, int32_t *rank, int32_t *world) {
    DeviceComm *dc = comm_of(device);
    if (!dc) return fail(TPLX_E_BADARG, "comm_info: device has no communicator");
    if (rank) *rank = dc->rank;
    if (world) *world = dc->world;
    return TPLX_OK;
}

extern "C" int32_t tplx_gpu_comm_destroy(int32_t device) {
    NcclApi *N = nccl_api();
    std::lock_guard<std::mutex> lk(g_mu);
    for (size_t i = 0; i < g_comms.size(); ++i)
        if (g_comms[i].first == device) {
            DeviceComm *dc = g_comms[i].second;
            cudaSetDevice(device);
            if (Device *d = get_device(device)) cudaStreamSynchronize(d->stream);
            if (N && dc->comm) N->CommDestroy(dc->comm);
            cudaFree(dc->buf);
            delete dc;
            g_comms.erase(g_comms.begin() + i);
            return TPLX_OK;
        }
    return TPLX_OK;
}

// fold the gathered partials in rank order: out = g[0] (+) g[1] (+) ... (one thread per accumulator)
__global__ void agg_combine_ranks_kernel(const uint64_t *__restrict__ g, uint32_t world, uint32_t na, const uint32_t *__restrict__ kinds_packed,
                                         uint64_t *__restrict__ out) {
    const uint32_t k = threadIdx.x;
    if (k >= na) return;
    const uint32_t kind = (kinds_packed[k / 4] >> (8 * (k % 4))) & 0xFF;
    uint64_t v = g[k];
    for (uint32_t r = 1; r < world; ++r) v = acc_combine(kind, v, g[(size_t)r * na + k]);
    out[k] = v;
}

extern "C" int32_t tplx_gpu_agg_finish(tplx_stage *s, int32_t device, const int64_t *local_bits, int64_t *out_bits) {
    NcclApi *N = nccl_api();
    Device *d = get_device(device);
    DeviceComm *dc = comm_of(device);
    if (!d) return fail(TPLX_E_NODEVICE, "agg_finish: device not initialised");
    if (!s || !local_bits || !out_bits || s->hdr.endpoint != TPLX_EP_AGGREGATE) return fail(TPLX_E_BADARG, "agg_finish: not an aggregate stage");
    const uint32_t na = (uint32_t)s->accs.size();
    if (!dc || dc->world == 1) {  // a single rank: nothing to combine
        memcpy(out_bits, local_bits, na * 8);
        return TPLX_OK;
    }
    if (!N) return fail(TPLX_E_UNSUPPORTED, "agg_finish: NCCL is not available");
    std::lock_guard<std::mutex> lk(d->mu);
    CU(cudaSetDevice(d->id));
    uint64_t *g = dc->buf;
    uint64_t *out = dc->buf + (size_t)dc->world * TPLX_MAX_ACCS;
    uint32_t *kinds = reinterpret_cast<uint32_t *>(out + TPLX_MAX_ACCS);
    uint32_t hk[TPLX_MAX_ACCS / 4] = {0};
    for (uint32_t k = 0; k < na; ++k) hk[k / 4] |= (uint32_t)s->accs[k].kind << (8 * (k % 4));
    CU(cudaMemcpyAsync(g + (size_t)dc->rank * na, local_bits, na * 8, cudaMemcpyHostToDevice, d->stream));
    CU(cudaMemcpyAsync(kinds, hk, sizeof(hk), cudaMemcpyHostToDevice, d->stream));
    NC(N->AllGather(g + (size_t)dc->rank * na, g, na, NCCL_UINT64, dc->comm, d->stream));
    agg_combine_ranks_kernel<<<1, 32, 0, d->stream>>>(g, (uint32_t)dc->world, na, kinds, out);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_bits, out, na * 8, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    return TPLX_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// aggregateByKey: hash-partitioned exchange on the device (LocalBackend::createFinalHashmap across GPUs)
// ---------------------------------------------------------------------------------------------------------------
extern "C" int32_t tplx_gpu_stage_hash_exchange(tplx_stage *s, int32_t device) {
    NcclApi *N = nccl_api();
    Device *d = get_device(device);
    DeviceComm *dc = comm_of(device);
    if (!d) return fail(TPLX_E_NODEVICE, "hash_exchange: device not initialised");
    if (!s || s->hdr.endpoint != TPLX_EP_HASH) return fail(TPLX_E_BADARG, "hash_exchange: not a hash stage");
    if (!dc || dc->world == 1) return TPLX_OK;  // a single rank owns every key
    if (!N) return fail(TPLX_E_UNSUPPORTED, "hash_exchange: NCCL is not available");
    std::lock_guard<std::mutex> lk(d->mu);
    CU(cudaSetDevice(d->id));
    StageDev *sd = nullptr;
    int32_t rc = stage_dev(s, d, &sd);
    if (rc) return rc;
    if (!sd->ht) {
        rc = hash_grow(d, s, sd, 1024, 1 << 20);
        if (rc) return rc;
    }
    const uint32_t G = (uint32_t)dc->world, me = (uint32_t)dc->rank;
    const uint32_t nk = s->hdr.n_keys, na = (uint32_t)s->accs.size();
    const uint32_t W = 1 + nk;  // sizes per (src, dst): rows, then string bytes of every key column (0 for fixed width)

    // 1. split the table by owner: G packed blocks (key columns + raw partials), all on the device
    std::vector<tplx_result *> parts(G, nullptr);
    auto drop_parts = [&]() {
        for (auto *p : parts)
            if (p) tplx_gpu_result_free(p);
    };
    std::vector<uint64_t> mine((size_t)G * W, 0);
    for (uint32_t g = 0; g < G; ++g) {
        rc = hash_export_locked(s, d, true, HashSel{(int32_t)g, G}, &parts[g]);
        if (rc) { drop_parts(); return rc; }
        mine[(size_t)g * W] = parts[g]->n_out;
        for (uint32_t k = 0; k < nk; ++k) mine[(size_t)g * W + 1 + k] = parts[g]->str_bytes[k];
    }
    // 2. everyone learns every (src, dst) size: one all-gather of G*W values per rank
    uint64_t *dsz = dc->buf + (size_t)G * TPLX_MAX_ACCS + 2 * TPLX_MAX_ACCS;
    std::vector<uint64_t> all((size_t)G * G * W, 0);
    CU(cudaMemcpyAsync(dsz + (size_t)me * G * W, mine.data(), (size_t)G * W * 8, cudaMemcpyHostToDevice, d->stream));
    NC(N->AllGather(dsz + (size_t)me * G * W, dsz, (size_t)G * W, NCCL_UINT64, dc->comm, d->stream));
    CU(cudaMemcpyAsync(all.data(), dsz, all.size() * 8, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    auto sz = [&](uint32_t src, uint32_t dst, uint32_t w) { return all[((size_t)src * G + dst) * W + w]; };

    // 3. receive buffers: per source rank one packed block in the layout hash_merge consumes
    struct Recv { std::vector<void *> data, offs; uint64_t n = 0; };
    std::vector<Recv> recv(G);
    std::vector<void *> owned;
    auto drop_owned = [&]() {
        for (void *p : owned) cudaFreeAsync(p, d->stream);
        owned.clear();
    };
    uint64_t total_keys = 0, total_heap = 0;
    for (uint32_t src = 0; src < G; ++src) {
        const uint64_t n = sz(src, me, 0);
        total_keys += n;
        for (uint32_t k = 0; k < nk; ++k) total_heap += s->out_cols[k].type == TPLX_T_STR ? sz(src, me, 1 + k) + 4 * n : 8 * n;
        if (src == me) continue;
        recv[src].n = n;
        recv[src].data.assign(nk + na, nullptr);
        recv[src].offs.assign(nk + na, nullptr);
        for (uint32_t c = 0; c < nk + na; ++c) {
            const bool str = c < nk && s->out_cols[c].type == TPLX_T_STR;
            void *p = nullptr;
            CU(cudaMallocAsync(&p, align_up(str ? sz(src, me, 1 + c) : n * 8, 16) + 16, d->stream));
            owned.push_back(p);
            recv[src].data[c] = p;
            if (str) {
                void *o = nullptr;
                CU(cudaMallocAsync(&o, (n + 1) * 4 + 16, d->stream));
                owned.push_back(o);
                recv[src].offs[c] = o;
            }
        }
    }
    // 4. the all-to-all: grouped sends / receives straight between device buffers (NVSwitch: every pair at full bandwidth)
    NC(N->GroupStart());
    for (uint32_t p = 0; p < G; ++p) {
        if (p == me) continue;
        const uint64_t ns = sz(me, p, 0), nr = sz(p, me, 0);
        for (uint32_t c = 0; c < nk + na; ++c) {
            const bool str = c < nk && s->out_cols[c].type == TPLX_T_STR;
            const OutCol &oc = parts[p]->out[c];
            if (str) {
                if (ns) {
                    NC(N->Send(oc.offsets, (ns + 1) * 4, NCCL_UINT8, (int)p, dc->comm, d->stream));
                    if (sz(me, p, 1 + c)) NC(N->Send(oc.bytes, sz(me, p, 1 + c), NCCL_UINT8, (int)p, dc->comm, d->stream));
                }
                if (nr) {
                    NC(N->Recv(recv[p].offs[c], (nr + 1) * 4, NCCL_UINT8, (int)p, dc->comm, d->stream));
                    if (sz(p, me, 1 + c)) NC(N->Recv(recv[p].data[c], sz(p, me, 1 + c), NCCL_UINT8, (int)p, dc->comm, d->stream));
                }
            } else {
                if (ns) NC(N->Send(oc.data, ns * 8, NCCL_UINT8, (int)p, dc->comm, d->stream));
                if (nr) NC(N->Recv(recv[p].data[c], nr * 8, NCCL_UINT8, (int)p, dc->comm, d->stream));
            }
        }
    }
    NC(N->GroupEnd());

    // 5. a fresh table for the keys this rank owns: its own share + what the peers sent, merged with the combine operation
    HashTable *old = sd->ht;
    sd->ht = nullptr;
    rc = hash_grow(d, s, sd, total_keys + 16, total_heap + (1 << 16));
    if (rc) { sd->ht = old; drop_parts(); drop_owned(); return rc; }
    auto merge_cols = [&](const std::vector<const void *> &data, const std::vector<const void *> &offs, const std::vector<uint64_t> &bytes, uint64_t n) -> int32_t {
        if (!n) return TPLX_OK;
        tplx_block blk;
        blk.dev = d;
        blk.n_rows = n;
        for (uint32_t c = 0; c < nk + na; ++c) {
            ColIn ci{};
            const bool str = c < nk && s->out_cols[c].type == TPLX_T_STR;
            ci.type = c < nk ? s->out_cols[c].type : (uint64_t)TPLX_T_I64;  // partials are moved as raw 8-byte patterns
            ci.data = data[c];
            ci.offsets = str ? static_cast<const uint32_t *>(offs[c]) : nullptr;
            blk.cols.push_back(ci);
            blk.data_bytes.push_back(str ? bytes[c] : n * 8);
        }
        return hash_merge_locked(s, &blk);
    };
    {
        std::vector<const void *> data(nk + na), offs(nk + na, nullptr);
        std::vector<uint64_t> bytes(nk + na, 0);
        for (uint32_t c = 0; c < nk + na; ++c) {
            const OutCol &oc = parts[me]->out[c];
            const bool str = c < nk && s->out_cols[c].type == TPLX_T_STR;
            data[c] = str ? static_cast<const void *>(oc.bytes) : static_cast<const void *>(oc.data);
            offs[c] = oc.offsets;
            bytes[c] = str ? parts[me]->str_bytes[c] : 0;
        }
        rc = merge_cols(data, offs, bytes, parts[me]->n_out);
    }
    for (uint32_t src = 0; src < G && !rc; ++src) {
        if (src == me || !recv[src].n) continue;
        std::vector<const void *> data(recv[src].data.begin(), recv[src].data.end()), offs(recv[src].offs.begin(), recv[src].offs.end());
        std::vector<uint64_t> bytes(nk + na, 0);
        for (uint32_t k = 0; k < nk; ++k) bytes[k] = sz(src, me, 1 + k);
        rc = merge_cols(data, offs, bytes, recv[src].n);
    }
    CU(cudaStreamSynchronize(d->stream));
    hash_table_destroy(old);
    drop_parts();
    drop_owned();
    CU(cudaStreamSynchronize(d->stream));
    return rc;
}
