// gpu_backend.cc — see gpu_backend.h. Links against libtplx_gpu.so only (no CUDA headers needed here).
#include "gpu_backend.h"

#include <algorithm>
#include <cstring>
#include <exception>
#include <mutex>
#include <set>
#include <thread>

namespace tuplex_b200 {

void GpuBackend::check(int32_t rc, const char *what) {
    if (rc != TPLX_OK) throw std::runtime_error(std::string(what) + ": " + tplx_gpu_last_error());
}

uint8_t GpuTransformStage::endpoint() const {
    if (descriptor.size() < sizeof(tplx_stage_header)) throw std::runtime_error("stage descriptor: truncated header");
    tplx_stage_header h;
    std::memcpy(&h, descriptor.data(), sizeof(h));
    return h.endpoint;
}

GpuBackend::GpuBackend(const std::vector<int32_t> &devices) : _devices(devices.empty() ? std::vector<int32_t>{0} : devices) {
    std::set<int32_t> uniq(_devices.begin(), _devices.end());
    std::vector<int32_t> u(uniq.begin(), uniq.end());
    check(tplx_gpu_init(u.data(), (int32_t)u.size()), "tplx_gpu_init");  // throws when no GPU: no CPU fallback
    if (uniq.size() == _devices.size() && _devices.size() > 1) {
        int32_t r = 0, w = 0;
        if (tplx_gpu_comm_info(_devices[0], &r, &w) != TPLX_OK)  // one communicator per process and device set
            check(tplx_gpu_comm_init_local(_devices.data(), (int32_t)_devices.size()), "tplx_gpu_comm_init_local");
        _comm = true;
    }
}

GpuBackend::~GpuBackend() = default;

namespace {
// accumulator kinds of the stage (tplx_acc section of the descriptor, include/tplx_ir.h)
std::vector<uint8_t> acc_kinds(const std::vector<uint8_t> &desc) {
    tplx_stage_header h;
    std::memcpy(&h, desc.data(), sizeof(h));
    auto pad8 = [](size_t v) { return (v + 7) / 8 * 8; };
    size_t off = sizeof(h) + pad8(h.n_in_cols) + pad8((size_t)h.n_out_cols * sizeof(tplx_outcol));
    std::vector<uint8_t> kinds;
    for (uint32_t k = 0; k < h.n_accs; ++k) {
        tplx_acc a;
        std::memcpy(&a, desc.data() + off + k * sizeof(tplx_acc), sizeof(a));
        kinds.push_back(a.kind);
    }
    return kinds;
}
int64_t combine(uint8_t kind, int64_t a, int64_t b) {
    auto f = [](int64_t v) { double d; std::memcpy(&d, &v, 8); return d; };
    auto u = [](double d) { int64_t v; std::memcpy(&v, &d, 8); return v; };
    switch (kind) {
        case TPLX_ACC_SUM_I64: return (int64_t)((uint64_t)a + (uint64_t)b);
        case TPLX_ACC_SUM_F64: return u(f(a) + f(b));
        case TPLX_ACC_MIN_I64: return b < a ? b : a;
        case TPLX_ACC_MAX_I64: return b > a ? b : a;
        case TPLX_ACC_MIN_F64: return f(b) < f(a) ? b : a;
        default: return f(b) > f(a) ? b : a;
    }
}
// exception partition = int64 numRows, then [rowNo, ecCode, opID, size, row bytes] records (IExceptionableTask.h:22-36)
void append_exceptions(std::vector<uint8_t> &dst, const std::vector<uint8_t> &src, int64_t row_no_offset) {
    if (src.size() < 8) return;
    int64_t n = 0;
    std::memcpy(&n, src.data(), 8);
    if (dst.empty()) dst.assign(8, 0);
    size_t p = 8;
    for (int64_t i = 0; i < n; ++i) {
        int64_t hdr[4];
        std::memcpy(hdr, src.data() + p, 32);
        hdr[0] += row_no_offset;
        const size_t rec = 32 + (size_t)hdr[3];
        const size_t at = dst.size();
        dst.resize(at + rec);
        std::memcpy(dst.data() + at, hdr, 32);
        std::memcpy(dst.data() + at + 32, src.data() + p + 32, (size_t)hdr[3]);
        p += rec;
    }
    int64_t total = 0;
    std::memcpy(&total, dst.data(), 8);
    total += n;
    std::memcpy(dst.data(), &total, 8);
}
std::vector<std::vector<uint8_t>> fetch_partitions(tplx_result *res, uint64_t psize, const char *what) {
    uint64_t need = 0;
    uint32_t np = 0;
    if (tplx_gpu_result_partitions(res, psize, nullptr, 0, &need, nullptr, 0, &np) != TPLX_OK) throw std::runtime_error(std::string(what) + ": " + tplx_gpu_last_error());
    std::vector<uint8_t> buf(std::max<uint64_t>(need, 8));
    std::vector<uint64_t> offs(np + 1);
    if (tplx_gpu_result_partitions(res, psize, buf.data(), buf.size(), &need, offs.data(), np, &np) != TPLX_OK)
        throw std::runtime_error(std::string(what) + ": " + tplx_gpu_last_error());
    std::vector<std::vector<uint8_t>> out;
    for (uint32_t p = 0; p < np; ++p) out.emplace_back(buf.begin() + offs[p], buf.begin() + offs[p + 1]);
    return out;
}
}  // namespace

void GpuBackend::execute(GpuTransformStage &st) {
    tplx_stage *stage = nullptr;
    check(tplx_gpu_stage_create(st.descriptor.data(), st.descriptor.size(), &stage), "tplx_gpu_stage_create");
    const uint8_t ep = st.endpoint();
    const uint32_t T = (uint32_t)_devices.size();
    const uint32_t np = (uint32_t)st.inputPartitions.size();
    st.tasks = T;

    struct Task {
        std::vector<std::vector<uint8_t>> out;
        std::vector<uint8_t> exc;
        std::vector<int64_t> agg;
        uint64_t n_out = 0, n_exc = 0;
        double kernel_ms = 0;
        std::exception_ptr error;
    };
    std::vector<Task> tasks(T);
    const std::vector<uint8_t> kinds = ep == TPLX_EP_AGGREGATE ? acc_kinds(st.descriptor) : std::vector<uint8_t>{};

    // one task = a contiguous run of partitions on one device, in order (row numbers are not reset between the input partitions
    // of a task, TransformTask.cc:885); K5 transposes them into a column block on the device
    auto run_task = [&](uint32_t t) {
        Task &tk = tasks[t];
        tplx_block *block = nullptr;
        tplx_result *res = nullptr;
        try {
            const uint32_t base = np / T, rem = np % T;
            const uint32_t lo = t * base + std::min(t, rem), hi = lo + base + (t < rem ? 1 : 0);
            std::vector<const uint8_t *> ptrs;
            std::vector<uint64_t> sizes;
            for (uint32_t p = lo; p < hi; ++p) {
                ptrs.push_back(st.inputPartitions[p].data());
                sizes.push_back(st.inputPartitions[p].size());
            }
            check(tplx_gpu_block_from_partitions(_devices[t], ptrs.data(), sizes.data(), (uint32_t)ptrs.size(), st.inputColumnTypes.data(),
                                                 (uint32_t)st.inputColumnTypes.size(), &block), "tplx_gpu_block_from_partitions");
            check(tplx_gpu_stage_run(stage, block, 0, &res), "tplx_gpu_stage_run");
            tplx_result_info info;
            check(tplx_gpu_result_info(res, &info), "tplx_gpu_result_info");
            tk.n_out = info.n_out_rows;
            tk.n_exc = info.n_exceptions;
            tk.kernel_ms = info.kernel_ms;
            if (ep == TPLX_EP_MEMORY) tk.out = fetch_partitions(res, st.partitionSize, "tplx_gpu_result_partitions");
            if (ep == TPLX_EP_AGGREGATE) {
                tk.agg.resize(kinds.size());
                check(tplx_gpu_result_fetch_aggregate(res, tk.agg.data()), "tplx_gpu_result_fetch_aggregate");
                if (_comm) {  // collective over the devices: every task ends up with the combined bits
                    std::vector<int64_t> all(kinds.size());
                    check(tplx_gpu_agg_finish(stage, _devices[t], tk.agg.data(), all.data()), "tplx_gpu_agg_finish");
                    tk.agg = all;
                }
            }
            uint64_t need = 0;
            check(tplx_gpu_result_exception_partition(res, nullptr, 0, &need), "tplx_gpu_result_exception_partition(size)");
            tk.exc.resize(need);
            check(tplx_gpu_result_exception_partition(res, tk.exc.data(), need, &need), "tplx_gpu_result_exception_partition");
            if (tk.n_exc && _onExceptions) _onExceptions(t, tk.exc);
            if (ep == TPLX_EP_HASH && _comm) check(tplx_gpu_stage_hash_exchange(stage, _devices[t]), "tplx_gpu_stage_hash_exchange");
        } catch (...) {
            tk.error = std::current_exception();
        }
        if (res) tplx_gpu_result_free(res);
        if (block) tplx_gpu_block_free(block);
    };
    if (T == 1) run_task(0);
    else {
        std::vector<std::thread> th;
        for (uint32_t t = 0; t < T; ++t) th.emplace_back(run_task, t);
        for (auto &x : th) x.join();
    }
    for (auto &tk : tasks)
        if (tk.error) {
            tplx_gpu_stage_destroy(stage);
            std::rethrow_exception(tk.error);
        }

    // ---- assemble the stage result in task order (LocalBackend.cc:1104-1152) ---------------------------------------------------
    st.outputPartitions.clear();
    st.exceptionPartition.clear();
    st.hashPartitions.clear();
    st.aggregate.clear();
    st.numOutputRows = st.numExceptionRows = 0;
    st.kernelMs = 0;
    int64_t row_no = 0;
    for (auto &tk : tasks) {
        for (auto &p : tk.out) st.outputPartitions.push_back(std::move(p));
        append_exceptions(st.exceptionPartition, tk.exc, row_no);
        row_no += (int64_t)(tk.n_out + tk.n_exc);
        st.numOutputRows += tk.n_out;
        st.numExceptionRows += tk.n_exc;
        st.kernelMs += tk.kernel_ms;
    }
    if (st.exceptionPartition.empty()) st.exceptionPartition.assign(8, 0);
    try {
        if (ep == TPLX_EP_AGGREGATE) {
            if (_comm) st.aggregate = tasks[0].agg;  // already combined on the devices
            else {
                st.aggregate = tasks[0].agg;
                for (uint32_t t = 1; t < T; ++t)
                    for (size_t k = 0; k < kinds.size(); ++k) st.aggregate[k] = combine(kinds[k], st.aggregate[k], tasks[t].agg[k]);
            }
            st.numOutputRows = 1;
        }
        if (ep == TPLX_EP_HASH) {
            std::set<int32_t> seen;
            st.numOutputRows = 0;
            for (int32_t dv : _devices) {
                if (!seen.insert(dv).second) continue;  // one table per device
                tplx_result *fin = nullptr;
                check(tplx_gpu_stage_hash_finish(stage, dv, &fin), "tplx_gpu_stage_hash_finish");
                tplx_result_info info;
                int32_t rc = tplx_gpu_result_info(fin, &info);
                std::vector<std::vector<uint8_t>> parts;
                if (rc == TPLX_OK) {
                    try {
                        parts = fetch_partitions(fin, st.partitionSize, "tplx_gpu_result_partitions(hash)");
                    } catch (...) {
                        tplx_gpu_result_free(fin);
                        throw;
                    }
                }
                tplx_gpu_result_free(fin);
                check(rc, "tplx_gpu_result_info(hash)");
                st.numOutputRows += info.n_out_rows;
                for (auto &p : parts) st.hashPartitions.push_back(std::move(p));
            }
        }
    } catch (...) {
        tplx_gpu_stage_destroy(stage);
        throw;
    }
    tplx_gpu_stage_destroy(stage);
}

void GpuBackend::execute(GpuHashJoinStage &st) {
    const uint32_t T = (uint32_t)_devices.size();
    const uint32_t np = (uint32_t)st.probePartitions.size();
    st.tasks = T;
    struct Task {
        std::vector<std::vector<uint8_t>> out;
        uint64_t n_out = 0;
        double kernel_ms = 0, build_ms = 0;
        std::exception_ptr error;
    };
    std::vector<Task> tasks(T);
    const uint32_t flags = (st.leftOuter ? TPLX_JOIN_LEFT_OUTER : 0) | (st.buildFirst ? TPLX_JOIN_BUILD_FIRST : 0);
    auto run_task = [&](uint32_t t) {
        Task &tk = tasks[t];
        tplx_block *bb = nullptr, *pb = nullptr;
        tplx_join *jn = nullptr;
        tplx_result *res = nullptr;
        try {
            std::vector<const uint8_t *> ptrs;
            std::vector<uint64_t> sizes;
            for (auto &p : st.buildPartitions) {
                ptrs.push_back(p.data());
                sizes.push_back(p.size());
            }
            check(tplx_gpu_block_from_partitions(_devices[t], ptrs.data(), sizes.data(), (uint32_t)ptrs.size(), st.buildColumnTypes.data(),
                                                 (uint32_t)st.buildColumnTypes.size(), &bb), "tplx_gpu_block_from_partitions(build)");
            check(tplx_gpu_join_build(bb, st.buildKey, &jn), "tplx_gpu_join_build");
            check(tplx_gpu_join_info(jn, nullptr, nullptr, &tk.build_ms, nullptr), "tplx_gpu_join_info");
            const uint32_t base = np / T, rem = np % T;
            const uint32_t lo = t * base + std::min(t, rem), hi = lo + base + (t < rem ? 1 : 0);
            ptrs.clear();
            sizes.clear();
            for (uint32_t p = lo; p < hi; ++p) {
                ptrs.push_back(st.probePartitions[p].data());
                sizes.push_back(st.probePartitions[p].size());
            }
            if (!ptrs.empty()) {
                check(tplx_gpu_block_from_partitions(_devices[t], ptrs.data(), sizes.data(), (uint32_t)ptrs.size(), st.probeColumnTypes.data(),
                                                     (uint32_t)st.probeColumnTypes.size(), &pb), "tplx_gpu_block_from_partitions(probe)");
                check(tplx_gpu_join_probe(jn, pb, st.probeKey, flags, &res), "tplx_gpu_join_probe");
                tplx_result_info info;
                check(tplx_gpu_result_info(res, &info), "tplx_gpu_result_info");
                tk.n_out = info.n_out_rows;
                tk.kernel_ms = info.kernel_ms;
                tk.out = fetch_partitions(res, st.partitionSize, "tplx_gpu_result_partitions(join)");
            }
        } catch (...) {
            tk.error = std::current_exception();
        }
        if (res) tplx_gpu_result_free(res);
        if (jn) tplx_gpu_join_destroy(jn);
        if (pb) tplx_gpu_block_free(pb);
        if (bb) tplx_gpu_block_free(bb);
    };
    if (T == 1) run_task(0);
    else {
        std::vector<std::thread> th;
        for (uint32_t t = 0; t < T; ++t) th.emplace_back(run_task, t);
        for (auto &x : th) x.join();
    }
    for (auto &tk : tasks)
        if (tk.error) std::rethrow_exception(tk.error);
    st.outputPartitions.clear();
    st.numOutputRows = 0;
    st.kernelMs = st.buildMs = 0;
    for (auto &tk : tasks) {
        for (auto &p : tk.out) st.outputPartitions.push_back(std::move(p));
        st.numOutputRows += tk.n_out;
        st.kernelMs += tk.kernel_ms;
        st.buildMs += tk.build_ms;
    }
}

}  // namespace tuplex_b200
