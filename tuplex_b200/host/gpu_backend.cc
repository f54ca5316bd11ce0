// gpu_backend.cc — see gpu_backend.h. Links against libtplx_gpu.so only (no CUDA headers needed here).
#include "gpu_backend.h"

namespace tuplex_b200 {

void GpuBackend::check(int32_t rc, const char *what) {
    if (rc != TPLX_OK) throw std::runtime_error(std::string(what) + ": " + tplx_gpu_last_error());
}

GpuBackend::GpuBackend(const std::vector<int32_t> &devices) : _devices(devices) {
    check(tplx_gpu_init(_devices.data(), (int32_t)_devices.size()), "tplx_gpu_init");  // throws when no GPU: no CPU fallback
}

GpuBackend::~GpuBackend() = default;

void GpuBackend::execute(GpuTransformStage &st) {
    tplx_stage *stage = nullptr;
    check(tplx_gpu_stage_create(st.descriptor.data(), st.descriptor.size(), &stage), "tplx_gpu_stage_create");
    tplx_block *block = nullptr;
    tplx_result *res = nullptr;
    try {
        // one task = all partitions of the stage in order (row numbers are not reset between input partitions of a
        // task, TransformTask.cc:885); K5 transposes them into a column block on the device
        std::vector<const uint8_t *> ptrs;
        std::vector<uint64_t> sizes;
        for (auto &p : st.inputPartitions) {
            ptrs.push_back(p.data());
            sizes.push_back(p.size());
        }
        check(tplx_gpu_block_from_partitions(device(), ptrs.data(), sizes.data(), (uint32_t)ptrs.size(), st.inputColumnTypes.data(),
                                             (uint32_t)st.inputColumnTypes.size(), &block), "tplx_gpu_block_from_partitions");
        check(tplx_gpu_stage_run(stage, block, 0, &res), "tplx_gpu_stage_run");
        tplx_result_info info;
        check(tplx_gpu_result_info(res, &info), "tplx_gpu_result_info");
        st.numOutputRows = info.n_out_rows;
        st.numExceptionRows = info.n_exceptions;
        st.kernelMs = info.kernel_ms;
        // normal-case output in Partition format, split like rowToMemorySink (TransformTask.h:47-92)
        uint64_t need = 0;
        uint32_t np = 0;
        check(tplx_gpu_result_partitions(res, st.partitionSize, nullptr, 0, &need, nullptr, 0, &np), "tplx_gpu_result_partitions(size)");
        std::vector<uint8_t> buf(need);
        std::vector<uint64_t> offs(np + 1);
        check(tplx_gpu_result_partitions(res, st.partitionSize, buf.data(), buf.size(), &need, offs.data(), np, &np), "tplx_gpu_result_partitions");
        st.outputPartitions.clear();
        for (uint32_t p = 0; p < np; ++p) st.outputPartitions.emplace_back(buf.begin() + offs[p], buf.begin() + offs[p + 1]);
        // exception rows for the unchanged resolve path (ResolveTask consumes exactly this format)
        check(tplx_gpu_result_exception_partition(res, nullptr, 0, &need), "tplx_gpu_result_exception_partition(size)");
        st.exceptionPartition.resize(need);
        check(tplx_gpu_result_exception_partition(res, st.exceptionPartition.data(), need, &need), "tplx_gpu_result_exception_partition");
    } catch (...) {
        if (res) tplx_gpu_result_free(res);
        if (block) tplx_gpu_block_free(block);
        tplx_gpu_stage_destroy(stage);
        throw;
    }
    tplx_gpu_result_free(res);
    tplx_gpu_block_free(block);
    tplx_gpu_stage_destroy(stage);
}

}  // namespace tuplex_b200
