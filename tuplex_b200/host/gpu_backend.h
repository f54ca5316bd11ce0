// gpu_backend.h — C++ host of the GPU executor, written against the C ABI only (include/tplx_gpu.h).
//
// In the reference, Context::Context switches on ContextOptions::BACKEND() and constructs an IBackend
// (tuplex/core/src/Context.cc:56-83; interface tuplex/core/include/ee/IBackend.h:29-46:
//   virtual Executor* driver(); virtual void execute(PhysicalStage*);).
// TransformStage::execute calls backend()->execute(this) (tuplex/core/src/physical/TransformStage.cc:700), the backend reads
// inputPartitions() / normalCaseInputSchema() / outputMode() and writes setMemoryResult(...) or setHashResult(...)
// (TransformStage.h:74-244,186-205,382-389). GpuBackend::execute does LocalBackend::executeTransformStage's job
// (tuplex/core/src/ee/local/LocalBackend.cc:815-1252) for a GpuTransformStage:
//   * tasks: the input partitions are split into contiguous groups, one task per device (LocalBackend.cc:679-735 creates one
//     task per partition group; :1531-1586 runs them on the executors), each task on its own host thread;
//   * memory endpoint: per-task output / exception partitions concatenated in task order, exception row numbers continued
//     across tasks (LocalBackend.cc:1104-1152);
//   * aggregate endpoint: per-task partials combined by tplx_gpu_agg_finish (NCCL over the devices) or, on one device, in task
//     order on the host (TransformTask.cc:278-299);
//   * hash endpoint: per-device tables exchanged by tplx_gpu_stage_hash_exchange, then materialised as partitions of
//     (key columns, aggregate columns) rows (TransformStage.cc:473-528,568-608);
//   * exception rows are handed to a resolve callback per task (the reference schedules ResolveTasks, LocalBackend.cc:1254-1400).
// Stage-level failures throw std::runtime_error like LocalBackend (LocalBackend.cc:896,908,1184,1212); row-level errors never
// throw — they come back as exception partitions.
#pragma once
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/tplx_gpu.h"

namespace tuplex_b200 {

// what a TransformStage hands to its backend (subset used on the hot path)
struct GpuTransformStage {
    std::vector<uint8_t> descriptor;                 // serialized tplx_stage_header + sections (replaces bitcode)
    std::vector<uint8_t> inputColumnTypes;           // normalCaseInputSchema() as tplx_type per column
    std::vector<std::vector<uint8_t>> inputPartitions;  // reference-format partitions: int64 numRows + rows
    uint64_t partitionSize = 32ull << 20;            // tuplex.partitionSize
    // ---- results -------------------------------------------------------------------------------------------
    // setMemoryResult: normal-case rows and exception rows of the whole stage, in input order
    std::vector<std::vector<uint8_t>> outputPartitions;
    std::vector<uint8_t> exceptionPartition;         // [numRows][rowNo, ecCode, opID, size, row]...
    // AGG_GENERAL: one 8-byte value per accumulator, combined over all tasks (raw bits: i64 or f64)
    std::vector<int64_t> aggregate;
    // setHashResult (AGG_BY_KEY / unique): the groups as rows (key columns, aggregate columns) in Partition format
    std::vector<std::vector<uint8_t>> hashPartitions;
    uint64_t numOutputRows = 0, numExceptionRows = 0;
    double kernelMs = 0;
    uint32_t tasks = 0;

    uint8_t endpoint() const;  // tplx_endpoint read from the descriptor header
};

// a JoinOperator between two executed stages (HashJoinStage, tuplex/core/include/physical/HashJoinStage.h): the build side's rows
// and the probe side's rows as reference-format partitions, key column per side, result as partitions
struct GpuHashJoinStage {
    std::vector<uint8_t> probeColumnTypes, buildColumnTypes;  // tplx_type per column, | TPLX_T_OPTION for Option[T] fields
    std::vector<std::vector<uint8_t>> probePartitions, buildPartitions;
    uint32_t probeKey = 0, buildKey = 0;
    bool leftOuter = false;   // JoinType::LEFT: unmatched probe rows are emitted with None build columns
    bool buildFirst = false;  // the build side is the LEFT dataset (JoinOperator::buildRight() == false): its columns come first
    uint64_t partitionSize = 32ull << 20;
    // ---- result ----
    std::vector<std::vector<uint8_t>> outputPartitions;  // | left non-key | key | right non-key | rows, probe order
    uint64_t numOutputRows = 0;
    double kernelMs = 0, buildMs = 0;
    uint32_t tasks = 0;
};

class GpuBackend {
public:
    // devices: one task per entry. Distinct devices get an NCCL communicator (tplx_gpu_comm_init_local); the same device may be
    // listed several times (several tasks on one GPU; combines then happen in task order on the host / in the one device table).
    explicit GpuBackend(const std::vector<int32_t> &devices = {0});
    ~GpuBackend();
    // IBackend::execute(PhysicalStage*): run the normal case of one stage over all its input partitions
    void execute(GpuTransformStage &stage);
    // HashJoinStage::execute's job on the GPU (K8): every task builds the table on its device from all build partitions
    // (broadcast of the small side, like the reference hands one hash map to every task through init_stage_f, CodeDefs.h:94) and
    // probes its contiguous run of probe partitions; outputs are concatenated in task order = probe order
    void execute(GpuHashJoinStage &stage);
    // called once per task that produced exception rows, with that task's exception partition (row numbers local to the task),
    // from the task's thread — the hook where the reference schedules its ResolveTask
    void setExceptionHandler(std::function<void(uint32_t task, const std::vector<uint8_t> &exceptionPartition)> fn) { _onExceptions = std::move(fn); }
    int32_t device() const { return _devices.empty() ? 0 : _devices[0]; }
    const std::vector<int32_t> &devices() const { return _devices; }

private:
    std::vector<int32_t> _devices;
    bool _comm = false;  // distinct devices joined in a communicator
    std::function<void(uint32_t, const std::vector<uint8_t> &)> _onExceptions;
    static void check(int32_t rc, const char *what);
};

}  // namespace tuplex_b200
