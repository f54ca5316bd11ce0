// gpu_backend.h — C++ host mirror of the reference's backend plug-in point, written against the C ABI only.
//
// In the reference, Context::Context switches on ContextOptions::BACKEND() and constructs an IBackend
// (tuplex/core/src/Context.cc:56-83; interface tuplex/core/include/ee/IBackend.h:29-46:
//   virtual Executor* driver(); virtual void execute(PhysicalStage*);).
// TransformStage::execute calls backend()->execute(this) (tuplex/core/src/physical/TransformStage.cc:700) and reads
// inputPartitions()/normalCaseInputSchema() and writes setMemoryResult(...) (TransformStage.h:74-244,186-205).
// GpuBackend::execute does the same job for a GpuTransformStage: the accessors below carry exactly the data the
// LocalBackend reads from a TransformStage, plus the stage descriptor that replaces the LLVM bitcode.
// Stage-level failures throw std::runtime_error like LocalBackend (LocalBackend.cc:896,908,1184,1212); row-level
// errors never throw — they come back as exception partitions.
#pragma once
#include <cstdint>
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
    // results (setMemoryResult)
    std::vector<std::vector<uint8_t>> outputPartitions;
    std::vector<uint8_t> exceptionPartition;         // [numRows][rowNo, ecCode, opID, size, row]...
    std::vector<int64_t> aggregate;                  // AGG_GENERAL: one 8-byte value per accumulator
    uint64_t numOutputRows = 0, numExceptionRows = 0;
    double kernelMs = 0;
};

class GpuBackend {
public:
    explicit GpuBackend(const std::vector<int32_t> &devices = {0});
    ~GpuBackend();
    // IBackend::execute(PhysicalStage*): run the normal case of one stage over all its input partitions
    void execute(GpuTransformStage &stage);
    int32_t device() const { return _devices.empty() ? 0 : _devices[0]; }

private:
    std::vector<int32_t> _devices;
    static void check(int32_t rc, const char *what);
};

}  // namespace tuplex_b200
