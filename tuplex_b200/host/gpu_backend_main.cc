// gpu_backend_main.cc — tiny driver around GpuBackend::execute, used by tests/test_gpu_cpp_host.py to exercise the
// C++ host mirror end to end in the reference's own language: files in (stage descriptor, reference-format input
// partitions), files out (output partitions, exception partition). Plays the part of TransformStage::execute ->
// backend()->execute(this) (tuplex/core/src/physical/TransformStage.cc:610-700).
//
//   tplx_host_run <descriptor.bin> <coltypes: e.g. 0,3,1> <partition_size> <out_prefix> <part0.bin> [part1.bin ...]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

#include "gpu_backend.h"

static std::vector<uint8_t> slurp(const char *path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static void dump(const std::string &path, const std::vector<uint8_t> &b) {
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char *>(b.data()), (std::streamsize)b.size());
}

int main(int argc, char **argv) {
    if (argc < 6) {
        std::fprintf(stderr, "usage: %s <descriptor.bin> <coltypes> <partition_size> <out_prefix> <part.bin>...\n", argv[0]);
        return 2;
    }
    try {
        tuplex_b200::GpuTransformStage st;
        st.descriptor = slurp(argv[1]);
        std::stringstream ss(argv[2]);
        for (std::string tok; std::getline(ss, tok, ',');) st.inputColumnTypes.push_back((uint8_t)std::atoi(tok.c_str()));
        st.partitionSize = std::strtoull(argv[3], nullptr, 10);
        const std::string prefix = argv[4];
        for (int i = 5; i < argc; ++i) st.inputPartitions.push_back(slurp(argv[i]));
        tuplex_b200::GpuBackend backend({0});
        backend.execute(st);
        for (size_t p = 0; p < st.outputPartitions.size(); ++p) dump(prefix + ".out" + std::to_string(p), st.outputPartitions[p]);
        dump(prefix + ".exc", st.exceptionPartition);
        std::printf("{\"out_rows\": %llu, \"exceptions\": %llu, \"out_partitions\": %zu, \"kernel_ms\": %.3f}\n",
                    (unsigned long long)st.numOutputRows, (unsigned long long)st.numExceptionRows, st.outputPartitions.size(), st.kernelMs);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());  // stage-level failure (LocalBackend throws std::runtime_error too)
        return 1;
    }
    return 0;
}
