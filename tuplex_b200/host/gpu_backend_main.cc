// gpu_backend_main.cc — tiny driver around GpuBackend::execute, used by tests/test_gpu_cpp_host.py to exercise the
// C++ host mirror end to end in the reference's own language: files in (stage descriptor, reference-format input
// partitions), files out (output partitions, exception partition). Plays the part of TransformStage::execute ->
// backend()->execute(this) (tuplex/core/src/physical/TransformStage.cc:610-700).
//
//   tplx_host_run <descriptor.bin> <coltypes: e.g. 0,3,1> <partition_size> <out_prefix> [--devices 0,1] <part0.bin> [part1.bin ...]
// --devices: one task per entry (the same device may repeat: several tasks on one GPU). Outputs: <prefix>.out<i> (memory endpoint),
// <prefix>.exc, <prefix>.agg (aggregate endpoint: raw 8-byte values), <prefix>.hash<i> (hash endpoint: groups as rows).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
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

// tplx_host_run --join <probe coltypes> <probe key> <build coltypes> <build key> <flags: 1 left outer, 2 build first> <partition_size>
//               <out_prefix> <devices: 0 or 0,1,..> <n_build_parts> <build part files...> <probe part files...>
static int main_join(int argc, char **argv) {
    if (argc < 12) {
        std::fprintf(stderr, "usage: %s --join <probe types> <probe key> <build types> <build key> <flags> <partition_size> <out_prefix> <devices> <n_build> <files...>\n", argv[0]);
        return 2;
    }
    try {
        tuplex_b200::GpuHashJoinStage st;
        auto types = [](const char *a) {
            std::vector<uint8_t> v;
            std::stringstream ss(a);
            for (std::string tok; std::getline(ss, tok, ',');) v.push_back((uint8_t)std::atoi(tok.c_str()));
            return v;
        };
        st.probeColumnTypes = types(argv[2]);
        st.probeKey = (uint32_t)std::atoi(argv[3]);
        st.buildColumnTypes = types(argv[4]);
        st.buildKey = (uint32_t)std::atoi(argv[5]);
        const int flags = std::atoi(argv[6]);
        st.leftOuter = flags & 1;
        st.buildFirst = flags & 2;
        st.partitionSize = std::strtoull(argv[7], nullptr, 10);
        const std::string prefix = argv[8];
        std::vector<int32_t> devices;
        for (uint8_t d : types(argv[9])) devices.push_back(d);
        const int nb = std::atoi(argv[10]);
        for (int i = 11; i < argc; ++i) (i < 11 + nb ? st.buildPartitions : st.probePartitions).push_back(slurp(argv[i]));
        tuplex_b200::GpuBackend backend(devices);
        backend.execute(st);
        for (size_t p = 0; p < st.outputPartitions.size(); ++p) dump(prefix + ".out" + std::to_string(p), st.outputPartitions[p]);
        std::printf("{\"out_rows\": %llu, \"out_partitions\": %zu, \"tasks\": %u, \"kernel_ms\": %.3f, \"build_ms\": %.3f}\n",
                    (unsigned long long)st.numOutputRows, st.outputPartitions.size(), st.tasks, st.kernelMs, st.buildMs);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}

int main(int argc, char **argv) {
    if (argc > 1 && std::string(argv[1]) == "--join") return main_join(argc, argv);
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
        std::vector<int32_t> devices{0};
        int first = 5;
        if (argc > 6 && std::string(argv[5]) == "--devices") {
            devices.clear();
            std::stringstream ds(argv[6]);
            for (std::string tok; std::getline(ds, tok, ',');) devices.push_back(std::atoi(tok.c_str()));
            first = 7;
        }
        for (int i = first; i < argc; ++i) st.inputPartitions.push_back(slurp(argv[i]));
        tuplex_b200::GpuBackend backend(devices);
        unsigned long long handed = 0;
        std::mutex mu;
        backend.setExceptionHandler([&](uint32_t, const std::vector<uint8_t> &part) {  // the resolve hook: count what a ResolveTask would get
            long long n = 0;
            std::memcpy(&n, part.data(), 8);
            std::lock_guard<std::mutex> lk(mu);
            handed += (unsigned long long)n;
        });
        backend.execute(st);
        for (size_t p = 0; p < st.outputPartitions.size(); ++p) dump(prefix + ".out" + std::to_string(p), st.outputPartitions[p]);
        for (size_t p = 0; p < st.hashPartitions.size(); ++p) dump(prefix + ".hash" + std::to_string(p), st.hashPartitions[p]);
        dump(prefix + ".exc", st.exceptionPartition);
        std::vector<uint8_t> agg(st.aggregate.size() * 8);
        if (!agg.empty()) std::memcpy(agg.data(), st.aggregate.data(), agg.size());
        dump(prefix + ".agg", agg);
        std::printf("{\"out_rows\": %llu, \"exceptions\": %llu, \"out_partitions\": %zu, \"hash_partitions\": %zu, \"n_aggregate\": %zu, "
                    "\"tasks\": %u, \"endpoint\": %u, \"exceptions_handed_to_resolve\": %llu, \"kernel_ms\": %.3f}\n",
                    (unsigned long long)st.numOutputRows, (unsigned long long)st.numExceptionRows, st.outputPartitions.size(), st.hashPartitions.size(),
                    st.aggregate.size(), st.tasks, (unsigned)st.endpoint(), handed, st.kernelMs);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());  // stage-level failure (LocalBackend throws std::runtime_error too)
        return 1;
    }
    return 0;
}
