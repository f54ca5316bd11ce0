// Reference cell splitter for the CSV oracle — TEST INFRASTRUCTURE ONLY.
//
// Drives the reference's own csvmonkey reader (compiled from where it lies:
// /root/reference/tuplex/core/include/physical/csvmonkey.h, passed with -I by oracle/Makefile; nothing of it is copied
// here) exactly as CSVReader::read does (core/src/physical/CSVReader.cc:414-417: CsvReader<>(cursor, delimiter,
// quotechar), yield_incomplete_row = false) over an in-memory cursor that ends the data with the '\n' the
// reference's VFCSVStreamCursor appends (CSVReader.cc:94-100,167-177). Output on stdout: per row
// [u32 count] ([u32 len] dequoted bytes)*  — the same dump oracle/csv_oracle.c produces with dump_cells = 1.
//
// usage: csv_ref <file> [delimiter] [quotechar]
#include <cstdio>
#include <cstdint>
#include <string>
#include <vector>
#include "csvmonkey.h"

struct MemCursor : public csvmonkey::StreamCursor {
    std::vector<char> data;
    size_t pos = 0, len = 0;
    const char *buf() override { return data.data() + pos; }
    size_t size() override { return len - pos; }
    void consume(size_t n) override { pos += std::min(n, len - pos); }
    bool fill() override { return false; }
};

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    MemCursor cur;
    char tmp[1 << 16];
    size_t k;
    while ((k = fread(tmp, 1, sizeof tmp, f)) > 0) cur.data.insert(cur.data.end(), tmp, tmp + k);
    fclose(f);
    cur.data.push_back('\n');
    cur.len = cur.data.size();
    cur.data.insert(cur.data.end(), 64, '\0');  // PCMPISTRI guard (csvmonkey.h:77-81)
    char delim = argc > 2 ? argv[2][0] : ',';
    char quote = argc > 3 ? argv[3][0] : '"';
    csvmonkey::CsvReader<> reader(cur, delim, quote);
    auto &row = reader.row();
    std::string out;
    while (reader.read_row()) {
        uint32_t cnt = (uint32_t)row.count;
        out.append((const char *)&cnt, 4);
        for (size_t i = 0; i < row.count; ++i) {
            std::string s = row.cells[i].ptr ? row.cells[i].as_str() : std::string();
            uint32_t l = (uint32_t)s.size();
            out.append((const char *)&l, 4);
            out.append(s);
        }
        if (out.size() > (1u << 20)) {
            fwrite(out.data(), 1, out.size(), stdout);
            out.clear();
        }
    }
    fwrite(out.data(), 1, out.size(), stdout);
    return 0;
}
