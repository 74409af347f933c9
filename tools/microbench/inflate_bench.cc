// inflate_bench.cc -- times csrc/host/fast_inflate.cc and zlib over every BGZF member of a BAM file (one thread).
//   g++ -O2 -std=c++17 -I metasnv_b200/csrc/host tools/microbench/inflate_bench.cc metasnv_b200/csrc/host/fast_inflate.cc -lz -o inflate_bench && ./inflate_bench some.bam
#include <zlib.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include "fast_inflate.hpp"
// time both decoders over every BGZF member of a BAM file
int main(int argc, char** argv) {
    int fd = open(argv[1], O_RDONLY); struct stat st; fstat(fd, &st);
    const uint8_t* m = (const uint8_t*)mmap(nullptr, st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    struct Mem { size_t off, clen, isize; }; std::vector<Mem> ms; size_t pos = 0, total = 0;
    while (pos + 18 < (size_t)st.st_size) { uint16_t bsize; memcpy(&bsize, m + pos + 16, 2); uint32_t isize; memcpy(&isize, m + pos + bsize + 1 - 4, 4);
        ms.push_back({pos + 18, (size_t)bsize + 1 - 18 - 8, isize}); total += isize; pos += bsize + 1; }
    std::vector<uint8_t> out(70000), out2(70000);
    for (int rep = 0; rep < 3; ++rep) {
        auto t0 = std::chrono::steady_clock::now(); size_t bad = 0;
        for (auto& x : ms) if (x.isize && !msnv::fast_inflate(m + x.off, x.clen, out.data(), x.isize)) ++bad;
        auto t1 = std::chrono::steady_clock::now();
        for (auto& x : ms) { if (!x.isize) continue; z_stream zs; memset(&zs, 0, sizeof zs); inflateInit2(&zs, -15); zs.next_in = (Bytef*)(m + x.off); zs.avail_in = x.clen; zs.next_out = out2.data(); zs.avail_out = x.isize; inflate(&zs, Z_FINISH); inflateEnd(&zs); }
        auto t2 = std::chrono::steady_clock::now();
        double a = std::chrono::duration<double>(t1 - t0).count(), b = std::chrono::duration<double>(t2 - t1).count();
        printf("%zu members %.1f MB out: fast %.3f s (%.0f MB/s)  zlib %.3f s (%.0f MB/s)  failures %zu\n", ms.size(), total / 1e6, a, total / 1e6 / a, b, total / 1e6 / b, bad);
    }
}
