// TEST INFRASTRUCTURE ONLY.  Drives the library's own tile_sort_gather_kernel (binning.cu, emulation build) directly:
// tile lists of every size class of the in-register sort (1 / 2 / 4 / 8 keys per thread), the boundaries between them,
// the shared-memory network beyond 1024 and its 8192 limit, with many equal depths -- against std::sort of the same
// 64-bit keys (depth bits << 32 | index: the order of a stable sort on depth over emission in index order).
#include <algorithm>
#include <random>
#include <vector>
#include "common.cuh"      // the emulation build's copy: brings its simt_emu.h

namespace h3dgs {
void tile_sort_gather_kernel(int gx, int rows, int shard_count, int shard_index, const uint2* ranges, const uint64_t* pairs,
                             const Record* records, uint64_t* keys_sorted, uint32_t* point_list, Record* sorted);
}

int main() {
    const std::vector<int> lens = {1, 2, 31, 32, 33, 100, 127, 128, 129, 200, 255, 256, 257, 400, 511, 512, 513, 777, 1000,
                                   1023, 1024, 1025, 1500, 2047, 2048, 3000, 0, 5, 8192};
    const int gx = (int)lens.size(), P = 20000;
    int bad = 0;
    for (unsigned seed = 1; seed <= 6; seed++) {
        std::mt19937 g(seed);
        std::vector<h3dgs::Record> records(P);
        for (auto& r : records) {
            float* f = reinterpret_cast<float*>(&r);
            for (int k = 0; k < 12; k++) f[k] = std::uniform_real_distribution<float>(0.1f, 40.0f)(g);
            r.b.w = __uint_as_float(1u);                       // kbits: one kid, no clamp flags
        }
        std::vector<uint2> ranges(gx);
        std::vector<uint64_t> pairs;
        size_t D = 0;
        int longest = 1;
        for (int t = 0; t < gx; t++) {
            const int n = lens[t];
            ranges[t] = n ? make_uint2((unsigned)D, (unsigned)(D + n)) : make_uint2(0u, 0u);
            std::vector<unsigned> idx(P);
            for (int i = 0; i < P; i++) idx[i] = (unsigned)i;
            std::shuffle(idx.begin(), idx.end(), g);           // unique Gaussian indices inside a tile, arbitrary emission order
            const int levels = 1 + (int)(g() % 7);              // few distinct depths: many ties, broken by the index
            for (int i = 0; i < n; i++) {
                const float depth = 0.5f + (float)(g() % levels);
                pairs.push_back(((uint64_t)__float_as_uint(depth) << 32) | idx[i]);
            }
            D += n;
            longest = std::max(longest, n);
        }
        std::vector<uint64_t> keys(D + 1, 0);
        std::vector<uint32_t> plist(D + 1, 0);
        std::vector<h3dgs::Record> sorted(D + 2);
        int m = 128;
        while (m < longest) m <<= 1;
        SIMT_LAUNCH((h3dgs::tile_sort_gather_kernel), gx, 128, (size_t)m * 8, 0)
            (gx, 1, 1, 0, ranges.data(), pairs.data(), records.data(), keys.data(), plist.data(), sorted.data());
        for (int t = 0; t < gx; t++) {
            const size_t a = ranges[t].x, b = ranges[t].y;
            std::vector<uint64_t> ref(pairs.begin() + a, pairs.begin() + b);
            std::sort(ref.begin(), ref.end());
            for (size_t i = a; i < b; i++) {
                const uint64_t want = ref[i - a];
                const uint64_t key_want = ((uint64_t)t << 32) | (want >> 32);       // sorted key: tile << 32 | depth bits
                const h3dgs::Record& src = records[(uint32_t)want];
                const bool ok = keys[i] == key_want && plist[i] == (uint32_t)want && sorted[i].a.x == src.a.x && sorted[i].c.w == src.c.w;
                if (!ok && bad++ < 5) fprintf(stderr, "seed %u tile %d (n=%d) entry %zu: got idx %u, want %u\n", seed, t, lens[t], i - a, plist[i], (uint32_t)want);
            }
        }
    }
    printf("tile_sort_test: %s\n", bad ? "MISMATCH" : "all tiles in order");
    return bad ? 1 : 0;
}
