// patterns.cpp -- site-pattern compression, the step immediately upstream of the likelihood path.
//
// Replaces PatternWeight (src/treesub.c:1386-1516).  The reference keeps a sorted array of pattern representatives and
// inserts every new pattern with a binary search + memmove: O(npatt^2) element moves (minutes at 1e6 patterns).  The
// RESULT it defines is: within each gene, the distinct alignment columns in ascending strcmp order of the column's
// characters + 1 (bytes compared as unsigned char, a zero byte ends the comparison), each represented by the first
// site that shows it; fpatt = column counts; pose = site -> pattern.  Here the same result is produced in
// O(ls * ns + npatt log npatt): columns are transposed into contiguous keys, de-duplicated with an open-addressing
// hash table (first occurrence wins, equality verified on the full key), and only the distinct keys are sorted.
// Host code: byte shuffling with no arithmetic; it runs once per data set, before the first evaluation.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/paml_b200.h"

namespace {

inline uint64_t mix64(uint64_t h, uint64_t v)
{
    h ^= v;
    h *= 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    return h;
}

inline uint64_t hash_key(const unsigned char* k, size_t B)
{
    uint64_t h = 0x243F6A8885A308D3ull ^ B;
    size_t i = 0;
    for (; i + 8 <= B; i += 8) { uint64_t w; memcpy(&w, k + i, 8); h = mix64(h, w); }
    if (i < B) { uint64_t w = 0; memcpy(&w, k + i, B - i); h = mix64(h, w); }
    h ^= h >> 32;
    return h;
}

}  // namespace

static int compress_impl(int ns, int width, long long ls, const unsigned char* const* z, const int* gene,
                         int ngene, long long* npatt_out, long long* posG, long long* lgene, int* pose,
                         double* fpatt, long long* rep)
{
    if (ns < 1 || width < 1 || ls < 1 || !z || ngene < 1 || !npatt_out || !posG || !pose || !fpatt || !rep) return 1;
    if (ls > 0x7fffffffLL) return 1;                          // pose[] is int in the reference
    const size_t B = (size_t)ns * width;                      // bytes per column
    // keys[h] = column h, sequence-major inside the column like the reference's zt (src/treesub.c:1424-1427), +1 per byte;
    // what follows a zero byte cannot influence strcmp: cleared so that memcmp gives the same order and equality
    std::vector<unsigned char> keys((size_t)ls * B);
    const long long TILE = 256;
    for (long long h0 = 0; h0 < ls; h0 += TILE) {
        const long long h1 = std::min(ls, h0 + TILE);
        for (int j = 0; j < ns; j++) {
            const unsigned char* row = z[j];
            for (long long h = h0; h < h1; h++)
                for (int k = 0; k < width; k++) keys[(size_t)h * B + (size_t)j * width + k] = (unsigned char)(row[h * width + k] + 1);
        }
    }
    for (long long h = 0; h < ls; h++) {
        unsigned char* k = &keys[(size_t)h * B];
        const void* zero = memchr(k, 0, B);
        if (zero) { size_t at = (const unsigned char*)zero - k; memset(k + at, 0, B - at); }
    }

    std::vector<long long> cnt(ngene, 0);
    for (long long h = 0; h < ls; h++) {
        const int g = gene ? gene[h] : 0;
        if (g < 0 || g >= ngene) return 2;
        cnt[g]++;
    }
    if (lgene) for (int g = 0; g < ngene; g++) lgene[g] = cnt[g];

    // sites of every gene, in site order (counting sort on the gene label)
    std::vector<long long> gstart(ngene + 1, 0), gsites((size_t)ls);
    for (int g = 0; g < ngene; g++) gstart[g + 1] = gstart[g] + cnt[g];
    {
        std::vector<long long> fill(gstart.begin(), gstart.end() - 1);
        for (long long h = 0; h < ls; h++) gsites[fill[gene ? gene[h] : 0]++] = h;
    }

    std::vector<int> site_pat((size_t)ls);                    // site -> distinct-key id within its gene (insertion order)
    long long npatt = 0;
    for (int g = 0; g < ngene; g++) {
        posG[g] = npatt;
        if (cnt[g] == 0) continue;                            // the caller reports empty genes like the reference does
        size_t cap = 16;
        while (cap < (size_t)cnt[g] * 2) cap <<= 1;
        std::vector<int> table(cap, -1);                      // slot -> id
        std::vector<long long> first;                         // id -> first site
        std::vector<double> count;                            // id -> sites
        for (long long q = gstart[g]; q < gstart[g + 1]; q++) {
            const long long h = gsites[q];
            const unsigned char* k = &keys[(size_t)h * B];
            size_t s = (size_t)hash_key(k, B) & (cap - 1);
            for (;;) {
                const int id = table[s];
                if (id < 0) {
                    table[s] = (int)first.size();
                    site_pat[h] = (int)first.size();
                    first.push_back(h); count.push_back(1.0);
                    break;
                }
                if (memcmp(&keys[(size_t)first[id] * B], k, B) == 0) { site_pat[h] = id; count[id] += 1.0; break; }
                s = (s + 1) & (cap - 1);
            }
        }
        const int nu = (int)first.size();
        std::vector<int> order(nu);
        for (int i = 0; i < nu; i++) order[i] = i;
        std::sort(order.begin(), order.end(), [&](int a, int b) {
            return memcmp(&keys[(size_t)first[a] * B], &keys[(size_t)first[b] * B], B) < 0;
        });
        std::vector<int> rank(nu);
        for (int r = 0; r < nu; r++) {
            rank[order[r]] = r;
            rep[npatt + r] = first[order[r]];
            fpatt[npatt + r] = count[order[r]];
        }
        for (long long q = gstart[g]; q < gstart[g + 1]; q++) { const long long h = gsites[q]; pose[h] = (int)(npatt + rank[site_pat[h]]); }
        npatt += nu;
    }
    posG[ngene] = npatt;
    *npatt_out = npatt;
    return 0;
}

extern "C" int pb200_compress_patterns(int ns, int width, long long ls, const unsigned char* const* z, const int* gene,
                                       int ngene, long long* npatt_out, long long* posG, long long* lgene, int* pose,
                                       double* fpatt, long long* rep)
{
    try {
        return compress_impl(ns, width, ls, z, gene, ngene, npatt_out, posG, lgene, pose, fpatt, rep);
    }
    catch (...) {
        return 3;                                             // out of memory
    }
}
