// cache.cu — the on-disk format either side of the path and handles that share one copy of the capacities.
//
// The reference reads an instance from text: `n m S`, then per arc `tail head (lb ub reward) x S` (Network::Network,
// /root/reference/Network.cpp:18-51) — O(m*S) tokens, 3e8 integers at C5.  The cache is what K1 reads, byte for byte:
//
//   [CacheHeader, 4096 bytes] [tail m x i32] [head m x i32] [reward0 m x i32] [vbar nvbar x i32] (padded to 4096)
//   [cap_u: S rows of m_pad fp64] [cap_l: S rows of m_pad fp64]              (m_pad = m rounded up to even, pad column 0)
//
// i.e. the scenario-major arrays exactly as they lie in HBM.  Loading is: mmap, build the model from the small
// arrays, stream the rows of the wanted scenario block into the device arrays through two pinned staging chunks with plain
// cudaMemcpyAsync (one per chunk).  A rank of a partition loads only its block: rows are contiguous in the file.
// sgufp_clone gives another host thread its own handle (stream, plan and result buffers) on the SAME device arrays.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "capi_internal.hpp"
#include "ctx.hpp"

using namespace sgufp;

namespace {

struct CacheHeader {
    char magic[8];            // "SGUFPC01"
    int32_t version, n, m, S, m_pad, nvbar, max_cap, max_lower;
    int64_t off_tail, off_head, off_rew, off_vbar, off_u, off_l, file_bytes;
};
constexpr size_t HEADER_BYTES = 4096, CHUNK_BYTES = (size_t)32 << 20;
thread_local std::string g_cache_error;

size_t up4k(size_t x) { return (x + 4095) & ~(size_t)4095; }

}  // namespace

extern "C" {

const char *sgufp_cache_last_error(void) { return g_cache_error.c_str(); }

int sgufp_cache_write(const char *path, int n, int m, int S, const int32_t *tail, const int32_t *head, const int32_t *upper,
                      const int32_t *lower, const int32_t *reward0, const int32_t *vbar, int nvbar) {
    if (!path || n < 2 || m < 1 || S < 0 || !tail || !head || !reward0 || (nvbar > 0 && !vbar) || (S > 0 && (!upper || !lower))) {
        g_cache_error = "bad sizes or null arrays"; return SGUFP_ERR_ARG;
    }
    CacheHeader H{};
    std::memcpy(H.magic, "SGUFPC01", 8);
    H.version = 1; H.n = n; H.m = m; H.S = S; H.m_pad = (m + 1) & ~1; H.nvbar = nvbar;
    for (size_t i = 0; i < (size_t)m * S; i++) {
        if (upper[i] < 0 || lower[i] < 0 || upper[i] >= (1 << 20) || lower[i] >= (1 << 20)) { g_cache_error = "capacities must lie in [0, 2^20) (DESIGN.md §5)"; return SGUFP_ERR_LIMITS; }
        H.max_cap = std::max(H.max_cap, std::max(upper[i], lower[i]));
        H.max_lower = std::max(H.max_lower, lower[i]);
    }
    size_t off = HEADER_BYTES;
    H.off_tail = (int64_t)off; off += (size_t)m * 4;
    H.off_head = (int64_t)off; off += (size_t)m * 4;
    H.off_rew = (int64_t)off; off += (size_t)m * 4;
    H.off_vbar = (int64_t)off; off += (size_t)nvbar * 4;
    off = up4k(off);
    const size_t rows = (size_t)S * H.m_pad * 8;
    H.off_u = (int64_t)off; off += rows;
    H.off_l = (int64_t)off; off += rows;
    H.file_bytes = (int64_t)off;
    FILE *f = std::fopen(path, "wb");
    if (!f) { g_cache_error = std::string("cannot open ") + path; return SGUFP_ERR_ARG; }
    std::vector<char> head_block(HEADER_BYTES, 0);
    std::memcpy(head_block.data(), &H, sizeof(H));
    bool ok = std::fwrite(head_block.data(), 1, HEADER_BYTES, f) == HEADER_BYTES;
    ok = ok && std::fwrite(tail, 4, m, f) == (size_t)m && std::fwrite(head, 4, m, f) == (size_t)m && std::fwrite(reward0, 4, m, f) == (size_t)m;
    ok = ok && (nvbar == 0 || std::fwrite(vbar, 4, nvbar, f) == (size_t)nvbar);
    { std::vector<char> pad((size_t)H.off_u - ((size_t)H.off_vbar + (size_t)nvbar * 4), 0); ok = ok && (pad.empty() || std::fwrite(pad.data(), 1, pad.size(), f) == pad.size()); }
    // arc-major int32 -> scenario-major fp64, a band of scenarios at a time (32 x 32 tiles keep both sides in cache)
    const int band = (int)std::max<size_t>(32, std::min<size_t>(4096, ((size_t)16 << 20) / ((size_t)H.m_pad * 8)) / 32 * 32);
    std::vector<double> buf((size_t)band * H.m_pad);
    for (int pass = 0; pass < 2 && ok; pass++) {
        const int32_t *src = pass == 0 ? upper : lower;
        for (int s0 = 0; s0 < S && ok; s0 += band) {
            const int ns = std::min(band, S - s0);
            std::fill(buf.begin(), buf.begin() + (size_t)ns * H.m_pad, 0.0);
            for (int a0 = 0; a0 < m; a0 += 32)
                for (int t0 = 0; t0 < ns; t0 += 32)
                    for (int a = a0; a < std::min(m, a0 + 32); a++) {
                        const int32_t *row = src + (size_t)a * S + s0;
                        for (int t = t0; t < std::min(ns, t0 + 32); t++) buf[(size_t)t * H.m_pad + a] = (double)row[t];
                    }
            ok = std::fwrite(buf.data(), 8, (size_t)ns * H.m_pad, f) == (size_t)ns * H.m_pad;
        }
    }
    ok = std::fclose(f) == 0 && ok;
    if (!ok) { g_cache_error = std::string("short write to ") + path; return SGUFP_ERR_ARG; }
    return SGUFP_OK;
}

int sgufp_cache_dims(const char *path, int *n, int *m, int *S, int *nvbar, int64_t *file_bytes) {
    FILE *f = path ? std::fopen(path, "rb") : nullptr;
    if (!f) { g_cache_error = std::string("cannot open ") + (path ? path : "(null)"); return SGUFP_ERR_ARG; }
    CacheHeader H{};
    const bool ok = std::fread(&H, sizeof(H), 1, f) == 1 && std::memcmp(H.magic, "SGUFPC01", 8) == 0 && H.version == 1;
    std::fclose(f);
    if (!ok) { g_cache_error = "not an SGUFPC01 cache file"; return SGUFP_ERR_ARG; }
    if (n) *n = H.n; if (m) *m = H.m; if (S) *S = H.S; if (nvbar) *nvbar = H.nvbar; if (file_bytes) *file_bytes = H.file_bytes;
    return SGUFP_OK;
}

int sgufp_create_from_cache(sgufp_ctx **out, const char *path, int device, int64_t scenario_offset, int64_t S_local) {
    if (!out) { g_cache_error = "out is null"; return SGUFP_ERR_ARG; }
    *out = nullptr;
    const int fd = path ? open(path, O_RDONLY) : -1;
    if (fd < 0) { g_cache_error = std::string("cannot open ") + (path ? path : "(null)"); return SGUFP_ERR_ARG; }
    struct stat sb;
    if (fstat(fd, &sb) != 0 || (size_t)sb.st_size < HEADER_BYTES) { close(fd); g_cache_error = "file too short"; return SGUFP_ERR_ARG; }
    const size_t bytes = (size_t)sb.st_size;
    void *map = mmap(nullptr, bytes, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) { g_cache_error = "mmap failed"; return SGUFP_ERR_ARG; }
    const char *base = static_cast<const char *>(map);
    CacheHeader H;
    std::memcpy(&H, base, sizeof(H));
    sgufp_ctx *c = nullptr;
    int32_t *pin[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    auto bail = [&](int code, const std::string &msg) {
        g_cache_error = msg;
        for (int i = 0; i < 2; i++) { if (pin[i]) cudaFreeHost(pin[i]); if (done[i]) cudaEventDestroy(done[i]); }
        if (c) sgufp_destroy(c);
        munmap(map, bytes);
        return code;
    };
    if (std::memcmp(H.magic, "SGUFPC01", 8) != 0 || H.version != 1 || (size_t)H.file_bytes > bytes || H.m_pad != ((H.m + 1) & ~1))
        return bail(SGUFP_ERR_ARG, "not an SGUFPC01 cache file (or truncated)");
    {   // the offsets are the writer's layout, nothing else: a damaged header must not send the reads outside the mapping
        if (H.n < 2 || H.m < 1 || H.S < 0 || H.nvbar < 0) return bail(SGUFP_ERR_ARG, "cache file header: bad sizes");
        size_t off = HEADER_BYTES;
        const size_t o_tail = off; off += (size_t)H.m * 4;
        const size_t o_head = off; off += (size_t)H.m * 4;
        const size_t o_rew = off; off += (size_t)H.m * 4;
        const size_t o_vbar = off; off += (size_t)H.nvbar * 4;
        off = up4k(off);
        const size_t rows = (size_t)H.S * H.m_pad * 8, o_u = off, o_l = off + rows, total = off + 2 * rows;
        if ((size_t)H.off_tail != o_tail || (size_t)H.off_head != o_head || (size_t)H.off_rew != o_rew || (size_t)H.off_vbar != o_vbar ||
            (size_t)H.off_u != o_u || (size_t)H.off_l != o_l || (size_t)H.file_bytes != total)
            return bail(SGUFP_ERR_ARG, "cache file header: offsets do not match the layout of its sizes");
    }
    if (S_local < 0) S_local = H.S - scenario_offset;
    if (scenario_offset < 0 || S_local < 0 || scenario_offset + S_local > H.S) return bail(SGUFP_ERR_ARG, "scenario block must lie inside [0, S) of the file");
    c = new sgufp_ctx();
    std::string e;
    if (int rc = c->M.build(H.n, H.m, reinterpret_cast<const int32_t *>(base + H.off_tail), reinterpret_cast<const int32_t *>(base + H.off_head),
                            reinterpret_cast<const int32_t *>(base + H.off_rew), reinterpret_cast<const int32_t *>(base + H.off_vbar), H.nvbar, e))
        return bail(rc, e);
    c->S = (int)S_local; c->scen_off = scenario_offset; c->S_total = std::max(1, H.S); c->m_pad = H.m_pad;
    c->max_cap = H.max_cap; c->max_lower = H.max_lower;
    { long long sa = 0; const int32_t *r = reinterpret_cast<const int32_t *>(base + H.off_rew); for (int a = 0; a < H.m; a++) sa += std::abs((long long)r[a]); c->sum_abs_r = (int)sa; }
    c->device = device;
    if (device == SGUFP_DEVICE_NONE) { munmap(map, bytes); *out = c; return SGUFP_OK; }
    if (int rc = init_device(c, device, e)) return bail(rc, e);
#define CUB(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return bail(SGUFP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } while (0)
    if (S_local > 0) {
        const size_t block = (size_t)S_local * H.m_pad * 8;
        c->caps = new CapStore();
        c->caps->device = device;
        CUB(cudaMalloc(&c->caps->d_u, block));
        CUB(cudaMalloc(&c->caps->d_l, block));
        c->d_u = c->caps->d_u; c->d_l = c->caps->d_l;
        const size_t chunk = std::min(CHUNK_BYTES, block);
        for (int i = 0; i < 2; i++) { CUB(cudaHostAlloc(reinterpret_cast<void **>(&pin[i]), chunk, cudaHostAllocDefault)); CUB(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming)); }
        int slot = 0;
        bool used[2] = {false, false};
        for (int pass = 0; pass < 2; pass++) {
            const char *src = base + (pass == 0 ? H.off_u : H.off_l) + (size_t)scenario_offset * H.m_pad * 8;
            char *dst = reinterpret_cast<char *>(pass == 0 ? c->d_u : c->d_l);
            for (size_t o = 0; o < block; o += chunk, slot ^= 1) {
                const size_t nb = std::min(chunk, block - o);
                if (used[slot]) CUB(cudaEventSynchronize(done[slot]));           // the previous copy out of this staging chunk has left it
                std::memcpy(pin[slot], src + o, nb);                             // page cache (or disk) -> pinned
                CUB(cudaMemcpyAsync(dst + o, pin[slot], nb, cudaMemcpyHostToDevice, c->st));   // one plain copy per chunk
                CUB(cudaEventRecord(done[slot], c->st));
                used[slot] = true;
            }
        }
        CUB(cudaStreamSynchronize(c->st));
        for (int i = 0; i < 2; i++) { cudaFreeHost(pin[i]); pin[i] = nullptr; cudaEventDestroy(done[i]); done[i] = nullptr; }
    }
#undef CUB
    munmap(map, bytes);
    *out = c;
    return SGUFP_OK;
}

int sgufp_clone(const sgufp_ctx *src, sgufp_ctx **out) {
    if (!src || !out) return SGUFP_ERR_ARG;
    *out = nullptr;
    if (src->part) { g_cache_error = "a partition handle cannot be cloned"; return SGUFP_ERR_ARG; }
    sgufp_ctx *c = new sgufp_ctx();
    c->M = src->M;
    c->S = src->S; c->m_pad = src->m_pad; c->max_cap = src->max_cap; c->max_lower = src->max_lower; c->sum_abs_r = src->sum_abs_r;
    c->scen_off = src->scen_off; c->S_total = src->S_total; c->device = src->device;
    if (src->device != SGUFP_DEVICE_NONE) {
        std::string e;
        if (int rc = init_device(c, src->device, e)) { g_cache_error = e; sgufp_destroy(c); return rc; }
        if (src->caps) {
            std::lock_guard<std::mutex> g(cap_store_mutex());
            src->caps->refs++;
            c->caps = src->caps; c->d_u = src->caps->d_u; c->d_l = src->caps->d_l;
        }
    }
    *out = c;
    return SGUFP_OK;
}

}  // extern "C"
