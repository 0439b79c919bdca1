// pgb_kernels.cu — sm_100a kernels of libpgb200 and their device-pointer launchers.
//
//   K0  k0_compact_samples   --include-sam keep-mask -> ascending kept-sample index list
//                            (warp ballot / popc ranks; block-level carry)
//   K1  k1_*                 device record index + per-line length exclusive prefix sum
//                            (pass one of the two-pass formatter)
//   K2  k2_format_kernel     decode + gather + format (pass two), body in k2_core.cuh
//   synth / fill             synthetic record generator and store-only calibration
//
// Reference path replaced: /root/reference/src/pfile.rs:149-192 (see k2_core.cuh).
// All work is integer/byte shuffling bounded by HBM bandwidth; no tensor cores.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "k2_core.cuh"
#include "pgb_internal.h"

// ------------------------------------------------------------------ K0 -----
// One CTA of 1024 threads walks the mask 1024 samples at a time.  Each warp ballots its
// 32 keep flags; a lane's rank is popc(ballot & lanemask_lt); warp totals are scanned by
// warp 0 and a running base carries across iterations.  N <= 2^32 samples, one-time per
// export (the kept list is identical for every variant, pfile.rs:128,171).
__global__ void __launch_bounds__(1024) k0_compact_samples(const uint8_t *__restrict__ keep, uint32_t n,
                                                           uint32_t *__restrict__ kidx, uint32_t *__restrict__ count) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_total;
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint32_t base = 0;
    for (uint64_t start = 0; start < n; start += 1024) {
        const uint64_t s = start + threadIdx.x;
        const bool k = (s < n) && keep[s] != 0;
        const uint32_t b = __ballot_sync(0xffffffffu, k);
        const uint32_t rank = __popc(b & ((1u << lane) - 1u));
        if (lane == 0) s_warp[wid] = __popc(b);
        __syncthreads();
        if (wid == 0) {
            const uint32_t v = s_warp[lane];
            uint32_t inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
                if ((int)lane >= d) inc += t;
            }
            s_warp[lane] = inc - v;
            if (lane == 31) s_total = inc;
        }
        __syncthreads();
        if (k) kidx[base + s_warp[wid] + rank] = (uint32_t)s;
        base += s_total;
        __syncthreads();
    }
    if (threadIdx.x < 8) kidx[base + threadIdx.x] = 0; // padding read by K2's 5th lookup
    if (threadIdx.x == 0) *count = base;
}

// ------------------------------------------------------------------ K1 -----
// L_i = P_i + 4K + 1 with P_i = prefix_off[i+1] - prefix_off[i]; line_off = exclusive
// scan of L.  Reduce-then-scan in three launches (no inter-CTA spinning):
//   k1_reduce   per-tile (2048 lines) sums
//   k1_scan     one CTA scans the tile sums in place, writes the grand total
//   k1_emit     per-tile local scan + tile base -> pgb_line_meta (also the record index
//               rec_off = var_row * pitch, replacing pfile.rs:165)
constexpr int K1_THREADS = 256;
constexpr int K1_ITEMS = 8;
constexpr int K1_TILE = K1_THREADS * K1_ITEMS;

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

__global__ void __launch_bounds__(K1_THREADS) k1_reduce(const uint64_t *__restrict__ prefix_off, uint64_t n,
                                                        uint64_t fixed, uint64_t *__restrict__ tile_sum) {
    __shared__ uint64_t s_w[K1_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * K1_TILE;
    uint64_t acc = 0;
#pragma unroll
    for (int k = 0; k < K1_ITEMS; k++) {
        const uint64_t i = base + (uint64_t)k * K1_THREADS + threadIdx.x;
        if (i < n) acc += (prefix_off[i + 1] - prefix_off[i]) + fixed;
    }
    acc = warp_sum_u64(acc);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
        for (int w = 0; w < K1_THREADS / 32; w++) t += s_w[w];
        tile_sum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) k1_scan(uint64_t *__restrict__ tile_sum, uint64_t n_tiles,
                                                pgb_line_meta *__restrict__ meta_last) {
    __shared__ uint64_t s_w[32];
    __shared__ uint64_t s_tot;
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint64_t carry = 0;
    for (uint64_t start = 0; start < n_tiles; start += 1024) {
        const uint64_t i = start + threadIdx.x;
        const uint64_t v = i < n_tiles ? tile_sum[i] : 0;
        uint64_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if ((int)lane >= d) inc += t;
        }
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const uint64_t wv = s_w[lane];
            uint64_t winc = wv;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint64_t t = __shfl_up_sync(0xffffffffu, winc, d);
                if ((int)lane >= d) winc += t;
            }
            s_w[lane] = winc - wv;
            if (lane == 31) s_tot = winc;
        }
        __syncthreads();
        if (i < n_tiles) tile_sum[i] = carry + s_w[wid] + (inc - v);
        carry += s_tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        pgb_line_meta m;
        m.line_off = carry;
        m.rec_off = 0;
        m.pfx_off = 0;
        m.pfx_len = 0;
        m.reserved = 0;
        *meta_last = m;
    }
}

__global__ void __launch_bounds__(K1_THREADS) k1_emit(const uint32_t *__restrict__ var_row,
                                                      const uint64_t *__restrict__ prefix_off, uint64_t prefix_base,
                                                      uint64_t n, uint64_t fixed, uint64_t pitch,
                                                      const uint64_t *__restrict__ tile_base,
                                                      pgb_line_meta *__restrict__ meta) {
    __shared__ uint64_t s_w[K1_THREADS / 32];
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint64_t first = (uint64_t)blockIdx.x * K1_TILE + (uint64_t)threadIdx.x * K1_ITEMS;
    uint64_t po[K1_ITEMS + 1];
#pragma unroll
    for (int k = 0; k <= K1_ITEMS; k++) {
        const uint64_t i = first + k;
        po[k] = i <= n ? prefix_off[i] : 0;
    }
    uint64_t tsum = 0;
#pragma unroll
    for (int k = 0; k < K1_ITEMS; k++)
        if (first + k < n) tsum += (po[k + 1] - po[k]) + fixed;
    uint64_t inc = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if ((int)lane >= d) inc += t;
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    uint64_t off = tile_base[blockIdx.x] + (inc - tsum);
    for (uint32_t w = 0; w < wid; w++) off += s_w[w];
#pragma unroll
    for (int k = 0; k < K1_ITEMS; k++) {
        const uint64_t i = first + k;
        if (i < n) {
            const uint64_t P = po[k + 1] - po[k];
            const uint64_t row = var_row ? (uint64_t)var_row[i] : i;
            uint4 a, b;
            a.x = (uint32_t)off; a.y = (uint32_t)(off >> 32);
            const uint64_t ro = row * pitch;
            a.z = (uint32_t)ro; a.w = (uint32_t)(ro >> 32);
            const uint64_t pf = po[k] - prefix_base;
            b.x = (uint32_t)pf; b.y = (uint32_t)(pf >> 32);
            b.z = (uint32_t)P; b.w = 0;
            uint4 *dst = reinterpret_cast<uint4 *>(meta + i);
            dst[0] = a;
            dst[1] = b;
            off += P + fixed;
        }
    }
}

// ------------------------------------------------------------------ K2 -----
// Persistent CTAs (one resident wave) walk the (line, tile) items warp by warp with a grid
// stride, so that the shared-memory tables are built once per CTA and neighbouring warps
// write neighbouring lines at the same time.
constexpr int K2_THREADS = 256;
constexpr int K2_WARPS = K2_THREADS / 32;

template <bool GATHER, int HINT, int REPL>
__global__ void __launch_bounds__(K2_THREADS) k2_format_kernel(const pgb_k2_params p) {
    __shared__ __align__(16) pgb_u4 s_lut4[256 * REPL];
    {
        const pgb_u4 e = pgb_lut_entry(threadIdx.x);
#pragma unroll
        for (int g = 0; g < REPL; g++) s_lut4[threadIdx.x * REPL + g] = e;
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t first = blockIdx.x * K2_WARPS + (threadIdx.x >> 5); // < 2^31: the grid is one wave
    uint64_t line = first / p.n_tiles;
    uint32_t tile = first - (uint32_t)line * p.n_tiles;
    while (line < p.n_lines) {
        pgb_k2_item<GATHER, HINT, REPL>(p, line, tile, lane, s_lut4);
        line += p.stride_lines;
        tile += p.stride_tiles;
        if (tile >= p.n_tiles) {
            tile -= p.n_tiles;
            line++;
        }
    }
}

// -------------------------------------------------------------- synth ------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// Integer hash generator documented in tools/synth.py (same bytes as its numpy twin).
__global__ void __launch_bounds__(256) synth_records_kernel(uint8_t *__restrict__ records, uint64_t pitch, uint64_t seed,
                                                            uint64_t row0, uint64_t n_rows, uint32_t n_samples,
                                                            uint32_t R) {
    const uint64_t total = n_rows * (uint64_t)R;
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t row = idx / R;
        const uint32_t j = (uint32_t)(idx - row * R);
        const uint64_t hv = splitmix64(seed * 0x2545F4914F6CDD1Dull + (row0 + row));
        const uint64_t p16 = 655ull + (((hv >> 40) * 32113ull) >> 24);
        const uint64_t q16 = 65536ull - p16;
        const uint64_t t0 = (q16 * q16 * 64881ull) >> 32;
        const uint64_t t1 = t0 + ((2ull * p16 * q16 * 64881ull) >> 32);
        const uint64_t hb = splitmix64(hv ^ ((uint64_t)(j + 1) * 0xD6E8FEB86659FD93ull));
        uint32_t byte = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint64_t u = (hb >> (16 * k)) & 0xFFFFull;
            uint32_t code = (u >= 64881ull) ? 3u : (uint32_t)(u >= t0) + (uint32_t)(u >= t1);
            if ((uint64_t)j * 4 + k >= n_samples) code = 0;
            byte |= code << (2 * k);
        }
        records[row * pitch + j] = (uint8_t)byte;
    }
}

__global__ void __launch_bounds__(256) fill_kernel(uint8_t *dst, uint64_t n16, int hint) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride)
        pgb_st16((uint64_t)(uintptr_t)dst + i * 16, 0x302F3009u, 0x312F3009u, 0x312F3109u, 0x2E2F2E09u, hint);
}

// ---------------------------------------------------------- launchers ------
static int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        pgb_set_error("%s: %s", what, cudaGetErrorString(e));
        return PGB_E_CUDA;
    }
    return PGB_OK;
}

extern "C" int pgb_dev_compact_samples(const uint8_t *keep_mask, uint32_t n_samples, uint32_t *kidx, uint32_t *count,
                                       void *stream) {
    if (!keep_mask || !kidx || !count) return PGB_E_ARG;
    k0_compact_samples<<<1, 1024, 0, (cudaStream_t)stream>>>(keep_mask, n_samples, kidx, count);
    return check_launch("k0_compact_samples");
}

extern "C" uint64_t pgb_dev_index_scratch_bytes(uint64_t n_lines) {
    uint64_t tiles = (n_lines + K1_TILE - 1) / K1_TILE;
    return (tiles + 1) * sizeof(uint64_t);
}

extern "C" int pgb_dev_index_lines(const uint32_t *var_row, const uint64_t *prefix_off, uint64_t prefix_base,
                                   uint64_t n_lines, uint32_t n_kept, uint64_t pitch, pgb_line_meta *meta,
                                   void *scratch, void *stream) {
    if (!prefix_off || !meta || !scratch) return PGB_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t fixed = 4ull * n_kept + 1ull;
    const uint64_t tiles = (n_lines + K1_TILE - 1) / K1_TILE;
    uint64_t *tile_sum = (uint64_t *)scratch;
    if (tiles > 0x7fffffffull) return PGB_E_ARG;
    if (tiles) {
        k1_reduce<<<(unsigned)tiles, K1_THREADS, 0, st>>>(prefix_off, n_lines, fixed, tile_sum);
        int rc = check_launch("k1_reduce");
        if (rc) return rc;
    }
    k1_scan<<<1, 1024, 0, st>>>(tile_sum, tiles, meta + n_lines);
    int rc = check_launch("k1_scan");
    if (rc) return rc;
    if (tiles) {
        k1_emit<<<(unsigned)tiles, K1_THREADS, 0, st>>>(var_row, prefix_off, prefix_base, n_lines, fixed, pitch, tile_sum,
                                                       meta);
        rc = check_launch("k1_emit");
    }
    return rc;
}

template <bool GATHER, int HINT, int REPL>
static int launch_k2(pgb_k2_params &p, cudaStream_t st) {
    const uint64_t n_items = p.n_lines * (uint64_t)p.n_tiles;
    if (n_items == 0) return PGB_OK;
    // one resident wave: SM count x CTAs per SM, cached per device
    static int wave[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (wave[dev] == 0) {
        int sms = 0, per_sm = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k2_format_kernel<GATHER, HINT, REPL>, K2_THREADS, 0) !=
                cudaSuccess || per_sm <= 0)
            per_sm = 4;
        wave[dev] = sms * per_sm;
    }
    uint64_t blocks = (n_items + K2_WARPS - 1) / K2_WARPS;
    if (blocks > (uint64_t)wave[dev]) blocks = (uint64_t)wave[dev];
    const uint64_t stride = blocks * K2_WARPS;
    p.stride_lines = stride / p.n_tiles;
    p.stride_tiles = (uint32_t)(stride - p.stride_lines * p.n_tiles);
    k2_format_kernel<GATHER, HINT, REPL><<<(unsigned)blocks, K2_THREADS, 0, st>>>(p);
    return check_launch("k2_format_kernel");
}

// variant: bits 0-3  store hint (0 => .cs streaming, the measured best; 1 => .cs; 2 => default write-back)
//          bits 4-7  LUT copies (0 => 8 interleaved, bank-conflict-free; 1 => a single copy)
//          bits 12-15 tile size = 4 KiB << n (0 => 16 KiB)
extern "C" int pgb_dev_format_lines(const uint8_t *records, const pgb_line_meta *meta, uint64_t n_lines,
                                    const uint8_t *prefix_blob, const uint32_t *kidx, uint32_t n_kept,
                                    uint32_t max_prefix_len, uint8_t *out, int variant, void *stream) {
    if (n_lines == 0) return PGB_OK;
    if (!records || !meta || !out || (!prefix_blob && max_prefix_len)) return PGB_E_ARG;
    pgb_k2_params p;
    p.records = records;
    p.meta = meta;
    p.prefix_blob = prefix_blob;
    p.kidx = kidx;
    p.out = out;
    p.n_lines = n_lines;
    p.K = n_kept;
    const int hint = (variant & 0xF) == 2 ? 0 : 1;
    const int single = (variant >> 4) & 0xF;
    const int tsel = (variant >> 12) & 0xF;
    p.tile_bytes = tsel ? (4096u << tsel) : 16384u;
    const uint64_t max_line = (uint64_t)max_prefix_len + 4ull * n_kept + 1ull;
    const uint64_t nt = (max_line + 511ull + p.tile_bytes - 1) / p.tile_bytes;
    if (nt > 0x7fffffffull) return PGB_E_ARG;
    p.n_tiles = (uint32_t)nt;
    p.stride_lines = 0;
    p.stride_tiles = 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool g = kidx != nullptr;
    if (hint == 1) {
        if (single) return g ? launch_k2<true, 1, 1>(p, st) : launch_k2<false, 1, 1>(p, st);
        return g ? launch_k2<true, 1, 8>(p, st) : launch_k2<false, 1, 8>(p, st);
    }
    if (single) return g ? launch_k2<true, 0, 1>(p, st) : launch_k2<false, 0, 1>(p, st);
    return g ? launch_k2<true, 0, 8>(p, st) : launch_k2<false, 0, 8>(p, st);
}

extern "C" int pgb_dev_synth_records(uint8_t *records, uint64_t pitch, uint64_t seed, uint64_t row0, uint64_t n_rows,
                                     uint32_t n_samples, void *stream) {
    const uint32_t R = pgb_record_bytes(n_samples);
    if (!records || pitch < R) return PGB_E_ARG;
    const uint64_t total = n_rows * (uint64_t)R;
    if (total == 0) return PGB_OK;
    uint64_t blocks = (total + 255) / 256;
    if (blocks > 148ull * 64) blocks = 148ull * 64;
    synth_records_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(records, pitch, seed, row0, n_rows, n_samples,
                                                                            R);
    return check_launch("synth_records_kernel");
}

extern "C" int pgb_dev_fill(uint8_t *dst, uint64_t bytes, int variant, void *stream) {
    if (!dst || ((uintptr_t)dst & 15)) return PGB_E_ARG;
    const uint64_t n16 = bytes / 16;
    if (!n16) return PGB_OK;
    uint64_t blocks = (n16 + 255) / 256;
    if (blocks > 148ull * 32) blocks = 148ull * 32;
    fill_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dst, n16, variant & 0xF);
    return check_launch("fill_kernel");
}
