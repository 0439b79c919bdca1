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

#include <algorithm>

#include "k2_batch.cuh"
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

// Prefix length of line i: prefix_off[i + 1] - prefix_off[i] (prefixes packed back to back), or, when the
// prefixes are rows of a raw .pvar image, prefix_len[i] bytes at prefix_off[i]; `fixed` carries the rest of
// the line (suffix + 4K + 1).
__device__ __forceinline__ uint64_t k1_plen(const uint64_t *__restrict__ prefix_off, const uint32_t *__restrict__ prefix_len,
                                            uint64_t i) {
    return prefix_len ? (uint64_t)prefix_len[i] : prefix_off[i + 1] - prefix_off[i];
}

__global__ void __launch_bounds__(K1_THREADS) k1_reduce(const uint64_t *__restrict__ prefix_off,
                                                        const uint32_t *__restrict__ prefix_len, uint64_t n,
                                                        uint64_t fixed, uint64_t *__restrict__ tile_sum) {
    __shared__ uint64_t s_w[K1_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * K1_TILE;
    uint64_t acc = 0;
#pragma unroll
    for (int k = 0; k < K1_ITEMS; k++) {
        const uint64_t i = base + (uint64_t)k * K1_THREADS + threadIdx.x;
        if (i < n) acc += k1_plen(prefix_off, prefix_len, i) + fixed;
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
                                                      const uint64_t *__restrict__ rec_off_in,
                                                      const uint64_t *__restrict__ prefix_off,
                                                      const uint32_t *__restrict__ prefix_len, uint32_t sfx_len,
                                                      uint64_t prefix_base, uint64_t n, uint64_t fixed, uint64_t pitch,
                                                      const uint64_t *__restrict__ tile_base,
                                                      pgb_line_meta *__restrict__ meta) {
    __shared__ uint64_t s_w[K1_THREADS / 32];
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint64_t first = (uint64_t)blockIdx.x * K1_TILE + (uint64_t)threadIdx.x * K1_ITEMS;
    uint64_t po[K1_ITEMS + 1];
    uint32_t pl[K1_ITEMS];
#pragma unroll
    for (int k = 0; k <= K1_ITEMS; k++) {
        const uint64_t i = first + k;
        po[k] = (prefix_len ? i < n : i <= n) ? prefix_off[i] : 0;
    }
#pragma unroll
    for (int k = 0; k < K1_ITEMS; k++) {
        const uint64_t i = first + k;
        pl[k] = i < n ? (prefix_len ? prefix_len[i] : (uint32_t)(po[k + 1] - po[k])) + sfx_len : 0u;
    }
    uint64_t tsum = 0;
#pragma unroll
    for (int k = 0; k < K1_ITEMS; k++)
        if (first + k < n) tsum += pl[k] + fixed;
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
            const uint64_t P = pl[k];
            const uint64_t row = var_row ? (uint64_t)var_row[i] : i;
            uint4 a, b;
            a.x = (uint32_t)off; a.y = (uint32_t)(off >> 32);
            // device record index: explicit offsets (standard-format .pgen, from the header walk of
            // pgen.rs:100-258) or row * pitch (fixed-width mode 0x02, pfile.rs:165)
            const uint64_t ro = rec_off_in ? rec_off_in[i] : row * pitch;
            a.z = (uint32_t)ro; a.w = (uint32_t)(ro >> 32);
            const uint64_t pf = po[k] - prefix_base;
            b.x = (uint32_t)pf; b.y = (uint32_t)(pf >> 32);
            b.z = (uint32_t)P;
            b.w = prefix_len ? 0u : 1u; // bit 0: the prefixes of this launch are packed back to back in prefix_blob
            uint4 *dst = reinterpret_cast<uint4 *>(meta + i);
            dst[0] = a;
            dst[1] = b;
            off += P + fixed;
        }
    }
}

// ------------------------------------------------------------------ K2 -----
// One CTA formats a batch of K2_ITEMS consecutive (line, tile) items, warp w taking items
// w, w + 8, ...  The grid is NOT persistent: small CTAs handed out in order by the hardware
// scheduler keep the set of lines being written compact and the warps staggered, which
// measured 5-15 % more HBM write throughput than one resident wave walking the lines in
// lockstep (tools/fill_sweep.py shows the same effect with pure stores).  Per CTA:
//   1. threads 0..K2_ITEMS-1 fetch their item's pgb_line_meta into shared memory while the
//      other threads build the text LUT;
//   2. every item's record slice and prefix are requested into L2 (prefetch.global.L2), so the
//      later items of each warp find their inputs on chip;
//   3. the warps format their items (k2_core.cuh).
constexpr int K2_THREADS = 256;
constexpr int K2_WARPS = K2_THREADS / 32;

template <bool GATHER, int HINT, int REPL, int IPW, bool ONE>
__global__ void __launch_bounds__(K2_THREADS) k2_format_kernel(const pgb_k2_params p) {
    constexpr int ITEMS = K2_WARPS * IPW;
    __shared__ __align__(16) pgb_u4 s_lut4[256 * REPL];
    __shared__ __align__(16) pgb_line_meta s_meta[ITEMS];
    __shared__ uint32_t s_tile[ITEMS];
    const uint64_t n_items = p.n_lines * (uint64_t)p.n_tiles;
    const uint64_t item0 = (uint64_t)blockIdx.x * ITEMS;
    if (threadIdx.x < ITEMS) {
        const uint64_t item = item0 + threadIdx.x;
        if (item < n_items) {
            uint64_t line = item;
            uint32_t tile = 0;
            if (p.n_tiles > 1) {
                line = item / p.n_tiles;
                tile = (uint32_t)(item - line * p.n_tiles);
            }
            const uint4 *q = reinterpret_cast<const uint4 *>(p.meta + line);
            uint4 *d = reinterpret_cast<uint4 *>(&s_meta[threadIdx.x]);
            d[0] = __ldg(q);
            d[1] = __ldg(q + 1);
            s_tile[threadIdx.x] = tile;
        }
    }
    if (REPL == 1) {
        s_lut4[threadIdx.x] = pgb_lut_entry(threadIdx.x);
    } else {
        // REPL interleaved copies, written without bank conflicts: stage one copy, then thread t
        // fills slots t, t + 256, ... (slot = entry * REPL + copy), consecutive lanes -> consecutive
        // 16-byte slots.
        __shared__ __align__(16) pgb_u4 s_one[REPL > 1 ? 256 : 1];
        s_one[threadIdx.x] = pgb_lut_entry(threadIdx.x);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < REPL; k++) {
            const uint32_t slot = threadIdx.x + 256u * k;
            s_lut4[slot] = s_one[slot / REPL];
        }
    }
    __syncthreads();
    { // L2 prefetch: 8 threads per item walk the 128-byte lines of its record slice and prefix
        static_assert(ITEMS * 8 <= K2_THREADS || ITEMS * 8 % K2_THREADS == 0, "prefetch mapping");
        for (uint32_t t = threadIdx.x; t < (uint32_t)ITEMS * 8u; t += K2_THREADS) {
            const uint32_t slot = t >> 3, part = t & 7u;
            if (item0 + slot >= n_items) continue;
            const pgb_line_meta m = s_meta[slot];
            const uint8_t *row = p.records + m.rec_off;
            uint64_t j0 = 0, j1 = ((uint64_t)p.row_bytes_hint);
            if (!GATHER && p.n_tiles > 1) {
                // record bytes of this tile: GT offsets [t0 - a_gs, t1 - a_gs) / 16
                const uint64_t a_ls = (uint64_t)(uintptr_t)p.out + m.line_off, a_gs = a_ls + m.pfx_len;
                const uint64_t t0 = (a_ls & ~511ull) + (uint64_t)s_tile[slot] * p.tile_bytes, t1 = t0 + p.tile_bytes;
                j0 = t0 > a_gs ? (t0 - a_gs) >> 4 : 0;
                const uint64_t je = t1 > a_gs ? ((t1 - a_gs) >> 4) + 2 : 0;
                j1 = je < j1 ? je : j1;
            } else if (GATHER) {
                if (p.n_tiles > 1 || p.K == 0) {
                    j1 = 0; // the span of a tile's kept samples is not known here
                } else {
                    j0 = __ldg(p.kidx) >> 2;
                    j1 = (__ldg(p.kidx + (p.K - 1)) >> 2) + 1;
                }
            }
            if (j1 > j0) {
                const uint64_t first = ((uint64_t)(uintptr_t)row + j0) & ~127ull;
                const uint64_t last = (uint64_t)(uintptr_t)row + j1;
                for (uint64_t a = first + 128ull * part; a < last && a < first + 128ull * 32; a += 128ull * 8)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
            }
            if (part >= 6 && s_tile[slot] == 0 && m.pfx_len) {
                const uint64_t pa = (uint64_t)(uintptr_t)(p.prefix_blob + m.pfx_off);
                const uint64_t a = (pa & ~127ull) + 128ull * (part - 6);
                if (a < pa + m.pfx_len) asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
            }
        }
    }
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll 1
    for (int j = 0; j < IPW; j++) {
        const uint32_t slot = j * K2_WARPS + warp;
        if (item0 + slot >= n_items) break;
        const pgb_line_meta m = s_meta[slot];
        if (ONE) pgb_k2_line<GATHER, HINT, REPL>(p, m, lane, s_lut4);
        else pgb_k2_item<GATHER, HINT, REPL>(p, m, s_tile[slot], lane, s_lut4);
    }
}

// ------------------------------------------------------------ K2 (batch) ---
// Short lines: a persistent, warp-specialised CTA walks batches of B consecutive lines through two
// shared-memory stages (k2_batch.cuh).  Warp 8 is the producer: it reads the pgb_line_meta of a
// batch one batch ahead, writes the stage's line table and hands the records and prefixes to the
// bulk-copy engine (completion on the stage's `full` mbarrier).  Warps 0-7 are consumers: a warp per
// line, they compact the kept samples, format the text into an image of the batch's contiguous
// output range and release the stage (`empty` mbarrier); the image leaves by ONE bulk async store
// that drains while the next batch is formatted into the other image.
__device__ __forceinline__ void k2b_mbar_wait(uint8_t *mbar, uint32_t parity) {
    const uint32_t bar = k2b_smem_addr(mbar);
    uint32_t done;
    do {
        // the suspend-time hint lets the hardware park the warp instead of spinning in the issue slots
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar), "r"(parity), "r"(20000u)
                     : "memory");
        if (!done) __nanosleep(64); // a waiting warp should not compete for issue slots
    } while (!done);
}

template <bool GATHER, bool IMG>
__global__ void __launch_bounds__(K2B_THREADS, 3) k2_batch_kernel(const pgb_k2b_params p) {
    extern __shared__ __align__(128) uint8_t k2b_smem[];
    uint8_t *smem = k2b_smem;
    const pgb_k2b_layout L = pgb_k2b_smem_layout(p.B, p.rowcap, p.pcap, p.vcap, p.outcap, GATHER, p.images, p.stages);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // record bytes that hold kept samples: all of the record, or the span of the kept-sample list
    uint32_t span_lo = 0, span_len = p.K ? p.R : 0u;
    if (GATHER && p.K) {
        span_lo = __ldg(p.kidx) >> 2;
        span_len = (__ldg(p.kidx + (p.K - 1u)) >> 2) + 1u - span_lo;
        if (tid < K2B_CONSUMERS) k2b_build_plan(p, smem, L, tid, span_lo);
    }
    if (tid == 0) {
        // full[s] at 8 * s: one arrival per producer lane; empty[s] at 32 + 8 * s: one arrival per consumer warp
        for (uint32_t st = 0; st < 3; st++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(k2b_smem_addr(smem + 8u * st)), "r"(32) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(k2b_smem_addr(smem + 32u + 8u * st)), "r"(K2B_WARPS)
                         : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    k2b_build_lut(smem, L, tid);
    __syncthreads(); // barriers initialised, tables built; from here on the two roles never meet at a CTA barrier

    // batches of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...; lines of batch b: [b * B, min(n_lines, (b+1) * B))
    const uint32_t n_lines = (uint32_t)p.n_lines; // < 2^32 (checked by the launcher)
    auto lines_of = [&](uint32_t b) -> uint32_t {
        if (b >= p.n_batches) return 0u;
        const uint32_t left = n_lines - b * p.B;
        return left < p.B ? left : p.B;
    };
    const uint32_t LPW = k2b_lines_per_warp(p.B);
    if (warp == K2B_WARPS) {
        // ------------------------------------------------ producer warp ----
        // The pgb_line_meta of the lane's line in the NEXT batch is kept as loaded: unpacking it right away
        // would make the warp wait for the load a batch early.
        uint4 ra = make_uint4(0, 0, 0, 0), rb = ra, first = ra;
        uint64_t base = 0, end_off = 0;
        auto fetch_meta = [&](uint32_t b) {
            const uint32_t nbl = lines_of(b);
            if (!nbl) return;
            const pgb_line_meta *m0 = p.meta + (uint64_t)b * p.B;
            if (lane < nbl) {
                const uint4 *q = reinterpret_cast<const uint4 *>(m0 + lane);
                ra = __ldg(q);
                rb = __ldg(q + 1);
            }
            base = k2b_ld_u64(&m0->line_off);
            end_off = k2b_ld_u64(&m0[nbl].line_off);
            first = __ldg(reinterpret_cast<const uint4 *>(m0) + 1); // pfx_off, pfx_len, reserved of the first line
        };
        uint32_t bt = blockIdx.x;
        fetch_meta(bt);
        for (uint32_t stage = 0, par = 0;; bt += gridDim.x) { // par: parity of the stage's current use
            const uint32_t nbl = lines_of(bt);
            if (!nbl) break;
            pgb_line_meta m;
            m.line_off = ((uint64_t)ra.y << 32) | ra.x;
            m.rec_off = ((uint64_t)ra.w << 32) | ra.z;
            m.pfx_off = ((uint64_t)rb.y << 32) | rb.x;
            m.pfx_len = rb.z;
            m.reserved = rb.w;
            const uint64_t base_n = base, end_n = end_off, pfx0_n = ((uint64_t)first.y << 32) | first.x;
            const bool packed_n = (first.w & 1u) != 0;
            // body offsets of the first line of the lane's consumer warp and of the line after the lane's
            const uint32_t fl = lane / LPW * LPW;
            const uint64_t first_off = ((uint64_t)__shfl_sync(0xffffffffu, ra.y, fl) << 32) | __shfl_sync(0xffffffffu, ra.x, fl);
            uint64_t next_off = ((uint64_t)__shfl_down_sync(0xffffffffu, ra.y, 1) << 32) | __shfl_down_sync(0xffffffffu, ra.x, 1);
            if (lane + 1u >= nbl) next_off = end_n;
            fetch_meta(bt + gridDim.x); // the index read of the batch after this one
            k2b_mbar_wait(smem + 32u + 8u * stage, par ^ 1u); // consumers are done with the stage
            k2b_produce(p, smem, L, stage, nbl, lane, m, first_off, next_off, base_n, end_n, pfx0_n, packed_n, span_lo, span_len);
            if (++stage == p.stages) { stage = 0; par ^= 1u; }
        }
        return;
    }
    // ---------------------------------------------------- consumer warps ----
    // Warp w formats lines [w * LPW, (w + 1) * LPW) of every batch into its own part of the image and stores that
    // contiguous byte range itself: the consumer warps never wait for one another.
    // the lane's gather plan for bytes lane and lane + 32 of a virtual record, in registers when that is all of it
    pgb_u4 preg0 = {0u, 0u, 0u, 0u}, preg1 = preg0;
    if (GATHER) {
        const uint32_t nb = (p.K + 3u) >> 2;
        const pgb_u4 *plan = reinterpret_cast<const pgb_u4 *>(smem + L.plan);
        if (nb <= 64u) {
            if (lane < nb) preg0 = pgb_lds4(plan + lane);
            if (lane + 32u < nb) preg1 = pgb_lds4(plan + lane + 32u);
        }
    }
    uint32_t bt = blockIdx.x;
    for (uint32_t n = 0, stage = 0, par = 0;; n++, bt += gridDim.x) {
        const uint32_t nbl = lines_of(bt);
        if (!nbl) break;
        const uint32_t img = p.images > 1 ? n & 1u : 0u;
        // the part of the image this warp is about to rewrite (batch n-2's, or n-1's with a single image) must have
        // left shared memory
        if (IMG && lane == 0) {
            if (p.images > 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        k2b_mbar_wait(smem + 8u * stage, par); // table written, records and prefixes landed
        const pgb_u4 h = pgb_lds4(reinterpret_cast<const pgb_u4 *>(smem + L.tab(stage) + 16u * warp));
        __syncwarp();
        const uint32_t ws = h.z, we = h.w;
        const uint64_t g_al = (uint64_t)(uintptr_t)p.out + (((uint64_t)h.y << 32) | h.x) - ws;
        const uint32_t l1 = (warp + 1u) * LPW < nbl ? (warp + 1u) * LPW : nbl;
        for (uint32_t l = warp * LPW; l < l1; l++) {
            k2b_line_gather<GATHER, IMG>(p, smem, L, stage, img, g_al, l, warp, lane, preg0, preg1);
            if (GATHER) __syncwarp();
            k2b_line_format<GATHER, IMG>(p, smem, L, stage, img, g_al, l, warp, lane);
            if (GATHER) __syncwarp();
        }
        if (IMG) {
            // the image was written through the generic proxy; the bulk store reads it through the async proxy
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
        // this warp no longer reads the stage's table, records or prefixes: let the producer refill it
        if (lane == 0) k2b_arrive(smem + 32u + 8u * stage);
        if (++stage == p.stages) { stage = 0; par ^= 1u; }
        if (IMG) {
            const uint8_t *outb = smem + L.outb(img);
            if (ws != 0xFFFFFFFFu) {
                const uint32_t h0 = (ws + 15u) & ~15u, h1 = we & ~15u;
                if (lane == 0 && h0 < h1)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g_al + h0),
                                 "r"(k2b_smem_addr(outb + h0)), "r"(h1 - h0)
                                 : "memory");
                k2b_store_edges(g_al, outb, ws, we, lane);
            }
            // one group per batch (possibly empty), so that wait_group.read 1 means "batch n-2 has left"
            if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    // shared memory must stay intact until the last bulk stores have read it
    if (IMG && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
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

// Cheap generator for the biobank shape (25 GB of records have to be produced on the device before the timed
// region): one murmur3 finaliser per 32-bit word (16 samples), documented in tools/synth.py (synth_records_fast is
// its numpy twin).  Every byte pattern occurs; the genotype distribution is uniform over the four codes.
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    return h ^ (h >> 16);
}

__global__ void __launch_bounds__(256) synth_fast_kernel(uint8_t *__restrict__ records, uint64_t pitch, uint32_t seed,
                                                         uint64_t row0, uint64_t n_rows, uint32_t n_samples, uint32_t R) {
    const uint32_t words = (R + 3u) / 4u;
    const uint64_t total = n_rows * (uint64_t)words;
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t row = idx / words;
        const uint32_t w = (uint32_t)(idx - row * words);
        const uint32_t v = (uint32_t)(row0 + row);
        const uint32_t h = fmix32(seed * 0x9E3779B1u ^ v * 0x85EBCA77u ^ (w + 1u) * 0xC2B2AE3Du);
        uint8_t *dst = records + row * pitch + 4ull * w;
        if ((((uintptr_t)dst) & 3u) == 0 && 4u * w + 4u <= R && 16u * w + 16u <= n_samples) { // a whole aligned word
            *reinterpret_cast<uint32_t *>(dst) = h;
            continue;
        }
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            const uint32_t j = 4u * w + k;
            if (j >= R) break;
            uint32_t byte = (h >> (8u * k)) & 0xFFu;
            const uint32_t first = 4u * j; // samples first .. first + 3; padding samples (>= n_samples) are 0
            if (first + 4u > n_samples) byte &= (1u << (2u * (n_samples - first))) - 1u;
            dst[k] = (uint8_t)byte;
        }
    }
}

__global__ void __launch_bounds__(256) fill_kernel(uint8_t *dst, uint64_t n16, int hint) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride)
        pgb_st16((uint64_t)(uintptr_t)dst + i * 16, 0x302F3009u, 0x312F3009u, 0x312F3109u, 0x2E2F2E09u, hint);
}

// Store-only calibration with K2-like write patterns: the output is cut into segments of
// seg_rows x 512 bytes; `coop` warps of a CTA share a segment (row-interleaved in bursts of
// `burst` rows); the (8 / coop) segment slots of all CTAs walk the segments with a grid stride.
// Separates DRAM/L2 write behaviour from K2's compute.
__global__ void __launch_bounds__(256) fill_segments_kernel(uint8_t *dst, uint64_t n_seg, uint32_t seg_rows, int hint,
                                                            uint32_t coop, uint32_t burst) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t groups = 8u / coop;                 // segments in flight per CTA
    const uint32_t sub = warp % coop;                  // this warp's share of the segment
    const uint64_t seg_stride = (uint64_t)gridDim.x * groups;
    for (uint64_t seg = (uint64_t)blockIdx.x * groups + warp / coop; seg < n_seg; seg += seg_stride) {
        const uint64_t base = (uint64_t)(uintptr_t)dst + seg * seg_rows * 512ull + lane * 16u;
        for (uint32_t r = sub * burst; r < seg_rows; r += burst * coop) {
            for (uint32_t u = 0; u < burst; u++)
                if (r + u < seg_rows)
                    pgb_st16(base + (uint64_t)(r + u) * 512ull, 0x302F3009u, 0x312F3009u, 0x312F3109u, 0x2E2F2E09u, hint);
        }
    }
}

// Plain grid-stride fill whose every thread issues `burst` stores one grid stride apart per iteration.
__global__ void __launch_bounds__(256) fill_strided_kernel(uint8_t *dst, uint64_t n16, int hint, uint32_t burst) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride * burst) {
        for (uint32_t u = 0; u < burst; u++)
            if (i + u * stride < n16)
                pgb_st16((uint64_t)(uintptr_t)dst + (i + u * stride) * 16, 0x302F3009u, 0x312F3009u, 0x312F3109u,
                         0x2E2F2E09u, hint);
    }
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

static int index_lines_impl(const uint32_t *var_row, const uint64_t *rec_off_in, const uint64_t *prefix_off,
                            const uint32_t *prefix_len, uint32_t sfx_len, uint64_t prefix_base, uint64_t n_lines,
                            uint32_t n_kept, uint64_t pitch, pgb_line_meta *meta, void *scratch, void *stream) {
    if (!prefix_off || !meta || !scratch || sfx_len > 4) return PGB_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t fixed = 4ull * n_kept + 1ull; // pl[] in k1_emit already includes the suffix
    const uint64_t tiles = (n_lines + K1_TILE - 1) / K1_TILE;
    uint64_t *tile_sum = (uint64_t *)scratch;
    if (tiles > 0x7fffffffull) return PGB_E_ARG;
    if (tiles) {
        k1_reduce<<<(unsigned)tiles, K1_THREADS, 0, st>>>(prefix_off, prefix_len, n_lines, fixed + sfx_len, tile_sum);
        int rc = check_launch("k1_reduce");
        if (rc) return rc;
    }
    k1_scan<<<1, 1024, 0, st>>>(tile_sum, tiles, meta + n_lines);
    int rc = check_launch("k1_scan");
    if (rc) return rc;
    if (tiles) {
        k1_emit<<<(unsigned)tiles, K1_THREADS, 0, st>>>(var_row, rec_off_in, prefix_off, prefix_len, sfx_len, prefix_base,
                                                       n_lines, fixed, pitch, tile_sum, meta);
        rc = check_launch("k1_emit");
    }
    return rc;
}

extern "C" int pgb_dev_index_lines(const uint32_t *var_row, const uint64_t *prefix_off, uint64_t prefix_base,
                                   uint64_t n_lines, uint32_t n_kept, uint64_t pitch, pgb_line_meta *meta,
                                   void *scratch, void *stream) {
    return index_lines_impl(var_row, nullptr, prefix_off, nullptr, 0, prefix_base, n_lines, n_kept, pitch, meta, scratch,
                            stream);
}

extern "C" int pgb_dev_index_lines_off(const uint64_t *rec_off, const uint64_t *prefix_off, uint64_t prefix_base,
                                       uint64_t n_lines, uint32_t n_kept, pgb_line_meta *meta, void *scratch,
                                       void *stream) {
    if (!rec_off) return PGB_E_ARG;
    return index_lines_impl(nullptr, rec_off, prefix_off, nullptr, 0, prefix_base, n_lines, n_kept, 0, meta, scratch, stream);
}

extern "C" int pgb_dev_index_lines_ex(const uint32_t *var_row, const uint64_t *rec_off, uint64_t pitch,
                                      const uint64_t *prefix_off, const uint32_t *prefix_len, uint32_t suffix_len,
                                      uint64_t prefix_base, uint64_t n_lines, uint32_t n_kept, pgb_line_meta *meta,
                                      void *scratch, void *stream) {
    return index_lines_impl(rec_off ? nullptr : var_row, rec_off, prefix_off, prefix_len, suffix_len, prefix_base, n_lines,
                            n_kept, pitch, meta, scratch, stream);
}

template <bool GATHER, int HINT, int REPL, int IPW, bool ONE>
static int launch_k2_one(const pgb_k2_params &p, cudaStream_t st) {
    const uint64_t n_items = p.n_lines * (uint64_t)p.n_tiles;
    if (n_items == 0) return PGB_OK;
    const uint64_t per_cta = (uint64_t)K2_WARPS * IPW;
    const uint64_t blocks = (n_items + per_cta - 1) / per_cta;
    if (blocks > 0x7fffffffull) {
        pgb_set_error("K2 grid of %llu CTAs exceeds the launch limit", (unsigned long long)blocks);
        return PGB_E_ARG;
    }
    k2_format_kernel<GATHER, HINT, REPL, IPW, ONE><<<(unsigned)blocks, K2_THREADS, 0, st>>>(p);
    return check_launch("k2_format_kernel");
}

template <bool GATHER, int HINT, int REPL, int IPW>
static int launch_k2(const pgb_k2_params &p, cudaStream_t st) {
    return p.n_tiles == 1 ? launch_k2_one<GATHER, HINT, REPL, IPW, true>(p, st)
                          : launch_k2_one<GATHER, HINT, REPL, IPW, false>(p, st);
}

// Items per warp: 1 or 4 (the two values the launcher's defaults use; other requests round to the nearer one).
template <bool GATHER, int HINT, int REPL>
static int launch_k2_ipw(const pgb_k2_params &p, cudaStream_t st, int ipw) {
    return ipw >= 3 ? launch_k2<GATHER, HINT, REPL, 4>(p, st) : launch_k2<GATHER, HINT, REPL, 1>(p, st);
}

// Batch path (k2_batch.cuh): plan the shared-memory budget; returns false when the lines are too long for it.
static bool plan_k2_batch(uint32_t K, uint32_t R, bool gather, uint32_t max_prefix_len, uint32_t suffix_len, int variant,
                          pgb_k2b_params *bp, uint32_t *smem_bytes) {
    const int bsel = (variant >> 20) & 0xF, msel = (variant >> 24) & 0xF;
    const uint64_t max_line = (uint64_t)max_prefix_len + 4ull * K + 1ull;
    const uint64_t budget = msel ? (uint64_t)msel * 16384ull : 72ull * 1024ull; // default: three CTAs per SM
    const uint64_t rowcap = ((uint64_t)R + 31ull + 15ull) & ~15ull;
    const uint64_t pcap = ((uint64_t)(max_prefix_len - std::min(max_prefix_len, suffix_len)) + 31ull + 15ull) & ~15ull;
    const uint64_t vcap = gather ? (((uint64_t)(K + 3u) / 4ull + 2ull + 15ull) & ~15ull) : 0ull;
    // Measured (profiles/README.md).  Gather-heavy chr22 shape: the consumer warps storing to global memory directly
    // with three input stages of 24 lines (0.178 ms) beat one bulk-stored shared-memory image per warp with two
    // stages (0.186 ms: the image's 1 KB per line costs a third of the lines a batch can hold).  Keep-all lines of
    // 1.2-2.4 KB (records of 75-150 bytes, so the image is nearly all of the batch): the image wins, 0.276 / 0.322 /
    // 0.463 ms against 0.310 / 0.398 / 0.564 ms direct for 300 / 400 / 600 samples x 1 M variants.
    const int isel = (variant >> 28) & 3;
    const uint64_t images = isel == 1 ? 1 : isel == 2 ? 2 : isel == 3 ? 0 : (gather ? 0 : 1);
    const uint64_t stages = (((variant >> 30) & 1) != 0) == gather ? 2 : 3;
    const uint64_t per_line = stages * (rowcap + pcap + 32ull) + images * max_line;
    const uint64_t fixed = 64 + 128 + stages * 128 + 128 + images * K2B_WARPS * 160 + vcap * 16 + vcap * K2B_WARPS;
    if (budget <= fixed + 2 * per_line) return false;
    uint64_t B = (budget - fixed) / per_line;
    if (B > 32) B = 32;
    if (B > 8) B &= ~7ull; // the same number of lines for each of the eight consumer warps
    if (bsel) B = std::min<uint64_t>(B, 2ull * bsel);
    if (B < 2) return false;
    bp->K = K;
    bp->R = R;
    bp->B = (uint32_t)B;
    bp->rowcap = (uint32_t)rowcap;
    bp->pcap = (uint32_t)pcap;
    bp->vcap = (uint32_t)vcap;
    bp->wcap = images ? (uint32_t)((k2b_lines_per_warp((uint32_t)B) * max_line + 32ull + 127ull) & ~127ull) : 0u;
    bp->outcap = K2B_WARPS * bp->wcap;
    bp->images = (uint32_t)images;
    bp->stages = (uint32_t)stages;
    *smem_bytes = pgb_k2b_smem_layout(bp->B, bp->rowcap, bp->pcap, bp->vcap, bp->outcap, gather, bp->images, bp->stages).total;
    return *smem_bytes <= 227u * 1024u;
}

template <bool GATHER, bool IMG>
static int launch_k2_batch(pgb_k2b_params &bp, uint32_t smem_bytes, int variant, cudaStream_t st) {
    const uint64_t batches = (bp.n_lines + bp.B - 1) / bp.B;
    if (bp.n_lines > 0xffffffffull) return PGB_E_ARG;
    bp.n_batches = (uint32_t)batches;
    cudaError_t e = cudaFuncSetAttribute(k2_batch_kernel<GATHER, IMG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    int per_sm = 0, dev = 0, sms = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k2_batch_kernel<GATHER, IMG>, K2B_THREADS, smem_bytes);
    if (e == cudaSuccess) e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess || per_sm < 1 || sms < 1) {
        pgb_set_error("k2_batch_kernel launch configuration (%u bytes of shared memory): %s", smem_bytes, cudaGetErrorString(e));
        return PGB_E_CUDA;
    }
    const uint64_t grid0 = (uint64_t)per_sm * (uint64_t)sms; // persistent grid: every SM full
    uint64_t grid = grid0;
    if (grid > batches) grid = batches;
    k2_batch_kernel<GATHER, IMG><<<(unsigned)grid, K2B_THREADS, smem_bytes, st>>>(bp);
    return check_launch("k2_batch_kernel");
}

// variant: bits 0-3   (ignored; was the store hint: .cs streaming stores are always used)
//          bits 4-7   LUT copies (0 => default, see below; 1 => 8 interleaved bank-conflict-free copies; 2 => one)
//          bits 8-11  items per warp (0 => default; 1-2 => 1, 3 and more => 4)
//          bits 12-15 tile size = 4 KiB << n (0 => 16 KiB)
//          bits 16-19 batch path (k2_batch.cuh): 0 => default (gather launches whose batch fits), 1 => off, 2 => on
//                     whenever the batch fits (also keep-all)
//          bits 20-23 batch path: lines per batch <= 2 * n (0 => up to 32)
//          bits 24-27 batch path: shared-memory budget = n * 16 KiB (0 => 72 KiB, three CTAs per SM)
//          bits 28-29 batch path: 3 => the consumer warps store to global memory directly (default for gather); 1 / 2 =>
//                     one / two shared-memory images per warp, bulk-stored with cp.async.bulk.global.shared::cta
//                     (1 is the default for keep-all)
//          bit  30    batch path: the other number of input stages (default: three for gather, two for keep-all)
extern "C" int pgb_dev_format_lines_ex(const uint8_t *records, uint32_t record_bytes, const pgb_line_meta *meta,
                                       uint64_t n_lines, const uint8_t *prefix_blob, uint32_t suffix, uint32_t suffix_len,
                                       const uint32_t *kidx, uint32_t n_kept, uint32_t max_prefix_len, uint8_t *out,
                                       int variant, void *stream) {
    if (n_lines == 0) return PGB_OK;
    if (!records || !meta || !out || (!prefix_blob && max_prefix_len > suffix_len) || suffix_len > 4) return PGB_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const bool gatherp = kidx != nullptr;
    const uint32_t kidx_vec = kidx && ((uintptr_t)kidx & 15u) == 0 ? 1u : 0u;
    if (!gatherp) record_bytes = (n_kept + 3u) / 4u; // keep-all: n_kept is the file's sample count
    const int bmode = (variant >> 16) & 0xF;
    // Default use of the batch path (measured, profiles/README.md): every gather launch whose batch fits in shared
    // memory, and keep-all launches with lines of at most 3 KiB (300 samples: 0.065 vs 0.079 ms; 600: 0.49 vs 0.53 ms;
    // 1 000 samples, 4 KiB lines: 0.82 vs 0.71 ms — the per-line kernels win from there on).
    const uint64_t max_line_b = (uint64_t)max_prefix_len + 4ull * n_kept + 1ull;
    if (bmode != 1 && (record_bytes || n_kept == 0) && (gatherp || max_line_b <= 3072 || bmode == 2)) {
        pgb_k2b_params bp;
        uint32_t smem_bytes = 0;
        if (plan_k2_batch(n_kept, record_bytes, gatherp, max_prefix_len, suffix_len, variant, &bp, &smem_bytes)) {
            bp.records = records;
            bp.meta = meta;
            bp.prefix_blob = prefix_blob;
            bp.kidx = kidx;
            bp.out = out;
            bp.n_lines = n_lines;
            bp.sfx = suffix;
            bp.sfx_len = suffix_len;
            bp.kidx_vec = kidx_vec;
            if (bp.images)
                return gatherp ? launch_k2_batch<true, true>(bp, smem_bytes, variant, st)
                               : launch_k2_batch<false, true>(bp, smem_bytes, variant, st);
            return gatherp ? launch_k2_batch<true, false>(bp, smem_bytes, variant, st)
                           : launch_k2_batch<false, false>(bp, smem_bytes, variant, st);
        }
    }
    pgb_k2_params p;
    p.records = records;
    p.meta = meta;
    p.prefix_blob = prefix_blob;
    p.kidx = kidx;
    p.out = out;
    p.n_lines = n_lines;
    p.K = n_kept;
    p.sfx = suffix;
    p.sfx_len = suffix_len;
    const int hint = (variant & 0xF) == 2 ? 0 : 1;
    const int lsel = (variant >> 4) & 0xF;
    int ipw = (variant >> 8) & 0xF;
    const int tsel = (variant >> 12) & 0xF;
    p.tile_bytes = tsel ? (4096u << tsel) : 16384u;
    const uint64_t max_line = (uint64_t)max_prefix_len + 4ull * n_kept + 1ull;
    const uint64_t nt = (max_line + 511ull + p.tile_bytes - 1) / p.tile_bytes;
    if (nt > 0x7fffffffull) return PGB_E_ARG;
    p.n_tiles = (uint32_t)nt;
    // Defaults measured on B200 (profiles/README.md, 2.5 GB of output per shape): lines > 7 KiB (and tiled
    // lines) -> 8 conflict-free LUT copies, one item per warp; 3-7 KiB -> one LUT copy (the 32 KB table costs
    // more to build than it saves), one item per warp; <= 3 KiB and gather -> one copy, four items per warp.
    const bool small_lines = nt == 1 && max_line <= 3072, mid_lines = nt == 1 && max_line <= 7168;
    const bool repl8 = lsel == 1 || (lsel == 0 && !gatherp && !mid_lines);
    if (ipw == 0) ipw = (gatherp || small_lines) ? 4 : 1;
    // bytes of a record worth prefetching: all of it for keep-all, up to the last kept sample otherwise
    p.row_bytes_hint = kidx ? 0u : (n_kept + 3u) / 4u + 1u;
    p.kidx_vec = kidx_vec;
    // 16 instantiations: {keep-all, gather} x {8, 1 LUT copies} x {1, 4 items per warp} x {whole lines, tiles}; all
    // use .cs streaming stores (the default write-back variant measured 4 % slower in round 1 and was dropped)
    (void)hint;
    const bool g = gatherp;
    if (repl8) return g ? launch_k2_ipw<true, 1, 8>(p, st, ipw) : launch_k2_ipw<false, 1, 8>(p, st, ipw);
    return g ? launch_k2_ipw<true, 1, 1>(p, st, ipw) : launch_k2_ipw<false, 1, 1>(p, st, ipw);
}

extern "C" int pgb_dev_format_lines(const uint8_t *records, const pgb_line_meta *meta, uint64_t n_lines,
                                    const uint8_t *prefix_blob, const uint32_t *kidx, uint32_t n_kept,
                                    uint32_t max_prefix_len, uint8_t *out, int variant, void *stream) {
    return pgb_dev_format_lines_ex(records, 0, meta, n_lines, prefix_blob, 0, 0, kidx, n_kept, max_prefix_len, out, variant,
                                   stream);
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

extern "C" int pgb_dev_synth_records_fast(uint8_t *records, uint64_t pitch, uint32_t seed, uint64_t row0, uint64_t n_rows,
                                          uint32_t n_samples, void *stream) {
    const uint32_t R = pgb_record_bytes(n_samples);
    if (!records || pitch < R) return PGB_E_ARG;
    const uint64_t total = n_rows * (uint64_t)((R + 3u) / 4u);
    if (total == 0) return PGB_OK;
    uint64_t blocks = (total + 255) / 256;
    if (blocks > 148ull * 64) blocks = 148ull * 64;
    synth_fast_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(records, pitch, seed, row0, n_rows, n_samples, R);
    return check_launch("synth_fast_kernel");
}

// variant: bits 0-3 store hint (0 default, 1 .cs streaming).
extern "C" int pgb_dev_fill(uint8_t *dst, uint64_t bytes, int variant, void *stream) {
    if (!dst || ((uintptr_t)dst & 15)) return PGB_E_ARG;
    const uint64_t n16 = bytes / 16;
    if (!n16) return PGB_OK;
    uint64_t blocks = (n16 + 255) / 256;
    if (blocks > 148ull * 32) blocks = 148ull * 32;
    fill_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dst, n16, variant & 0xF);
    return check_launch("fill_kernel");
}

extern "C" int pgb_dev_fill_pattern(uint8_t *dst, uint64_t bytes, uint32_t seg_rows, uint32_t coop, uint32_t burst,
                                    uint32_t n_ctas, int hint, void *stream) {
    if (!dst || ((uintptr_t)dst & 15) || !n_ctas || !burst) return PGB_E_ARG;
    if (coop != 1 && coop != 2 && coop != 4 && coop != 8) return PGB_E_ARG;
    if (seg_rows == 0) {
        fill_strided_kernel<<<n_ctas, 256, 0, (cudaStream_t)stream>>>(dst, bytes / 16, hint, burst);
        return check_launch("fill_strided_kernel");
    }
    const uint64_t n_seg = bytes / (seg_rows * 512ull);
    if (!n_seg) return PGB_OK;
    fill_segments_kernel<<<n_ctas, 256, 0, (cudaStream_t)stream>>>(dst, n_seg, seg_rows, hint, coop, burst);
    return check_launch("fill_segments_kernel");
}
