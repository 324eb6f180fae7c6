// Bucket sort of the (digit, point) pairs of a Pippenger pass, staged through shared memory (north_star (2):
// "signed-digit window decomposition, a radix sort of (digit, point-index) pairs staged through shared memory").
// Replaces the first version of the pipeline (k_digits -> k_scan -> k_scatter in msm.cuh, still used for short MSMs):
// that one wrote every digit to HBM, read it back, and paid one global atomic plus one random 4-byte write per pair
// (13.6 M of each at 2^20 points; k_digits ran at 22 % and k_scatter at 11 % of DRAM throughput,
// profiles/r01e_tail_kernels_ncu.csv).
//
// Two-level MSD radix sort on the bucket key, most significant digit first:
//   pass 0  k_sort_tiles<COUNT>   a CTA owns a tile of SORT_TILE scalars: digits are produced in registers (scalar source ->
//                                 from-Montgomery -> signed radix-2^c digits) and counted per PARTITION (the top bits of the
//                                 bucket key) in a shared-memory histogram; the histogram row goes to tile_count[tile][*]
//           k_sort_tile_scan      column-wise exclusive scan of that matrix: position of every (tile, partition) run inside
//                                 its partition, and the partition sizes; the existing k_scan turns those into offsets
//   pass 1  k_sort_tiles<WRITE>   the same tile, digits recomputed (cheaper than a round trip of 13 words per scalar
//                                 through HBM): shared-memory cursors, preset to the global position of the tile's run in each
//                                 partition, hand out slots; (key low bits, table index | sign) pairs are written as
//                                 contiguous runs -- no global atomics, streaming writes
//   pass 2  k_sort_buckets        one CTA per partition: histogram of the key's low bits in shared memory, exclusive scan
//                                 -> the bucket offsets k_accumulate reads (written coalesced), then the pairs are
//                                 scattered through shared-memory cursors into their final order inside the partition's
//                                 own (L2-resident) output range
// Output: exactly what k_scatter produced -- offsets[0 .. nkeys] and entries[] grouped by bucket -- so the accumulation
// and reduction kernels are unchanged.  The order of the entries INSIDE a bucket differs from run to run (atomics);
// the bucket sums, and with them every result, do not (the group is commutative, outputs are normalised).
// Degenerate inputs (constant scalar vectors: all lanes of a warp land in one bin) take one aggregated atomic per warp;
// when one partition collects far more than its share (the same constant vectors: 13 partitions hold everything) a single CTA
// per partition would serialise the bucket pass, so the tile scan records the largest partition and, above a threshold,
// the write and bucket passes stand down and the first-version sort (global atomics aggregated per warp, built for exactly
// that case) runs instead -- decided on the device through a gate word, no host round trip.
#pragma once
#include "msm.cuh"

namespace accmsm {

#ifndef ACC_SORT_TILE
#define ACC_SORT_TILE 1024
#endif
#ifndef ACC_SORT_THREADS
#define ACC_SORT_THREADS 512
#endif
constexpr int SORT_TILE = ACC_SORT_TILE;        // scalars per tile (x windows pairs: 13.3 K pairs = 104 KB of staging at 13 windows; measured: 1024 / 512 threads beats 512 / 256)
constexpr int SORT_THREADS = ACC_SORT_THREADS;
constexpr int SORT_BUCKET_THREADS = 512;
constexpr uint32_t SORT_MAX_PARTS = 8192;   // partitions = shared-memory histogram bins of the tile passes
constexpr uint32_t SORT_MAX_LB = 12;        // low key bits = shared-memory bins of the bucket pass (<= 4096)
constexpr uint32_t SORT_STAGE_MAX = 24576;  // pairs a tile may stage in shared memory (192 KB); more: direct scattered writes

struct SortPlan {
    uint32_t lb;              // low key bits resolved inside a partition
    uint32_t nparts;          // ceil(nkeys / 2^lb)
    uint32_t tiles_per_job;   // ceil(n / SORT_TILE)
    uint32_t ntiles;          // njobs * tiles_per_job
    uint32_t stage_pairs;     // staging capacity of the write pass (0 = no staging), SORT_TILE * nwin when it fits
    uint32_t bucket_cap;      // staging capacity (entries) of the bucket pass
};

// slot for one pair in bin `bin` of a shared-memory counter array; all lanes of the warp that are active call it.
// Warp-uniform bins (constant scalar vectors) take ONE atomic for the whole warp.
ACC_D uint32_t sort_take_slot(uint32_t *bins, uint32_t bin, bool valid) {
    const unsigned am = __activemask();
    const uint32_t lane = threadIdx.x & 31u, leader = __ffs(am) - 1;
    const uint32_t first = __shfl_sync(am, bin, leader);
    const bool uniform = __all_sync(am, valid && bin == first);
    uint32_t pos = 0;
    if (uniform) {
        if (lane == leader) pos = atomicAdd(&bins[bin], (uint32_t)__popc(am));
        pos = __shfl_sync(am, pos, leader) + __popc(am & ((1u << lane) - 1u));
    } else if (valid) {
        pos = atomicAdd(&bins[bin], 1u);
    }
    return pos;
}

// in-place exclusive scan of a[0 .. n) in shared memory by the whole CTA (blockDim.x = NT, n <= NT * 32); `scratch` holds
// NT / 32 words.  Returns the total.  Ends with a barrier.
template <int NT> ACC_D uint32_t sort_block_scan(uint32_t *a, uint32_t n, uint32_t *scratch) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t per = (n + NT - 1) / NT, k0 = tid * per;
    uint32_t sum = 0;
    for (uint32_t j = 0; j < per; j++) if (k0 + j < n) sum += a[k0 + j];
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += o; }
    if (lane == 31) scratch[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < NT / 32 ? scratch[lane] : 0u, wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= (uint32_t)d) wi += o; }
        if (lane < NT / 32) scratch[lane] = wi - w;
        if (lane == NT / 32 - 1) scratch[NT / 32] = wi;
    }
    __syncthreads();
    uint32_t run = scratch[wid] + incl - sum;
    for (uint32_t j = 0; j < per; j++) if (k0 + j < n) { uint32_t v = a[k0 + j]; a[k0 + j] = run; run += v; }
    const uint32_t total = scratch[NT / 32];
    __syncthreads();
    return total;
}

// WRITE = false: tile_hist[tile][p] = tile_count[tile][p] = pairs of the tile that fall into partition p
// WRITE = true : the tile's pairs are grouped by partition in shared memory (local offsets = scan of the tile's own
//                histogram row) and copied out as contiguous runs to part_offs[p] + tile_count[tile][p] (tile_count now
//                holds the exclusive column scan): (key & (2^lb - 1), entry) pairs, coalesced stores
template <class Src, bool WRITE>
__global__ void __launch_bounds__(SORT_THREADS) k_sort_tiles(Src src, MsmShape sh, const uint8_t *__restrict__ base_is_identity,
                                                               SortPlan pl, uint32_t t_begin, uint32_t t_end,
                                                               uint32_t *__restrict__ tile_count, uint32_t *__restrict__ tile_hist,
                                                               const uint32_t *__restrict__ part_offs, uint2 *__restrict__ pairs, SortGate gate) {
    extern __shared__ uint32_t sort_smem[];
    if (WRITE && gate.max_part && *gate.max_part > gate.thr) return;     // skewed input: the first-version sort takes over
    __shared__ uint32_t scan_scratch[SORT_THREADS / 32 + 1];
    uint32_t *cursor = sort_smem;                       // nparts
    uint32_t *delta = sort_smem + pl.nparts;            // nparts (WRITE): global position - local position of a partition's run
    uint2 *stage = reinterpret_cast<uint2 *>(sort_smem + 2 * pl.nparts);
    // tiles [t_begin, t_end) of every job: the whole vector, or the tiles of one chunk of a host upload still in flight
    const uint32_t per_job = t_end - t_begin;
    const uint32_t job = blockIdx.x / per_job, t = t_begin + blockIdx.x % per_job;
    const uint32_t tile = job * pl.tiles_per_job + t;
    const size_t row = (size_t)tile * pl.nparts;
    const bool staged = WRITE && pl.stage_pairs != 0;
    uint32_t total = 0;
    if (WRITE) {
        for (uint32_t p = threadIdx.x; p < pl.nparts; p += SORT_THREADS) cursor[p] = staged ? tile_hist[row + p] : part_offs[p] + tile_count[row + p];
        __syncthreads();
        if (staged) {
            total = sort_block_scan<SORT_THREADS>(cursor, pl.nparts, scan_scratch);
            for (uint32_t p = threadIdx.x; p < pl.nparts; p += SORT_THREADS) delta[p] = part_offs[p] + tile_count[row + p] - cursor[p];
            __syncthreads();
        }
    } else {
        for (uint32_t p = threadIdx.x; p < pl.nparts; p += SORT_THREADS) cursor[p] = 0u;
        __syncthreads();
    }
    const uint32_t i0 = t * SORT_TILE, i1 = min(sh.n, i0 + SORT_TILE);
    const uint32_t half = 1u << (sh.c - 1), key_job = job * sh.sets_per_job * sh.nb, lo_mask = (1u << pl.lb) - 1u;
    for (uint32_t i = i0 + threadIdx.x; i < i1; i += SORT_THREADS) {
        fe_t s = src.canonical(job, i);
        const uint32_t base_index = msm_base_index(sh, job, i);
        if (base_is_identity && base_is_identity[base_index]) s = Fp<0>::zero();     // identity bases contribute nothing
        uint32_t carry = 0;
        for (uint32_t w = 0; w < sh.nwin; w++) {
            const uint32_t raw = extract_bits(s.l, w * sh.c, sh.c) + carry;
            uint32_t mag, sign = 0;
            if (raw > half) { mag = (1u << sh.c) - raw; carry = 1; sign = 0x80000000u; }   // negative digit, borrow from the next window
            else { mag = raw; carry = 0; }
            const uint32_t key = key_job + w * sh.hist_stride + mag - 1u;       // only meaningful when mag != 0
            const uint32_t pos = sort_take_slot(cursor, mag ? key >> pl.lb : 0u, mag != 0);
            if (WRITE && mag) {
                const uint32_t entry = (w * sh.ent_stride + base_index) | sign;
                if (staged) stage[pos] = make_uint2(key, entry);
                else pairs[pos] = make_uint2(key & lo_mask, entry);
            }
        }
    }
    __syncthreads();
    if (!WRITE) {
        for (uint32_t p = threadIdx.x; p < pl.nparts; p += SORT_THREADS) { const uint32_t v = cursor[p]; tile_count[row + p] = v; tile_hist[row + p] = v; }
    } else if (staged) {
        // consecutive staged pairs of one partition go to consecutive global slots: coalesced, mostly full sectors
        for (uint32_t idx = threadIdx.x; idx < total; idx += SORT_THREADS) {
            const uint2 pr = stage[idx];
            pairs[idx + delta[pr.x >> pl.lb]] = make_uint2(pr.x & lo_mask, pr.y);
        }
    }
}

// Column-wise exclusive scan of tile_count[ntiles][nparts] (in place) and the column totals part_count[nparts].
// Block = 8 partitions x 128 tile segments: segment sums -> scan over the segments in shared memory -> running write.
constexpr uint32_t TS_PARTS = 8, TS_SEGS = 128;
__global__ void __launch_bounds__(TS_PARTS * TS_SEGS) k_sort_tile_scan(uint32_t *__restrict__ tile_count, uint32_t ntiles, uint32_t nparts,
                                                                        uint32_t *__restrict__ part_count, uint32_t *__restrict__ max_part) {
    __shared__ uint32_t seg_sum[TS_SEGS][TS_PARTS + 1];
    const uint32_t px = threadIdx.x % TS_PARTS, ty = threadIdx.x / TS_PARTS, p = blockIdx.x * TS_PARTS + px;
    const uint32_t seg = (ntiles + TS_SEGS - 1) / TS_SEGS, a = min(ntiles, ty * seg), b = min(ntiles, a + seg);
    uint32_t sum = 0;
    if (p < nparts) for (uint32_t t = a; t < b; t++) sum += tile_count[(size_t)t * nparts + p];
    seg_sum[ty][px] = sum;
    __syncthreads();
    if (ty == 0) {
        uint32_t run = 0;
        for (uint32_t j = 0; j < TS_SEGS; j++) { uint32_t v = seg_sum[j][px]; seg_sum[j][px] = run; run += v; }
        if (p < nparts) { part_count[p] = run; atomicMax(max_part, run); }
    }
    __syncthreads();
    uint32_t run = seg_sum[ty][px];
    if (p < nparts) for (uint32_t t = a; t < b; t++) {
        const size_t idx = (size_t)t * nparts + p;
        const uint32_t v = tile_count[idx];
        tile_count[idx] = run;
        run += v;
    }
}

// One CTA per partition: pairs[part_offs[p] .. part_offs[p + 1]) -> entries in bucket order + bucket offsets.
// Partitions of up to bucket_cap pairs are permuted in shared memory and written out coalesced; larger ones (skewed
// scalar distributions) scatter straight to HBM.
__global__ void __launch_bounds__(SORT_BUCKET_THREADS) k_sort_buckets(const uint2 *__restrict__ pairs, const uint32_t *__restrict__ part_offs,
                                                                       uint32_t lb, uint32_t nkeys, uint32_t bucket_cap,
                                                                       uint32_t *__restrict__ offsets, uint32_t *__restrict__ entries, SortGate gate) {
    extern __shared__ uint32_t sort_smem[];
    if (gate.max_part && *gate.max_part > gate.thr) return;
    __shared__ uint32_t scan_scratch[SORT_BUCKET_THREADS / 32 + 1];
    const uint32_t p = blockIdx.x, nb = 1u << lb, tid = threadIdx.x;
    uint32_t *bins = sort_smem, *stage = sort_smem + nb;
    const uint32_t b0 = part_offs[p], b1 = part_offs[p + 1], m = b1 - b0;
    const bool staged = m <= bucket_cap;
    for (uint32_t k = tid; k < nb; k += SORT_BUCKET_THREADS) bins[k] = 0;
    __syncthreads();
    constexpr uint32_t U = 4;                                          // independent loads in flight per thread
    for (uint32_t e0 = b0; e0 < b1; e0 += U * SORT_BUCKET_THREADS) {   // whole warps stay in step (aggregated atomics)
        uint32_t k[U]; bool valid[U];
#pragma unroll
        for (uint32_t u = 0; u < U; u++) { const uint32_t e = e0 + u * SORT_BUCKET_THREADS + tid; valid[u] = e < b1; k[u] = valid[u] ? __ldg(&pairs[e].x) : 0u; }
#pragma unroll
        for (uint32_t u = 0; u < U; u++) sort_take_slot(bins, k[u], valid[u]);
    }
    __syncthreads();
    sort_block_scan<SORT_BUCKET_THREADS>(bins, nb, scan_scratch);     // bins = offsets of the buckets inside the partition
    const uint32_t key0 = p << lb;
    for (uint32_t k = tid; k < nb; k += SORT_BUCKET_THREADS) if (key0 + k < nkeys) offsets[key0 + k] = b0 + bins[k];
    if (p == gridDim.x - 1 && tid == 0) offsets[nkeys] = b1;
    __syncthreads();
    for (uint32_t e0 = b0; e0 < b1; e0 += U * SORT_BUCKET_THREADS) {
        uint2 pr[U]; bool valid[U];
#pragma unroll
        for (uint32_t u = 0; u < U; u++) { const uint32_t e = e0 + u * SORT_BUCKET_THREADS + tid; valid[u] = e < b1; pr[u] = valid[u] ? __ldg(&pairs[e]) : make_uint2(0u, 0u); }
#pragma unroll
        for (uint32_t u = 0; u < U; u++) {
            const uint32_t pos = sort_take_slot(bins, pr[u].x, valid[u]);
            if (valid[u]) { if (staged) stage[pos] = pr[u].y; else entries[b0 + pos] = pr[u].y; }
        }
    }
    if (staged) {
        __syncthreads();
        for (uint32_t idx = tid; idx < m; idx += SORT_BUCKET_THREADS) entries[b0 + idx] = stage[idx];
    }
}

}  // namespace accmsm
