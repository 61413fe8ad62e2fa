// gk_refine.cu -- prefix-doubling refinement for k-mers longer than one key word (SURVEY.md 8a A4,
// north_star subsystem 2 "extended with prefix-doubling refinement").
//
// After the windows are sorted by their first h symbols (head flags mark the h-groups), the rank
// of a start is the sorted position of its group's first member.  A window of h2 <= 2h symbols is
// then ordered by the pair (rank_h[start], rank_h[start + h2 - h]): both halves are h-windows that
// lie inside the same record.  Only members of groups with more than one element can move, so the
// pair sort runs on that subset and the rest keeps its slot.
//
//   head_positions   group id (= position of the group's head) for every sorted position, and
//                    rank_of_start[idx[p]] = that id.  Tile max-scan, 3 kernels.
//   valid_flags      which windows still fit their record at the next length
//   gid_flags        head flags after compaction (group id changes) + "group has >1 member" marks
//   pair_keys        key2 = (gid << 32) | rank_of_start[start + delta]
//   key2_flags       head flags inside the re-sorted subset, scattered to their slots
// Indices are 32-bit here: the multi-level path is limited to byte arrays below 2^32.
#include "gk_common.cuh"

namespace gk {

constexpr uint8_t kFlagHead = 1;
constexpr uint8_t kFlagPass = 4;
constexpr uint8_t kFlagMulti = 8;  // member of a group with more than one element

constexpr int kHpThreads = 256;
constexpr int kHpPerThread = 16;
constexpr int kHpTile = kHpThreads * kHpPerThread;

// last head position inside each tile (kNoHead if none)
__global__ void __launch_bounds__(kHpThreads)
tile_last_head_kernel(const uint8_t *__restrict__ flags, uint64_t n, uint32_t *__restrict__ tile_last)
{
    __shared__ uint32_t s_max[kHpThreads / 32];
    const uint64_t p0 = (uint64_t)blockIdx.x * kHpTile + (uint64_t)threadIdx.x * kHpPerThread;
    uint32_t last = 0;  // store position+1 so that 0 means "none"
    for (int i = 0; i < kHpPerThread; ++i) {
        const uint64_t p = p0 + i;
        if (p < n && (flags[p] & kFlagHead)) last = (uint32_t)p + 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t v = __shfl_xor_sync(0xffffffffu, last, o);
        last = v > last ? v : last;
    }
    if (lane_id() == 0) s_max[threadIdx.x >> 5] = last;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t m = 0;
        for (int w = 0; w < kHpThreads / 32; ++w) m = s_max[w] > m ? s_max[w] : m;
        tile_last[blockIdx.x] = m;
    }
}

// one CTA: carry[t] = last head (+1) in tiles < t (exclusive running max)
__global__ void __launch_bounds__(1024)
tile_carry_kernel(const uint32_t *__restrict__ tile_last, uint64_t n_tiles, uint32_t *__restrict__ carry)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    for (uint64_t base = 0; base < n_tiles; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const uint32_t v = (i < n_tiles) ? tile_last[i] : 0;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc = u > inc ? u : inc;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t pre = s_carry;
        for (uint32_t w = 0; w < warp; ++w) pre = s_warp[w] > pre ? s_warp[w] : pre;
        // exclusive: max over everything strictly before i
        uint32_t excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = 0;
        excl = excl > pre ? excl : pre;
        if (i < n_tiles) carry[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = inc > pre ? inc : pre;
        __syncthreads();
    }
}

// gid[p] = position of the head of p's group; optionally rank_of_start[idx[p]] = gid[p]
__global__ void __launch_bounds__(kHpThreads)
head_positions_kernel(const uint8_t *__restrict__ flags, const uint32_t *__restrict__ idx, uint64_t n,
                      const uint32_t *__restrict__ carry, uint32_t *__restrict__ gid,
                      uint32_t *__restrict__ rank_of_start)
{
    __shared__ uint32_t s_warp[kHpThreads / 32];
    const uint64_t p0 = (uint64_t)blockIdx.x * kHpTile + (uint64_t)threadIdx.x * kHpPerThread;
    uint32_t local[kHpPerThread];
    uint32_t run = 0;  // position+1 of the latest head seen by this thread
#pragma unroll
    for (int i = 0; i < kHpPerThread; ++i) {
        const uint64_t p = p0 + i;
        if (p < n && (flags[p] & kFlagHead)) run = (uint32_t)p + 1;
        local[i] = run;
    }
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    uint32_t inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc = u > inc ? u : inc;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t pre = carry[blockIdx.x];
    for (uint32_t w = 0; w < warp; ++w) pre = s_warp[w] > pre ? s_warp[w] : pre;
    uint32_t excl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) excl = 0;
    excl = excl > pre ? excl : pre;
#pragma unroll
    for (int i = 0; i < kHpPerThread; ++i) {
        const uint64_t p = p0 + i;
        if (p < n) {
            const uint32_t h = (local[i] > excl ? local[i] : excl) - 1;  // position 0 is always a head
            gid[p] = h;
            if (rank_of_start) rank_of_start[idx[p]] = h;
        }
    }
}

// flags[p] = kFlagPass iff the window of `len` symbols at idx[p] stays inside its record
__global__ void __launch_bounds__(256)
valid_flags_kernel(const uint32_t *__restrict__ idx, uint64_t n, const uint64_t *__restrict__ seg_starts,
                   uint32_t n_seg, uint64_t sba_len, uint32_t len, uint8_t *__restrict__ flags)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        const uint64_t s = idx[p];
        const uint32_t seg = upper_seg(seg_starts, n_seg, s);
        const uint64_t seg_end = (seg + 1 < n_seg) ? seg_starts[seg + 1] - 1 : sba_len;
        flags[p] = (s + len <= seg_end) ? kFlagPass : 0;
    }
}

// head flag where the group id changes; kFlagMulti on every member of a group with > 1 element
__global__ void __launch_bounds__(256)
gid_flags_kernel(const uint32_t *__restrict__ gid, uint64_t n, uint8_t *__restrict__ flags)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        const uint32_t g = gid[p];
        const bool head = (p == 0) || gid[p - 1] != g;
        const bool next_head = (p + 1 == n) || gid[p + 1] != g;
        flags[p] = (head ? kFlagHead : 0) | ((head && next_head) ? 0 : kFlagMulti);
    }
}

__global__ void __launch_bounds__(256)
pair_keys_kernel(const uint32_t *__restrict__ sub_idx, const uint32_t *__restrict__ sub_gid, uint64_t m,
                 const uint32_t *__restrict__ rank_of_start, uint32_t delta, uint64_t *__restrict__ keys)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += stride)
        keys[r] = ((uint64_t)sub_gid[r] << 32) | rank_of_start[(uint64_t)sub_idx[r] + delta];
}

// the re-sorted subset goes back to its slots with fresh head flags
__global__ void __launch_bounds__(256)
key2_scatter_kernel(const uint64_t *__restrict__ keys_sorted, const uint32_t *__restrict__ sub_idx_sorted,
                    const uint32_t *__restrict__ slots, uint64_t m, uint32_t *__restrict__ idx,
                    uint8_t *__restrict__ flags)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += stride) {
        const bool head = (r == 0) || keys_sorted[r] != keys_sorted[r - 1];
        const uint32_t slot = slots[r];
        idx[slot] = sub_idx_sorted[r];
        flags[slot] = head ? kFlagHead : 0;
    }
}

static int grid_for(uint64_t items)
{
    uint64_t blocks = (items + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int head_positions_device(const uint8_t *d_flags, const uint32_t *d_idx, uint64_t n, uint32_t *d_gid,
                          uint32_t *d_rank_of_start, cudaStream_t st)
{
    if (n == 0) return GK_OK;
    const uint64_t tiles = (n + kHpTile - 1) / kHpTile;
    DeviceBuffer temp;
    GK_TRY(temp.alloc((size_t)tiles * 8, st));
    uint32_t *d_last = temp.as<uint32_t>();
    uint32_t *d_carry = d_last + tiles;
    tile_last_head_kernel<<<(unsigned)tiles, kHpThreads, 0, st>>>(d_flags, n, d_last);
    GK_LAUNCH_CHECK();
    tile_carry_kernel<<<1, 1024, 0, st>>>(d_last, tiles, d_carry);
    GK_LAUNCH_CHECK();
    head_positions_kernel<<<(unsigned)tiles, kHpThreads, 0, st>>>(d_flags, d_idx, n, d_carry, d_gid,
                                                                  d_rank_of_start);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int valid_flags_device(const uint32_t *d_idx, uint64_t n, const uint64_t *d_seg_starts, uint32_t n_seg,
                       uint64_t sba_len, uint32_t len, uint8_t *d_flags, cudaStream_t st)
{
    if (n == 0) return GK_OK;
    valid_flags_kernel<<<grid_for(n), 256, 0, st>>>(d_idx, n, d_seg_starts, n_seg, sba_len, len, d_flags);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int gid_flags_device(const uint32_t *d_gid, uint64_t n, uint8_t *d_flags, cudaStream_t st)
{
    if (n == 0) return GK_OK;
    gid_flags_kernel<<<grid_for(n), 256, 0, st>>>(d_gid, n, d_flags);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int pair_keys_device(const uint32_t *d_sub_idx, const uint32_t *d_sub_gid, uint64_t m,
                     const uint32_t *d_rank_of_start, uint32_t delta, uint64_t *d_keys, cudaStream_t st)
{
    if (m == 0) return GK_OK;
    pair_keys_kernel<<<grid_for(m), 256, 0, st>>>(d_sub_idx, d_sub_gid, m, d_rank_of_start, delta, d_keys);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int key2_scatter_device(const uint64_t *d_keys_sorted, const uint32_t *d_sub_idx_sorted,
                        const uint32_t *d_slots, uint64_t m, uint32_t *d_idx, uint8_t *d_flags,
                        cudaStream_t st)
{
    if (m == 0) return GK_OK;
    key2_scatter_kernel<<<grid_for(m), 256, 0, st>>>(d_keys_sorted, d_sub_idx_sorted, d_slots, m, d_idx,
                                                     d_flags);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

}  // namespace gk
