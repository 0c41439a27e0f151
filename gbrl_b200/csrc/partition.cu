// partition.cu -- stable per-node partition of the row order after a level's splits.
//
// Restates TreeNode::splitNode (node.cpp:64-149): rows with x[f] > thr go to the right child, the others
// to the left, each side keeping ascending sample order.  All nodes of a level are partitioned by one
// global pass: flag[k] = row at position k goes right; R = exclusive prefix sum of flag over the whole
// order array; for a row of node h at position k
//     left :  new = seg_start(h) + (k - seg_start(h)) - (R[k] - R[seg_start(h)])
//     right:  new = seg_start(h) + n_left(h)          + (R[k] - R[seg_start(h)])
// Rows of nodes that are not split at this level keep their position.  The comparison is done on the raw
// fp32 feature value, exactly like the reference.
#include "engine.cuh"

namespace gb {

constexpr int PART_CHUNK = 2048;   // rows per CTA (256 threads x 8)

__global__ void __launch_bounds__(256)
part_flag_kernel(const float *__restrict__ X, int F, const uint16_t *__restrict__ codesT, long long stride, int row_offset,
                 const int *__restrict__ order, const int *__restrict__ nid, NodeArrays na, uint8_t *__restrict__ flag,
                 int *__restrict__ chunk_sums, int N) {
    __shared__ int s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const int k0 = blockIdx.x * PART_CHUNK;
    int mine = 0;
    for (int k = k0 + threadIdx.x; k < min(N, k0 + PART_CHUNK); k += 256) {
        const int i = order[k];
        const int h = nid[i];
        uint8_t fl = 0;
        // x > thr[f][j]  <=>  code(x) > j  (candidates.cu): 2 coalesced-ish bytes of the feature-major codes per row
        if (na.state[h] == NODE_SPLIT) {
            if (codesT != nullptr) fl = (int)codesT[(size_t)na.split_f[h] * stride + row_offset + i] > na.split_j[h] ? 1 : 0;
            else fl = X[(size_t)i * F + na.split_f[h]] > na.split_thr[h] ? 1 : 0;
        }
        flag[k] = fl;
        mine += fl;
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_sum, mine);
    __syncthreads();
    if (threadIdx.x == 0) chunk_sums[blockIdx.x] = s_sum;
}

// exclusive scan of the chunk sums (single CTA, sequential over tiles of 1024)
__global__ void __launch_bounds__(1024) part_scan_chunks_kernel(int *chunk_sums, int n_chunks) {
    __shared__ int s[1024];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int b = 0; b < n_chunks; b += 1024) {
        const int i = b + threadIdx.x;
        const int v = i < n_chunks ? chunk_sums[i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < n_chunks) chunk_sums[i] = s_carry + s[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry += s[1023];
        __syncthreads();
    }
}

// R[k] for every position (each thread owns 8 consecutive positions)
__global__ void __launch_bounds__(256)
part_rscan_kernel(const uint8_t *__restrict__ flag, const int *__restrict__ chunk_sums, int *__restrict__ rscan, int N) {
    __shared__ int s_warp[8];
    const int k0 = blockIdx.x * PART_CHUNK + threadIdx.x * 8;
    int f[8], loc = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { f[j] = (k0 + j < N) ? flag[k0 + j] : 0; loc += f[j]; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = loc;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += s_warp[w];
    int run = chunk_sums[blockIdx.x] + wbase + inc - loc;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (k0 + j < N) rscan[k0 + j] = run;
        run += f[j];
    }
}

__global__ void __launch_bounds__(256)
part_scatter_kernel(const int *__restrict__ order_in, int *__restrict__ order_out, int *__restrict__ nid,
                    const uint8_t *__restrict__ flag, const int *__restrict__ rscan, NodeArrays na, int N) {
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= N) return;
    const int i = order_in[k];
    const int h = nid[i];
    if (na.state[h] != NODE_SPLIT) { order_out[k] = i; return; }
    const int s0 = na.seg_start[h];
    const int rbefore = rscan[k] - rscan[s0];
    if (flag[k]) {
        const int nl = na.seg_len[2 * h + 1];
        order_out[s0 + nl + rbefore] = i;
        nid[i] = 2 * h + 2;
    } else {
        order_out[s0 + (k - s0) - rbefore] = i;
        nid[i] = 2 * h + 1;
    }
}

// after this call ws.order[0] is the new current order
void launch_partition(Model &m, const float *X, int level, int cur, cudaStream_t s) {
    (void)level; (void)cur;
    Workspace &ws = m.ws;
    const int N = ws.N;
    if (N == 0) return;
    const int n_chunks = ceil_div(N, PART_CHUNK);
    GB_LAUNCH(part_flag_kernel, n_chunks, 256, 0, s, X, ws.F, ws.use_codesT ? ws.codesT.as<uint16_t>() : nullptr, ws.codesT_stride,
              ws.row_offset, ws.order[0].as<int>(),
              ws.nid.as<int>(), ws.na, ws.rflag.as<uint8_t>(), ws.chunk_sums.as<int>(), N);
    GB_LAUNCH(part_scan_chunks_kernel, 1, 1024, 0, s, ws.chunk_sums.as<int>(), n_chunks);
    GB_LAUNCH(part_rscan_kernel, n_chunks, 256, 0, s, ws.rflag.as<uint8_t>(), ws.chunk_sums.as<int>(), ws.rscan.as<int>(), N);
    GB_LAUNCH(part_scatter_kernel, ceil_div(N, 256), 256, 0, s, ws.order[0].as<int>(), ws.order[1].as<int>(), ws.nid.as<int>(),
              ws.rflag.as<uint8_t>(), ws.rscan.as<int>(), ws.na, N);
    std::swap(ws.order[0].p, ws.order[1].p);
    std::swap(ws.order[0].bytes, ws.order[1].bytes);
}

}  // namespace gb
