// partition.cu -- stable per-node partition of the row order after a level's splits.
//
// Restates TreeNode::splitNode (node.cpp:64-149): rows with x[f] > thr go to the right child, the others
// to the left, each side keeping ascending sample order.  All nodes of a level are partitioned by one
// global pass (two launches): flag[k] = row at position k goes right; R = exclusive prefix sum of flag over the
// whole order array (in-chunk prefixes + chunk prefixes); for a row of node h at position k
//     left :  new = seg_start(h) + (k - seg_start(h)) - (R[k] - R[seg_start(h)])
//     right:  new = seg_start(h) + n_left(h)          + (R[k] - R[seg_start(h)])
// Rows of nodes that are not split at this level keep their position.  The comparison is done on the raw
// fp32 feature value, exactly like the reference.
#include "engine.cuh"

namespace gb {

constexpr int PART_CHUNK = 2048;   // rows per CTA (256 threads x 8 consecutive positions)

// Pass 1 (one CTA per chunk): side flag of every position + exclusive prefix of the flags INSIDE the chunk, packed as
// (flag << 15) | prefix in one u16 per position; the chunk's total; the last CTA to finish turns the totals into exclusive
// chunk prefixes (so that no separate scan launch is needed).
__global__ void __launch_bounds__(256)
part_flag_scan_kernel(const float *__restrict__ X, int F, const uint16_t *__restrict__ codesT, long long stride, int row_offset,
                      const int *__restrict__ order, const int *__restrict__ pnode, NodeArrays na, uint16_t *__restrict__ rloc,
                      int *__restrict__ chunk_sums, int *__restrict__ done_counter, int N, int n_chunks) {
    __shared__ int s_warp[8];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k0 = blockIdx.x * PART_CHUNK + threadIdx.x * 8;
    int f[8], loc = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = k0 + j;
        int fl = 0;
        if (k < N) {
            const int h = pnode[k];                  // node of the row at this position (coalesced; the row id is only needed for the code)
            // x > thr[f][j]  <=>  code(x) > j  (candidates.cu): 2 bytes of the feature-major codes per row
            if (na.state[h] == NODE_SPLIT) {
                const int i = order[k];
                if (codesT != nullptr) fl = (int)codesT[(size_t)na.split_f[h] * stride + row_offset + i] > na.split_j[h] ? 1 : 0;
                else fl = X[(size_t)i * F + na.split_f[h]] > na.split_thr[h] ? 1 : 0;          // node.cpp:89
            }
        }
        f[j] = fl; loc += fl;
    }
    int inc = loc;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { if (w < warp) wbase += s_warp[w]; total += s_warp[w]; }
    int run = wbase + inc - loc;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (k0 + j < N) rloc[k0 + j] = (uint16_t)((f[j] << 15) | run);
        run += f[j];
    }
    if (threadIdx.x == 0) {
        chunk_sums[blockIdx.x] = total;
        __threadfence();
        s_last = (atomicAdd(done_counter, 1) == n_chunks - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    // ---- last CTA: exclusive scan of the chunk totals, in place (every thread owns a contiguous run of chunks)
    __threadfence();
    const int per = (n_chunks + 255) / 256;
    const int c0 = min(n_chunks, (int)threadIdx.x * per), c1 = min(n_chunks, c0 + per);
    int mine = 0;
    for (int c = c0; c < c1; ++c) mine += __ldcg(chunk_sums + c);
    int incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += s_warp[w];
    int acc = base + incl - mine;
    for (int c = c0; c < c1; ++c) { const int v = __ldcg(chunk_sums + c); chunk_sums[c] = acc; acc += v; }
    if (threadIdx.x == 0) *done_counter = 0;
}

// Pass 2: stable scatter.  R[k] = chunk prefix + in-chunk prefix; rows right of the split go behind the node's left rows.
__global__ void __launch_bounds__(256)
part_scatter_kernel(const int *__restrict__ order_in, int *__restrict__ order_out, const int *__restrict__ pnode_in, int *__restrict__ pnode_out,
                    const uint16_t *__restrict__ rloc, const int *__restrict__ chunk_prefix, NodeArrays na, int N) {
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= N) return;
    const int i = order_in[k];
    const int h = pnode_in[k];
    if (na.state[h] != NODE_SPLIT) { order_out[k] = i; pnode_out[k] = h; return; }
    const int s0 = na.seg_start[h];
    const unsigned int pk = rloc[k], p0 = rloc[s0];
    const int rbefore = (chunk_prefix[k / PART_CHUNK] + (int)(pk & 0x7fffu)) - (chunk_prefix[s0 / PART_CHUNK] + (int)(p0 & 0x7fffu));
    if (pk >> 15) {
        const int nl = na.seg_len[2 * h + 1];
        order_out[s0 + nl + rbefore] = i;
        pnode_out[s0 + nl + rbefore] = 2 * h + 2;
    } else {
        order_out[s0 + (k - s0) - rbefore] = i;
        pnode_out[s0 + (k - s0) - rbefore] = 2 * h + 1;
    }
}

// after this call ws.order_p[0] is the new current order
void launch_partition(Model &m, const float *X, int level, int cur, cudaStream_t s) {
    (void)level; (void)cur;
    Workspace &ws = m.ws;
    const int N = ws.N;
    if (N == 0) return;
    const int n_chunks = ceil_div(N, PART_CHUNK);
    // ws.rflag holds the packed u16 (flag, in-chunk prefix) per position; chunk_sums[n_chunks] is the done counter (zero between launches)
    GB_LAUNCH(part_flag_scan_kernel, n_chunks, 256, 0, s, X, ws.F, ws.use_codesT ? ws.codesT.as<uint16_t>() : nullptr, ws.codesT_stride,
              ws.row_offset, ws.order_p[0], ws.pnode_p[0], ws.na, ws.rflag.as<uint16_t>(), ws.chunk_sums.as<int>(),
              ws.chunk_sums.as<int>() + ws.chunk_cap, N, n_chunks);
    GB_LAUNCH(part_scatter_kernel, ceil_div(N, 256), 256, 0, s, ws.order_p[0], ws.order_p[1], ws.pnode_p[0], ws.pnode_p[1],
              ws.rflag.as<uint16_t>(), ws.chunk_sums.as<int>(), ws.na, N);
    std::swap(ws.order_p[0], ws.order_p[1]);
    std::swap(ws.pnode_p[0], ws.pnode_p[1]);
}

}  // namespace gb
