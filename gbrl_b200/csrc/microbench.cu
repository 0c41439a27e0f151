// microbench.cu -- hardware budgets quoted in DESIGN.md (not part of the reference surface).
//   0: shared-memory ATOMS.ADD (int32) throughput, bank-conflict free by construction (lane == bank),
//      random row per lane  -> G lane-atomics / s on the whole chip
//   1: same but fully random addresses (birthday bank conflicts)
//   2: conflict-free, three independent planes per element (count / lo / hi pattern of hist_kernel)
//   3: global REDG.ADD.64 throughput, spread addresses
//   4: streaming read bandwidth (GB/s) of a 1 GiB buffer with 16-byte loads
#include "engine.cuh"

namespace gb {

template <int MODE>
__global__ void __launch_bounds__(256) atoms_kernel(int iters, unsigned int *sink) {
    extern __shared__ int sh[];
    const int PLANE = NB * FT;
    for (int i = threadIdx.x; i < 3 * PLANE; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned int x = 1234567u + 977u * (blockIdx.x * blockDim.x + threadIdx.x);
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
        x = x * 1664525u + 1013904223u;
        const int row = (x >> 12) & (NB - 1);
        if (MODE == 0) atomicAdd(&sh[row * FT + lane], 1);
        else if (MODE == 1) atomicAdd(&sh[(x >> 9) & (PLANE - 1)], 1);
        else {
            atomicAdd(&sh[row * FT + lane], 1);
            atomicAdd(&sh[PLANE + row * FT + lane], (int)(x & 0xffff));
            atomicAdd(&sh[2 * PLANE + row * FT + lane], (int)(x >> 20));
        }
    }
    __syncthreads();
    unsigned int acc = 0;
    for (int i = threadIdx.x; i < 3 * PLANE; i += blockDim.x) acc += sh[i];
    if (acc == 0xdeadbeefu) sink[0] = acc;
}

__global__ void __launch_bounds__(256) redg_kernel(int iters, long long *buf, int n) {
    unsigned int x = 7654321u + 31u * (blockIdx.x * blockDim.x + threadIdx.x);
    for (int it = 0; it < iters; ++it) {
        x = x * 1664525u + 1013904223u;
        red_add64(buf + ((x >> 8) % n), 1);
    }
}

__global__ void __launch_bounds__(256) stream_read_kernel(const uint4 *buf, size_t n, unsigned int *sink) {
    unsigned int acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = buf[i];
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0xdeadbeefu) sink[0] = acc;
}

double microbench(int which, int iters) {
    cudaEvent_t e0, e1;
    GB_CUDA(cudaEventCreate(&e0)); GB_CUDA(cudaEventCreate(&e1));
    unsigned int *sink; GB_CUDA(cudaMalloc(&sink, 4));
    cudaDeviceProp prop; int dev; GB_CUDA(cudaGetDevice(&dev)); GB_CUDA(cudaGetDeviceProperties(&prop, dev));
    const int sms = prop.multiProcessorCount;
    double result = 0.0;
    float ms = 0.0f;
    const size_t smem = 3 * NB * FT * sizeof(int);
    if (which >= 0 && which <= 2) {
        const int grid = sms * 2;
        auto run = [&](int n) {
            if (which == 0) { cudaFuncSetAttribute(atoms_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); atoms_kernel<0><<<grid, 256, smem>>>(n, sink); }
            if (which == 1) { cudaFuncSetAttribute(atoms_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); atoms_kernel<1><<<grid, 256, smem>>>(n, sink); }
            if (which == 2) { cudaFuncSetAttribute(atoms_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); atoms_kernel<2><<<grid, 256, smem>>>(n, sink); }
        };
        run(iters / 10 + 1);
        GB_CUDA(cudaDeviceSynchronize());
        GB_CUDA(cudaEventRecord(e0));
        run(iters);
        GB_CUDA(cudaEventRecord(e1));
        GB_CUDA(cudaEventSynchronize(e1));
        GB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double n_at = (double)grid * 256.0 * iters * (which == 2 ? 3.0 : 1.0);
        result = n_at / (ms * 1e-3) / 1e9;
    } else if (which == 3) {
        const int n = 1 << 22;
        long long *buf; GB_CUDA(cudaMalloc(&buf, (size_t)n * 8)); GB_CUDA(cudaMemset(buf, 0, (size_t)n * 8));
        redg_kernel<<<sms * 4, 256>>>(iters / 10 + 1, buf, n);
        GB_CUDA(cudaDeviceSynchronize());
        GB_CUDA(cudaEventRecord(e0));
        redg_kernel<<<sms * 4, 256>>>(iters, buf, n);
        GB_CUDA(cudaEventRecord(e1)); GB_CUDA(cudaEventSynchronize(e1));
        GB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        result = (double)sms * 4 * 256.0 * iters / (ms * 1e-3) / 1e9;
        cudaFree(buf);
    } else if (which == 4) {
        const size_t bytes = (size_t)1 << 30;
        uint4 *buf; GB_CUDA(cudaMalloc(&buf, bytes)); GB_CUDA(cudaMemset(buf, 1, bytes));
        stream_read_kernel<<<sms * 8, 256>>>(buf, bytes / 16, sink);
        GB_CUDA(cudaDeviceSynchronize());
        GB_CUDA(cudaEventRecord(e0));
        for (int r = 0; r < (iters > 0 ? iters : 1); ++r) stream_read_kernel<<<sms * 8, 256>>>(buf, bytes / 16, sink);
        GB_CUDA(cudaEventRecord(e1)); GB_CUDA(cudaEventSynchronize(e1));
        GB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        result = (double)bytes * (iters > 0 ? iters : 1) / (ms * 1e-3) / 1e9;
        cudaFree(buf);
    } else {
        throw Error("unknown microbench id");
    }
    GB_CUDA(cudaGetLastError());
    cudaFree(sink); cudaEventDestroy(e0); cudaEventDestroy(e1);
    return result;
}

}  // namespace gb
