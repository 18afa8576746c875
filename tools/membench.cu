// tools/membench.cu -- what a B200 sustains for the wavefront's state access pattern: gather a 64-byte record
// per thread through an index queue, touch it, write it back (optionally also write a second 64-byte record).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/membench.cu -o gpurun_out/membench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#define P_RESULT_BASE ((size_t)(1 << 24) * 4)

__global__ void k_gather(const int* __restrict__ queue, int n, float4* __restrict__ rec, float4* __restrict__ rec2, int write2, int work)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
        const int s = queue[i];
        float4* p = rec + (size_t)s * 4;
        float4 a = p[0], b = p[1], c = p[2], d = p[3];
        float acc = a.x + b.y + c.z + d.w;
        for (int k = 0; k < work; ++k) acc = acc * 1.0001f + 0.5f;   // dependent FP32 chain: `work` x 2 ops
        a.x = acc; b.y = acc; c.z = acc; d.w = acc;
        p[0] = a; p[1] = b; p[2] = c; p[3] = d;
        if (write2 == 1)      // 64 bytes into a 128-byte-stride line
        {
            float4* q = rec2 + (size_t)s * 8;
            q[0] = a; q[1] = b; q[2] = c; q[3] = d;
        }
        else if (write2 == 2) // 64 bytes into a dense 64-byte-stride record
        {
            float4* q = rec2 + (size_t)s * 4;
            q[0] = a; q[1] = b; q[2] = c; q[3] = d;
        }
        else if (write2 == 3) // 64 bytes dense + read of a 32-byte dense record (the pending result)
        {
            float4* q = rec2 + (size_t)s * 4;
            const float4* r = rec2 + (size_t)P_RESULT_BASE + (size_t)s * 2;
            float4 x = r[0], y = r[1];
            a.y += x.x + y.y;
            q[0] = a; q[1] = b; q[2] = c; q[3] = d;
        }
    }
}

int main()
{
    const int P = 1 << 24;
    float4 *rec, *rec2; int* queue;
    cudaMalloc(&rec, (size_t)P * 64); cudaMalloc(&rec2, (size_t)P * 128); cudaMalloc(&queue, (size_t)P * 4);
    cudaMemset(rec, 0, (size_t)P * 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<int> h;
    for (int mode = 0; mode < 3; ++mode)
    {
        h.clear();
        unsigned long long st = 88172645463325252ull;
        auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
        if (mode == 0) for (int i = 0; i < P; ++i) h.push_back(i);                              // identity
        if (mode == 1) for (int i = 0; i < P; ++i) if (rnd() % 100 < 60) h.push_back(i);        // sorted, 60 % dense (a lobe queue)
        if (mode == 2) { for (int i = 0; i < P; ++i) if (rnd() % 100 < 60) h.push_back(i);      // same, in chunks of 32 shuffled
                         int nc = (int)h.size() / 32; for (int c = nc - 1; c > 0; --c) { int o = rnd() % (c + 1); for (int k = 0; k < 32; ++k) std::swap(h[c * 32 + k], h[o * 32 + k]); } }
        const int n = (int)h.size();
        cudaMemcpy(queue, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
        for (int write2 = 0; write2 < 4; ++write2)
            for (int work : { 200 })
                for (int threads : { 128 })
                {
                    const int grid = 148 * (threads == 128 ? 24 : 16);
                    k_gather<<<grid, threads>>>(queue, n, rec, rec2, write2, work);
                    cudaEventRecord(e0);
                    for (int r = 0; r < 3; ++r) k_gather<<<grid, threads>>>(queue, n, rec, rec2, write2, work);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
                    const double bytes = (double)n * (64 + 64 + 4 + (write2 ? 64 : 0) + (write2 == 3 ? 32 : 0));
                    printf("queue %-22s n=%8d write2=%d work=%4d threads=%3d  %7.3f ms  %7.1f GB/s  %6.2f G records/s\n",
                           mode == 0 ? "identity" : mode == 1 ? "sorted 60% dense" : "60%, 32-chunks shuffled", n, write2, work, threads, ms, bytes / ms / 1e6, n / ms / 1e6);
                }
    }
    return 0;
}
