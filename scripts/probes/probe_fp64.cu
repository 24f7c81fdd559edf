// Microbenchmark: FP64 issue rates on sm_100a -- DFMA (CUDA cores), DMMA (mma.sync f64), and both mixed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_fp64 probe_fp64.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ILP>
__global__ void dfma_kernel(double *out, int iters, double a, double b)
{
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double *c, const double *a, const double *b)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double *c, const double *a, const double *b)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void dmma884_kernel(double *out, int iters, double a, double b)
{
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma1688_kernel(double *out, int iters, double a, double b)
{
    double c[ILP][4], av[4] = {a, a, b, b}, bv[2] = {b, a};
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) dmma1688(c[i], av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma16816_kernel(double *out, int iters, double a, double b)
{
    double c[ILP][4], av[8] = {a, a, b, b, a, b, a, b}, bv[4] = {b, a, a, b};
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) dmma16816(c[i], av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// both in one warp: per iteration NM dmma884 + NF dfma
template <int NM, int NF>
__global__ void mixed_kernel(double *out, int iters, double a, double b)
{
    double c[NM > 0 ? NM : 1][2], acc[NF > 0 ? NF : 1];
#pragma unroll
    for (int i = 0; i < NM; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
#pragma unroll
    for (int i = 0; i < NF; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < (NM > NF ? NM : NF); i++) {
            if (i < NM) dmma884(c[i][0], c[i][1], a, b);
            if (i < NF) acc[i] = fma(acc[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NM; i++) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NF; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shared-memory double atomics and LDS.128 broadcast rates
__global__ void smem_atomic_kernel(double *out, int iters)
{
    __shared__ double buf[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) buf[i] = 0;
    __syncthreads();
    for (int it = 0; it < iters; it++) atomicAdd(&buf[(threadIdx.x * 9 + it * 37) & 4095], 1.0);
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = buf[threadIdx.x];
}
__global__ void smem_rmw_kernel(double *out, int iters)
{
    __shared__ double buf[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) buf[i] = 0;
    __syncthreads();
    for (int it = 0; it < iters; it++) { double *p = &buf[(threadIdx.x + it * 256) & 4095]; *p = *p + 1.0; }
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = buf[threadIdx.x];
}

template <class F>
float timeit(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 32 * 1024));
    const int iters = 20000;
    for (int wps = 4; wps <= 32; wps *= 2) {          // warps per SM
        int block = 128, grid = sms * wps * 32 / block;
        float ms;
        ms = timeit([&] { dfma_kernel<8><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
        printf("warps/SM %2d  DFMA ilp8        : %8.3f ms  %7.2f TFLOP/s (2 flop/fma)\n", wps, ms, 2.0 * grid * block * 8.0 * iters / ms * 1e-9);
        ms = timeit([&] { dmma884_kernel<4><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
        printf("warps/SM %2d  DMMA m8n8k4 ilp4 : %8.3f ms  %7.2f TFLOP/s\n", wps, ms, 2.0 * 256 * (grid * block / 32) * 4.0 * iters / ms * 1e-9);
        ms = timeit([&] { dmma1688_kernel<4><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
        printf("warps/SM %2d  DMMA m16n8k8 ilp4: %8.3f ms  %7.2f TFLOP/s\n", wps, ms, 2.0 * 1024 * (grid * block / 32) * 4.0 * iters / ms * 1e-9);
        ms = timeit([&] { dmma16816_kernel<4><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
        printf("warps/SM %2d  DMMA m16n8k16 il4: %8.3f ms  %7.2f TFLOP/s\n", wps, ms, 2.0 * 2048 * (grid * block / 32) * 4.0 * iters / ms * 1e-9);
        ms = timeit([&] { mixed_kernel<4, 8><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
        printf("warps/SM %2d  mixed 4 DMMA+8 DFMA: %8.3f ms  %7.2f TFLOP/s total (DMMA part %.2f, DFMA part %.2f)\n", wps, ms,
               (2.0 * 256 * 4 + 2.0 * 32 * 8) * (grid * block / 32) * (double) iters / ms * 1e-9,
               2.0 * 256 * 4 * (grid * block / 32) * (double) iters / ms * 1e-9, 2.0 * 32 * 8 * (grid * block / 32) * (double) iters / ms * 1e-9);
        ms = timeit([&] { mixed_kernel<2, 16><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
        printf("warps/SM %2d  mixed 2 DMMA+16 DFMA: %8.3f ms  %7.2f TFLOP/s total\n", wps, ms,
               (2.0 * 256 * 2 + 2.0 * 32 * 16) * (grid * block / 32) * (double) iters / ms * 1e-9);
    }
    {
        int grid = sms * 4, block = 256, it2 = 20000;
        float ms = timeit([&] { smem_atomic_kernel<<<grid, block>>>(out, it2); });
        printf("shared atomicAdd(double): %8.3f ms  %7.2f G atomics/s  (%.2f per clk per SM)\n", ms, (double) grid * block * it2 / ms * 1e-6,
               (double) grid * block * it2 / (ms * 1e-3) / sms / (p.clockRate * 1e3));
        ms = timeit([&] { smem_rmw_kernel<<<grid, block>>>(out, it2); });
        printf("shared plain RMW(double): %8.3f ms  %7.2f G rmw/s  (%.2f per clk per SM)\n", ms, (double) grid * block * it2 / ms * 1e-6,
               (double) grid * block * it2 / (ms * 1e-3) / sms / (p.clockRate * 1e3));
    }
    return 0;
}
