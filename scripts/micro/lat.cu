// Single-warp latency microbenchmarks on B200 (cycles per dependent operation):
// DFMA, DADD, sqrt, division, LDS round trip, STS->syncwarp->LDS, shuffle, L1/L2 hit loads,
// effect of prefetch.global.L1.   nvcc -arch=sm_100a -O3 -o lat lat.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_arith(double *out, double x, int n, long long *cyc) {
    double a = x, b = x * 0.5, c = 1.0000001;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) a = fma(a, c, b);
    long long t1 = clock64();
    double d = a;
    for (int i = 0; i < n; ++i) d = d + b;
    long long t2 = clock64();
    double e = fabs(d) + 2.0;
    for (int i = 0; i < n; ++i) e = sqrt(e) + 3.0;
    long long t3 = clock64();
    double f = e;
    for (int i = 0; i < n; ++i) f = 7.0 / f + 1.0;
    long long t4 = clock64();
    double g = f;
    for (int i = 0; i < n; ++i) g = __shfl_sync(0xffffffffu, g, (threadIdx.x + 1) & 31) + 1.0;
    long long t5 = clock64();
    if (threadIdx.x == 0) {
        cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4;
    }
    out[threadIdx.x] = g;
}

__global__ void k_smem(double *out, int n, long long *cyc) {
    __shared__ double s[64];
    int lane = threadIdx.x;
    s[lane] = lane;
    s[lane + 32] = 1.0;
    __syncwarp();
    // dependent LDS chain (index from loaded value)
    int idx = lane;
    long long t0 = clock64();
    double v = 0;
    for (int i = 0; i < n; ++i) { v = s[idx & 63]; idx = (int)v + 1; }
    long long t1 = clock64();
    // STS -> syncwarp -> LDS of a neighbour's value, dependent
    double x = v;
    for (int i = 0; i < n; ++i) {
        s[lane] = x;
        __syncwarp();
        x = s[(lane + 1) & 31] + 1.0;
        __syncwarp();
    }
    long long t2 = clock64();
    if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
    out[lane] = x;
}

// walk a buffer: one new 128-byte line per iteration, loads dependent through the address
__global__ void k_gmem(const long long *buf, long long *out, int n, int stride_elems, int prefetch, long long *cyc) {
    int lane = threadIdx.x;
    long long off = 0, acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        if (prefetch && lane == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(buf + off + (long long)prefetch * stride_elems));
        long long v = buf[off + lane % 6];      // value is 0: keeps the chain dependent
        acc += v;
        off += stride_elems + v;
    }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
    out[lane] = acc;
}

// same but with a store to the line just read (write-through interaction)
__global__ void k_gmem_rw(long long *buf, long long *out, int n, int stride_elems, long long *cyc) {
    int lane = threadIdx.x;
    long long off = 0, acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        long long v = buf[off + lane % 6];
        acc += v;
        if (lane < 6) buf[off + lane] = v;       // store back (same value)
        off += stride_elems + v;
    }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
    out[lane] = acc;
}

int main() {
    double *out; long long *cyc, *buf, *lout;
    cudaMalloc(&out, 1024); cudaMalloc(&cyc, 1024); cudaMalloc(&lout, 1024);
    size_t nb = 64 << 20;
    cudaMalloc(&buf, nb); cudaMemset(buf, 0, nb);
    long long h[8];
    int n = 4096;
    for (int rep = 0; rep < 2; ++rep) {
        k_arith<<<1, 32>>>(out, 1.0, n, cyc);
        cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        if (rep) printf("dependent DFMA %.1f  DADD %.1f  sqrt+add %.1f  div+add %.1f  shfl+add %.1f cycles\n", h[0] / (double)n, h[1] / (double)n,
                        h[2] / (double)n, h[3] / (double)n, h[4] / (double)n);
        k_smem<<<1, 32>>>(out, n, cyc);
        cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        if (rep) printf("dependent LDS %.1f   STS->syncwarp->LDS(+add)->syncwarp %.1f cycles\n", h[0] / (double)n, h[1] / (double)n);
    }
    int strides[3] = {6, 16, 64};      // elements of 8 bytes: 48 B (band block), 128 B (a line), 512 B
    for (int s = 0; s < 3; ++s)
        for (int pf = 0; pf <= 32; pf += 16)
            for (int rep = 0; rep < 2; ++rep) {
                k_gmem<<<1, 32>>>(buf, lout, n, strides[s], pf, cyc);
                cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
                printf("global walk stride %4d B, prefetch %2d ahead, pass %d (0 = cold L2 line fill, 1 = L2 hits): %.1f cycles/step\n",
                       strides[s] * 8, pf, rep, h[0] / (double)n);
            }
    for (int rep = 0; rep < 2; ++rep) {
        k_gmem_rw<<<1, 32>>>(buf, lout, n, 6, cyc);
        cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("global walk stride 48 B with store-back, pass %d: %.1f cycles/step\n", rep, h[0] / (double)n);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
