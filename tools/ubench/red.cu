// L2 reduction (RED) throughput microbenchmark: how many atomic-add operations per second does the B200 L2 sustain,
// by operand width, footprint (L2-resident vs streaming) and per-instruction contiguity?  Decides how S is accumulated in k_sgm.cu.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k_red(unsigned long long* buf, size_t n64, int passes) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; p++) {
        if (MODE == 0) {  // RED.64, 8 B per lane, contiguous 256 B per warp instruction
            for (size_t i = tid; i < n64; i += stride) asm volatile("red.global.add.u64 [%0], %1;" ::"l"(buf + i), "l"(0x0001000100010001ull) : "memory");
        } else if (MODE == 1) {  // RED.32, 4 B per lane
            unsigned int* b32 = (unsigned int*)buf;
            for (size_t i = tid; i < 2 * n64; i += stride) asm volatile("red.global.add.u32 [%0], %1;" ::"l"(b32 + i), "r"(0x00010001u) : "memory");
        } else if (MODE == 2) {  // two RED.64 per lane on a 16-byte slot (each instruction touches half of every sector)
            for (size_t i = tid; i < n64 / 2; i += stride) {
                asm volatile("red.global.add.u64 [%0], %1;" ::"l"(buf + 2 * i), "l"(0x0001000100010001ull) : "memory");
                asm volatile("red.global.add.u64 [%0], %1;" ::"l"(buf + 2 * i + 1), "l"(0x0001000100010001ull) : "memory");
            }
        } else if (MODE == 3) {  // plain 8-byte stores (reference)
            for (size_t i = tid; i < n64; i += stride) buf[i] = i;
        } else if (MODE == 4) {  // v4.f32 vector reduction, 16 B per lane
            float* f = (float*)buf;
            for (size_t i = tid; i < n64 / 2; i += stride) asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(f + 4 * i), "f"(1.0f) : "memory");
        } else if (MODE == 5) {  // v2.f32, 8 B per lane
            float* f = (float*)buf;
            for (size_t i = tid; i < n64; i += stride) asm volatile("red.global.add.v2.f32 [%0], {%1,%1};" ::"l"(f + 2 * i), "f"(1.0f) : "memory");
        } else if (MODE == 6) {  // v8 f16x2?  (v4.f16x2 = 16 B per lane)
            for (size_t i = tid; i < n64 / 2; i += stride) asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1,%1,%1,%1};" ::"l"(buf + 2 * i), "r"(0x3c003c00u) : "memory");
        }
    }
}
template <int MODE>
void run(const char* name, unsigned long long* buf, size_t bytes, int passes) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const size_t n64 = bytes / 8;
    k_red<MODE><<<148 * 8, 512>>>(buf, n64, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k_red<MODE><<<148 * 8, 512>>>(buf, n64, passes);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-28s footprint %7.1f MB  %6.3f ms/pass  %7.1f GB/s payload  (%s)\n", name, bytes / 1e6, ms / passes, bytes * (double)passes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    unsigned long long* buf; const size_t big = 315ull << 20;
    cudaMalloc(&buf, big); cudaMemset(buf, 0, big);
    for (size_t bytes : {(size_t)32 << 20, big}) {
        run<3>("STG.64", buf, bytes, 8);
        run<0>("RED.64 contiguous", buf, bytes, 8);
        run<1>("RED.32 contiguous", buf, bytes, 8);
        run<2>("RED.64 x2 interleaved", buf, bytes, 8);
        run<5>("RED.v2.f32", buf, bytes, 8);
        run<4>("RED.v4.f32", buf, bytes, 8);
        run<6>("RED.v4.f16x2", buf, bytes, 8);
    }
    return 0;
}
