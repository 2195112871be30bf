// pipes.cu — instruction-throughput microbenchmark for the integer ops the K1 kernels are built from (B200, sm_100a).
// Prints warp-instructions per clock per SM for each op, and for a few mixes, with 16 warps/SM x 8 independent chains.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 2048
#define CH 8

template <int OP>
__device__ __forceinline__ void step(uint32_t (&v)[CH], uint32_t a, uint32_t b, uint32_t sbase, int lane, uint32_t& sink) {
#pragma unroll
    for (int i = 0; i < CH; i++) {
        if (OP == 0) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(v[i]) : "r"(a), "r"(b + i));
        if (OP == 1) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(v[i]) : "r"(a), "r"(b));
        if (OP == 2) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(v[i]) : "r"(a), "r"(b));
        if (OP == 3) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(a), "r"(b));
        if (OP == 4) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(a), "r"(b));
        if (OP == 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(a), "r"(b));
        if (OP == 6) asm volatile("add.u32 %0, %0, %1;" : "+r"(v[i]) : "r"(a));
        if (OP == 7) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(a), "r"(b));
        if (OP == 8) asm volatile("shfl.sync.up.b32 %0, %0, 5, 0x0, 0xffffffff;" : "+r"(v[i]));
        if (OP == 9) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[i]) : "r"(sbase + v[i]));
        if (OP == 10) { uint32_t y; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v[i]), "=r"(y) : "r"(sbase + v[i])); sink ^= y; }
        if (OP == 11) { uint32_t y, z, w; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[i]), "=r"(y), "=r"(z), "=r"(w) : "r"(sbase + v[i])); sink ^= y ^ z ^ w; }
        if (OP == 12) asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(a), "r"(0));
        if (OP == 13) asm volatile("min.u16x2 %0, %0, %1;" : "+r"(v[i]) : "r"(a));
        if (OP == 14) { asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(v[i]) : "r"(a), "r"(b + i)); i++; asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(a), "r"(b)); }   // mix SAD + PRMT 1:1
        if (OP == 15) { asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(v[i]) : "r"(a), "r"(b)); i++; asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(a), "r"(b)); }            // mix IDP + PRMT 1:1
        if (OP == 16) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(a), "r"(b)); i++; asm volatile("add.u32 %0, %0, %1;" : "+r"(v[i]) : "r"(a)); }                              // mix IMAD + IADD 1:1
        if (OP == 17) { asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(v[i]) : "r"(a), "r"(b + i)); i++; asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(v[i]) : "r"(a), "r"(b)); }  // mix SAD + IDP
        if (OP == 18) { asm volatile("shfl.sync.up.b32 %0, %0, 5, 0x0, 0xffffffff;" : "+r"(v[i])); i++; asm volatile("add.u32 %0, %0, %1;" : "+r"(v[i]) : "r"(a)); }                              // mix SHFL + IADD
        if (OP == 19) { asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[i]) : "r"(sbase + v[i])); i++; asm volatile("add.u32 %0, %0, %1;" : "+r"(v[i]) : "r"(a)); }  // LDS + IADD
    }
}

template <int OP>
__global__ void __launch_bounds__(512) k(uint32_t* out, uint32_t a, uint32_t b, long long* cyc) {
    __shared__ uint32_t sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 4;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
    uint32_t v[CH];
    uint32_t sink = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) v[i] = (OP == 9 || OP == 19) ? (lane * 4 + i * 128) : OP == 10 ? (lane * 8 + i * 256) : OP == 11 ? (lane * 16 + i * 512) : threadIdx.x * 17 + i;
        long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) step<OP>(v, a, b, sbase, lane, sink);
    long long t1 = clock64();
    uint32_t s = sink;
#pragma unroll
    for (int i = 0; i < CH; i++) s ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, uint32_t* out, long long* cyc) {
    const int grid = 148, threads = 512;
    k<OP><<<grid, threads>>>(out, 0x01020304u, 0x3210u, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<grid, threads>>>(out, 0x01020304u, 0x3210u, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; i++) avg += h[i]; avg /= grid;
    double winstr = (double)ITERS * CH * (threads / 32);
    printf("%-28s %8.3f warp-instr/clk/SM  (%.0f cycles, %.3f ms)  err=%s\n", name, winstr / avg, avg, ms, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    run<0>("VABSDIFF4.U8.ACC", out, cyc);
    run<12>("VABSDIFF4.U8", out, cyc);
    run<1>("IDP.2A", out, cyc);
    run<2>("IDP.4A", out, cyc);
    run<3>("PRMT", out, cyc);
    run<4>("SHF", out, cyc);
    run<5>("LOP3", out, cyc);
    run<6>("IADD", out, cyc);
    run<7>("IMAD", out, cyc);
    run<13>("VIMNMX.U16x2", out, cyc);
    run<8>("SHFL.UP", out, cyc);
    run<9>("LDS.32", out, cyc);
    run<10>("LDS.64", out, cyc);
    run<11>("LDS.128", out, cyc);
    run<14>("mix SAD.ACC+PRMT", out, cyc);
    run<15>("mix IDP.2A+PRMT", out, cyc);
    run<16>("mix IMAD+IADD", out, cyc);
    run<17>("mix SAD.ACC+IDP.2A", out, cyc);
    run<18>("mix SHFL+IADD", out, cyc);
    run<19>("mix LDS.32+IADD", out, cyc);
    return 0;
}
