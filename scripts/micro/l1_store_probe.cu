// One thread: does a global store evict the line from the SM's L1?  (Decides whether the topology walker's read-after-write of its
// position lists can hit L1.)   nvcc -arch=sm_100a -O2 -o /tmp/l1probe scripts/micro/l1_store_probe.cu && /tmp/l1probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(int* buf, long long* out) {
    if (threadIdx.x) return;
    volatile int sink = 0;
    int* p = buf + 4096;
    long long t0, t1;
    // cold load (L2 or DRAM), then an L1 hit
    t0 = clock64(); int a = *(volatile int*)p; sink += a; t1 = clock64(); out[0] = t1 - t0;
    t0 = clock64(); a = *(volatile int*)p; sink += a; t1 = clock64(); out[1] = t1 - t0;
    // plain (cacheable) loads: asm keeps them ld.global.ca
    int v;
    asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(v) : "l"(p + 32)); sink += v;                       // warm another line into L1
    t0 = clock64(); asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(v) : "l"(p + 32)); sink += v; t1 = clock64(); out[2] = t1 - t0;   // L1 hit
    asm volatile("st.global.s32 [%0], %1;" :: "l"(p + 33), "r"(v + 1) : "memory");                          // store into the same line
    t0 = clock64(); asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(v) : "l"(p + 32)); sink += v; t1 = clock64(); out[3] = t1 - t0;   // after the store
    t0 = clock64(); asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(v) : "l"(p + 32)); sink += v; t1 = clock64(); out[4] = t1 - t0;   // and once more
    // L2 hit for comparison: ld.cg of a line that was touched before
    t0 = clock64(); asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p + 32)); sink += v; t1 = clock64(); out[5] = t1 - t0;
    out[6] = sink;
}
int main() {
    int* buf; long long* out; long long h[7];
    cudaMalloc(&buf, 1 << 20); cudaMemset(buf, 0, 1 << 20); cudaMalloc(&out, sizeof(h));
    for (int rep = 0; rep < 3; ++rep) {
        probe<<<1, 32>>>(buf, out); cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("cold %lld | volatile again %lld | L1 hit (ld.ca) %lld | ld.ca after a store to the line %lld | again %lld | L2 hit (ld.cg) %lld clk\n", h[0], h[1], h[2], h[3], h[4], h[5]);
    }
    return 0;
}
