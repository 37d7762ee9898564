// tcgen05.mma issue-rate probe (sm_100a): clocks per MMA for the operand layouts / tile shapes the whitening kernels
// could use.  Not part of the library; built and run by hand:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/mma_probe tools/mma_probe.cu && tools/mma_probe
// One elected thread issues REPS MMAs (M = 128) back to back on static shared-memory contents, cycling over G
// accumulators, then commits and waits; (t1 - t0) / REPS is reported for CTA 0 and as the max over all 148 CTAs.
// Layouts (both operands): 0 = K-major SWIZZLE_NONE, 1 = MN-major SWIZZLE_NONE, 2 = K-major SWIZZLE_128B,
// 3 = MN-major SWIZZLE_128B.  ATMEM = 1: A operand read from TMEM (K-major by definition), B from shared memory.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout_type << 61);
}

template <int KIND>   // 0: i8 (u8 x u8 -> s32), 1: tf32 -> f32
__device__ __forceinline__ uint32_t make_idesc(int N, bool mn_major) {
    uint32_t d = (KIND == 0) ? (2u << 4) : ((1u << 4) | (2u << 7) | (2u << 10));
    if (mn_major) d |= (1u << 15) | (1u << 16);
    return d | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int KIND, bool ATMEM>
__device__ __forceinline__ void mma(uint32_t tmem_d, uint64_t adesc, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    if constexpr (ATMEM) {
        if constexpr (KIND == 0)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(tmem_d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(tmem_d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    } else {
        if constexpr (KIND == 0)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    }
}

template <int KIND, bool ATMEM>
__global__ void __launch_bounds__(128, 1) probe(int N, int layout, int G, int reps, int ksteps, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x;
    for (int i = tid; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u * (uint32_t)(i & 3);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    uint32_t leader = 0;
    if (tid < 32)                                     // elect.sync: see cleora_b200/csrc/gram_tc.cu (elect_one)
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\t@p mov.u32 %0, 1;\n\t}" : "+r"(leader));
    if (leader) {
        const bool mn = layout == 1 || layout == 3;
        const uint32_t idesc = make_idesc<KIND>(N, mn);
        // A occupies [0, 64 KB), B [64 KB, 192 KB).  One k-step of one MMA consumes 32 bytes of K per row.
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 64 * 1024;
        uint32_t lbo, sbo_a, sbo_b, lt, kstep_a, kstep_b;
        if (layout == 0)      { lbo = 128; sbo_a = sbo_b = 256; lt = 0; kstep_a = 128 / 8 * 256; kstep_b = N / 8 * 256; }
        else if (layout == 1) { lbo = 128; sbo_a = sbo_b = 512; lt = 0; kstep_a = 128 / 16 * 512; kstep_b = N / 16 * 512; }
        else if (layout == 2) { lbo = 16; sbo_a = sbo_b = 1024; lt = 2; kstep_a = kstep_b = 32; }          // within the 128-byte atom
        else                  { lbo = 4096; sbo_a = sbo_b = 1024; lt = 2; kstep_a = 4096; kstep_b = 4096 * ((N + 127) / 128); }
        // everything an MMA needs is in registers before the clock starts: the first probe built the descriptors
        // inside the loop and measured ~180 clk per MMA for EVERY shape -- the issuing thread, not the tensor pipe
        uint64_t ad[4], bd[4];
        uint32_t at[4], dc[8];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            ad[ks] = make_desc(a0 + (ks % ksteps) * kstep_a, lbo, sbo_a, lt);
            bd[ks] = make_desc(b0 + (ks % ksteps) * kstep_b, lbo, sbo_b, lt);
            at[ks] = tmem + 512 - 64 + (uint32_t)(ks * 8);              // A in TMEM: 8 columns per k-step
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) dc[j] = tmem + (uint32_t)((j % G) * N);
#pragma unroll
        for (int j = 0; j < 8; ++j) mma<KIND, ATMEM>(dc[j], ad[j & 3], at[j & 3], bd[j & 3], idesc, j >= G ? 1u : 0u);
        long long t0 = clock64();
        for (int r = 8; r < reps; r += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) mma<KIND, ATMEM>(dc[j], ad[j & 3], at[j & 3], bd[j & 3], idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
            if (clock64() - t0 > 4000000000LL) break;                  // watchdog (~2 s): never hang the box
        }
        long long t1 = clock64();
        out[blockIdx.x] = done ? t1 - t0 : -1;   // covers reps - 8 MMAs
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int KIND, bool ATMEM>
static void run(const char *kind, int N, int layout, int G, int grid) {
    const int reps = 2048, ksteps = 4;
    long long *d_out;
    cudaMalloc(&d_out, sizeof(long long) * grid);
    auto k = probe<KIND, ATMEM>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int rep = 0; rep < 2; ++rep) k<<<grid, 128, 200 * 1024>>>(N, layout, G, reps, ksteps, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s,N=%d,layout=%d,atmem=%d,G=%d,grid=%d,ERROR %s\n", kind, N, layout, (int)ATMEM, G, grid, cudaGetErrorString(e)); exit(1); }
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), d_out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (auto v : h) mx = v > mx ? v : mx;
    printf("%s,N=%d,layout=%d,atmem=%d,G=%d,grid=%d,clk_per_mma_cta0=%.1f,clk_per_mma_max=%.1f,floor=%.0f\n", kind, N, layout,
           (int)ATMEM, G, grid, (double)h[0] / (reps - 8), (double)mx / (reps - 8), N / 2.0);
    cudaFree(d_out);
}

// One configuration per process (an illegal combination poisons the context): `mma_probe` prints the number of
// configurations, `mma_probe <i>` runs configuration i.
int main(int argc, char **argv) {
    struct Cfg { int kind, atmem, N, layout, G, grid; };
    std::vector<Cfg> cfgs;
    for (int grid : {1, 148})
        for (int N : {64, 128, 256}) {
            const int G = 512 / N > 7 ? 7 : (448 / N < 1 ? 1 : 448 / N);
            for (int layout = 0; layout < 4; ++layout)
                for (int kind = 0; kind < 2; ++kind) cfgs.push_back({kind, 0, N, layout, G, grid});
            for (int kind = 0; kind < 2; ++kind)
                for (int layout : {0, 2}) cfgs.push_back({kind, 1, N, layout, G, grid});
            cfgs.push_back({0, 0, N, 0, 1, grid});       // one accumulator: dependent chain
        }
    if (argc < 2) { printf("%zu\n", cfgs.size()); return 0; }
    const Cfg c = cfgs[(size_t)atoi(argv[1]) % cfgs.size()];
    if (c.kind == 0 && !c.atmem) run<0, false>("i8", c.N, c.layout, c.G, c.grid);
    if (c.kind == 1 && !c.atmem) run<1, false>("tf32", c.N, c.layout, c.G, c.grid);
    if (c.kind == 0 && c.atmem) run<0, true>("i8", c.N, c.layout, c.G, c.grid);
    if (c.kind == 1 && c.atmem) run<1, true>("tf32", c.N, c.layout, c.G, c.grid);
    return 0;
}
