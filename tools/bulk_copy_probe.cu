// Probe: throughput of per-row `cp.async.bulk` (UBLKCP) gathers on B200 as a function of row length and alignment.
// Every warp of a persistent grid issues one bulk copy per lane from a pseudo-random row of a large buffer into its private
// shared-memory ring (2 stages), waits on the stage's mbarrier and discards the data. Reports copies/s and GB/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bulk_copy_probe tools/bulk_copy_probe.cu && ./tools/bulk_copy_probe
// Used to decide between bulk copies and 128-bit register loads in csrc/gather.cu (profiles/r2_bulk_copy_probe.txt).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int ROWS_PER_STAGE>
__global__ void __launch_bounds__(256) k_probe(const unsigned char* __restrict__ base, size_t n_rows, uint32_t pitch, uint32_t bytes, uint32_t skew,
                                               int groups, unsigned long long* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t slot = (bytes + 127) / 128 * 128;
    unsigned char* mine = smem + (size_t)warp * (2 * ROWS_PER_STAGE * slot + 128);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(mine);
    unsigned char* slots = mine + 128;
    if (lane == 0) {
        for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bars + i)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint64_t rng = (blockIdx.x * 8ull + warp) * 0x9E3779B97F4A7C15ull + 12345;
    auto issue = [&](int g) {
        const int st = g & 1;
        const uint32_t bar = smem_u32(bars + st);
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * ROWS_PER_STAGE) : "memory");
        uint64_t r = rng + (uint64_t)g * 1000003ull + lane * 7919ull;
        r ^= r >> 33; r *= 0xff51afd7ed558ccdull; r ^= r >> 33;
        const size_t row = r % n_rows;
        if (lane < ROWS_PER_STAGE)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(slots + (size_t)(st * ROWS_PER_STAGE + lane) * slot)),
                         "l"(base + row * pitch + skew), "r"(bytes), "r"(bar)
                         : "memory");
    };
    issue(0);
    unsigned acc = 0;
    for (int g = 0; g < groups; ++g) {
        if (g + 1 < groups) issue(g + 1);
        const int st = g & 1;
        const uint32_t parity = (g >> 1) & 1;
        asm volatile(
            "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(bars + st)), "r"(parity)
            : "memory");
        acc += slots[(size_t)(st * ROWS_PER_STAGE) * slot + lane];
        __syncwarp();
    }
    if (acc == 0xffffffffu) *sink = acc;
}

template <int RPS>
static void run(const unsigned char* d, size_t n_rows, uint32_t pitch, uint32_t bytes, uint32_t skew, int sms, unsigned long long* sink) {
    const uint32_t slot = (bytes + 127) / 128 * 128;
    const size_t smem = 8 * (2 * RPS * slot + 128);
    if (smem > 200 * 1024) return;
    cudaFuncSetAttribute(k_probe<RPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_probe<RPS>, 256, smem);
    if (per_sm < 1) return;
    const int grid = sms * per_sm, groups = 400;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_probe<RPS><<<grid, 256, smem>>>(d, n_rows, pitch, bytes, skew, groups, sink);
    cudaEventRecord(e0);
    k_probe<RPS><<<grid, 256, smem>>>(d, n_rows, pitch, bytes, skew, groups, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double copies = (double)grid * 8 * groups * RPS;
    printf("bytes %5u pitch %5u skew %3u rows/stage %2d CTAs/SM %d: %7.2f G copies/s  %8.1f GB/s  (%.1f clk per copy per SM at 1.9 GHz)  %s\n", bytes, pitch, skew, RPS,
           per_sm, copies / ms / 1e6, copies * bytes / ms / 1e6, 1.9e9 * sms / (copies / (ms * 1e-3)), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t total = 4ull << 30;   // larger than L2
    unsigned char* d = nullptr;
    unsigned long long* sink = nullptr;
    cudaMalloc(&d, total + 4096);
    cudaMalloc(&sink, 8);
    cudaMemset(d, 1, total + 4096);
    struct Case { uint32_t bytes, pitch, skew; };
    const std::vector<Case> cases = {{128, 128, 0}, {256, 256, 0}, {400, 400, 0}, {400, 512, 0}, {400, 512, 16}, {512, 512, 0}, {512, 512, 16}, {512, 528, 0},
                                     {1024, 1024, 0}, {1024, 1040, 0}, {2048, 2048, 0}, {160, 160, 0}, {800, 800, 0}};
    for (const Case& c : cases) {
        const size_t n_rows = total / c.pitch;
        run<8>(d, n_rows, c.pitch, c.bytes, c.skew, sms, sink);
        run<32>(d, n_rows, c.pitch, c.bytes, c.skew, sms, sink);
    }
    // L2-resident working set (64 MB): the rate of the copy engine itself
    printf("-- L2-resident (64 MB) --\n");
    for (const Case& c : cases) {
        const size_t n_rows = (64ull << 20) / c.pitch;
        run<8>(d, n_rows, c.pitch, c.bytes, c.skew, sms, sink);
    }
    return 0;
}
