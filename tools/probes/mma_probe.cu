// Micro-benchmark: issue rate of tcgen05.mma (SS mode, cta_group::1, M=128) on sm_100a as a function of kind (tf32 /
// bf16), N and accumulator reuse.  Operands are static shared-memory tiles (no TMA, no epilogue): the number is the
// floor any implicit-GEMM main loop of this library can reach per MMA.  Build: see tools/probes/build.sh.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../latent-pose-reenactment_b200/csrc/ptx.cuh"

using namespace b200lp;

template <int N, int KIND, int ELECT>   // KIND 0 tf32, 1 bf16; ELECT 1: warp-uniform loop, elect.sync leader issues
__global__ void __launch_bounds__(128, 1) probe(long long* out, int iters, int n_acc, int a_tiles) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    // zero operands: [a_tiles x 16 KB A][32 KB B]
    for (int i = threadIdx.x; i < (a_tiles * 16384 + 32768) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&tmem_slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (ELECT ? threadIdx.x < 32 : threadIdx.x == 0) {
        const bool leader = ELECT ? elect_one() : true;
        constexpr uint32_t idesc = KIND == 0 ? make_idesc_tf32(128, N, 0, 0) : make_idesc_bf16(128, N, 0, 0);
        const uint64_t dhi = make_smem_desc(0, 16, 1024, 2);
        const uint32_t b_addr = base + a_tiles * 16384;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t a_addr = base + (i & (a_tiles - 1)) * 16384;
            const uint64_t da0 = dhi | ((a_addr >> 4) & 0x3FFF), db0 = dhi | ((b_addr >> 4) & 0x3FFF);
            const uint32_t d = tmem + (i & (n_acc - 1)) * N;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (leader) {
                    if (KIND == 0) umma_tf32_ss(d, da0 + 2 * k, db0 + 2 * k, idesc, 1u);
                    else umma_f16_ss(d, da0 + 2 * k, db0 + 2 * k, idesc, 1u);
                }
            }
        }
        if (leader) umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (leader) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int N, int KIND, int ELECT>
void run(const char* name, int grid, int n_acc, int a_tiles) {
    const int iters = 4096;
    long long* d;
    cudaMalloc(&d, sizeof(long long) * grid);
    const int smem = a_tiles * 16384 + 32768 + 1024;
    cudaFuncSetAttribute(probe<N, KIND, ELECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<N, KIND, ELECT><<<grid, 128, smem>>>(d, 64, n_acc, a_tiles);   // warm
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<N, KIND, ELECT><<<grid, 128, smem>>>(d, iters, n_acc, a_tiles);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[256];
    cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    const double per = double(mx) / (iters * 4.0);
    const double kelem = KIND == 0 ? 8 : 16;
    const double tf = 2.0 * 128 * N * kelem * iters * 4.0 * grid / (ms * 1e-3) / 1e12;
    printf("%-5s elect=%d N=%3d grid=%3d acc=%d a_tiles=%d : %.1f clk/MMA (nominal %d)  %.0f TFLOP/s chip-equivalent  [%s]\n", name, ELECT, N,
           grid, n_acc, a_tiles, per, N / 2, tf, cudaGetErrorString(err));
    cudaFree(d);
}

template <int E>
void all() {
    for (int grid : {1, 148}) {
        run<64, 0, E>("tf32", grid, 1, 1);  run<64, 0, E>("tf32", grid, 4, 4);
        run<128, 0, E>("tf32", grid, 1, 1); run<128, 0, E>("tf32", grid, 2, 4);
        run<256, 0, E>("tf32", grid, 1, 1); run<256, 0, E>("tf32", grid, 2, 4);
        run<64, 1, E>("bf16", grid, 1, 1);  run<64, 1, E>("bf16", grid, 4, 4);
        run<128, 1, E>("bf16", grid, 1, 1); run<128, 1, E>("bf16", grid, 2, 4);
        run<256, 1, E>("bf16", grid, 1, 1); run<256, 1, E>("bf16", grid, 2, 4);
    }
}

int main() {
    all<0>();
    all<1>();
    return 0;
}
