// Weight gradient of the 3x3 / 1x1 convolutions on the sm_100a tensor cores.
//
//   dW[(tap,ci)][co] = sum_pixels  X[pixel + tap][ci] * dY[pixel][co]
//   GEMM view:  M = (tap, ci) rows (128 per CTA = four 32-channel blocks, each with its own tap shift),
//               N = Cout tile, K = pixels (split across CTAs: "split-K", deterministic two-pass reduction).
//   Both operands are read straight from the NHWC tensors: a TMA box [32 pixels][32 channels] lands in shared
//   memory as 32 rows (K) x 128 bytes (32 channels of M or N): the canonical MN-major SWIZZLE_128B_BASE32B UMMA operand
//   (TMA swizzle mode 128B_ATOM_32B; the only MN-major layout the tensor core accepts for 32-bit operands),
//   so no transposed copy of the activations or gradients ever exists in HBM.
//   Zero padding = TMA out-of-bounds zero fill on the shifted X box.
//
// Replaces torch autograd's conv backward-filter for the convs listed in include/b200lp.h.
#include "common.cuh"
#include "ptx.cuh"

namespace b200lp {

constexpr int kWgM = 128;          // (tap,ci) rows per CTA
constexpr int kWgThreads = 192;
constexpr int kWgMaxStages = 8;
constexpr int kWgMaxSmem = 226 * 1024;   // 227 KB opt-in limit minus the static barriers

struct WgradParams {
    float* ws;            // [splits][rows_total][Cout]
    int N, H, W, Cin, Cout;
    int ksize;
    int pw, ph, pn;       // pixel box: pw*ph*pn == kstep
    int kstep;            // pixels per pipeline stage (32 or 64)
    int blk_bytes;        // one [kstep pixels][32 ch] box = kstep * 128 bytes
    int stages;           // smem ring depth
    int steps_w, steps_h; // boxes per image row / column
    int total_steps;      // K steps over the whole batch
    int steps_per_split;
    int cblks;            // Cin / 32
    int rows_total;       // taps * Cin
    int blocks_total;     // taps * cblks
    int grouped;          // 1: grouped conv — the M blocks of N tile n are the taps of channel block n (cblks == 1)
};

template <int BLOCK_N>
struct WgCfg {
    static constexpr int kBlocks = 4 + BLOCK_N / 32;               // [kstep][32 ch] boxes per stage (A: 4, B: N/32)
    static constexpr uint32_t kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kWgThreads, 2)
conv_wgrad_tf32_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                       const WgradParams p) {
    using Cfg = WgCfg<BLOCK_N>;
    const int kStages = p.stages;
    const int kWgBlkBytes = p.blk_bytes;
    const int kABytes = 4 * kWgBlkBytes;
    const int kBBytes = (BLOCK_N / 32) * kWgBlkBytes;
    const int kStageBytes = kABytes + kBBytes;

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kWgMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kWgMaxStages];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_slot;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_tile = blockIdx.x;
    const int m_tile = blockIdx.y;
    const int split = blockIdx.z;

    const int ks_begin = split * p.steps_per_split;
    int ks_end = ks_begin + p.steps_per_split;
    if (ks_end > p.total_steps) ks_end = p.total_steps;
    const int nsteps = ks_end - ks_begin;   // >= 1 by construction of the grid

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmDY);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<Cfg::kTmemCols>(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // warp-uniform loop, one elected lane issues (see conv_igemm.cu: a divergent `if (lane == 0)` costs an
        // ELECT/branch loop per TMA and MMA instruction)
        const bool leader = elect_one();
        {
            const int pad = p.ksize >> 1;
            // the four 32-row blocks of this M tile: (tap shift, channel block); blocks past the end are skipped
            int blk_dx[4], blk_dy[4], blk_c[4];
            int nvalid = 0;
            for (int j = 0; j < 4; ++j) {
                const int b = m_tile * 4 + j;
                if (b < p.blocks_total) {
                    const int tap = b / p.cblks;
                    blk_c[j] = (b - tap * p.cblks) * 32 + (p.grouped ? n_tile * BLOCK_N : 0);
                    blk_dy[j] = tap / p.ksize - pad;
                    blk_dx[j] = tap % p.ksize - pad;
                    ++nvalid;
                } else {
                    blk_c[j] = 0; blk_dx[j] = 0; blk_dy[j] = 0;
                }
            }
            const uint32_t tx_bytes = static_cast<uint32_t>(nvalid * kWgBlkBytes + kBBytes);
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < nsteps; ++i) {
                int ks = ks_begin + i;
                const int tw = ks % p.steps_w;
                ks /= p.steps_w;
                const int th = ks % p.steps_h;
                const int tn = ks / p.steps_h;
                const int w0 = tw * p.pw, h0 = th * p.ph, n0 = tn * p.pn;
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                uint8_t* sa = smem_al + stage * kStageBytes;
                uint8_t* sb = sa + kABytes;
                if (leader) {
                    mbar_expect_tx(&full_bar[stage], tx_bytes);
                    for (int j = 0; j < nvalid; ++j)
                        tma_load_4d(sa + j * kWgBlkBytes, &tmX, &full_bar[stage], blk_c[j], w0 + blk_dx[j],
                                    h0 + blk_dy[j], n0);
#pragma unroll
                    for (int j = 0; j < BLOCK_N / 32; ++j)
                        tma_load_4d(sb + j * kWgBlkBytes, &tmDY, &full_bar[stage], n_tile * BLOCK_N + j * 32, w0, h0,
                                    n0);
                }
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        const bool leader = elect_one();
        {
            constexpr uint32_t idesc = make_idesc_tf32(kWgM, BLOCK_N, 1, 1);  // both operands MN-major
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < nsteps; ++i) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_base + stage * kStageBytes;
                const uint32_t b_addr = a_addr + kABytes;
                const int kiters = p.kstep >> 3;
                for (int k = 0; leader && k < kiters; ++k) {
                    // MN-major 32-bit operands must use SWIZZLE_128B_BASE32B (layout type 1): atoms of 4 K-rows x 128 B;
                    // LBO = stride between 32-channel blocks, SBO = stride between 4-row K groups (512 B);
                    // one K=8 MMA spans two atoms, so stepping K by 8 = +1024 bytes.
                    const uint64_t da = make_smem_desc(a_addr + k * 1024, kWgBlkBytes, 512, 1);
                    const uint64_t db = make_smem_desc(b_addr + k * 1024, kWgBlkBytes, 512, 1);
                    umma_tf32_ss(tmem_base, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
                }
                if (leader) umma_commit(&empty_bar[stage]);
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            if (leader) umma_commit(&tmem_full_bar);
        }
    } else {
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;
        const int row = m_tile * kWgM + m;
        const bool valid = row < p.rows_total;
        float* orow = p.ws + (static_cast<size_t>(split) * p.rows_total + row) * p.Cout + n_tile * BLOCK_N;
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c0, v);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                           __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                    *reinterpret_cast<float4*>(orow + c0 + j) = o;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// dw[co][ci][tap] = scale * sum_s ws[s][tap*Cin+ci][co]
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int splits, int rows,
                                    int Cout, int Cin, int taps, float scale) {
    // thread per (row, co) with co fastest for coalesced reads
    const long total = static_cast<long>(rows) * Cout;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        float acc = 0.f;
        for (int s = 0; s < splits; ++s) acc += ws[static_cast<long>(s) * total + i];
        const int co = i % Cout;
        const int row = i / Cout;
        const int tap = row / Cin;
        const int ci = row - tap * Cin;
        dw[(static_cast<long>(co) * Cin + ci) * taps + tap] = acc * scale;
    }
}


// Tiled form of the reduction for the fused "weight gradient -> parameter .grad" path (b200lp_conv_wgrad_sn_acc):
//   G[co][ci][tap] = sum_s ws[s][tap*Cin+ci][co]                    (deterministic order over the K splits)
//   grad[co][ci][tap] += (*inv_sigma) * G                           (or = when !accumulate)
//   dot_part[block]   = sum over the block's elements of G * w      (partials of <G, W>, summed in order by sn_rank1)
// A block owns a 32(ci) x 32(co) tile for TPB taps: workspace rows are read coalesced along co, transposed through
// shared memory, and grad / w are accessed along their contiguous (ci, tap) axis — wgrad_reduce_kernel's OIHW
// stores are one 4-byte element per 32-byte sector.
template <int TPB>
__global__ void __launch_bounds__(256, 2)
wgrad_reduce_acc_kernel(const float* __restrict__ ws, float* __restrict__ grad, const float* __restrict__ w,
                        const float* __restrict__ inv_sigma, float* __restrict__ dot_part, int splits, int Cin,
                        int Cout, int taps, int accumulate) {
    constexpr int kStride = 32 * TPB + 1;
    __shared__ float tile[32 * kStride];
    __shared__ float red[8];
    const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32, tap0 = blockIdx.z * TPB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long total = static_cast<long>(taps) * Cin * Cout;
    // phase 1: lane = co; warp w owns the (ci, tap) rows r = w + 8 i.  All R row loads of one split are independent and
    // in flight together (one load at a time per warp left this kernel latency-bound: 46 us for a 512 x 512 x 9 weight)
    constexpr int R = 32 * TPB / 8;
    float acc[R];
#pragma unroll
    for (int i = 0; i < R; ++i) acc[i] = 0.f;
    const float* src = ws + (static_cast<long>(tap0) * Cin + ci0 + warp) * Cout + co0 + lane;
    for (int s = 0; s < splits; ++s) {
#pragma unroll
        for (int i = 0; i < R; ++i)   // r = warp + 8 i  ->  tap t = i / 4, channel cl = 8 (i % 4) + warp
            acc[i] += __ldg(src + (static_cast<long>(i / 4) * Cin + 8 * (i % 4)) * Cout);
        src += total;
    }
#pragma unroll
    for (int i = 0; i < R; ++i) tile[lane * kStride + (8 * (i % 4) + warp) * TPB + i / 4] = acc[i];
    __syncthreads();
    // phase 2: one output channel per warp pass, lanes along the (ci, tap) axis
    const float sc = inv_sigma ? __ldg(inv_sigma) : 1.f;
    float dot = 0.f;
    // all 4 x TPB loads of grad / w are issued before the first dependent store (a load -> store chain per element
    // serialised this phase on the memory latency)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        int idx[2][TPB];                      // element offsets fit 32 bits (a weight has < 2^31 elements)
        float gold[2][TPB], wv[2][TPB];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int col = warp + 8 * (2 * half + q);
            const int base = ((co0 + col) * Cin + ci0) * taps + tap0;
#pragma unroll
            for (int k = 0; k < TPB; ++k) {
                const int j = lane + 32 * k;
                const int cl = j / TPB, t = j - cl * TPB;
                idx[q][k] = base + cl * taps + t;
                gold[q][k] = accumulate ? grad[idx[q][k]] : 0.f;
                wv[q][k] = w ? __ldg(w + idx[q][k]) : 0.f;
            }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int col = warp + 8 * (2 * half + q);
#pragma unroll
            for (int k = 0; k < TPB; ++k) {
                const float g = tile[col * kStride + lane + 32 * k];
                dot += g * wv[q][k];
                grad[idx[q][k]] = gold[q][k] + sc * g;
            }
        }
    }
    if (dot_part) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if (lane == 0) red[warp] = dot;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tsum = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) tsum += red[k];
            dot_part[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = tsum;
        }
    }
}

// grad[r][c] -= s^2 * <G, W> * u[r] * v[c]   (the rank-1 term of the spectral-norm gradient);
// <G, W> = sum of `nparts` partials in index order.  grid = (column chunks of 1024, row groups of 8)
__global__ void __launch_bounds__(256)
sn_rank1_acc_kernel(float* __restrict__ grad, const float* __restrict__ part, int nparts,
                    const float* __restrict__ inv_sigma, const float* __restrict__ u, const float* __restrict__ v,
                    int rows, int cols) {
    __shared__ float sh[8];
    __shared__ float kk;
    // every block sums the same partials in the same order: identical k everywhere, no atomics
    float a = 0.f;
    for (int i = threadIdx.x; i < nparts; i += 256) a += part[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tsum = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) tsum += sh[k];
        const float s = __ldg(inv_sigma);
        kk = tsum * s * s;
    }
    __syncthreads();
    const float k = kk;
    const int c = blockIdx.x * 1024 + threadIdx.x * 4;
    if (c >= cols) return;
    const int r0 = blockIdx.y * 8;
    const bool vec = (cols & 3) == 0 && ((reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (vec) {
        const float4 vv = *reinterpret_cast<const float4*>(v + c);
        for (int r = r0; r < r0 + 8 && r < rows; ++r) {
            const float ku = k * __ldg(u + r);
            float4* gp = reinterpret_cast<float4*>(grad + static_cast<long>(r) * cols + c);
            float4 g = *gp;
            g.x -= ku * vv.x; g.y -= ku * vv.y; g.z -= ku * vv.z; g.w -= ku * vv.w;
            *gp = g;
        }
    } else {
        for (int r = r0; r < r0 + 8 && r < rows; ++r) {
            const float ku = k * __ldg(u + r);
            for (int j = c; j < c + 4 && j < cols; ++j) grad[static_cast<long>(r) * cols + j] -= ku * __ldg(v + j);
        }
    }
}

struct WgPlan {
    int block_n, m_tiles, n_tiles, splits, steps_per_split, total_steps, pw, ph, pn, kstep, stages;
};

static int plan_wgrad(int N, int H, int W, int Cin, int Cout, int ksize, WgPlan* pl, int kstep_req = 0,
                      int stages_req = 0, int splits_req = 0, bool grouped = false) {
    if (!(ksize == 1 || ksize == 3) || Cin % 32 || Cout % 32 || Cin <= 0 || Cout <= 0 || N <= 0) return -1;
    if (ilog2_exact(H) < 1 || ilog2_exact(W) < 1) return -1;
    pl->block_n = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : 32));
    if (grouped) {
        if (Cin != Cout || ksize != 3) return -1;
        pl->block_n = 32;                       // one 32-channel block on both sides
    }
    pl->n_tiles = Cout / pl->block_n;
    // pixels per pipeline stage: 64 amortises the per-stage TMA issue / barrier cost when the whole batch offers
    // enough K (and the tile is not already 48 KB per 32 pixels)
    // (measured on B200, profiles/r01_conv_wgrad_tuning_sweep.log: 32-pixel stages with shallow rings — more CTAs per
    //  SM — beat 64-pixel stages on every shape of the step except one tie)
    //  After the elected-lane issue fix (profiles/r01_kernel_diag_elect_issue.log) Cout <= 64 tiles prefer 64-pixel stages in
    //  a 2-deep ring (336 vs 277 TFLOP/s at 64->64 @256x256); the wider tiles keep 32-pixel stages.
    int kstep = kstep_req ? kstep_req : (pl->block_n <= 64 ? 64 : 32);
    if (kstep != 32 && kstep != 64) return -1;
    if (static_cast<long>(N) * H * W < kstep) kstep = 32;
    if (!kstep_req && kstep == 64 && static_cast<long>(H) * W < 64 && N % (64 / (H * W))) kstep = 32;   // ragged batch
    pl->kstep = kstep;
    pl->pw = W < kstep ? W : kstep;
    pl->ph = (kstep / pl->pw) < H ? (kstep / pl->pw) : H;
    pl->pn = kstep / (pl->pw * pl->ph);
    const int steps_n = (N + pl->pn - 1) / pl->pn;
    pl->total_steps = (W / pl->pw) * (H / pl->ph) * steps_n;
    const int stage_bytes = (4 + pl->block_n / 32) * kstep * 128;
    int stages = stages_req ? stages_req : ((pl->block_n == 256 || kstep == 64) ? 2 : 3);   // <= ~100 KB per CTA: two CTAs share an SM
    if (stages > kWgMaxStages) stages = kWgMaxStages;
    while (stages > 1 && stages * stage_bytes + 1024 > kWgMaxSmem) --stages;
    pl->stages = stages;
    const int blocks_total = ksize * ksize * (grouped ? 1 : Cin / 32);
    pl->m_tiles = (blocks_total + 3) / 4;
    const int base = pl->m_tiles * pl->n_tiles;
    // split-K so that the grid is at most one full wave of 2 CTAs x 148 SMs (one CTA over the wave costs a whole
    // extra wave), with at least 4 K steps per CTA
    int splits = splits_req ? splits_req : (2 * 148) / base;
    if (!splits_req && splits > pl->total_steps / 4) splits = pl->total_steps / 4;
    // at most 64 splits, except for single-tile grids (the Cin = 3 stems' 27(+5)-column patch GEMM at 256 x 256: 64 CTAs
    // ran 49 us): those may take one CTA per SM
    if (splits > (base == 1 ? 148 : 64)) splits = base == 1 ? 148 : 64;
    if (splits > pl->total_steps) splits = pl->total_steps;
    if (splits < 1) splits = 1;
    pl->steps_per_split = (pl->total_steps + splits - 1) / splits;
    pl->splits = (pl->total_steps + pl->steps_per_split - 1) / pl->steps_per_split;  // no empty split
    return 0;
}

template <int BLOCK_N>
static int launch_wgrad(const CUtensorMap& tmX, const CUtensorMap& tmDY, const WgradParams& p, const WgPlan& pl,
                        cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        B200LP_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_tf32_kernel<BLOCK_N>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, kWgMaxSmem));
        B200LP_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_tf32_kernel<BLOCK_N>,
                                               cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_set = true;
    }
    const int smem_bytes = pl.stages * (4 + BLOCK_N / 32) * pl.kstep * 128 + 1024;
    dim3 grid(pl.n_tiles, pl.m_tiles, pl.splits);
    conv_wgrad_tf32_kernel<BLOCK_N><<<grid, kWgThreads, smem_bytes, stream>>>(tmX, tmDY, p);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

}  // namespace b200lp

using namespace b200lp;

extern "C" int64_t b200lp_conv_wgrad_workspace(int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                                               int32_t ksize) {
    WgPlan pl;
    if (plan_wgrad(N, H, W, Cin, Cout, ksize, &pl)) {
        set_error("conv_wgrad_workspace: unsupported shape N=%d H=%d W=%d Cin=%d Cout=%d k=%d", N, H, W, Cin, Cout,
                  ksize);
        return B200LP_EINVAL;
    }
    return static_cast<int64_t>(pl.splits) * ksize * ksize * Cin * Cout * 4;   // default plan (tuning knobs = auto)
}

// plans and launches the tensor-core kernel (split-K partial sums -> a->workspace); the caller adds the reduction
static int wgrad_main(const b200lp_wgrad_args* a, void* stream, WgPlan* pl_out, int64_t extra_ws_bytes,
                      bool grouped = false) {
    B200LP_REQUIRE(a && a->x && a->dy && a->dw && a->workspace, "conv_wgrad: null pointer");
    B200LP_REQUIRE(grouped || a->grouped == 0, "conv_wgrad: grouped weights go through b200lp_gconv3x3_wgrad_tc");
    WgPlan& pl = *pl_out;
    B200LP_REQUIRE(plan_wgrad(a->N, a->H, a->W, a->Cin, a->Cout, a->ksize, &pl, a->kstep, a->stages, a->splits, grouped) == 0,
                   "conv_wgrad: unsupported shape N=%d H=%d W=%d Cin=%d Cout=%d k=%d", a->N, a->H, a->W, a->Cin,
                   a->Cout, a->ksize);
    B200LP_REQUIRE(a->N % pl.pn == 0, "conv_wgrad: N=%d must be a multiple of %d for %dx%d planes", a->N, pl.pn,
                   a->H, a->W);
    const int taps = a->ksize * a->ksize;
    const int64_t need = static_cast<int64_t>(pl.splits) * taps * (grouped ? 32 : a->Cin) * a->Cout * 4 + extra_ws_bytes;
    B200LP_REQUIRE(a->workspace_bytes >= need, "conv_wgrad: workspace %lld < %lld bytes",
                   (long long)a->workspace_bytes, (long long)need);

    WgradParams p;
    p.ws = a->workspace;
    p.N = a->N; p.H = a->H; p.W = a->W; p.Cin = a->Cin; p.Cout = a->Cout;
    p.ksize = a->ksize;
    p.pw = pl.pw; p.ph = pl.ph; p.pn = pl.pn;
    p.kstep = pl.kstep;
    p.blk_bytes = pl.kstep * 128;
    p.stages = pl.stages;
    p.steps_w = a->W / pl.pw;
    p.steps_h = a->H / pl.ph;
    p.total_steps = pl.total_steps;
    p.steps_per_split = pl.steps_per_split;
    p.grouped = grouped ? 1 : 0;
    p.cblks = grouped ? 1 : a->Cin / 32;
    p.rows_total = taps * p.cblks * 32;
    p.blocks_total = taps * p.cblks;

    CUtensorMap tmX, tmDY;
    const uint32_t box[4] = {32u, (uint32_t)pl.pw, (uint32_t)pl.ph, (uint32_t)pl.pn};
    {
        const uint64_t dims[4] = {(uint64_t)a->Cin, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->N};
        const uint64_t strides[3] = {(uint64_t)a->Cin * 4, (uint64_t)a->W * a->Cin * 4,
                                     (uint64_t)a->H * a->W * a->Cin * 4};
        int r = encode_tmap_f32(&tmX, a->x, 4, dims, strides, box, true);
        if (r) return r;
    }
    {
        const uint64_t dims[4] = {(uint64_t)a->Cout, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->N};
        const uint64_t strides[3] = {(uint64_t)a->Cout * 4, (uint64_t)a->W * a->Cout * 4,
                                     (uint64_t)a->H * a->W * a->Cout * 4};
        int r = encode_tmap_f32(&tmDY, a->dy, 4, dims, strides, box, true);
        if (r) return r;
    }
    cudaStream_t s = as_stream(stream);
    int r;
    switch (pl.block_n) {
        case 256: r = launch_wgrad<256>(tmX, tmDY, p, pl, s); break;
        case 128: r = launch_wgrad<128>(tmX, tmDY, p, pl, s); break;
        case 64: r = launch_wgrad<64>(tmX, tmDY, p, pl, s); break;
        default: r = launch_wgrad<32>(tmX, tmDY, p, pl, s); break;
    }
    return r;
}

extern "C" int32_t b200lp_conv_wgrad(const b200lp_wgrad_args* a, void* stream) {
    WgPlan pl;
    int r = wgrad_main(a, stream, &pl, 0);
    if (r) return r;
    const int taps = a->ksize * a->ksize;
    const int rows_total = taps * a->Cin;
    const long total = static_cast<long>(rows_total) * a->Cout;
    int blocks = static_cast<int>((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    wgrad_reduce_kernel<<<blocks, 256, 0, as_stream(stream)>>>(a->workspace, a->dw, pl.splits, rows_total, a->Cout,
                                                               a->Cin, taps, a->scale);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

// number of <G, W> partial sums wgrad_reduce_acc_kernel writes, and whether a block takes all taps of its tile
static int wgrad_acc_parts(int Cin, int Cout, int ksize, bool* all_taps) {
    const int tiles = (Cin / 32) * (Cout / 32);
    const int taps = ksize * ksize;
    // fewer than 256 tiles (up to 256 x 512 channels): one block per tap — 9x the blocks; measured with one block per
    // tile: 64 blocks 50 us, 128 blocks 28 us, 256 blocks 16.5 us (profiles/r01_ncu_launch_list_graph_step.csv).  The
    // strided OIHW accesses of the per-tap form stay in L2 (a weight is <= 4.7 MB there)
    *all_taps = taps == 1 || tiles >= 256;
    return *all_taps ? tiles : tiles * taps;
}

extern "C" int64_t b200lp_conv_wgrad_sn_acc_workspace(int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                                                      int32_t ksize) {
    const int64_t base = b200lp_conv_wgrad_workspace(N, H, W, Cin, Cout, ksize);
    if (base < 0) return base;
    bool all_taps;
    return base + static_cast<int64_t>(wgrad_acc_parts(Cin, Cout, ksize, &all_taps)) * 4;
}

extern "C" int32_t b200lp_conv_wgrad_sn_acc(const b200lp_wgrad_args* a, const float* w, const float* inv_sigma,
                                            const float* u, const float* v, int32_t accumulate, void* stream) {
    B200LP_REQUIRE(a && (inv_sigma == nullptr || (w && u && v)), "conv_wgrad_sn_acc: w, u, v are required with inv_sigma");
    B200LP_REQUIRE(a->splits == 0 && a->kstep == 0 && a->stages == 0, "conv_wgrad_sn_acc: tuning knobs are not supported");
    bool all_taps;
    const int taps = a->ksize * a->ksize;
    const int nparts = (a->Cin > 0 && a->Cout > 0) ? wgrad_acc_parts(a->Cin, a->Cout, a->ksize, &all_taps) : 0;
    WgPlan pl;
    int r = wgrad_main(a, stream, &pl, static_cast<int64_t>(nparts) * 4);
    if (r) return r;
    cudaStream_t s = as_stream(stream);
    float* parts = a->workspace + static_cast<size_t>(pl.splits) * taps * a->Cin * a->Cout;
    const bool sn = inv_sigma != nullptr;
    dim3 grid(a->Cin / 32, a->Cout / 32, all_taps ? 1 : taps);
    if (taps == 9 && all_taps)
        wgrad_reduce_acc_kernel<9><<<grid, 256, 0, s>>>(a->workspace, a->dw, sn ? w : nullptr, inv_sigma,
                                                        sn ? parts : nullptr, pl.splits, a->Cin, a->Cout, taps, accumulate);
    else
        wgrad_reduce_acc_kernel<1><<<grid, 256, 0, s>>>(a->workspace, a->dw, sn ? w : nullptr, inv_sigma,
                                                        sn ? parts : nullptr, pl.splits, a->Cin, a->Cout, taps, accumulate);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    if (sn) {
        const int rows = a->Cout, cols = a->Cin * taps;
        dim3 g2((cols + 1023) / 1024, (rows + 7) / 8);
        sn_rank1_acc_kernel<<<g2, 256, 0, s>>>(a->dw, parts, nparts, inv_sigma, u, v, rows, cols);
        B200LP_CHECK_CUDA(cudaGetLastError());
        count_launch();
    }
    return B200LP_OK;
}

namespace b200lp {
// dw[co][cig][tap] (+)= sum_s ws[s][tap*32 + (co%32/cpg)*cpg + cig][co]: the block-diagonal entries of the per-block
// [9 x 32 ci] x [32 co] products, summed over the K splits in index order.  One thread per (co, tap, cig) with co fastest
// across the warp (coalesced workspace reads; the OIHW stores hit a <= 1.2 MB tensor).
__global__ void __launch_bounds__(256)
gconv_wgrad_tc_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int splits, int C, int cpg, int accumulate) {
    const long total = static_cast<long>(C) * cpg * 9;
    const long split_stride = 288L * C;
    for (long i = blockIdx.x * 256L + threadIdx.x; i < total; i += gridDim.x * 256L) {
        const int co = static_cast<int>(i % C);
        const int r = static_cast<int>(i / C);
        const int cig = r % cpg, tap = r / cpg;
        const int row = tap * 32 + ((co & 31) / cpg) * cpg + cig;
        const float* src = ws + static_cast<long>(row) * C + co;
        float acc = 0.f;
        for (int s = 0; s < splits; ++s) acc += __ldg(src + s * split_stride);
        float* o = dw + (static_cast<long>(co) * cpg + cig) * 9 + tap;
        *o = accumulate ? *o + acc : acc;
    }
}

__global__ void __launch_bounds__(256)
zero_stuff2_kernel(const float4* __restrict__ x, float4* __restrict__ out, long total4, int W2, int H2, int c4) {
    for (long i = blockIdx.x * 256L + threadIdx.x; i < total4; i += gridDim.x * 256L) {
        const int c = static_cast<int>(i % c4);
        long pix = i / c4;
        const int w = static_cast<int>(pix % W2);
        pix /= W2;
        const int h = static_cast<int>(pix % H2);
        const long n = pix / H2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (((w | h) & 1) == 0) v = __ldg(x + ((n * (H2 >> 1) + (h >> 1)) * (W2 >> 1) + (w >> 1)) * c4 + c);
        out[i] = v;
    }
}
}  // namespace b200lp

extern "C" int64_t b200lp_gconv3x3_wgrad_tc_workspace(int32_t N, int32_t H, int32_t W, int32_t C) {
    WgPlan pl;
    if (plan_wgrad(N, H, W, C, C, 3, &pl, 0, 0, 0, true)) {
        set_error("gconv3x3_wgrad_tc_workspace: unsupported shape N=%d H=%d W=%d C=%d", N, H, W, C);
        return B200LP_EINVAL;
    }
    return static_cast<int64_t>(pl.splits) * 288 * C * 4;
}

extern "C" int32_t b200lp_gconv3x3_wgrad_tc(const b200lp_wgrad_args* a, int32_t accumulate, void* stream) {
    B200LP_REQUIRE(a && a->grouped > 0 && a->grouped <= 32 && 32 % a->grouped == 0 && a->Cin == a->Cout && a->ksize == 3,
                   "gconv3x3_wgrad_tc: needs grouped = cpg in {1,2,4,8,16,32}, Cin == Cout, ksize 3");
    B200LP_REQUIRE(a->splits == 0 && a->kstep == 0 && a->stages == 0, "gconv3x3_wgrad_tc: tuning knobs are not supported");
    WgPlan pl;
    int r = wgrad_main(a, stream, &pl, 0, true);
    if (r) return r;
    const long total = static_cast<long>(a->Cout) * a->grouped * 9;
    int blocks = static_cast<int>((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    gconv_wgrad_tc_reduce_kernel<<<blocks, 256, 0, as_stream(stream)>>>(a->workspace, a->dw, pl.splits, a->Cout, a->grouped,
                                                                        accumulate);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_zero_stuff2(const float* x, float* out, int32_t N, int32_t H, int32_t W, int32_t C, void* stream) {
    B200LP_REQUIRE(x && out && N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "zero_stuff2: bad args");
    const long total4 = static_cast<long>(N) * (2 * H) * (2 * W) * (C / 4);
    long blocks = (total4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    zero_stuff2_kernel<<<static_cast<int>(blocks), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(out), total4, 2 * W, 2 * H, C / 4);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}
