// Implicit-GEMM 3x3 / 1x1 convolution (stride 1, zero "same" padding) on the sm_100a tensor cores.
//
//   GEMM view:  M = output pixels (a 128-pixel spatial patch per CTA),  N = Cout,  K = taps * Cin
//   A operand:  NHWC activations, fetched by TMA as a 4-D box (32 channels x bw x bh x bn pixels) whose origin is
//               shifted by the filter tap; out-of-image pixels are zero-filled by the TMA unit (= zero padding).
//               In shared memory the box is 128 rows (pixels) x 128 bytes (32 tf32 channels), 128B-swizzled:
//               exactly the canonical K-major SWIZZLE_128B UMMA operand.
//   B operand:  packed weights [Cout][tap][Cin] (K contiguous), TMA 2-D box (32 x BLOCK_N), same canonical layout.
//   D:          fp32 accumulator, 128 lanes x BLOCK_N columns of TMEM; tcgen05.mma.kind::tf32, one issuing thread.
//
//   Warp roles (192 threads):  warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue
//   (tcgen05.ld -> bias / residual / relu / tf32 rounding -> 128-byte vector stores, one pixel row per thread).
//
// Two operand precisions share the kernel (template parameter MODE):
//   MODE 0  "tf32"  : fp32 tensors holding TF32-rounded values, one tcgen05.mma.kind::tf32 per K=8 step.
//   MODE 1  "bf16x3": every operand is a (hi, lo) pair of bf16 tensors with hi + lo == the fp32 value to 2^-17;
//                     three tcgen05.mma.kind::f16 per K=16 step (Ah*Bh + Ah*Bl + Al*Bh) accumulate in the same
//                     fp32 TMEM tile.  1.5x the tensor time of MODE 0, ~45x smaller operand-rounding error: this is
//                     what keeps the generator output within 1e-3 of the fp32 reference.
//
// Replaces torch's nn.Conv2d forward (and, with transposed packing, conv backward-data) at the call sites listed
// in include/b200lp.h.
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace b200lp {

constexpr int kBlockM = 128;      // pixels per CTA tile
constexpr int kRowBytes = 128;    // one smem row = 128 bytes = swizzle span = 32 tf32 or 64 bf16 channels
constexpr int kConvThreads = 192;

struct ConvParams {
    const float* out_scale;   // optional device scalar: accumulator *= *out_scale before bias (1/sigma of spectral norm)
    const float* bias;
    const float* residual;
    float* y;
    __nv_bfloat16* y_split;   // optional extra output: (hi, lo) bf16 planes of y, lo plane at +split_stride elements
    long long split_stride;
    int N, H, W, Cin, Cout;
    int ksize;
    int bw, bh, bn;        // patch shape: bw*bh*bn == 128
    int tiles_w, tiles_h;  // patches per image row / column
    int cblks;             // channel blocks per tap: ceil(Cin / channels-per-row)
    int num_kb;            // taps * cblks
    int residual_mode;
    int relu;
    int round_out;
    uint32_t a_bytes;      // bytes one A box delivers (may be < 16 KB when bn > N)
    int stages;            // depth of the smem ring (<= ConvCfg::kMaxStages)
    int total_tiles;       // pixel patches x Cout tiles x K splits
    int splits;            // split-K factor (1 = none): small-plane layers have too few tiles to fill 148 SMs
    int kb_per_split;      // K blocks per split
    float* ws;             // split-K partial sums [splits][N*H*W][Cout] (raw accumulators), NULL when splits == 1
    long long ws_stride;   // N*H*W*Cout
    int a_stages, b_stages;   // halo kernel: depths of the activation-slab ring and of the weight-tile ring
    int kcin;              // input channels per filter tap in the packed weights: Cin, or BLOCK_N when grouped
    int grouped;           // 1: block-diagonal (grouped) convolution — N tile n reads input channels [n*BLOCK_N, (n+1)*BLOCK_N)
};

template <int BLOCK_N, int MODE>
struct ConvCfg {
    static constexpr int kParts = MODE == 0 ? 1 : 2;               // operand planes (tf32: 1, bf16 hi/lo: 2)
    static constexpr int kChanPerRow = MODE == 0 ? 32 : 64;        // channels in one 128-byte row
    static constexpr int kABytes = kBlockM * kRowBytes;            // 16 KB per plane
    static constexpr int kBBytes = BLOCK_N * kRowBytes;
    static constexpr int kStageBytes = kParts * (kABytes + kBBytes);
    static constexpr int kMaxStages = 8;
    static constexpr int kMaxSmemBytes = 226 * 1024;   // 227 KB opt-in limit minus the static barriers
    // two accumulators in TMEM: the epilogue drains one while the MMAs of the next tile fill the other
    static constexpr uint32_t kAccCols = BLOCK_N < 32 ? 32 : BLOCK_N;
    static constexpr uint32_t kTmemCols = 2 * kAccCols;
};

struct ConvMaps {
    CUtensorMap a[2];   // activation planes (hi, lo)
    CUtensorMap b[2];   // weight planes
};

// Epilogue of one 32-column chunk of one pixel row: v = raw accumulators of channels [c0, c0+32) of this thread's pixel.
// Either raw partial sums into the split-K workspace, or  y = round(relu(acc * oscale + bias + residual))  (+ the (hi, lo)
// bf16 planes of y for a following bf16x3 conv).  `elem` = pix * Cout + first channel of the tile.
__device__ __forceinline__ void epilogue_store_chunk(const ConvParams& p, const uint32_t (&v)[32], float oscale,
                                                     float* yrow, float* wsrow, const float* rrow, const float* brow,
                                                     size_t elem, int c0) {
    if (wsrow) {          // split-K: raw partial sums; the epilogue runs in splitk_epilogue_kernel
#pragma unroll
        for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(wsrow + c0 + j) =
                make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                            __uint_as_float(v[j + 3]));
        return;
    }
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        float4 o;
        o.x = __uint_as_float(v[j + 0]) * oscale;
        o.y = __uint_as_float(v[j + 1]) * oscale;
        o.z = __uint_as_float(v[j + 2]) * oscale;
        o.w = __uint_as_float(v[j + 3]) * oscale;
        if (brow) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(brow + c0 + j));
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
        }
        if (rrow) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(rrow + c0 + j));
            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        if (p.relu) {
            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
        }
        if (p.round_out) {
            o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w);
        }
        *reinterpret_cast<float4*>(yrow + c0 + j) = o;
        if (p.y_split) {   // (hi, lo) bf16 planes of the same values, for a following bf16x3 conv
            __nv_bfloat16* sp = p.y_split + elem + c0 + j;
            const __nv_bfloat16 h0 = __float2bfloat16_rn(o.x), h1 = __float2bfloat16_rn(o.y),
                                h2 = __float2bfloat16_rn(o.z), h3 = __float2bfloat16_rn(o.w);
            __nv_bfloat162 hi01 = __halves2bfloat162(h0, h1), hi23 = __halves2bfloat162(h2, h3);
            __nv_bfloat162 lo01 = __halves2bfloat162(__float2bfloat16_rn(o.x - __bfloat162float(h0)),
                                                     __float2bfloat16_rn(o.y - __bfloat162float(h1)));
            __nv_bfloat162 lo23 = __halves2bfloat162(__float2bfloat16_rn(o.z - __bfloat162float(h2)),
                                                     __float2bfloat16_rn(o.w - __bfloat162float(h3)));
            uint2 hv, lv;
            hv.x = *reinterpret_cast<uint32_t*>(&hi01); hv.y = *reinterpret_cast<uint32_t*>(&hi23);
            lv.x = *reinterpret_cast<uint32_t*>(&lo01); lv.y = *reinterpret_cast<uint32_t*>(&lo23);
            *reinterpret_cast<uint2*>(sp) = hv;
            *reinterpret_cast<uint2*>(sp + p.split_stride) = lv;
        }
    }
}

// Persistent, warp-specialised: every CTA walks the tile list  tile = blockIdx.x, blockIdx.x + gridDim.x, ...
// (consecutive tiles = the N tiles of one pixel patch, so co-resident CTAs share the activation tile in L2).
// The shared-memory ring and the two TMEM accumulators stay live across tiles: no per-tile prologue, and the
// epilogue of tile i overlaps the main loop of tile i+1.
template <int BLOCK_N, int MODE>
__global__ void __launch_bounds__(kConvThreads, 2)
conv_igemm_kernel(const __grid_constant__ ConvMaps tm, const ConvParams p) {
    using Cfg = ConvCfg<BLOCK_N, MODE>;
    const int kStages = p.stages;

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[Cfg::kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[Cfg::kMaxStages];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_slot;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_tiles = p.Cout / BLOCK_N;

    if (warp == 0 && lane == 0) {
#pragma unroll
        for (int q = 0; q < Cfg::kParts; ++q) {
            tma_prefetch_desc(&tm.a[q]);
            tma_prefetch_desc(&tm.b[q]);
        }
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], 4);      // one arrival per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc<Cfg::kTmemCols>(&tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        // The whole warp walks the loop (warp-uniform control flow) and one elected lane issues: under a divergent
        // `if (lane == 0)` ptxas wraps every TMA / MMA / commit in an ELECT + BRA.U.ANY loop (~8 extra instructions
        // per MMA on the single issuing thread, measured 76 clk per MMA against a 64-clk N=128 instruction).
        const bool leader = elect_one();
        {
            const int pad = p.ksize >> 1;
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int ks = tile % p.splits;
                const int mn = tile / p.splits;
                const int n_tile = mn % n_tiles;
                int m_tile = mn / n_tiles;
                const int tw = m_tile % p.tiles_w;
                m_tile /= p.tiles_w;
                const int th = m_tile % p.tiles_h;
                const int tn = m_tile / p.tiles_h;
                const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
                const int kb0 = ks * p.kb_per_split;
                const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    const int tap = kb / p.cblks;
                    const int cb = kb - tap * p.cblks;
                    const int dy = tap / p.ksize - pad;
                    const int dx = tap % p.ksize - pad;
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    // stage layout: [A plane 0][A plane 1]...[B plane 0][B plane 1]..., every plane 1024-byte aligned
                    uint8_t* sa = smem_al + stage * Cfg::kStageBytes;
                    uint8_t* sb = sa + Cfg::kParts * Cfg::kABytes;
                    const int kcoord = tap * p.kcin + cb * Cfg::kChanPerRow;   // column of the packed weight matrix
                    const int ccoord = cb * Cfg::kChanPerRow + (p.grouped ? n_tile * BLOCK_N : 0);
                    if (leader) {
                        mbar_expect_tx(&full_bar[stage], Cfg::kParts * (p.a_bytes + Cfg::kBBytes));
#pragma unroll
                        for (int q = 0; q < Cfg::kParts; ++q) {
                            tma_load_4d(sa + q * Cfg::kABytes, &tm.a[q], &full_bar[stage], ccoord, w0 + dx, h0 + dy, n0);
                            tma_load_2d(sb + q * Cfg::kBBytes, &tm.b[q], &full_bar[stage], kcoord, n_tile * BLOCK_N);
                        }
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform loop, elected lane issues) =====================
        const bool leader = elect_one();
        {
            constexpr uint32_t idesc = MODE == 0 ? make_idesc_tf32(kBlockM, BLOCK_N, 0, 0)
                                                 : make_idesc_bf16(kBlockM, BLOCK_N, 0, 0);
            // K-major SWIZZLE_128B descriptors: 8-row groups 1024 B apart (SBO); only the start address changes
            const uint64_t desc_hi = make_smem_desc(0, 16, 1024, 2);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);      // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + acc * Cfg::kAccCols;
                const int nkb = min(p.num_kb, (tile % p.splits + 1) * p.kb_per_split) - (tile % p.splits) * p.kb_per_split;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_base + stage * Cfg::kStageBytes;
                    const uint32_t b_addr = a_addr + Cfg::kParts * Cfg::kABytes;
                    const uint64_t da0 = desc_hi | ((a_addr >> 4) & 0x3FFFu);
                    const uint64_t db0 = desc_hi | ((b_addr >> 4) & 0x3FFFu);
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            // one MMA consumes 32 bytes of K (8 tf32 or 16 bf16): +32 B inside the 128-byte swizzle span
                            const uint64_t da = da0 + 2 * k, db = db0 + 2 * k;
                            if (MODE == 0) {
                                umma_tf32_ss(tmem_acc, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                            } else {
                                const uint64_t da_lo = da + (Cfg::kABytes >> 4), db_lo = db + (Cfg::kBBytes >> 4);
                                umma_f16_ss(tmem_acc, da_lo, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);   // Al * Bh
                                umma_f16_ss(tmem_acc, da, db_lo, idesc, 1u);                            // Ah * Bl
                                umma_f16_ss(tmem_acc, da, db, idesc, 1u);                               // Ah * Bh
                            }
                        }
                        umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above have read it
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
                if (leader) umma_commit(&tmem_full_bar[acc]);    // accumulator complete
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quarter = warp & 3;               // TMEM lane quarter this warp may access
        const int m = quarter * 32 + lane;          // row of the tile = pixel of the patch
        const int wl = m % p.bw;
        const int hl = (m / p.bw) % p.bh;
        const int nl = m / (p.bw * p.bh);
        const float oscale = p.out_scale ? __ldg(p.out_scale) : 1.0f;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int ks = tile % p.splits;
            const int mn = tile / p.splits;
            const int n_tile = mn % n_tiles;
            int m_tile = mn / n_tiles;
            const int tw = m_tile % p.tiles_w;
            m_tile /= p.tiles_w;
            const int th = m_tile % p.tiles_h;
            const int tn = m_tile / p.tiles_h;
            const int n = tn * p.bn + nl, h = th * p.bh + hl, w = tw * p.bw + wl;
            const bool valid = n < p.N;
            const size_t pix = (static_cast<size_t>(n) * p.H + h) * p.W + w;
            float* yrow = p.y + pix * p.Cout + n_tile * BLOCK_N;
            float* wsrow = p.ws ? p.ws + ks * p.ws_stride + pix * p.Cout + n_tile * BLOCK_N : nullptr;
            const float* rrow = nullptr;
            if (p.residual_mode == 1) {
                rrow = p.residual + pix * p.Cout + n_tile * BLOCK_N;
            } else if (p.residual_mode == 2) {
                const size_t rp = (static_cast<size_t>(n) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
                rrow = p.residual + rp * p.Cout + n_tile * BLOCK_N;
            }
            const float* brow = p.bias ? p.bias + n_tile * BLOCK_N : nullptr;

            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + acc * Cfg::kAccCols + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_acc + c0, v);
                tmem_ld_wait();
                if (c0 + 32 >= BLOCK_N) {
                    // all of this warp's TMEM reads of the accumulator are done: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
                }
                if (valid) epilogue_store_chunk(p, v, oscale, yrow, wsrow, rrow, brow, pix * p.Cout + n_tile * BLOCK_N, c0);
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

// ====================================================================================================================
// Halo variant for 3x3 layers on planes >= 16 x 8 (all the FLOP-heavy layers of the step).
//
// The per-tap kernel above re-fetches the activation patch once per filter tap (9x) and the weight tile once per
// 128-pixel patch; every instantiation of it measured at the L2->SM fabric limit (~40 B/clk/SM, ncu: tensor pipe 32 % /
// 50 % / 72 % active at BLOCK_N 64 / 128 / 256 == bytes-per-FLOP of the tile).  This variant cuts the bytes:
//   * a CTA tile is 8 pixels wide and PH = 16*MT rows tall (MT accumulators of 128 pixels each share every weight tile);
//   * per 32-channel block the activations arrive as THREE column-shifted slabs (dx = -1, 0, +1), each
//     8 x (PH + 2) pixels = (PH + 2) KB.  A slab row is one 8-pixel x 128-byte swizzle atom (1024 B), so the A operand
//     of tap (dy, dx) for sub-tile mt is the same slab at byte offset (16*mt + dy + 1) * 1024: 1024-byte aligned, the
//     plain K-major SWIZZLE_128B descriptor with SBO = 1024 — only the start address moves.
//   Activation bytes per 128 pixels and channel block: 3*(16*MT+2)/MT KB instead of 144 KB; weight bytes / MT.
// Two independent rings (activation slabs, weight tiles) with one producer warp each; the MMA warp walks
// channel block -> dx -> dy -> sub-tile -> 4 K steps.
template <int BLOCK_N, int MT, int MODE>
struct HaloCfg {
    static constexpr int kParts = MODE == 0 ? 1 : 2;
    static constexpr int kChanPerRow = MODE == 0 ? 32 : 64;
    static constexpr int kPH = 16 * MT;                               // patch height
    static constexpr int kSlabBytes = (kPH + 2) * 1024;               // one plane of one dx slab
    static constexpr int kAUnitBytes = kParts * kSlabBytes;
    static constexpr int kBPlaneBytes = BLOCK_N * kRowBytes;
    static constexpr int kBUnitBytes = kParts * kBPlaneBytes;
    static constexpr int kMaxAStages = 6, kMaxBStages = 8;
    static constexpr int kMaxSmemBytes = 226 * 1024;
    static constexpr uint32_t kTileCols = MT * BLOCK_N;               // accumulator columns of one CTA tile
    static constexpr int kAccBufs = 2 * kTileCols <= 512 ? 2 : 1;     // double-buffer when TMEM allows
    static constexpr uint32_t kTmemCols = kAccBufs * kTileCols < 32 ? 32 : kAccBufs * kTileCols;
    static_assert(kTileCols <= 512, "accumulators exceed TMEM");
    // MMA-issuing warps: the issue loop (barrier wait, ~3 uniform-datapath instructions per descriptor, commit) costs about
    // as many cycles as a 64-column MMA executes (ncu: the issuing warp never waits on data, 22 % MIO-queue stalls), so
    // with MT >= 2 the sub-tiles are split between two issuing warps, each owning its accumulators (+12..23 %; four
    // issuers at MT = 4 measured no further gain — profiles/r01_kernel_diag_halo_issuers.log).
    static constexpr int kIssuers = MT >= 2 ? 2 : 1;
    static constexpr int kThreads = 224 + 32 * (kIssuers - 1);
};

// warp 0: slab producer, 1: MMA issuer, 2..5: epilogue, 6: weight producer, 7: second MMA issuer (MT >= 2)

template <int BLOCK_N, int MT, int MODE>
__global__ void __launch_bounds__(HaloCfg<BLOCK_N, MT, MODE>::kThreads, 1)
conv_halo_kernel(const __grid_constant__ ConvMaps tm, const ConvParams p) {
    using Cfg = HaloCfg<BLOCK_N, MT, MODE>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t a_full[Cfg::kMaxAStages];
    __shared__ __align__(8) uint64_t a_empty[Cfg::kMaxAStages];
    __shared__ __align__(8) uint64_t b_full[Cfg::kMaxBStages];
    __shared__ __align__(8) uint64_t b_empty[Cfg::kMaxBStages];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_slot;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    const int SA = p.a_stages, SB = p.b_stages;
    const uint32_t b_ring_off = SA * Cfg::kAUnitBytes;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_tiles = p.Cout / BLOCK_N;

    if (warp == 0 && lane == 0) {
#pragma unroll
        for (int q = 0; q < Cfg::kParts; ++q) {
            tma_prefetch_desc(&tm.a[q]);
            tma_prefetch_desc(&tm.b[q]);
        }
        for (int s = 0; s < SA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], Cfg::kIssuers); }
        for (int s = 0; s < SB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], Cfg::kIssuers); }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], Cfg::kIssuers);   // one commit per issuing warp
            mbar_init(&tmem_empty_bar[a], 4);              // one arrival per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<Cfg::kTmemCols>(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ===================== activation-slab producer =====================
        const bool leader = elect_one();
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                int m_tile = tile / n_tiles;
                const int tw = m_tile % p.tiles_w;
                m_tile /= p.tiles_w;
                const int th = m_tile % p.tiles_h;
                const int n = m_tile / p.tiles_h;
                const int w0 = tw * 8, h0 = th * Cfg::kPH;
                const int cbase = p.grouped ? (tile % n_tiles) * BLOCK_N : 0;     // grouped: the N tile's own channels
                for (int cb = 0; cb < p.cblks; ++cb) {
                    for (int dx = -1; dx <= 1; ++dx) {
                        mbar_wait(&a_empty[stage], phase ^ 1u);
                        uint8_t* sa = smem_al + stage * Cfg::kAUnitBytes;
                        if (leader) {
                            mbar_expect_tx(&a_full[stage], Cfg::kAUnitBytes);
#pragma unroll
                            for (int q = 0; q < Cfg::kParts; ++q)
                                tma_load_4d(sa + q * Cfg::kSlabBytes, &tm.a[q], &a_full[stage],
                                            cbase + cb * Cfg::kChanPerRow, w0 + dx, h0 - 1, n);
                        }
                        if (++stage == SA) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 6) {
        // ===================== weight-tile producer =====================
        const bool leader = elect_one();
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int n_tile = tile % n_tiles;
                for (int cb = 0; cb < p.cblks; ++cb) {
                    for (int dxi = 0; dxi < 3; ++dxi) {
                        for (int dyi = 0; dyi < 3; ++dyi) {
                            mbar_wait(&b_empty[stage], phase ^ 1u);
                            uint8_t* sb = smem_al + b_ring_off + stage * Cfg::kBUnitBytes;
                            const int kcoord = (dyi * 3 + dxi) * p.kcin + cb * Cfg::kChanPerRow;
                            if (leader) {
                                mbar_expect_tx(&b_full[stage], Cfg::kBUnitBytes);
#pragma unroll
                                for (int q = 0; q < Cfg::kParts; ++q)
                                    tma_load_2d(sb + q * Cfg::kBPlaneBytes, &tm.b[q], &b_full[stage], kcoord,
                                                n_tile * BLOCK_N);
                            }
                            if (++stage == SB) { stage = 0; phase ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1 || warp == 7) {
        // ===================== MMA issuer(s) =====================
        const bool leader = elect_one();
        constexpr int kMtPer = MT / Cfg::kIssuers;          // sub-tiles (accumulators) owned by this issuing warp
        const int mt_first = warp == 1 ? 0 : kMtPer;
        {
            constexpr uint32_t idesc = MODE == 0 ? make_idesc_tf32(kBlockM, BLOCK_N, 0, 0)
                                                 : make_idesc_bf16(kBlockM, BLOCK_N, 0, 0);
            const uint64_t desc_hi = make_smem_desc(0, 16, 1024, 2);
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const int acc = Cfg::kAccBufs == 2 ? (it & 1) : 0;
                const uint32_t acc_phase = Cfg::kAccBufs == 2 ? ((it >> 1) & 1) : (it & 1);
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + acc * Cfg::kTileCols;
                for (int cb = 0; cb < p.cblks; ++cb) {
                    for (int dxi = 0; dxi < 3; ++dxi) {
                        mbar_wait(&a_full[sa], pa);
                        const uint32_t a_addr = smem_base + sa * Cfg::kAUnitBytes;
                        for (int dyi = 0; dyi < 3; ++dyi) {
                            mbar_wait(&b_full[sb], pb);
                            tc_fence_after();
                            const uint32_t b_addr = smem_base + b_ring_off + sb * Cfg::kBUnitBytes;
                            const uint64_t db0 = desc_hi | ((b_addr >> 4) & 0x3FFFu);
                            const uint32_t first = (cb | dxi | dyi) == 0 ? 0u : 1u;
                            if (leader) {
#pragma unroll
                            for (int j = 0; j < kMtPer; ++j) {
                                const int mt = mt_first + j;
                                const uint32_t a_tap = a_addr + (16 * mt + dyi) * 1024;
                                const uint64_t da0 = desc_hi | ((a_tap >> 4) & 0x3FFFu);
                                const uint32_t d = tmem_acc + mt * BLOCK_N;
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const uint64_t da = da0 + 2 * k, db = db0 + 2 * k;
                                    const uint32_t accum = k > 0 ? 1u : first;
                                    if (MODE == 0) {
                                        umma_tf32_ss(d, da, db, idesc, accum);
                                    } else {
                                        const uint64_t da_lo = da + (Cfg::kSlabBytes >> 4);
                                        const uint64_t db_lo = db + (Cfg::kBPlaneBytes >> 4);
                                        umma_f16_ss(d, da_lo, db, idesc, accum);   // Al * Bh
                                        umma_f16_ss(d, da, db_lo, idesc, 1u);      // Ah * Bl
                                        umma_f16_ss(d, da, db, idesc, 1u);         // Ah * Bh
                                    }
                                }
                            }
                            umma_commit(&b_empty[sb]);
                            }
                            if (++sb == SB) { sb = 0; pb ^= 1u; }
                        }
                        if (leader) umma_commit(&a_empty[sa]);
                        if (++sa == SA) { sa = 0; pa ^= 1u; }
                    }
                }
                if (leader) umma_commit(&tmem_full_bar[acc]);
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;          // row of a 128-pixel sub-tile: 8 wide x 16 tall
        const int wl = m & 7;
        const int hl = m >> 3;
        const float oscale = p.out_scale ? __ldg(p.out_scale) : 1.0f;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int acc = Cfg::kAccBufs == 2 ? (it & 1) : 0;
            const uint32_t acc_phase = Cfg::kAccBufs == 2 ? ((it >> 1) & 1) : (it & 1);
            const int n_tile = tile % n_tiles;
            int m_tile = tile / n_tiles;
            const int tw = m_tile % p.tiles_w;
            m_tile /= p.tiles_w;
            const int th = m_tile % p.tiles_h;
            const int n = m_tile / p.tiles_h;
            const float* brow = p.bias ? p.bias + n_tile * BLOCK_N : nullptr;
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int mt = 0; mt < MT; ++mt) {
                const int h = th * Cfg::kPH + mt * 16 + hl, w = tw * 8 + wl;
                const size_t pix = (static_cast<size_t>(n) * p.H + h) * p.W + w;
                float* yrow = p.y + pix * p.Cout + n_tile * BLOCK_N;
                const float* rrow = nullptr;
                if (p.residual_mode == 1) {
                    rrow = p.residual + pix * p.Cout + n_tile * BLOCK_N;
                } else if (p.residual_mode == 2) {
                    const size_t rp = (static_cast<size_t>(n) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
                    rrow = p.residual + rp * p.Cout + n_tile * BLOCK_N;
                }
                const uint32_t tmem_acc = tmem_base + acc * Cfg::kTileCols + mt * BLOCK_N +
                                          (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
                for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_acc + c0, v);
                    tmem_ld_wait();
                    if (mt == MT - 1 && c0 + 32 >= BLOCK_N) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
                    }
                    epilogue_store_chunk(p, v, oscale, yrow, nullptr, rrow, brow, pix * p.Cout + n_tile * BLOCK_N, c0);
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

// ====================================================================================================================
// CTA-pair variant of the halo kernel (tcgen05 cta_group::2).
//
// With the issue loop fixed, the halo kernel is bound by the SM's 128 B/clk of shared-memory bandwidth, which the MMA's
// operand reads (A 4 KB + B 32*BLOCK_N B per instruction) share with the TMA writes of the next operands
// (profiles/README.md: t_MMA ~ max(N/2, (A + B + writes) / 128 B) clk explains every row of the sweep).  A CTA pair
// executes one M = 256 MMA: each CTA supplies its own 128 pixels of A and only HALF of the weight tile, so per SM the
// weight bytes read AND written halve.  Pair = two horizontally adjacent 8-pixel-wide tiles (rank r takes columns
// w0 + 8r); protocol (CUTLASS' 2-SM scheme): both CTAs run producers into their own rings but count transaction bytes on
// the leader's full barriers; the leader's MMA warps issue, and their commits multicast to the empty / tmem_full barriers
// of both CTAs; epilogue warps of both CTAs arrive on the leader's tmem_empty barriers.
template <int BLOCK_N, int MT, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(HaloCfg<BLOCK_N, MT, MODE>::kThreads, 1)
conv_halo2_kernel(const __grid_constant__ ConvMaps tm, const ConvParams p) {
    using Cfg = HaloCfg<BLOCK_N, MT, MODE>;
    constexpr int kBHalfPlane = Cfg::kBPlaneBytes / 2;              // this CTA's half of the weight tile, one plane
    constexpr int kBHalfUnit = Cfg::kParts * kBHalfPlane;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t a_full[Cfg::kMaxAStages];
    __shared__ __align__(8) uint64_t a_empty[Cfg::kMaxAStages];
    __shared__ __align__(8) uint64_t b_full[Cfg::kMaxBStages];
    __shared__ __align__(8) uint64_t b_empty[Cfg::kMaxBStages];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_slot;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    const int SA = p.a_stages, SB = p.b_stages;
    const uint32_t b_ring_off = SA * Cfg::kAUnitBytes;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_tiles = p.Cout / BLOCK_N;
    const uint32_t rank = cluster_ctarank();            // 0 = leader
    const int pair0 = blockIdx.x >> 1, pair_step = gridDim.x >> 1;
    const int total_pairs = p.total_tiles;               // tiles of the pair grid: (W/16) x (H/PH) x N x n_tiles

    if (warp == 0 && lane == 0) {
#pragma unroll
        for (int q = 0; q < Cfg::kParts; ++q) {
            tma_prefetch_desc(&tm.a[q]);
            tma_prefetch_desc(&tm.b[q]);
        }
        for (int s = 0; s < SA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], Cfg::kIssuers); }
        for (int s = 0; s < SB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], Cfg::kIssuers); }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], Cfg::kIssuers);
            mbar_init(&tmem_empty_bar[a], 8);             // 4 epilogue warps of each CTA of the pair
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_2cta<Cfg::kTmemCols>(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                    // peer barriers initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ===================== activation-slab producer (both CTAs) =====================
        const bool leader = elect_one();
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = pair0; tile < total_pairs; tile += pair_step) {
            int m_tile = tile / n_tiles;
            const int tw = m_tile % p.tiles_w;
            m_tile /= p.tiles_w;
            const int th = m_tile % p.tiles_h;
            const int n = m_tile / p.tiles_h;
            const int w0 = tw * 16 + static_cast<int>(rank) * 8, h0 = th * Cfg::kPH;
            const int cbase = p.grouped ? (tile % n_tiles) * BLOCK_N : 0;
            for (int cb = 0; cb < p.cblks; ++cb) {
                for (int dx = -1; dx <= 1; ++dx) {
                    mbar_wait(&a_empty[stage], phase ^ 1u);
                    uint8_t* sa = smem_al + stage * Cfg::kAUnitBytes;
                    if (leader) {
                        if (rank == 0) mbar_expect_tx(&a_full[stage], 2 * Cfg::kAUnitBytes);   // both CTAs' slabs
#pragma unroll
                        for (int q = 0; q < Cfg::kParts; ++q)
                            tma_load_4d_2cta(sa + q * Cfg::kSlabBytes, &tm.a[q], &a_full[stage],
                                             cbase + cb * Cfg::kChanPerRow, w0 + dx, h0 - 1, n);
                    }
                    if (++stage == SA) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 6) {
        // ===================== weight-tile producer (both CTAs, half a tile each) =====================
        const bool leader = elect_one();
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = pair0; tile < total_pairs; tile += pair_step) {
            const int n_tile = tile % n_tiles;
            const int row0 = n_tile * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / 2);
            for (int cb = 0; cb < p.cblks; ++cb) {
                for (int dxi = 0; dxi < 3; ++dxi) {
                    for (int dyi = 0; dyi < 3; ++dyi) {
                        mbar_wait(&b_empty[stage], phase ^ 1u);
                        uint8_t* sb = smem_al + b_ring_off + stage * kBHalfUnit;
                        const int kcoord = (dyi * 3 + dxi) * p.kcin + cb * Cfg::kChanPerRow;
                        if (leader) {
                            if (rank == 0) mbar_expect_tx(&b_full[stage], 2 * kBHalfUnit);
#pragma unroll
                            for (int q = 0; q < Cfg::kParts; ++q)
                                tma_load_2d_2cta(sb + q * kBHalfPlane, &tm.b[q], &b_full[stage], kcoord, row0);
                        }
                        if (++stage == SB) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1 || warp == 7) {
        // ===================== MMA issuer(s): leader CTA only =====================
        if (rank == 0) {
            const bool leader = elect_one();
            constexpr int kMtPer = MT / Cfg::kIssuers;
            const int mt_first = warp == 1 ? 0 : kMtPer;
            constexpr uint32_t idesc = MODE == 0 ? make_idesc_tf32(256, BLOCK_N, 0, 0) : make_idesc_bf16(256, BLOCK_N, 0, 0);
            const uint64_t desc_hi = make_smem_desc(0, 16, 1024, 2);
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            int it = 0;
            for (int tile = pair0; tile < total_pairs; tile += pair_step, ++it) {
                const int acc = Cfg::kAccBufs == 2 ? (it & 1) : 0;
                const uint32_t acc_phase = Cfg::kAccBufs == 2 ? ((it >> 1) & 1) : (it & 1);
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + acc * Cfg::kTileCols;
                for (int cb = 0; cb < p.cblks; ++cb) {
                    for (int dxi = 0; dxi < 3; ++dxi) {
                        mbar_wait(&a_full[sa], pa);
                        const uint32_t a_addr = smem_base + sa * Cfg::kAUnitBytes;
                        for (int dyi = 0; dyi < 3; ++dyi) {
                            mbar_wait(&b_full[sb], pb);
                            tc_fence_after();
                            const uint32_t b_addr = smem_base + b_ring_off + sb * kBHalfUnit;
                            const uint64_t db0 = desc_hi | ((b_addr >> 4) & 0x3FFFu);
                            const uint32_t first = (cb | dxi | dyi) == 0 ? 0u : 1u;
                            if (leader) {
#pragma unroll
                                for (int j = 0; j < kMtPer; ++j) {
                                    const int mt = mt_first + j;
                                    const uint32_t a_tap = a_addr + (16 * mt + dyi) * 1024;
                                    const uint64_t da0 = desc_hi | ((a_tap >> 4) & 0x3FFFu);
                                    const uint32_t d = tmem_acc + mt * BLOCK_N;
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        const uint64_t da = da0 + 2 * k, db = db0 + 2 * k;
                                        const uint32_t accum = k > 0 ? 1u : first;
                                        if (MODE == 0) {
                                            umma_tf32_ss_2cta(d, da, db, idesc, accum);
                                        } else {
                                            const uint64_t da_lo = da + (Cfg::kSlabBytes >> 4);
                                            const uint64_t db_lo = db + (kBHalfPlane >> 4);
                                            umma_f16_ss_2cta(d, da_lo, db, idesc, accum);   // Al * Bh
                                            umma_f16_ss_2cta(d, da, db_lo, idesc, 1u);      // Ah * Bl
                                            umma_f16_ss_2cta(d, da, db, idesc, 1u);         // Ah * Bh
                                        }
                                    }
                                }
                                umma_commit_2cta(&b_empty[sb]);
                            }
                            if (++sb == SB) { sb = 0; pb ^= 1u; }
                        }
                        if (leader) umma_commit_2cta(&a_empty[sa]);
                        if (++sa == SA) { sa = 0; pa ^= 1u; }
                    }
                }
                if (leader) umma_commit_2cta(&tmem_full_bar[acc]);
            }
        }
    } else {
        // ===================== epilogue (warps 2..5 of both CTAs) =====================
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;
        const int wl = m & 7;
        const int hl = m >> 3;
        const float oscale = p.out_scale ? __ldg(p.out_scale) : 1.0f;
        int it = 0;
        for (int tile = pair0; tile < total_pairs; tile += pair_step, ++it) {
            const int acc = Cfg::kAccBufs == 2 ? (it & 1) : 0;
            const uint32_t acc_phase = Cfg::kAccBufs == 2 ? ((it >> 1) & 1) : (it & 1);
            const int n_tile = tile % n_tiles;
            int m_tile = tile / n_tiles;
            const int tw = m_tile % p.tiles_w;
            m_tile /= p.tiles_w;
            const int th = m_tile % p.tiles_h;
            const int n = m_tile / p.tiles_h;
            const float* brow = p.bias ? p.bias + n_tile * BLOCK_N : nullptr;
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int mt = 0; mt < MT; ++mt) {
                const int h = th * Cfg::kPH + mt * 16 + hl, w = tw * 16 + static_cast<int>(rank) * 8 + wl;
                const size_t pix = (static_cast<size_t>(n) * p.H + h) * p.W + w;
                float* yrow = p.y + pix * p.Cout + n_tile * BLOCK_N;
                const float* rrow = nullptr;
                if (p.residual_mode == 1) {
                    rrow = p.residual + pix * p.Cout + n_tile * BLOCK_N;
                } else if (p.residual_mode == 2) {
                    const size_t rp = (static_cast<size_t>(n) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
                    rrow = p.residual + rp * p.Cout + n_tile * BLOCK_N;
                }
                const uint32_t tmem_acc = tmem_base + acc * Cfg::kTileCols + mt * BLOCK_N +
                                          (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
                for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_acc + c0, v);
                    tmem_ld_wait();
                    if (mt == MT - 1 && c0 + 32 >= BLOCK_N) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[acc]);
                    }
                    epilogue_store_chunk(p, v, oscale, yrow, nullptr, rrow, brow, pix * p.Cout + n_tile * BLOCK_N, c0);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // the peer may still be reading this CTA's shared memory / arriving on its barriers
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2cta<Cfg::kTmemCols>(tmem_base);
    }
}

// Split-K second pass: y = epilogue( sum_s ws[s] ), the same epilogue as the main kernel, deterministic summation order.
__global__ void __launch_bounds__(256)
splitk_epilogue_kernel(const ConvParams p, long long total4) {
    const float oscale = p.out_scale ? __ldg(p.out_scale) : 1.0f;
    const int c4 = p.Cout >> 2;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total4; i += gridDim.x * 256LL) {
        const int c = static_cast<int>(i % c4) * 4;
        const long long pix = i / c4;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < p.splits; ++s) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p.ws + s * p.ws_stride + i * 4));
            o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
        }
        o.x *= oscale; o.y *= oscale; o.z *= oscale; o.w *= oscale;
        if (p.bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c));
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
        }
        if (p.residual_mode == 1) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(p.residual + i * 4));
            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        } else if (p.residual_mode == 2) {
            const int w = static_cast<int>(pix % p.W);
            const int h = static_cast<int>((pix / p.W) % p.H);
            const long long n = pix / (static_cast<long long>(p.W) * p.H);
            const long long rp = (n * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
            const float4 r = __ldg(reinterpret_cast<const float4*>(p.residual + rp * p.Cout + c));
            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        if (p.round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
        *reinterpret_cast<float4*>(p.y + i * 4) = o;
        if (p.y_split) {
            __nv_bfloat16* sp = p.y_split + i * 4;
            const __nv_bfloat16 h0 = __float2bfloat16_rn(o.x), h1 = __float2bfloat16_rn(o.y),
                                h2 = __float2bfloat16_rn(o.z), h3 = __float2bfloat16_rn(o.w);
            __nv_bfloat162 hi01 = __halves2bfloat162(h0, h1), hi23 = __halves2bfloat162(h2, h3);
            __nv_bfloat162 lo01 = __halves2bfloat162(__float2bfloat16_rn(o.x - __bfloat162float(h0)),
                                                     __float2bfloat16_rn(o.y - __bfloat162float(h1)));
            __nv_bfloat162 lo23 = __halves2bfloat162(__float2bfloat16_rn(o.z - __bfloat162float(h2)),
                                                     __float2bfloat16_rn(o.w - __bfloat162float(h3)));
            uint2 hv, lv;
            hv.x = *reinterpret_cast<uint32_t*>(&hi01); hv.y = *reinterpret_cast<uint32_t*>(&hi23);
            lv.x = *reinterpret_cast<uint32_t*>(&lo01); lv.y = *reinterpret_cast<uint32_t*>(&lo23);
            *reinterpret_cast<uint2*>(sp) = hv;
            *reinterpret_cast<uint2*>(sp + p.split_stride) = lv;
        }
    }
}

// K splits so that a layer with few (pixel patch x Cout) tiles still fills the machine: at most one wave of 2 x #SM CTAs,
// at least 4 K blocks per split, at most 32 splits.
static int choose_splits(int mn_tiles, int num_kb) {
    if (mn_tiles >= 100) return 1;
    int s = (2 * 148) / mn_tiles;
    if (s > num_kb / 4) s = num_kb / 4;
    if (s > 32) s = 32;
    return s < 1 ? 1 : s;
}

template <int BLOCK_N, int MODE>
static int launch_conv(const ConvMaps& tm, ConvParams p, int m_tiles, int stages, int ctas_per_sm, cudaStream_t stream) {
    using Cfg = ConvCfg<BLOCK_N, MODE>;
    static bool attr_set = false;
    if (!attr_set) {
        B200LP_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BLOCK_N, MODE>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kMaxSmemBytes));
        B200LP_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BLOCK_N, MODE>,
                                               cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_set = true;
    }
    // CTAs per SM: two when the double accumulators (2*BLOCK_N columns each) of two CTAs fit the 512 TMEM columns AND
    // half of the shared memory still holds a ring of >= 3 stages; otherwise one CTA with the deepest ring that fits.
    const int max_cpsm = (2 * Cfg::kTmemCols <= 512) ? 2 : 1;
    auto ring = [](int cpsm) {
        const int budget = (cpsm == 2 ? 113 * 1024 : Cfg::kMaxSmemBytes) - 1024;
        int st = budget / Cfg::kStageBytes;
        return st > Cfg::kMaxStages ? Cfg::kMaxStages : st;
    };
    if (ctas_per_sm <= 0) ctas_per_sm = (max_cpsm == 2 && ring(2) >= 3) ? 2 : 1;
    if (ctas_per_sm > max_cpsm) ctas_per_sm = max_cpsm;
    int max_stages = ring(ctas_per_sm);
    if (max_stages < 1) max_stages = 1;
    if (stages <= 0 || stages > max_stages) stages = max_stages;
    p.stages = stages;
    p.total_tiles = m_tiles * (p.Cout / BLOCK_N) * p.splits;
    const int smem_bytes = stages * Cfg::kStageBytes + 1024;   // + alignment slack
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        B200LP_CHECK_CUDA(cudaGetDevice(&dev));
        B200LP_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    int grid = num_sms * ctas_per_sm;
    if (grid > p.total_tiles) grid = p.total_tiles;
    conv_igemm_kernel<BLOCK_N, MODE><<<grid, kConvThreads, smem_bytes, stream>>>(tm, p);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    if (p.splits > 1) {
        const long long total4 = p.ws_stride / 4;
        long long blocks = (total4 + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        splitk_epilogue_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(p, total4);
        B200LP_CHECK_CUDA(cudaGetLastError());
        count_launch();
    }
    return B200LP_OK;
}

// ring depths of the halo kernel: the deepest slab ring (<= 4 slabs when they are small, else 3) that leaves room for
// >= 4 weight tiles; returns false when even 2 + 2 do not fit.
template <int BLOCK_N, int MT, int MODE>
static bool halo_pick_stages(int& a_stages, int& b_stages) {
    using Cfg = HaloCfg<BLOCK_N, MT, MODE>;
    const int budget = Cfg::kMaxSmemBytes - 1024;
    auto b_fit = [&](int a) {
        int b = (budget - a * Cfg::kAUnitBytes) / Cfg::kBUnitBytes;
        return b > Cfg::kMaxBStages ? Cfg::kMaxBStages : b;
    };
    if (a_stages <= 0) {
        a_stages = 2;
        for (int a = (Cfg::kAUnitBytes <= 20 * 1024 ? 4 : 3); a >= 2; --a)
            if (b_fit(a) >= 4) { a_stages = a; break; }   // a deep weight ring matters more than a third slab (halo_tune)
    }
    if (a_stages > Cfg::kMaxAStages) a_stages = Cfg::kMaxAStages;
    const int bmax = b_fit(a_stages);
    if (bmax < 2) return false;
    if (b_stages <= 0 || b_stages > bmax) b_stages = bmax;
    return true;
}

template <int BLOCK_N, int MT, int MODE>
static int launch_halo(const ConvMaps& tm, ConvParams p, int a_stages, int b_stages, cudaStream_t stream) {
    using Cfg = HaloCfg<BLOCK_N, MT, MODE>;
    static bool attr_set = false;
    if (!attr_set) {
        B200LP_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_kernel<BLOCK_N, MT, MODE>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kMaxSmemBytes));
        attr_set = true;
    }
    B200LP_REQUIRE((halo_pick_stages<BLOCK_N, MT, MODE>(a_stages, b_stages)),
                   "conv_fwd: halo rings (%d slabs, block_n %d, variant %d) do not fit shared memory", a_stages, BLOCK_N, MT);
    p.a_stages = a_stages;
    p.b_stages = b_stages;
    p.tiles_w = p.W / 8;
    p.tiles_h = p.H / Cfg::kPH;
    p.total_tiles = p.tiles_w * p.tiles_h * p.N * (p.Cout / BLOCK_N);
    const int smem_bytes = a_stages * Cfg::kAUnitBytes + b_stages * Cfg::kBUnitBytes + 1024;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        B200LP_CHECK_CUDA(cudaGetDevice(&dev));
        B200LP_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int grid = num_sms < p.total_tiles ? num_sms : p.total_tiles;
    conv_halo_kernel<BLOCK_N, MT, MODE><<<grid, Cfg::kThreads, smem_bytes, stream>>>(tm, p);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

template <int BLOCK_N, int MT, int MODE>
static int launch_halo2(const ConvMaps& tm, ConvParams p, int a_stages, int b_stages, cudaStream_t stream) {
    using Cfg = HaloCfg<BLOCK_N, MT, MODE>;
    static bool attr_set = false;
    if (!attr_set) {
        B200LP_CHECK_CUDA(cudaFuncSetAttribute(conv_halo2_kernel<BLOCK_N, MT, MODE>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kMaxSmemBytes));
        attr_set = true;
    }
    // ring depths: the weight ring holds half tiles, so the same budget buys twice the depth (capped at kMaxBStages)
    const int budget = Cfg::kMaxSmemBytes - 1024;
    const int b_half = Cfg::kBUnitBytes / 2;
    auto b_fit = [&](int a) {
        int b = (budget - a * Cfg::kAUnitBytes) / b_half;
        return b > Cfg::kMaxBStages ? Cfg::kMaxBStages : b;
    };
    if (a_stages <= 0) {
        a_stages = 2;
        for (int a = (Cfg::kAUnitBytes <= 20 * 1024 ? 4 : 3); a >= 2; --a)
            if (b_fit(a) >= 4) { a_stages = a; break; }
    }
    if (a_stages > Cfg::kMaxAStages) a_stages = Cfg::kMaxAStages;
    const int bmax = b_fit(a_stages);
    B200LP_REQUIRE(bmax >= 2, "conv_fwd: pair-kernel rings (%d slabs, block_n %d, variant %d) do not fit shared memory",
                   a_stages, BLOCK_N, MT);
    if (b_stages <= 0 || b_stages > bmax) b_stages = bmax;
    p.a_stages = a_stages;
    p.b_stages = b_stages;
    p.tiles_w = p.W / 16;                       // pair tiles: 16 pixels wide
    p.tiles_h = p.H / Cfg::kPH;
    p.total_tiles = p.tiles_w * p.tiles_h * p.N * (p.Cout / BLOCK_N);
    const int smem_bytes = a_stages * Cfg::kAUnitBytes + b_stages * b_half + 1024;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        B200LP_CHECK_CUDA(cudaGetDevice(&dev));
        B200LP_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    int pairs = num_sms / 2;
    if (pairs > p.total_tiles) pairs = p.total_tiles;
    conv_halo2_kernel<BLOCK_N, MT, MODE><<<2 * pairs, Cfg::kThreads, smem_bytes, stream>>>(tm, p);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

}  // namespace b200lp

using namespace b200lp;

// geometry shared by b200lp_conv_fwd and b200lp_conv_fwd_workspace
static int conv_pick_block_n(const b200lp_conv_args* a, int m_tiles) {
    int block_n = a->block_n;
    if (a->grouped) return a->precision == 0 ? 32 : 64;      // one 128-byte operand row of channels per N tile
    if (block_n == 0) {
        if (a->precision == 0 && a->Cout % 256 == 0 && static_cast<long>(m_tiles) * (a->Cout / 256) >= 100) block_n = 256;
        else if (a->Cout % 128 == 0) block_n = 128;
        else if (a->Cout % 64 == 0) block_n = 64;
        else block_n = 32;
    }
    return block_n;
}
static int conv_m_tiles(const b200lp_conv_args* a) {
    const int bw = a->W < 16 ? a->W : 16;
    const int bh = (kBlockM / bw) < a->H ? (kBlockM / bw) : a->H;
    const int bn = kBlockM / (bw * bh);
    return (a->W / bw) * (a->H / bh) * ((a->N + bn - 1) / bn);
}
static int conv_splits(const b200lp_conv_args* a, int block_n, int m_tiles) {
    if (a->splits == 1) return 1;
    const int chan_per_row = a->precision == 0 ? 32 : 64;
    const int kcin = a->grouped ? block_n : a->Cin;
    const int num_kb = a->ksize * a->ksize * ((kcin + chan_per_row - 1) / chan_per_row);
    int s = a->splits > 1 ? a->splits : choose_splits(m_tiles * (a->Cout / block_n), num_kb);
    if (s > num_kb) s = num_kb;
    const int per = (num_kb + s - 1) / s;
    return (num_kb + per - 1) / per;      // no empty split
}

// Halo-kernel sub-tiles per CTA (0 = use the per-tap kernel).  Auto: 3x3 layers on planes >= 16 x 8 that are not split
// along K.
static int conv_halo_mt(const b200lp_conv_args* a, int block_n, int splits, bool* pair) {
    *pair = false;
    if (a->variant < 0 || a->ksize != 3 || a->W < 8 || a->H < 16 || splits != 1) return 0;
    if (a->variant > 0) {
        const int mt = a->variant >= 10 ? a->variant - 10 : a->variant;     // 11 / 12: CTA-pair kernel, 1 / 2 sub-tiles
        if ((mt != 1 && mt != 2 && mt != 4) || a->H % (16 * mt)) return -1;
        if (a->variant >= 10 && (mt == 4 || a->W % 16 || block_n < 64)) return -1;
        *pair = a->variant >= 10;
        return mt;
    }
    // CTA pairs (cta_group::2) first: +4..9 % tf32, +6..19 % bf16x3 over the single-CTA halo kernel on every shape of
    // the sweep (profiles/r01_kernel_diag_pair.log).  Two sub-tiles for block_n <= 128 in tf32; one otherwise.
    if (a->W % 16 == 0 && block_n >= 64 && !a->grouped) {
        int mt = (a->precision == 0 && block_n <= 128) ? 2 : 1;
        const long pair_tiles1 = static_cast<long>(a->N) * (a->H / 16) * (a->W / 16) * (a->Cout / block_n);
        while (mt > 1 && (a->H % (16 * mt) || pair_tiles1 / mt < 60)) mt >>= 1;
        if (pair_tiles1 / mt >= 37) {
            *pair = true;
            return mt;
        }
    }
    // measured (halo_tune): two sub-tiles sharing every weight tile win whenever >= ~100 CTA tiles remain (even below one
    // tile per SM: 512->256 @64x64 runs 774 vs 742 TFLOP/s); four sub-tiles never beat two once both have two issuers.
    // bf16x3 with block_n 128: the 68-KB slab pairs leave only a 2-deep ring, one sub-tile is faster.
    // block_n 256 x 2 sub-tiles fills TMEM (no double buffering: the epilogue is exposed), which only pays with a long K
    // loop: 512->256 yes (774 vs 742), 256->256 no (the step's VGG / discriminator 256-channel layers got 16 % slower).
    int mt_max = (a->precision == 0 || block_n <= 64) ? 2 : 1;
    if (block_n == 256 && a->Cin < 512) mt_max = 1;
    const long tiles128 = static_cast<long>(a->N) * (a->H / 16) * (a->W / 8) * (a->Cout / block_n);
    int mt = mt_max;
    while (mt > 1 && (a->H % (16 * mt) || tiles128 / mt < 100)) mt >>= 1;
    return mt;
}

extern "C" int64_t b200lp_conv_fwd_workspace(const b200lp_conv_args* a) {
    if (!a || a->Cout <= 0 || a->N <= 0 || a->H <= 0 || a->W <= 0 || (a->ksize != 1 && a->ksize != 3)) return B200LP_EINVAL;
    const int m_tiles = conv_m_tiles(a);
    const int block_n = conv_pick_block_n(a, m_tiles);
    if (a->Cout % block_n) return B200LP_EINVAL;
    const int s = conv_splits(a, block_n, m_tiles);
    return s > 1 ? static_cast<int64_t>(s) * a->N * a->H * a->W * a->Cout * 4 : 0;
}

extern "C" int32_t b200lp_conv_fwd(const b200lp_conv_args* a, void* stream) {
    B200LP_REQUIRE(a && a->x && a->wp && a->y, "conv_fwd: null pointer");
    B200LP_REQUIRE(a->ksize == 1 || a->ksize == 3, "conv_fwd: ksize %d not in {1,3}", a->ksize);
    B200LP_REQUIRE(a->Cin % 32 == 0 && a->Cout % 32 == 0 && a->Cin > 0 && a->Cout > 0,
                   "conv_fwd: Cin=%d Cout=%d must be positive multiples of 32", a->Cin, a->Cout);
    B200LP_REQUIRE(ilog2_exact(a->H) >= 1 && ilog2_exact(a->W) >= 1 && a->N > 0,
                   "conv_fwd: H=%d W=%d must be powers of two >= 2, N=%d > 0", a->H, a->W, a->N);
    B200LP_REQUIRE(a->residual_mode >= 0 && a->residual_mode <= 2 && (a->residual_mode == 0 || a->residual),
                   "conv_fwd: bad residual mode %d", a->residual_mode);
    B200LP_REQUIRE(a->precision == 0 || a->precision == 1, "conv_fwd: precision %d not in {0 tf32, 1 bf16x3}",
                   a->precision);
    const int mode = a->precision;
    const int chan_per_row = mode == 0 ? 32 : 64;
    const int elem_bytes = mode == 0 ? 4 : 2;
    B200LP_REQUIRE(mode == 0 || a->Cin % 64 == 0 || a->Cin == 32,
                   "conv_fwd: bf16x3 needs Cin %% 64 == 0 (or Cin == 32), got %d", a->Cin);
    B200LP_REQUIRE(!a->grouped || (a->Cin == a->Cout && a->ksize == 3 && (a->block_n == 0 || a->block_n == chan_per_row)),
                   "conv_fwd: grouped mode needs Cin == Cout, ksize 3 and block_n auto (got Cin=%d Cout=%d k=%d block_n=%d)",
                   a->Cin, a->Cout, a->ksize, a->block_n);

    ConvParams p;
    p.out_scale = a->out_scale;
    p.bias = a->bias;
    p.residual = a->residual_mode ? a->residual : nullptr;
    p.y = a->y;
    p.y_split = static_cast<__nv_bfloat16*>(a->y_split);
    p.split_stride = static_cast<long long>(a->N) * a->H * a->W * a->Cout;
    p.N = a->N; p.H = a->H; p.W = a->W; p.Cin = a->Cin; p.Cout = a->Cout;
    p.ksize = a->ksize;
    p.bw = a->W < 16 ? a->W : 16;
    p.bh = (kBlockM / p.bw) < a->H ? (kBlockM / p.bw) : a->H;
    p.bn = kBlockM / (p.bw * p.bh);
    p.tiles_w = a->W / p.bw;
    p.tiles_h = a->H / p.bh;
    p.grouped = a->grouped ? 1 : 0;
    p.kcin = a->grouped ? chan_per_row : a->Cin;           // grouped: wp is [Cout][taps][block_n], block_n = one operand row
    p.cblks = (p.kcin + chan_per_row - 1) / chan_per_row;
    p.num_kb = a->ksize * a->ksize * p.cblks;
    p.residual_mode = a->residual_mode;
    p.relu = a->relu;
    p.round_out = a->round_tf32;
    const int bn_box = p.bn < a->N ? p.bn : a->N;
    p.a_bytes = static_cast<uint32_t>(kRowBytes * p.bw * p.bh * bn_box);
    const int tiles_n = (a->N + p.bn - 1) / p.bn;
    const int m_tiles = p.tiles_w * p.tiles_h * tiles_n;

    const int block_n = conv_pick_block_n(a, m_tiles);
    B200LP_REQUIRE((block_n == 32 || block_n == 64 || block_n == 128 || (block_n == 256 && mode == 0)) &&
                       a->Cout % block_n == 0,
                   "conv_fwd: block_n=%d incompatible with Cout=%d (precision %d)", block_n, a->Cout, mode);

    p.splits = conv_splits(a, block_n, m_tiles);
    p.ws = nullptr;
    p.ws_stride = static_cast<long long>(a->N) * a->H * a->W * a->Cout;
    if (p.splits > 1) {
        const int64_t need = static_cast<int64_t>(p.splits) * p.ws_stride * 4;
        if (a->workspace && a->workspace_bytes >= need) p.ws = a->workspace;
        else p.splits = 1;                      // no (or too small a) workspace: run unsplit
    }
    p.kb_per_split = (p.num_kb + p.splits - 1) / p.splits;

    bool halo_pair = false;
    const int halo_mt = conv_halo_mt(a, block_n, p.splits, &halo_pair);
    B200LP_REQUIRE(halo_mt >= 0, "conv_fwd: variant %d needs a 3x3 layer with W >= 8 (pair kernel: W %% 16 == 0), "
                   "H %% (16 * sub-tiles) == 0, no split-K", a->variant);

    ConvMaps tm;
    const uint64_t ktot = (uint64_t)a->ksize * a->ksize * p.kcin;
    const int parts = mode == 0 ? 1 : 2;
    for (int q = 0; q < parts; ++q) {
        {
            const uint64_t dims[4] = {(uint64_t)a->Cin, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->N};
            const uint64_t strides[3] = {(uint64_t)a->Cin * elem_bytes, (uint64_t)a->W * a->Cin * elem_bytes,
                                         (uint64_t)a->H * a->W * a->Cin * elem_bytes};
            // per-tap kernel: one 128-pixel patch per box; halo kernel: one 8 x (16*MT + 2) column-shifted slab
            const uint32_t box_tap[4] = {(uint32_t)chan_per_row, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)bn_box};
            const uint32_t box_halo[4] = {(uint32_t)chan_per_row, 8u, (uint32_t)(16 * halo_mt + 2), 1u};
            const uint32_t* box = halo_mt ? box_halo : box_tap;
            const char* base = static_cast<const char*>(a->x) +
                               (size_t)q * a->N * a->H * a->W * a->Cin * elem_bytes;   // lo plane follows hi plane
            int r = encode_tmap(&tm.a[q], base, mode == 0 ? kTmapF32 : kTmapBF16, 4, dims, strides, box, false);
            if (r) return r;
        }
        {
            const uint64_t dims[2] = {ktot, (uint64_t)a->Cout};
            const uint64_t strides[1] = {ktot * elem_bytes};
            // CTA-pair kernel: each CTA of the pair fetches half of the weight tile
            const uint32_t box[2] = {(uint32_t)chan_per_row, (uint32_t)(halo_pair ? block_n / 2 : block_n)};
            const char* base = static_cast<const char*>(a->wp) + (size_t)q * ktot * a->Cout * elem_bytes;
            int r = encode_tmap(&tm.b[q], base, mode == 0 ? kTmapF32 : kTmapBF16, 2, dims, strides, box, false);
            if (r) return r;
        }
    }
    if (mode == 0) { tm.a[1] = tm.a[0]; tm.b[1] = tm.b[0]; }
    cudaStream_t s = as_stream(stream);
    const int st = a->stages, cps = a->ctas_per_sm;
    if (halo_pair) {
        const int ast = a->a_stages;
#define B200LP_PAIR_CASE(BN, MT, MODE) \
    if (block_n == BN && halo_mt == MT && mode == MODE) return launch_halo2<BN, MT, MODE>(tm, p, ast, st, s);
        B200LP_PAIR_CASE(256, 1, 0) B200LP_PAIR_CASE(256, 2, 0)
        B200LP_PAIR_CASE(128, 1, 0) B200LP_PAIR_CASE(128, 2, 0)
        B200LP_PAIR_CASE(64, 1, 0) B200LP_PAIR_CASE(64, 2, 0)
        B200LP_PAIR_CASE(128, 1, 1) B200LP_PAIR_CASE(128, 2, 1)
        B200LP_PAIR_CASE(64, 1, 1) B200LP_PAIR_CASE(64, 2, 1)
#undef B200LP_PAIR_CASE
        B200LP_REQUIRE(false, "conv_fwd: no pair kernel for block_n %d, variant %d, precision %d", block_n, a->variant, mode);
    }
    if (halo_mt) {
        const int ast = a->a_stages;
#define B200LP_HALO_CASE(BN, MT, MODE) \
    if (block_n == BN && halo_mt == MT && mode == MODE) return launch_halo<BN, MT, MODE>(tm, p, ast, st, s);
        B200LP_HALO_CASE(256, 1, 0) B200LP_HALO_CASE(256, 2, 0)
        B200LP_HALO_CASE(128, 1, 0) B200LP_HALO_CASE(128, 2, 0)
        B200LP_HALO_CASE(64, 1, 0) B200LP_HALO_CASE(64, 2, 0) B200LP_HALO_CASE(64, 4, 0)
        B200LP_HALO_CASE(32, 1, 0) B200LP_HALO_CASE(32, 2, 0) B200LP_HALO_CASE(32, 4, 0)
        B200LP_HALO_CASE(128, 1, 1) B200LP_HALO_CASE(128, 2, 1)
        B200LP_HALO_CASE(64, 1, 1) B200LP_HALO_CASE(64, 2, 1)
        B200LP_HALO_CASE(32, 1, 1) B200LP_HALO_CASE(32, 2, 1)
#undef B200LP_HALO_CASE
        B200LP_REQUIRE(false, "conv_fwd: no halo kernel for block_n %d, variant %d, precision %d", block_n, halo_mt, mode);
    }
    if (mode == 0) {
        switch (block_n) {
            case 256: return launch_conv<256, 0>(tm, p, m_tiles, st, cps, s);
            case 128: return launch_conv<128, 0>(tm, p, m_tiles, st, cps, s);
            case 64: return launch_conv<64, 0>(tm, p, m_tiles, st, cps, s);
            default: return launch_conv<32, 0>(tm, p, m_tiles, st, cps, s);
        }
    }
    switch (block_n) {
        case 128: return launch_conv<128, 1>(tm, p, m_tiles, st, cps, s);
        case 64: return launch_conv<64, 1>(tm, p, m_tiles, st, cps, s);
        default: return launch_conv<32, 1>(tm, p, m_tiles, st, cps, s);
    }
}

// ------------------------------------------------------------------------------------------------ weight packing
namespace b200lp {
// One thread per packed element, indexed by DESTINATION (coalesced writes); the strided reads of the OIHW source are
// served by L2 (a weight tensor is at most 9.4 MB).  A shared-memory-transposing variant was measured slower on the
// small tensors (too few blocks) — profiles/r01_bench_graph_first.json.
template <bool SPLIT>
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                        float* __restrict__ wp, int Cout, int Cin, int taps, int transpose) {
    const float s = scale ? __ldg(scale) : 1.0f;
    const long total = static_cast<long>(Cout) * Cin * taps;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        float v;
        if (!transpose) {
            const int ci = i % Cin;
            const int tap = (i / Cin) % taps;
            const int co = i / (static_cast<long>(Cin) * taps);
            v = w[(static_cast<long>(co) * Cin + ci) * taps + tap];
        } else {
            const int co = i % Cout;
            const int tapf = (i / Cout) % taps;
            const int ci = i / (static_cast<long>(Cout) * taps);
            v = w[(static_cast<long>(co) * Cin + ci) * taps + (taps - 1 - tapf)];
        }
        const float f = v * s;
        if (!SPLIT) {
            wp[i] = round_tf32(f);
        } else {   // (hi, lo) bf16 planes: hi + lo == f to 2^-17
            __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(wp);
            const __nv_bfloat16 h = __float2bfloat16_rn(f);
            out[i] = h;
            out[total + i] = __float2bfloat16_rn(f - __bfloat162float(h));
        }
    }
}

// Multi-tensor variant: every conv weight of a network re-packed by ONE launch after the optimizer step (the step has
// ~100 packed copies: forward + transposed, tf32 + bf16 planes; one 16-us launch each was 5 % of the step).
struct PackItem {
    const float* w;     // OIHW source
    void* wp;           // packed destination
    long long Cout, Cin, taps, transpose, precision, total;
};

__device__ __forceinline__ float pack_gather(const float* __restrict__ w, long i, int Cout, int Cin, int taps, int transpose) {
    if (!transpose) {
        const int ci = i % Cin;
        const int tap = (i / Cin) % taps;
        const int co = i / (static_cast<long>(Cin) * taps);
        return w[(static_cast<long>(co) * Cin + ci) * taps + tap];
    }
    const int co = i % Cout;
    const int tapf = (i / Cout) % taps;
    const int ci = i / (static_cast<long>(Cout) * taps);
    return w[(static_cast<long>(co) * Cin + ci) * taps + (taps - 1 - tapf)];
}

__global__ void __launch_bounds__(256)
pack_conv_weight_multi_kernel(const PackItem* __restrict__ table, const int* __restrict__ chunk_item,
                              const long long* __restrict__ chunk_off, long long chunk_elems) {
    const PackItem it = table[chunk_item[blockIdx.x]];
    const long off = chunk_off[blockIdx.x];
    long end = off + chunk_elems;
    if (end > it.total) end = it.total;
    const int Cout = static_cast<int>(it.Cout), Cin = static_cast<int>(it.Cin), taps = static_cast<int>(it.taps);
    for (long i = off + threadIdx.x; i < end; i += 256) {
        const float f = pack_gather(it.w, i, Cout, Cin, taps, static_cast<int>(it.transpose));
        if (it.precision == 0) {
            static_cast<float*>(it.wp)[i] = round_tf32(f);
        } else {
            __nv_bfloat16* out = static_cast<__nv_bfloat16*>(it.wp);
            const __nv_bfloat16 h = __float2bfloat16_rn(f);
            out[i] = h;
            out[it.total + i] = __float2bfloat16_rn(f - __bfloat162float(h));
        }
    }
}

// Tiled multi-tensor packing: one block per 32 (co) x 32 (ci) x taps tile.  The tile's source rows (one co: 32*taps
// contiguous floats of the OIHW tensor) are read coalesced into shared memory, then written as 128-byte runs along the
// packed layout's contiguous axis (ci forward, co transposed) — the element-wise kernel above reads with a stride of
// `taps` (forward) or Cin*taps (transposed) floats: 8-9x the sectors.
__global__ void __launch_bounds__(256)
pack_conv_weight_tiles_kernel(const PackItem* __restrict__ table, const int* __restrict__ tile_item,
                              const int* __restrict__ tile_index) {
    __shared__ float tile[32 * 289];
    const PackItem it = table[tile_item[blockIdx.x]];
    const int Cout = static_cast<int>(it.Cout), Cin = static_cast<int>(it.Cin), taps = static_cast<int>(it.taps);
    const int ci_tiles = (Cin + 31) >> 5;
    const int t = tile_index[blockIdx.x];
    const int co0 = (t / ci_tiles) << 5, ci0 = (t % ci_tiles) << 5;
    const int nco = min(32, Cout - co0), nci = min(32, Cin - ci0);
    const int run = nci * taps;                                    // contiguous floats per source row
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < nco; r += 8) {                          // a warp per source row, lanes along the contiguous run
        const float* src = it.w + (static_cast<long>(co0 + r) * Cin + ci0) * taps;
        float v[9];                                                // all loads of the row in flight before the first store
#pragma unroll
        for (int k = 0; k < 9; ++k) v[k] = lane + 32 * k < run ? __ldg(src + lane + 32 * k) : 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k)
            if (lane + 32 * k < run) tile[r * 289 + lane + 32 * k] = v[k];
    }
    __syncthreads();
    const bool split = it.precision != 0;
    float* outf = static_cast<float*>(it.wp);
    __nv_bfloat16* outh = static_cast<__nv_bfloat16*>(it.wp);
    // a warp per 128-byte destination run: (co, tap) rows with lanes along ci (forward), (ci, tap) rows with lanes along
    // co (transposed); no per-element division
    const int nrows = (it.transpose ? nci : nco), nlan = (it.transpose ? nco : nci);
    for (int rr = warp; rr < nrows; rr += 8) {
        for (int tap = 0; tap < taps; ++tap) {
            if (lane >= nlan) continue;
            float f;
            long dst;
            if (!it.transpose) {         // rr = co, lane = ci: dest [co][tap][ci]
                f = tile[rr * 289 + lane * taps + tap];
                dst = (static_cast<long>(co0 + rr) * taps + tap) * Cin + ci0 + lane;
            } else {                     // rr = ci, lane = co: dest [ci][taps-1-tap][co]
                f = tile[lane * 289 + rr * taps + tap];
                dst = (static_cast<long>(ci0 + rr) * taps + (taps - 1 - tap)) * Cout + co0 + lane;
            }
            if (!split) {
                outf[dst] = round_tf32(f);
            } else {
                const __nv_bfloat16 h = __float2bfloat16_rn(f);
                outh[dst] = h;
                outh[it.total + dst] = __float2bfloat16_rn(f - __bfloat162float(h));
            }
        }
    }
}
}  // namespace b200lp

extern "C" int32_t b200lp_pack_conv_weight(const float* w_oihw, const float* scale, void* wp_out, int32_t Cout,
                                           int32_t Cin, int32_t ksize, int32_t transpose, int32_t precision,
                                           void* stream) {
    float* wp = static_cast<float*>(wp_out);
    B200LP_REQUIRE(w_oihw && wp && Cout > 0 && Cin > 0 && (ksize == 1 || ksize == 3), "pack_conv_weight: bad args");
    const long total = static_cast<long>(Cout) * Cin * ksize * ksize;
    const int threads = 256;
    long blocks = (total + threads - 1) / threads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (precision == 0)
        pack_conv_weight_kernel<false><<<(int)blocks, threads, 0, as_stream(stream)>>>(w_oihw, scale, wp, Cout, Cin,
                                                                                     ksize * ksize, transpose);
    else
        pack_conv_weight_kernel<true><<<(int)blocks, threads, 0, as_stream(stream)>>>(w_oihw, scale, wp, Cout, Cin,
                                                                                    ksize * ksize, transpose);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

namespace b200lp {
// Grouped 3x3 weight [C][cpg][9] -> block-diagonal dense tiles [C][9][B] (B = 32 tf32 / 64 bf16 planes); one thread per
// packed element (coalesced writes; the source is at most 1.2 MB and stays in L2).
template <bool SPLIT>
__global__ void pack_gconv_weight_kernel(const float* __restrict__ w, float* __restrict__ wp, int C, int cpg, int B,
                                         int transpose) {
    const long total = static_cast<long>(C) * 9 * B;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int j = static_cast<int>(i % B);
        const int tap = static_cast<int>((i / B) % 9);
        const int r = static_cast<int>(i / (9L * B));           // output row: co (forward) or ci (data-gradient)
        const int other = (r / B) * B + j;                       // the block's j-th channel on the contracted side
        float f = 0.f;
        if (other / cpg == r / cpg) {
            const int co = transpose ? other : r, ci = transpose ? r : other;
            f = w[(static_cast<long>(co) * cpg + (ci % cpg)) * 9 + (transpose ? 8 - tap : tap)];
        }
        if (!SPLIT) {
            wp[i] = round_tf32(f);
        } else {
            __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(wp);
            const __nv_bfloat16 h = __float2bfloat16_rn(f);
            out[i] = h;
            out[total + i] = __float2bfloat16_rn(f - __bfloat162float(h));
        }
    }
}
}  // namespace b200lp

extern "C" int32_t b200lp_pack_conv_weight_tiles(const void* table_dev, const int32_t* tile_item_dev,
                                                 const int32_t* tile_index_dev, int32_t n_tiles, void* stream) {
    B200LP_REQUIRE(table_dev && tile_item_dev && tile_index_dev && n_tiles > 0, "pack_conv_weight_tiles: bad args");
    pack_conv_weight_tiles_kernel<<<n_tiles, 256, 0, as_stream(stream)>>>(static_cast<const PackItem*>(table_dev),
                                                                         tile_item_dev, tile_index_dev);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_pack_gconv_weight(const float* w, void* wp_out, int32_t C, int32_t cpg, int32_t transpose,
                                            int32_t precision, void* stream) {
    const int B = precision == 0 ? 32 : 64;
    B200LP_REQUIRE(w && wp_out && C > 0 && cpg > 0 && C % B == 0 && B % cpg == 0 && (precision == 0 || precision == 1),
                   "pack_gconv_weight: C=%d cpg=%d precision=%d (C %% %d == 0 and cpg | %d required)", C, cpg, precision, B, B);
    const long total = static_cast<long>(C) * 9 * B;
    long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    float* wp = static_cast<float*>(wp_out);
    if (precision == 0)
        pack_gconv_weight_kernel<false><<<(int)blocks, 256, 0, as_stream(stream)>>>(w, wp, C, cpg, B, transpose);
    else
        pack_gconv_weight_kernel<true><<<(int)blocks, 256, 0, as_stream(stream)>>>(w, wp, C, cpg, B, transpose);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_pack_conv_weight_multi(const void* table_dev, const int32_t* chunk_item_dev,
                                                 const int64_t* chunk_off_dev, int32_t n_chunks, int64_t chunk_elems,
                                                 void* stream) {
    B200LP_REQUIRE(table_dev && chunk_item_dev && chunk_off_dev && n_chunks > 0 && chunk_elems > 0,
                   "pack_conv_weight_multi: bad args");
    pack_conv_weight_multi_kernel<<<n_chunks, 256, 0, as_stream(stream)>>>(
        static_cast<const PackItem*>(table_dev), chunk_item_dev, reinterpret_cast<const long long*>(chunk_off_dev),
        chunk_elems);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}
