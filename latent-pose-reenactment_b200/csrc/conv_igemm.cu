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
// Replaces torch's nn.Conv2d forward (and, with transposed packing, conv backward-data) at the call sites listed
// in include/b200lp.h.
#include "common.cuh"
#include "ptx.cuh"

namespace b200lp {

constexpr int kBlockM = 128;      // pixels per CTA tile
constexpr int kBlockK = 32;       // tf32 elements per pipeline stage row (= 128 bytes = swizzle span)
constexpr int kUmmaK = 8;         // K of one tcgen05.mma.kind::tf32
constexpr int kConvThreads = 192;

struct ConvParams {
    const float* bias;
    const float* residual;
    float* y;
    int N, H, W, Cin, Cout;
    int ksize;
    int bw, bh, bn;        // patch shape: bw*bh*bn == 128
    int tiles_w, tiles_h;  // patches per image row / column
    int cblks;             // Cin / 32
    int num_kb;            // taps * cblks
    int residual_mode;
    int relu;
    int round_out;
    uint32_t a_bytes;      // bytes one A box delivers (may be < 16 KB when bn > N)
};

template <int BLOCK_N>
struct ConvCfg {
    static constexpr int kABytes = kBlockM * kBlockK * 4;   // 16 KB
    static constexpr int kBBytes = BLOCK_N * kBlockK * 4;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024;  // + alignment slack
    static constexpr uint32_t kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const ConvParams p) {
    using Cfg = ConvCfg<BLOCK_N>;
    constexpr int kStages = Cfg::kStages;

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kStages];
    __shared__ __align__(8) uint64_t empty_bar[kStages];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_slot;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    // stage s: A at s*kStageBytes, B right after it (both 1024-byte aligned)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int n_tile = blockIdx.x;
    int m_tile = blockIdx.y;
    const int tw = m_tile % p.tiles_w;
    m_tile /= p.tiles_w;
    const int th = m_tile % p.tiles_h;
    const int tn = m_tile / p.tiles_h;
    const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full_bar, 1);
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
        if (lane == 0) {
            const int pad = p.ksize >> 1;
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < p.num_kb; ++kb) {
                const int tap = kb / p.cblks;
                const int cb = kb - tap * p.cblks;
                const int dy = tap / p.ksize - pad;
                const int dx = tap % p.ksize - pad;
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                uint8_t* sa = smem_al + stage * Cfg::kStageBytes;
                uint8_t* sb = sa + Cfg::kABytes;
                mbar_expect_tx(&full_bar[stage], p.a_bytes + Cfg::kBBytes);
                tma_load_4d(sa, &tmA, &full_bar[stage], cb * kBlockK, w0 + dx, h0 + dy, n0);
                tma_load_2d(sb, &tmB, &full_bar[stage], kb * kBlockK, n_tile * BLOCK_N);
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(kBlockM, BLOCK_N, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < p.num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_base + stage * Cfg::kStageBytes;
                const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                    // K-major SWIZZLE_128B: 8-row groups are 1024 B apart (SBO); stepping K by 8 tf32 = +32 bytes
                    const uint64_t da = make_smem_desc(a_addr + k * (kUmmaK * 4), 16, 1024, 2);
                    const uint64_t db = make_smem_desc(b_addr + k * (kUmmaK * 4), 16, 1024, 2);
                    umma_tf32_ss(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above have read it
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            umma_commit(&tmem_full_bar);  // accumulator complete
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quarter = warp & 3;               // TMEM lane quarter this warp may access
        const int m = quarter * 32 + lane;          // row of the tile = pixel of the patch
        const int wl = m % p.bw;
        const int hl = (m / p.bw) % p.bh;
        const int nl = m / (p.bw * p.bh);
        const int n = n0 + nl, h = h0 + hl, w = w0 + wl;
        const bool valid = n < p.N;
        const size_t pix = (static_cast<size_t>(n) * p.H + h) * p.W + w;
        float* yrow = p.y + pix * p.Cout + n_tile * BLOCK_N;
        const float* rrow = nullptr;
        if (p.residual_mode == 1) {
            rrow = p.residual + pix * p.Cout + n_tile * BLOCK_N;
        } else if (p.residual_mode == 2) {
            const size_t rp = (static_cast<size_t>(n) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
            rrow = p.residual + rp * p.Cout + n_tile * BLOCK_N;
        }
        const float* brow = p.bias ? p.bias + n_tile * BLOCK_N : nullptr;

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
                    float4 o;
                    o.x = __uint_as_float(v[j + 0]);
                    o.y = __uint_as_float(v[j + 1]);
                    o.z = __uint_as_float(v[j + 2]);
                    o.w = __uint_as_float(v[j + 3]);
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

template <int BLOCK_N>
static int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, int m_tiles,
                       cudaStream_t stream) {
    using Cfg = ConvCfg<BLOCK_N>;
    static bool attr_set = false;
    if (!attr_set) {
        B200LP_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_tf32_kernel<BLOCK_N>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        attr_set = true;
    }
    dim3 grid(p.Cout / BLOCK_N, m_tiles, 1);
    conv_igemm_tf32_kernel<BLOCK_N><<<grid, kConvThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, p);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

}  // namespace b200lp

using namespace b200lp;

extern "C" int32_t b200lp_conv_fwd(const b200lp_conv_args* a, void* stream) {
    B200LP_REQUIRE(a && a->x && a->wp && a->y, "conv_fwd: null pointer");
    B200LP_REQUIRE(a->ksize == 1 || a->ksize == 3, "conv_fwd: ksize %d not in {1,3}", a->ksize);
    B200LP_REQUIRE(a->Cin % 32 == 0 && a->Cout % 32 == 0 && a->Cin > 0 && a->Cout > 0,
                   "conv_fwd: Cin=%d Cout=%d must be positive multiples of 32", a->Cin, a->Cout);
    B200LP_REQUIRE(ilog2_exact(a->H) >= 1 && ilog2_exact(a->W) >= 1 && a->N > 0,
                   "conv_fwd: H=%d W=%d must be powers of two >= 2, N=%d > 0", a->H, a->W, a->N);
    B200LP_REQUIRE(a->residual_mode >= 0 && a->residual_mode <= 2 && (a->residual_mode == 0 || a->residual),
                   "conv_fwd: bad residual mode %d", a->residual_mode);

    ConvParams p;
    p.bias = a->bias;
    p.residual = a->residual_mode ? a->residual : nullptr;
    p.y = a->y;
    p.N = a->N; p.H = a->H; p.W = a->W; p.Cin = a->Cin; p.Cout = a->Cout;
    p.ksize = a->ksize;
    p.bw = a->W < 16 ? a->W : 16;
    p.bh = (kBlockM / p.bw) < a->H ? (kBlockM / p.bw) : a->H;
    p.bn = kBlockM / (p.bw * p.bh);
    p.tiles_w = a->W / p.bw;
    p.tiles_h = a->H / p.bh;
    p.cblks = a->Cin / kBlockK;
    p.num_kb = a->ksize * a->ksize * p.cblks;
    p.residual_mode = a->residual_mode;
    p.relu = a->relu;
    p.round_out = a->round_tf32;
    const int bn_box = p.bn < a->N ? p.bn : a->N;
    p.a_bytes = static_cast<uint32_t>(kBlockK * 4 * p.bw * p.bh * bn_box);
    const int tiles_n = (a->N + p.bn - 1) / p.bn;
    const int m_tiles = p.tiles_w * p.tiles_h * tiles_n;

    int block_n = a->block_n;
    if (block_n == 0) {
        if (a->Cout % 256 == 0 && static_cast<long>(m_tiles) * (a->Cout / 256) >= 148) block_n = 256;
        else if (a->Cout % 128 == 0) block_n = 128;
        else if (a->Cout % 64 == 0) block_n = 64;
        else block_n = 32;
    }
    B200LP_REQUIRE((block_n == 32 || block_n == 64 || block_n == 128 || block_n == 256) && a->Cout % block_n == 0,
                   "conv_fwd: block_n=%d incompatible with Cout=%d", block_n, a->Cout);

    CUtensorMap tmA, tmB;
    {
        const uint64_t dims[4] = {(uint64_t)a->Cin, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->N};
        const uint64_t strides[3] = {(uint64_t)a->Cin * 4, (uint64_t)a->W * a->Cin * 4,
                                     (uint64_t)a->H * a->W * a->Cin * 4};
        const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)bn_box};
        int r = encode_tmap_f32(&tmA, a->x, 4, dims, strides, box);
        if (r) return r;
    }
    {
        const uint64_t ktot = (uint64_t)a->ksize * a->ksize * a->Cin;
        const uint64_t dims[2] = {ktot, (uint64_t)a->Cout};
        const uint64_t strides[1] = {ktot * 4};
        const uint32_t box[2] = {(uint32_t)kBlockK, (uint32_t)block_n};
        int r = encode_tmap_f32(&tmB, a->wp, 2, dims, strides, box);
        if (r) return r;
    }
    cudaStream_t s = as_stream(stream);
    switch (block_n) {
        case 256: return launch_conv<256>(tmA, tmB, p, m_tiles, s);
        case 128: return launch_conv<128>(tmA, tmB, p, m_tiles, s);
        case 64: return launch_conv<64>(tmA, tmB, p, m_tiles, s);
        default: return launch_conv<32>(tmA, tmB, p, m_tiles, s);
    }
}

// ------------------------------------------------------------------------------------------------ weight packing
namespace b200lp {
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                        float* __restrict__ wp, int Cout, int Cin, int taps, int transpose) {
    const float s = scale ? __ldg(scale) : 1.0f;
    const long total = static_cast<long>(Cout) * Cin * taps;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        // i indexes the packed destination so that writes are coalesced
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
        wp[i] = round_tf32(v * s);
    }
}
}  // namespace b200lp

extern "C" int32_t b200lp_pack_conv_weight(const float* w_oihw, const float* scale, float* wp, int32_t Cout,
                                           int32_t Cin, int32_t ksize, int32_t transpose, void* stream) {
    B200LP_REQUIRE(w_oihw && wp && Cout > 0 && Cin > 0 && (ksize == 1 || ksize == 3), "pack_conv_weight: bad args");
    const long total = static_cast<long>(Cout) * Cin * ksize * ksize;
    const int threads = 256;
    long blocks = (total + threads - 1) / threads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    pack_conv_weight_kernel<<<(int)blocks, threads, 0, as_stream(stream)>>>(w_oihw, scale, wp, Cout, Cin,
                                                                            ksize * ksize, transpose);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}
