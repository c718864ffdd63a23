// Identity encoder (torchvision ResNeXt50-32x4d, train-mode BatchNorm over the B*K identity frames) — the kernels that
// are not tensor-core GEMMs, forward AND backward.  Shared with the pose encoder's backward (MobileNetV2).
//
// Replaces, for `embedders/unsupervised_pose_separate_embResNeXt_segmentation.py:26-27,37-54`
// (Embedder.get_identity_embedding) and its autograd backward, cuDNN's BatchNorm forward / backward kernels
// (bn_fw_tr / bn_bw: 18 ms of the 64.6 ms meta-training step in round 1's profile), the grouped 3x3 convolutions
// (g = 32, 4..32 channels per group: 100-300 us per cuDNN launch), NCHW<->NHWC conversions, ReLU / add / max-pool
// kernels.  The 1x1 convolutions (94 % of the encoder's FLOPs) and the 7x7 stem (through an im2col patch matrix) run on
// the tcgen05 implicit-GEMM kernels of conv_igemm.cu / conv_wgrad.cu.
//
// Data flow (activations NHWC fp32, a layer = [M = N*H*W rows][C channels]):
//   conv (raw output r) -> col_stats(r) -> bn_finalize -> (mean, rstd, scale, shift)
//   consumers apply BatchNorm + ReLU on load (grouped conv) or through ONE materialising pass (bn_act: writes the
//   tf32-rounded fp32 operand for the weight-gradient GEMM and / or the (hi, lo) bf16 planes for the bf16x3 forward GEMM)
//   backward: bn_bwd_reduce (sum dz, sum dz*xhat per channel; the ReLU mask is recomputed, never stored)
//             -> bn_bwd_finalize (dgamma, dbeta added in place; the two correction coefficients) -> bn_bwd_apply.
// Reductions are two-stage with a fixed order (no atomics): results are bit-reproducible.
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace b200lp {

namespace {

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return fminf(fmaxf(v, 0.f), 6.f);
    return v;
}

// dz = dy * [activation passes]; mode 0: none, 1: mask tensor value > 0, 2: relu(pre) with pre = raw*scale+shift,
// 3: relu6(pre): 0 < pre < 6 (torch's hardtanh backward: gradient where min < x < max)
// mode 4: one byte per float4 written by bn_act (bit k = activation k passed) -> +1 / -1 so that `> 0` selects
__device__ __forceinline__ float4 mask_byte4(unsigned b) {
    return make_float4((b & 1u) ? 1.f : -1.f, (b & 2u) ? 1.f : -1.f, (b & 4u) ? 1.f : -1.f, (b & 8u) ? 1.f : -1.f);
}
__device__ __forceinline__ float mask_apply(float dy, float pre_or_mask, int mode) {
    if (mode == 0) return dy;
    if (mode == 3) return (pre_or_mask > 0.f && pre_or_mask < 6.f) ? dy : 0.f;
    return pre_or_mask > 0.f ? dy : 0.f;
}

// one float4 as (hi, lo) bf16 planes; lo plane at +split_stride elements
__device__ __forceinline__ void store_split4(__nv_bfloat16* ys, long long split_stride, size_t off, float4 o) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(o.x), h1 = __float2bfloat16_rn(o.y), h2 = __float2bfloat16_rn(o.z),
                        h3 = __float2bfloat16_rn(o.w);
    __nv_bfloat162 hi01 = __halves2bfloat162(h0, h1), hi23 = __halves2bfloat162(h2, h3);
    __nv_bfloat162 lo01 = __halves2bfloat162(__float2bfloat16_rn(o.x - __bfloat162float(h0)),
                                             __float2bfloat16_rn(o.y - __bfloat162float(h1)));
    __nv_bfloat162 lo23 = __halves2bfloat162(__float2bfloat16_rn(o.z - __bfloat162float(h2)),
                                             __float2bfloat16_rn(o.w - __bfloat162float(h3)));
    uint2 hv, lv;
    hv.x = *reinterpret_cast<uint32_t*>(&hi01); hv.y = *reinterpret_cast<uint32_t*>(&hi23);
    lv.x = *reinterpret_cast<uint32_t*>(&lo01); lv.y = *reinterpret_cast<uint32_t*>(&lo23);
    *reinterpret_cast<uint2*>(ys + off) = hv;
    *reinterpret_cast<uint2*>(ys + split_stride + off) = lv;
}

__device__ __forceinline__ float4 round4(float4 v) {
    return make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
}

}  // namespace

// row chunks of an [M][C] matrix for the two-stage per-channel reductions: <= 4 x 148 chunks of >= 16 rows
static int row_chunks(long M, int* rows_per_chunk) {
    long rpc = (M + 591) / 592;
    if (rpc < 16) rpc = 16;
    *rows_per_chunk = static_cast<int>(rpc);
    return static_cast<int>((M + rpc - 1) / rpc);
}

// ------------------------------------------------------------------------------------------------ column statistics
// part[chunk][0][c] = sum over the chunk's rows of x[r][c]; part[chunk][1][c] = sum of squares.
// block 256 = Qb channel quads x (256 / Qb) row lanes per quad slice (C / 4 may exceed 256: slices are looped).
__global__ void __launch_bounds__(256)
col_stats_kernel(const float* __restrict__ x, float* __restrict__ part, long M, int C, int rows_per_chunk) {
    __shared__ float4 red[2][256];
    const int Q = C >> 2;
    const long r0 = static_cast<long>(blockIdx.x) * rows_per_chunk;
    const long r1 = min(r0 + static_cast<long>(rows_per_chunk), M);
    for (int qs = 0; qs < Q; qs += 256) {
        const int Qb = min(256, Q - qs);
        const int ppi = 256 / Qb;
        const int q = threadIdx.x % Qb, ps = threadIdx.x / Qb;
        float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ps < ppi) {
            const float* base = x + static_cast<size_t>(qs + q) * 4;
            long r = r0 + ps;
            for (; r + 3L * ppi < r1; r += 4L * ppi) {       // four rows in flight
                const float4 a = ldg4(base + r * C), b = ldg4(base + (r + ppi) * C), c = ldg4(base + (r + 2L * ppi) * C),
                             d = ldg4(base + (r + 3L * ppi) * C);
                s1.x += (a.x + b.x) + (c.x + d.x); s1.y += (a.y + b.y) + (c.y + d.y);
                s1.z += (a.z + b.z) + (c.z + d.z); s1.w += (a.w + b.w) + (c.w + d.w);
                s2.x += (a.x * a.x + b.x * b.x) + (c.x * c.x + d.x * d.x);
                s2.y += (a.y * a.y + b.y * b.y) + (c.y * c.y + d.y * d.y);
                s2.z += (a.z * a.z + b.z * b.z) + (c.z * c.z + d.z * d.z);
                s2.w += (a.w * a.w + b.w * b.w) + (c.w * c.w + d.w * d.w);
            }
            for (; r < r1; r += ppi) {
                const float4 a = ldg4(base + r * C);
                s1.x += a.x; s1.y += a.y; s1.z += a.z; s1.w += a.w;
                s2.x += a.x * a.x; s2.y += a.y * a.y; s2.z += a.z * a.z; s2.w += a.w * a.w;
            }
        }
        red[0][threadIdx.x] = s1;
        red[1][threadIdx.x] = s2;
        __syncthreads();
        if (ps == 0) {
#pragma unroll
            for (int wh = 0; wh < 2; ++wh) {
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int r = 0; r < ppi; ++r) {
                    const float4 g = red[wh][r * Qb + q];
                    t.x += g.x; t.y += g.y; t.z += g.z; t.w += g.w;
                }
                *reinterpret_cast<float4*>(part + (static_cast<size_t>(blockIdx.x) * 2 + wh) * C + (qs + q) * 4) = t;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ BN + activation
// y = act(x*scale + shift (+ res [*res_scale + res_shift]))  ->  fp32 (tf32-rounded on request) and / or (hi, lo) planes
__global__ void __launch_bounds__(256)
bn_act_kernel(const float4* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
              const float4* __restrict__ res, const float* __restrict__ res_scale, const float* __restrict__ res_shift,
              float4* __restrict__ y, __nv_bfloat16* __restrict__ ys, long long split_stride, long total4, int C4, int act,
              int round_out, unsigned char* __restrict__ mask_out) {
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < static_cast<unsigned>(total4); i += gridDim.x * 256u) {   // 32-bit index math
        const int c = static_cast<int>(i % static_cast<unsigned>(C4)) * 4;
        float4 v = __ldg(x + i);
        if (scale) {
            const float4 sc = ldg4(scale + c), sh = ldg4(shift + c);
            v.x = v.x * sc.x + sh.x; v.y = v.y * sc.y + sh.y; v.z = v.z * sc.z + sh.z; v.w = v.w * sc.w + sh.w;
        }
        if (res) {
            float4 r = __ldg(res + i);
            if (res_scale) {
                const float4 sc = ldg4(res_scale + c), sh = ldg4(res_shift + c);
                r.x = r.x * sc.x + sh.x; r.y = r.y * sc.y + sh.y; r.z = r.z * sc.z + sh.z; r.w = r.w * sc.w + sh.w;
            }
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        v.x = act_apply(v.x, act); v.y = act_apply(v.y, act); v.z = act_apply(v.z, act); v.w = act_apply(v.w, act);
        if (ys) store_split4(ys, split_stride, static_cast<size_t>(i) * 4, v);
        if (y) y[i] = round_out ? round4(v) : v;
        // 4 bits per float4: where the activation passed (the backward pass reads this byte instead of the 16-byte output)
        if (mask_out)
            mask_out[i] = static_cast<unsigned char>((v.x > 0.f ? 1u : 0u) | (v.y > 0.f ? 2u : 0u) | (v.z > 0.f ? 4u : 0u) |
                                                     (v.w > 0.f ? 8u : 0u));
    }
}

// ------------------------------------------------------------------------------------------------ BN backward
// part[chunk][0][c] = sum dz, part[chunk][1][c] = sum dz * xhat,  xhat = (x_raw - mean) * rstd
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ mask_src, const float* __restrict__ x_raw,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ scale,
                     const float* __restrict__ shift, float* __restrict__ part, long M, int C, int rows_per_chunk,
                     int mask_mode) {
    __shared__ float4 red[2][256];
    const int Q = C >> 2;
    const long r0 = static_cast<long>(blockIdx.x) * rows_per_chunk;
    const long r1 = min(r0 + static_cast<long>(rows_per_chunk), M);
    for (int qs = 0; qs < Q; qs += 256) {
        const int Qb = min(256, Q - qs);
        const int ppi = 256 / Qb;
        const int q = threadIdx.x % Qb, ps = threadIdx.x / Qb;
        const int c = (qs + q) * 4;
        float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ps < ppi) {
            const float4 mu = ldg4(mean + c), rs = ldg4(rstd + c);
            float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (mask_mode >= 2) { sc = ldg4(scale + c); sh = ldg4(shift + c); }
            for (long r = r0 + ps; r < r1; r += ppi) {
                const size_t off = static_cast<size_t>(r) * C + c;
                const float4 g = ldg4(dy + off), xr = ldg4(x_raw + off);
                float4 m = xr;
                if (mask_mode == 1) m = ldg4(mask_src + off);
                else if (mask_mode == 4) m = mask_byte4(reinterpret_cast<const unsigned char*>(mask_src)[off >> 2]);
                else if (mask_mode >= 2) m = make_float4(xr.x * sc.x + sh.x, xr.y * sc.y + sh.y, xr.z * sc.z + sh.z, xr.w * sc.w + sh.w);
                const float d0 = mask_apply(g.x, m.x, mask_mode), d1 = mask_apply(g.y, m.y, mask_mode),
                            d2 = mask_apply(g.z, m.z, mask_mode), d3 = mask_apply(g.w, m.w, mask_mode);
                s1.x += d0; s1.y += d1; s1.z += d2; s1.w += d3;
                s2.x += d0 * ((xr.x - mu.x) * rs.x); s2.y += d1 * ((xr.y - mu.y) * rs.y);
                s2.z += d2 * ((xr.z - mu.z) * rs.z); s2.w += d3 * ((xr.w - mu.w) * rs.w);
            }
        }
        red[0][threadIdx.x] = s1;
        red[1][threadIdx.x] = s2;
        __syncthreads();
        if (ps == 0) {
#pragma unroll
            for (int wh = 0; wh < 2; ++wh) {
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int r = 0; r < ppi; ++r) {
                    const float4 g = red[wh][r * Qb + q];
                    t.x += g.x; t.y += g.y; t.z += g.z; t.w += g.w;
                }
                *reinterpret_cast<float4*>(part + (static_cast<size_t>(blockIdx.x) * 2 + wh) * C + c) = t;
            }
        }
        __syncthreads();
    }
}

// merges the partials (fp64, fixed order): dgamma (+)= sum dz*xhat, dbeta (+)= sum dz,
// coef[0][c] = sum dz / count, coef[1][c] = sum dz*xhat / count   (zero when the layer normalised with running statistics)
__global__ void __launch_bounds__(1024)
bn_bwd_finalize_kernel(const float* __restrict__ part, int nparts, double count, float* __restrict__ dgamma,
                       float* __restrict__ dbeta, int accumulate, float* __restrict__ coef, int C, int batch_stats) {
    __shared__ double red[2][32][33];
    const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    double s1 = 0.0, s2 = 0.0;
    if (c < C)
#pragma unroll 8
        for (int i = pl; i < nparts; i += 32) {          // independent loads: 16 in flight per thread
            s1 += static_cast<double>(part[(static_cast<size_t>(i) * 2 + 0) * C + c]);
            s2 += static_cast<double>(part[(static_cast<size_t>(i) * 2 + 1) * C + c]);
        }
    red[0][pl][cl] = s1;
    red[1][pl][cl] = s2;
    __syncthreads();
    if (pl == 0 && c < C) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int r = 0; r < 32; ++r) { a += red[0][r][cl]; b += red[1][r][cl]; }
        if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + static_cast<float>(b);
        if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + static_cast<float>(a);
        coef[c] = batch_stats ? static_cast<float>(a / count) : 0.f;
        coef[C + c] = batch_stats ? static_cast<float>(b / count) : 0.f;
    }
}

// dx = gamma*rstd * (dz - coef0 - xhat*coef1)   (+ dz itself for the identity branch of a residual block)
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ mask_src, const float* __restrict__ x_raw,
                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ scale,
                    const float* __restrict__ shift, const float* __restrict__ gamma, const float* __restrict__ coef,
                    float* __restrict__ dx, float* __restrict__ dz_out, long total4, int C, int mask_mode, int round_out) {
    const int C4 = C >> 2;
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < static_cast<unsigned>(total4); i += gridDim.x * 256u) {   // 32-bit index math
        const int c = static_cast<int>(i % static_cast<unsigned>(C4)) * 4;
        const size_t off = static_cast<size_t>(i) * 4;
        const float4 g = ldg4(dy + off), xr = ldg4(x_raw + off);
        float4 m = xr;
        if (mask_mode == 1) m = ldg4(mask_src + off);
        else if (mask_mode == 4) m = mask_byte4(reinterpret_cast<const unsigned char*>(mask_src)[i]);
        else if (mask_mode >= 2) {
            const float4 sc = ldg4(scale + c), sh = ldg4(shift + c);
            m = make_float4(xr.x * sc.x + sh.x, xr.y * sc.y + sh.y, xr.z * sc.z + sh.z, xr.w * sc.w + sh.w);
        }
        const float4 dz = make_float4(mask_apply(g.x, m.x, mask_mode), mask_apply(g.y, m.y, mask_mode),
                                      mask_apply(g.z, m.z, mask_mode), mask_apply(g.w, m.w, mask_mode));
        const float4 mu = ldg4(mean + c), rs = ldg4(rstd + c), gm = ldg4(gamma + c), c0 = ldg4(coef + c), c1 = ldg4(coef + C + c);
        float4 o;
        o.x = gm.x * rs.x * (dz.x - c0.x - (xr.x - mu.x) * rs.x * c1.x);
        o.y = gm.y * rs.y * (dz.y - c0.y - (xr.y - mu.y) * rs.y * c1.y);
        o.z = gm.z * rs.z * (dz.z - c0.z - (xr.z - mu.z) * rs.z * c1.z);
        o.w = gm.w * rs.w * (dz.w - c0.w - (xr.w - mu.w) * rs.w * c1.w);
        *reinterpret_cast<float4*>(dx + off) = round_out ? round4(o) : o;
        if (dz_out) *reinterpret_cast<float4*>(dz_out + off) = dz;
    }
}

// ------------------------------------------------------------------------------------------------ grouped 3x3 conv
// y[n,ho,wo,g*CPG+co] = sum_{kh,kw,ci} f(x[n, ho*s+kh-1, wo*s+kw-1, g*CPG+ci]) * w[g*CPG+co][ci][kh][kw]
// f = relu(x*scale+shift) (the producer's BatchNorm + ReLU, applied on load; zero padding AFTER f) or identity.
// FP32 on the CUDA cores: 4..32 channels per group do not fill a tensor-core tile, and the op is ~6 % of the encoder's
// FLOPs.  A lane owns PPT output pixels along W of one group; a quarter-warp = 8 neighbouring groups (their CPG channels
// are contiguous: coalesced global loads), weights of the 8 groups sit interleaved in shared memory so that one LDS.128
// per (tap, co, 4 input channels) serves 4*PPT FMAs.  TRANSPOSED: the same kernel computes the stride-1 data gradient
// (w^T with flipped taps, gathered while staging the weights).
template <int CPG> struct GconvCfg;
// PPT = output pixels per lane; GL = groups per block (their weights for ALL input channels stay in shared memory for the
// whole kernel: 4.6 / 18 / 74 / 147 KB); CI4C = chunk of input-channel quads of the stride-2 data-gradient kernel
template <> struct GconvCfg<4>  { static constexpr int PPT = 4, GL = 8, CI4C = 1; };
template <> struct GconvCfg<8>  { static constexpr int PPT = 4, GL = 8, CI4C = 2; };
template <> struct GconvCfg<16> { static constexpr int PPT = 2, GL = 8, CI4C = 2; };
template <> struct GconvCfg<32> { static constexpr int PPT = 2, GL = 4, CI4C = 1; };

template <int CPG, int STRIDE, bool TRANSPOSED>
__global__ void __launch_bounds__(256)
gconv3x3_fwd_kernel(const float* __restrict__ x, const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                    const float* __restrict__ w, float* __restrict__ y, float* __restrict__ part, int N, int H, int W,
                    int C, int Ho, int Wo) {
    constexpr int PPT = GconvCfg<CPG>::PPT, GL = GconvCfg<CPG>::GL, PL = 32 / GL, CI4 = CPG / 4;
    constexpr int NCOL = (PPT - 1) * STRIDE + 3;
    constexpr int SPT = 8 * PL;                       // pixel strips per tile (8 warps x PL lanes)
    extern __shared__ float4 sw[];                    // [tap 9][ci4 CI4][co CPG][gl GL]  (+ stats scratch reuses it)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % GL, pl = lane / GL;
    const int g = blockIdx.y * GL + gl;               // absolute group
    const int cbase = g * CPG;
    // the weights of this block's GL groups, staged once: sw[((tap*CI4 + ci4)*CPG + co)*GL + gl] = 4 input channels
    for (int i = threadIdx.x; i < 9 * CI4 * CPG * GL; i += 256) {
        const int sgl = i % GL;
        const int co = (i / GL) % CPG;
        const int ci4 = (i / (GL * CPG)) % CI4;
        const int tap = i / (GL * CPG * CI4);
        const int sg = blockIdx.y * GL + sgl;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // forward: w[sg*CPG + co][ci4*4 + j][tap];  transposed: output channel `co` is an INPUT channel of the forward
            // conv, the reduction runs over its output channels, taps flipped: w[sg*CPG + ci4*4 + j][co][8 - tap]
            const size_t idx = TRANSPOSED ? (static_cast<size_t>(sg * CPG + ci4 * 4 + j) * CPG + co) * 9 + (8 - tap)
                                          : (static_cast<size_t>(sg * CPG + co) * CPG + ci4 * 4 + j) * 9 + tap;
            v[j] = __ldg(w + idx);
        }
        sw[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();
    const unsigned strips_w = (Wo + PPT - 1) / PPT;
    const unsigned total_strips = static_cast<unsigned>(N) * Ho * strips_w;     // < 2^31 (host check)
    float st1[CPG], st2[CPG];
#pragma unroll
    for (int j = 0; j < CPG; ++j) { st1[j] = 0.f; st2[j] = 0.f; }

    for (unsigned tile = blockIdx.x; tile * SPT < total_strips; tile += gridDim.x) {
        const unsigned strip = tile * SPT + warp * PL + pl;
        if (strip >= total_strips) continue;
        const unsigned sw_i = strip % strips_w;
        const unsigned t2 = strip / strips_w;
        const int ho = static_cast<int>(t2 % Ho);
        const int n = static_cast<int>(t2 / Ho);
        const int wo0 = static_cast<int>(sw_i) * PPT;
        float acc[PPT][CPG];
#pragma unroll
        for (int u = 0; u < PPT; ++u)
#pragma unroll
            for (int j = 0; j < CPG; ++j) acc[u][j] = 0.f;
#pragma unroll 1
        for (int ci4 = 0; ci4 < CI4; ++ci4) {
            float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in_scale) { sc = ldg4(in_scale + cbase + ci4 * 4); sh = ldg4(in_shift + cbase + ci4 * 4); }
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int hi = ho * STRIDE + kh - 1;
                const bool hok = hi >= 0 && hi < H;
                const float* rowp = x + (static_cast<size_t>(n * H + (hok ? hi : 0)) * W) * C + cbase + ci4 * 4;
                float4 xin[NCOL];
#pragma unroll
                for (int col = 0; col < NCOL; ++col) {
                    const int wi = wo0 * STRIDE + col - 1;
                    const bool ok = hok && wi >= 0 && wi < W;
                    float4 v = ldg4(rowp + static_cast<size_t>(ok ? wi : 0) * C);
                    if (in_scale) {
                        v.x = fmaxf(v.x * sc.x + sh.x, 0.f); v.y = fmaxf(v.y * sc.y + sh.y, 0.f);
                        v.z = fmaxf(v.z * sc.z + sh.z, 0.f); v.w = fmaxf(v.w * sc.w + sh.w, 0.f);
                    }
                    xin[col] = ok ? v : make_float4(0.f, 0.f, 0.f, 0.f);    // zero padding AFTER the activation
                }
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float4* wrow = sw + ((kh * 3 + kw) * CI4 + ci4) * CPG * GL + gl;
#pragma unroll
                    for (int co = 0; co < CPG; ++co) {
                        const float4 wv = wrow[co * GL];
#pragma unroll
                        for (int u = 0; u < PPT; ++u) {
                            const float4 xv = xin[u * STRIDE + kw];
                            acc[u][co] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc[u][co]))));
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            const int wo = wo0 + u;
            if (wo < Wo) {
                float* o = y + (static_cast<size_t>(n * Ho + ho) * Wo + wo) * C + cbase;
#pragma unroll
                for (int j = 0; j < CPG; j += 4)
                    *reinterpret_cast<float4*>(o + j) = make_float4(acc[u][j], acc[u][j + 1], acc[u][j + 2], acc[u][j + 3]);
#pragma unroll
                for (int j = 0; j < CPG; ++j) { st1[j] += acc[u][j]; st2[j] += acc[u][j] * acc[u][j]; }
            }
        }
    }
    if (part) {
        // per-channel sums of this block: the PL pixel lanes of a warp (shuffle), then the 8 warps (shared memory, fixed order)
        __syncthreads();
        float* red = reinterpret_cast<float*>(sw);           // [2][8 warps][GL][CPG]
#pragma unroll
        for (int j = 0; j < CPG; ++j) {
            float a = st1[j], b = st2[j];
#pragma unroll
            for (int m = GL; m < 32; m <<= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, m);
                b += __shfl_xor_sync(0xffffffffu, b, m);
            }
            if (pl == 0) {
                red[((0 * 8 + warp) * GL + gl) * CPG + j] = a;
                red[((1 * 8 + warp) * GL + gl) * CPG + j] = b;
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * GL * CPG; i += 256) {
            const int wh = i / (GL * CPG), r = i % (GL * CPG);        // r = gl*CPG + j: channel within the block's groups
            float s_ = 0.f;
#pragma unroll
            for (int wq = 0; wq < 8; ++wq) s_ += red[((wh * 8 + wq) * GL) * CPG + r];
            part[(static_cast<size_t>(blockIdx.x) * 2 + wh) * C + blockIdx.y * GL * CPG + r] = s_;
        }
    }
}

// stride-2 data gradient (three layers of the net): dx[n,hi,wi,g*CPG+ci] = sum over the taps with (hi+1-kh), (wi+1-kw)
// even of dy[n,(hi+1-kh)/2,(wi+1-kw)/2, g*CPG+co] * w[g*CPG+co][ci][kh][kw].  One input pixel x one group per lane; the
// output channels of the forward conv are walked in chunks of 4*CO4C (weights of a chunk in shared memory), later chunks
// add to what the earlier ones stored.
template <int CPG>
__global__ void __launch_bounds__(256)
gconv3x3_dgrad_s2_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int N, int H,
                         int W, int C, int Ho, int Wo) {
    constexpr int CO4C = GconvCfg<CPG>::CI4C;
    extern __shared__ float4 sw[];                    // [tap 9][co4l CO4C][ci CPG][gl 8] = 4 co values
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane & 7, pl = lane >> 3;
    const int g = blockIdx.y * 8 + gl;
    const int cbase = g * CPG;
    const unsigned total = static_cast<unsigned>(N) * H * W;          // < 2^31 (host check)
    for (int cc = 0; cc < CPG / 4; cc += CO4C) {
        __syncthreads();
        for (int i = threadIdx.x; i < 9 * CO4C * CPG * 8; i += 256) {
            const int sgl = i & 7;
            const int ci = (i >> 3) % CPG;
            const int co4l = ((i >> 3) / CPG) % CO4C;
            const int tap = (i >> 3) / (CPG * CO4C);
            const int sg = blockIdx.y * 8 + sgl;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[j] = __ldg(w + (static_cast<size_t>(sg * CPG + (cc + co4l) * 4 + j) * CPG + ci) * 9 + tap);
            sw[i] = make_float4(v[0], v[1], v[2], v[3]);
        }
        __syncthreads();
        for (unsigned p = blockIdx.x * 32u + warp * 4 + pl; p < total; p += gridDim.x * 32u) {
            const int wi = static_cast<int>(p % static_cast<unsigned>(W));
            const unsigned t2 = p / static_cast<unsigned>(W);
            const int hi = static_cast<int>(t2 % static_cast<unsigned>(H));
            const int n = static_cast<int>(t2 / static_cast<unsigned>(H));
            float* o = dx + static_cast<size_t>(p) * C + cbase;
            float acc[CPG];
            if (cc == 0) {
#pragma unroll
                for (int j = 0; j < CPG; ++j) acc[j] = 0.f;
            } else {
#pragma unroll
                for (int j = 0; j < CPG; j += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(o + j);
                    acc[j] = t.x; acc[j + 1] = t.y; acc[j + 2] = t.z; acc[j + 3] = t.w;
                }
            }
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int h2 = hi + 1 - kh;
                if (h2 < 0 || (h2 & 1) || (h2 >> 1) >= Ho) continue;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int w2 = wi + 1 - kw;
                    if (w2 < 0 || (w2 & 1) || (w2 >> 1) >= Wo) continue;
                    const float* src = dy + (static_cast<size_t>(n * Ho + (h2 >> 1)) * Wo + (w2 >> 1)) * C + cbase + cc * 4;
#pragma unroll
                    for (int co4l = 0; co4l < CO4C; ++co4l) {
                        const float4 g4 = ldg4(src + co4l * 4);
                        const float4* wrow = sw + (((kh * 3 + kw) * CO4C + co4l) * CPG) * 8 + gl;
#pragma unroll
                        for (int ci = 0; ci < CPG; ++ci) {
                            const float4 wv = wrow[ci * 8];
                            acc[ci] = fmaf(g4.x, wv.x, fmaf(g4.y, wv.y, fmaf(g4.z, wv.z, fmaf(g4.w, wv.w, acc[ci]))));
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < CPG; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        }
    }
}

// weight gradient: dW[g*CPG+co][ci][tap] = sum_{n,ho,wo} dy[n,ho,wo,g*CPG+co] * f(x[n, ho*s+kh-1, wo*s+kw-1, g*CPG+ci]).
// A thread owns (group, 4 output channels, 4 input channels, ONE filter row kh) = 48 accumulators and walks whole output
// ROWS of its chunk: the (n, ho) decomposition and the input-row test happen once per row, the inner loop over wo has no
// divisions and keeps 8 loads (2 pixels x (dy quad + three taps of x)) in flight.  (Earlier versions: 72 accumulators with
// a branch per tap serialised on the L2 latency — 1.2 ms per layer; per-pixel index arithmetic — 0.32 ms.)
// Partial sums go to ws[chunk][tap][co_abs][ci] (16-byte stores), merged in a fixed order by the reduce kernels.
// block = TS owner threads x (256 / TS) pixel lanes;  grid = (row chunks, owners / TS)
template <int CPG>
__global__ void __launch_bounds__(256)
gconv3x3_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                      const float* __restrict__ dy, float* __restrict__ ws, int N, int H, int W, int C, int stride, int Ho,
                      int Wo, int rows_per_chunk, int TS) {
    extern __shared__ float4 red4[];                  // [lanes-1][TS][12] float4 for the cross-lane merge
    const int t = threadIdx.x % TS, lanes = blockDim.x / TS, ln = threadIdx.x / TS;
    const int owner = blockIdx.y * TS + t;            // (g, co4, ci4, kh): kh fastest, then ci4
    const int kh = owner % 3;
    const int ci4 = (owner / 3) % (CPG / 4);
    const int co4 = (owner / (3 * (CPG / 4))) % (CPG / 4);
    const int g = owner / (3 * (CPG / 4) * (CPG / 4));
    const int cin = g * CPG + ci4 * 4, cout = g * CPG + co4 * 4;
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool act = in_scale != nullptr;
    if (act) { sc = ldg4(in_scale + cin); sh = ldg4(in_shift + cin); }
    float4 acc[4][3];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[c][k] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int rows_total = N * Ho;
    const int r0 = blockIdx.x * rows_per_chunk;
    const int r1 = min(r0 + rows_per_chunk, rows_total);
    for (int row = r0; row < r1; ++row) {
        const int n = row / Ho, ho = row - n * Ho;
        const int hi = ho * stride + kh - 1;
        if (hi < 0 || hi >= H) continue;              // this filter row reads padding for the whole output row
        const float* xrow = x + (static_cast<size_t>(n * H + hi) * W) * C + cin;
        const float* dyrow = dy + (static_cast<size_t>(row) * Wo) * C + cout;
        for (int wb = ln * 2; wb < Wo; wb += 2 * lanes) {
            float4 d[2], v[2][3];
            bool ok[2][3];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int wo = wb + u;
                const bool pv = wo < Wo;
                d[u] = ldg4(dyrow + static_cast<size_t>(pv ? wo : wb) * C);
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int wi = wo * stride + kw - 1;
                    const bool in = pv && wi >= 0 && wi < W;
                    ok[u][kw] = in;
                    v[u][kw] = ldg4(xrow + static_cast<size_t>(in ? wi : 0) * C);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float dd[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    if (!ok[u][kw]) continue;             // zero padding applies to the activation, not to the raw input
                    float4 a = v[u][kw];
                    if (act) {
                        a.x = fmaxf(a.x * sc.x + sh.x, 0.f); a.y = fmaxf(a.y * sc.y + sh.y, 0.f);
                        a.z = fmaxf(a.z * sc.z + sh.z, 0.f); a.w = fmaxf(a.w * sc.w + sh.w, 0.f);
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float4& o = acc[c][kw];
                        o.x = fmaf(dd[c], a.x, o.x); o.y = fmaf(dd[c], a.y, o.y); o.z = fmaf(dd[c], a.z, o.z); o.w = fmaf(dd[c], a.w, o.w);
                    }
                }
            }
        }
    }
    // merge the pixel lanes (fixed order), lane 0 writes
    if (lanes > 1) {
        if (ln > 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int k = 0; k < 3; ++k) red4[(static_cast<size_t>(ln - 1) * TS + t) * 12 + c * 3 + k] = acc[c][k];
        }
        __syncthreads();
        if (ln == 0) {
            for (int l = 1; l < lanes; ++l)
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const float4 o = red4[(static_cast<size_t>(l - 1) * TS + t) * 12 + c * 3 + k];
                        acc[c][k].x += o.x; acc[c][k].y += o.y; acc[c][k].z += o.z; acc[c][k].w += o.w;
                    }
        }
    }
    if (ln == 0) {
        float* base = ws + static_cast<size_t>(blockIdx.x) * 9 * C * CPG;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int k = 0; k < 3; ++k)
                *reinterpret_cast<float4*>(base + (static_cast<size_t>(kh * 3 + k) * C + cout + c) * CPG + ci4 * 4) = acc[c][k];
    }
}

// dW[co][ci][tap] (+)= sum over chunks of ws[chunk][tap][co][ci], in two stages with a fixed order: stage 1 — block
// (x, j) sums chunks j, j + CG, ... of 256 consecutive (tap, co, ci) elements (coalesced) into ws2[j][...]; stage 2 — sums
// the <= 32 stage-1 rows and transposes to [co][ci][tap].  (One thread per output walking ALL chunks serialised 2368
// dependent L2 round trips: 0.42 ms per layer.)
__global__ void __launch_bounds__(256)
gconv3x3_wgrad_reduce1_kernel(const float* __restrict__ ws, float* __restrict__ ws2, int chunks, int CG, unsigned total) {
    const unsigned i = blockIdx.x * 256u + threadIdx.x;
    if (i >= total) return;
    const float* src = ws + i;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int k = blockIdx.y;
    for (; k + 3 * CG < chunks; k += 4 * CG) {
        s0 += __ldg(src + static_cast<size_t>(k) * total);
        s1 += __ldg(src + static_cast<size_t>(k + CG) * total);
        s2 += __ldg(src + static_cast<size_t>(k + 2 * CG) * total);
        s3 += __ldg(src + static_cast<size_t>(k + 3 * CG) * total);
    }
    for (; k < chunks; k += CG) s0 += __ldg(src + static_cast<size_t>(k) * total);
    ws2[static_cast<size_t>(blockIdx.y) * total + i] = (s0 + s1) + (s2 + s3);
}

__global__ void __launch_bounds__(256)
gconv3x3_wgrad_reduce2_kernel(const float* __restrict__ ws2, float* __restrict__ dw, int CG, int C, int cpg, int accumulate) {
    const unsigned total = static_cast<unsigned>(C) * cpg * 9;
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < total; i += gridDim.x * 256u) {
        const unsigned tap = i % 9u;
        const unsigned cc = i / 9u;                               // co*cpg + ci
        const float* src = ws2 + static_cast<size_t>(tap) * C * cpg + cc;
        float s = 0.f;
        for (int k = 0; k < CG; ++k) s += __ldg(src + static_cast<size_t>(k) * total);
        dw[i] = (accumulate ? dw[i] : 0.f) + s;
    }
}

// ------------------------------------------------------------------------------------------------ 7x7 stride-2 stem
// x NCHW (N,3,H,W) -> patch matrix col[(n,ho,wo)][KP]: column c*49 + kh*7 + kw (147 real, zero up to KP) — the operand of a
// 1x1 tensor-core GEMM against the (64, 147 -> KP) stem weight.  fp32 (tf32-rounded: weight-gradient operand) and / or
// (hi, lo) bf16 planes (bf16x3 forward operand).
__global__ void __launch_bounds__(256)
im2col7x7_s2_kernel(const float* __restrict__ x, float* __restrict__ col, __nv_bfloat16* __restrict__ cols,
                    long long split_stride, int N, int H, int W, int Ho, int Wo, int KP) {
    const int K4 = KP >> 2;
    const long total4 = static_cast<long>(N) * Ho * Wo * K4;
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < static_cast<unsigned>(total4); i += gridDim.x * 256u) {   // 32-bit index math
        const int k0 = static_cast<int>(i % static_cast<unsigned>(K4)) * 4;
        const unsigned p = i / static_cast<unsigned>(K4);
        const int wo = static_cast<int>(p % static_cast<unsigned>(Wo));
        const int ho = static_cast<int>((p / static_cast<unsigned>(Wo)) % static_cast<unsigned>(Ho));
        const long n = p / (static_cast<unsigned>(Wo) * static_cast<unsigned>(Ho));
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + j;
            float t = 0.f;
            if (k < 147) {
                const int c = k / 49, r = k - c * 49;
                const int kh = r / 7, kw = r - kh * 7;
                const int hi = ho * 2 + kh - 3, wi = wo * 2 + kw - 3;
                if (hi >= 0 && hi < H && wi >= 0 && wi < W) t = __ldg(x + ((n * 3 + c) * H + hi) * W + wi);
            }
            v[j] = t;
        }
        const float4 o = make_float4(v[0], v[1], v[2], v[3]);
        if (cols) store_split4(cols, split_stride, static_cast<size_t>(i) * 4, o);
        if (col) *reinterpret_cast<float4*>(col + static_cast<size_t>(i) * 4) = round4(o);
    }
}

// ------------------------------------------------------------------------------------------------ max-pool 3x3 s2 p1
// y = maxpool(relu(x*scale+shift)) over NHWC, idx = tap (kh*3+kw) of the first maximum (for the backward gather)
__global__ void __launch_bounds__(256)
maxpool3x3s2_fwd_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                        float* __restrict__ y, __nv_bfloat16* __restrict__ ys, long long split_stride,
                        unsigned char* __restrict__ idx, int N, int H, int W, int C, int Ho, int Wo, int round_out) {
    const int C4 = C >> 2;
    const long total4 = static_cast<long>(N) * Ho * Wo * C4;
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < static_cast<unsigned>(total4); i += gridDim.x * 256u) {   // 32-bit index math
        const int c = static_cast<int>(i % static_cast<unsigned>(C4)) * 4;
        const unsigned p = i / static_cast<unsigned>(C4);
        const int wo = static_cast<int>(p % static_cast<unsigned>(Wo));
        const int ho = static_cast<int>((p / static_cast<unsigned>(Wo)) % static_cast<unsigned>(Ho));
        const long n = p / (static_cast<unsigned>(Wo) * static_cast<unsigned>(Ho));
        const float4 sc = ldg4(scale + c), sh = ldg4(shift + c);
        float best[4] = {-1.f, -1.f, -1.f, -1.f};          // activations are >= 0: any in-image tap beats -1
        int bi[4] = {0, 0, 0, 0};
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int hi = ho * 2 + kh - 1;
            if (hi < 0 || hi >= H) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int wi = wo * 2 + kw - 1;
                if (wi < 0 || wi >= W) continue;
                const float4 v = ldg4(x + ((n * H + hi) * W + wi) * C + c);
                const float a[4] = {fmaxf(v.x * sc.x + sh.x, 0.f), fmaxf(v.y * sc.y + sh.y, 0.f),
                                    fmaxf(v.z * sc.z + sh.z, 0.f), fmaxf(v.w * sc.w + sh.w, 0.f)};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (a[j] > best[j]) { best[j] = a[j]; bi[j] = kh * 3 + kw; }
            }
        }
        const float4 o = make_float4(best[0], best[1], best[2], best[3]);
        if (ys) store_split4(ys, split_stride, static_cast<size_t>(i) * 4, o);
        if (y) *reinterpret_cast<float4*>(y + static_cast<size_t>(i) * 4) = round_out ? round4(o) : o;
        if (idx) *reinterpret_cast<uchar4*>(idx + static_cast<size_t>(i) * 4) = make_uchar4(bi[0], bi[1], bi[2], bi[3]);
    }
}

// dx[n,hi,wi,c] = sum over the windows (ho,wo) that contain (hi,wi) and whose recorded maximum is this tap of dy[n,ho,wo,c]
__global__ void __launch_bounds__(256)
maxpool3x3s2_bwd_kernel(const float* __restrict__ dy, const unsigned char* __restrict__ idx, float* __restrict__ dx, int N,
                        int H, int W, int C, int Ho, int Wo) {
    const int C4 = C >> 2;
    const long total4 = static_cast<long>(N) * H * W * C4;
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < static_cast<unsigned>(total4); i += gridDim.x * 256u) {   // 32-bit index math
        const int c = static_cast<int>(i % static_cast<unsigned>(C4)) * 4;
        const unsigned p = i / static_cast<unsigned>(C4);
        const int wi = static_cast<int>(p % static_cast<unsigned>(W));
        const int hi = static_cast<int>((p / static_cast<unsigned>(W)) % static_cast<unsigned>(H));
        const long n = p / (static_cast<unsigned>(W) * static_cast<unsigned>(H));
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int h2 = hi + 1 - kh;
            if (h2 < 0 || (h2 & 1) || (h2 >> 1) >= Ho) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int w2 = wi + 1 - kw;
                if (w2 < 0 || (w2 & 1) || (w2 >> 1) >= Wo) continue;
                const size_t off = ((n * Ho + (h2 >> 1)) * Wo + (w2 >> 1)) * C + c;
                const uchar4 t = *reinterpret_cast<const uchar4*>(idx + off);
                const float4 g = ldg4(dy + off);
                const int tap = kh * 3 + kw;
                if (t.x == tap) o.x += g.x;
                if (t.y == tap) o.y += g.y;
                if (t.z == tap) o.z += g.z;
                if (t.w == tap) o.w += g.w;
            }
        }
        *reinterpret_cast<float4*>(dx + static_cast<size_t>(i) * 4) = o;
    }
}

// ------------------------------------------------------------------------------------------------ stride-2 subsample
// y[n,h,w,:] = x[n,2h,2w,:] for the fp32 copy and / or the (hi, lo) planes (the operand of a stride-2 1x1 convolution)
__global__ void __launch_bounds__(256)
subsample2_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ xs, long long in_split_stride,
                  float* __restrict__ y, __nv_bfloat16* __restrict__ ys, long long out_split_stride, int N, int Ho, int Wo,
                  int C) {
    const int C4 = C >> 2;
    const long total4 = static_cast<long>(N) * Ho * Wo * C4;
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < static_cast<unsigned>(total4); i += gridDim.x * 256u) {   // 32-bit index math
        const int c = static_cast<int>(i % static_cast<unsigned>(C4)) * 4;
        const unsigned p = i / static_cast<unsigned>(C4);
        const int wo = static_cast<int>(p % static_cast<unsigned>(Wo));
        const int ho = static_cast<int>((p / static_cast<unsigned>(Wo)) % static_cast<unsigned>(Ho));
        const long n = p / (static_cast<unsigned>(Wo) * static_cast<unsigned>(Ho));
        const size_t src = ((n * 2 * Ho + 2 * ho) * (2 * Wo) + 2 * wo) * C + c;
        if (y) *reinterpret_cast<float4*>(y + static_cast<size_t>(i) * 4) = ldg4(x + src);
        if (ys) {
            *reinterpret_cast<uint2*>(ys + static_cast<size_t>(i) * 4) = *reinterpret_cast<const uint2*>(xs + src);
            *reinterpret_cast<uint2*>(ys + out_split_stride + static_cast<size_t>(i) * 4) =
                *reinterpret_cast<const uint2*>(xs + in_split_stride + src);
        }
    }
}

// dx[n,2h,2w,:] += dsub[n,h,w,:]
__global__ void __launch_bounds__(256)
scatter_add2_kernel(const float* __restrict__ dsub, float* __restrict__ dx, int N, int Ho, int Wo, int C) {
    const int C4 = C >> 2;
    const long total4 = static_cast<long>(N) * Ho * Wo * C4;
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < static_cast<unsigned>(total4); i += gridDim.x * 256u) {   // 32-bit index math
        const int c = static_cast<int>(i % static_cast<unsigned>(C4)) * 4;
        const unsigned p = i / static_cast<unsigned>(C4);
        const int wo = static_cast<int>(p % static_cast<unsigned>(Wo));
        const int ho = static_cast<int>((p / static_cast<unsigned>(Wo)) % static_cast<unsigned>(Ho));
        const long n = p / (static_cast<unsigned>(Wo) * static_cast<unsigned>(Ho));
        float4* d = reinterpret_cast<float4*>(dx + ((n * 2 * Ho + 2 * ho) * (2 * Wo) + 2 * wo) * C + c);
        const float4 g = ldg4(dsub + static_cast<size_t>(i) * 4);
        float4 v = *d;
        v.x += g.x; v.y += g.y; v.z += g.z; v.w += g.w;
        *d = v;
    }
}

// ------------------------------------------------------------------------------------------------ global average pool
// y[n][c] = mean over p of x[n,p,c];  grid = (channel-quad groups of 32, N), block = 32 lanes x 8 warps over the pixels
__global__ void __launch_bounds__(256)
avgpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int HW, int C) {
    __shared__ float4 red[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + lane) * 4;
    const long n = blockIdx.y;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C)
        for (int p = warp; p < HW; p += 8) {
            const float4 v = ldg4(x + (n * HW + p) * C + c);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    red[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && c < C) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 8; ++r) { const float4 v = red[r][lane]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
        const float inv = 1.f / static_cast<float>(HW);
        *reinterpret_cast<float4*>(y + n * C + c) = make_float4(t.x * inv, t.y * inv, t.z * inv, t.w * inv);
    }
}

// dx[n,p,c] = dy[n][c] / HW
__global__ void __launch_bounds__(256)
avgpool_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, long total4, int HW, int C4) {
    const float inv = 1.f / static_cast<float>(HW);
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < static_cast<unsigned>(total4); i += gridDim.x * 256u) {   // 32-bit index math
        const int c4 = static_cast<int>(i % static_cast<unsigned>(C4));
        const long n = i / (static_cast<unsigned>(C4) * static_cast<unsigned>(HW));
        const float4 g = ldg4(dy + (n * C4 + c4) * 4);
        *reinterpret_cast<float4*>(dx + static_cast<size_t>(i) * 4) = make_float4(g.x * inv, g.y * inv, g.z * inv, g.w * inv);
    }
}

// ------------------------------------------------------------------------------------------------ small strided SGEMM
// C[i][j] (+)= alpha * sum_k A(i,k) * B(k,j) (+ bias[j]),  A(i,k) = A[i*sai + k*sak],  B(k,j) = B[k*sbk + j*sbj]
// (fp32, 64x64 tile, K step 16; alpha = *alpha_dev or 1).  For the classifier / projector layers and their gradients
// (M <= 64 rows): y = x W^T, dgrad = dy W, wgrad = dy^T x.  The tile loaders run along whichever axis of the operand
// is contiguous (AK / BK: the k axis), so every operand streams in full 32-byte sectors.
template <bool AK, bool BK>
__global__ void __launch_bounds__(256)
sgemm_strided_kernel(const float* __restrict__ A, long sai, long sak, const float* __restrict__ B, long sbk, long sbj,
                     float* __restrict__ Cm, const float* __restrict__ alpha_dev, const float* __restrict__ bias, int M, int N,
                     int K, int accumulate, int k_per_split, float* __restrict__ ws) {
    __shared__ float As[16][65], Bs[16][65];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    // split-K (gridDim.z > 1): this block reduces k in [kb, ke) and stores raw partial sums to ws[z][M][N]; a second
    // kernel merges them in a fixed order (a (8 x 13056) @ (13056 x 768) product would otherwise run on 12 blocks)
    const int kb = blockIdx.z * k_per_split;
    const int ke = min(kb + k_per_split, K);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = kb; k0 < ke; k0 += 16) {
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            const int kk = AK ? (e & 15) : (e >> 6), r = AK ? (e >> 4) : (e & 63);
            const int m = m0 + r, k = k0 + kk;
            As[kk][r] = (m < M && k < ke) ? __ldg(A + m * sai + k * sak) : 0.f;
        }
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            const int kk = BK ? (e & 15) : (e >> 6), r = BK ? (e >> 4) : (e & 63);
            const int n = n0 + r, k = k0 + kk;
            Bs[kk][r] = (n < N && k < ke) ? __ldg(B + k * sbk + n * sbj) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    const float alpha = alpha_dev ? __ldg(alpha_dev) : 1.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) {
                if (gridDim.z > 1) {
                    ws[(static_cast<size_t>(blockIdx.z) * M + m) * N + n] = acc[i][j];
                } else {
                    float* o = Cm + static_cast<size_t>(m) * N + n;
                    *o = (accumulate ? *o : 0.f) + alpha * acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256)
sgemm_splitk_reduce_kernel(const float* __restrict__ ws, float* __restrict__ Cm, const float* __restrict__ alpha_dev,
                           const float* __restrict__ bias, int splits, int M, int N, int accumulate) {
    const float alpha = alpha_dev ? __ldg(alpha_dev) : 1.f;
    const int total = M * N;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
        float s_ = 0.f;
        for (int z = 0; z < splits; ++z) s_ += __ldg(ws + static_cast<size_t>(z) * total + i);
        Cm[i] = (accumulate ? Cm[i] : 0.f) + alpha * s_ + (bias ? __ldg(bias + i % N) : 0.f);
    }
}

static int ew_blocks(long total4) {      // callers keep total4 < 2^31 (the kernels index with 32-bit integers)
    long blocks = (total4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

// row chunks of the grouped weight gradient: about 148 x 8 blocks in total, whole output rows per chunk
static int gconv_wgrad_plan(int rows, int C, int cpg, int* rows_per_chunk, int* TS, int* slices) {
    const int owners = 3 * (C * cpg / 16);          // (group, output-channel quad, input-channel quad, filter row)
    int ts = owners < 256 ? owners : 256;
    while (owners % ts) --ts;                       // the largest divisor <= 256 (96 / 192 / 256 / 256 for 32 groups)
    *TS = ts;
    *slices = owners / ts;
    long chunks = (148L * 8) / *slices;
    if (chunks < 1) chunks = 1;
    int rpc = static_cast<int>((rows + chunks - 1) / chunks);
    if (rpc < 1) rpc = 1;
    *rows_per_chunk = rpc;
    return (rows + rpc - 1) / rpc;
}

template <int CPG, int STRIDE, bool TR>
static int launch_gconv(const float* x, const float* sc, const float* sh, const float* w, float* y, float* part, int N,
                        int H, int W, int C, int Ho, int Wo, int blocks_x, cudaStream_t st) {
    constexpr int GL = GconvCfg<CPG>::GL;
    constexpr int kSmem = 9 * (CPG / 4) * CPG * GL * 16;
    constexpr int kRed = 2 * 8 * GL * CPG * 4;
    constexpr int kBytes = kSmem > kRed ? kSmem : kRed;
    static bool attr_set = false;
    if (!attr_set) {
        B200LP_CHECK_CUDA(cudaFuncSetAttribute(gconv3x3_fwd_kernel<CPG, STRIDE, TR>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, kBytes));
        attr_set = true;
    }
    dim3 grid(blocks_x, C / (GL * CPG));
    gconv3x3_fwd_kernel<CPG, STRIDE, TR><<<grid, 256, kBytes, st>>>(x, sc, sh, w, y, part, N, H, W, C, Ho, Wo);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

// persistent blocks: each stages its groups' weights once, then walks pixel tiles; about 4 / 4 / 2 / 1 resident blocks
// per SM (the weight tile is 4.6 / 18 / 74 / 147 KB of shared memory)
static int gconv_blocks_x(int N, int Ho, int Wo, int cpg, int C) {
    const int ppt = (cpg <= 8) ? 4 : 2;
    const int gl = cpg == 32 ? 4 : 8;
    const int spt = 8 * (32 / gl);
    const long strips = static_cast<long>(N) * Ho * ((Wo + ppt - 1) / ppt);
    long tiles = (strips + spt - 1) / spt;
    const int per_sm = cpg <= 8 ? 4 : (cpg == 16 ? 2 : 1);
    long cap = (148L * per_sm) / (C / (gl * cpg));
    if (cap < 1) cap = 1;
    if (tiles > cap) tiles = cap;
    return static_cast<int>(tiles < 1 ? 1 : tiles);
}

}  // namespace b200lp

using namespace b200lp;

extern "C" int32_t b200lp_col_stats_parts(int64_t M) {
    if (M <= 0) return B200LP_EINVAL;
    int rpc;
    return row_chunks(M, &rpc);
}

extern "C" int32_t b200lp_col_stats(const float* x, float* part, int64_t M, int32_t C, void* stream) {
    B200LP_REQUIRE(x && part && M > 0 && C > 0 && C % 4 == 0, "col_stats: bad args M=%lld C=%d", (long long)M, C);
    int rpc;
    const int chunks = row_chunks(M, &rpc);
    col_stats_kernel<<<chunks, 256, 0, as_stream(stream)>>>(x, part, M, C, rpc);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_bn_act(const float* x, const float* scale, const float* shift, const float* res,
                                 const float* res_scale, const float* res_shift, float* y, void* y_split, int64_t M,
                                 int32_t C, int32_t act, int32_t round_tf32, uint8_t* mask_out, void* stream) {
    B200LP_REQUIRE(x && (y || y_split) && M > 0 && C > 0 && C % 4 == 0 && act >= 0 && act <= 2, "bn_act: bad args");
    B200LP_REQUIRE((scale == nullptr) == (shift == nullptr) && (res_scale == nullptr) == (res_shift == nullptr) &&
                       (res || !res_scale), "bn_act: scale / shift go together (res_scale needs res)");
    const long total4 = M * (C / 4);
    bn_act_kernel<<<ew_blocks(total4), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(x), scale, shift, reinterpret_cast<const float4*>(res), res_scale, res_shift,
        reinterpret_cast<float4*>(y), static_cast<__nv_bfloat16*>(y_split), static_cast<long long>(M) * C, total4, C / 4,
        act, round_tf32, mask_out);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_bn_bwd(const float* dy, const void* mask_src_v, const float* x_raw, const float* mean,
                                 const float* rstd, const float* scale, const float* shift, const float* gamma,
                                 float* dgamma, float* dbeta, int32_t accumulate, float* dx, float* dz_out, float* workspace,
                                 int64_t workspace_bytes, int64_t M, int32_t C, int32_t mask_mode, int32_t batch_stats,
                                 int32_t round_tf32, void* stream) {
    const float* mask_src = static_cast<const float*>(mask_src_v);
    B200LP_REQUIRE(dy && x_raw && mean && rstd && gamma && dx && workspace && M > 0 && C > 0 && C % 4 == 0 &&
                       mask_mode >= 0 && mask_mode <= 4, "bn_bwd: bad args");
    B200LP_REQUIRE((mask_mode != 1 && mask_mode != 4) || mask_src, "bn_bwd: mask_mode 1 / 4 needs mask_src");
    B200LP_REQUIRE(mask_mode < 2 || mask_mode == 4 || (scale && shift), "bn_bwd: mask_mode 2/3 needs scale and shift");
    int rpc;
    const int chunks = row_chunks(M, &rpc);
    const int64_t need = (static_cast<int64_t>(chunks) * 2 * C + 2 * C) * 4;
    B200LP_REQUIRE(workspace_bytes >= need, "bn_bwd: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
    cudaStream_t st = as_stream(stream);
    float* part = workspace;
    float* coef = workspace + static_cast<size_t>(chunks) * 2 * C;
    bn_bwd_reduce_kernel<<<chunks, 256, 0, st>>>(dy, mask_src, x_raw, mean, rstd, scale, shift, part, M, C, rpc, mask_mode);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    bn_bwd_finalize_kernel<<<(C + 31) / 32, 1024, 0, st>>>(part, chunks, static_cast<double>(M), dgamma, dbeta, accumulate,
                                                           coef, C, batch_stats);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    const long total4 = M * (C / 4);
    bn_bwd_apply_kernel<<<ew_blocks(total4), 256, 0, st>>>(dy, mask_src, x_raw, mean, rstd, scale, shift, gamma, coef, dx,
                                                           dz_out, total4, C, mask_mode, round_tf32);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int64_t b200lp_bn_bwd_workspace(int64_t M, int32_t C) {
    if (M <= 0 || C <= 0) return B200LP_EINVAL;
    int rpc;
    const int chunks = row_chunks(M, &rpc);
    return (static_cast<int64_t>(chunks) * 2 * C + 2 * C) * 4;
}

extern "C" int32_t b200lp_gconv3x3_parts(int32_t N, int32_t H, int32_t W, int32_t C, int32_t cpg, int32_t stride) {
    if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || (stride != 1 && stride != 2) ||
        !(cpg == 4 || cpg == 8 || cpg == 16 || cpg == 32) || C % (8 * cpg))
        return B200LP_EINVAL;
    return gconv_blocks_x(N, (H - 1) / stride + 1, (W - 1) / stride + 1, cpg, C);
}

extern "C" int32_t b200lp_gconv3x3_fwd(const float* x, const float* in_scale, const float* in_shift, const float* w,
                                       float* y, float* part, int32_t N, int32_t H, int32_t W, int32_t C, int32_t cpg,
                                       int32_t stride, int32_t transposed, void* stream) {
    B200LP_REQUIRE(x && w && y && N > 0 && H > 0 && W > 0 && C > 0, "gconv3x3_fwd: bad args");
    B200LP_REQUIRE((cpg == 4 || cpg == 8 || cpg == 16 || cpg == 32) && C % (8 * cpg) == 0,
                   "gconv3x3_fwd: channels per group %d not in {4,8,16,32} or C=%d not a multiple of 8 groups", cpg, C);
    B200LP_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), "gconv3x3_fwd: in_scale and in_shift go together");
    B200LP_REQUIRE(stride == 1 || (stride == 2 && !transposed), "gconv3x3_fwd: stride %d (transposed %d) unsupported",
                   stride, transposed);
    B200LP_REQUIRE(static_cast<long>(N) * H * W < (1L << 31), "gconv3x3_fwd: more than 2^31 pixels");
    const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    const int bx = gconv_blocks_x(N, Ho, Wo, cpg, C);
    cudaStream_t st = as_stream(stream);
#define B200LP_GCONV_CASE(CPG)                                                                                         \
    if (cpg == CPG) {                                                                                                  \
        if (transposed) return launch_gconv<CPG, 1, true>(x, in_scale, in_shift, w, y, part, N, H, W, C, Ho, Wo, bx, st); \
        if (stride == 1) return launch_gconv<CPG, 1, false>(x, in_scale, in_shift, w, y, part, N, H, W, C, Ho, Wo, bx, st); \
        return launch_gconv<CPG, 2, false>(x, in_scale, in_shift, w, y, part, N, H, W, C, Ho, Wo, bx, st);              \
    }
    B200LP_GCONV_CASE(4) B200LP_GCONV_CASE(8) B200LP_GCONV_CASE(16) B200LP_GCONV_CASE(32)
#undef B200LP_GCONV_CASE
    return B200LP_EINVAL;
}

template <int CPG>
static int launch_gconv_dgrad_s2(const float* dy, const float* w, float* dx, int N, int H, int W, int C, int Ho, int Wo,
                                 cudaStream_t st) {
    constexpr int kBytes = 9 * GconvCfg<CPG>::CI4C * CPG * 8 * 16;
    static bool attr_set = false;
    if (!attr_set) {
        B200LP_CHECK_CUDA(cudaFuncSetAttribute(gconv3x3_dgrad_s2_kernel<CPG>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBytes));
        attr_set = true;
    }
    const long P = static_cast<long>(N) * H * W;
    long bx = (P + 31) / 32;
    const long cap = (148L * 8) / (C / (8 * CPG));
    if (bx > cap) bx = cap;
    dim3 grid(static_cast<unsigned>(bx < 1 ? 1 : bx), C / (8 * CPG));
    gconv3x3_dgrad_s2_kernel<CPG><<<grid, 256, kBytes, st>>>(dy, w, dx, N, H, W, C, Ho, Wo);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_gconv3x3_dgrad(const float* dy, const float* w, float* dx, int32_t N, int32_t H, int32_t W,
                                         int32_t C, int32_t cpg, int32_t stride, void* stream) {
    // (N, H, W) = shape of dx (the forward INPUT); dy is (N, Ho, Wo, C)
    if (stride == 1) return b200lp_gconv3x3_fwd(dy, nullptr, nullptr, w, dx, nullptr, N, H, W, C, cpg, 1, 1, stream);
    B200LP_REQUIRE(dy && w && dx && N > 0 && H > 0 && W > 0 && stride == 2, "gconv3x3_dgrad: bad args");
    B200LP_REQUIRE(static_cast<long>(N) * H * W < (1L << 31), "gconv3x3_dgrad: more than 2^31 pixels");
    B200LP_REQUIRE((cpg == 4 || cpg == 8 || cpg == 16 || cpg == 32) && C % (8 * cpg) == 0, "gconv3x3_dgrad: bad cpg %d / C %d", cpg, C);
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    cudaStream_t st = as_stream(stream);
    if (cpg == 4) return launch_gconv_dgrad_s2<4>(dy, w, dx, N, H, W, C, Ho, Wo, st);
    if (cpg == 8) return launch_gconv_dgrad_s2<8>(dy, w, dx, N, H, W, C, Ho, Wo, st);
    if (cpg == 16) return launch_gconv_dgrad_s2<16>(dy, w, dx, N, H, W, C, Ho, Wo, st);
    return launch_gconv_dgrad_s2<32>(dy, w, dx, N, H, W, C, Ho, Wo, st);
}

extern "C" int64_t b200lp_gconv3x3_wgrad_workspace(int32_t N, int32_t H, int32_t W, int32_t C, int32_t cpg,
                                                   int32_t stride) {
    if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || (stride != 1 && stride != 2) ||
        !(cpg == 4 || cpg == 8 || cpg == 16 || cpg == 32) || C % (8 * cpg))
        return B200LP_EINVAL;
    const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    int ppc, ts, slices;
    const int chunks = gconv_wgrad_plan(N * Ho, C, cpg, &ppc, &ts, &slices);
    const int cg = chunks < 32 ? chunks : 32;
    return static_cast<int64_t>(chunks + cg) * 9 * C * cpg * 4;       // chunk partials + the stage-1 rows of the reduction
}

template <int CPG>
static int launch_gconv_wgrad(const float* x, const float* sc, const float* sh, const float* dy, float* ws, int N, int H,
                              int W, int C, int stride, int Ho, int Wo, int chunks, int ppc, int ts, int slices,
                              cudaStream_t st) {
    const int lanes = 256 / ts;
    const int bytes = (lanes > 1 ? (lanes - 1) * ts * 12 * 16 : 16);
    static bool attr_set = false;
    if (!attr_set) {
        B200LP_CHECK_CUDA(cudaFuncSetAttribute(gconv3x3_wgrad_kernel<CPG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
        attr_set = true;
    }
    dim3 grid(chunks, slices);
    gconv3x3_wgrad_kernel<CPG><<<grid, lanes * ts, bytes, st>>>(x, sc, sh, dy, ws, N, H, W, C, stride, Ho, Wo, ppc, ts);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_gconv3x3_wgrad(const float* x, const float* in_scale, const float* in_shift, const float* dy,
                                         float* dw, int32_t accumulate, float* workspace, int64_t workspace_bytes,
                                         int32_t N, int32_t H, int32_t W, int32_t C, int32_t cpg, int32_t stride,
                                         void* stream) {
    B200LP_REQUIRE(x && dy && dw && workspace, "gconv3x3_wgrad: null pointer");
    const int64_t need = b200lp_gconv3x3_wgrad_workspace(N, H, W, C, cpg, stride);
    B200LP_REQUIRE(need > 0, "gconv3x3_wgrad: unsupported shape N=%d H=%d W=%d C=%d cpg=%d stride=%d", N, H, W, C, cpg, stride);
    B200LP_REQUIRE(workspace_bytes >= need, "gconv3x3_wgrad: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
    B200LP_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), "gconv3x3_wgrad: in_scale and in_shift go together");
    B200LP_REQUIRE(static_cast<long>(N) * H * W < (1L << 31), "gconv3x3_wgrad: more than 2^31 pixels");
    const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    int ppc, ts, slices;
    const int chunks = gconv_wgrad_plan(N * Ho, C, cpg, &ppc, &ts, &slices);
    cudaStream_t st = as_stream(stream);
    int r;
    if (cpg == 4) r = launch_gconv_wgrad<4>(x, in_scale, in_shift, dy, workspace, N, H, W, C, stride, Ho, Wo, chunks, ppc, ts, slices, st);
    else if (cpg == 8) r = launch_gconv_wgrad<8>(x, in_scale, in_shift, dy, workspace, N, H, W, C, stride, Ho, Wo, chunks, ppc, ts, slices, st);
    else if (cpg == 16) r = launch_gconv_wgrad<16>(x, in_scale, in_shift, dy, workspace, N, H, W, C, stride, Ho, Wo, chunks, ppc, ts, slices, st);
    else r = launch_gconv_wgrad<32>(x, in_scale, in_shift, dy, workspace, N, H, W, C, stride, Ho, Wo, chunks, ppc, ts, slices, st);
    if (r) return r;
    const unsigned total = static_cast<unsigned>(C) * cpg * 9;
    const int cg = chunks < 32 ? chunks : 32;
    float* ws2 = workspace + static_cast<size_t>(chunks) * total;
    dim3 g1((total + 255) / 256, cg);
    gconv3x3_wgrad_reduce1_kernel<<<g1, 256, 0, st>>>(workspace, ws2, chunks, cg, total);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    gconv3x3_wgrad_reduce2_kernel<<<ew_blocks((total + 3) / 4), 256, 0, st>>>(ws2, dw, cg, C, cpg, accumulate);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_im2col7x7_s2(const float* x_nchw, float* col, void* col_split, int32_t N, int32_t H, int32_t W,
                                       int32_t KP, void* stream) {
    B200LP_REQUIRE(x_nchw && (col || col_split) && N > 0 && H > 0 && W > 0 && KP >= 148 && KP % 4 == 0,
                   "im2col7x7_s2: bad args (KP=%d must be a multiple of 4 >= 148)", KP);
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;      // kernel 7, padding 3, stride 2
    const long total4 = static_cast<long>(N) * Ho * Wo * (KP / 4);
    im2col7x7_s2_kernel<<<ew_blocks(total4), 256, 0, as_stream(stream)>>>(
        x_nchw, col, static_cast<__nv_bfloat16*>(col_split), static_cast<long long>(N) * Ho * Wo * KP, N, H, W, Ho, Wo, KP);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_maxpool3x3s2_fwd(const float* x, const float* scale, const float* shift, float* y, void* y_split,
                                           uint8_t* idx, int32_t N, int32_t H, int32_t W, int32_t C, int32_t round_tf32,
                                           void* stream) {
    B200LP_REQUIRE(x && scale && shift && (y || y_split) && N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0,
                   "maxpool3x3s2_fwd: bad args");
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long total4 = static_cast<long>(N) * Ho * Wo * (C / 4);
    maxpool3x3s2_fwd_kernel<<<ew_blocks(total4), 256, 0, as_stream(stream)>>>(
        x, scale, shift, y, static_cast<__nv_bfloat16*>(y_split), static_cast<long long>(N) * Ho * Wo * C, idx, N, H, W, C, Ho,
        Wo, round_tf32);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_maxpool3x3s2_bwd(const float* dy, const uint8_t* idx, float* dx, int32_t N, int32_t H, int32_t W,
                                           int32_t C, void* stream) {
    B200LP_REQUIRE(dy && idx && dx && N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "maxpool3x3s2_bwd: bad args");
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long total4 = static_cast<long>(N) * H * W * (C / 4);
    maxpool3x3s2_bwd_kernel<<<ew_blocks(total4), 256, 0, as_stream(stream)>>>(dy, idx, dx, N, H, W, C, Ho, Wo);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_subsample2(const float* x, const void* x_split, float* y, void* y_split, int32_t N, int32_t Ho,
                                     int32_t Wo, int32_t C, void* stream) {
    B200LP_REQUIRE(((x && y) || (x_split && y_split)) && N > 0 && Ho > 0 && Wo > 0 && C > 0 && C % 4 == 0, "subsample2: bad args");
    B200LP_REQUIRE((x == nullptr) == (y == nullptr) && (x_split == nullptr) == (y_split == nullptr),
                   "subsample2: input / output planes go together");
    const long total4 = static_cast<long>(N) * Ho * Wo * (C / 4);
    subsample2_kernel<<<ew_blocks(total4), 256, 0, as_stream(stream)>>>(
        x, static_cast<const __nv_bfloat16*>(x_split), static_cast<long long>(N) * 4 * Ho * Wo * C, y,
        static_cast<__nv_bfloat16*>(y_split), static_cast<long long>(N) * Ho * Wo * C, N, Ho, Wo, C);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_scatter_add2(const float* dsub, float* dx, int32_t N, int32_t Ho, int32_t Wo, int32_t C,
                                       void* stream) {
    B200LP_REQUIRE(dsub && dx && N > 0 && Ho > 0 && Wo > 0 && C > 0 && C % 4 == 0, "scatter_add2: bad args");
    const long total4 = static_cast<long>(N) * Ho * Wo * (C / 4);
    scatter_add2_kernel<<<ew_blocks(total4), 256, 0, as_stream(stream)>>>(dsub, dx, N, Ho, Wo, C);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_avgpool_fwd(const float* x, float* y, int32_t N, int32_t HW, int32_t C, void* stream) {
    B200LP_REQUIRE(x && y && N > 0 && HW > 0 && C > 0 && C % 4 == 0, "avgpool_fwd: bad args");
    dim3 grid((C / 4 + 31) / 32, N);
    avgpool_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, y, HW, C);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_avgpool_bwd(const float* dy, float* dx, int32_t N, int32_t HW, int32_t C, void* stream) {
    B200LP_REQUIRE(dy && dx && N > 0 && HW > 0 && C > 0 && C % 4 == 0, "avgpool_bwd: bad args");
    const long total4 = static_cast<long>(N) * HW * (C / 4);
    avgpool_bwd_kernel<<<ew_blocks(total4), 256, 0, as_stream(stream)>>>(dy, dx, total4, HW, C / 4);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

namespace b200lp {
// M <= 8 rows (the batch of a projector / classifier layer): one thread per output column keeps the 8 row sums, the A rows
// of a 128-wide k chunk sit in shared memory (broadcast reads), B streams once — the 64 x 64 tile kernel above spends 8x
// the FMAs and A traffic on rows that do not exist.  BK: B(k, n) = B[n*ldb + k] (k contiguous), else B[k*ldb + n].
// grid = (column blocks of 256, k splits); splits > 1: raw partial sums to ws[z][M][N] (merged by the reduce kernel).
template <bool BK>
__global__ void __launch_bounds__(256)
sgemm_skinny_kernel(const float* __restrict__ A, long lda, const float* __restrict__ B, long ldb, float* __restrict__ Cm,
                    const float* __restrict__ alpha_dev, const float* __restrict__ bias, int M, int N, int K, int accumulate,
                    int k_per_split, float* __restrict__ ws) {
    __shared__ float As[8][128];
    const int n = blockIdx.x * 256 + threadIdx.x;
    const int kb = blockIdx.y * k_per_split;
    const int ke = min(kb + k_per_split, K);
    float acc[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) acc[m] = 0.f;
    for (int k0 = kb; k0 < ke; k0 += 128) {
        for (int e = threadIdx.x; e < 8 * 128; e += 256) {
            const int m = e >> 7, kk = e & 127;
            As[m][kk] = (m < M && k0 + kk < ke) ? __ldg(A + m * lda + k0 + kk) : 0.f;
        }
        __syncthreads();
        if (n < N) {
            const int kn = min(128, ke - k0);
            if (BK) {
                const float* bp = B + static_cast<size_t>(n) * ldb + k0;
                const bool vec = ((reinterpret_cast<uintptr_t>(bp) & 15) == 0);
                int kk = 0;
                if (vec)
                    for (; kk + 4 <= kn; kk += 4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + kk));
#pragma unroll
                        for (int m = 0; m < 8; ++m)
                            acc[m] = fmaf(As[m][kk], b4.x, fmaf(As[m][kk + 1], b4.y, fmaf(As[m][kk + 2], b4.z, fmaf(As[m][kk + 3], b4.w, acc[m]))));
                    }
                for (; kk < kn; ++kk) {
                    const float b = __ldg(bp + kk);
#pragma unroll
                    for (int m = 0; m < 8; ++m) acc[m] = fmaf(As[m][kk], b, acc[m]);
                }
            } else {
                const float* bp = B + static_cast<size_t>(k0) * ldb + n;
#pragma unroll 4
                for (int kk = 0; kk < kn; ++kk) {
                    const float b = __ldg(bp + static_cast<size_t>(kk) * ldb);
#pragma unroll
                    for (int m = 0; m < 8; ++m) acc[m] = fmaf(As[m][kk], b, acc[m]);
                }
            }
        }
        __syncthreads();
    }
    if (n >= N) return;
    const float alpha = alpha_dev ? __ldg(alpha_dev) : 1.f;
    for (int m = 0; m < M; ++m) {
        if (gridDim.y > 1) {
            ws[(static_cast<size_t>(blockIdx.y) * M + m) * N + n] = acc[m];
        } else {
            float* o = Cm + static_cast<size_t>(m) * N + n;
            *o = (accumulate ? *o : 0.f) + alpha * acc[m] + (bias ? __ldg(bias + n) : 0.f);
        }
    }
}

}  // namespace b200lp

static int skinny_splits(int N, int K, int* k_per_split) {
    const int nb = (N + 255) / 256;
    int splits = (2 * 148 + nb - 1) / nb;
    const int kchunks = (K + 127) / 128;
    if (splits > kchunks) splits = kchunks;
    if (splits < 1) splits = 1;
    int kps = (kchunks + splits - 1) / splits * 128;
    *k_per_split = kps;
    return (K + kps - 1) / kps;
}

static int sgemm_splits(int M, int N, int K) {
    const long tiles = static_cast<long>((M + 63) / 64) * ((N + 63) / 64);
    if (tiles >= 148 || K < 512) return 1;
    long s_ = (2 * 148 + tiles - 1) / tiles;
    if (s_ > K / 128) s_ = K / 128;
    if (s_ > 128) s_ = 128;
    return static_cast<int>(s_ < 1 ? 1 : s_);
}

extern "C" int64_t b200lp_sgemm_strided_workspace(int32_t M, int32_t N, int32_t K) {
    if (M <= 0 || N <= 0 || K <= 0) return B200LP_EINVAL;
    int sp = sgemm_splits(M, N, K);
    if (M <= 8) {
        int kps;
        const int sk = skinny_splits(N, K, &kps);
        if (sk > sp) sp = sk;
    }
    return sp > 1 ? static_cast<int64_t>(sp) * M * N * 4 : 0;
}

extern "C" int32_t b200lp_sgemm_strided(const float* A, int64_t sai, int64_t sak, const float* B, int64_t sbk, int64_t sbj,
                                        float* C, const float* alpha_dev, const float* bias, int32_t M, int32_t N, int32_t K,
                                        int32_t accumulate, float* workspace, int64_t workspace_bytes, void* stream) {
    B200LP_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "sgemm_strided: bad args");
    // measurement switch (A/B runs only): B200LP_NO_SKINNY_SGEMM=1 sends batch-sized products through the 64 x 64 tile kernel
    static const bool no_skinny = getenv("B200LP_NO_SKINNY_SGEMM") != nullptr;
    if (!no_skinny && M <= 8 && sak == 1 && (sbk == 1 || sbj == 1)) {   // skinny path: A row-major, B contiguous along k or n
        cudaStream_t st = as_stream(stream);
        int kps;
        int sp = skinny_splits(N, K, &kps);
        if (sp > 1 && (!workspace || workspace_bytes < static_cast<int64_t>(sp) * M * N * 4)) { sp = 1; kps = (K + 127) / 128 * 128; }
        dim3 grid((N + 255) / 256, sp);
        if (sbk == 1 && sbj != 1)
            sgemm_skinny_kernel<true><<<grid, 256, 0, st>>>(A, sai, B, sbj, C, alpha_dev, bias, M, N, K, accumulate, kps, workspace);
        else
            sgemm_skinny_kernel<false><<<grid, 256, 0, st>>>(A, sai, B, sbk, C, alpha_dev, bias, M, N, K, accumulate, kps, workspace);
        B200LP_CHECK_CUDA(cudaGetLastError());
        count_launch();
        if (sp > 1) {
            int blocks = (M * N + 255) / 256;
            if (blocks > 148 * 8) blocks = 148 * 8;
            sgemm_splitk_reduce_kernel<<<blocks, 256, 0, st>>>(workspace, C, alpha_dev, bias, sp, M, N, accumulate);
            B200LP_CHECK_CUDA(cudaGetLastError());
            count_launch();
        }
        return B200LP_OK;
    }
    int splits = sgemm_splits(M, N, K);
    if (splits > 1 && (!workspace || workspace_bytes < static_cast<int64_t>(splits) * M * N * 4)) splits = 1;
    int kps = (K + splits - 1) / splits;
    kps = (kps + 15) / 16 * 16;
    splits = (K + kps - 1) / kps;
    dim3 grid((N + 63) / 64, (M + 63) / 64, splits);
    cudaStream_t st = as_stream(stream);
    const bool ak = sak == 1, bk = sbk == 1 && sbj != 1;
#define B200LP_SGEMM(AKV, BKV) sgemm_strided_kernel<AKV, BKV><<<grid, 256, 0, st>>>(A, sai, sak, B, sbk, sbj, C, alpha_dev, bias, M, N, K, accumulate, kps, workspace)
    if (ak && bk) B200LP_SGEMM(true, true);
    else if (ak) B200LP_SGEMM(true, false);
    else if (bk) B200LP_SGEMM(false, true);
    else B200LP_SGEMM(false, false);
#undef B200LP_SGEMM
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    if (splits > 1) {
        int blocks = (M * N + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        sgemm_splitk_reduce_kernel<<<blocks, 256, 0, st>>>(workspace, C, alpha_dev, bias, splits, M, N, accumulate);
        B200LP_CHECK_CUDA(cudaGetLastError());
        count_launch();
    }
    return B200LP_OK;
}
