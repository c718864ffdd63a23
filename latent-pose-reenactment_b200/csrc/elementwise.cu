// HBM-bound kernels of the hot path: instance-norm statistics, AdaIN+ReLU(+2x upsample) apply and backward,
// ReLU / pooling / upsample-backward, L1 feature-loss reductions, layout conversion, bias gradients.
// All activations are NHWC fp32; every kernel streams 16-byte vectors along the channel axis (coalesced),
// keeps per-(n,c) scalars in registers, and sizes its grid as a multiple of the 148 SMs.
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace b200lp {

constexpr int kEwThreads = 256;

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ---------------------------------------------------------------------------------------------------------------
// Instance-norm statistics.  x[N][HW][C];  pass 1: each block reduces a chunk of pixels for all channels to
// (count, mean, M2) partials using shifted sums (no E[x^2]-E[x]^2 cancellation); pass 2 merges partials with
// Chan's formula in fp64 and writes mean / rstd.
// Replaces nn.InstanceNorm2d statistics (generators/common/blocks.py:12,19).
// ---------------------------------------------------------------------------------------------------------------
struct Moments { float cnt, mean, m2; };

__device__ __forceinline__ Moments merge(Moments a, Moments b) {
    if (b.cnt == 0.f) return a;
    if (a.cnt == 0.f) return b;
    Moments r;
    r.cnt = a.cnt + b.cnt;
    const float d = b.mean - a.mean;
    const float f = b.cnt / r.cnt;
    r.mean = a.mean + d * f;
    r.m2 = a.m2 + b.m2 + d * d * a.cnt * f;
    return r;
}

// block 256 = (C/4 lanes) x (256/(C/4) pixel rows)   [C/4 <= 256]; sm: [rows][C][3] floats
__device__ __forceinline__ void in_stats_partial_body(const float* __restrict__ x, float* __restrict__ part, int HW, int C,
                                                      int pix_per_chunk, int n, int chunk, int nchunks, float* sm) {
    const int cq = C >> 2;
    const int lane_c = threadIdx.x % cq;
    const int row = threadIdx.x / cq;
    const int rows = blockDim.x / cq;
    const int p0 = chunk * pix_per_chunk;
    int p1 = p0 + pix_per_chunk;
    if (p1 > HW) p1 = HW;
    const float* base = x + (static_cast<size_t>(n) * HW) * C + lane_c * 4;

    float4 k = make_float4(0, 0, 0, 0), s = k, q = k;
    float cnt = 0.f;
    if (row < rows) {
        int p = p0 + row;
        if (p < p1) {
            k = ld4(base + static_cast<size_t>(p) * C);
            cnt = 1.f;
            p += rows;
#define B200LP_ACC(v)                                                                        \
    {                                                                                        \
        const float dx = (v).x - k.x, dy = (v).y - k.y, dz = (v).z - k.z, dw = (v).w - k.w;  \
        s.x += dx; s.y += dy; s.z += dz; s.w += dw;                                          \
        q.x += dx * dx; q.y += dy * dy; q.z += dz * dz; q.w += dw * dw;                      \
        cnt += 1.f;                                                                          \
    }
            for (; p + 3 * rows < p1; p += 4 * rows) {      // four independent 16-byte loads in flight per thread
                const float4 v0 = ld4(base + static_cast<size_t>(p) * C);
                const float4 v1 = ld4(base + static_cast<size_t>(p + rows) * C);
                const float4 v2 = ld4(base + static_cast<size_t>(p + 2 * rows) * C);
                const float4 v3 = ld4(base + static_cast<size_t>(p + 3 * rows) * C);
                B200LP_ACC(v0) B200LP_ACC(v1) B200LP_ACC(v2) B200LP_ACC(v3)
            }
            for (; p < p1; p += rows) {
                const float4 v = ld4(base + static_cast<size_t>(p) * C);
                B200LP_ACC(v)
            }
#undef B200LP_ACC
        }
    }
    // to (cnt, mean, M2)
    float mean[4], m2[4];
    const float inv = cnt > 0.f ? 1.f / cnt : 0.f;
    mean[0] = k.x + s.x * inv; m2[0] = q.x - s.x * s.x * inv;
    mean[1] = k.y + s.y * inv; m2[1] = q.y - s.y * s.y * inv;
    mean[2] = k.z + s.z * inv; m2[2] = q.z - s.z * s.z * inv;
    mean[3] = k.w + s.w * inv; m2[3] = q.w - s.w * s.w * inv;
    if (row < rows) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float* d = sm + (static_cast<size_t>(row) * C + lane_c * 4 + j) * 3;
            d[0] = cnt; d[1] = mean[j]; d[2] = m2[j] > 0.f ? m2[j] : 0.f;
        }
    }
    __syncthreads();
    // merge rows: one thread per channel
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        Moments acc{0.f, 0.f, 0.f};
        for (int r = 0; r < rows; ++r) {
            const float* d = sm + (static_cast<size_t>(r) * C + c) * 3;
            acc = merge(acc, Moments{d[0], d[1], d[2]});
        }
        float* o = part + ((static_cast<size_t>(n) * nchunks + chunk) * C + c) * 3;
        o[0] = acc.cnt; o[1] = acc.mean; o[2] = acc.m2;
    }
}

// grid (chunks, N)
__global__ void in_stats_partial_kernel(const float* __restrict__ x, float* __restrict__ part, int HW, int C,
                                        int pix_per_chunk) {
    extern __shared__ float sm[];
    in_stats_partial_body(x, part, HW, C, pix_per_chunk, blockIdx.y, blockIdx.x, gridDim.x, sm);
}

// mean / rstd of one (n, c) plane from its chunk partials: lanes stride over the chunks, then a shuffle tree of Chan
// merges (fp64).  Whole warp; every lane returns the same values.  `part_n` = partials of sample n, [chunks][C][3].
__device__ __forceinline__ void in_stats_merge_warp(const float* part_n, int chunks, int C, int c, float eps, float* mean_out,
                                                    float* rstd_out) {
    const int lane = threadIdx.x & 31;
    double cnt = 0.0, mu = 0.0, m2 = 0.0;
    for (int k = lane; k < chunks; k += 32) {
        const float* d = part_n + (static_cast<size_t>(k) * C + c) * 3;
        const double bc = __ldcg(d), bm = __ldcg(d + 1), b2 = __ldcg(d + 2);
        if (bc == 0.0) continue;
        const double tot = cnt + bc;
        const double dl = bm - mu;
        mu += dl * (bc / tot);
        m2 += b2 + dl * dl * cnt * (bc / tot);
        cnt = tot;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double oc = __shfl_xor_sync(0xffffffffu, cnt, o);
        const double om = __shfl_xor_sync(0xffffffffu, mu, o);
        const double o2 = __shfl_xor_sync(0xffffffffu, m2, o);
        const double tot = cnt + oc;
        if (tot > 0.0) {
            const double dl = om - mu;
            const double nm = (cnt * mu + oc * om) / tot;
            m2 = m2 + o2 + dl * dl * (cnt * oc / tot);
            mu = nm;
            cnt = tot;
        }
    }
    const double var = cnt > 0.0 ? m2 / cnt : 0.0;   // biased variance, as nn.InstanceNorm2d
    *mean_out = static_cast<float>(mu);
    *rstd_out = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

// one warp per (n, c): lanes stride over the chunk partials, then a shuffle tree of Chan merges (fp64)
__global__ void in_stats_final_kernel(const float* __restrict__ part, float* __restrict__ mean,
                                      float* __restrict__ rstd, int chunks, int C, int NC, float eps) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // n*C + c
    const int lane = threadIdx.x & 31;
    if (i >= NC) return;
    const int n = i / C, c = i - n * C;
    double cnt = 0.0, mu = 0.0, m2 = 0.0;
    for (int k = lane; k < chunks; k += 32) {
        const float* d = part + ((static_cast<size_t>(n) * chunks + k) * C + c) * 3;
        const double bc = d[0], bm = d[1], b2 = d[2];
        if (bc == 0.0) continue;
        const double tot = cnt + bc;
        const double dl = bm - mu;
        mu += dl * (bc / tot);
        m2 += b2 + dl * dl * cnt * (bc / tot);
        cnt = tot;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double oc = __shfl_xor_sync(0xffffffffu, cnt, o);
        const double om = __shfl_xor_sync(0xffffffffu, mu, o);
        const double o2 = __shfl_xor_sync(0xffffffffu, m2, o);
        const double tot = cnt + oc;
        if (tot > 0.0) {
            const double dl = om - mu;
            // symmetric merge: every lane ends with the same value (operands enter the formula identically on both sides
            // up to the sign of dl, which is squared / multiplied by the partner's weight)
            const double nm = (cnt * mu + oc * om) / tot;
            m2 = m2 + o2 + dl * dl * (cnt * oc / tot);
            mu = nm;
            cnt = tot;
        }
    }
    if (lane == 0) {
        const double var = cnt > 0.0 ? m2 / cnt : 0.0;   // biased variance, as nn.InstanceNorm2d
        mean[i] = static_cast<float>(mu);
        rstd[i] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
}

static void stats_plan(int N, int HW, int C, int* chunks, int* pix_per_chunk) {
    const int rows = kEwThreads / (C / 4) > 0 ? kEwThreads / (C / 4) : 1;
    int want = (4 * 148 + N - 1) / N;         // ~4 blocks per SM over the whole grid
    int max_chunks = HW / (rows * 4);         // at least 4 pixels per thread
    if (max_chunks < 1) max_chunks = 1;
    if (want > max_chunks) want = max_chunks;
    if (want < 1) want = 1;
    *pix_per_chunk = (HW + want - 1) / want;
    *chunks = (HW + *pix_per_chunk - 1) / *pix_per_chunk;
}

// ---------------------------------------------------------------------------------------------------------------
// AdaIN + ReLU (+ nearest 2x upsample) apply:  y = [tf32]( relu( (x-mean)*rstd*gamma + beta ) )
// Replaces blocks.py:18-26 (AdaptiveNorm2d.forward), :73 (ReLU) and :75 (Upsample).
// grid (chunks, N), block = (C/4 lanes) x rows
// ---------------------------------------------------------------------------------------------------------------
// write one float4 as fp32 (optional) and/or as (hi, lo) bf16 planes (optional; lo plane at +split_stride elements)
__device__ __forceinline__ void store_f32_and_split(float* y, __nv_bfloat16* ys, long long split_stride, size_t off,
                                                    float4 o) {
    if (y) st4(y + off, o);
    if (ys) {
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
        *reinterpret_cast<uint2*>(ys + off) = hv;
        *reinterpret_cast<uint2*>(ys + split_stride + off) = lv;
    }
}

// mean_n / rstd_n: this sample's [C] statistics — global or shared memory (plain loads)
template <bool UP, bool ROUND>
__device__ __forceinline__ void adain_apply_body(const float* __restrict__ x, const float* mean_n, const float* rstd_n,
                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                 long affine_stride, float* __restrict__ y, __nv_bfloat16* __restrict__ ys,
                                                 long long split_stride, int H, int W, int C, int pix_per_chunk, int n,
                                                 int chunk) {
    const int cq = C >> 2;
    const int lane_c = threadIdx.x % cq;
    const int row = threadIdx.x / cq;
    const int rows = blockDim.x / cq;
    if (row >= rows) return;
    const int HW = H * W;
    const int p0 = chunk * pix_per_chunk;
    int p1 = p0 + pix_per_chunk;
    if (p1 > HW) p1 = HW;
    const int c = lane_c * 4;
    const float4 mu = *reinterpret_cast<const float4*>(mean_n + c);
    const float4 rs = *reinterpret_cast<const float4*>(rstd_n + c);
    float4 g, b;
    {
        const float* gp = gamma + n * affine_stride + c;
        const float* bp = beta + n * affine_stride + c;
        g = make_float4(__ldg(gp), __ldg(gp + 1), __ldg(gp + 2), __ldg(gp + 3));
        b = make_float4(__ldg(bp), __ldg(bp + 1), __ldg(bp + 2), __ldg(bp + 3));
    }
    const float* xb = x + static_cast<size_t>(n) * HW * C + c;
    for (int p = p0 + row; p < p1; p += rows) {
        const float4 v = ld4(xb + static_cast<size_t>(p) * C);
        float4 o;
        // same operation order as the reference: normalise, then scale, then shift
        o.x = fmaxf(((v.x - mu.x) * rs.x) * g.x + b.x, 0.f);
        o.y = fmaxf(((v.y - mu.y) * rs.y) * g.y + b.y, 0.f);
        o.z = fmaxf(((v.z - mu.z) * rs.z) * g.z + b.z, 0.f);
        o.w = fmaxf(((v.w - mu.w) * rs.w) * g.w + b.w, 0.f);
        float4 of = o;   // fp32 copy: tf32-rounded when it feeds a TF32 MMA (weight gradient); split planes keep full o
        if (ROUND) { of.x = round_tf32(o.x); of.y = round_tf32(o.y); of.z = round_tf32(o.z); of.w = round_tf32(o.w); }
        if (!UP) {
            const size_t off = (static_cast<size_t>(n) * HW + p) * C + c;
            if (y) st4(y + off, of);
            if (ys) store_f32_and_split(nullptr, ys, split_stride, off, o);
        } else {
            const int h = p / W, w = p - h * W;
            const size_t off = ((static_cast<size_t>(n) * (2 * H) + 2 * h) * (2 * W) + 2 * w) * C + c;
            const size_t row2 = static_cast<size_t>(2 * W) * C;
            if (y) { st4(y + off, of); st4(y + off + C, of); st4(y + off + row2, of); st4(y + off + row2 + C, of); }
            if (ys) {
                store_f32_and_split(nullptr, ys, split_stride, off, o);
                store_f32_and_split(nullptr, ys, split_stride, off + C, o);
                store_f32_and_split(nullptr, ys, split_stride, off + row2, o);
                store_f32_and_split(nullptr, ys, split_stride, off + row2 + C, o);
            }
        }
    }
}

template <bool UP, bool ROUND>
__global__ void adain_relu_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                  const float* __restrict__ rstd, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, long affine_stride, float* __restrict__ y,
                                  __nv_bfloat16* __restrict__ ys, long long split_stride, int H, int W, int C,
                                  int pix_per_chunk) {
    adain_apply_body<UP, ROUND>(x, mean + static_cast<size_t>(blockIdx.y) * C, rstd + static_cast<size_t>(blockIdx.y) * C, gamma,
                                beta, affine_stride, y, ys, split_stride, H, W, C, pix_per_chunk, blockIdx.y, blockIdx.x);
}

// Barrier among the `expected` CTAs that work on one sample (all co-resident: the host sizes the grid by the occupancy
// query).  ctr[0] counts arrivals, ctr[1] departures; the last CTA to leave resets both, so the pair is reusable by the next
// launch on the stream without a memset.
__device__ __forceinline__ void sample_barrier(unsigned* ctr, unsigned expected) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(ctr, 1u);
        while (*reinterpret_cast<volatile unsigned*>(ctr) < expected) __nanosleep(40);
        __threadfence();
        if (atomicAdd(ctr + 1, 1u) == expected - 1) {      // everyone has seen the full count: safe to reset
            ctr[1] = 0u;
            __threadfence();
            ctr[0] = 0u;
        }
    }
    __syncthreads();
}

// The AdaIN site as ONE launch: statistics partials -> per-sample barrier -> CTA `chunk` merges the channels chunk,
// chunk + nchunks, ... of its sample (one warp per channel; every CTA merging ALL channels in fp64 cost 3x the whole
// two-kernel form) and publishes mean / rstd -> second barrier -> apply, re-reading the pixels this very CTA just streamed
// (L2-resident unless the tensor exceeds the L2).  grid (chunks, N), all CTAs co-resident.
template <bool UP, bool ROUND>
__global__ void __launch_bounds__(kEwThreads)
adain_fused_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                   long affine_stride, float* __restrict__ y, __nv_bfloat16* __restrict__ ys, long long split_stride,
                   float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ part, unsigned* __restrict__ ctr,
                   int H, int W, int C, int pix_per_chunk, float eps) {
    extern __shared__ float sm[];          // [rows][C][3]
    const int n = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
    in_stats_partial_body(x, part, H * W, C, pix_per_chunk, n, chunk, nchunks, sm);
    sample_barrier(ctr + 4 * n, nchunks);
    const float* part_n = part + static_cast<size_t>(n) * nchunks * C * 3;
    const int warp = threadIdx.x >> 5;
    for (int c = chunk + nchunks * warp; c < C; c += nchunks * (kEwThreads / 32)) {
        float m, r;
        in_stats_merge_warp(part_n, nchunks, C, c, eps, &m, &r);
        if ((threadIdx.x & 31) == 0) { mean[static_cast<size_t>(n) * C + c] = m; rstd[static_cast<size_t>(n) * C + c] = r; }
    }
    sample_barrier(ctr + 4 * n + 2, nchunks);
    adain_apply_body<UP, ROUND>(x, mean + static_cast<size_t>(n) * C, rstd + static_cast<size_t>(n) * C, gamma, beta,
                                affine_stride, y, ys, split_stride, H, W, C, pix_per_chunk, n, chunk);
}

// Backward pass 1: per-(n,c) partial sums of dz and dz*xhat, dz = dy_in * [z > 0], where dy_in is dy summed over
// the 2x2 upsampled block when UP.  part[n][chunk][C][2]
template <bool UP>
__global__ void adain_bwd_partial_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                         const float* __restrict__ rstd, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, long affine_stride,
                                         const float* __restrict__ dy, float* __restrict__ part, int H, int W, int C,
                                         int pix_per_chunk) {
    extern __shared__ float sm[];  // [rows][C][2]
    const int cq = C >> 2;
    const int lane_c = threadIdx.x % cq;
    const int row = threadIdx.x / cq;
    const int rows = blockDim.x / cq;
    const int n = blockIdx.y;
    const int HW = H * W;
    const int p0 = blockIdx.x * pix_per_chunk;
    int p1 = p0 + pix_per_chunk;
    if (p1 > HW) p1 = HW;
    const int c = lane_c * 4;
    float sb[4] = {0, 0, 0, 0}, sg[4] = {0, 0, 0, 0};
    if (row < rows) {
        const float4 mu = ld4(mean + static_cast<size_t>(n) * C + c);
        const float4 rs = ld4(rstd + static_cast<size_t>(n) * C + c);
        const float* gp = gamma + n * affine_stride + c;
        const float* bp = beta + n * affine_stride + c;
        const float g[4] = {__ldg(gp), __ldg(gp + 1), __ldg(gp + 2), __ldg(gp + 3)};
        const float b[4] = {__ldg(bp), __ldg(bp + 1), __ldg(bp + 2), __ldg(bp + 3)};
        const float m_[4] = {mu.x, mu.y, mu.z, mu.w};
        const float r_[4] = {rs.x, rs.y, rs.z, rs.w};
        const float* xb = x + static_cast<size_t>(n) * HW * C + c;
        for (int p = p0 + row; p < p1; p += rows) {
            const float4 v4 = ld4(xb + static_cast<size_t>(p) * C);
            const float v[4] = {v4.x, v4.y, v4.z, v4.w};
            float d[4];
            if (!UP) {
                const float4 d4 = ld4(dy + (static_cast<size_t>(n) * HW + p) * C + c);
                d[0] = d4.x; d[1] = d4.y; d[2] = d4.z; d[3] = d4.w;
            } else {
                const int h = p / W, w = p - h * W;
                const float* db = dy + ((static_cast<size_t>(n) * (2 * H) + 2 * h) * (2 * W) + 2 * w) * C + c;
                const float4 a0 = ld4(db), a1 = ld4(db + C), a2 = ld4(db + static_cast<size_t>(2 * W) * C),
                             a3 = ld4(db + static_cast<size_t>(2 * W) * C + C);
                d[0] = (a0.x + a1.x) + (a2.x + a3.x);
                d[1] = (a0.y + a1.y) + (a2.y + a3.y);
                d[2] = (a0.z + a1.z) + (a2.z + a3.z);
                d[3] = (a0.w + a1.w) + (a2.w + a3.w);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float xh = (v[j] - m_[j]) * r_[j];
                const float z = xh * g[j] + b[j];
                const float dz = z > 0.f ? d[j] : 0.f;
                sb[j] += dz;
                sg[j] += dz * xh;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sm[(static_cast<size_t>(row) * C + c + j) * 2 + 0] = sb[j];
            sm[(static_cast<size_t>(row) * C + c + j) * 2 + 1] = sg[j];
        }
    }
    __syncthreads();
    for (int cc = threadIdx.x; cc < C; cc += blockDim.x) {
        float a = 0.f, g2 = 0.f;
        for (int r = 0; r < rows; ++r) {
            a += sm[(static_cast<size_t>(r) * C + cc) * 2 + 0];
            g2 += sm[(static_cast<size_t>(r) * C + cc) * 2 + 1];
        }
        float* o = part + ((static_cast<size_t>(n) * gridDim.x + blockIdx.x) * C + cc) * 2;
        o[0] = a; o[1] = g2;
    }
}

__global__ void adain_bwd_final_kernel(const float* __restrict__ part, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, int chunks, int C, int NC) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NC) return;
    const int n = i / C, c = i - n * C;
    double a = 0.0, g = 0.0;
    for (int k = 0; k < chunks; ++k) {
        const float* d = part + ((static_cast<size_t>(n) * chunks + k) * C + c) * 2;
        a += d[0]; g += d[1];
    }
    dbeta[i] = static_cast<float>(a);
    dgamma[i] = static_cast<float>(g);
}

// Backward pass 2: dx = rstd*gamma*(dz - dbeta/HW - xhat*dgamma/HW)
template <bool UP>
__global__ void adain_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                       const float* __restrict__ rstd, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, long affine_stride,
                                       const float* __restrict__ dy, const float* __restrict__ dgamma,
                                       const float* __restrict__ dbeta, float* __restrict__ dx, int H, int W, int C,
                                       int pix_per_chunk, const float* __restrict__ add, int round_out) {
    const int cq = C >> 2;
    const int lane_c = threadIdx.x % cq;
    const int row = threadIdx.x / cq;
    const int rows = blockDim.x / cq;
    if (row >= rows) return;
    const int n = blockIdx.y;
    const int HW = H * W;
    const int p0 = blockIdx.x * pix_per_chunk;
    int p1 = p0 + pix_per_chunk;
    if (p1 > HW) p1 = HW;
    const int c = lane_c * 4;
    const float4 mu = ld4(mean + static_cast<size_t>(n) * C + c);
    const float4 rs = ld4(rstd + static_cast<size_t>(n) * C + c);
    const float4 dg4 = ld4(dgamma + static_cast<size_t>(n) * C + c);
    const float4 db4 = ld4(dbeta + static_cast<size_t>(n) * C + c);
    const float* gp = gamma + n * affine_stride + c;
    const float* bp = beta + n * affine_stride + c;
    const float g[4] = {__ldg(gp), __ldg(gp + 1), __ldg(gp + 2), __ldg(gp + 3)};
    const float b[4] = {__ldg(bp), __ldg(bp + 1), __ldg(bp + 2), __ldg(bp + 3)};
    const float m_[4] = {mu.x, mu.y, mu.z, mu.w};
    const float r_[4] = {rs.x, rs.y, rs.z, rs.w};
    const float inv = 1.f / static_cast<float>(HW);
    const float mg[4] = {dg4.x * inv, dg4.y * inv, dg4.z * inv, dg4.w * inv};
    const float mb[4] = {db4.x * inv, db4.y * inv, db4.z * inv, db4.w * inv};
    const float* xb = x + static_cast<size_t>(n) * HW * C + c;
    for (int p = p0 + row; p < p1; p += rows) {
        const float4 v4 = ld4(xb + static_cast<size_t>(p) * C);
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
        float d[4];
        if (!UP) {
            const float4 d4 = ld4(dy + (static_cast<size_t>(n) * HW + p) * C + c);
            d[0] = d4.x; d[1] = d4.y; d[2] = d4.z; d[3] = d4.w;
        } else {
            const int h = p / W, w = p - h * W;
            const float* dbp = dy + ((static_cast<size_t>(n) * (2 * H) + 2 * h) * (2 * W) + 2 * w) * C + c;
            const float4 a0 = ld4(dbp), a1 = ld4(dbp + C), a2 = ld4(dbp + static_cast<size_t>(2 * W) * C),
                         a3 = ld4(dbp + static_cast<size_t>(2 * W) * C + C);
            d[0] = (a0.x + a1.x) + (a2.x + a3.x);
            d[1] = (a0.y + a1.y) + (a2.y + a3.y);
            d[2] = (a0.z + a1.z) + (a2.z + a3.z);
            d[3] = (a0.w + a1.w) + (a2.w + a3.w);
        }
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float xh = (v[j] - m_[j]) * r_[j];
            const float z = xh * g[j] + b[j];
            const float dz = z > 0.f ? d[j] : 0.f;
            o[j] = r_[j] * g[j] * (dz - mb[j] - xh * mg[j]);
        }
        const size_t off = (static_cast<size_t>(n) * HW + p) * C + c;
        if (add) {          // a second gradient of the same tensor (the block's skip branch): merged here, not by at::add
            const float4 a4 = ld4(add + off);
            o[0] += a4.x; o[1] += a4.y; o[2] += a4.z; o[3] += a4.w;
        }
        if (round_out) {    // the result only feeds tf32 MMAs (which would truncate)
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = round_tf32(o[j]);
        }
        st4(dx + off, make_float4(o[0], o[1], o[2], o[3]));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// simple streaming kernels
// ---------------------------------------------------------------------------------------------------------------
__global__ void relu_round_kernel(const float4* __restrict__ x, float4* __restrict__ y, long n4) {
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        float4 v = __ldg(x + i);
        v.x = round_tf32(fmaxf(v.x, 0.f)); v.y = round_tf32(fmaxf(v.y, 0.f));
        v.z = round_tf32(fmaxf(v.z, 0.f)); v.w = round_tf32(fmaxf(v.w, 0.f));
        y[i] = v;
    }
}
__global__ void relu_bwd_kernel(const float4* __restrict__ y, const float4* __restrict__ dy, float4* __restrict__ dx,
                                long n4) {
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const float4 m = __ldg(y + i);
        float4 d = __ldg(dy + i);
        d.x = m.x > 0.f ? d.x : 0.f; d.y = m.y > 0.f ? d.y : 0.f;
        d.z = m.z > 0.f ? d.z : 0.f; d.w = m.w > 0.f ? d.w : 0.f;
        dx[i] = d;
    }
}

// y[n,h,w,:] = scale * (sum of the 2x2 block of x) (+ addend);  x is [N,2H,2W,C], y [N,H,W,C]
template <bool ROUND>
__global__ void pool2_kernel(const float* __restrict__ x, const float* __restrict__ addend, float* __restrict__ y,
                             int H, int W, int C, long total4, float scale) {
    const int cq = C >> 2;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total4;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % cq) * 4;
        long p = i / cq;
        const int w = static_cast<int>(p % W);
        p /= W;
        const int h = static_cast<int>(p % H);
        const long n = p / H;
        const float* xb = x + ((n * (2 * H) + 2 * h) * (2L * W) + 2 * w) * C + c;
        const float4 a0 = ld4(xb), a1 = ld4(xb + C), a2 = ld4(xb + 2L * W * C), a3 = ld4(xb + 2L * W * C + C);
        float4 o;
        o.x = ((a0.x + a1.x) + (a2.x + a3.x)) * scale;
        o.y = ((a0.y + a1.y) + (a2.y + a3.y)) * scale;
        o.z = ((a0.z + a1.z) + (a2.z + a3.z)) * scale;
        o.w = ((a0.w + a1.w) + (a2.w + a3.w)) * scale;
        if (addend) {
            const float4 ad = ld4(addend + i * 4);
            o.x += ad.x; o.y += ad.y; o.z += ad.z; o.w += ad.w;
        }
        if (ROUND) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
        st4(y + i * 4, o);
    }
}

// dx[n,h,w,:] = scale * dy[n,h/2,w/2,:]   dx is [N,2H,2W,C], dy [N,H,W,C]
__global__ void unpool2_kernel(const float* __restrict__ dy, float* __restrict__ dx, int H, int W, int C,
                               long total4, float scale) {
    const int cq = C >> 2;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total4;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % cq) * 4;
        long p = i / cq;
        const int w = static_cast<int>(p % W);
        p /= W;
        const int h = static_cast<int>(p % H);
        const long n = p / H;
        float4 v = ld4(dy + i * 4);
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
        float* db = dx + ((n * (2 * H) + 2 * h) * (2L * W) + 2 * w) * C + c;
        st4(db, v); st4(db + C, v); st4(db + 2L * W * C, v); st4(db + 2L * W * C + C, v);
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// out[0] += scale * sum |a-b|
__global__ void l1_sum_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float* __restrict__ out,
                              long n4, float scale) {
    float acc = 0.f;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const float4 u = __ldg(a + i), v = __ldg(b + i);
        acc += (fabsf(u.x - v.x) + fabsf(u.y - v.y)) + (fabsf(u.z - v.z) + fabsf(u.w - v.w));
    }
    __shared__ float ws[kEwThreads / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < kEwThreads / 32 ? ws[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(out, v * scale);
    }
}
template <bool ACC>
__global__ void l1_bwd_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                              const float* __restrict__ gscale, float scale2, float4* __restrict__ da, long n4) {
    const float g = (gscale ? __ldg(gscale) : 1.f) * scale2;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const float4 u = __ldg(a + i), v = __ldg(b + i);
        float4 d;
        d.x = (u.x > v.x ? g : (u.x < v.x ? -g : 0.f));
        d.y = (u.y > v.y ? g : (u.y < v.y ? -g : 0.f));
        d.z = (u.z > v.z ? g : (u.z < v.z ? -g : 0.f));
        d.w = (u.w > v.w ? g : (u.w < v.w ? -g : 0.f));
        if (ACC) {
            const float4 o = da[i];
            d.x += o.x; d.y += o.y; d.z += o.z; d.w += o.w;
        }
        da[i] = d;
    }
}

// VGG backward tap: d_out = [a > 0] * (d_in + sign(a - b) * g)   (a = post-ReLU feature, so [a>0] is the ReLU mask)
// = l1_bwd (accumulating) followed by relu_bwd, in one pass over a, b, d.
template <bool HAS_IN>
__global__ void l1_relu_bwd_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                   const float* __restrict__ gscale, float scale2, const float4* __restrict__ d_in,
                                   float4* __restrict__ d_out, long n4) {
    const float g = (gscale ? __ldg(gscale) : 1.f) * scale2;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const float4 u = __ldg(a + i), v = __ldg(b + i);
        float4 d = HAS_IN ? __ldg(d_in + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        d.x = u.x > 0.f ? d.x + (u.x > v.x ? g : (u.x < v.x ? -g : 0.f)) : 0.f;
        d.y = u.y > 0.f ? d.y + (u.y > v.y ? g : (u.y < v.y ? -g : 0.f)) : 0.f;
        d.z = u.z > 0.f ? d.z + (u.z > v.z ? g : (u.z < v.z ? -g : 0.f)) : 0.f;
        d.w = u.w > 0.f ? d.w + (u.w > v.w ? g : (u.w < v.w ? -g : 0.f)) : 0.f;
        d_out[i] = d;
    }
}

// VGG tap, forward half of the fused L1: out[0] += scale * sum |a-b| AND one code byte per float4 for the backward pass:
// 2 bits per element — 0: a <= 0 (ReLU mask closed), 1 / 2 / 3: a > 0 and sign(a-b) = -1 / 0 / +1.  The backward tap then
// needs neither feature map (0.25 B instead of 8 B per element, and the `b` branch's activations are not kept alive).
__device__ __forceinline__ unsigned l1_code4(const float4 u, const float4 v) {
    auto cd = [](float p, float q) -> unsigned { return p > 0.f ? (p > q ? 3u : (p < q ? 1u : 2u)) : 0u; };
    return cd(u.x, v.x) | (cd(u.y, v.y) << 2) | (cd(u.z, v.z) << 4) | (cd(u.w, v.w) << 6);
}
__device__ __forceinline__ float l1_abs4(const float4 u, const float4 v) {
    return (fabsf(u.x - v.x) + fabsf(u.y - v.y)) + (fabsf(u.z - v.z) + fabsf(u.w - v.w));
}

// One thread handles 16 consecutive elements: eight 16-byte loads in flight, one 32-bit store of four code bytes.
__global__ void l1_sum_code_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float* __restrict__ out,
                                   unsigned char* __restrict__ code, long n4, float scale) {
    float acc = 0.f;
    const long n16 = n4 >> 2;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n16;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const float4 u0 = __ldg(a + 4 * i), u1 = __ldg(a + 4 * i + 1), u2 = __ldg(a + 4 * i + 2), u3 = __ldg(a + 4 * i + 3);
        const float4 v0 = __ldg(b + 4 * i), v1 = __ldg(b + 4 * i + 1), v2 = __ldg(b + 4 * i + 2), v3 = __ldg(b + 4 * i + 3);
        acc += (l1_abs4(u0, v0) + l1_abs4(u1, v1)) + (l1_abs4(u2, v2) + l1_abs4(u3, v3));
        reinterpret_cast<unsigned*>(code)[i] =
            l1_code4(u0, v0) | (l1_code4(u1, v1) << 8) | (l1_code4(u2, v2) << 16) | (l1_code4(u3, v3) << 24);
    }
    if (blockIdx.x == 0)                                          // ragged tail (n4 % 4 float4s)
        for (long i = (n16 << 2) + threadIdx.x; i < n4; i += blockDim.x) {
            const float4 u = __ldg(a + i), v = __ldg(b + i);
            acc += l1_abs4(u, v);
            code[i] = static_cast<unsigned char>(l1_code4(u, v));
        }
    __shared__ float ws[kEwThreads / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < kEwThreads / 32 ? ws[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(out, v * scale);
    }
}

// backward half: d_out = code ? tf32(d_in + (code - 2) * g) : 0      (== l1_relu_bwd on the features the code came from)
template <bool HAS_IN>
__global__ void l1_code_bwd_kernel(const unsigned char* __restrict__ code, const float* __restrict__ gscale, float scale2,
                                   const float4* __restrict__ d_in, float4* __restrict__ d_out, long n4) {
    const float g = (gscale ? __ldg(gscale) : 1.f) * scale2;
    auto one = [g](float din, unsigned cc) -> float {
        return cc ? round_tf32(din + (static_cast<float>(cc) - 2.f) * g) : 0.f;
    };
    auto quad = [&](float4 d, unsigned c) -> float4 {
        return make_float4(one(d.x, c & 3u), one(d.y, (c >> 2) & 3u), one(d.z, (c >> 4) & 3u), one(d.w, (c >> 6) & 3u));
    };
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const long n16 = n4 >> 2;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n16;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const unsigned c = __ldg(reinterpret_cast<const unsigned*>(code) + i);
        const float4 d0 = HAS_IN ? __ldg(d_in + 4 * i) : zero, d1 = HAS_IN ? __ldg(d_in + 4 * i + 1) : zero,
                     d2 = HAS_IN ? __ldg(d_in + 4 * i + 2) : zero, d3 = HAS_IN ? __ldg(d_in + 4 * i + 3) : zero;
        d_out[4 * i] = quad(d0, c & 255u);
        d_out[4 * i + 1] = quad(d1, (c >> 8) & 255u);
        d_out[4 * i + 2] = quad(d2, (c >> 16) & 255u);
        d_out[4 * i + 3] = quad(d3, c >> 24);
    }
    if (blockIdx.x == 0)
        for (long i = (n16 << 2) + threadIdx.x; i < n4; i += blockDim.x)
            d_out[i] = quad(HAS_IN ? __ldg(d_in + i) : zero, code[i]);
}

// The same tap when a 2x2 average pool follows it (VGG conv1_2 / 2_2 / 3_4 / 4_4): one thread owns a 2x2 pixel block of
// 4 channels, so the pass that already reads both feature maps also writes their pooled (tf32-rounded) versions — the
// two avgpool2 launches per tap and their 8 B/element of reads disappear.  Codes keep the flat float4 indexing.
__global__ void l1_sum_code_pool_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float* __restrict__ out,
                                        unsigned char* __restrict__ code, float4* __restrict__ ap, float4* __restrict__ bp,
                                        long total, int Ho, int Wo, int C4, float scale) {
    float acc = 0.f;
    const long rowq = 2L * Wo * C4;                     // float4s per full-resolution image row
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C4);
        long t = i / C4;
        const int wo = static_cast<int>(t % Wo);
        t /= Wo;
        const int ho = static_cast<int>(t % Ho);
        const long n = t / Ho;
        const long i00 = ((n * 2 * Ho + 2 * ho) * 2L * Wo + 2 * wo) * C4 + c;
        const long idx[4] = {i00, i00 + C4, i00 + rowq, i00 + rowq + C4};
        float4 u[4], v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { u[k] = __ldg(a + idx[k]); v[k] = __ldg(b + idx[k]); }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            acc += l1_abs4(u[k], v[k]);
            code[idx[k]] = static_cast<unsigned char>(l1_code4(u[k], v[k]));
        }
        auto pool = [](const float4 (&q)[4]) -> float4 {
            return make_float4(round_tf32(((q[0].x + q[1].x) + (q[2].x + q[3].x)) * 0.25f),
                               round_tf32(((q[0].y + q[1].y) + (q[2].y + q[3].y)) * 0.25f),
                               round_tf32(((q[0].z + q[1].z) + (q[2].z + q[3].z)) * 0.25f),
                               round_tf32(((q[0].w + q[1].w) + (q[2].w + q[3].w)) * 0.25f));
        };
        ap[i] = pool(u);
        bp[i] = pool(v);
    }
    __shared__ float ws[kEwThreads / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < kEwThreads / 32 ? ws[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(out, v * scale);
    }
}

// backward of that pair: d_out[2x2 block] = code ? tf32(0.25 * d_low + (code - 2) * g) : 0   (avgpool2_bwd + l1_code_bwd)
__global__ void l1_code_bwd_unpool_kernel(const unsigned char* __restrict__ code, const float* __restrict__ gscale,
                                          float scale2, const float4* __restrict__ d_low, float4* __restrict__ d_out,
                                          long total, int Ho, int Wo, int C4) {
    const float g = (gscale ? __ldg(gscale) : 1.f) * scale2;
    auto one = [g](float din, unsigned cc) -> float {
        return cc ? round_tf32(din + (static_cast<float>(cc) - 2.f) * g) : 0.f;
    };
    const long rowq = 2L * Wo * C4;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C4);
        long t = i / C4;
        const int wo = static_cast<int>(t % Wo);
        t /= Wo;
        const int ho = static_cast<int>(t % Ho);
        const long n = t / Ho;
        const long i00 = ((n * 2 * Ho + 2 * ho) * 2L * Wo + 2 * wo) * C4 + c;
        const long idx[4] = {i00, i00 + C4, i00 + rowq, i00 + rowq + C4};
        float4 d = __ldg(d_low + i);
        d.x *= 0.25f; d.y *= 0.25f; d.z *= 0.25f; d.w *= 0.25f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned cc = code[idx[k]];
            d_out[idx[k]] = make_float4(one(d.x, cc & 3u), one(d.y, (cc >> 2) & 3u), one(d.z, (cc >> 4) & 3u), one(d.w, cc >> 6));
        }
    }
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int HW, long total) {
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        const long p = i / C;             // n*HW + hw
        const long n = p / HW;
        const long hw = p - n * HW;
        y[i] = __ldg(x + (n * C + c) * HW + hw);
    }
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int HW, long total) {
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        const long hw = i % HW;
        const long nc = i / HW;
        const int c = static_cast<int>(nc % C);
        const long n = nc / C;
        y[i] = __ldg(x + (n * HW + hw) * C + c);
    }
}

// db[c] += sum over pixels of dy[p][c]   (db zeroed by the caller);  block = (C/4 lanes) x rows
__global__ void bias_grad_kernel(const float* __restrict__ dy, float* __restrict__ db, long pixels, int C,
                                 long pix_per_block) {
    extern __shared__ float sm[];  // [rows][C]
    const int cq = C >> 2;
    const int lane_c = threadIdx.x % cq;
    const int row = threadIdx.x / cq;
    const int rows = blockDim.x / cq;
    const long p0 = blockIdx.x * pix_per_block;
    long p1 = p0 + pix_per_block;
    if (p1 > pixels) p1 = pixels;
    float4 acc = make_float4(0, 0, 0, 0);
    if (row < rows) {
        for (long p = p0 + row; p < p1; p += rows) {
            const float4 v = ld4(dy + p * C + lane_c * 4);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        st4(sm + static_cast<size_t>(row) * C + lane_c * 4, acc);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < rows; ++r) s += sm[static_cast<size_t>(r) * C + c];
        atomicAdd(db + c, s);
    }
}

// dx[p][c] = y[p][c] > 0 ? dy[p][c] + add[p][c] : 0   (ReLU backward, optionally merging a second incoming gradient),
// optionally dq = 0.25 * dx (the gradient behind a 2x2 average pool, consumed at the low resolution) and the column sums
// db_a[c] += sum_p dx, db_b[c] += sum_p dx (bias gradients of the convolutions whose output gradient dx is).
// Same blocking as bias_grad_kernel: block = (C/4 channel quads) x rows, a pixel range per block, two rows in flight.
__global__ void __launch_bounds__(kEwThreads)
relu_bwd_fused_kernel(const float* __restrict__ y, const float* __restrict__ dy, const float* __restrict__ add,
                      float* __restrict__ dx, float* __restrict__ dq, float* __restrict__ db_a, float* __restrict__ db_b,
                      long pixels, int C, long pix_per_block, int round_out) {
    extern __shared__ float sm[];  // [rows][C]
    const int cq = C >> 2;
    const int lane_c = threadIdx.x % cq;
    const int row = threadIdx.x / cq;
    const int rows = blockDim.x / cq;
    const long p0 = blockIdx.x * pix_per_block;
    long p1 = p0 + pix_per_block;
    if (p1 > pixels) p1 = pixels;
    float4 acc = make_float4(0, 0, 0, 0);
    if (row < rows) {
        auto one = [&](long p, float4 m, float4 g, float4 a2) {
            float4 v;
            v.x = m.x > 0.f ? g.x + a2.x : 0.f; v.y = m.y > 0.f ? g.y + a2.y : 0.f;
            v.z = m.z > 0.f ? g.z + a2.z : 0.f; v.w = m.w > 0.f ? g.w + a2.w : 0.f;
            const size_t off = static_cast<size_t>(p) * C + lane_c * 4;
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;          // bias sums of the unrounded values
            if (round_out) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
            st4(dx + off, v);
            if (dq) st4(dq + off, make_float4(0.25f * v.x, 0.25f * v.y, 0.25f * v.z, 0.25f * v.w));
        };
        const float4 zero = make_float4(0, 0, 0, 0);
        long p = p0 + row;
        for (; p + rows < p1; p += 2L * rows) {
            const size_t o0 = static_cast<size_t>(p) * C + lane_c * 4, o1 = static_cast<size_t>(p + rows) * C + lane_c * 4;
            const float4 m0 = ld4(y + o0), m1 = ld4(y + o1), g0 = ld4(dy + o0), g1 = ld4(dy + o1);
            const float4 a0 = add ? ld4(add + o0) : zero, a1 = add ? ld4(add + o1) : zero;
            one(p, m0, g0, a0);
            one(p + rows, m1, g1, a1);
        }
        for (; p < p1; p += rows) {
            const size_t o0 = static_cast<size_t>(p) * C + lane_c * 4;
            one(p, ld4(y + o0), ld4(dy + o0), add ? ld4(add + o0) : zero);
        }
    }
    if (db_a == nullptr && db_b == nullptr) return;
    if (row < rows) st4(sm + static_cast<size_t>(row) * C + lane_c * 4, acc);
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float t = 0.f;
        for (int r = 0; r < rows; ++r) t += sm[static_cast<size_t>(r) * C + c];
        if (db_a) atomicAdd(db_a + c, t);
        if (db_b) atomicAdd(db_b + c, t);
    }
}

// dst[i] <- src[i] for a table of small buffers (one block per buffer): the running-average copies of BatchNorm /
// spectral-norm buffers (runners/holycow.py:106-109) are ~370 tiny tensors per step
struct CopyItem {
    void* dst;
    const void* src;
    long long nbytes;
};
__global__ void __launch_bounds__(256)
copy_multi_kernel(const CopyItem* __restrict__ table) {
    const CopyItem it = table[blockIdx.x];
    const bool words = ((reinterpret_cast<uintptr_t>(it.dst) | reinterpret_cast<uintptr_t>(it.src) |
                         static_cast<uintptr_t>(it.nbytes)) & 3) == 0;
    if (words) {
        const long long n = it.nbytes >> 2;
        const uint32_t* s = static_cast<const uint32_t*>(it.src);
        uint32_t* d = static_cast<uint32_t*>(it.dst);
        for (long long i = threadIdx.x; i < n; i += 256) d[i] = s[i];
    } else {
        const uint8_t* s = static_cast<const uint8_t*>(it.src);
        uint8_t* d = static_cast<uint8_t*>(it.dst);
        for (long long i = threadIdx.x; i < it.nbytes; i += 256) d[i] = s[i];
    }
}

static int grid_for(long work_items, int threads) {
    long b = (work_items + threads - 1) / threads;
    const long cap = 148L * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<int>(b);
}

}  // namespace b200lp

using namespace b200lp;

extern "C" int64_t b200lp_in_stats_workspace(int32_t N, int32_t HW, int32_t C) {
    if (N <= 0 || HW <= 0 || C <= 0 || C % 4 || C / 4 > kEwThreads) return B200LP_EINVAL;
    int chunks, ppc;
    stats_plan(N, HW, C, &chunks, &ppc);
    return static_cast<int64_t>(N) * chunks * C * 3 * 4;
}

extern "C" int32_t b200lp_in_stats(const float* x, float* mean, float* rstd, float* workspace,
                                   int64_t workspace_bytes, int32_t N, int32_t HW, int32_t C, float eps,
                                   void* stream) {
    B200LP_REQUIRE(x && mean && rstd && workspace, "in_stats: null pointer");
    B200LP_REQUIRE(N > 0 && HW > 0 && C > 0 && C % 4 == 0 && C / 4 <= kEwThreads, "in_stats: bad shape N=%d HW=%d C=%d",
                   N, HW, C);
    int chunks, ppc;
    stats_plan(N, HW, C, &chunks, &ppc);
    B200LP_REQUIRE(workspace_bytes >= static_cast<int64_t>(N) * chunks * C * 12, "in_stats: workspace too small");
    const int rows = kEwThreads / (C / 4);
    const size_t smem = static_cast<size_t>(rows) * C * 3 * 4;
    cudaStream_t s = as_stream(stream);
    in_stats_partial_kernel<<<dim3(chunks, N), kEwThreads, smem, s>>>(x, workspace, HW, C, ppc);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    const int NC = N * C;
    in_stats_final_kernel<<<(NC + 3) / 4, 128, 0, s>>>(workspace, mean, rstd, chunks, C, NC, eps);   // 4 warps = 4 planes / block
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_adain_relu(const float* x, const float* mean, const float* rstd, const float* gamma,
                                     const float* beta, int64_t affine_stride, float* y, void* y_split, int32_t N,
                                     int32_t H, int32_t W, int32_t C, int32_t upsample2, int32_t round_tf32,
                                     void* stream) {
    B200LP_REQUIRE(x && mean && rstd && gamma && beta && (y || y_split), "adain_relu: null pointer");
    B200LP_REQUIRE(N > 0 && H > 0 && W > 0 && C % 4 == 0 && C / 4 <= kEwThreads, "adain_relu: bad shape");
    int chunks, ppc;
    stats_plan(N, H * W, C, &chunks, &ppc);
    dim3 grid(chunks, N);
    cudaStream_t s = as_stream(stream);
    __nv_bfloat16* ys = static_cast<__nv_bfloat16*>(y_split);
    const long long split_stride = static_cast<long long>(N) * H * W * C * (upsample2 ? 4 : 1);
#define LAUNCH(UP, RD)                                                                                             \
    adain_relu_kernel<UP, RD><<<grid, kEwThreads, 0, s>>>(x, mean, rstd, gamma, beta, affine_stride, y, ys,        \
                                                          split_stride, H, W, C, ppc)
    if (upsample2) { if (round_tf32) LAUNCH(true, true); else LAUNCH(true, false); }
    else { if (round_tf32) LAUNCH(false, true); else LAUNCH(false, false); }
#undef LAUNCH
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

// chunk count of a fused (barrier) launch: the plan of the separate kernels, capped so that every CTA of the grid is
// resident at once (occupancy query of the very kernel that will run)
template <typename Kern>
static int fused_chunks(Kern kern, size_t smem, int N, int HW, int C, int* pix_per_chunk) {
    int chunks, ppc;
    stats_plan(N, HW, C, &chunks, &ppc);
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            return -1;
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kEwThreads, smem) != cudaSuccess || occ < 1) return -1;
    const int cap = (occ * num_sms) / N;
    if (cap < 1) return -1;
    if (chunks > cap) {
        ppc = (HW + cap - 1) / cap;
        chunks = (HW + ppc - 1) / ppc;
    }
    *pix_per_chunk = ppc;
    return chunks;
}

extern "C" int32_t b200lp_adain_relu_fused(const float* x, const float* gamma, const float* beta, int64_t affine_stride,
                                           float* y, void* y_split, float* mean, float* rstd, float* workspace,
                                           int64_t workspace_bytes, uint32_t* sync, int32_t N, int32_t H, int32_t W,
                                           int32_t C, float eps, int32_t upsample2, int32_t round_tf32, void* stream) {
    B200LP_REQUIRE(x && gamma && beta && (y || y_split) && mean && rstd && workspace && sync, "adain_relu_fused: null pointer");
    B200LP_REQUIRE(N > 0 && H > 0 && W > 0 && C % 4 == 0 && C / 4 <= kEwThreads, "adain_relu_fused: bad shape");
    const int rows = kEwThreads / (C / 4);
    const size_t smem = static_cast<size_t>(rows) * C * 3 * 4;
    cudaStream_t s = as_stream(stream);
    __nv_bfloat16* ys = static_cast<__nv_bfloat16*>(y_split);
    const long long split_stride = static_cast<long long>(N) * H * W * C * (upsample2 ? 4 : 1);
    int ppc = 0, chunks = 0;
#define LAUNCH(UP, RD)                                                                                                  \
    {                                                                                                                   \
        chunks = fused_chunks(adain_fused_kernel<UP, RD>, smem, N, H * W, C, &ppc);                                     \
        B200LP_REQUIRE(chunks >= 1, "adain_relu_fused: the grid cannot be made co-resident (N=%d)", N);                 \
        B200LP_REQUIRE(workspace_bytes >= static_cast<int64_t>(N) * chunks * C * 12, "adain_relu_fused: workspace too small"); \
        adain_fused_kernel<UP, RD><<<dim3(chunks, N), kEwThreads, smem, s>>>(x, gamma, beta, affine_stride, y, ys,      \
                                                                             split_stride, mean, rstd, workspace, sync, \
                                                                             H, W, C, ppc, eps);                        \
    }
    if (upsample2) { if (round_tf32) LAUNCH(true, true) else LAUNCH(true, false) }
    else { if (round_tf32) LAUNCH(false, true) else LAUNCH(false, false) }
#undef LAUNCH
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int64_t b200lp_adain_relu_bwd_workspace(int32_t N, int32_t HW, int32_t C) {
    if (N <= 0 || HW <= 0 || C <= 0 || C % 4 || C / 4 > kEwThreads) return B200LP_EINVAL;
    int chunks, ppc;
    stats_plan(N, HW, C, &chunks, &ppc);
    return static_cast<int64_t>(N) * chunks * C * 2 * 4;
}

extern "C" int32_t b200lp_adain_relu_bwd(const float* x, const float* mean, const float* rstd, const float* gamma,
                                         const float* beta, int64_t affine_stride, const float* dy, float* dx,
                                         float* dgamma, float* dbeta, float* workspace, int64_t workspace_bytes,
                                         int32_t N, int32_t H, int32_t W, int32_t C, int32_t upsample2,
                                         const float* add, int32_t round_tf32, void* stream) {
    B200LP_REQUIRE(x && mean && rstd && gamma && beta && dy && dx && dgamma && dbeta && workspace,
                   "adain_relu_bwd: null pointer");
    B200LP_REQUIRE(N > 0 && H > 0 && W > 0 && C % 4 == 0 && C / 4 <= kEwThreads, "adain_relu_bwd: bad shape");
    int chunks, ppc;
    stats_plan(N, H * W, C, &chunks, &ppc);
    B200LP_REQUIRE(workspace_bytes >= static_cast<int64_t>(N) * chunks * C * 8, "adain_relu_bwd: workspace too small");
    const int rows = kEwThreads / (C / 4);
    const size_t smem = static_cast<size_t>(rows) * C * 2 * 4;
    dim3 grid(chunks, N);
    cudaStream_t s = as_stream(stream);
    if (upsample2)
        adain_bwd_partial_kernel<true><<<grid, kEwThreads, smem, s>>>(x, mean, rstd, gamma, beta, affine_stride, dy,
                                                                      workspace, H, W, C, ppc);
    else
        adain_bwd_partial_kernel<false><<<grid, kEwThreads, smem, s>>>(x, mean, rstd, gamma, beta, affine_stride, dy,
                                                                       workspace, H, W, C, ppc);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    const int NC = N * C;
    adain_bwd_final_kernel<<<(NC + 127) / 128, 128, 0, s>>>(workspace, dgamma, dbeta, chunks, C, NC);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    if (upsample2)
        adain_bwd_apply_kernel<true><<<grid, kEwThreads, 0, s>>>(x, mean, rstd, gamma, beta, affine_stride, dy, dgamma,
                                                                 dbeta, dx, H, W, C, ppc, add, round_tf32);
    else
        adain_bwd_apply_kernel<false><<<grid, kEwThreads, 0, s>>>(x, mean, rstd, gamma, beta, affine_stride, dy,
                                                                  dgamma, dbeta, dx, H, W, C, ppc, add, round_tf32);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_relu_round(const float* x, float* y, int64_t n, void* stream) {
    B200LP_REQUIRE(x && y && n > 0 && n % 4 == 0, "relu_round: bad args");
    relu_round_kernel<<<grid_for(n / 4, kEwThreads), kEwThreads, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), n / 4);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_relu_bwd(const float* y, const float* dy, float* dx, int64_t n, void* stream) {
    B200LP_REQUIRE(y && dy && dx && n > 0 && n % 4 == 0, "relu_bwd: bad args");
    relu_bwd_kernel<<<grid_for(n / 4, kEwThreads), kEwThreads, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(y), reinterpret_cast<const float4*>(dy), reinterpret_cast<float4*>(dx), n / 4);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

// x [N,2H,2W,C] -> y [N,H,W,C]; (H, W are the OUTPUT dims)
extern "C" int32_t b200lp_avgpool2(const float* x, const float* addend, float* y, int32_t N, int32_t H, int32_t W,
                                   int32_t C, int32_t round_tf32, void* stream) {
    B200LP_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C % 4 == 0, "avgpool2: bad args");
    const long total4 = static_cast<long>(N) * H * W * (C / 4);
    const int g = grid_for(total4, kEwThreads);
    if (round_tf32)
        pool2_kernel<true><<<g, kEwThreads, 0, as_stream(stream)>>>(x, addend, y, H, W, C, total4, 0.25f);
    else
        pool2_kernel<false><<<g, kEwThreads, 0, as_stream(stream)>>>(x, addend, y, H, W, C, total4, 0.25f);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

// dy [N,H,W,C] -> dx [N,2H,2W,C]
extern "C" int32_t b200lp_avgpool2_bwd(const float* dy, float* dx, int32_t N, int32_t H, int32_t W, int32_t C,
                                       void* stream) {
    B200LP_REQUIRE(dy && dx && N > 0 && H > 0 && W > 0 && C % 4 == 0, "avgpool2_bwd: bad args");
    const long total4 = static_cast<long>(N) * H * W * (C / 4);
    unpool2_kernel<<<grid_for(total4, kEwThreads), kEwThreads, 0, as_stream(stream)>>>(dy, dx, H, W, C, total4, 0.25f);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

// dy [N,2H,2W,C] -> dx [N,H,W,C]  (H, W are the low-resolution dims)
extern "C" int32_t b200lp_upsample2_bwd(const float* dy, float* dx, int32_t N, int32_t H, int32_t W, int32_t C,
                                        void* stream) {
    B200LP_REQUIRE(dy && dx && N > 0 && H > 0 && W > 0 && C % 4 == 0, "upsample2_bwd: bad args");
    const long total4 = static_cast<long>(N) * H * W * (C / 4);
    pool2_kernel<false><<<grid_for(total4, kEwThreads), kEwThreads, 0, as_stream(stream)>>>(dy, nullptr, dx, H, W, C,
                                                                                          total4, 1.0f);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_l1_sum(const float* a, const float* b, float* out, int64_t n, float scale, void* stream) {
    B200LP_REQUIRE(a && b && out && n > 0 && n % 4 == 0, "l1_sum: bad args");
    int g = grid_for(n / 4, kEwThreads);
    if (g > 148 * 4) g = 148 * 4;
    l1_sum_kernel<<<g, kEwThreads, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(a),
                                                          reinterpret_cast<const float4*>(b), out, n / 4, scale);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_l1_sum_code(const float* a, const float* b, float* out, uint8_t* code, int64_t n, float scale,
                                      void* stream) {
    B200LP_REQUIRE(a && b && out && code && n > 0 && n % 4 == 0, "l1_sum_code: bad args");
    int g = grid_for(n / 16 + 1, kEwThreads);          // 16 elements per thread
    if (g > 148 * 4) g = 148 * 4;
    l1_sum_code_kernel<<<g, kEwThreads, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(a),
                                                               reinterpret_cast<const float4*>(b), out, code, n / 4, scale);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_l1_code_bwd(const uint8_t* code, const float* gscale, float scale2, const float* d_in,
                                      float* d_out, int64_t n, void* stream) {
    B200LP_REQUIRE(code && d_out && n > 0 && n % 4 == 0, "l1_code_bwd: bad args");
    const int g = grid_for(n / 16 + 1, kEwThreads);    // 16 elements per thread
    if (d_in)
        l1_code_bwd_kernel<true><<<g, kEwThreads, 0, as_stream(stream)>>>(code, gscale, scale2,
                                                                          reinterpret_cast<const float4*>(d_in),
                                                                          reinterpret_cast<float4*>(d_out), n / 4);
    else
        l1_code_bwd_kernel<false><<<g, kEwThreads, 0, as_stream(stream)>>>(code, gscale, scale2, nullptr,
                                                                           reinterpret_cast<float4*>(d_out), n / 4);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_l1_sum_code_pool(const float* a, const float* b, float* out, uint8_t* code, float* a_pool,
                                           float* b_pool, int32_t N, int32_t H, int32_t W, int32_t C, float scale,
                                           void* stream) {
    B200LP_REQUIRE(a && b && out && code && a_pool && b_pool && N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 &&
                       C > 0 && C % 4 == 0, "l1_sum_code_pool: bad args (H, W even, C %% 4 == 0)");
    const long total = static_cast<long>(N) * (H / 2) * (W / 2) * (C / 4);
    int g = grid_for(total, kEwThreads);
    if (g > 148 * 8) g = 148 * 8;
    l1_sum_code_pool_kernel<<<g, kEwThreads, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), out, code, reinterpret_cast<float4*>(a_pool),
        reinterpret_cast<float4*>(b_pool), total, H / 2, W / 2, C / 4, scale);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_l1_code_bwd_unpool(const uint8_t* code, const float* gscale, float scale2, const float* d_low,
                                             float* d_out, int32_t N, int32_t H, int32_t W, int32_t C, void* stream) {
    B200LP_REQUIRE(code && d_low && d_out && N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 4 == 0,
                   "l1_code_bwd_unpool: bad args (H, W even, C %% 4 == 0)");
    const long total = static_cast<long>(N) * (H / 2) * (W / 2) * (C / 4);
    l1_code_bwd_unpool_kernel<<<grid_for(total, kEwThreads), kEwThreads, 0, as_stream(stream)>>>(
        code, gscale, scale2, reinterpret_cast<const float4*>(d_low), reinterpret_cast<float4*>(d_out), total, H / 2, W / 2,
        C / 4);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_l1_bwd(const float* a, const float* b, const float* gscale, float scale2, float* da,
                                 int64_t n, int32_t accumulate, void* stream) {
    B200LP_REQUIRE(a && b && da && n > 0 && n % 4 == 0, "l1_bwd: bad args");
    const int g = grid_for(n / 4, kEwThreads);
    if (accumulate)
        l1_bwd_kernel<true><<<g, kEwThreads, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(a),
                                                                     reinterpret_cast<const float4*>(b), gscale, scale2,
                                                                     reinterpret_cast<float4*>(da), n / 4);
    else
        l1_bwd_kernel<false><<<g, kEwThreads, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(a),
                                                                      reinterpret_cast<const float4*>(b), gscale,
                                                                      scale2, reinterpret_cast<float4*>(da), n / 4);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_l1_relu_bwd(const float* a, const float* b, const float* gscale, float scale2,
                                      const float* d_in, float* d_out, int64_t n, void* stream) {
    B200LP_REQUIRE(a && b && d_out && n > 0 && n % 4 == 0, "l1_relu_bwd: bad args");
    const int g = grid_for(n / 4, kEwThreads);
    if (d_in)
        l1_relu_bwd_kernel<true><<<g, kEwThreads, 0, as_stream(stream)>>>(
            reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), gscale, scale2,
            reinterpret_cast<const float4*>(d_in), reinterpret_cast<float4*>(d_out), n / 4);
    else
        l1_relu_bwd_kernel<false><<<g, kEwThreads, 0, as_stream(stream)>>>(
            reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), gscale, scale2, nullptr,
            reinterpret_cast<float4*>(d_out), n / 4);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_nchw_to_nhwc(const float* x, float* y, int32_t N, int32_t C, int32_t H, int32_t W,
                                       void* stream) {
    B200LP_REQUIRE(x && y && N > 0 && C > 0 && H > 0 && W > 0, "nchw_to_nhwc: bad args");
    const long total = static_cast<long>(N) * C * H * W;
    nchw_to_nhwc_kernel<<<grid_for(total, kEwThreads), kEwThreads, 0, as_stream(stream)>>>(x, y, C, H * W, total);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_nhwc_to_nchw(const float* x, float* y, int32_t N, int32_t C, int32_t H, int32_t W,
                                       void* stream) {
    B200LP_REQUIRE(x && y && N > 0 && C > 0 && H > 0 && W > 0, "nhwc_to_nchw: bad args");
    const long total = static_cast<long>(N) * C * H * W;
    nhwc_to_nchw_kernel<<<grid_for(total, kEwThreads), kEwThreads, 0, as_stream(stream)>>>(x, y, C, H * W, total);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_copy_multi(const void* table_dev, int32_t count, void* stream) {
    B200LP_REQUIRE(table_dev && count > 0, "copy_multi: bad args");
    copy_multi_kernel<<<count, 256, 0, as_stream(stream)>>>(static_cast<const CopyItem*>(table_dev));
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

namespace b200lp {
// wide, short matrices (the 13056-column projector output of a batch of 8): one thread per column walks the rows
__global__ void __launch_bounds__(256)
bias_grad_wide_kernel(const float* __restrict__ dy, float* __restrict__ db, int rows, int C, int accumulate) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += __ldg(dy + static_cast<size_t>(r) * C + c);
    db[c] = (accumulate ? db[c] : 0.f) + s;
}
}  // namespace b200lp

static int32_t bias_grad_impl(const float* dy, float* db, int64_t pixels, int32_t C, int32_t accumulate, void* stream);

extern "C" int32_t b200lp_bias_grad(const float* dy, float* db, int64_t pixels, int32_t C, void* stream) {
    return bias_grad_impl(dy, db, pixels, C, 0, stream);
}

extern "C" int32_t b200lp_bias_grad_acc(const float* dy, float* db, int64_t pixels, int32_t C, void* stream) {
    return bias_grad_impl(dy, db, pixels, C, 1, stream);
}

extern "C" int32_t b200lp_relu_bwd_fused(const float* y, const float* dy, const float* add, float* dx, float* dq,
                                         float* db_a, float* db_b, int64_t pixels, int32_t C, int32_t round_tf32,
                                         void* stream) {
    B200LP_REQUIRE(y && dy && dx && pixels > 0 && C > 0 && C % 4 == 0 && C / 4 <= kEwThreads,
                   "relu_bwd_fused: bad args (C=%d must be a multiple of 4, <= %d)", C, 4 * kEwThreads);
    const int rows = kEwThreads / (C / 4);
    long blocks = 148 * 4;
    long ppb = (pixels + blocks - 1) / blocks;
    if (ppb < 2L * rows) ppb = 2L * rows;
    blocks = (pixels + ppb - 1) / ppb;
    const size_t smem = (db_a || db_b) ? static_cast<size_t>(rows) * C * 4 : 0;
    relu_bwd_fused_kernel<<<static_cast<int>(blocks), kEwThreads, smem, as_stream(stream)>>>(y, dy, add, dx, dq, db_a, db_b,
                                                                                             pixels, C, ppb, round_tf32);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

static int32_t bias_grad_impl(const float* dy, float* db, int64_t pixels, int32_t C, int32_t accumulate, void* stream) {
    B200LP_REQUIRE(dy && db && pixels > 0 && C > 0, "bias_grad: bad args");
    cudaStream_t s = as_stream(stream);
    if (C % 4 != 0 || C / 4 > kEwThreads) {
        B200LP_REQUIRE(pixels <= 4096, "bias_grad: C=%d needs the wide kernel, which is for short matrices (rows=%lld)", C,
                       (long long)pixels);
        bias_grad_wide_kernel<<<(C + 255) / 256, 256, 0, s>>>(dy, db, static_cast<int>(pixels), C, accumulate);
        B200LP_CHECK_CUDA(cudaGetLastError());
        count_launch();
        return B200LP_OK;
    }
    if (!accumulate) B200LP_CHECK_CUDA(cudaMemsetAsync(db, 0, static_cast<size_t>(C) * 4, s));
    const int rows = kEwThreads / (C / 4);
    long blocks = 148 * 4;
    long ppb = (pixels + blocks - 1) / blocks;
    if (ppb < rows) ppb = rows;
    blocks = (pixels + ppb - 1) / ppb;
    bias_grad_kernel<<<static_cast<int>(blocks), kEwThreads, static_cast<size_t>(rows) * C * 4, s>>>(dy, db, pixels, C,
                                                                                                   ppb);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}
