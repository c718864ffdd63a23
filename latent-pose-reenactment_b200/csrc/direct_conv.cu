// CUDA-core direct convolutions for the two degenerate, HBM-bound shapes of the hot path:
//   * 3x3 with Cin = 3 (image stems: discriminator down_block.0, VGG features.0)  — K = 27, not a tensor-core shape
//   * the generator tail 3x3 Cin -> 4 with tanh / rgb*segm composition            — N = 4
// plus their data / weight gradients.  Image-side tensors are NCHW (the plugin boundary), feature side NHWC.
#include "common.cuh"
#include "ptx.cuh"

namespace b200lp {

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ------------------------------------------------------------------------------------------------ Cin = 3 forward
// Thread = 4 horizontally adjacent pixels x 8 output channels; block 256 threads = (256/groups pixel quads) x
// (Cout/8 channel groups), group index fastest.  The `groups` lanes of one pixel hold interleaved 4-channel quads (thread
// g owns channels q*4*groups + 4g .. +3, q = 0, 1): one store instruction writes 16*groups contiguous bytes per pixel
// (whole 128-byte lines at Cout = 64) and one weight read is a contiguous 16*groups-byte broadcast.
// Why 4 pixels per thread: every FMA needs a weight from shared memory; with one pixel per thread the kernel ran at the
// LDS.128 rate (108 broadcast loads per thread-pixel, 4 clk each per warp: 97 us of the measured 103 us for
// 8 x 256^2 x 64), 4x the FMA time.  Four pixels share each weight load and the 3 x 6 input window.
// Why 8 channels: 4 x 16 accumulators needed 255 registers = 8 warps per SM, and ncu showed the issue slots 38 % busy
// (latency-bound, 117 us); 4 x 8 keeps the same loads-per-FMA ratio at half the registers.
// Blocks are persistent over pixel chunks: the 27 x Cout weights are staged in shared memory once per block.
__global__ void __launch_bounds__(256, 2)
conv3x3_c3_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ wscale,
                      const float* __restrict__ bias, const float* __restrict__ pre_scale,
                      const float* __restrict__ pre_shift, float* __restrict__ y, int N, int H, int W, int Cout,
                      int relu, int round_out) {
    extern __shared__ float sm[];           // [27][Cout] weights, then [Cout] bias
    float* sw = sm;
    float* sb = sm + 27 * Cout;
    const float s = wscale ? __ldg(wscale) : 1.f;
    for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) {
        const int co = i % Cout, t = i / Cout;          // t = c*9 + tap (OIHW inner order)
        sw[i] = __ldg(w + co * 27 + t) * s;
    }
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) sb[i] = bias ? __ldg(bias + i) : 0.f;
    __syncthreads();

    // block tile = quads_w pixel quads x rows image rows; its (rows+2) x (4*quads_w+2) x 3 input patch is staged in shared
    // memory by the whole block (coalesced, zero-filled outside the image): the per-thread global loads of the first
    // versions left every warp waiting on L2 latency once per tile (ncu: 35 % of samples on the first dependent FFMA).
    float* patch = sb + Cout;
    const int groups = Cout >> 3;
    const int qpb = blockDim.x / groups;                 // pixel quads per block tile
    const int g = threadIdx.x % groups;
    const int ql = threadIdx.x / groups;
    const int qstride = 4 * groups;                      // channels between this thread's two quads
    const int wq4 = W >> 2;
    const int quads_w = wq4 < 8 ? wq4 : 8;
    const int rows = qpb / quads_w;
    const int pw = 4 * quads_w + 2, ph = rows + 2;
    const int tiles_w = wq4 / quads_w;
    const int tiles_h = (H + rows - 1) / rows;
    const long total_tiles = static_cast<long>(N) * tiles_h * tiles_w;
    const int qx = ql % quads_w, qy = ql / quads_w;
    float ps[3], pb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        ps[c] = pre_scale ? __ldg(pre_scale + c) : 1.f;
        pb[c] = pre_shift ? __ldg(pre_shift + c) : 0.f;
    }
    for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int tw = static_cast<int>(tile % tiles_w);
        const int th = static_cast<int>((tile / tiles_w) % tiles_h);
        const long n = tile / (static_cast<long>(tiles_w) * tiles_h);
        const int wbase = tw * 4 * quads_w, hbase = th * rows;
        __syncthreads();                                  // previous tile's readers are done with the patch
        for (int i = threadIdx.x; i < 3 * ph * pw; i += blockDim.x) {
            const int col = i % pw, row = (i / pw) % ph, c = i / (pw * ph);
            const int hh = hbase + row - 1, ww = wbase + col - 1;
            float v = 0.f;
            if (hh >= 0 && hh < H && ww >= 0 && ww < W)
                v = __ldg(x + ((n * 3 + c) * static_cast<long>(H) + hh) * W + ww) * ps[c] + pb[c];
            patch[i] = v;
        }
        __syncthreads();
        const int hq = hbase + qy;
        if (hq >= H) continue;
        const int w0 = wbase + 4 * qx;
        float acc[4][8];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float4 b = *reinterpret_cast<const float4*>(sb + q * qstride + g * 4);
#pragma unroll
            for (int px = 0; px < 4; ++px) {
                acc[px][q * 4 + 0] = b.x; acc[px][q * 4 + 1] = b.y; acc[px][q * 4 + 2] = b.z; acc[px][q * 4 + 3] = b.w;
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const float* prow = patch + (c * ph + qy + kh) * pw + 4 * qx;
                float in[6];                              // columns w0-1 .. w0+4 of this input row
#pragma unroll
                for (int j = 0; j < 6; ++j) in[j] = prow[j];
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float* wr = sw + (c * 9 + kh * 3 + kw) * Cout + g * 4;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const float4 wv = *reinterpret_cast<const float4*>(wr + q * qstride);
#pragma unroll
                        for (int px = 0; px < 4; ++px) {
                            const float v = in[px + kw];
                            acc[px][q * 4 + 0] += v * wv.x; acc[px][q * 4 + 1] += v * wv.y;
                            acc[px][q * 4 + 2] += v * wv.z; acc[px][q * 4 + 3] += v * wv.w;
                        }
                    }
                }
            }
        }
        const long P = (n * H + hq) * static_cast<long>(W) + w0;
#pragma unroll
        for (int px = 0; px < 4; ++px) {
            float* yo = y + (P + px) * Cout + g * 4;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                float4 o = make_float4(acc[px][q * 4], acc[px][q * 4 + 1], acc[px][q * 4 + 2], acc[px][q * 4 + 3]);
                if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                if (round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
                *reinterpret_cast<float4*>(yo + q * qstride) = o;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ Cin = 3 dgrad
// dx[n,c,h,w] = pre_scale[c] * sum_{kh,kw,co} dy[n,h-kh+1,w-kw+1,co] * w[co][c][kh][kw]
__global__ void __launch_bounds__(256)
conv3x3_c3_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, const float* __restrict__ wscale,
                        const float* __restrict__ pre_scale, float* __restrict__ dx, int N, int H, int W, int Cout) {
    extern __shared__ float sm[];   // [9 taps][3][Cout]
    const float s = wscale ? __ldg(wscale) : 1.f;
    for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) {
        const int co = i % Cout;
        const int c = (i / Cout) % 3;
        const int tap = i / (3 * Cout);
        sm[i] = __ldg(w + (co * 3 + c) * 9 + tap) * s;
    }
    __syncthreads();
    const long P = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long total = static_cast<long>(N) * H * W;
    if (P >= total) return;
    const int wq = static_cast<int>(P % W);
    const int hq = static_cast<int>((P / W) % H);
    const long n = P / (static_cast<long>(W) * H);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int kh = 0; kh < 3; ++kh) {
        const int hh = hq - kh + 1;
        if (hh < 0 || hh >= H) continue;
        for (int kw = 0; kw < 3; ++kw) {
            const int ww = wq - kw + 1;
            if (ww < 0 || ww >= W) continue;
            const float* dp = dy + ((n * H + hh) * W + ww) * Cout;
            const float* w0 = sm + (kh * 3 + kw) * 3 * Cout;
            for (int co = 0; co < Cout; co += 4) {
                const float4 d = ldg4(dp + co);
                const float4 u0 = *reinterpret_cast<const float4*>(w0 + co);
                const float4 u1 = *reinterpret_cast<const float4*>(w0 + Cout + co);
                const float4 u2 = *reinterpret_cast<const float4*>(w0 + 2 * Cout + co);
                a0 += d.x * u0.x + d.y * u0.y + d.z * u0.z + d.w * u0.w;
                a1 += d.x * u1.x + d.y * u1.y + d.z * u1.z + d.w * u1.w;
                a2 += d.x * u2.x + d.y * u2.y + d.z * u2.z + d.w * u2.w;
            }
        }
    }
    const long HW = static_cast<long>(H) * W;
    const long o = n * 3 * HW + static_cast<long>(hq) * W + wq;
    dx[o] = a0 * (pre_scale ? __ldg(pre_scale + 0) : 1.f);
    dx[o + HW] = a1 * (pre_scale ? __ldg(pre_scale + 1) : 1.f);
    dx[o + 2 * HW] = a2 * (pre_scale ? __ldg(pre_scale + 2) : 1.f);
}

// ------------------------------------------------------------------------------------------------ Cin = 3 wgrad
// dw[co][c][tap] += scale * sum_p dy[p][co] * x[p+tap][c];  block = Cout channel lanes x (256/Cout) pixel lanes
__global__ void __launch_bounds__(256)
conv3x3_c3_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                        float scale, int N, int H, int W, int Cout, int pix_per_block) {
    extern __shared__ float sm[];   // [lanes][Cout][27]
    const int co = threadIdx.x % Cout;
    const int pl = threadIdx.x / Cout;
    const int lanes = blockDim.x / Cout;
    const long total = static_cast<long>(N) * H * W;
    const long p0 = static_cast<long>(blockIdx.x) * pix_per_block;
    long p1 = p0 + pix_per_block;
    if (p1 > total) p1 = total;
    float acc[27];
#pragma unroll
    for (int t = 0; t < 27; ++t) acc[t] = 0.f;
    const long HW = static_cast<long>(H) * W;
    for (long P = p0 + pl; P < p1; P += lanes) {
        const int wq = static_cast<int>(P % W);
        const int hq = static_cast<int>((P / W) % H);
        const long n = P / HW;
        const float d = __ldg(dy + P * Cout + co);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* xp = x + (n * 3 + c) * HW;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int hh = hq + kh - 1, ww = wq + kw - 1;
                    float v = 0.f;
                    if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = __ldg(xp + static_cast<long>(hh) * W + ww);
                    acc[c * 9 + kh * 3 + kw] += d * v;
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < 27; ++t) sm[(pl * Cout + co) * 27 + t] = acc[t];
    __syncthreads();
    for (int i = threadIdx.x; i < Cout * 27; i += blockDim.x) {
        float s = 0.f;
        for (int l = 0; l < lanes; ++l) s += sm[l * Cout * 27 + i];
        atomicAdd(dw + i, s * scale);   // i = co*27 + (c*9+tap) = OIHW order
    }
}

// ------------------------------------------------------------------------------------------------ generator tail
// forward: thread per pixel; weights [tap][ci] as float4 over the 4 output channels in shared memory
__global__ void __launch_bounds__(128)
gen_tail_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ wscale,
                    const float* __restrict__ bias, float* __restrict__ rgbs, float* __restrict__ segm,
                    float* __restrict__ t_out, int N, int H, int W, int Cin) {
    extern __shared__ float4 sw4[];   // [9][Cin]
    const float s = wscale ? __ldg(wscale) : 1.f;
    for (int i = threadIdx.x; i < 9 * Cin; i += blockDim.x) {
        const int ci = i % Cin, tap = i / Cin;
        float4 v;
        v.x = __ldg(w + (0 * Cin + ci) * 9 + tap) * s;
        v.y = __ldg(w + (1 * Cin + ci) * 9 + tap) * s;
        v.z = __ldg(w + (2 * Cin + ci) * 9 + tap) * s;
        v.w = __ldg(w + (3 * Cin + ci) * 9 + tap) * s;
        sw4[i] = v;
    }
    __syncthreads();
    const long P = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long HW = static_cast<long>(H) * W;
    if (P >= N * HW) return;
    const int wq = static_cast<int>(P % W);
    const int hq = static_cast<int>((P / W) % H);
    const long n = P / HW;
    float a0 = __ldg(bias + 0), a1 = __ldg(bias + 1), a2 = __ldg(bias + 2), a3 = __ldg(bias + 3);
    for (int kh = 0; kh < 3; ++kh) {
        const int hh = hq + kh - 1;
        if (hh < 0 || hh >= H) continue;
        for (int kw = 0; kw < 3; ++kw) {
            const int ww = wq + kw - 1;
            if (ww < 0 || ww >= W) continue;
            const float* xp = x + ((n * H + hh) * W + ww) * Cin;
            const float4* wt = sw4 + (kh * 3 + kw) * Cin;
#pragma unroll 4
            for (int ci = 0; ci < Cin; ci += 4) {
                const float4 v = ldg4(xp + ci);
                const float4 u0 = wt[ci], u1 = wt[ci + 1], u2 = wt[ci + 2], u3 = wt[ci + 3];
                a0 += v.x * u0.x + v.y * u1.x + v.z * u2.x + v.w * u3.x;
                a1 += v.x * u0.y + v.y * u1.y + v.z * u2.y + v.w * u3.y;
                a2 += v.x * u0.z + v.y * u1.z + v.z * u2.z + v.w * u3.z;
                a3 += v.x * u0.w + v.y * u1.w + v.z * u2.w + v.w * u3.w;
            }
        }
    }
    const float t0 = tanhf(a0), t1 = tanhf(a1), t2 = tanhf(a2), t3 = tanhf(a3);
    *reinterpret_cast<float4*>(t_out + P * 4) = make_float4(t0, t1, t2, t3);
    // generators/vector_pose_unsupervised_segmentation_noBottleneck.py:170-181
    const float sg = t3 * 0.5f + 0.5f;
    const long hw = static_cast<long>(hq) * W + wq;
    rgbs[(n * 3 + 0) * HW + hw] = (t0 * 0.75f + 0.5f) * sg;
    rgbs[(n * 3 + 1) * HW + hw] = (t1 * 0.75f + 0.5f) * sg;
    rgbs[(n * 3 + 2) * HW + hw] = (t2 * 0.75f + 0.5f) * sg;
    segm[n * HW + hw] = sg;
}

// Composition stage of the tensor-core tail: a32[n,h,w,0:4] (+ bias) is the 4-channel conv output computed by the
// tcgen05 bf16x3 kernel on a weight zero-padded to 32 output channels; t = tanh(a), then the same rgb*segm composition
// as gen_tail_fwd_kernel.  Thread per pixel.
__global__ void __launch_bounds__(256)
gen_tail_compose_kernel(const float* __restrict__ a32, const float* __restrict__ bias, float* __restrict__ rgbs,
                        float* __restrict__ segm, float* __restrict__ t_out, long NP, long HW, int stride) {
    const long P = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (P >= NP) return;
    const float4 a = ldg4(a32 + P * stride);
    const float t0 = tanhf(a.x + __ldg(bias + 0)), t1 = tanhf(a.y + __ldg(bias + 1)), t2 = tanhf(a.z + __ldg(bias + 2)),
                t3 = tanhf(a.w + __ldg(bias + 3));
    *reinterpret_cast<float4*>(t_out + P * 4) = make_float4(t0, t1, t2, t3);
    const float sg = t3 * 0.5f + 0.5f;
    const long n = P / HW, hw = P - n * HW;
    rgbs[(n * 3 + 0) * HW + hw] = (t0 * 0.75f + 0.5f) * sg;
    rgbs[(n * 3 + 1) * HW + hw] = (t1 * 0.75f + 0.5f) * sg;
    rgbs[(n * 3 + 2) * HW + hw] = (t2 * 0.75f + 0.5f) * sg;
    segm[n * HW + hw] = sg;
}

// da = d(loss)/d(pre-tanh) from d(fake_rgbs), d(fake_segm)
__global__ void gen_tail_bwd_act_kernel(const float* __restrict__ t, const float* __restrict__ d_rgbs,
                                        const float* __restrict__ d_segm, float* __restrict__ da, long NP, long HW,
                                        int da_stride) {
    const long P = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (P >= NP) return;
    const long n = P / HW, hw = P - n * HW;
    const float4 tv = ldg4(t + P * 4);
    const float sg = tv.w * 0.5f + 0.5f;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f, gs = 0.f;
    if (d_rgbs) {
        g0 = __ldg(d_rgbs + (n * 3 + 0) * HW + hw);
        g1 = __ldg(d_rgbs + (n * 3 + 1) * HW + hw);
        g2 = __ldg(d_rgbs + (n * 3 + 2) * HW + hw);
    }
    if (d_segm) gs = __ldg(d_segm + n * HW + hw);
    const float r0 = tv.x * 0.75f + 0.5f, r1 = tv.y * 0.75f + 0.5f, r2 = tv.z * 0.75f + 0.5f;
    const float dt0 = 0.75f * sg * g0, dt1 = 0.75f * sg * g1, dt2 = 0.75f * sg * g2;
    const float dt3 = 0.5f * (r0 * g0 + r1 * g1 + r2 * g2 + gs);
    float4 o;
    o.x = (1.f - tv.x * tv.x) * dt0;
    o.y = (1.f - tv.y * tv.y) * dt1;
    o.z = (1.f - tv.z * tv.z) * dt2;
    o.w = (1.f - tv.w * tv.w) * dt3;
    float* dst = da + P * da_stride;
    if (da_stride == 32) {   // the padded form feeds the tensor cores (weight / data gradient): round, do not truncate
        o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w);
    }
    *reinterpret_cast<float4*>(dst) = o;
    // da_stride 32: zero-padded to a 32-channel NHWC tensor so that the tensor-core weight-gradient kernel can take it
    for (int j = 4; j < da_stride; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(0.f, 0.f, 0.f, 0.f);
}

// dx[p][ci] = sum_tap sum_j da[p - off(tap)][j] * w[j][ci][tap];   thread = (pixel, 16-channel group)
__global__ void __launch_bounds__(256)
gen_tail_bwd_data_kernel(const float* __restrict__ da, const float* __restrict__ w, const float* __restrict__ wscale,
                         float* __restrict__ dx, int N, int H, int W, int Cin, int da_stride) {
    extern __shared__ float4 sw4[];   // [9][Cin] float4 over j
    const float s = wscale ? __ldg(wscale) : 1.f;
    for (int i = threadIdx.x; i < 9 * Cin; i += blockDim.x) {
        const int ci = i % Cin, tap = i / Cin;
        float4 v;
        v.x = __ldg(w + (0 * Cin + ci) * 9 + tap) * s;
        v.y = __ldg(w + (1 * Cin + ci) * 9 + tap) * s;
        v.z = __ldg(w + (2 * Cin + ci) * 9 + tap) * s;
        v.w = __ldg(w + (3 * Cin + ci) * 9 + tap) * s;
        sw4[i] = v;
    }
    __syncthreads();
    const int groups = Cin >> 4;
    const int ppb = blockDim.x / groups;
    const int g = threadIdx.x / ppb;
    const int pl = threadIdx.x - g * ppb;
    const long P = static_cast<long>(blockIdx.x) * ppb + pl;
    const long HW = static_cast<long>(H) * W;
    if (P >= N * HW) return;
    const int wq = static_cast<int>(P % W);
    const int hq = static_cast<int>((P / W) % H);
    const long n = P / HW;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    for (int kh = 0; kh < 3; ++kh) {
        const int hh = hq - kh + 1;
        if (hh < 0 || hh >= H) continue;
        for (int kw = 0; kw < 3; ++kw) {
            const int ww = wq - kw + 1;
            if (ww < 0 || ww >= W) continue;
            const float4 d = ldg4(da + ((n * H + hh) * W + ww) * da_stride);
            const float4* wt = sw4 + (kh * 3 + kw) * Cin + g * 16;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float4 u = wt[j];
                acc[j] += d.x * u.x + d.y * u.y + d.z * u.z + d.w * u.w;
            }
        }
    }
    float* o = dx + P * Cin + g * 16;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(o + q * 4) = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
}

// dw[j][ci][tap] += sum_p da[p][j] * x[p+off(tap)][ci];  db[j] += sum_p da[p][j]
// block = Cin channel lanes x (256/Cin) pixel lanes
__global__ void __launch_bounds__(256)
gen_tail_bwd_weight_kernel(const float* __restrict__ x, const float* __restrict__ da, float* __restrict__ dw,
                           float* __restrict__ db, int N, int H, int W, int Cin, int pix_per_block) {
    extern __shared__ float sm[];   // [lanes][Cin][36] + [lanes][4]
    const int ci = threadIdx.x % Cin;
    const int pl = threadIdx.x / Cin;
    const int lanes = blockDim.x / Cin;
    const long HW = static_cast<long>(H) * W;
    const long total = N * HW;
    const long p0 = static_cast<long>(blockIdx.x) * pix_per_block;
    long p1 = p0 + pix_per_block;
    if (p1 > total) p1 = total;
    float acc[36];
#pragma unroll
    for (int t = 0; t < 36; ++t) acc[t] = 0.f;
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
    for (long P = p0 + pl; P < p1; P += lanes) {
        const int wq = static_cast<int>(P % W);
        const int hq = static_cast<int>((P / W) % H);
        const long n = P / HW;
        const float4 d = ldg4(da + P * 4);
        bsum[0] += d.x; bsum[1] += d.y; bsum[2] += d.z; bsum[3] += d.w;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int hh = hq + kh - 1, ww = wq + kw - 1;
                float v = 0.f;
                if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = __ldg(x + ((n * H + hh) * W + ww) * Cin + ci);
                const int tap = kh * 3 + kw;
                acc[0 * 9 + tap] += d.x * v;
                acc[1 * 9 + tap] += d.y * v;
                acc[2 * 9 + tap] += d.z * v;
                acc[3 * 9 + tap] += d.w * v;
            }
        }
    }
    float* sacc = sm;
    float* sbs = sm + static_cast<size_t>(lanes) * Cin * 36;
#pragma unroll
    for (int t = 0; t < 36; ++t) sacc[(static_cast<size_t>(pl) * Cin + ci) * 36 + t] = acc[t];
    if (ci == 0) { sbs[pl * 4 + 0] = bsum[0]; sbs[pl * 4 + 1] = bsum[1]; sbs[pl * 4 + 2] = bsum[2]; sbs[pl * 4 + 3] = bsum[3]; }
    __syncthreads();
    for (int i = threadIdx.x; i < Cin * 36; i += blockDim.x) {
        float s = 0.f;
        for (int l = 0; l < lanes; ++l) s += sacc[static_cast<size_t>(l) * Cin * 36 + i];
        const int c = i / 36, r = i - c * 36;       // r = j*9 + tap
        const int j = r / 9, tap = r - j * 9;
        atomicAdd(dw + (static_cast<size_t>(j) * Cin + c) * 9 + tap, s);
    }
    if (threadIdx.x < 4) {
        float s = 0.f;
        for (int l = 0; l < lanes; ++l) s += sbs[l * 4 + threadIdx.x];
        atomicAdd(db + threadIdx.x, s);
    }
}

// ------------------------------------------------------------------------------------------------ Cin = 3 via GEMM
// The weight- and data-gradient of the 3-channel stem convs are tiny-K GEMMs over 524288 pixels; as CUDA-core loops
// they were the slowest kernels of the step (profiles/r01_bench_graph_first.json).  They run on the tensor cores
// instead, through an explicit 27(+5 zero)-column patch matrix:
//   col[p][c*9 + kh*3 + kw] = x[n, c, h+kh-1, w+kw-1]         (im2col, NHWC with 32 "channels", tf32-rounded)
//   dW = conv_wgrad(col, dy, 1x1)[:, :27]                     dcol = conv1x1(dy, W^T) ;  dx = col2im(dcol)
__global__ void __launch_bounds__(256)
im2col3x3_c3_kernel(const float* __restrict__ x, float* __restrict__ col, int N, int H, int W) {
    const long P = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long HW = static_cast<long>(H) * W;
    if (P >= N * HW) return;
    const int wq = static_cast<int>(P % W);
    const int hq = static_cast<int>((P / W) % H);
    const long n = P / HW;
    float v[32];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float* xp = x + (n * 3 + c) * HW;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int hh = hq + kh - 1, ww = wq + kw - 1;
                float t = 0.f;
                if (hh >= 0 && hh < H && ww >= 0 && ww < W) t = __ldg(xp + static_cast<long>(hh) * W + ww);
                v[c * 9 + kh * 3 + kw] = round_tf32(t);
            }
        }
    }
#pragma unroll
    for (int j = 27; j < 32; ++j) v[j] = 0.f;
    float4* dst = reinterpret_cast<float4*>(col + P * 32);
#pragma unroll
    for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// dx[n,c,h,w] = pre_scale[c] * sum_{kh,kw} dcol[n, h-kh+1, w-kw+1][c*9 + kh*3 + kw]
__global__ void __launch_bounds__(256)
col2im3x3_c3_kernel(const float* __restrict__ dcol, const float* __restrict__ pre_scale, float* __restrict__ dx, int N,
                    int H, int W) {
    const long P = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long HW = static_cast<long>(H) * W;
    if (P >= N * HW) return;
    const int wq = static_cast<int>(P % W);
    const int hq = static_cast<int>((P / W) % H);
    const long n = P / HW;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        const int hh = hq - kh + 1;
        if (hh < 0 || hh >= H) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const int ww = wq - kw + 1;
            if (ww < 0 || ww >= W) continue;
            const float* r = dcol + ((n * H + hh) * W + ww) * 32 + kh * 3 + kw;
            a0 += __ldg(r);
            a1 += __ldg(r + 9);
            a2 += __ldg(r + 18);
        }
    }
    const long o = n * 3 * HW + static_cast<long>(hq) * W + wq;
    dx[o] = a0 * (pre_scale ? __ldg(pre_scale + 0) : 1.f);
    dx[o + HW] = a1 * (pre_scale ? __ldg(pre_scale + 1) : 1.f);
    dx[o + 2 * HW] = a2 * (pre_scale ? __ldg(pre_scale + 2) : 1.f);
}

}  // namespace b200lp

using namespace b200lp;

extern "C" int32_t b200lp_im2col3x3_c3(const float* x_nchw, float* col_nhwc32, int32_t N, int32_t H, int32_t W,
                                       void* stream) {
    B200LP_REQUIRE(x_nchw && col_nhwc32 && N > 0 && H > 0 && W > 0, "im2col3x3_c3: bad args");
    const long total = static_cast<long>(N) * H * W;
    im2col3x3_c3_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, as_stream(stream)>>>(x_nchw, col_nhwc32, N, H, W);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_col2im3x3_c3(const float* dcol_nhwc32, const float* pre_scale, float* dx_nchw, int32_t N,
                                       int32_t H, int32_t W, void* stream) {
    B200LP_REQUIRE(dcol_nhwc32 && dx_nchw && N > 0 && H > 0 && W > 0, "col2im3x3_c3: bad args");
    const long total = static_cast<long>(N) * H * W;
    col2im3x3_c3_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, as_stream(stream)>>>(dcol_nhwc32, pre_scale,
                                                                                              dx_nchw, N, H, W);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

namespace b200lp {
// ------------------------------------------------------------------------------------------------ Cin = 3 forward, tcgen05
// The same layer as an implicit GEMM on the tensor cores: M = 128 consecutive pixels, N = 64 output channels, K = 27 taps
// padded to 32.  There is no TMA-able layout for a 3-channel NCHW patch, so the 128 pixel threads BUILD the A tile:
// each gathers its 27 inputs (input normalisation applied, zero padding, tf32 rounding) and writes its 128-byte row in
// the canonical K-major SWIZZLE_128B order (16-byte chunk j of row r at ((j ^ (r & 7)) << 4)), a proxy fence hands the
// tile to the async proxy, one thread issues four K = 8 MMAs into a 64-column TMEM accumulator, and the same 128 threads
// run the epilogue (1/sigma, bias, ReLU, tf32 rounding, 256 contiguous bytes per pixel).  The CUDA-core kernel above
// needs 1728 FMAs per pixel from shared-memory weights (72 us for 8 x 256^2); here the layer is bound by the 134 MB it writes.
// Several CTAs per SM overlap each other's gather / MMA / store phases (24 KB of shared memory, 64 TMEM columns each).
constexpr int kC3Cout = 64;

__global__ void __launch_bounds__(128, 6)
conv3x3_c3_tc_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ wscale,
                     const float* __restrict__ bias, const float* __restrict__ pre_scale,
                     const float* __restrict__ pre_shift, float* __restrict__ y, int N, int H, int W, int relu,
                     int round_out, int total_tiles) {
    __shared__ __align__(1024) uint8_t sA[128 * 128];
    __shared__ __align__(1024) uint8_t sB[kC3Cout * 128];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int r = threadIdx.x;                    // pixel row of the tile == TMEM lane
    const int warp = r >> 5;
    if (r == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<64>(&tmem_slot);
    // B tile: row co = 27 (+5 zero) weights, k = c*9 + kh*3 + kw (the OIHW inner order), tf32-rounded, swizzled like A
    if (r < kC3Cout) {
        float wk[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) wk[k] = k < 27 ? round_tf32(__ldg(w + r * 27 + k)) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(sB + r * 128 + ((j ^ (r & 7)) << 4)) =
                make_float4(wk[4 * j], wk[4 * j + 1], wk[4 * j + 2], wk[4 * j + 3]);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const float s = wscale ? __ldg(wscale) : 1.f;
    float ps[3], pb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        ps[c] = pre_scale ? __ldg(pre_scale + c) : 1.f;
        pb[c] = pre_shift ? __ldg(pre_shift + c) : 0.f;
    }
    constexpr uint32_t idesc = make_idesc_tf32(128, kC3Cout, 0, 0);
    const uint64_t desc_hi = make_smem_desc(0, 16, 1024, 2);
    const uint64_t da0 = desc_hi | ((smem_u32(sA) >> 4) & 0x3FFFu);
    const uint64_t db0 = desc_hi | ((smem_u32(sB) >> 4) & 0x3FFFu);
    const long P = static_cast<long>(N) * H * W;
    const long HW = static_cast<long>(H) * W;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const long p = static_cast<long>(tile) * 128 + r;
        const bool valid = p < P;
        const long pp = valid ? p : 0;
        const int n = static_cast<int>(pp / HW);
        const int hw = static_cast<int>(pp - n * HW);
        const int h = hw / W, wq = hw - h * W;
        // ---- gather: 27 inputs of this pixel (consecutive threads = consecutive w: coalesced, overlaps served by L1)
        float v[32];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* xc = x + (static_cast<long>(n) * 3 + c) * HW;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int hh = h + kh - 1;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int ww = wq + kw - 1;
                    const bool in = valid && hh >= 0 && hh < H && ww >= 0 && ww < W;
                    const float t = in ? __ldg(xc + static_cast<long>(hh) * W + ww) : 0.f;
                    v[c * 9 + kh * 3 + kw] = in ? round_tf32(t * ps[c] + pb[c]) : 0.f;
                }
            }
        }
#pragma unroll
        for (int k = 27; k < 32; ++k) v[k] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(sA + r * 128 + ((j ^ (r & 7)) << 4)) =
                make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        fence_proxy_async_smem();                 // generic-proxy writes -> visible to the tensor core's async proxy
        __syncthreads();
        if (r == 0) {
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_tf32_ss(tmem, da0 + 2 * k, db0 + 2 * k, idesc, k > 0 ? 1u : 0u);
            umma_commit(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1u;
        tc_fence_after();
        // ---- epilogue: 32 channels at a time through the (now free) A tile so that the global stores are coalesced: each
        // thread parks its pixel's 128 bytes in shared memory (same XOR swizzle: conflict-free), then the block writes the
        // 128 x 128-byte half-rows with consecutive threads on consecutive 16-byte chunks (a warp = 4 whole lines)
        const long p0 = static_cast<long>(tile) * 128;
#pragma unroll 1
        for (int c0 = 0; c0 < kC3Cout; c0 += 32) {
            uint32_t a[32];
            tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, a);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 o = make_float4(__uint_as_float(a[j]) * s, __uint_as_float(a[j + 1]) * s,
                                       __uint_as_float(a[j + 2]) * s, __uint_as_float(a[j + 3]) * s);
                if (bias) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c0 + j));
                    o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                }
                if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                if (round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
                *reinterpret_cast<float4*>(sA + r * 128 + (((j >> 2) ^ (r & 7)) << 4)) = o;
            }
            __syncthreads();
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int q = (r >> 3) + 16 * it, j = r & 7;            // pixel q of the tile, 16-byte chunk j
                if (p0 + q < P)
                    *reinterpret_cast<float4*>(y + (p0 + q) * kC3Cout + c0 + 4 * j) =
                        *reinterpret_cast<const float4*>(sA + q * 128 + ((j ^ (q & 7)) << 4));
            }
            __syncthreads();
        }
        tc_fence_before();
        __syncthreads();                          // every lane has drained the accumulator and the A tile is free again
    }
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<64>(tmem);
    }
}
}  // namespace b200lp

extern "C" int32_t b200lp_conv3x3_c3_fwd_tc(const float* x_nchw, const float* w_oihw, const float* wscale,
                                            const float* bias, const float* pre_scale, const float* pre_shift,
                                            float* y_nhwc, int32_t N, int32_t H, int32_t W, int32_t Cout, int32_t relu,
                                            int32_t round_tf32, void* stream) {
    B200LP_REQUIRE(x_nchw && w_oihw && y_nhwc && N > 0 && H > 0 && W > 0, "conv3x3_c3_fwd_tc: bad args");
    B200LP_REQUIRE(Cout == kC3Cout, "conv3x3_c3_fwd_tc: Cout=%d (the tensor-core stem takes 64 output channels)", Cout);
    const long P = static_cast<long>(N) * H * W;
    B200LP_REQUIRE(P < (1L << 31) * 64, "conv3x3_c3_fwd_tc: too many pixels");
    const int total_tiles = static_cast<int>((P + 127) / 128);
    const int blocks = total_tiles < 148 * 6 ? total_tiles : 148 * 6;
    conv3x3_c3_tc_kernel<<<blocks, 128, 0, as_stream(stream)>>>(x_nchw, w_oihw, wscale, bias, pre_scale, pre_shift, y_nhwc,
                                                               N, H, W, relu, round_tf32, total_tiles);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_conv3x3_c3_fwd(const float* x_nchw, const float* w_oihw, const float* wscale,
                                         const float* bias, const float* pre_scale, const float* pre_shift,
                                         float* y_nhwc, int32_t N, int32_t H, int32_t W, int32_t Cout, int32_t relu,
                                         int32_t round_tf32, void* stream) {
    B200LP_REQUIRE(x_nchw && w_oihw && y_nhwc, "conv3x3_c3_fwd: null pointer");
    B200LP_REQUIRE(N > 0 && H > 0 && W > 0 && W % 4 == 0 && Cout % 16 == 0 && Cout >= 16 && Cout <= 128 &&
                       (256 % (Cout / 16)) == 0,
                   "conv3x3_c3_fwd: bad shape (Cout=%d, W=%d: Cout in {16,32,64,128}, W %% 4 == 0)", Cout, W);
    const int groups = Cout / 8;
    const int qpb = 256 / groups;                                 // pixel quads per block tile
    const int wq4 = W / 4;
    const int quads_w = wq4 < 8 ? wq4 : 8;
    B200LP_REQUIRE(qpb % quads_w == 0 && wq4 % quads_w == 0, "conv3x3_c3_fwd: W=%d not tileable (Cout=%d)", W, Cout);
    const int rows = qpb / quads_w;
    const long total_tiles = static_cast<long>(N) * ((H + rows - 1) / rows) * (wq4 / quads_w);
    long blocks = total_tiles < 148 * 6 ? total_tiles : 148 * 6;   // persistent: weights staged once per block
    const size_t smem = (27 * Cout + Cout + 3 * (rows + 2) * (4 * quads_w + 2)) * sizeof(float);
    conv3x3_c3_fwd_kernel<<<static_cast<int>(blocks), 256, smem, as_stream(stream)>>>(
        x_nchw, w_oihw, wscale, bias, pre_scale, pre_shift, y_nhwc, N, H, W, Cout, relu, round_tf32);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_conv3x3_c3_dgrad(const float* dy_nhwc, const float* w_oihw, const float* wscale,
                                           const float* pre_scale, float* dx_nchw, int32_t N, int32_t H, int32_t W,
                                           int32_t Cout, void* stream) {
    B200LP_REQUIRE(dy_nhwc && w_oihw && dx_nchw && N > 0 && H > 0 && W > 0 && Cout % 4 == 0 && Cout <= 256,
                   "conv3x3_c3_dgrad: bad args");
    const long total = static_cast<long>(N) * H * W;
    const int blocks = static_cast<int>((total + 255) / 256);
    conv3x3_c3_dgrad_kernel<<<blocks, 256, 27 * Cout * 4, as_stream(stream)>>>(dy_nhwc, w_oihw, wscale, pre_scale,
                                                                               dx_nchw, N, H, W, Cout);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_conv3x3_c3_wgrad(const float* x_nchw, const float* dy_nhwc, float* dw_oihw,
                                           float wscale_host, int32_t N, int32_t H, int32_t W, int32_t Cout,
                                           void* stream) {
    B200LP_REQUIRE(x_nchw && dy_nhwc && dw_oihw && N > 0 && H > 0 && W > 0 && Cout > 0 && 256 % Cout == 0,
                   "conv3x3_c3_wgrad: bad args (Cout=%d must divide 256)", Cout);
    cudaStream_t s = as_stream(stream);
    B200LP_CHECK_CUDA(cudaMemsetAsync(dw_oihw, 0, static_cast<size_t>(Cout) * 27 * 4, s));
    const long total = static_cast<long>(N) * H * W;
    const int lanes = 256 / Cout;
    int ppb = 1024;
    if (ppb < lanes) ppb = lanes;
    const int blocks = static_cast<int>((total + ppb - 1) / ppb);
    conv3x3_c3_wgrad_kernel<<<blocks, 256, static_cast<size_t>(lanes) * Cout * 27 * 4, s>>>(
        x_nchw, dy_nhwc, dw_oihw, wscale_host, N, H, W, Cout, ppb);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_gen_tail_fwd(const float* x_nhwc, const float* w_oihw, const float* wscale,
                                       const float* bias, float* fake_rgbs_nchw, float* fake_segm_nchw, float* t_out,
                                       int32_t N, int32_t H, int32_t W, int32_t Cin, void* stream) {
    B200LP_REQUIRE(x_nhwc && w_oihw && bias && fake_rgbs_nchw && fake_segm_nchw && t_out, "gen_tail_fwd: null pointer");
    B200LP_REQUIRE(N > 0 && H > 0 && W > 0 && Cin % 4 == 0 && Cin <= 512, "gen_tail_fwd: bad shape");
    const long total = static_cast<long>(N) * H * W;
    const int blocks = static_cast<int>((total + 127) / 128);
    gen_tail_fwd_kernel<<<blocks, 128, 9 * Cin * 16, as_stream(stream)>>>(x_nhwc, w_oihw, wscale, bias, fake_rgbs_nchw,
                                                                          fake_segm_nchw, t_out, N, H, W, Cin);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_gen_tail_compose(const float* a_nhwc, const float* bias, float* fake_rgbs_nchw,
                                           float* fake_segm_nchw, float* t_out, int32_t N, int32_t H, int32_t W,
                                           int32_t a_stride, void* stream) {
    B200LP_REQUIRE(a_nhwc && bias && fake_rgbs_nchw && fake_segm_nchw && t_out && N > 0 && H > 0 && W > 0 &&
                       a_stride >= 4 && a_stride % 4 == 0,
                   "gen_tail_compose: bad args");
    const long HW = static_cast<long>(H) * W;
    const long NP = N * HW;
    gen_tail_compose_kernel<<<static_cast<int>((NP + 255) / 256), 256, 0, as_stream(stream)>>>(
        a_nhwc, bias, fake_rgbs_nchw, fake_segm_nchw, t_out, NP, HW, a_stride);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_gen_tail_bwd_act(const float* t, const float* d_rgbs, const float* d_segm, float* da,
                                           int32_t N, int32_t H, int32_t W, int32_t da_stride, void* stream) {
    B200LP_REQUIRE(t && da && N > 0 && H > 0 && W > 0 && (da_stride == 4 || da_stride == 32), "gen_tail_bwd_act: bad args");
    const long HW = static_cast<long>(H) * W;
    const long NP = N * HW;
    gen_tail_bwd_act_kernel<<<static_cast<int>((NP + 255) / 256), 256, 0, as_stream(stream)>>>(t, d_rgbs, d_segm, da,
                                                                                               NP, HW, da_stride);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_gen_tail_bwd_data(const float* da, const float* w_oihw, const float* wscale, float* dx_nhwc,
                                            int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t da_stride,
                                            void* stream) {
    B200LP_REQUIRE(da && w_oihw && dx_nhwc && N > 0 && H > 0 && W > 0 && Cin % 16 == 0 && Cin <= 128 &&
                       256 % (Cin / 16) == 0 && (da_stride == 4 || da_stride == 32),
                   "gen_tail_bwd_data: bad args");
    const int groups = Cin / 16;
    const int ppb = 256 / groups;
    const long total = static_cast<long>(N) * H * W;
    const int blocks = static_cast<int>((total + ppb - 1) / ppb);
    gen_tail_bwd_data_kernel<<<blocks, 256, 9 * Cin * 16, as_stream(stream)>>>(da, w_oihw, wscale, dx_nhwc, N, H, W,
                                                                               Cin, da_stride);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_gen_tail_bwd_weight(const float* x_nhwc, const float* da, float* dw_oihw, float* dbias,
                                              int32_t N, int32_t H, int32_t W, int32_t Cin, void* stream) {
    B200LP_REQUIRE(x_nhwc && da && dw_oihw && dbias && N > 0 && H > 0 && W > 0 && Cin > 0 && 256 % Cin == 0,
                   "gen_tail_bwd_weight: bad args (Cin=%d must divide 256)", Cin);
    cudaStream_t s = as_stream(stream);
    B200LP_CHECK_CUDA(cudaMemsetAsync(dw_oihw, 0, static_cast<size_t>(4) * Cin * 9 * 4, s));
    B200LP_CHECK_CUDA(cudaMemsetAsync(dbias, 0, 16, s));
    const long total = static_cast<long>(N) * H * W;
    const int lanes = 256 / Cin;
    int ppb = 1024;
    const int blocks = static_cast<int>((total + ppb - 1) / ppb);
    const size_t smem = (static_cast<size_t>(lanes) * Cin * 36 + lanes * 4) * 4;
    gen_tail_bwd_weight_kernel<<<blocks, 256, smem, s>>>(x_nhwc, da, dw_oihw, dbias, N, H, W, Cin, ppb);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}
