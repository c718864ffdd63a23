// Pose-encoder (torchvision MobileNetV2) BACKWARD kernels — FP32 on the CUDA cores, like the forward (csrc/mobilenet.cu).
//
// Replaces autograd's backward of `Embedder.get_pose_embedding` (embedders/unsupervised_pose_separate_embResNeXt_
// segmentation.py:56-58) in meta-training: ~350 cuDNN / ATen launches (depthwise grad kernels, 1x1 GEMMs, BatchNorm
// backward, hardtanh backward, adds).  BatchNorm backward itself is b200lp_bn_bwd (csrc/encoder.cu, ReLU6 mask mode).
//
//   pw_wgrad    dW[co][ci] (+)= sum_m dy[m][co] * f(x[m][ci])        1x1 conv / linear weight gradient, split over the rows,
//                                                                   f = producer BatchNorm (+ReLU6) applied on load
//   dw_dgrad    dx[n,hi,wi,c] = sum_taps dy[n,ho,wo,c] * w[c][tap]   depthwise 3x3 data gradient (stride 1 / 2, gather)
//   dw_wgrad    dW[c][tap] (+)= sum_p dy[p][c] * f(x[p + tap][c])     depthwise weight gradient
//   stem_wgrad  dW[32][27] (+)= sum_p dy[p][co] * patch(x)[p][27]    3x3 stride-2 stem on the NCHW image
//   transpose2d dst[c][r] = src[r][c]                                the 1x1 data gradient runs as pw_conv on W^T
// Every reduction is two-stage with a fixed order (no atomics).
#include "common.cuh"

namespace b200lp {

namespace {
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float relu6f(float v) { return fminf(fmaxf(v, 0.f), 6.f); }
}  // namespace

__global__ void __launch_bounds__(256)
transpose2d_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int C) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < C) ? __ldg(src + static_cast<size_t>(r) * C + c) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < C && r < R) dst[static_cast<size_t>(c) * R + r] = tile[threadIdx.x][i];
    }
}

// ------------------------------------------------------------------------------------------------ 1x1 weight gradient
// part[z][co][ci] = sum over the z-th row chunk of dy[m][co] * f(x[m][ci]);  64 x 64 output tile, 16 rows per step,
// 256 threads x (4 x 4) outputs.  grid = (ci tiles, co tiles, row chunks)
__global__ void __launch_bounds__(256)
pw_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ in_scale,
                const float* __restrict__ in_shift, int in_relu6, float* __restrict__ part, int M, int Cout, int Cin,
                int rows_per_chunk) {
    __shared__ float As[16][68], Bs[16][68];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int ci0 = blockIdx.x * 64, co0 = blockIdx.y * 64;
    const int m0 = blockIdx.z * rows_per_chunk;
    const int m1 = min(m0 + rows_per_chunk, M);
    // loader: 16 rows x 16 float4 per operand tile
    const int lr = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 4;
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool bin = ci0 + lc < Cin;          // Cin % 4 == 0: a float4 is entirely inside or outside
    const bool ain = co0 + lc < Cout;
    if (in_scale && bin) { sc = ldg4(in_scale + ci0 + lc); sh = ldg4(in_shift + ci0 + lc); }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int mk = m0; mk < m1; mk += 16) {
        const int m = mk + lr;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < m1) {
            if (ain) a = ldg4(dy + static_cast<size_t>(m) * Cout + co0 + lc);
            if (bin) {
                b = ldg4(x + static_cast<size_t>(m) * Cin + ci0 + lc);
                if (in_scale) {
                    b.x = b.x * sc.x + sh.x; b.y = b.y * sc.y + sh.y; b.z = b.z * sc.z + sh.z; b.w = b.w * sc.w + sh.w;
                    if (in_relu6) { b.x = relu6f(b.x); b.y = relu6f(b.y); b.z = relu6f(b.z); b.w = relu6f(b.w); }
                }
            }
        }
        __syncthreads();
        *reinterpret_cast<float4*>(&As[lr][lc]) = a;
        *reinterpret_cast<float4*>(&Bs[lr][lc]) = b;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
    }
    float* base = part + static_cast<size_t>(blockIdx.z) * Cout * Cin;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + ty * 4 + i;
        if (co >= Cout) continue;
        const int ci = ci0 + tx * 4;
        if (ci < Cin) *reinterpret_cast<float4*>(base + static_cast<size_t>(co) * Cin + ci) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
}

// dst[i] (+)= sum_z part[z][i]   (fixed order)
__global__ void __launch_bounds__(256)
sum_parts_kernel(const float* __restrict__ part, float* __restrict__ dst, int nparts, long total, int accumulate) {
    for (long i = blockIdx.x * 256L + threadIdx.x; i < total; i += gridDim.x * 256L) {
        float s = 0.f;
#pragma unroll 8
        for (int z = 0; z < nparts; ++z) s += __ldg(part + static_cast<size_t>(z) * total + i);   // 8 loads in flight, fixed order
        dst[i] = (accumulate ? dst[i] : 0.f) + s;
    }
}

// ------------------------------------------------------------------------------------------------ depthwise 3x3 backward
// dx[n,hi,wi,c] = sum over (kh,kw) with ho*stride + kh - 1 == hi (and the same for w) of dy[n,ho,wo,c] * w[c][kh*3+kw]
__global__ void __launch_bounds__(256)
dw_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int N, int H, int W, int C,
                int stride, int Ho, int Wo) {
    const unsigned C4 = C >> 2;
    const unsigned total4 = static_cast<unsigned>(N) * H * W * C4;
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < total4; i += gridDim.x * 256u) {
        const unsigned c = (i % C4) * 4;
        const unsigned p = i / C4;
        const int wi = static_cast<int>(p % static_cast<unsigned>(W));
        const unsigned t2 = p / static_cast<unsigned>(W);
        const int hi = static_cast<int>(t2 % static_cast<unsigned>(H));
        const unsigned n = t2 / static_cast<unsigned>(H);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int h2 = hi + 1 - kh;
            if (h2 < 0 || h2 % stride) continue;
            const int ho = h2 / stride;
            if (ho >= Ho) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int w2 = wi + 1 - kw;
                if (w2 < 0 || w2 % stride) continue;
                const int wo = w2 / stride;
                if (wo >= Wo) continue;
                const float4 g = ldg4(dy + (static_cast<size_t>(n * Ho + ho) * Wo + wo) * C + c);
                const int t = kh * 3 + kw;
                o.x = fmaf(g.x, __ldg(w + (c + 0) * 9 + t), o.x); o.y = fmaf(g.y, __ldg(w + (c + 1) * 9 + t), o.y);
                o.z = fmaf(g.z, __ldg(w + (c + 2) * 9 + t), o.z); o.w = fmaf(g.w, __ldg(w + (c + 3) * 9 + t), o.w);
            }
        }
        *reinterpret_cast<float4*>(dx + static_cast<size_t>(i) * 4) = o;
    }
}

// part[chunk][tap][c] = sum over the chunk's output pixels of dy[p][c] * relu6(x[p (+) tap][c]*scale+shift)
// block = Qb channel quads x (256 / Qb) pixel lanes (C / 4 <= 256);  grid = pixel chunks
__global__ void __launch_bounds__(256)
dw_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                const float* __restrict__ dy, float* __restrict__ part, int N, int H, int W, int C, int stride, int Ho,
                int Wo, int pix_per_chunk) {
    extern __shared__ float4 red[];                  // [lanes][Qb][9]
    const int Qb = C >> 2;
    const int lanes = blockDim.x / Qb;
    const int q = threadIdx.x % Qb, ln = threadIdx.x / Qb;
    const int c = q * 4;
    float4 acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ln < lanes) {
        const float4 sc = ldg4(in_scale + c), sh = ldg4(in_shift + c);
        const unsigned P = static_cast<unsigned>(N) * Ho * Wo;
        const unsigned p0 = blockIdx.x * static_cast<unsigned>(pix_per_chunk);
        const unsigned p1 = min(p0 + static_cast<unsigned>(pix_per_chunk), P);
        for (unsigned p = p0 + ln; p < p1; p += lanes) {
            const int wo = static_cast<int>(p % static_cast<unsigned>(Wo));
            const unsigned t2 = p / static_cast<unsigned>(Wo);
            const int ho = static_cast<int>(t2 % static_cast<unsigned>(Ho));
            const unsigned n = t2 / static_cast<unsigned>(Ho);
            const float4 g = ldg4(dy + static_cast<size_t>(p) * C + c);
            float4 v[9];
            bool ok[9];
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int hi = ho * stride + kh - 1;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int wi = wo * stride + kw - 1;
                    const bool in = hi >= 0 && hi < H && wi >= 0 && wi < W;
                    ok[kh * 3 + kw] = in;
                    v[kh * 3 + kw] = ldg4(x + (static_cast<size_t>(n * H + (in ? hi : 0)) * W + (in ? wi : 0)) * C + c);
                }
            }
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                if (!ok[t]) continue;
                const float4 a = v[t];
                acc[t].x = fmaf(g.x, relu6f(a.x * sc.x + sh.x), acc[t].x);
                acc[t].y = fmaf(g.y, relu6f(a.y * sc.y + sh.y), acc[t].y);
                acc[t].z = fmaf(g.z, relu6f(a.z * sc.z + sh.z), acc[t].z);
                acc[t].w = fmaf(g.w, relu6f(a.w * sc.w + sh.w), acc[t].w);
            }
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) red[(static_cast<size_t>(ln) * Qb + q) * 9 + t] = acc[t];
    }
    __syncthreads();
    if (ln == 0) {
        for (int l = 1; l < lanes; ++l)
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float4 o = red[(static_cast<size_t>(l) * Qb + q) * 9 + t];
                acc[t].x += o.x; acc[t].y += o.y; acc[t].z += o.z; acc[t].w += o.w;
            }
        float* base = part + static_cast<size_t>(blockIdx.x) * 9 * C;
#pragma unroll
        for (int t = 0; t < 9; ++t) *reinterpret_cast<float4*>(base + static_cast<size_t>(t) * C + c) = acc[t];
    }
}

// dW[c][tap] (+)= sum_chunks part[chunk][tap][c]
__global__ void __launch_bounds__(256)
dw_wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, int chunks, int C, int accumulate) {
    // threads along c (the partials' contiguous axis), 8 chunk loads in flight; one launch summed 592 chunks with a
    // dependent, 36-byte-strided load per step (52..73 us for <= 8640 outputs)
    const int total = C * 9;
    for (int j = blockIdx.x * 256 + threadIdx.x; j < total; j += gridDim.x * 256) {
        const int t = j / C, c = j - t * C;
        float s = 0.f;
#pragma unroll 8
        for (int k = 0; k < chunks; ++k) s += __ldg(part + static_cast<size_t>(k) * total + j);
        const int i = c * 9 + t;
        dw[i] = (accumulate ? dw[i] : 0.f) + s;
    }
}

// ------------------------------------------------------------------------------------------------ stem weight gradient
// part[chunk][co][27] = sum over the chunk's output pixels of dy[p][co] * x_nchw[n][c][2ho+kh-1][2wo+kw-1]
// block 224 = 27 patch positions x 8 output-channel quads (+8 idle)
__global__ void __launch_bounds__(224)
mbv2_stem_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ part, int N, int H,
                       int W, int Ho, int Wo, int pix_per_chunk) {
    const int k = threadIdx.x >> 3, q = threadIdx.x & 7;      // k = c*9 + kh*3 + kw
    if (k >= 27) return;
    const int c = k / 9, kh = (k % 9) / 3, kw = k % 3;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const unsigned P = static_cast<unsigned>(N) * Ho * Wo;
    const unsigned p0 = blockIdx.x * static_cast<unsigned>(pix_per_chunk);
    const unsigned p1 = min(p0 + static_cast<unsigned>(pix_per_chunk), P);
    for (unsigned p = p0; p < p1; ++p) {
        const int wo = static_cast<int>(p % static_cast<unsigned>(Wo));
        const unsigned t2 = p / static_cast<unsigned>(Wo);
        const int ho = static_cast<int>(t2 % static_cast<unsigned>(Ho));
        const unsigned n = t2 / static_cast<unsigned>(Ho);
        const int hi = ho * 2 + kh - 1, wi = wo * 2 + kw - 1;
        if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
        const float v = __ldg(x + ((static_cast<size_t>(n) * 3 + c) * H + hi) * W + wi);
        const float4 g = ldg4(dy + static_cast<size_t>(p) * 32 + q * 4);
        acc.x = fmaf(g.x, v, acc.x); acc.y = fmaf(g.y, v, acc.y); acc.z = fmaf(g.z, v, acc.z); acc.w = fmaf(g.w, v, acc.w);
    }
    float* base = part + static_cast<size_t>(blockIdx.x) * 32 * 27;
    base[(q * 4 + 0) * 27 + k] = acc.x; base[(q * 4 + 1) * 27 + k] = acc.y;
    base[(q * 4 + 2) * 27 + k] = acc.z; base[(q * 4 + 3) * 27 + k] = acc.w;
}

static int pix_chunks(long P, int min_per_chunk, int* ppc_out) {
    long ppc = (P + 591) / 592;
    if (ppc < min_per_chunk) ppc = min_per_chunk;
    *ppc_out = static_cast<int>(ppc);
    return static_cast<int>((P + ppc - 1) / ppc);
}

// row chunks of the 1x1 weight gradient: enough (tile x chunk) blocks for ~4 per SM, >= 64 rows per chunk
static int pw_wgrad_plan(long M, int Cout, int Cin, int* rows_per_chunk) {
    const long tiles = static_cast<long>((Cout + 63) / 64) * ((Cin + 63) / 64);
    long chunks = (148L * 4 + tiles - 1) / tiles;
    long rpc = (M + chunks - 1) / chunks;
    if (rpc < 64) rpc = 64;
    rpc = (rpc + 15) / 16 * 16;
    *rows_per_chunk = static_cast<int>(rpc);
    return static_cast<int>((M + rpc - 1) / rpc);
}

static int blocks_for(long total) {
    long b = (total + 255) / 256;
    if (b > 148 * 8) b = 148 * 8;
    return static_cast<int>(b < 1 ? 1 : b);
}

}  // namespace b200lp

using namespace b200lp;

extern "C" int32_t b200lp_transpose2d(const float* src, float* dst, int32_t R, int32_t C, void* stream) {
    B200LP_REQUIRE(src && dst && R > 0 && C > 0, "transpose2d: bad args");
    dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
    transpose2d_kernel<<<grid, block, 0, as_stream(stream)>>>(src, dst, R, C);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int64_t b200lp_pw_wgrad_workspace(int64_t M, int32_t Cin, int32_t Cout) {
    if (M <= 0 || Cin <= 0 || Cout <= 0) return B200LP_EINVAL;
    int rpc;
    return static_cast<int64_t>(pw_wgrad_plan(M, Cout, Cin, &rpc)) * Cout * Cin * 4;
}

extern "C" int32_t b200lp_pw_wgrad(const float* dy, const float* x, const float* in_scale, const float* in_shift,
                                   int32_t in_relu6, float* dw, int32_t accumulate, float* workspace,
                                   int64_t workspace_bytes, int64_t M, int32_t Cin, int32_t Cout, void* stream) {
    B200LP_REQUIRE(dy && x && dw && workspace && M > 0 && Cin > 0 && Cout > 0 && Cin % 4 == 0 && Cout % 4 == 0,
                   "pw_wgrad: bad args M=%lld Cin=%d Cout=%d (channels must be multiples of 4)", (long long)M, Cin, Cout);
    B200LP_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), "pw_wgrad: in_scale and in_shift go together");
    B200LP_REQUIRE(M < (1LL << 31) - 64, "pw_wgrad: M too large");
    int rpc;
    const int chunks = pw_wgrad_plan(M, Cout, Cin, &rpc);
    B200LP_REQUIRE(workspace_bytes >= static_cast<int64_t>(chunks) * Cout * Cin * 4, "pw_wgrad: workspace too small");
    cudaStream_t st = as_stream(stream);
    dim3 grid((Cin + 63) / 64, (Cout + 63) / 64, chunks);
    pw_wgrad_kernel<<<grid, 256, 0, st>>>(dy, x, in_scale, in_shift, in_relu6, workspace, static_cast<int>(M), Cout, Cin, rpc);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    const long total = static_cast<long>(Cout) * Cin;
    sum_parts_kernel<<<blocks_for(total), 256, 0, st>>>(workspace, dw, chunks, total, accumulate);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_dw_dgrad(const float* dy, const float* w, float* dx, int32_t N, int32_t H, int32_t W, int32_t C,
                                   int32_t stride, void* stream) {
    B200LP_REQUIRE(dy && w && dx && N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && (stride == 1 || stride == 2),
                   "dw_dgrad: bad args");
    const long total4 = static_cast<long>(N) * H * W * (C / 4);
    B200LP_REQUIRE(total4 < (1L << 31), "dw_dgrad: tensor too large");
    const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    dw_dgrad_kernel<<<blocks_for(total4), 256, 0, as_stream(stream)>>>(dy, w, dx, N, H, W, C, stride, Ho, Wo);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int64_t b200lp_dw_wgrad_workspace(int32_t N, int32_t H, int32_t W, int32_t C, int32_t stride) {
    if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || (stride != 1 && stride != 2)) return B200LP_EINVAL;
    const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    int ppc;
    return static_cast<int64_t>(pix_chunks(static_cast<long>(N) * Ho * Wo, 8, &ppc)) * 9 * C * 4;
}

extern "C" int32_t b200lp_dw_wgrad(const float* x, const float* in_scale, const float* in_shift, const float* dy, float* dw,
                                   int32_t accumulate, float* workspace, int64_t workspace_bytes, int32_t N, int32_t H,
                                   int32_t W, int32_t C, int32_t stride, void* stream) {
    B200LP_REQUIRE(x && in_scale && in_shift && dy && dw && workspace && N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 &&
                       C / 4 <= 256 && (stride == 1 || stride == 2), "dw_wgrad: bad args (C=%d)", C);
    B200LP_REQUIRE(static_cast<long>(N) * H * W < (1L << 31), "dw_wgrad: more than 2^31 pixels");
    const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    int ppc;
    const int chunks = pix_chunks(static_cast<long>(N) * Ho * Wo, 8, &ppc);
    B200LP_REQUIRE(workspace_bytes >= static_cast<int64_t>(chunks) * 9 * C * 4, "dw_wgrad: workspace too small");
    cudaStream_t st = as_stream(stream);
    const int qb = C / 4;
    const int lanes = 256 / qb;
    const int smem = lanes * qb * 9 * 16;
    static bool attr_set = false;
    if (!attr_set) {
        B200LP_CHECK_CUDA(cudaFuncSetAttribute(dw_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_set = true;
    }
    dw_wgrad_kernel<<<chunks, lanes * qb, smem, st>>>(x, in_scale, in_shift, dy, workspace, N, H, W, C, stride, Ho, Wo, ppc);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    dw_wgrad_reduce_kernel<<<blocks_for(static_cast<long>(C) * 9), 256, 0, st>>>(workspace, dw, chunks, C, accumulate);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int64_t b200lp_mbv2_stem_wgrad_workspace(int32_t N, int32_t H, int32_t W) {
    if (N <= 0 || H <= 0 || W <= 0) return B200LP_EINVAL;
    int ppc;
    return static_cast<int64_t>(pix_chunks(static_cast<long>(N) * ((H - 1) / 2 + 1) * ((W - 1) / 2 + 1), 64, &ppc)) * 32 * 27 * 4;
}

extern "C" int32_t b200lp_mbv2_stem_wgrad(const float* x_nchw, const float* dy, float* dw, int32_t accumulate,
                                          float* workspace, int64_t workspace_bytes, int32_t N, int32_t H, int32_t W,
                                          void* stream) {
    B200LP_REQUIRE(x_nchw && dy && dw && workspace && N > 0 && H > 0 && W > 0, "mbv2_stem_wgrad: bad args");
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long P = static_cast<long>(N) * Ho * Wo;
    B200LP_REQUIRE(P < (1L << 31), "mbv2_stem_wgrad: too many pixels");
    int ppc;
    const int chunks = pix_chunks(P, 64, &ppc);
    B200LP_REQUIRE(workspace_bytes >= static_cast<int64_t>(chunks) * 32 * 27 * 4, "mbv2_stem_wgrad: workspace too small");
    cudaStream_t st = as_stream(stream);
    mbv2_stem_wgrad_kernel<<<chunks, 224, 0, st>>>(x_nchw, dy, workspace, N, H, W, Ho, Wo, ppc);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    sum_parts_kernel<<<blocks_for(32 * 27), 256, 0, st>>>(workspace, dw, chunks, 32 * 27, accumulate);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}
