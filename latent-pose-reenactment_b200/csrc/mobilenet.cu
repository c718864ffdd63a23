// Pose-encoder (torchvision MobileNetV2) forward as hand-written FP32 kernels.
//
// Replaces, for `embedders/unsupervised_pose_separate_embResNeXt_segmentation.py:56-58` (Embedder.get_pose_embedding,
// the only embedder call of drive.py:87 and of every fine-tuning step), torchvision's
//     Conv2d(1x1 | depthwise 3x3 | 3x3 stem) -> BatchNorm2d(eps 1e-5, momentum 0.1) -> ReLU6      x 52 layers
// i.e. ~300 cuDNN / ATen launches (13 us per BatchNorm call) for 0.39 GMAC per image, by ~125 launches:
//
//   conv kernel  : reads the PREVIOUS layer's raw conv output and applies that layer's BatchNorm + ReLU6 on load
//                  (per-channel scale / shift), writes ITS raw output once and the per-channel sum / sum-of-squares
//                  partials of it (train-mode batch statistics) from the same registers
//   bn_finalize  : partials -> batch mean / biased variance (fp64 merge) -> scale = gamma*rstd, shift = beta-mean*scale,
//                  running_mean / running_var (unbiased) / num_batches_tracked updates; eval mode: from running stats
//   bn_apply     : block outputs (linear bottleneck + residual) are materialised once: y = x*scale + shift (+ skip)
//
// Everything is FP32 on the CUDA cores (these are K <= 960 contractions on <= 128x128 planes; the pose vector feeds
// every AdaIN gain of the generator, whose RGB output must stay within 1e-3 of the fp32 reference).
// Activations NHWC; statistics partials are [part][2][C] (sum row, then sum-of-squares row), merged in a fixed order:
// results are bit-reproducible run to run (no atomics).
#include "common.cuh"

namespace b200lp {

__device__ __forceinline__ float relu6f(float v) { return fminf(fmaxf(v, 0.f), 6.f); }

// ------------------------------------------------------------------------------------------------ pointwise conv
// y[m][n] = sum_k f(x[m][k]) * w[n][k] (+ bias[n]);  f = identity | BN | ReLU6(BN) of the producer layer.
// SIMT SGEMM: BM x 64 tile (BM = 128 when the grid still fills the GPU, else 64), K step 16, 256 threads,
// (BM/16) x 4 outputs per thread; the NEXT TWO K tiles travel global -> registers while the current one is multiplied
// (two register sets: the small layers run one 8-warp block per SM, so nothing else hides the load latency — with one
// tile in flight an iteration cost ~1000 clk against ~320 clk of FMAs; a 32-row tile variant was measured slower).
constexpr int kPwBN = 64, kPwBK = 16;

template <int BM>
__global__ void __launch_bounds__(256)
pw_conv_kernel(const float* __restrict__ x, const float* __restrict__ in_scale, const float* __restrict__ in_shift,
               int in_relu6, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y,
               float* __restrict__ part, int M, int K, int N, int k_per_split, float* __restrict__ ws) {
    constexpr int TM = BM / 16;          // rows per thread
    constexpr int AL = BM / 64;          // A float4 loads per thread and K tile
    __shared__ float As[kPwBK][BM + 4];
    __shared__ float Bs[kPwBK][kPwBN + 4];
    __shared__ float red[2][16][kPwBN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * kPwBN;
    // split-K (gridDim.z > 1): this block contracts k in [kb, ke) and leaves raw partial sums in ws[z][M][N]; bias, output
    // and statistics are produced by pw_splitk_reduce_kernel
    const int kb = blockIdx.z * k_per_split;
    const int ke = min(K, kb + k_per_split);
    // loader mapping: one float4 (4 consecutive k) of one row per thread (AL rows of A, one of B)
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float4 ra0[AL], rb0, ra1[AL], rb1;
    // raw loads only: the producer's BatchNorm + ReLU6 is applied when the tile is STORED to shared memory, one or two
    // iterations later, so that no instruction waits on the load where it is issued
    auto load_tile = [&](int k0, float4 (&ra)[AL], float4& rb) {
        const int k = kb + k0 + lk;
        const bool kin = k < ke;      // K % 4 == 0 and k_per_split % 16 == 0: a float4 is entirely inside or outside
#pragma unroll
        for (int q = 0; q < AL; ++q) {
            const int m = m0 + lrow + 64 * q;
            ra[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kin && m < M) ra[q] = __ldg(reinterpret_cast<const float4*>(x + static_cast<size_t>(m) * K + k));
        }
        rb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kin && n0 + lrow < N) rb = __ldg(reinterpret_cast<const float4*>(w + static_cast<size_t>(n0 + lrow) * K + k));
    };
    auto store_tile = [&](int k0, const float4 (&ra)[AL], const float4& rb) {
        const int k = kb + k0 + lk;
        const bool kin = k < ke;
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kin && in_scale) {
            sc = __ldg(reinterpret_cast<const float4*>(in_scale + k));
            sh = __ldg(reinterpret_cast<const float4*>(in_shift + k));
        }
#pragma unroll
        for (int q = 0; q < AL; ++q) {
            float4 a = ra[q];
            if (in_scale && kin && m0 + lrow + 64 * q < M) {      // rows / columns outside the matrix stay zero
                a.x = a.x * sc.x + sh.x; a.y = a.y * sc.y + sh.y; a.z = a.z * sc.z + sh.z; a.w = a.w * sc.w + sh.w;
                if (in_relu6) { a.x = relu6f(a.x); a.y = relu6f(a.y); a.z = relu6f(a.z); a.w = relu6f(a.w); }
            }
            As[lk + 0][lrow + 64 * q] = a.x; As[lk + 1][lrow + 64 * q] = a.y;
            As[lk + 2][lrow + 64 * q] = a.z; As[lk + 3][lrow + 64 * q] = a.w;
        }
        Bs[lk + 0][lrow] = rb.x; Bs[lk + 1][lrow] = rb.y; Bs[lk + 2][lrow] = rb.z; Bs[lk + 3][lrow] = rb.w;
    };

    auto compute = [&]() {
#pragma unroll
        for (int kk = 0; kk < kPwBK; ++kk) {
            float ar[TM];
#pragma unroll
            for (int q = 0; q < TM / 4; ++q) {
                const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * TM + 4 * q]);
                ar[4 * q] = av.x; ar[4 * q + 1] = av.y; ar[4 * q + 2] = av.z; ar[4 * q + 3] = av.w;
            }
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
    };
    const int nk = (ke - kb + kPwBK - 1) / kPwBK;
    // iteration `it`: shared memory holds tile it, `rs` holds tile it + 1 (loaded one iteration ago), `rl` is free
    auto step = [&](float4 (&ral)[AL], float4& rbl, const float4 (&ras)[AL], const float4& rbs, int it) {
        if (it + 2 < nk) load_tile((it + 2) * kPwBK, ral, rbl);
        compute();
        __syncthreads();
        if (it + 1 < nk) {
            store_tile((it + 1) * kPwBK, ras, rbs);
            __syncthreads();
        }
    };
    load_tile(0, ra0, rb0);
    store_tile(0, ra0, rb0);
    if (nk > 1) load_tile(kPwBK, ra1, rb1);
    __syncthreads();
    for (int it = 0; it < nk; it += 2) {
        step(ra0, rb0, ra1, rb1, it);
        if (it + 1 < nk) step(ra1, rb1, ra0, rb0, it + 1);
    }

    const int n = n0 + tx * 4;
    if (ws) {
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int m = m0 + ty * TM + i;
            if (m < M && n < N)
                *reinterpret_cast<float4*>(ws + (static_cast<size_t>(blockIdx.z) * M + m) * N + n) =
                    make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
        return;
    }
    float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m < M && n < N) {      // N % 4 == 0
            float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
#pragma unroll
            for (int j = 0; j < 4; ++j) { cs[j] += acc[i][j]; cq[j] += acc[i][j] * acc[i][j]; }
            if (bias) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + n));
                o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
            }
            *reinterpret_cast<float4*>(y + static_cast<size_t>(m) * N + n) = o;
        }
    }
    if (part) {      // batch-statistics partials of this row tile (rows / columns outside the matrix contribute 0)
#pragma unroll
        for (int j = 0; j < 4; ++j) { red[0][ty][tx * 4 + j] = cs[j]; red[1][ty][tx * 4 + j] = cq[j]; }
        __syncthreads();
        if (tid < 2 * kPwBN) {
            const int which = tid >> 6, c = tid & 63;
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < 16; ++r) s += red[which][r][c];
            if (n0 + c < N) part[(static_cast<size_t>(blockIdx.x) * 2 + which) * N + n0 + c] = s;
        }
    }
}

// y = sum_z ws[z] (+ bias) and the statistics partials of one BM-row tile, splits summed in index order (deterministic).
// block = 64 columns x 4 row lanes; grid = (row tiles of `bm`, column tiles of 64)
__global__ void __launch_bounds__(256)
pw_splitk_reduce_kernel(const float* __restrict__ ws, const float* __restrict__ bias, float* __restrict__ y,
                        float* __restrict__ part, int splits, int M, int N, int bm) {
    __shared__ float red[2][4][kPwBN];
    const int c = threadIdx.x & 63, rl = threadIdx.x >> 6;
    const int n = blockIdx.y * kPwBN + c;
    const int m0 = blockIdx.x * bm;
    float s1 = 0.f, s2 = 0.f;
    if (n < N) {
        const float b = bias ? __ldg(bias + n) : 0.f;
        for (int r = rl; r < bm && m0 + r < M; r += 4) {
            const size_t off = static_cast<size_t>(m0 + r) * N + n;
            float v = 0.f;
#pragma unroll 8
            for (int z = 0; z < splits; ++z) v += __ldg(ws + static_cast<size_t>(z) * M * N + off);   // independent loads, fixed order
            y[off] = v + b;
            s1 += v;
            s2 += v * v;
        }
    }
    if (part) {
        red[0][rl][c] = s1;
        red[1][rl][c] = s2;
        __syncthreads();
        if (threadIdx.x < 2 * kPwBN) {
            const int which = threadIdx.x >> 6, cc = threadIdx.x & 63;
            const float t = ((red[which][0][cc] + red[which][1][cc]) + red[which][2][cc]) + red[which][3][cc];
            if (blockIdx.y * kPwBN + cc < N) part[(static_cast<size_t>(blockIdx.x) * 2 + which) * N + blockIdx.y * kPwBN + cc] = t;
        }
    }
}

// K splits of a layer whose (row tile x column tile) grid leaves most SMs idle: the small planes of a batch of 8 ran 8..64
// blocks, each walking K = 384..2048 serially at ~1 us per 16-wide step (68 us for 512 x 960 -> 160, 78 us for the 8-row
// classifier).  At least 64 k per split, at most one wave of blocks.
static int pw_splits(long M, int N, int K, int bm, int* k_per_split) {
    const long tiles = ((M + bm - 1) / bm) * ((N + kPwBN - 1) / kPwBN);
    int s = 1;
    if (tiles * 2 <= 148 && K >= 128) {
        s = static_cast<int>(148 / tiles);
        if (s > K / 64) s = K / 64;
        if (s < 1) s = 1;
    }
    int kps = ((K + s - 1) / s + kPwBK - 1) / kPwBK * kPwBK;
    *k_per_split = kps;
    return (K + kps - 1) / kps;
}

// row-tile height: 128 while (row tiles x column tiles) still covers the 148 SMs, else 64
static int pw_tile_m(long M, int N) {
    const long blocks128 = ((M + 127) / 128) * ((N + kPwBN - 1) / kPwBN);
    return blocks128 >= 148 ? 128 : 64;
}

// ------------------------------------------------------------------------------------------------ depthwise 3x3
// y[n,ho,wo,c] = sum_tap relu6(x[n, ho*s+kh-1, wo*s+kw-1, c]*scale[c] + shift[c]) * w[c][tap]   (zero padding AFTER the
// activation, as in the reference where the padded tensor is the ReLU6 output).
// A block is `ppi` pixel lanes x Q = C/4 channel quads (blockDim = ppi*Q <= 256: every lane works whatever C is — with
// one warp per 32 quads the C = 32 layer ran 8 of 32 lanes); thread (ps, q) keeps quad q and walks pixels
// p0 + ps, + 2 ppi, ... two at a time.  grid = pixel chunks.
__global__ void __launch_bounds__(256, 2)
dw_conv3x3_kernel(const float* __restrict__ x, const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                  const float* __restrict__ w, float* __restrict__ y, float* __restrict__ part, int N, int H, int W,
                  int C, int stride, int Ho, int Wo, int pix_per_chunk, int c_base) {
    __shared__ float4 red[2][256];
    const int Qb = (C >> 2) - c_base > 256 ? 256 : (C >> 2) - c_base;     // quads handled by this launch slice
    const int ppi = blockDim.x / Qb;
    const int q = threadIdx.x % Qb, ps = threadIdx.x / Qb;
    const int c = (c_base + q) * 4;
    float4 wt[9];
#pragma unroll
    for (int t = 0; t < 9; ++t)
        wt[t] = make_float4(__ldg(w + (c + 0) * 9 + t), __ldg(w + (c + 1) * 9 + t), __ldg(w + (c + 2) * 9 + t),
                            __ldg(w + (c + 3) * 9 + t));
    const float4 sc = __ldg(reinterpret_cast<const float4*>(in_scale + c));
    const float4 sh = __ldg(reinterpret_cast<const float4*>(in_shift + c));
    const int P = N * Ho * Wo;                       // < 2^31 (checked by the host)
    const int p0 = blockIdx.x * pix_per_chunk;
    const int p1 = min(p0 + pix_per_chunk, P);
    const int HoWo = Ho * Wo;
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = make_float4(0.f, 0.f, 0.f, 0.f);
    // two output pixels per iteration, all 18 (predicated) loads issued before the first use: with a branch per tap
    // the loads serialised on the L2 latency (4.5 us per pixel per warp measured)
    for (int pb = p0 + ps; pb < p1; pb += 2 * ppi) {
        float4 v[2][9];
        bool ok[2][9];
        bool pvalid[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int p = pb + ppi * u;
            pvalid[u] = p < p1;
            const int pp = pvalid[u] ? p : p0;
            const int n = pp / HoWo;
            const int r = pp - n * HoWo;
            const int ho = r / Wo, wo = r - ho * Wo;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int hi = ho * stride + kh - 1;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int wi = wo * stride + kw - 1;
                    const bool in = pvalid[u] && hi >= 0 && hi < H && wi >= 0 && wi < W;
                    ok[u][kh * 3 + kw] = in;
                    const size_t off = (static_cast<size_t>(n * H + (in ? hi : 0)) * W + (in ? wi : 0)) * C + c;
                    v[u][kh * 3 + kw] = __ldg(reinterpret_cast<const float4*>(x + off));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (!pvalid[u]) continue;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                if (!ok[u][t]) continue;       // zero padding applies to the activation, not to the raw input
                const float4 g = v[u][t];
                const float4 k = wt[t];
                a.x = fmaf(relu6f(g.x * sc.x + sh.x), k.x, a.x);
                a.y = fmaf(relu6f(g.y * sc.y + sh.y), k.y, a.y);
                a.z = fmaf(relu6f(g.z * sc.z + sh.z), k.z, a.z);
                a.w = fmaf(relu6f(g.w * sc.w + sh.w), k.w, a.w);
            }
            *reinterpret_cast<float4*>(y + static_cast<size_t>(pb + ppi * u) * C + c) = a;
            s1.x += a.x; s1.y += a.y; s1.z += a.z; s1.w += a.w;
            s2.x += a.x * a.x; s2.y += a.y * a.y; s2.z += a.z * a.z; s2.w += a.w * a.w;
        }
    }
    if (part) {
        red[0][threadIdx.x] = s1;
        red[1][threadIdx.x] = s2;
        __syncthreads();
        if (ps == 0) {          // pixel lane 0 merges its quad's partial sums in a fixed order
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
    }
}

// ------------------------------------------------------------------------------------------------ stem 3x3 stride 2
// x NCHW (N,3,H,W) -> y NHWC (N,H/2,W/2,32) raw; thread = (output pixel, 8-channel group), block = 64 pixels x 4 groups
__global__ void __launch_bounds__(256)
mbv2_stem_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
                 float* __restrict__ part, int N, int H, int W, int Ho, int Wo) {
    __shared__ float sw[27][32];          // [c*9 + tap][co]
    __shared__ float red[2][64][33];
    for (int i = threadIdx.x; i < 27 * 32; i += 256) {
        const int co = i / 27, t = i - co * 27;      // w is [32][3][3][3] = [co][c*9 + tap]
        sw[t][co] = __ldg(w + i);
    }
    __syncthreads();
    const int g = threadIdx.x & 3, pl = threadIdx.x >> 2;
    const long P = static_cast<long>(N) * Ho * Wo;
    const long p = static_cast<long>(blockIdx.x) * 64 + pl;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const bool valid = p < P;
    if (valid) {
        const int wo = static_cast<int>(p % Wo);
        const int ho = static_cast<int>((p / Wo) % Ho);
        const long n = p / (static_cast<long>(Wo) * Ho);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int hi = ho * 2 + kh - 1;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int wi = wo * 2 + kw - 1;
                    float v = 0.f;
                    if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = __ldg(x + ((n * 3 + c) * H + hi) * W + wi);
                    const float* wr = &sw[c * 9 + kh * 3 + kw][g * 8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, wr[j], acc[j]);
                }
            }
        }
        float* o = y + p * 32 + g * 8;
        *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    if (part) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            red[0][pl][g * 8 + j] = valid ? acc[j] : 0.f;
            red[1][pl][g * 8 + j] = valid ? acc[j] * acc[j] : 0.f;
        }
        __syncthreads();
        if (threadIdx.x < 64) {
            const int which = threadIdx.x >> 5, c = threadIdx.x & 31;
            float s = 0.f;
            for (int r = 0; r < 64; ++r) s += red[which][r][c];
            part[(static_cast<size_t>(blockIdx.x) * 2 + which) * 32 + c] = s;
        }
    }
}

// ------------------------------------------------------------------------------------------------ BatchNorm pieces
// training: mean / biased var of `count` samples from [nparts][2][C] partials; eval: running statistics.
// grid = ceil(C / 32); block = 32 channels x 32 partial lanes (the stem and the 128 x 128 layers have 2048 partials).
__global__ void __launch_bounds__(1024)
bn_finalize_kernel(const float* __restrict__ part, int nparts, double count, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ running_mean, float* __restrict__ running_var,
                   long long* __restrict__ num_batches_tracked, float momentum, float eps, float* __restrict__ scale,
                   float* __restrict__ shift, float* __restrict__ mean_out, float* __restrict__ rstd_out, int C,
                   int training) {
    __shared__ double red[2][32][33];
    const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    if (training) {
        double s1 = 0.0, s2 = 0.0;
        if (c < C)
#pragma unroll 8
            for (int i = pl; i < nparts; i += 32) {
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
            const double mean = a / count;
            double var = b / count - mean * mean;
            if (var < 0.0) var = 0.0;
            const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
            const float sc = __ldg(gamma + c) * rstd;
            scale[c] = sc;
            shift[c] = __ldg(beta + c) - static_cast<float>(mean) * sc;
            if (mean_out) { mean_out[c] = static_cast<float>(mean); rstd_out[c] = rstd; }
            if (running_mean) {
                const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
                running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
            }
        }
        if (num_batches_tracked && blockIdx.x == 0 && threadIdx.x == 0) *num_batches_tracked += 1;
    } else if (pl == 0 && c < C) {
        // 1/sqrt in double like the training branch keeps eval and train numerics on the same footing
        const float rs = static_cast<float>(1.0 / sqrt(static_cast<double>(running_var[c]) + static_cast<double>(eps)));
        const float sc = __ldg(gamma + c) * rs;
        scale[c] = sc;
        shift[c] = __ldg(beta + c) - running_mean[c] * sc;
        if (mean_out) { mean_out[c] = running_mean[c]; rstd_out[c] = rs; }
    }
}

// y = x*scale[c] + shift[c] (+ residual) (relu6)     — float4 over [M][C]
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float4* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                const float4* __restrict__ residual, float4* __restrict__ y, long total4, int C4, int relu6) {
    for (long i = blockIdx.x * 256L + threadIdx.x; i < total4; i += gridDim.x * 256L) {
        const int c = static_cast<int>(i % C4) * 4;
        float4 v = __ldg(x + i);
        const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + c));
        const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + c));
        v.x = v.x * sc.x + sh.x; v.y = v.y * sc.y + sh.y; v.z = v.z * sc.z + sh.z; v.w = v.w * sc.w + sh.w;
        if (residual) {
            const float4 r = __ldg(residual + i);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        if (relu6) { v.x = relu6f(v.x); v.y = relu6f(v.y); v.z = relu6f(v.z); v.w = relu6f(v.w); }
        y[i] = v;
    }
}

// y[n][c] = mean over HW of relu6(x[n,p,c]*scale[c] + shift[c])      (features.18 BN + ReLU6 + adaptive_avg_pool2d(1))
// grid = (channel-quad groups, N); block = 32 lanes x 8 warps (warps split the pixels)
__global__ void __launch_bounds__(256)
bn_relu6_avgpool_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                        float* __restrict__ y, int HW, int C) {
    __shared__ float4 red[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + lane) * 4;
    const long n = blockIdx.y;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + c));
        const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + c));
        for (int p = warp; p < HW; p += 8) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(x + (n * HW + p) * C + c));
            s.x += relu6f(v.x * sc.x + sh.x); s.y += relu6f(v.y * sc.y + sh.y);
            s.z += relu6f(v.z * sc.z + sh.z); s.w += relu6f(v.w * sc.w + sh.w);
        }
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

static int dw_plan(long P, int* pix_per_chunk) {
    long ppc = (P + 1183) / 1184;          // <= 148 x 8 chunks
    // >= 8 pixels per chunk (was 32: the 8 x 8 planes of a batch of 8 ran as 16 blocks whose single pixel lane — C = 960
    // fills the block with channel quads — walked its 32 pixels serially: 38 us per layer)
    if (ppc < 8) ppc = 8;
    *pix_per_chunk = static_cast<int>(ppc);
    return static_cast<int>((P + ppc - 1) / ppc);
}

}  // namespace b200lp

using namespace b200lp;

extern "C" int32_t b200lp_pw_conv_parts(int64_t M, int32_t Cout) {
    const int bm = pw_tile_m(M, Cout);
    return static_cast<int32_t>((M + bm - 1) / bm);
}

extern "C" int64_t b200lp_pw_conv_workspace(int64_t M, int32_t Cin, int32_t Cout) {
    if (M <= 0 || Cin <= 0 || Cout <= 0) return B200LP_EINVAL;
    int kps;
    const int sp = pw_splits(M, Cout, Cin, pw_tile_m(M, Cout), &kps);
    return sp > 1 ? static_cast<int64_t>(sp) * M * Cout * 4 : 0;
}

extern "C" int32_t b200lp_pw_conv_ws(const float* x, const float* in_scale, const float* in_shift, int32_t in_relu6,
                                     const float* w, const float* bias, float* y, float* part, int64_t M, int32_t Cin,
                                     int32_t Cout, float* workspace, int64_t workspace_bytes, void* stream) {
    B200LP_REQUIRE(x && w && y && M > 0 && Cin > 0 && Cout > 0 && Cin % 4 == 0 && Cout % 4 == 0,
                   "pw_conv: bad args M=%lld Cin=%d Cout=%d (channels must be multiples of 4)", (long long)M, Cin, Cout);
    B200LP_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), "pw_conv: in_scale and in_shift go together");
    B200LP_REQUIRE(M < (1LL << 31) - 128, "pw_conv: M too large");
    const int bm = pw_tile_m(M, Cout);
    int kps;
    int sp = pw_splits(M, Cout, Cin, bm, &kps);
    if (sp > 1 && (!workspace || workspace_bytes < static_cast<int64_t>(sp) * M * Cout * 4)) { sp = 1; kps = (Cin + kPwBK - 1) / kPwBK * kPwBK; }
    float* ws = sp > 1 ? workspace : nullptr;
    dim3 grid(static_cast<unsigned>((M + bm - 1) / bm), static_cast<unsigned>((Cout + kPwBN - 1) / kPwBN), sp);
    cudaStream_t st = as_stream(stream);
    if (bm == 128)
        pw_conv_kernel<128><<<grid, 256, 0, st>>>(x, in_scale, in_shift, in_relu6, w, bias, y, part, static_cast<int>(M), Cin,
                                                   Cout, kps, ws);
    else
        pw_conv_kernel<64><<<grid, 256, 0, st>>>(x, in_scale, in_shift, in_relu6, w, bias, y, part, static_cast<int>(M), Cin,
                                                  Cout, kps, ws);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    if (sp > 1) {
        dim3 g2(grid.x, grid.y);
        pw_splitk_reduce_kernel<<<g2, 256, 0, st>>>(ws, bias, y, part, sp, static_cast<int>(M), Cout, bm);
        B200LP_CHECK_CUDA(cudaGetLastError());
        count_launch();
    }
    return B200LP_OK;
}

extern "C" int32_t b200lp_pw_conv(const float* x, const float* in_scale, const float* in_shift, int32_t in_relu6,
                                  const float* w, const float* bias, float* y, float* part, int64_t M, int32_t Cin,
                                  int32_t Cout, void* stream) {
    return b200lp_pw_conv_ws(x, in_scale, in_shift, in_relu6, w, bias, y, part, M, Cin, Cout, nullptr, 0, stream);
}

extern "C" int32_t b200lp_dw_conv3x3_parts(int32_t N, int32_t H, int32_t W, int32_t stride) {
    if (N <= 0 || H <= 0 || W <= 0 || (stride != 1 && stride != 2)) return B200LP_EINVAL;
    const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    int ppc;
    return dw_plan(static_cast<long>(N) * Ho * Wo, &ppc);
}

extern "C" int32_t b200lp_dw_conv3x3(const float* x, const float* in_scale, const float* in_shift, const float* w,
                                     float* y, float* part, int32_t N, int32_t H, int32_t W, int32_t C, int32_t stride,
                                     void* stream) {
    B200LP_REQUIRE(x && in_scale && in_shift && w && y && N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 &&
                       (stride == 1 || stride == 2),
                   "dw_conv3x3: bad args N=%d H=%d W=%d C=%d stride=%d", N, H, W, C, stride);
    const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;     // kernel 3, padding 1
    B200LP_REQUIRE(static_cast<long>(N) * H * W < (1L << 31), "dw_conv3x3: more than 2^31 pixels");
    int ppc;
    const int chunks = dw_plan(static_cast<long>(N) * Ho * Wo, &ppc);
    // channel quads are cut into slices of <= 256 (one launch each; every MobileNetV2 layer has C/4 <= 240)
    for (int c_base = 0; c_base < C / 4; c_base += 256) {
        const int qb = C / 4 - c_base > 256 ? 256 : C / 4 - c_base;
        const int threads = (256 / qb) * qb;
        dw_conv3x3_kernel<<<chunks, threads, 0, as_stream(stream)>>>(x, in_scale, in_shift, w, y, part, N, H, W, C, stride,
                                                                     Ho, Wo, ppc, c_base);
        B200LP_CHECK_CUDA(cudaGetLastError());
        count_launch();
    }
    return B200LP_OK;
}

extern "C" int32_t b200lp_mbv2_stem_parts(int32_t N, int32_t H, int32_t W) {
    if (N <= 0 || H <= 0 || W <= 0) return B200LP_EINVAL;
    const long P = static_cast<long>(N) * ((H - 1) / 2 + 1) * ((W - 1) / 2 + 1);
    return static_cast<int32_t>((P + 63) / 64);
}

extern "C" int32_t b200lp_mbv2_stem(const float* x_nchw, const float* w, float* y_nhwc, float* part, int32_t N, int32_t H,
                                    int32_t W, void* stream) {
    B200LP_REQUIRE(x_nchw && w && y_nhwc && N > 0 && H > 0 && W > 0, "mbv2_stem: bad args");
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long P = static_cast<long>(N) * Ho * Wo;
    mbv2_stem_kernel<<<static_cast<unsigned>((P + 63) / 64), 256, 0, as_stream(stream)>>>(x_nchw, w, y_nhwc, part, N, H, W,
                                                                                          Ho, Wo);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_bn_finalize(const float* part, int32_t nparts, int64_t count, const float* gamma,
                                      const float* beta, float* running_mean, float* running_var,
                                      int64_t* num_batches_tracked, float momentum, float eps, float* scale,
                                      float* shift, float* mean_out, float* rstd_out, int32_t C, int32_t training,
                                      void* stream) {
    B200LP_REQUIRE(gamma && beta && scale && shift && C > 0, "bn_finalize: bad args");
    B200LP_REQUIRE((mean_out == nullptr) == (rstd_out == nullptr), "bn_finalize: mean_out and rstd_out go together");
    B200LP_REQUIRE(training ? (part && nparts > 0 && count > 0) : (running_mean && running_var),
                   "bn_finalize: training needs partials, eval needs running statistics");
    bn_finalize_kernel<<<(C + 31) / 32, 1024, 0, as_stream(stream)>>>(
        part, nparts, static_cast<double>(count), gamma, beta, running_mean, running_var,
        reinterpret_cast<long long*>(num_batches_tracked), momentum, eps, scale, shift, mean_out, rstd_out, C, training);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_bn_apply(const float* x, const float* scale, const float* shift, const float* residual,
                                   float* y, int64_t M, int32_t C, int32_t relu6, void* stream) {
    B200LP_REQUIRE(x && scale && shift && y && M > 0 && C > 0 && C % 4 == 0, "bn_apply: bad args");
    const long total4 = M * (C / 4);
    long blocks = (total4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    bn_apply_kernel<<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(x), scale, shift, reinterpret_cast<const float4*>(residual),
        reinterpret_cast<float4*>(y), total4, C / 4, relu6);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_bn_relu6_avgpool(const float* x, const float* scale, const float* shift, float* y, int32_t N,
                                           int32_t HW, int32_t C, void* stream) {
    B200LP_REQUIRE(x && scale && shift && y && N > 0 && HW > 0 && C > 0 && C % 4 == 0, "bn_relu6_avgpool: bad args");
    dim3 grid((C / 4 + 31) / 32, N);
    bn_relu6_avgpool_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, scale, shift, y, HW, C);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}
