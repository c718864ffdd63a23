// Scalar-loss reductions and the small dense pieces around them — the last torch ops of the training step.
//
//   dice          criterions/dice.py:30-34        -log(2 sum(f*r) / (sum f^2 + sum r^2)) * w      (f broadcast over r's channels)
//   adversarial   criterions/adversarial.py:42-47 hinge discriminator loss, -mean generator loss   ((B,) score vectors)
//   crop          criterions/idt_embed.py:62-83   affine_grid + grid_sample(bilinear, reflection, align_corners=False) of a box
//   disc. head    discriminators/no_landmarks.py:101-105   relu -> spatial sum -> SN-linear(512 -> 1) + <feat, embed[label]>
//
// All reductions are two-stage with a fixed order (bit-reproducible); the backward kernels are gathers (no atomics).
#include "common.cuh"

namespace b200lp {

namespace {
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// block-wide sum of three values (256 threads); result valid in thread 0
__device__ __forceinline__ void block_sum3(float& a, float& b, float& c) {
    __shared__ float red[3][8];
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[0][warp] = a; red[1][warp] = b; red[2][warp] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = b = c = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { a += red[0][i]; b += red[1][i]; c += red[2][i]; }
    }
}
}  // namespace

// ------------------------------------------------------------------------------------------------ dice
// f (B, HW), r (B, CR, HW):  part[block] = (sum f*r, sum f^2, sum r^2) over the block's slice of (b, p)
__global__ void __launch_bounds__(256)
dice_partial_kernel(const float* __restrict__ f, const float* __restrict__ r, float* __restrict__ part, int B, int CR,
                    int HW) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    const unsigned total = static_cast<unsigned>(B) * HW;
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < total; i += gridDim.x * 256u) {
        const unsigned b = i / static_cast<unsigned>(HW), p = i - b * HW;
        const float fv = __ldg(f + i);
        s1 += fv * fv;
        for (int c = 0; c < CR; ++c) {
            const float rv = __ldg(r + (static_cast<size_t>(b) * CR + c) * HW + p);
            s0 += fv * rv;
            s2 += rv * rv;
        }
    }
    block_sum3(s0, s1, s2);
    if (threadIdx.x == 0) { part[blockIdx.x * 3 + 0] = s0; part[blockIdx.x * 3 + 1] = s1; part[blockIdx.x * 3 + 2] = s2; }
}

// sums[0..2] = totals (fp64 merge), loss[0] = -log(2 s0 / (s1 + s2)) * weight
__global__ void dice_final_kernel(const float* __restrict__ part, int nparts, float weight, float* __restrict__ sums,
                                  float* __restrict__ loss) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = 0; i < nparts; ++i) { a += part[i * 3]; b += part[i * 3 + 1]; c += part[i * 3 + 2]; }
    sums[0] = static_cast<float>(a); sums[1] = static_cast<float>(b); sums[2] = static_cast<float>(c);
    loss[0] = static_cast<float>(-log(2.0 * a / (b + c)) * static_cast<double>(weight));
}

// df[b,p] = g * (-w) * ( sum_c r[b,c,p] / s0 - 2 f[b,p] / (s1 + s2) )        (d/df of -w (log(2 s0) - log(s1 + s2)))
__global__ void __launch_bounds__(256)
dice_bwd_kernel(const float* __restrict__ f, const float* __restrict__ r, const float* __restrict__ sums,
                const float* __restrict__ g, float weight, float* __restrict__ df, int B, int CR, int HW) {
    const float gs = -weight * __ldg(g);
    const float inv_n = 1.f / __ldg(sums), inv_d = 2.f / (__ldg(sums + 1) + __ldg(sums + 2));
    const unsigned total = static_cast<unsigned>(B) * HW;
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < total; i += gridDim.x * 256u) {
        const unsigned b = i / static_cast<unsigned>(HW), p = i - b * HW;
        float rs = 0.f;
        for (int c = 0; c < CR; ++c) rs += __ldg(r + (static_cast<size_t>(b) * CR + c) * HW + p);
        df[i] = gs * (rs * inv_n - __ldg(f + i) * inv_d);
    }
}

// ------------------------------------------------------------------------------------------------ adversarial (hinge)
// out[0] = loss_G, out[1] = loss_D over (B,) score vectors.  relativistic = 0: gan; 1: rgan; 2: ragan
// (get_dis_preds, criterions/adversarial.py:20-29).  One block.
__global__ void adversarial_fwd_kernel(const float* __restrict__ fake_g, const float* __restrict__ fake_d,
                                       const float* __restrict__ real, float* __restrict__ out, int B, int relativistic) {
    __shared__ float mean_real, mean_fd, mean_fg;
    float a = 0.f, b = 0.f, c = 0.f;
    for (int i = threadIdx.x; i < B; i += blockDim.x) { a += real[i]; b += fake_d[i]; c += fake_g[i]; }
    block_sum3(a, b, c);
    if (threadIdx.x == 0) { mean_real = a / B; mean_fd = b / B; mean_fg = c / B; }
    __syncthreads();
    float lg = 0.f, ld = 0.f, z = 0.f;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        float rp_d, fp_d, rp_g, fp_g;      // (real_pred, fake_pred) as the D loss / the G loss see them
        if (relativistic == 0) { rp_d = real[i]; fp_d = fake_d[i]; rp_g = real[i]; fp_g = fake_g[i]; }
        else if (relativistic == 1) { rp_d = real[i] - fake_d[i]; fp_d = fake_d[i] - real[i]; rp_g = real[i] - fake_g[i]; fp_g = fake_g[i] - real[i]; }
        else { rp_d = real[i] - mean_fd; fp_d = fake_d[i] - mean_real; rp_g = real[i] - mean_fg; fp_g = fake_g[i] - mean_real; }
        ld += fmaxf(1.f - rp_d, 0.f) + fmaxf(1.f + fp_d, 0.f);
        lg += relativistic == 0 ? -fp_g : fmaxf(1.f + rp_g, 0.f) + fmaxf(1.f - fp_g, 0.f);
    }
    block_sum3(lg, ld, z);
    if (threadIdx.x == 0) { out[0] = lg / B; out[1] = ld / B; }
}

// gan type only (the shipped configs): d fake_g = -gG / B;  d real = -gD/B [1 - real > 0];  d fake_d = gD/B [1 + fake_d > 0]
__global__ void adversarial_bwd_kernel(const float* __restrict__ fake_d, const float* __restrict__ real,
                                       const float* __restrict__ g_g, const float* __restrict__ g_d,
                                       float* __restrict__ d_fake_g, float* __restrict__ d_fake_d, float* __restrict__ d_real,
                                       int B) {
    const float gg = g_g ? __ldg(g_g) / B : 0.f, gd = g_d ? __ldg(g_d) / B : 0.f;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        if (d_fake_g) d_fake_g[i] = -gg;
        if (d_real) d_real[i] = (1.f - real[i] > 0.f) ? -gd : 0.f;
        if (d_fake_d) d_fake_d[i] = (1.f + fake_d[i] > 0.f) ? gd : 0.f;
    }
}

// ------------------------------------------------------------------------------------------------ box crop (grid_sample)
// torch semantics: affine_grid(theta, align_corners=False) with theta from the box [t, b, l, r] (pixels), then
// grid_sample(bilinear, padding_mode='reflection', align_corners=False).
__device__ __forceinline__ float reflect_coord(float x, int size) {
    // reflect about -0.5 and size - 0.5 (align_corners=False), then clip to [0, size - 1]
    const float lo = -0.5f, span = static_cast<float>(size);
    x = fabsf(x - lo);
    const float extra = fmodf(x, span);
    const int flips = static_cast<int>(floorf(x / span));
    x = (flips & 1) ? (span - extra + lo) : (extra + lo);
    return fminf(fmaxf(x, 0.f), static_cast<float>(size - 1));
}

__device__ __forceinline__ float src_coord(int o, int osize, float lo, float hi, int isize) {
    // normalised output coordinate -> theta -> input pixel coordinate (align_corners=False)
    const float xn = (2.f * o + 1.f) / osize - 1.f;
    const float xs = (hi - lo) / isize * xn + ((lo + hi) / isize - 1.f);
    return ((xs + 1.f) * isize - 1.f) * 0.5f;
}

// boxes (B, 4) = [t, b, l, r] (device);  x (B, C, H, W) -> y (B, C, OH, OW)
__global__ void __launch_bounds__(256)
crop_bilinear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ boxes, float* __restrict__ y, int B, int C,
                         int H, int W, int OH, int OW) {
    const unsigned total = static_cast<unsigned>(B) * C * OH * OW;
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < total; i += gridDim.x * 256u) {
        const unsigned ox = i % static_cast<unsigned>(OW);
        const unsigned t1 = i / static_cast<unsigned>(OW);
        const unsigned oy = t1 % static_cast<unsigned>(OH);
        const unsigned bc = t1 / static_cast<unsigned>(OH);
        const unsigned b = bc / static_cast<unsigned>(C);
        const float* bx = boxes + b * 4;
        const float fy = reflect_coord(src_coord(oy, OH, bx[0], bx[1], H), H);
        const float fx = reflect_coord(src_coord(ox, OW, bx[2], bx[3], W), W);
        const int y0 = static_cast<int>(floorf(fy)), x0 = static_cast<int>(floorf(fx));
        const float wy1 = fy - y0, wx1 = fx - x0, wy0 = 1.f - wy1, wx0 = 1.f - wx1;
        const float* img = x + static_cast<size_t>(bc) * H * W;
        const bool y1in = y0 + 1 < H, x1in = x0 + 1 < W;
        float v = wy0 * wx0 * __ldg(img + y0 * W + x0);
        if (x1in) v += wy0 * wx1 * __ldg(img + y0 * W + x0 + 1);
        if (y1in) v += wy1 * wx0 * __ldg(img + (y0 + 1) * W + x0);
        if (y1in && x1in) v += wy1 * wx1 * __ldg(img + (y0 + 1) * W + x0 + 1);
        y[i] = v;
    }
}

// dx (B, C, H, W) = adjoint of the crop.  Gather form: the sampling positions are monotonic in the output index (boxes
// inside the image: no reflection is active — checked by the host wrapper), so the outputs that touch input row iy are
// those with source coordinate in (iy - 1, iy + 1): a contiguous index range found from the inverse affine map.
__global__ void __launch_bounds__(256)
crop_bilinear_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ boxes, float* __restrict__ dx, int B, int C,
                         int H, int W, int OH, int OW) {
    const unsigned total = static_cast<unsigned>(B) * C * H * W;
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < total; i += gridDim.x * 256u) {
        const unsigned ix = i % static_cast<unsigned>(W);
        const unsigned t1 = i / static_cast<unsigned>(W);
        const unsigned iy = t1 % static_cast<unsigned>(H);
        const unsigned bc = t1 / static_cast<unsigned>(H);
        const unsigned b = bc / static_cast<unsigned>(C);
        const float* bx = boxes + b * 4;
        // source coordinate = a * o + c  (a > 0)
        const float ay = (bx[1] - bx[0]) / OH, cy = bx[0] + 0.5f * ay - 0.5f;
        const float ax = (bx[3] - bx[2]) / OW, cx = bx[2] + 0.5f * ax - 0.5f;
        int oy0 = static_cast<int>(ceilf((static_cast<float>(iy) - 1.f - cy) / ay)) - 1;
        int oy1 = static_cast<int>(floorf((static_cast<float>(iy) + 1.f - cy) / ay)) + 1;
        int ox0 = static_cast<int>(ceilf((static_cast<float>(ix) - 1.f - cx) / ax)) - 1;
        int ox1 = static_cast<int>(floorf((static_cast<float>(ix) + 1.f - cx) / ax)) + 1;
        oy0 = max(oy0, 0); ox0 = max(ox0, 0); oy1 = min(oy1, OH - 1); ox1 = min(ox1, OW - 1);
        const float* g = dy + static_cast<size_t>(bc) * OH * OW;
        float acc = 0.f;
        for (int oy = oy0; oy <= oy1; ++oy) {
            const float fy = src_coord(oy, OH, bx[0], bx[1], H);
            const float wy = 1.f - fabsf(fy - static_cast<float>(iy));
            if (wy <= 0.f) continue;
            for (int ox = ox0; ox <= ox1; ++ox) {
                const float fx = src_coord(ox, OW, bx[2], bx[3], W);
                const float wx = 1.f - fabsf(fx - static_cast<float>(ix));
                if (wx > 0.f) acc += wy * wx * __ldg(g + oy * OW + ox);
            }
        }
        dx[i] = acc;
    }
}

// ------------------------------------------------------------------------------------------------ discriminator head
// feat (B, P, C) NHWC raw;  o[b][c] = sum_p relu(feat[b,p,c]);  score[b] = s * <o[b], w> + bias + <o[b], embed[b]>
// grid = B, block 256 (C <= 1024: four channels per thread at most)
__global__ void __launch_bounds__(256)
disc_head_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ embed, const float* __restrict__ w,
                     const float* __restrict__ inv_sigma, const float* __restrict__ bias, float* __restrict__ o,
                     float* __restrict__ score, int P, int C) {
    const int b = blockIdx.x;
    float lin = 0.f, proj = 0.f, z = 0.f;
    for (int c = threadIdx.x; c < C; c += 256) {
        float s = 0.f;
        for (int p = 0; p < P; ++p) s += fmaxf(__ldg(feat + (static_cast<size_t>(b) * P + p) * C + c), 0.f);
        o[static_cast<size_t>(b) * C + c] = s;
        lin += s * __ldg(w + c);
        if (embed) proj += s * __ldg(embed + static_cast<size_t>(b) * C + c);
    }
    block_sum3(lin, proj, z);
    if (threadIdx.x == 0) score[b] = lin * __ldg(inv_sigma) + __ldg(bias) + proj;
}

// d_feat[b,p,c] = [feat > 0] * g[b] * (s*w[c] + embed[b,c]);  d_embed[b,c] = g[b]*o[b,c]
__global__ void __launch_bounds__(256)
disc_head_bwd_kernel(const float* __restrict__ feat, const float* __restrict__ embed, const float* __restrict__ w,
                     const float* __restrict__ inv_sigma, const float* __restrict__ o, const float* __restrict__ g,
                     float* __restrict__ d_feat, float* __restrict__ d_embed, int P, int C) {
    const int b = blockIdx.x;
    const float gb = __ldg(g + b), s = __ldg(inv_sigma);
    for (int c = threadIdx.x; c < C; c += 256) {
        const float k = gb * (s * __ldg(w + c) + (embed ? __ldg(embed + static_cast<size_t>(b) * C + c) : 0.f));
        for (int p = 0; p < P; ++p) {
            const size_t off = (static_cast<size_t>(b) * P + p) * C + c;
            d_feat[off] = __ldg(feat + off) > 0.f ? k : 0.f;
        }
        if (d_embed) d_embed[static_cast<size_t>(b) * C + c] = gb * __ldg(o + static_cast<size_t>(b) * C + c);
    }
}

// parameter gradients of the head's linear layer (one block): dw[c] (+)= s * sum_b g[b] o[b,c];  ds = sum_b g[b] <o[b], w>;
// dbias (+)= sum_b g[b]
__global__ void __launch_bounds__(256)
disc_head_wgrad_kernel(const float* __restrict__ o, const float* __restrict__ w, const float* __restrict__ inv_sigma,
                       const float* __restrict__ g, float* __restrict__ dw, float* __restrict__ ds,
                       float* __restrict__ dbias, int accumulate, int B, int C) {
    const float s = __ldg(inv_sigma);
    float dsum = 0.f, gsum = 0.f, z = 0.f;
    for (int c = threadIdx.x; c < C; c += 256) {
        float a = 0.f;
        for (int b = 0; b < B; ++b) a += __ldg(g + b) * __ldg(o + static_cast<size_t>(b) * C + c);
        if (dw) dw[c] = (accumulate ? dw[c] : 0.f) + s * a;
        dsum += a * __ldg(w + c);
    }
    if (threadIdx.x == 0)
        for (int b = 0; b < B; ++b) gsum += __ldg(g + b);
    block_sum3(dsum, gsum, z);
    if (threadIdx.x == 0) {
        if (ds) ds[0] = dsum;
        if (dbias) dbias[0] = (accumulate ? dbias[0] : 0.f) + gsum;
    }
}

static int loss_blocks(long total) {
    long blocks = (total + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    return static_cast<int>(blocks < 1 ? 1 : blocks);
}

}  // namespace b200lp

using namespace b200lp;

extern "C" int64_t b200lp_dice_workspace(int32_t B, int32_t HW) {
    if (B <= 0 || HW <= 0) return B200LP_EINVAL;
    return static_cast<int64_t>(loss_blocks(static_cast<long>(B) * HW)) * 3 * 4;
}

extern "C" int32_t b200lp_dice_fwd(const float* fake, const float* real, float weight, float* sums, float* loss,
                                   float* workspace, int64_t workspace_bytes, int32_t B, int32_t CR, int32_t HW,
                                   void* stream) {
    B200LP_REQUIRE(fake && real && sums && loss && workspace && B > 0 && CR > 0 && HW > 0, "dice_fwd: bad args");
    B200LP_REQUIRE(static_cast<long>(B) * CR * HW < (1L << 31), "dice_fwd: tensor too large");
    const int blocks = loss_blocks(static_cast<long>(B) * HW);
    B200LP_REQUIRE(workspace_bytes >= static_cast<int64_t>(blocks) * 12, "dice_fwd: workspace too small");
    cudaStream_t st = as_stream(stream);
    dice_partial_kernel<<<blocks, 256, 0, st>>>(fake, real, workspace, B, CR, HW);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    dice_final_kernel<<<1, 32, 0, st>>>(workspace, blocks, weight, sums, loss);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_dice_bwd(const float* fake, const float* real, const float* sums, const float* grad, float weight,
                                   float* d_fake, int32_t B, int32_t CR, int32_t HW, void* stream) {
    B200LP_REQUIRE(fake && real && sums && grad && d_fake && B > 0 && CR > 0 && HW > 0, "dice_bwd: bad args");
    dice_bwd_kernel<<<loss_blocks(static_cast<long>(B) * HW), 256, 0, as_stream(stream)>>>(fake, real, sums, grad, weight,
                                                                                           d_fake, B, CR, HW);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_adversarial_fwd(const float* fake_g, const float* fake_d, const float* real, float* out2,
                                          int32_t B, int32_t relativistic, void* stream) {
    B200LP_REQUIRE(fake_g && fake_d && real && out2 && B > 0 && relativistic >= 0 && relativistic <= 2, "adversarial_fwd: bad args");
    adversarial_fwd_kernel<<<1, 256, 0, as_stream(stream)>>>(fake_g, fake_d, real, out2, B, relativistic);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_adversarial_bwd(const float* fake_d, const float* real, const float* grad_g, const float* grad_d,
                                          float* d_fake_g, float* d_fake_d, float* d_real, int32_t B, void* stream) {
    B200LP_REQUIRE(fake_d && real && B > 0, "adversarial_bwd: bad args");
    adversarial_bwd_kernel<<<1, 256, 0, as_stream(stream)>>>(fake_d, real, grad_g, grad_d, d_fake_g, d_fake_d, d_real, B);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_crop_bilinear_fwd(const float* x, const float* boxes, float* y, int32_t B, int32_t C, int32_t H,
                                            int32_t W, int32_t OH, int32_t OW, void* stream) {
    B200LP_REQUIRE(x && boxes && y && B > 0 && C > 0 && H > 1 && W > 1 && OH > 0 && OW > 0, "crop_bilinear_fwd: bad args");
    const long total = static_cast<long>(B) * C * OH * OW;
    B200LP_REQUIRE(total < (1L << 31) && static_cast<long>(B) * C * H * W < (1L << 31), "crop_bilinear_fwd: tensor too large");
    crop_bilinear_fwd_kernel<<<loss_blocks(total), 256, 0, as_stream(stream)>>>(x, boxes, y, B, C, H, W, OH, OW);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_crop_bilinear_bwd(const float* dy, const float* boxes, float* dx, int32_t B, int32_t C, int32_t H,
                                            int32_t W, int32_t OH, int32_t OW, void* stream) {
    B200LP_REQUIRE(dy && boxes && dx && B > 0 && C > 0 && H > 1 && W > 1 && OH > 0 && OW > 0, "crop_bilinear_bwd: bad args");
    const long total = static_cast<long>(B) * C * H * W;
    B200LP_REQUIRE(total < (1L << 31) && static_cast<long>(B) * C * OH * OW < (1L << 31), "crop_bilinear_bwd: tensor too large");
    crop_bilinear_bwd_kernel<<<loss_blocks(total), 256, 0, as_stream(stream)>>>(dy, boxes, dx, B, C, H, W, OH, OW);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_disc_head_fwd(const float* feat, const float* embed, const float* w, const float* inv_sigma,
                                        const float* bias, float* o, float* score, int32_t B, int32_t P, int32_t C,
                                        void* stream) {
    B200LP_REQUIRE(feat && w && inv_sigma && bias && o && score && B > 0 && P > 0 && C > 0, "disc_head_fwd: bad args");
    disc_head_fwd_kernel<<<B, 256, 0, as_stream(stream)>>>(feat, embed, w, inv_sigma, bias, o, score, P, C);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_disc_head_bwd(const float* feat, const float* embed, const float* w, const float* inv_sigma,
                                        const float* o, const float* grad, float* d_feat, float* d_embed, float* dw,
                                        float* ds, float* dbias, int32_t accumulate, int32_t B, int32_t P, int32_t C,
                                        void* stream) {
    B200LP_REQUIRE(feat && w && inv_sigma && o && grad && B > 0 && P > 0 && C > 0, "disc_head_bwd: bad args");
    cudaStream_t st = as_stream(stream);
    if (d_feat) {
        disc_head_bwd_kernel<<<B, 256, 0, st>>>(feat, embed, w, inv_sigma, o, grad, d_feat, d_embed, P, C);
        B200LP_CHECK_CUDA(cudaGetLastError());
        count_launch();
    }
    if (dw || ds || dbias) {
        disc_head_wgrad_kernel<<<1, 256, 0, st>>>(o, w, inv_sigma, grad, dw, ds, dbias, accumulate, B, C);
        B200LP_CHECK_CUDA(cudaGetLastError());
        count_launch();
    }
    return B200LP_OK;
}
