// Spectral normalisation as three multi-tensor launches per network pass (instead of ~14 tiny launches per weight):
//
//   torch.nn.utils.spectral_norm (SpectralNorm.compute_weight), one power iteration in training mode:
//       v <- normalize(W^T u) ;  u <- normalize(W v) ;  sigma = u^T W v ;  W_bar = W / sigma
//   call sites: generators/common/blocks.py:78-100, generators/vector_pose_..._noBottleneck.py:84-86,
//               discriminators/no_landmarks.py:54-66 — executed on EVERY forward call (3x per step for D).
//
//   kernel A  (sn_wt_u)   : partial t = W^T u per slab of 128 rows (blocks over column chunks x row slabs)
//   kernel A2 (sn_t_sum)  : t = sum of the slab partials (fixed order), |t|^2 partials per column chunk
//   kernel B  (sn_w_v)    : s = W v_hat, v_hat = t/|t|      (blocks over row chunks of every weight)
//   kernel C  (sn_final)  : u_new = s/|s|, sigma = |s| (= u_new^T W v_hat), writes u, v, 1/sigma, and a snapshot
//                           copy of (u, v) for the backward pass (the buffers are overwritten by the next pass)
//   eval mode: kernel B with v_hat = v (no normalisation) and kernel C computing sigma = u^T s without updates.
//
// All weights of a network pass are described by one by-value parameter block (<= 32 tensors), norms are reduced
// through per-block partial sums (deterministic; no atomics).
//
// Also here: the spectral-norm correction of a weight gradient
//       dL/dW_orig = s*G - s^2 <G, W> u v^T ,   s = 1/sigma,  G = dL/d(W*s)
#include "common.cuh"

namespace b200lp {

constexpr int kSnMaxTensors = 32;   // 32 x 112-byte items = 3.5 KB of kernel parameters (limit 4 KB)
constexpr int kSnThreads = 256;
constexpr int kSnColsPerBlock = 32;   // kernel A: 32 columns x 8 row groups
constexpr int kSnRowsPerBlock = 8;    // kernel B: one warp per row
constexpr int kSnSlabRows = 128;      // kernel A: rows per block (tall matrices - the 13056 x 768 projector - get many slabs)

struct SnItem {
    const float* w;     // [rows][cols]
    float* u;           // [rows]  module buffer (in/out)
    float* v;           // [cols]  module buffer (in/out)
    float* snap_u;      // [rows]  copy of the vectors used for this pass's sigma (for backward), may be NULL
    float* snap_v;      // [cols]
    float* t;           // [cols]  scratch: W^T u
    float* tpart;       // [slabs][cols] scratch: per-slab partial W^T u
    float* s;           // [rows]  scratch: W v
    float* part;        // scratch: partial sums, >= max(blocksA, blocksB) floats
    float* inv_sigma;   // [1]
    int rows, cols;
    int blkA0, blkB0;   // first block index of this tensor in kernels A / B
    int blkA20;         // ... in kernel A2
    int slabs;          // ceil(rows / kSnSlabRows)
    float eps;
    int pad;
};

struct SnBatch {
    SnItem it[kSnMaxTensors];
    int count;
    int training;
};
static_assert(sizeof(SnBatch) <= 4096, "SnBatch travels as a by-value kernel parameter");

__device__ __forceinline__ int find_tensor(const SnBatch& b, int blk, int kernel) {
    int t = 0;
    for (int i = 1; i < b.count; ++i) {
        const int first = kernel == 0 ? b.it[i].blkA0 : (kernel == 1 ? b.it[i].blkB0 : b.it[i].blkA20);
        if (blk >= first) t = i;
    }
    return t;
}

// tpart[slab][j] = sum over the slab's rows i of W[i][j] u[i]
__global__ void __launch_bounds__(kSnThreads)
sn_wt_u_kernel(const __grid_constant__ SnBatch b) {
    const int ti = find_tensor(b, blockIdx.x, 0);
    const SnItem& it = b.it[ti];
    const int local = blockIdx.x - it.blkA0;
    const int nA = (it.cols + kSnColsPerBlock - 1) / kSnColsPerBlock;
    const int slab = local / nA, cb = local - slab * nA;
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;     // 32 columns x 8 row groups
    const int col = cb * kSnColsPerBlock + cx;
    const int r0 = slab * kSnSlabRows;
    const int r1 = r0 + kSnSlabRows < it.rows ? r0 + kSnSlabRows : it.rows;
    float acc = 0.f;
    if (col < it.cols) {
        const float* wc = it.w + col;
#pragma unroll 4
        for (int i = r0 + ry; i < r1; i += 8) acc += __ldg(wc + static_cast<size_t>(i) * it.cols) * __ldg(it.u + i);
    }
    __shared__ float red[8][33];
    red[ry][cx] = acc;
    __syncthreads();
    if (ry == 0 && col < it.cols) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][cx];
        it.tpart[static_cast<size_t>(slab) * it.cols + col] = t;
    }
}

// t[j] = sum_slab tpart[slab][j] (fixed order);  part[block] = sum_j t[j]^2 over this block's 256 columns
__global__ void __launch_bounds__(kSnThreads)
sn_t_sum_kernel(const __grid_constant__ SnBatch b) {
    const int ti = find_tensor(b, blockIdx.x, 2);
    const SnItem& it = b.it[ti];
    const int cb = blockIdx.x - it.blkA20;
    const int col = cb * kSnThreads + threadIdx.x;
    float t = 0.f;
    if (col < it.cols) {
        for (int k = 0; k < it.slabs; ++k) t += it.tpart[static_cast<size_t>(k) * it.cols + col];
        it.t[col] = t;
    }
    float sq = t * t;
    __shared__ float sh[kSnThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
#pragma unroll
        for (int k = 0; k < kSnThreads / 32; ++k) tot += sh[k];
        it.part[cb] = tot;
    }
}

// s[i] = sum_j W[i][j] vhat[j];  training: vhat = t / max(|t|, eps) (|t| from kernel A's partials), eval: vhat = v
__global__ void __launch_bounds__(kSnThreads)
sn_w_v_kernel(const __grid_constant__ SnBatch b) {
    const int ti = find_tensor(b, blockIdx.x, 1);
    const SnItem& it = b.it[ti];
    const int rb = blockIdx.x - it.blkB0;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    float inv = 1.f;
    const float* vec = it.v;
    if (b.training) {
        const int nA = (it.cols + kSnThreads - 1) / kSnThreads;      // |t|^2 partials of sn_t_sum_kernel
        float sq = 0.f;
        for (int k = 0; k < nA; ++k) sq += it.part[k];          // same order in every block: deterministic
        inv = 1.f / fmaxf(sqrtf(sq), it.eps);
        vec = it.t;
    }
    const int row = rb * kSnRowsPerBlock + wrp;
    float acc = 0.f;
    if (row < it.rows) {
        const float* wr = it.w + static_cast<size_t>(row) * it.cols;
        for (int j = lane; j < it.cols; j += 32) acc += __ldg(wr + j) * (vec[j] * inv);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0 && row < it.rows) it.s[row] = acc;
}

// one block per tensor
__global__ void __launch_bounds__(kSnThreads)
sn_final_kernel(const __grid_constant__ SnBatch b) {
    const SnItem& it = b.it[blockIdx.x];
    __shared__ float sh[kSnThreads / 32];
    __shared__ float bc[2];
    // |t| again (for v_hat) and |s|^2 or u.s
    float a = 0.f;
    if (b.training) {
        const int nA = (it.cols + kSnThreads - 1) / kSnThreads;
        if (threadIdx.x == 0) {
            float sq = 0.f;
            for (int k = 0; k < nA; ++k) sq += it.part[k];
            bc[0] = 1.f / fmaxf(sqrtf(sq), it.eps);
        }
        for (int i = threadIdx.x; i < it.rows; i += blockDim.x) a += it.s[i] * it.s[i];
    } else {
        for (int i = threadIdx.x; i < it.rows; i += blockDim.x) a += it.u[i] * it.s[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int k = 0; k < kSnThreads / 32; ++k) tot += sh[k];
        bc[1] = tot;
    }
    __syncthreads();
    if (b.training) {
        const float inv_t = bc[0];
        const float ns = sqrtf(bc[1]);
        const float inv_s = 1.f / fmaxf(ns, it.eps);
        const float sigma = bc[1] * inv_s;                       // u_new . s = |s|^2 / max(|s|, eps)
        for (int j = threadIdx.x; j < it.cols; j += blockDim.x) {
            const float vh = it.t[j] * inv_t;
            it.v[j] = vh;
            if (it.snap_v) it.snap_v[j] = vh;
        }
        for (int i = threadIdx.x; i < it.rows; i += blockDim.x) {
            const float un = it.s[i] * inv_s;
            it.u[i] = un;
            if (it.snap_u) it.snap_u[i] = un;
        }
        if (threadIdx.x == 0) *it.inv_sigma = 1.f / sigma;
    } else {
        if (it.snap_v) for (int j = threadIdx.x; j < it.cols; j += blockDim.x) it.snap_v[j] = it.v[j];
        if (it.snap_u) for (int i = threadIdx.x; i < it.rows; i += blockDim.x) it.snap_u[i] = it.u[i];
        if (threadIdx.x == 0) *it.inv_sigma = 1.f / bc[1];
    }
}

// ---- weight-gradient correction -------------------------------------------------------------------------------
// pass 1: part[block] = sum over this block's elements of G*W           (G = raw tensor-core weight gradient)
__global__ void __launch_bounds__(256)
sn_dot_kernel(const float* __restrict__ g, const float* __restrict__ w, float* __restrict__ part, long n) {
    float acc = 0.f;
    for (long i = blockIdx.x * 256L + threadIdx.x; i < n; i += gridDim.x * 256L) acc += __ldg(g + i) * __ldg(w + i);
    __shared__ float sh[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += sh[k];
        part[blockIdx.x] = t;
    }
}
// pass 2: dw[i][j] = s*G[i][j] - s^2 * c * u[i] * v[j],  c = sum(part)
__global__ void __launch_bounds__(256)
sn_fix_kernel(const float* __restrict__ g, const float* __restrict__ part, int nparts, const float* __restrict__ inv_sigma,
              const float* __restrict__ u, const float* __restrict__ v, float* __restrict__ dw, int cols, long n) {
    __shared__ float cs;
    if (threadIdx.x == 0) {
        float c = 0.f;
        for (int k = 0; k < nparts; ++k) c += part[k];
        const float s = __ldg(inv_sigma);
        cs = c * s * s;
    }
    __syncthreads();
    const float s = __ldg(inv_sigma), k = cs;
    for (long i = blockIdx.x * 256L + threadIdx.x; i < n; i += gridDim.x * 256L) {
        const long r = i / cols;
        const int c = static_cast<int>(i - r * cols);
        dw[i] = s * __ldg(g + i) - k * __ldg(u + r) * __ldg(v + c);
    }
}

}  // namespace b200lp

using namespace b200lp;

extern "C" int32_t b200lp_sn_max_tensors(void) { return kSnMaxTensors; }

extern "C" int64_t b200lp_sn_scratch_floats(int32_t rows, int32_t cols) {
    const int nA2 = (cols + kSnThreads - 1) / kSnThreads;
    const int slabs = (rows + kSnSlabRows - 1) / kSnSlabRows;
    return static_cast<int64_t>(cols) + rows + nA2 + static_cast<int64_t>(slabs) * cols;   // t, s, |t|^2 partials, slab partials
}

extern "C" int32_t b200lp_sn_sigma_multi(const b200lp_sn_item* items, int32_t count, int32_t training, void* stream) {
    B200LP_REQUIRE(items && count > 0 && count <= kSnMaxTensors, "sn_sigma_multi: count %d not in [1,%d]", count,
                   kSnMaxTensors);
    SnBatch b;
    b.count = count;
    b.training = training;
    int blkA = 0, blkB = 0, blkA2 = 0;
    for (int i = 0; i < count; ++i) {
        const b200lp_sn_item& s = items[i];
        B200LP_REQUIRE(s.w && s.u && s.v && s.scratch && s.inv_sigma && s.rows > 0 && s.cols > 0,
                       "sn_sigma_multi: bad item %d", i);
        SnItem& d = b.it[i];
        d.w = s.w; d.u = s.u; d.v = s.v; d.snap_u = s.snap_u; d.snap_v = s.snap_v;
        const int nA2 = (s.cols + kSnThreads - 1) / kSnThreads;
        d.t = s.scratch; d.s = s.scratch + s.cols; d.part = s.scratch + s.cols + s.rows;
        d.tpart = d.part + nA2;
        d.inv_sigma = s.inv_sigma;
        d.rows = s.rows; d.cols = s.cols; d.eps = s.eps; d.pad = 0;
        d.slabs = (s.rows + kSnSlabRows - 1) / kSnSlabRows;
        d.blkA0 = blkA; d.blkB0 = blkB; d.blkA20 = blkA2;
        blkA += ((s.cols + kSnColsPerBlock - 1) / kSnColsPerBlock) * d.slabs;
        blkB += (s.rows + kSnRowsPerBlock - 1) / kSnRowsPerBlock;
        blkA2 += nA2;
    }
    cudaStream_t st = as_stream(stream);
    if (training) {
        sn_wt_u_kernel<<<blkA, kSnThreads, 0, st>>>(b);
        B200LP_CHECK_CUDA(cudaGetLastError());
        count_launch();
        sn_t_sum_kernel<<<blkA2, kSnThreads, 0, st>>>(b);
        B200LP_CHECK_CUDA(cudaGetLastError());
        count_launch();
    }
    sn_w_v_kernel<<<blkB, kSnThreads, 0, st>>>(b);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    sn_final_kernel<<<count, kSnThreads, 0, st>>>(b);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int64_t b200lp_sn_wgrad_fix_workspace(int64_t n) {
    long blocks = (n + 256L * 8 - 1) / (256L * 8);
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    return blocks * 4;
}

extern "C" int32_t b200lp_sn_wgrad_fix(const float* g, const float* w, const float* inv_sigma, const float* u,
                                       const float* v, float* dw, float* workspace, int32_t rows, int32_t cols,
                                       void* stream) {
    B200LP_REQUIRE(g && w && inv_sigma && u && v && dw && workspace && rows > 0 && cols > 0, "sn_wgrad_fix: bad args");
    const long n = static_cast<long>(rows) * cols;
    const int blocks = static_cast<int>(b200lp_sn_wgrad_fix_workspace(n) / 4);
    cudaStream_t st = as_stream(stream);
    sn_dot_kernel<<<blocks, 256, 0, st>>>(g, w, workspace, n);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    sn_fix_kernel<<<blocks, 256, 0, st>>>(g, workspace, blocks, inv_sigma, u, v, dw, cols, n);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}
