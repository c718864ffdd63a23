// Fused multi-tensor optimizer step + weight EMA (HBM-bound, one pass over parameters, gradients and moments).
//
// Replaces, per training step (runners/holycow.py:99-109,244,252,257):
//   torch.optim.Adam.step / the vendored RAdam.step (utils/radam.py:29-95)  — ~10 foreach launches per optimizer
//   TrainingModule.update_running_average: p_avg = p_avg*alpha + p*(1-alpha)  — 2 more passes over E/G parameters
// with two launches: `opt_tick` (advances the step counter ON THE DEVICE and derives the step-dependent scalars, so the
// whole training step can be captured in a CUDA graph and replayed) and `adam_ema_multi` (float4 streaming update).
//
// The tensors of one optimizer live in flat buffers (parameters excepted): grads are views into the data-parallel
// gradient bucket, exp_avg / exp_avg_sq are views into two flat state buffers, so a "tensor table" entry is five
// pointers + a length, and work is split into fixed-size chunks (chunk -> tensor, offset) like apex's multi_tensor_apply.
#include "common.cuh"

namespace b200lp {

struct OptTensor {
    float* p;
    const float* g;
    float* m;
    float* v;
    float* ema;      // NULL: no running average for this tensor
    long long n;
};

// state[0] = step (as float), [1] = step_size (lr folded in), [2] = rectified flag / 1, [3] = 1/sqrt(bias_correction2),
// [4] = learning rate, [5] = running-average alpha.  lr and alpha are read from the DEVICE vector, so a captured CUDA
// graph of the step follows a learning-rate schedule: the host only rewrites state[4..5] between replays.
__global__ void opt_tick_kernel(float* __restrict__ state, float beta1, float beta2, int mode, int degenerated_to_sgd) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double lr = static_cast<double>(state[4]);
    const double t = static_cast<double>(state[0]) + 1.0;
    state[0] = static_cast<float>(t);
    const double b1t = pow(static_cast<double>(beta1), t), b2t = pow(static_cast<double>(beta2), t);
    if (mode == 0) {                      // torch.optim.Adam (no amsgrad, no weight decay)
        state[1] = static_cast<float>(lr / (1.0 - b1t));
        state[2] = 1.f;
        state[3] = static_cast<float>(1.0 / sqrt(1.0 - b2t));
    } else {                              // RAdam, utils/radam.py:29-95
        const double n_max = 2.0 / (1.0 - beta2) - 1.0;
        const double n_sma = n_max - 2.0 * t * b2t / (1.0 - b2t);
        if (n_sma >= 5.0) {
            const double r = sqrt((1.0 - b2t) * (n_sma - 4.0) / (n_max - 4.0) * (n_sma - 2.0) / n_sma * n_max / (n_max - 2.0));
            state[1] = static_cast<float>(lr * r / (1.0 - b1t));
            state[2] = 1.f;
        } else {
            state[1] = degenerated_to_sgd ? static_cast<float>(lr / (1.0 - b1t)) : 0.f;
            state[2] = 0.f;
        }
        state[3] = 1.f;
    }
}

__global__ void __launch_bounds__(256)
adam_ema_multi_kernel(const OptTensor* __restrict__ table, const int* __restrict__ chunk_tensor,
                      const long long* __restrict__ chunk_off, long long chunk_elems, const float* __restrict__ state,
                      float beta1, float beta2, float eps, int mode) {
    const OptTensor t = table[chunk_tensor[blockIdx.x]];
    const float ema_alpha = state[5];
    const long long off = chunk_off[blockIdx.x];
    long long end = off + chunk_elems;
    if (end > t.n) end = t.n;
    const float step_size = state[1];
    const bool rect = state[2] != 0.f;
    const float inv_bc2_sqrt = state[3];
    const float one_m_b1 = 1.f - beta1, one_m_b2 = 1.f - beta2, one_m_a = 1.f - ema_alpha;
    // 16-byte path when every stream of this chunk is aligned (gradients / moments are views into flat buffers, so
    // alignment depends on the sizes of the tensors before this one); the scalar loop takes the rest
    long long vec_end = off;
    const uintptr_t al = reinterpret_cast<uintptr_t>(t.p + off) | reinterpret_cast<uintptr_t>(t.g + off) |
                         reinterpret_cast<uintptr_t>(t.m + off) | reinterpret_cast<uintptr_t>(t.v + off) |
                         (t.ema ? reinterpret_cast<uintptr_t>(t.ema + off) : 0);
    if ((al & 15) == 0) {
        const long long n4 = (end - off) >> 2;
        vec_end = off + (n4 << 2);
        const float4* g4 = reinterpret_cast<const float4*>(t.g + off);
        float4* m4 = reinterpret_cast<float4*>(t.m + off);
        float4* v4 = reinterpret_cast<float4*>(t.v + off);
        float4* p4 = reinterpret_cast<float4*>(t.p + off);
        float4* e4 = t.ema ? reinterpret_cast<float4*>(t.ema + off) : nullptr;
        for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
            const float4 gq = g4[i];
            float4 mq = m4[i], vq = v4[i], pq = p4[i];
            const float gg[4] = {gq.x, gq.y, gq.z, gq.w};
            float mm[4] = {mq.x, mq.y, mq.z, mq.w}, vv[4] = {vq.x, vq.y, vq.z, vq.w}, pp[4] = {pq.x, pq.y, pq.z, pq.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                vv[k] = vv[k] * beta2 + one_m_b2 * gg[k] * gg[k];
                if (mode == 0) {
                    mm[k] = mm[k] + (gg[k] - mm[k]) * one_m_b1;
                    pp[k] -= step_size * mm[k] / (sqrtf(vv[k]) * inv_bc2_sqrt + eps);
                } else {
                    mm[k] = mm[k] * beta1 + one_m_b1 * gg[k];
                    pp[k] -= rect ? step_size * mm[k] / (sqrtf(vv[k]) + eps) : step_size * mm[k];
                }
            }
            m4[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
            v4[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
            p4[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
            if (e4) {
                float4 eq = e4[i];
                eq.x = eq.x * ema_alpha + pp[0] * one_m_a; eq.y = eq.y * ema_alpha + pp[1] * one_m_a;
                eq.z = eq.z * ema_alpha + pp[2] * one_m_a; eq.w = eq.w * ema_alpha + pp[3] * one_m_a;
                e4[i] = eq;
            }
        }
    }
    for (long long i = vec_end + threadIdx.x; i < end; i += blockDim.x) {
        const float g = t.g[i];
        float m = t.m[i], v = t.v[i], p = t.p[i];
        v = v * beta2 + one_m_b2 * g * g;
        if (mode == 0) {
            m = m + (g - m) * one_m_b1;                                  // exp_avg.lerp_(grad, 1 - beta1)
            p -= step_size * m / (sqrtf(v) * inv_bc2_sqrt + eps);
        } else {
            m = m * beta1 + one_m_b1 * g;
            p -= rect ? step_size * m / (sqrtf(v) + eps) : step_size * m;
        }
        t.m[i] = m; t.v[i] = v; t.p[i] = p;
        if (t.ema) t.ema[i] = t.ema[i] * ema_alpha + p * one_m_a;
    }
}

// plain EMA over a tensor table (used when the optimizer is not the fused one): ema = ema*alpha + p*(1-alpha)
__global__ void __launch_bounds__(256)
ema_multi_kernel(const OptTensor* __restrict__ table, const int* __restrict__ chunk_tensor,
                 const long long* __restrict__ chunk_off, long long chunk_elems, float ema_alpha) {
    const OptTensor t = table[chunk_tensor[blockIdx.x]];
    const long long off = chunk_off[blockIdx.x];
    long long end = off + chunk_elems;
    if (end > t.n) end = t.n;
    if (!t.ema) return;
    const float oma = 1.f - ema_alpha;
    long long vec_end = off;
    if (((reinterpret_cast<uintptr_t>(t.ema + off) | reinterpret_cast<uintptr_t>(t.p + off)) & 15) == 0) {
        const long long n4 = (end - off) >> 2;
        vec_end = off + (n4 << 2);
        float4* e4 = reinterpret_cast<float4*>(t.ema + off);
        const float4* p4 = reinterpret_cast<const float4*>(t.p + off);
        for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
            float4 e = e4[i];
            const float4 q = p4[i];
            e.x = e.x * ema_alpha + q.x * oma; e.y = e.y * ema_alpha + q.y * oma;
            e.z = e.z * ema_alpha + q.z * oma; e.w = e.w * ema_alpha + q.w * oma;
            e4[i] = e;
        }
    }
    for (long long i = vec_end + threadIdx.x; i < end; i += blockDim.x) t.ema[i] = t.ema[i] * ema_alpha + t.p[i] * oma;
}

}  // namespace b200lp

using namespace b200lp;

extern "C" int32_t b200lp_adam_ema_multi(const void* table_dev, const int32_t* chunk_tensor_dev,
                                         const int64_t* chunk_off_dev, int32_t n_chunks, int64_t chunk_elems,
                                         float* state_dev, float beta1, float beta2, float eps, int32_t mode,
                                         int32_t degenerated_to_sgd, void* stream) {
    B200LP_REQUIRE(table_dev && chunk_tensor_dev && chunk_off_dev && state_dev && n_chunks > 0 && chunk_elems > 0 &&
                       (mode == 0 || mode == 1),
                   "adam_ema_multi: bad args");
    cudaStream_t st = as_stream(stream);
    opt_tick_kernel<<<1, 32, 0, st>>>(state_dev, beta1, beta2, mode, degenerated_to_sgd);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    adam_ema_multi_kernel<<<n_chunks, 256, 0, st>>>(static_cast<const OptTensor*>(table_dev), chunk_tensor_dev,
                                                    reinterpret_cast<const long long*>(chunk_off_dev), chunk_elems,
                                                    state_dev, beta1, beta2, eps, mode);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}

extern "C" int32_t b200lp_ema_multi(const void* table_dev, const int32_t* chunk_tensor_dev, const int64_t* chunk_off_dev,
                                    int32_t n_chunks, int64_t chunk_elems, float ema_alpha, void* stream) {
    B200LP_REQUIRE(table_dev && chunk_tensor_dev && chunk_off_dev && n_chunks > 0 && chunk_elems > 0, "ema_multi: bad args");
    ema_multi_kernel<<<n_chunks, 256, 0, as_stream(stream)>>>(static_cast<const OptTensor*>(table_dev), chunk_tensor_dev,
                                                             reinterpret_cast<const long long*>(chunk_off_dev),
                                                             chunk_elems, ema_alpha);
    B200LP_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return B200LP_OK;
}
