// Fused multi-tensor optimizer step on the fp32 master parameters (trains.py:131-133 `optimizer.step()` with the optimizers of
// trains.py:225-233): ONE launch updates all 128 parameter tensors of the model (torch.optim's foreach path: ~10 launches per step
// plus host-side list handling), reads every hyper-parameter from device memory (so a captured CUDA graph follows the learning-rate
// schedule) and keeps the arithmetic of torch.optim bit for bit:
//   SGD  : g' = fma(wd, p, g); buf = buf*momentum + (1-dampening)*g' (first step: buf starts at 0); p = fma(-lr, nesterov ? g' + momentum*buf : buf, p)
//   Adam : g' = fma(wd, p, g); m = lerp(m, g', 1-beta1); v = v*beta2 + (1-beta2)*g'^2; p -= lr/(1-beta1^t) * m / (sqrt(v)/sqrt(1-beta2^t) + eps)
// The bf16 weight images the conv kernels read are re-packed from the updated masters by nunet_pack_weights_batched (second launch).
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/nunet_b200.h"

namespace nunet {

constexpr int kOptChunk = 8192;   // elements per CTA (256 threads x 32)

__global__ void __launch_bounds__(256) optim_sgd_kernel(const nunet_optim_tensor_t* __restrict__ tensors, const int2* __restrict__ chunks,
                                                        const float* __restrict__ hyper) {
    const int2 ck = chunks[blockIdx.x];
    nunet_optim_tensor_t t = tensors[ck.x];
    if (!t.grad) return;
    const bool first = (t.numel & NUNET_OPTIM_FIRST_STEP) != 0;
    t.numel &= ~NUNET_OPTIM_FIRST_STEP;
    const float lr = hyper[0], momentum = hyper[1], dampening = hyper[2], wd = hyper[3];
    const bool nesterov = hyper[4] != 0.f;
    const float nlr = -lr, gscale = 1.f - dampening;
    const long long base = (long long)ck.y * kOptChunk;
    const long long end = base + kOptChunk < t.numel ? base + kOptChunk : t.numel;
    for (long long i = base + threadIdx.x; i < end; i += 256) {
        const float p = t.param[i];
        float g = t.grad[i];
        if (wd != 0.f) g = __fmaf_rn(wd, p, g);                          // grad.add(param, alpha=weight_decay)
        float upd = g;
        if (momentum != 0.f) {
            float b = g;                                                  // first step: buf = clone(grad)
            if (!first) {
                b = __fmul_rn(t.state1[i], momentum);                     // buf.mul_(momentum)
                b = (gscale == 1.f) ? __fadd_rn(b, g) : __fmaf_rn(gscale, g, b);   // .add_(grad, alpha=1-dampening)
            }
            t.state1[i] = b;
            upd = nesterov ? __fmaf_rn(momentum, b, g) : b;               // grad.add(buf, alpha=momentum)
        }
        t.param[i] = __fmaf_rn(nlr, upd, p);                              // param.add_(upd, alpha=-lr)
    }
}

// hyper: lr, beta1, beta2, eps, weight_decay ; step: one device counter per tensor; every block reads step[tensor] + 1 as the step
// number, the counters are advanced by a separate tiny kernel launched behind this one
__global__ void __launch_bounds__(256) optim_adam_kernel(const nunet_optim_tensor_t* __restrict__ tensors, const int2* __restrict__ chunks,
                                                         const float* __restrict__ hyper, const int* __restrict__ step) {
    const int2 ck = chunks[blockIdx.x];
    const nunet_optim_tensor_t t = tensors[ck.x];
    if (!t.grad) return;
    const float lr = hyper[0], beta1 = hyper[1], beta2 = hyper[2], eps = hyper[3], wd = hyper[4];
    const double tstep = (double)(step[ck.x] + 1);     // per-tensor step count, as torch keeps it (a tensor without gradient skips the step)
    // torch.optim.adam (_multi_tensor_adam, capturable=False): python-float bias corrections, applied as fp32 scalars
    const double bc1 = 1.0 - pow((double)beta1, tstep), bc2 = 1.0 - pow((double)beta2, tstep);
    const float step_size_neg = (float)(-(double)lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    const float w1 = 1.f - beta1, w2 = 1.f - beta2;
    const long long base = (long long)ck.y * kOptChunk;
    const long long end = base + kOptChunk < t.numel ? base + kOptChunk : t.numel;
    for (long long i = base + threadIdx.x; i < end; i += 256) {
        const float p = t.param[i];
        float g = t.grad[i];
        if (wd != 0.f) g = __fmaf_rn(wd, p, g);
        const float m0 = t.state1[i];
        // exp_avg.lerp_(grad, 1-beta1): ATen lerp = weight < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w)
        const float diff = __fsub_rn(g, m0);
        const float m = (w1 < 0.5f) ? __fmaf_rn(w1, diff, m0) : __fsub_rn(g, __fmul_rn(diff, 1.f - w1));
        // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
        const float v = __fmaf_rn(__fmul_rn(w2, g), g, __fmul_rn(t.state2[i], beta2));
        t.state1[i] = m;
        t.state2[i] = v;
        const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), eps);          // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
        t.param[i] = __fmaf_rn(step_size_neg, __fdiv_rn(m, denom), p);                   // param.addcdiv_(exp_avg, denom, value=-step_size)
    }
}
__global__ void optim_step_inc_kernel(const nunet_optim_tensor_t* __restrict__ tensors, int ntensors, int* step) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ntensors; i += gridDim.x * blockDim.x)
        if (tensors[i].grad) step[i] += 1;
}

}  // namespace nunet

using namespace nunet;

extern "C" int nunet_optim_chunk_elems(void) { return kOptChunk; }

extern "C" int nunet_optim_step(const nunet_optim_tensor_t* tensors_dev, int32_t ntensors, const int32_t* chunks_dev, int32_t nchunks,
                                const float* hyper_dev, int32_t algo, int32_t* step_dev, void* stream) {
    if (!tensors_dev || ntensors < 1 || !chunks_dev || nchunks < 1 || !hyper_dev || (algo != NUNET_OPTIM_SGD && algo != NUNET_OPTIM_ADAM)) return NUNET_E_BADARG;
    if (algo == NUNET_OPTIM_ADAM && !step_dev) return NUNET_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (algo == NUNET_OPTIM_SGD) {
        optim_sgd_kernel<<<nchunks, 256, 0, st>>>(tensors_dev, reinterpret_cast<const int2*>(chunks_dev), hyper_dev);
    } else {
        optim_adam_kernel<<<nchunks, 256, 0, st>>>(tensors_dev, reinterpret_cast<const int2*>(chunks_dev), hyper_dev, step_dev);
        optim_step_inc_kernel<<<(ntensors + 255) / 256, 256, 0, st>>>(tensors_dev, ntensors, step_dev);
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e;
}
