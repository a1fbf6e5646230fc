// Fused multi-tensor Adam: ONE launch updates every parameter of an optimizer.
//
// Replaces torch.optim.Adam of the reference's three optimizers (models/registration_model.py:114-115, 135, 168-171:
// optimizer_G / optimizer_F / optimizer_R, lr = opt.lr, betas = (opt.beta1, opt.beta2), eps 1e-8, no weight decay,
// no amsgrad).  Same update, operation for operation in fp32:
//     step += 1;  m = m + (g - m) * (1 - beta1);  v = v * beta2 + (1 - beta2) * g * g
//     p -= (lr / (1 - beta1^step)) * m / (sqrt(v) / sqrt(1 - beta2^step) + eps)
// The step counter and (optionally) the learning rate live on the device, so a captured CUDA graph keeps following
// the scheduler (the host fills the lr tensor in place) - what torch's capturable=True does with ~6 multi-tensor
// kernels per optimizer.  Work list: each CTA takes one CHUNK-element piece of one tensor, found in a device table.
#include "common.cuh"
#include "dfmir_b200.h"

namespace {

constexpr int CHUNK = 4096;    // elements per CTA (256 threads x 4 x float4)

struct AdamTensor { float* p; const float* g; float* m; float* v; long long n; };

__global__ void adam_step_inc_kernel(float* step) { step[0] += 1.0f; }

__global__ void __launch_bounds__(256)
adam_multi_kernel(const AdamTensor* __restrict__ tensors, const int2* __restrict__ work, const float* __restrict__ step_dev,
                  const float* __restrict__ lr_dev, double lr_host, double beta1d, double beta2d, double epsd) {
  const int2 wk = work[blockIdx.x];                         // {tensor, chunk}
  const AdamTensor t = tensors[wk.x];
  const long long beg = (long long)wk.y * CHUNK;
  const long long end = beg + CHUNK < t.n ? beg + CHUNK : t.n;
  // the scalars of the step in double, as torch.optim.Adam forms them on the host (1 - 0.999f would already be off
  // by 5e-5 relative): one thread, broadcast through shared memory
  __shared__ float sc[5];
  if (threadIdx.x == 0) {
    const double step = (double)step_dev[0];
    const double lr = lr_dev ? (double)lr_dev[0] : lr_host;
    const double bc1 = 1.0 - pow(beta1d, step), bc2 = 1.0 - pow(beta2d, step);
    sc[0] = (float)(lr / bc1); sc[1] = (float)sqrt(bc2); sc[2] = (float)(1.0 - beta1d); sc[3] = (float)(1.0 - beta2d); sc[4] = (float)beta2d;
  }
  __syncthreads();
  const float step_size = sc[0], bc2_sqrt = sc[1], omb1 = sc[2], omb2 = sc[3], beta2 = sc[4], eps = (float)epsd;
  auto upd = [&](float& p, float g, float& m, float& v) {
    m = m + (g - m) * omb1;
    v = v * beta2 + omb2 * (g * g);
    p -= step_size * (m / (sqrtf(v) / bc2_sqrt + eps));
  };
  const bool vec = (((uintptr_t)t.p | (uintptr_t)t.g | (uintptr_t)t.m | (uintptr_t)t.v) & 15) == 0;
  if (vec) {
    const long long end4 = beg + ((end - beg) & ~3LL);
    for (long long i = beg + 4 * threadIdx.x; i < end4; i += 4 * 256) {
      float4 p = *reinterpret_cast<float4*>(t.p + i), m = *reinterpret_cast<float4*>(t.m + i), v = *reinterpret_cast<float4*>(t.v + i);
      const float4 g = *reinterpret_cast<const float4*>(t.g + i);
      upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
      *reinterpret_cast<float4*>(t.p + i) = p; *reinterpret_cast<float4*>(t.m + i) = m; *reinterpret_cast<float4*>(t.v + i) = v;
    }
    for (long long i = end4 + threadIdx.x; i < end; i += 256) upd(t.p[i], t.g[i], t.m[i], t.v[i]);
  } else {
    for (long long i = beg + threadIdx.x; i < end; i += 256) upd(t.p[i], t.g[i], t.m[i], t.v[i]);
  }
}

}  // namespace

extern "C" int dfmir_adam_chunk_elems(void) { return CHUNK; }

// tensors: device array of n_work-independent records {param*, grad*, exp_avg*, exp_avg_sq*, int64 numel} (40 bytes each);
// work: device array of n_work {tensor index, chunk index} pairs covering every tensor in pieces of
// dfmir_adam_chunk_elems() elements; step: device float, incremented here before the update; lr_dev: device float or
// null (then lr_host is used).
extern "C" int dfmir_adam_multi(const void* tensors, const void* work, int n_work, float* step, const float* lr_dev, double lr_host,
                                double beta1, double beta2, double eps, void* stream) {
  DFMIR_CHECK_ARG(tensors && work && step && n_work >= 0, "dfmir_adam_multi: null pointer");
  DFMIR_CHECK_ARG(beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && eps >= 0., "dfmir_adam_multi: bad hyper-parameters");
  static_assert(sizeof(AdamTensor) == 40, "record layout");
  cudaStream_t st = (cudaStream_t)stream;
  adam_step_inc_kernel<<<1, 1, 0, st>>>(step);
  DFMIR_CHECK_LAUNCH("dfmir_adam_multi(step)");
  if (n_work == 0) return DFMIR_OK;
  adam_multi_kernel<<<n_work, 256, 0, st>>>((const AdamTensor*)tensors, (const int2*)work, step, lr_dev, lr_host, beta1, beta2, eps);
  DFMIR_CHECK_LAUNCH("dfmir_adam_multi");
  return DFMIR_OK;
}
