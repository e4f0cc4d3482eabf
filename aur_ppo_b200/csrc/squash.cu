// Tanh-squashed Gaussian policy head (row Q): PPOGaussianPolicyBase.sample, src/nets/nets.py:90-105.
//   action = given, or mean + exp(log_std) * N(0,1) (rsample);  y = tanh(action)
//   log_prob = sum_k [ Normal(mean, std).log_prob(action) - log(1 - y^2 + 1e-6) ]   (keepdim -> [B,1])
//   returns (y, log_prob, tanh(mean), Normal.entropy() [B,A] unsummed)
// Elementwise and HBM-bound (5 reads/writes of [B,A] fp32): one thread per row, A <= 16 dims in registers.
// Noise: Philox4x32-10 keyed by the seed, counter = (row, stream_id, dim block), Box-Muller as in policy.cuh.
#include "policy.cuh"

namespace aur {

constexpr int SQ_MAX_A = 16;

__global__ void __launch_bounds__(256) squashed_sample_kernel(long long B, int A, const float* __restrict__ mean,
                                                            const float* __restrict__ log_std, const float* __restrict__ action_in,
                                                            uint64_t seed, uint64_t stream_id, float* __restrict__ y_out,
                                                            float* __restrict__ logp_out, float* __restrict__ mean_out,
                                                            float* __restrict__ ent_out, float* __restrict__ pre_out) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float LOG_SQRT_2PI = 0.91893853320467267f;
  float lp = 0.0f;
  for (int k0 = 0; k0 < A; k0 += 4) {
    float z[POL_OUT_MAX] = {0.f, 0.f, 0.f, 0.f};
    if (!action_in) {
      const Philox r = philox4x32_10((uint32_t)b, (uint32_t)((uint64_t)b >> 32), (uint32_t)stream_id,
                                     (uint32_t)(stream_id >> 32) ^ ((uint32_t)(k0 >> 2) << 24), (uint32_t)seed, (uint32_t)(seed >> 32));
      normal4(r, z);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + j;
      if (k < A) {
        const long long e = b * A + k;
        const float mu = mean[e], ls = log_std[e], sd = expf(ls);
        const float x = action_in ? action_in[e] : mu + sd * z[j];
        const float y = tanhf(x);
        const float d = x - mu;
        lp += (-(d * d) / (2.0f * (sd * sd)) - logf(sd) - LOG_SQRT_2PI) - logf((1.0f - y * y) + 1e-6f);
        y_out[e] = y;
        mean_out[e] = tanhf(mu);
        ent_out[e] = 0.5f + LOG_SQRT_2PI + logf(sd);
        if (pre_out) pre_out[e] = x;
      }
    }
  }
  logp_out[b] = lp;
}

}  // namespace aur

extern "C" int aur_squashed_gaussian_sample(int64_t B, int32_t A, const float* mean, const float* log_std, const float* action_in,
                                            uint64_t seed, uint64_t stream_id, float* action_out, float* logp_out, float* mean_out,
                                            float* entropy_out, float* pre_tanh_out, void* stream) {
  using namespace aur;
  if (B < 0 || A < 1 || A > SQ_MAX_A) { set_error("aur_squashed_gaussian_sample: need B >= 0 and 1 <= A <= %d", SQ_MAX_A); return AUR_ERR_ARG; }
  if (B == 0) return 0;
  if (!mean || !log_std || !action_out || !logp_out || !mean_out || !entropy_out) {
    set_error("aur_squashed_gaussian_sample: null buffer"); return AUR_ERR_ARG;
  }
  squashed_sample_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>((long long)B, A, mean, log_std, action_in, seed,
                                                                                     stream_id, action_out, logp_out, mean_out,
                                                                                     entropy_out, pre_tanh_out);
  AUR_LAUNCH_OK("squashed_sample_kernel");
  return 0;
}
