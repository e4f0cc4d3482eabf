/* aur_ppo.h -- C ABI of libaurppo.so, the sm_100a implementation of the PPO hot
 * path of biirving/aur_ppo (rollout step -> GAE -> minibatch update).
 *
 * The reference has no FFI of its own: its boundary is the Python surface
 * (src/run_ppo.py:14-41 flags, `ppo(params).train()` src/ppo.py:44,169,
 * `actor_critic.evaluate/value` src/models/actor_critic.py:31-51).  Each entry
 * point below replaces the stock-PyTorch/gym code at the cited reference lines;
 * aur_ppo_b200/ (the Python host mirror) binds them with ctypes, and
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.
 *   - All data pointers are DEVICE pointers into caller-owned memory (torch
 *     tensors); `stream` is a cudaStream_t passed as void* (NULL = default).
 *     The caller sets the current CUDA device before calling.
 *   - Return 0 on success, <0 on error: AUR_ERR_ARG (bad argument),
 *     AUR_ERR_UNSUPPORTED (shape outside the compiled kernels -- there is NO
 *     CPU or library fallback), or -(1000 + cudaError_t).  aur_last_error()
 *     gives a thread-local human-readable message.
 *   - Nothing is allocated that the caller must free; functions are
 *     re-entrant per (device, stream) and launch asynchronously.
 */
#ifndef AUR_PPO_H
#define AUR_PPO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AUR_ABI_VERSION 1
#define AUR_ERR_ARG (-1)
#define AUR_ERR_UNSUPPORTED (-2)

#define AUR_ENV_CARTPOLE 0 /* gym CartPole-v1 */
#define AUR_ENV_PENDULUM 1 /* gym Pendulum-v1 */

int aur_abi_version(void);
const char* aur_last_error(void);
/* number of this library's kernels launched by the calling thread since the last reset */
int64_t aur_launch_count(void);
void aur_launch_count_reset(void);

/* ---------------------------------------------------------------- GAE ----
 * Replaces ppo.advantages -> run_gae (src/ppo.py:125-142, use_gae != 0) or
 * normal_advantage (src/ppo.py:145-157, use_gae == 0): the reverse scan over
 * the [T,N] rollout buffers with done masking, then returns = adv + values
 * (GAE) / adv = returns - values (Monte-Carlo).
 *
 *   rewards, values, terminals : [T,N] fp32 row-major (torch_buffer layout,
 *                                src/ppo.py:26-29); terminals[t] is the done
 *                                flag that arrived WITH obs[t] (ppo.py:204)
 *   next_value, next_done      : [N] fp32 (critic(next_obs), ppo.py:161)
 *   adv_out, ret_out           : [T,N] fp32
 *   gamma, gae_lambda          : Python floats of the reference; converted
 *                                exactly as torch does for tensor*scalar
 *                                ((float)gamma, (float)(gamma*gae_lambda))
 * Arithmetic is fp32, sequential in t, one rounding per reference operation
 * (no FMA contraction): results are bit-identical to the reference's torch
 * CPU loop.  T >= 1, N >= 1; N == 0 or T == 0 is a no-op returning 0. */
int aur_gae_f32(int32_t T, int64_t N, const float* rewards, const float* values, const float* terminals,
                const float* next_value, const float* next_done, double gamma, double gae_lambda,
                int32_t use_gae, float* adv_out, float* ret_out, void* stream);

/* Which kernel aur_gae_f32 would pick for this shape/alignment: 1 = bulk-async
 * (TMA 1-D) pipelined kernel, 0 = plain column kernel.  For tests/bench. */
int aur_gae_kernel_kind(int32_t T, int64_t N, const float* rewards, const float* values, const float* terminals,
                        const float* adv_out, const float* ret_out);

#ifdef __cplusplus
}
#endif
#endif /* AUR_PPO_H */
