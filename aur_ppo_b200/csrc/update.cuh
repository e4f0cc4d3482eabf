// Declarations shared by the SIMT (update.cu) and tensor-core (update_tc.cu) minibatch-update kernels.
#pragma once
#include "policy.cuh"

namespace aur {

constexpr int UPD_THREADS = 256;
constexpr int UPD_S = 256;             // samples per tile (one per thread in the per-sample phases)
constexpr int UPD_LD = UPD_S + 4;      // feature-major row stride: 16-B aligned, 4 banks per row
constexpr int UPD_H = 64;
constexpr int UPD_PSTRIDE = 4800;      // floats per CTA partial: net gradients + AUR_NUM_STATS
constexpr int UPD_STAT_OFF = UPD_PSTRIDE - AUR_NUM_STATS;
constexpr int UPD_SW = 4800;           // smem floats reserved for one net (padded layout)
constexpr int UPD_WD = 2 * UPD_H * UPD_H;   // a 64x64 matrix with every entry duplicated: [k][n][2]
constexpr int UPD_SMEM_FLOATS = UPD_SW + 2 * UPD_WD + UPD_H + 2 * UPD_H * UPD_LD + 2 * 4 * UPD_LD + 4 * UPD_H;
constexpr size_t UPD_SMEM = sizeof(float) * UPD_SMEM_FLOATS;
constexpr int MOM_CTAS = 148;

struct UpdDev {
  long long m_local;
  const int32_t* idx;
  long long idx_offset;
  const float *obs, *actions, *logprobs, *advantages, *returns, *values, *params;
  int obs_dim, act_dim, continuous, norm_adv, clip_vloss;
  float clip, clip_lo, clip_hi, ent_c, vf_c, inv_m;
  const double* moments;
  float* partials;      // [2][gridDim.x][UPD_PSTRIDE]
};


// tensor-core implementation (update_tc.cu): both nets per CTA, grid = gx CTAs; same partial layout
int launch_ppo_grad_tc(const UpdDev& d, int gx, cudaStream_t s);
size_t ppo_grad_tc_smem_bytes();

}  // namespace aur
