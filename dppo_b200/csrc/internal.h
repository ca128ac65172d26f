// Host-side internals shared by the translation units of libdppo_b200.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/dppo_b200.h"

namespace dppo {

// SM count of a B200: a placeholder until the device attribute has been read (contexts and plans overwrite it) and the
// fall-back when that query fails; launch shapes are sized from the queried value
constexpr int kDefaultSmCount = 148;


void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define DPPO_CUDA(call)                                   \
  do {                                                    \
    cudaError_t e__ = (call);                             \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

// One evaluated denoising step (device table, one row per network evaluation i = 0..S-1).
//   DDPM : x0 = f0*x - f1*eps ; mu = f2*clamp(x0) + f3*x
//   DDIM : x0 = (x - f1*eps)/f0 ; eps' = (x - f0*clamp(x0))/f1 ; mu = f2*clamp(x0) + f3*eps'
struct StepRow {
  int32_t t;         // diffusion timestep fed to the time embedding
  int32_t ft;        // 1 = fine-tuned step (actor_ft unless use_base_policy)
  int32_t slot;      // chain slot the step's OUTPUT is written to, -1 = not recorded
  int32_t pad;
  float f0, f1, f2, f3;
  float std_train;   // exp(0.5 logvar) before any floor (stochastic eta)
  float f2_det, f3_det, std_det;  // DDIM deterministic (eta = 0) variants of f2, f3, sigma
};

// Geometry of the packed MLP (same for actor and actor_ft).
struct MlpGeom {
  int D, Dc_in, Dc, td, H, nb, act, ln, CH, CO;  // Dc_in = cond_dim, Dc = features entering layer 0 (cond_dim or cond_out)
  int MT, KCH;   // H/128, H/64
  int KC0;       // 64-wide K chunks of layer 0 ([x | cond])
  int KCc, MTc;  // cond_mlp layer 0: K chunks of cond_dim, M tiles of CH
  int nsplit;    // 2 = hi+lo tiles (3-MMA split), 1 = hi only
  // fp32 side table offsets (floats)
  size_t off_tb;      // [K][H] layer-0 bias incl. time embedding
  size_t off_blk;     // per block: b1[H] b2[H] (g1 be1 g2 be2)[H each if ln]
  size_t blk_stride;
  size_t off_bout;    // [128]
  size_t off_bc0;     // [CH]
  size_t off_bc1;     // [128]
  size_t n_side;
  // tile stream offsets (bytes)
  size_t off_cond_tiles;  // cond_mlp tiles (prologue), 0-sized without cond_mlp
  size_t off_step_tiles;  // per-step tiles
  size_t n_cond_tiles, n_step_tiles;  // counts of 16 KiB tiles
  size_t blob_bytes;
};

struct UnetPlan;
struct UPackJob;
struct USideJob;
struct ULayer;

struct PackedNet {
  uint8_t* tiles = nullptr;  // device
  float* side = nullptr;     // device
  bool packed = false;
  std::vector<const float*> raw;  // the caller's fp32 parameter tensors as of the last pack (device pointers)
};

}  // namespace dppo

struct dppo_ctx {
  int device = 0;
  int precision = 0;
  int sm_count = dppo::kDefaultSmCount;
  dppo_mlp_desc net{};
  dppo::MlpGeom g{};
  // schedule
  int K = 0, ft = 0, S = 0, use_ddim = 0;
  float eta = 1.f, x0_clip = 1.f, randn_clip = 3.f, final_clip = -1.f, eps_clip = -1.f, min_logprob_std = 0.1f;
  std::vector<dppo::StepRow> rows;        // host copy, S rows
  dppo::StepRow* d_rows = nullptr;        // device copy
  dppo::PackedNet nets[2];
  // Unet1D denoiser (kind == 1): host plan + its device-side job / layer tables (unet_plan.h)
  int kind = 0;  // 0 = DiffusionMLP, 1 = Unet1D
  dppo::UnetPlan* unet = nullptr;
  dppo::UPackJob* d_unet_jobs = nullptr;
  dppo::USideJob* d_unet_side_jobs = nullptr;
  dppo::ULayer* d_unet_layers = nullptr;
  const float** d_unet_params[2] = {nullptr, nullptr};
  int sample_dim = 0;  // Ta * Da of either denoiser kind
  int small_clusters = 0;  // 16-CTA clusters of the weights-stationary small-batch kernel the device co-schedules
  int chain_clusters[3][4] = {};  // co-resident clusters of the MLP chain kernel per (tile envs 16/32/64, cluster size 1/2/4/8)
  bool chain_clusters_known = false;
  int* d_nonfinite = nullptr;  // device flag OR-ed by the chain kernels when a sampled action element is not finite
  uint8_t* h_stage = nullptr;  // page-locked staging area of dppo_sample_chain_host (pageable caller buffers)
  size_t h_stage_bytes = 0;
  // completion ticket of dppo_sample_chain_host (small-batch kernel): page-locked word the kernel's last cluster stores
  // done_seq into; done_want = the running call asks for it, done_armed = the launched kernel will write it
  unsigned* h_done = nullptr;
  unsigned done_seq = 0;
  int done_want = 0, done_armed = 0;
  int force_ne = 0, force_c = 0;          // launch-shape override of the chain kernel (0 = cost model), dppo_debug_set_shape
  unsigned long long* d_prof = nullptr;  // optional cycle counters written by the chain kernel (dppo_debug_set_prof)
};
