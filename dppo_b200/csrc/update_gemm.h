// Update-path tensor-core kernels (sm_100a, tcgen05 + TMEM + bulk TMA copies): argument blocks and launchers shared by
// update_gemm.cu (kernels) and update_plan.cu (the per-minibatch program behind dppo_update_*).
//
// Replaces torch autograd over cuBLASLt for the PPO update of the MLP denoiser and the critic:
//   actor_ft forward / backward   reference dppo/model/diffusion/diffusion_vpg.py:398-461 (get_logprobs_subsample),
//                                 dppo/model/diffusion/mlp_diffusion.py:218-250, dppo/model/common/mlp.py:84-154
//   critic forward / backward     reference dppo/model/common/critic.py:40-54
//   loss.backward()               reference dppo/agent/finetune/train_ppo_diffusion_agent.py:360-364
//
// Operand images ("OpMat").  Every activation / gradient matrix X[R][F] that feeds a GEMM lives in HBM as bf16 hi + lo
// planes (x = hi + lo to 16 mantissa bits, the chain kernel's 3-product split) cut into 16 KiB tiles of 128 rows x 64
// features in the SWIZZLE_128B K-major image tcgen05.mma reads: tile (rt, fc, plane) at ((rt*FCp + fc)*2 + plane)*16 KiB,
// element (r, k) of a tile at r*128 + (((k>>3) ^ (r&7)) << 4) + (k&7)*2.  The SAME image is
//   - the K-major A operand (M = 128 rows, K = 64 features) of the forward / dgrad GEMMs (contraction over features), and
//   - an MN-major operand (MN = 64 features, K = 128 rows) of the wgrad GEMM (contraction over rows),
// so nothing is ever transposed and every operand load is a plain cp.async.bulk of a contiguous tile.  Rows >= R and
// features >= F of an image are zero (producers write whole tiles), FCp = number of 64-feature chunks rounded up to 2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dppo {

constexpr int kUActNone = 0, kUActRelu = 1, kUActMish = 2;
constexpr uint32_t kImg = 16384;  // one operand image (128 rows x 64 features, one plane)

inline int opmat_chunks(int features) { return ((features + 63) / 64 + 1) & ~1; }
inline size_t opmat_bytes(int rows, int features) {
  return size_t((rows + 127) / 128) * opmat_chunks(features) * 2 * kImg;
}

// out[r][n] = sum_k A[r][k] * Wt[n][k]  (+ epilogue), A = operand images, Wt = packed weight tiles
struct RowGemmArgs {
  const uint8_t* A;   // operand images of the input, FCa chunks per row tile
  int FCa;
  const uint8_t* B;   // packed weights: [NT][KC][plane][NTILE x 128 B]
  int R, RT;          // rows, row tiles of 128
  int KC;             // 64-wide contraction chunks
  int N, NTILE, NT;   // real output features; tile width (multiple of 16, <= 256); number of column tiles
  int NTP;            // tile width B was PACKED for (0 = NTILE); a multiple of NTILE: a call on few rows runs narrower tiles
                      // (more work units for the same bytes) on the same packed weights
  // ---- epilogue, per element v = acc[r][n].  fp32 side tensors are row-major [R][ld] (mode 0) or "tiled" (mode 1):
  // [row tile][n / 8][128 rows][8 floats], ld = feature count (multiple of 8) - the layout in which a warp of the
  // epilogue (lane = row, 8 consecutive features per access) reads / writes 1 KiB contiguous.
  const float* bias;  // [N] or null:                 v += bias[n]
  const float* pre;   // fp32 or null:                 v *= act'(p), p = pre[r][n] (or its LayerNorm, below)
  int ld_pre, pre_mode, act_grad;
  const uint32_t* mask_in;  // ReLU derivative as bits (written by a forward epilogue) instead of `pre`, or null
  const float* ln_stats;  // [R][2] (mean, rstd) or null: p = (pre - mean) * rstd * ln_g[n] + ln_b[n]
  const float* ln_g;
  const float* ln_b;
  const float* res;   // fp32 or null:                 v += res[r][n]
  int ld_res, res_mode;
  float* out_f32;     // fp32 or null
  int ld_out, out_mode;
  uint8_t* out_op;    // operand images of act_out(v) or null, feature n lands at column op_col0 + n
  int FCo, op_col0, act_out;
  float* stat_out;    // [R][2] or null, accumulated with atomics (the caller zeroes it): (sum v, sum v^2) of the fp32 output
                      // row; with ln_stats set: (sum dz g, sum dz g xhat) of the LayerNorm backward
  uint32_t* mask_out; // bit (n % 32) of word [row tile][n / 32][row] = (v > 0), or null
  int mask_words;     // 32-feature words per row of the mask tensors
  int nstage;
  unsigned long long* prof;  // optional [grid][8] role cycle counters (bring-up), null in production
};
inline size_t f32_tiled_floats(int rows, int features) { return size_t((rows + 127) / 128) * 128 * ((features + 7) / 8 * 8); }
inline size_t mask_words_total(int rows, int features) { return size_t((rows + 127) / 128) * 128 * ((features + 31) / 32); }

// dW[n][k] += sum_r G[r][n] * X[r][k];  db[n] += sum_r G[r][n]
struct WgradArgs {
  const uint8_t* G;   // images of the gradient w.r.t. the layer's output  [R][N_out]
  int FCg;
  const uint8_t* X;   // images of the layer's input                       [R][K_in]
  int FCx;
  int R;
  int n_units;        // 64-row units = ceil(R / 64)
  int MT;             // 128-feature tiles of N_out
  int NTn, NCH;       // column tiles over K_in, 64-feature chunks per column tile (<= 4)
  int S, units_per_split;
  int N_out, K_in;
  float* dW;          // fp32 [N_out][ld_dw], accumulated with red.global.add
  int ld_dw;
  float* db;          // [N_out] or null
  int nstage;
  uint64_t desc_mn;   // MN-major shared-memory descriptor flags (set by the launcher)
};

// one column segment of a gathered / packed operand matrix
struct PackSeg {
  const float* src;   // fp32 source; null = one-hot of the row's denoising index (width = ft)
  int64_t ld;         // row stride of the source (floats)
  int mode;           // 0: row r;  1: row b = idx / ft;  2: row idx (= b * ft + d, e.g. chains[b, d] with ld = D... see pack)
  int width;          // columns copied
  int dst_col;        // first destination column
};
struct PackArgs {
  PackSeg seg[4];
  int n_seg;
  const int64_t* inds;   // flat (b * ft + d) indices of the rows, or null
  const int64_t* dinds;  // per-row denoising index when inds == null (one-hot segments), may be null
  int ft;
  int64_t chain_stride;  // floats between consecutive b of the chains buffer ((ft+1) * D) for mode 2
  int64_t chain_d;       // floats between consecutive d (D)
  int R, FCp;
  const float* scale;    // optional device scalar multiplied into every copied value
  float scale_imm;       // immediate multiplier on top of it (0 is read as 1)
  uint8_t* out;
};

int launch_row_gemm(const RowGemmArgs& a, int sm_count, cudaStream_t st);
int launch_wgrad(const WgradArgs& a, int sm_count, cudaStream_t st);
int launch_pack_rows(const PackArgs& a, cudaStream_t st);
// operand images -> fp32 [R][F] (hi + lo); bring-up / tests
int launch_unpack_rows(const uint8_t* img, int FCp, int R, int F, float* out, int ld, cudaStream_t st);
// several weight matrices in one launch (jobs live in device memory)
struct PackWJob {
  const float* W;
  long long s_row, s_col;
  int rows, K, NTILE, NT, KC;
  uint8_t* out;
};
int launch_pack_weights(const PackWJob* d_jobs, int n_jobs, long long max_total, cudaStream_t st);
void set_mn_desc_override(uint32_t lbo_bytes, uint32_t sbo_bytes);
void set_row_gemm_tuning(int ntile_cap, int max_stages);  // bring-up: cap the column tile / ring depth (0 = default)
int row_gemm_ntile_cap();
void set_row_gemm_fast_epilogue(bool on);  // bring-up: force the generic epilogue
// fp32 [rows][K] (element (j, c) at W[j * s_row + c * s_col]) -> packed B tiles [NT][KC][plane][NTILE x 128 B]
int launch_pack_weight(const float* W, int64_t s_row, int64_t s_col, int rows, int K, int NTILE, uint8_t* out, cudaStream_t st);
size_t packed_weight_bytes(int rows, int K, int NTILE);
int row_gemm_ntile(int N);
// LayerNorm(eps) over F features + activation -> operand images; stats[r] = (mean, rstd)
int launch_ln_fwd(const float* x, int ld, int R, int F, const float* g, const float* b, float eps, int act, float* stats,
                  uint8_t* out_op, int FCo, cudaStream_t st);
// dz (gradient w.r.t. the LayerNorm output, fp32) -> dx (+ res) as fp32 (optional) and operand images; dg / db accumulated
int launch_ln_bwd(const float* dz, int ld_dz, const float* x, int ld_x, const float* stats, const float* g, int R, int F,
                  const float* res, int ld_res, float* out_f32, int ld_out, uint8_t* out_op, int FCo, float* dg, float* db,
                  int sm_count, cudaStream_t st);

// the same two on tiled fp32 tensors (update_gemm.h, RowGemmArgs): one CTA per 128-row tile, coalesced throughout
// `sums`: [R][2] row sums accumulated by the producing GEMM's epilogue (RowGemmArgs.stat_out)
int launch_ln_fwd_tiled(const float* x, const float* sums, int R, int F, const float* g, const float* b, float eps, int act,
                        float* stats, uint8_t* out_op, int FCo, cudaStream_t st);
int launch_ln_bwd_tiled(const float* dz, const float* sums, const float* x, const float* stats, const float* g, int R, int F,
                        const float* res, float* out_f32, uint8_t* out_op, int FCo, float* dg, float* db, cudaStream_t st);

}  // namespace dppo
