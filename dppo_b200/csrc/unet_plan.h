// Unet1D lowered to a program of dense tensor-core layers (host plan + device-visible layer records).
//
// Reference: Unet1D / ResidualBlock1D dppo/model/diffusion/unet.py:27-327, Conv1dBlock / Downsample1d / Upsample1d
// dppo/model/diffusion/modules.py:30-95.  Over the short action horizon (Ta = 4 ... 16) a Conv1d is a small banded
// block-Toeplitz matrix; the plan lowers every conv / transposed conv / Linear of the net to a dense map
// [in features] -> [out features] and describes one network evaluation as a list of layers
//     accumulate 1-2 GEMMs into TMEM  ->  epilogue (bias, GroupNorm, activation, FiLM, residual)  ->  next operand.
// Activations are indexed channel-major (feature = channel * T + t), which makes every GroupNorm group a run of
// consecutive features, i.e. consecutive TMEM lanes.
#pragma once
#include <stdint.h>

#include <vector>

#include "../../include/dppo_b200.h"

namespace dppo {

struct UGemm {
  uint32_t tile_off;      // index of its first 16 KiB tile in the per-network tile stream
  uint16_t mt, kc;        // M tiles (128 output features each), K chunks (64 input features each, all segments)
  uint16_t src_chunk[2];  // operand chunk index where each K segment starts
  uint16_t src_n[2];      // chunks per K segment (src_n[1] = 0: one segment)
  uint16_t acc_tile;      // accumulator tile index (TMEM column = acc_tile * NE)
  uint16_t pad;
};

enum : int32_t { U_EPI_OPERAND = 0, U_EPI_FILM = 1, U_EPI_EPS = 2 };
enum : int32_t { U_RES_NONE = 0, U_RES_ACC = 1, U_RES_SLOT = 2 };

struct ULayer {
  int32_t n_gemm;
  UGemm g[2];
  int32_t kind;                    // U_EPI_*
  int32_t acc_tile, mt, nf;        // accumulator tiles read by the epilogue, valid output features
  int32_t bias_off, bias_tstride;  // floats into the side table; tstride != 0: row t of a per-timestep table
  int32_t gn_size;                 // features per GroupNorm group (0 = no norm; <= 32; not a power of two: segmented path)
  int32_t gamma_off, beta_off;
  float gn_eps;
  int32_t act;                     // 1 = apply the net's activation
  int32_t film;                    // 0 none, 1 additive, 2 scale + bias (read from the FiLM buffer of this block)
  int32_t film_c, film_tshift;     // channels, log2(T): channel of feature f = f >> tshift
  int32_t res;                     // U_RES_*
  int32_t res_acc_tile, res_bias_off, res_chunk;
  int32_t dst_chunk;               // operand chunk the result is written to (U_EPI_OPERAND)
  int32_t track;                   // 0 = main path (x -> eps), 1 = FiLM conditioning encoders (depend on t and obs only)
  // main path only: the operand chunks written by the PREVIOUS main-path layer (cyclically: the first layer of a step
  // waits for the sample x written by the posterior).  The track-split kernel hands those over M tile by M tile.
  int32_t wait_chunk, wait_tiles;
};

enum : int32_t { U_PACK_LINEAR = 0, U_PACK_CONV = 1, U_PACK_CONVT = 2 };

// one GEMM's weight lowering (host description; W is resolved from the parameter list at pack time)
struct UPackJob {
  int32_t param;      // index of the weight tensor in the dppo_pack_unet parameter list
  int32_t kind;       // U_PACK_*
  int32_t out_f;      // valid output features
  int32_t Cout, Tout, out_tmajor;  // out feature r -> (co, to): tmajor ? (r % Cout, r / Cout) : (r / Tout, r % Tout)
  int32_t nseg, seg_f[2], seg_c0[2], seg_chunks[2];  // K segments: valid features, first input channel, chunks
  int32_t Tin, Cin_total;
  int32_t ks, stride, pad;
  int32_t lin_in, lin_col0;  // Linear: row stride of W and first column used
  int32_t MT, KC;
  uint32_t tile_off;
};

// value of the lowered dense matrix at (output feature r, padded input column k)
__host__ __device__ inline float unet_dense_at(const UPackJob& j, const float* W, int r, int k) {
  if (r >= j.out_f) return 0.f;
  int seg = 0, kk = k;
  if (k >= j.seg_chunks[0] * 64) {
    if (j.nseg < 2) return 0.f;
    seg = 1, kk = k - j.seg_chunks[0] * 64;
  }
  if (kk >= j.seg_f[seg]) return 0.f;
  if (j.kind == U_PACK_LINEAR) return W[(size_t)r * j.lin_in + j.lin_col0 + kk];
  const int co = j.out_tmajor ? r % j.Cout : r / j.Tout;
  const int to = j.out_tmajor ? r / j.Cout : r % j.Tout;
  const int ci = j.seg_c0[seg] + kk / j.Tin, ti = kk % j.Tin;
  if (j.kind == U_PACK_CONV) {
    const int q = ti - to * j.stride + j.pad;
    if (q < 0 || q >= j.ks) return 0.f;
    return W[((size_t)co * j.Cin_total + ci) * j.ks + q];
  }
  const int q = to - ti * j.stride + j.pad;  // ConvTranspose1d, weight [Cin][Cout][ks]
  if (q < 0 || q >= j.ks) return 0.f;
  return W[((size_t)ci * j.Cout + co) * j.ks + q];
}

enum : int32_t { U_SIDE_COPY = 0, U_SIDE_TIME = 1 };
struct USideJob {
  int32_t kind;     // U_SIDE_COPY: side[dst .. dst+pad_to) = param[0..n) zero padded
                    // U_SIDE_TIME: side[dst + t*pad_to + f] = b[f] + W[f, :e] . g(time_mlp(t)), t < K, f < n
  int32_t param;    // COPY: source tensor; TIME: the Linear's weight (bias = param + 1)
  int32_t n, dst, pad_to;
  int32_t lin_in;   // TIME: row stride of W
  int32_t act_first;  // TIME: g = activation (small encoder) else identity
};

struct UnetPlan {
  dppo_unet_desc d{};
  int K = 0, nsplit = 2;
  int D = 0, e = 0;
  int n_params = 0;
  // operand chunk map (units of 64 features)
  int chunk_x = 0, KX = 0, chunk_state = 0, KS = 0, chunk_state_act = -1, chunk_condh = 0, n_condh = 0;
  int total_chunks = 0;
  int KA = 0, MTmax = 0;   // widest activation in chunks (even), its M tiles
  int film_dim = 0;        // floats per env of the FiLM buffer
  int gn_segmented = 0;    // 1: some GroupNorm group size is not a power of two (kernel needs its partial-sum scratch)
  std::vector<ULayer> layers;
  std::vector<UPackJob> jobs;
  std::vector<USideJob> side_jobs;
  size_t n_side = 0, n_tiles = 0;
  size_t track_tiles[2] = {0, 0};  // weight tiles per evaluation on each track
  int track_layers[2] = {0, 0};
  double macs_dense = 0;   // MACs per sample and evaluation of the lowered net (zero padding excluded)
};

// returns DPPO_OK or an error (message via set_error)
int unet_build_plan(const dppo_unet_desc& d, int K, int precision, UnetPlan* out);

}  // namespace dppo
