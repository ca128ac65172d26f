// Update-path tensor-core kernels (sm_100a): see update_gemm.h for the operand-image format and the reference lines.
//
//   ugemm_rows_kernel   out[r][n] = sum_k A[r][k] Wt[n][k]  with a fused epilogue (bias, activation derivative of a saved
//                       pre-activation [optionally through LayerNorm], residual add, fp32 store, activation + bf16 hi/lo
//                       split into the next GEMM's operand images).  Forward Linear layers and dgrad (Wt = W^T tiles).
//                       Persistent CTAs over 128 x NTILE output tiles; A (rows x 64 features, hi + lo) and Wt tiles arrive by
//                       cp.async.bulk into a shared-memory ring; three tcgen05.mma products per K step
//                       (a_hi w_hi + a_hi w_lo + a_lo w_hi, fp32 accumulate) into one of two TMEM accumulators, so the
//                       epilogue of tile i runs under the MMAs of tile i + 1.
//   ugemm_wgrad_kernel  dW[n][k] += sum_r G[r][n] X[r][k]: the same images read as MN-major operands (contraction over
//                       rows), split over row ranges, accumulated with vector red.global.add; the bias gradient is one more
//                       N = 16 MMA against a tile of ones.
//   pack_rows_kernel    gathers (obs[b], chains[b, d], one-hot(d), plain matrices) into operand images
//   pack_weight_kernel  fp32 weights -> K-major B tiles (optionally transposed: dgrad)
//   ln_fwd / ln_bwd     LayerNorm + activation around the GEMMs (row statistics need the whole feature row)
// Warp roles of the two GEMM kernels (320 threads): warp 0 producer, warp 1 MMA issuer + TMEM allocator, warps 2..9 epilogue.
#include <set>

#include "common.cuh"
#include "internal.h"
#include "update_gemm.h"

namespace dppo {

constexpr int kUThreads = 320;
constexpr int kUEpiThreads = 256;
constexpr uint32_t kHalfImg = 8192;  // 64 rows of an operand image

// MN-major SWIZZLE_128B descriptor of operand images stacked 8 KiB apart: 64 features (128 B) contiguous, 8-row groups
// 1024 B apart (stride byte offset), 64-feature blocks 8192 B apart (leading byte offset).  cute/atom/mma_traits_sm100.hpp:
// Swizzle<3,4,3> o ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units.
constexpr uint64_t kDescSw128MnMajor =
    (uint64_t(kHalfImg >> 4) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
constexpr uint32_t kIdescMnMajor = (1u << 15) | (1u << 16);  // A and B both MN-major

// four K steps of one smem stage in one statement (K-major: +32 B per step, MN-major: +2048 B), optional ring release
__device__ __forceinline__ void umma_x4_step_p(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc_first,
                                               uint32_t leader, uint32_t commit_bar, uint32_t step16) {
  asm volatile(
      "{\n\t.reg .pred p, q, t, c;\n\t.reg .b64 s, a1, a2, a3, b1, b2, b3;\n\t"
      "setp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\tsetp.eq.u32 t, 1, 1;\n\t"
      "setp.ne.and.b32 c, %6, 0, q;\n\t"
      "cvt.u64.u32 s, %7;\n\t"
      "add.u64 a1, %1, s;\n\tadd.u64 a2, a1, s;\n\tadd.u64 a3, a2, s;\n\t"
      "add.u64 b1, %2, s;\n\tadd.u64 b2, b1, s;\n\tadd.u64 b3, b2, s;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, t;\n\t"
      "@c tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t}" ::"r"(d_tmem),
      "l"(a), "l"(b), "r"(idesc), "r"(acc_first), "r"(leader), "r"(commit_bar), "r"(step16)
      : "memory");
}

__device__ __forceinline__ float relu_f(float x) { return fmaxf(x, 0.f); }
// d/dx [x tanh(softplus(x))] = tanh(sp) + x sigmoid(x) (1 - tanh(sp)^2)   (ATen's mish_backward formula)
__device__ __forceinline__ float mish_grad_f(float x) {
  float e, r1, r2;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(x, 20.0f) * 1.4426950408889634f));
  const float n = e * (e + 2.0f);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(n + 2.0f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r2) : "f"(e + 1.0f));
  const float th = n * r1;
  return th + x * (e * r2) * (1.0f - th * th);
}
__device__ __forceinline__ float act_apply(int act, float x) {
  return act == kUActRelu ? relu_f(x) : (act == kUActMish ? mish_f(x) : x);
}
__device__ __forceinline__ float act_grad(int act, float p) {
  return act == kUActRelu ? (p > 0.f ? 1.f : 0.f) : (act == kUActMish ? mish_grad_f(p) : 1.f);
}

// 8 consecutive features of one row -> one 16-byte chunk of the hi image and one of the lo image
__device__ __forceinline__ void store_op8(uint8_t* img_hi, uint32_t rloc, uint32_t kk, const float (&y)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(y[2 * i], y[2 * i + 1]);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(y[2 * i] - __low2float(h2), y[2 * i + 1] - __high2float(h2));
    h[i] = *reinterpret_cast<const uint32_t*>(&h2);
    l[i] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  const uint32_t off = rloc * 128u + ((((kk >> 3) ^ (rloc & 7u))) << 4);
  *reinterpret_cast<uint4*>(img_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(img_hi + kImg + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void load8(const float* p, bool vec, int n, int N, float (&o)[8]) {
  if (vec && n + 7 < N) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    o[0] = a.x, o[1] = a.y, o[2] = a.z, o[3] = a.w, o[4] = b.x, o[5] = b.y, o[6] = b.z, o[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = n + i < N ? p[i] : 0.f;
  }
}
__device__ __forceinline__ void store8(float* p, bool vec, int n, int N, const float (&o)[8]) {
  if (vec && n + 7 < N) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(o[4], o[5], o[6], o[7]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (n + i < N) p[i] = o[i];
  }
}

// ============================================================================================== row GEMM (forward / dgrad)
struct UBars {
  uint64_t *full, *empty, *tfull, *tempty;
  uint32_t* tmem_slot;
};
constexpr int kUMaxStages = 6;
__device__ __forceinline__ UBars carve_bars(uint8_t* p) {
  UBars b;
  b.full = reinterpret_cast<uint64_t*>(p);
  b.empty = b.full + kUMaxStages;
  b.tfull = b.empty + kUMaxStages;
  b.tempty = b.tfull + 2;
  b.tmem_slot = reinterpret_cast<uint32_t*>(b.tempty + 2);
  return b;
}
constexpr uint32_t kUBarBytes = (2 * kUMaxStages + 4) * 8 + 16;

// address of the 8 floats (row, features n .. n+7) of an fp32 side tensor
__device__ __forceinline__ size_t side_off(int mode, int ld, int row, int n) {
  return mode ? ((size_t(row >> 7) * (ld >> 3) + (n >> 3)) * 128 + (row & 127)) * 8 : size_t(row) * ld + n;
}

// One 128 x NTILE accumulator tile.  Thread = one row (TMEM lane) x 32 columns of each 64-column chunk; the operand
// images of a chunk are assembled in shared memory (`stage`, 32 KiB: hi image then lo image, the global tile layout) and
// copied out with fully coalesced 16-byte stores by all 256 epilogue threads.
__device__ __forceinline__ void rows_epilogue(const RowGemmArgs& a, uint32_t tmem_d, int rt, int nt, int warp, int lane,
                                              uint8_t* stage, long long (&pc)[4]) {
  const int q = warp & 3, half = (warp - 2) >> 2;
  const uint32_t rloc = uint32_t(q * 32 + lane);
  const int row = rt * 128 + int(rloc);
  const bool valid = row < a.R;
  const int et = int(threadIdx.x) - 64;
  const int ncc = (a.NTILE + 63) >> 6;
  float mean = 0.f, rstd = 1.f;
  if (a.ln_stats && valid) mean = a.ln_stats[2 * size_t(row)], rstd = a.ln_stats[2 * size_t(row) + 1];
  const bool vec_pre = a.pre_mode || ((reinterpret_cast<uintptr_t>(a.pre) & 15) == 0 && (a.ld_pre & 3) == 0);
  const bool vec_res = a.res_mode || ((reinterpret_cast<uintptr_t>(a.res) & 15) == 0 && (a.ld_res & 3) == 0);
  const bool vec_out = a.out_mode || ((reinterpret_cast<uintptr_t>(a.out_f32) & 15) == 0 && (a.ld_out & 3) == 0);
  for (int cc = 0; cc < ncc; ++cc) {
    const int cb = cc * 64 + half * 32;  // first column (within the tile) of this thread's 32
    float v[32];
    const long long tp0 = clock64();
    if (cb < a.NTILE) {
      float lo16[16];
      tmem_ld16(tmem_d + (uint32_t(q * 32) << 16) + uint32_t(cb), lo16);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = lo16[i];
    }
    if (cb + 16 < a.NTILE) {
      float hi16[16];
      tmem_ld16(tmem_d + (uint32_t(q * 32) << 16) + uint32_t(cb + 16), hi16);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[16 + i] = hi16[i];
    }
    const int n0 = nt * a.NTILE + cb;
    const long long tp1 = clock64();
    pc[0] += tp1 - tp0;
    uint32_t mword = 0, mout = 0;
    if (a.mask_in && valid && n0 < a.N) mword = a.mask_in[(size_t(rt) * a.mask_words + (n0 >> 5)) * 128 + rloc];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int n = n0 + g * 8;  // first output feature of this group
      const bool live = cb + g * 8 < a.NTILE && n < a.N;
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = live ? v[g * 8 + i] : 0.f;
      if (live) {
        if (a.bias) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (n + i < a.N) x[i] += __ldg(a.bias + n + i);
        }
        if (a.mask_in) {
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = ((mword >> (g * 8 + i)) & 1u) ? x[i] : 0.f;
        } else if (a.pre) {
          float p[8];
          if (valid) load8(a.pre + side_off(a.pre_mode, a.ld_pre, row, n), vec_pre, n, a.N, p);
          else {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = 0.f;
          }
          if (a.ln_stats) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int nn = n + i < a.N ? n + i : a.N - 1;
              p[i] = (p[i] - mean) * rstd * __ldg(a.ln_g + nn) + __ldg(a.ln_b + nn);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] *= act_grad(a.act_grad, p[i]);
        }
        if (a.res && valid) {
          float r8[8];
          load8(a.res + side_off(a.res_mode, a.ld_res, row, n), vec_res, n, a.N, r8);
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] += r8[i];
        }
        if (a.out_f32 && valid) store8(a.out_f32 + side_off(a.out_mode, a.ld_out, row, n), vec_out, n, a.N, x);
#pragma unroll
        for (int i = 0; i < 8; ++i) mout |= (valid && n + i < a.N && x[i] > 0.f) ? (1u << (g * 8 + i)) : 0u;
      }
      if (a.out_op) {
        float y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = (live && valid && n + i < a.N) ? act_apply(a.act_out, x[i]) : 0.f;
        store_op8(stage, rloc, uint32_t(half * 32 + g * 8), y);
      }
    }
    if (a.mask_out && n0 < a.N && cb < a.NTILE) a.mask_out[(size_t(rt) * a.mask_words + (n0 >> 5)) * 128 + rloc] = mout;
    const long long tp2 = clock64();
    pc[1] += tp2 - tp1;
    if (a.out_op) {
      named_bar_sync(1, kUEpiThreads);  // the chunk's two images are complete in shared memory
      const long long tp3 = clock64();
      pc[2] += tp3 - tp2;
      const int fc = ((a.op_col0 + nt * a.NTILE) >> 6) + cc;
      uint4* dst = reinterpret_cast<uint4*>(a.out_op + (size_t(rt) * a.FCo + fc) * 2 * kImg);
      const uint4* src = reinterpret_cast<const uint4*>(stage);
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[et + i * kUEpiThreads] = src[et + i * kUEpiThreads];
      named_bar_sync(1, kUEpiThreads);  // everybody has read the staging buffer before the next chunk overwrites it
      pc[3] += clock64() - tp3;
    }
  }
}


__device__ __forceinline__ bool row_valid(const RowGemmArgs& a, int rt, int warp, int lane) {
  return rt * 128 + (warp & 3) * 32 + lane < a.R;
}

// ---------------------------------------------------------------------------------------------- specialised epilogue
// The epilogue is instruction-issue bound (8 warps on 4 schedulers, 32 accumulator values per thread and chunk): the
// generic version above spends ~2.5 k cycles per 64-column chunk on run-time flag checks alone (role counters in
// profiles/r2).  The variants the update program actually launches are therefore compiled as straight-line code:
//   V & 1        bias            (V >> 1) & 3   0 none | 1 ReLU bit mask in | 2 relu'(pre) | 3 mish'(pre)   (pre: tiled fp32)
//   (V >> 3) & 1 residual in     (V >> 4) & 1   fp32 out (tiled)       (V >> 5) & 3   activation of the operand output
//   (V >> 7) & 1 bit mask out    (V >> 8) & 1   operand-image output   (V >> 9) & 1   pre goes through LayerNorm first
// Preconditions (checked by the launcher): NTILE and N multiples of 64, side tensors tiled.
constexpr int kEpiGeneric = -1;
constexpr int epi_variant(bool bias, int pre, bool res, bool outf, int acto, bool masko, bool outop = true, bool ln = false,
                          bool stats = false) {
  return (bias ? 1 : 0) | (pre << 1) | ((res ? 1 : 0) << 3) | ((outf ? 1 : 0) << 4) | (acto << 5) | ((masko ? 1 : 0) << 7) |
         ((outop ? 1 : 0) << 8) | ((ln ? 1 : 0) << 9) | ((stats ? 1 : 0) << 10);
}

__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

template <int V>
__device__ __forceinline__ void rows_epilogue_fast(const RowGemmArgs& a, uint32_t tmem_d, int rt, int nt, int warp, int lane,
                                                   uint8_t* stage) {
  constexpr bool BIAS = V & 1, RES = (V >> 3) & 1, OUTF = (V >> 4) & 1, MASKO = (V >> 7) & 1, OUTOP = (V >> 8) & 1, LN = (V >> 9) & 1;
  constexpr bool STATS = (V >> 10) & 1;
  constexpr int PRE = (V >> 1) & 3, ACTO = (V >> 5) & 3;
  float mean = 0.f, rstd = 1.f;
  if (LN && row_valid(a, rt, warp, lane)) {
    const int row_ = rt * 128 + (warp & 3) * 32 + lane;
    mean = a.ln_stats[2 * size_t(row_)], rstd = a.ln_stats[2 * size_t(row_) + 1];
  }
  const int q = warp & 3, half = (warp - 2) >> 2;
  const uint32_t rloc = uint32_t(q * 32 + lane);
  const int row = rt * 128 + int(rloc);
  const bool valid = row < a.R;
  const int et = int(threadIdx.x) - 64;
  const int ncc = a.NTILE >> 6;
  const int G = a.N >> 3;  // 8-feature groups per row of the tiled side tensors (all of width N)
  for (int cc = 0; cc < ncc; ++cc) {
    const int cb = cc * 64 + half * 32;
    const int n0 = nt * a.NTILE + cb;
    uint32_t raw[32];
    tmem_ld32_issue(tmem_d + (uint32_t(q * 32) << 16) + uint32_t(cb), raw);
    // side inputs are requested before the accumulator is needed
    const size_t side = ((size_t(rt) * G + (n0 >> 3)) * 128 + rloc) * 8;  // group g of this chunk: + g * 1024 floats
    float4 pr[8], rs[8];
    uint32_t mword = 0;
    if (valid) {
      if (PRE == 1) mword = a.mask_in[(size_t(rt) * a.mask_words + (n0 >> 5)) * 128 + rloc];
      if (PRE >= 2) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          pr[2 * g] = *reinterpret_cast<const float4*>(a.pre + side + g * 1024);
          pr[2 * g + 1] = *reinterpret_cast<const float4*>(a.pre + side + g * 1024 + 4);
        }
      }
      if (RES) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          rs[2 * g] = *reinterpret_cast<const float4*>(a.res + side + g * 1024);
          rs[2 * g + 1] = *reinterpret_cast<const float4*>(a.res + side + g * 1024 + 4);
        }
      }
    }
    float x[32];
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(raw[i]);
    if (BIAS) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + n0) + j);
        x[4 * j] += b4.x, x[4 * j + 1] += b4.y, x[4 * j + 2] += b4.z, x[4 * j + 3] += b4.w;
      }
    }
    uint32_t mout = 0;
    float st1 = 0.f, st2 = 0.f;
    if (valid) {
      if (PRE == 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = (mword & (1u << i)) ? x[i] : 0.f;
      } else if (PRE >= 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float p4[4] = {pr[j].x, pr[j].y, pr[j].z, pr[j].w};
          float xh[4] = {0.f, 0.f, 0.f, 0.f}, gm[4] = {0.f, 0.f, 0.f, 0.f};
          if (LN) {
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.ln_g + n0) + j);
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.ln_b + n0) + j);
            gm[0] = g4.x, gm[1] = g4.y, gm[2] = g4.z, gm[3] = g4.w;
            const float bt[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) xh[i] = (p4[i] - mean) * rstd, p4[i] = xh[i] * gm[i] + bt[i];
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            x[4 * j + i] = PRE == 2 ? (p4[i] > 0.f ? x[4 * j + i] : 0.f) : x[4 * j + i] * mish_grad_f(p4[i]);
            if (LN && STATS) {
              const float d = x[4 * j + i] * gm[i];
              st1 += d, st2 += d * xh[i];
            }
          }
        }
      }
      if (RES) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[4 * j] += rs[j].x, x[4 * j + 1] += rs[j].y, x[4 * j + 2] += rs[j].z, x[4 * j + 3] += rs[j].w;
      }
      if (STATS && !LN) {
#pragma unroll
        for (int i = 0; i < 32; ++i) st1 += x[i], st2 += x[i] * x[i];
      }
      if (STATS) {
        atomicAdd(a.stat_out + 2 * size_t(row), st1);
        atomicAdd(a.stat_out + 2 * size_t(row) + 1, st2);
      }
      if (OUTF) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          *reinterpret_cast<float4*>(a.out_f32 + side + g * 1024) = make_float4(x[8 * g], x[8 * g + 1], x[8 * g + 2], x[8 * g + 3]);
          *reinterpret_cast<float4*>(a.out_f32 + side + g * 1024 + 4) = make_float4(x[8 * g + 4], x[8 * g + 5], x[8 * g + 6], x[8 * g + 7]);
        }
      }
      if (MASKO) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mout |= x[i] > 0.f ? (1u << i) : 0u;
        a.mask_out[(size_t(rt) * a.mask_words + (n0 >> 5)) * 128 + rloc] = mout;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = ACTO == kUActRelu ? fmaxf(x[i], 0.f) : (ACTO == kUActMish ? mish_f(x[i]) : x[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = 0.f;
    }
    if (OUTOP) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float y[8] = {x[8 * g], x[8 * g + 1], x[8 * g + 2], x[8 * g + 3], x[8 * g + 4], x[8 * g + 5], x[8 * g + 6], x[8 * g + 7]};
        store_op8(stage, rloc, uint32_t(half * 32 + g * 8), y);
      }
      named_bar_sync(1, kUEpiThreads);
      const int fc = ((a.op_col0 + nt * a.NTILE) >> 6) + cc;
      uint4* dst = reinterpret_cast<uint4*>(a.out_op + (size_t(rt) * a.FCo + fc) * 2 * kImg);
      const uint4* src = reinterpret_cast<const uint4*>(stage);
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[et + i * kUEpiThreads] = src[et + i * kUEpiThreads];
      named_bar_sync(1, kUEpiThreads);
    }
  }
}

template <int V>
__global__ void __launch_bounds__(kUThreads, 1) ugemm_rows_kernel(const RowGemmArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t b_plane = uint32_t(a.NTILE) * 128u;
  const uint32_t stage_bytes = 2 * kImg + 2 * b_plane;
  uint8_t* epi_stage = smem + size_t(a.nstage) * stage_bytes;  // 32 KiB: operand images of one output chunk
  const UBars bar = carve_bars(epi_stage + 2 * kImg);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < a.nstage; ++i) mbar_init(&bar.full[i], 1), mbar_init(&bar.empty[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&bar.tfull[i], 1), mbar_init(&bar.tempty[i], kUEpiThreads);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(bar.tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *bar.tmem_slot;
  const int n_tiles = a.RT * a.NT;

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ producer
    uint32_t stage = 0, phase = 0;
    long long w_empty = 0;
    const long long t_begin = clock64();
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int rt = t / a.NT, nt = t - rt * a.NT;
      const uint8_t* a_src = a.A + size_t(rt) * a.FCa * 2 * kImg;
      // packed tile that holds this (possibly narrower) tile, and the rows of it
      const uint32_t p_plane = uint32_t(a.NTP) * 128u;
      const int sub_per = a.NTP / a.NTILE, pnt = nt / sub_per;
      const uint8_t* b_src = a.B + size_t(pnt) * a.KC * 2 * p_plane + size_t(nt - pnt * sub_per) * b_plane;
      for (int kc = 0; kc < a.KC; ++kc) {
        const long long tw = clock64();
        mbar_wait(&bar.empty[stage], phase ^ 1);
        w_empty += clock64() - tw;
        if (lane == 0) {
          uint8_t* dst = smem + size_t(stage) * stage_bytes;
          mbar_arrive_expect_tx(&bar.full[stage], stage_bytes);
          bulk_g2s(dst, a_src + size_t(kc) * 2 * kImg, 2 * kImg, &bar.full[stage]);
          if (sub_per == 1) {
            bulk_g2s(dst + 2 * kImg, b_src + size_t(kc) * 2 * b_plane, 2 * b_plane, &bar.full[stage]);
          } else {  // hi and lo rows of the narrow tile are b_plane bytes each, p_plane apart in the packed tile
            bulk_g2s(dst + 2 * kImg, b_src + size_t(kc) * 2 * p_plane, b_plane, &bar.full[stage]);
            bulk_g2s(dst + 2 * kImg + b_plane, b_src + size_t(kc) * 2 * p_plane + p_plane, b_plane, &bar.full[stage]);
          }
        }
        __syncwarp();
        if (++stage == uint32_t(a.nstage)) stage = 0, phase ^= 1;
      }
    }
    if (a.prof && lane == 0) a.prof[blockIdx.x * 16 + 0] = w_empty, a.prof[blockIdx.x * 16 + 1] = clock64() - t_begin;
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    long long w_full = 0, w_tempty = 0;
    const long long t_begin = clock64();
    const uint32_t idesc = umma_idesc_bf16(128, a.NTILE);
    const uint32_t leader = elect_one() ? 1u : 0u;
    uint32_t stage = 0, phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const uint32_t buf = uint32_t(it) & 1u, use = uint32_t(it) >> 1;
      long long tw = clock64();
      mbar_wait(&bar.tempty[buf], (use & 1u) ^ 1u);  // the epilogue drained this accumulator
      w_tempty += clock64() - tw;
      tc_fence_after();
      const uint32_t d = tmem + buf * 256u;
      for (int kc = 0; kc < a.KC; ++kc) {
        tw = clock64();
        mbar_wait(&bar.full[stage], phase);
        w_full += clock64() - tw;
        tc_fence_after();
        const uint32_t base = smem_u32(smem + size_t(stage) * stage_bytes);
        const uint64_t a_hi = umma_desc(base), a_lo = umma_desc(base + kImg);
        const uint64_t b_hi = umma_desc(base + 2 * kImg), b_lo = umma_desc(base + 2 * kImg + b_plane);
        umma_x4_step_p(d, a_hi, b_hi, idesc, kc > 0 ? 1u : 0u, leader, 0u, 2u);
        umma_x4_step_p(d, a_hi, b_lo, idesc, 1u, leader, 0u, 2u);
        umma_x4_step_p(d, a_lo, b_hi, idesc, 1u, leader, smem_u32(&bar.empty[stage]), 2u);
        if (++stage == uint32_t(a.nstage)) stage = 0, phase ^= 1;
      }
      umma_commit_p(&bar.tfull[buf], leader);
    }
    if (a.prof && lane == 0)
      a.prof[blockIdx.x * 16 + 2] = w_full, a.prof[blockIdx.x * 16 + 3] = w_tempty, a.prof[blockIdx.x * 16 + 4] = clock64() - t_begin;
  } else {
    // ------------------------------------------------------------------------------------------ epilogue
    long long w_tfull = 0, w_arrive = 0;
    long long pc[4] = {0, 0, 0, 0};
    const long long t_begin = clock64();
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int rt = t / a.NT, nt = t - rt * a.NT;
      const uint32_t buf = uint32_t(it) & 1u, use = uint32_t(it) >> 1;
      const long long tw = clock64();
      mbar_wait(&bar.tfull[buf], use & 1u);
      w_tfull += clock64() - tw;
      tc_fence_after();
      if (V == kEpiGeneric) rows_epilogue(a, tmem + buf * 256u, rt, nt, warp, lane, epi_stage, pc);
      else rows_epilogue_fast<V>(a, tmem + buf * 256u, rt, nt, warp, lane, epi_stage);
      const long long ta = clock64();
      tc_fence_before();
      mbar_arrive(&bar.tempty[buf]);
      w_arrive += clock64() - ta;
    }
    if (a.prof && threadIdx.x == 64) {
      unsigned long long* o = a.prof + blockIdx.x * 16;
      o[5] = w_tfull, o[6] = clock64() - t_begin, o[7] = w_arrive;
      o[8] = pc[0], o[9] = pc[1], o[10] = pc[2], o[11] = pc[3];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ============================================================================================== wgrad
__global__ void __launch_bounds__(kUThreads, 1) ugemm_wgrad_kernel(const WgradArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t stage_bytes = 4 * kHalfImg + uint32_t(a.NCH) * 2 * kHalfImg;
  uint8_t* ones = smem + size_t(a.nstage) * stage_bytes;  // 8 KiB of bf16 1.0 (B operand of the bias-gradient MMA)
  const UBars bar = carve_bars(ones + kHalfImg);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < a.nstage; ++i) mbar_init(&bar.full[i], 1), mbar_init(&bar.empty[i], 1);
    mbar_init(&bar.tfull[0], 1), mbar_init(&bar.tempty[0], kUEpiThreads);
    fence_mbar_init();
  }
  for (uint32_t i = threadIdx.x; i < kHalfImg / 4; i += kUThreads) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;
  fence_proxy_async_smem();
  if (warp == 1) tmem_alloc(bar.tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *bar.tmem_slot;
  const int per_split = a.MT * a.NTn;
  const int n_items = per_split * a.S;
  const int fcx_used = (a.K_in + 63) / 64;

  if (warp == 0) {
    uint32_t stage = 0, phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int s = item / per_split, rem = item - s * per_split;
      const int mt = rem / a.NTn, nt = rem - mt * a.NTn;
      const int u0 = s * a.units_per_split, u1 = min(a.n_units, u0 + a.units_per_split);
      const int nch = min(a.NCH, fcx_used - nt * a.NCH);
      for (int u = u0; u < u1; ++u) {
        mbar_wait(&bar.empty[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* dst = smem + size_t(stage) * stage_bytes;
          mbar_arrive_expect_tx(&bar.full[stage], 4 * kHalfImg + uint32_t(nch) * 2 * kHalfImg);
          const size_t half_off = size_t(u & 1) * kHalfImg;
          const uint8_t* g_tile = a.G + (size_t(u >> 1) * a.FCg + size_t(mt) * 2) * 2 * kImg + half_off;
          for (int j = 0; j < 2; ++j) {
            bulk_g2s(dst + j * kHalfImg, g_tile + size_t(j) * 2 * kImg, kHalfImg, &bar.full[stage]);
            bulk_g2s(dst + (2 + j) * kHalfImg, g_tile + size_t(j) * 2 * kImg + kImg, kHalfImg, &bar.full[stage]);
          }
          const uint8_t* x_tile = a.X + (size_t(u >> 1) * a.FCx + size_t(nt) * a.NCH) * 2 * kImg + half_off;
          uint8_t* xb = dst + 4 * kHalfImg;
          for (int j = 0; j < nch; ++j) {
            bulk_g2s(xb + j * kHalfImg, x_tile + size_t(j) * 2 * kImg, kHalfImg, &bar.full[stage]);
            bulk_g2s(xb + (a.NCH + j) * kHalfImg, x_tile + size_t(j) * 2 * kImg + kImg, kHalfImg, &bar.full[stage]);
          }
        }
        __syncwarp();
        if (++stage == uint32_t(a.nstage)) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t ones_lo = smem_u32(ones);
    uint32_t stage = 0, phase = 0;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int s = item / per_split, rem = item - s * per_split;
      const int mt = rem / a.NTn, nt = rem - mt * a.NTn;
      (void)mt;
      const int u0 = s * a.units_per_split, u1 = min(a.n_units, u0 + a.units_per_split);
      if (u0 >= u1) continue;
      const int nch = min(a.NCH, fcx_used - nt * a.NCH);
      const uint32_t idesc = umma_idesc_bf16(128, nch * 64) | kIdescMnMajor;
      const uint32_t idesc_b = umma_idesc_bf16(128, 16) | kIdescMnMajor;
      const bool with_bias = a.db != nullptr && nt == 0;
      mbar_wait(&bar.tempty[0], (uint32_t(it) & 1u) ^ 1u);
      tc_fence_after();
      for (int u = u0; u < u1; ++u) {
        mbar_wait(&bar.full[stage], phase);
        tc_fence_after();
        const uint32_t base = smem_u32(smem + size_t(stage) * stage_bytes);
        const uint64_t g_hi = umma_desc(base, a.desc_mn), g_lo = umma_desc(base + 2 * kHalfImg, a.desc_mn);
        const uint64_t x_hi = umma_desc(base + 4 * kHalfImg, a.desc_mn);
        const uint64_t x_lo = umma_desc(base + (4 + a.NCH) * kHalfImg, a.desc_mn);
        const uint32_t first = u > u0 ? 1u : 0u;
        if (with_bias) {
          const uint64_t o = umma_desc(ones_lo, a.desc_mn);
          umma_x4_step_p(tmem + 256u, g_hi, o, idesc_b, first, leader, 0u, 128u);
          umma_x4_step_p(tmem + 256u, g_lo, o, idesc_b, 1u, leader, 0u, 128u);
        }
        umma_x4_step_p(tmem, g_hi, x_hi, idesc, first, leader, 0u, 128u);
        umma_x4_step_p(tmem, g_hi, x_lo, idesc, 1u, leader, 0u, 128u);
        umma_x4_step_p(tmem, g_lo, x_hi, idesc, 1u, leader, smem_u32(&bar.empty[stage]), 128u);
        if (++stage == uint32_t(a.nstage)) stage = 0, phase ^= 1;
      }
      umma_commit_p(&bar.tfull[0], leader);
      ++it;
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int rloc = q * 32 + lane;
    const bool vec = (reinterpret_cast<uintptr_t>(a.dW) & 15) == 0 && (a.ld_dw & 3) == 0;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int s = item / per_split, rem = item - s * per_split;
      const int mt = rem / a.NTn, nt = rem - mt * a.NTn;
      const int u0 = s * a.units_per_split, u1 = min(a.n_units, u0 + a.units_per_split);
      if (u0 >= u1) continue;
      const int nch = min(a.NCH, fcx_used - nt * a.NCH);
      mbar_wait(&bar.tfull[0], uint32_t(it) & 1u);
      tc_fence_after();
      const int n = mt * 128 + rloc;
      const bool valid = n < a.N_out;
      for (int c = half; c < nch * 2; c += 2) {
        float v[32];
        tmem_ld32(tmem + (uint32_t(q * 32) << 16) + uint32_t(c * 32), v);
        const int k0 = (nt * a.NCH) * 64 + c * 32;
        if (valid) {
          float* dst = a.dW + size_t(n) * a.ld_dw + k0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (vec && k0 + j + 3 < a.K_in) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]),
                           "f"(v[j + 2]), "f"(v[j + 3])
                           : "memory");
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (k0 + j + i < a.K_in) atomicAdd(dst + j + i, v[j + i]);
            }
          }
        }
      }
      if (a.db != nullptr && nt == 0 && half == 0) {
        float v[16];
        tmem_ld16(tmem + (uint32_t(q * 32) << 16) + 256u, v);
        if (valid) atomicAdd(a.db + n, v[0]);
      }
      tc_fence_before();
      mbar_arrive(&bar.tempty[0]);
      ++it;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ============================================================================================== gather / pack
// one thread per (row, group of 8 features); rows >= R and features outside every segment are written as zeros
__global__ void pack_rows_kernel(const PackArgs a) {
  const int groups = a.FCp * 8;
  const long long total = (long long)((a.R + 127) / 128) * 128 * groups;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int r = int(idx / groups), g = int(idx - (long long)r * groups);
  float y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) y[i] = 0.f;
  if (r < a.R) {
    long long b = r, d = 0;
    if (a.inds) {
      const long long f = a.inds[r];
      b = f / a.ft, d = f - b * a.ft;
    } else if (a.dinds) {
      d = a.dinds[r];
    }
    const float sc = (a.scale ? *a.scale : 1.f) * (a.scale_imm != 0.f ? a.scale_imm : 1.f);
    for (int s = 0; s < a.n_seg; ++s) {
      const PackSeg& sg = a.seg[s];
      if (g * 8 + 7 < sg.dst_col || g * 8 >= sg.dst_col + sg.width) continue;
      const float* src = nullptr;
      if (sg.src) src = sg.mode == 0 ? sg.src + (long long)r * sg.ld
                        : sg.mode == 1 ? sg.src + b * sg.ld
                                       : sg.src + b * a.chain_stride + d * a.chain_d;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = g * 8 + i - sg.dst_col;
        if (c >= 0 && c < sg.width) y[i] = sg.src ? src[c] * sc : (c == int(d) ? 1.f : 0.f);
      }
    }
  }
  uint8_t* img = a.out + (size_t(r >> 7) * a.FCp + (g >> 3)) * 2 * kImg;
  store_op8(img, uint32_t(r & 127), uint32_t((g & 7) * 8), y);
}

// fp32 weights -> B tiles [NT][KC][plane][NTILE x 128 B] (K-major SWIZZLE_128B, rows = output features of the GEMM)
__global__ void pack_weight_kernel(const float* __restrict__ W, long long s_row, long long s_col, int rows, int K, int NTILE,
                                   int NT, int KC, uint8_t* __restrict__ out) {
  const long long total = (long long)NT * KC * NTILE * 8;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = int(idx & 7);
  long long t = idx >> 3;
  const int j = int(t % NTILE);
  t /= NTILE;
  const int kc = int(t % KC), nt = int(t / KC);
  const int jg = nt * NTILE + j;
  float y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = kc * 64 + cg * 8 + i;
    y[i] = (jg < rows && c < K) ? W[jg * s_row + c * s_col] : 0.f;
  }
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(y[2 * i], y[2 * i + 1]);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(y[2 * i] - __low2float(h2), y[2 * i + 1] - __high2float(h2));
    h[i] = *reinterpret_cast<const uint32_t*>(&h2), l[i] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  const size_t plane = size_t(NTILE) * 128;
  uint8_t* tile = out + (size_t(nt) * KC + kc) * 2 * plane;
  const uint32_t off = uint32_t(j) * 128u + (uint32_t(cg ^ (j & 7)) << 4);
  *reinterpret_cast<uint4*>(tile + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(tile + plane + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void pack_weights_multi_kernel(const PackWJob* __restrict__ jobs) {
  const PackWJob jb = jobs[blockIdx.y];
  const long long total = (long long)jb.NT * jb.KC * jb.NTILE * 8;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cg = int(idx & 7);
    long long t = idx >> 3;
    const int j = int(t % jb.NTILE);
    t /= jb.NTILE;
    const int kc = int(t % jb.KC), nt = int(t / jb.KC);
    const int jg = nt * jb.NTILE + j;
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = kc * 64 + cg * 8 + i;
      y[i] = (jg < jb.rows && c < jb.K) ? jb.W[jg * jb.s_row + c * jb.s_col] : 0.f;
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h2 = __floats2bfloat162_rn(y[2 * i], y[2 * i + 1]);
      const __nv_bfloat162 l2 = __floats2bfloat162_rn(y[2 * i] - __low2float(h2), y[2 * i + 1] - __high2float(h2));
      h[i] = *reinterpret_cast<const uint32_t*>(&h2), l[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    const size_t plane = size_t(jb.NTILE) * 128;
    uint8_t* tile = jb.out + (size_t(nt) * jb.KC + kc) * 2 * plane;
    const uint32_t off = uint32_t(j) * 128u + (uint32_t(cg ^ (j & 7)) << 4);
    *reinterpret_cast<uint4*>(tile + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(tile + plane + off) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

__global__ void unpack_rows_kernel(const uint8_t* __restrict__ img, int FCp, int R, int F, float* __restrict__ out, int ld) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)R * F) return;
  const int r = int(idx / F), f = int(idx - (long long)r * F);
  const uint8_t* tile = img + (size_t(r >> 7) * FCp + (f >> 6)) * 2 * kImg;
  const uint32_t rl = uint32_t(r & 127), kk = uint32_t(f & 63);
  const uint32_t off = rl * 128u + (((kk >> 3) ^ (rl & 7u)) << 4) + (kk & 7u) * 2u;
  out[size_t(r) * ld + f] = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(tile + off)) +
                            __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(tile + kImg + off));
}

// ============================================================================================== LayerNorm around the GEMMs
// One warp per row (F <= 1024, multiple of 256): lane l owns features [8l, 8l+8) + 256 j.  Two-pass statistics in registers.
// reference: nn.LayerNorm(hidden_dim, eps=1e-6) inside TwoLayerPreActivationResNetLinear, dppo/model/common/mlp.py:137-152
constexpr int kLnMaxJ = 4;
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, int ld, int R, int F,
                                                     const float* __restrict__ g, const float* __restrict__ b, float eps,
                                                     int act, float* __restrict__ stats, uint8_t* __restrict__ out_op, int FCo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  const int RT128 = ((R + 127) / 128) * 128;
  if (r >= RT128) return;
  const int nj = F / 256;
  float v[kLnMaxJ][8];
  float mean = 0.f, rstd = 0.f;
  if (r < R) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kLnMaxJ; ++j)
      if (j < nj) {
        const float* p = x + size_t(r) * ld + j * 256 + lane * 8;
        const float4 a0 = *reinterpret_cast<const float4*>(p), a1 = *reinterpret_cast<const float4*>(p + 4);
        v[j][0] = a0.x, v[j][1] = a0.y, v[j][2] = a0.z, v[j][3] = a0.w, v[j][4] = a1.x, v[j][5] = a1.y, v[j][6] = a1.z, v[j][7] = a1.w;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[j][i];
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    mean = s / float(F);
    float s2 = 0.f;
#pragma unroll
    for (int j = 0; j < kLnMaxJ; ++j)
      if (j < nj) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d = v[j][i] - mean;
          s2 += d * d;
        }
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    rstd = rsqrtf(s2 / float(F) + eps);
    if (lane == 0) stats[2 * size_t(r)] = mean, stats[2 * size_t(r) + 1] = rstd;
  }
#pragma unroll
  for (int j = 0; j < kLnMaxJ; ++j)
    if (j < nj) {
      float y[8];
      const int f0 = j * 256 + lane * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        y[i] = r < R ? act_apply(act, (v[j][i] - mean) * rstd * __ldg(g + f0 + i) + __ldg(b + f0 + i)) : 0.f;
      uint8_t* img = out_op + (size_t(r >> 7) * FCo + (f0 >> 6)) * 2 * kImg;
      store_op8(img, uint32_t(r & 127), uint32_t(f0 & 63), y);
    }
}

// dz = gradient w.r.t. z = xhat * g + b (already multiplied by act'(z) in the dgrad epilogue);
// dx = rstd * (dz g - mean(dz g) - xhat mean(dz g xhat)) [+ res];  dg += sum_r dz xhat;  db += sum_r dz
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dz, int ld_dz, const float* __restrict__ x,
                                                     int ld_x, const float* __restrict__ stats, const float* __restrict__ g,
                                                     int R, int F, const float* __restrict__ res, int ld_res,
                                                     float* __restrict__ out_f32, int ld_out, uint8_t* __restrict__ out_op,
                                                     int FCo, float* __restrict__ dg, float* __restrict__ db) {
  __shared__ float s_dg[1024], s_db[1024];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nj = F / 256;
  const int RT128 = ((R + 127) / 128) * 128;
  for (int i = threadIdx.x; i < F; i += blockDim.x) s_dg[i] = 0.f, s_db[i] = 0.f;
  __syncthreads();
  float adg[kLnMaxJ][8], adb[kLnMaxJ][8];
#pragma unroll
  for (int j = 0; j < kLnMaxJ; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) adg[j][i] = 0.f, adb[j][i] = 0.f;
  for (int r = blockIdx.x * 8 + warp; r < RT128; r += gridDim.x * 8) {
    float dxh[kLnMaxJ][8], xh[kLnMaxJ][8];
    float rstd = 0.f, m1 = 0.f, m2 = 0.f;
    if (r < R) {
      const float mean = stats[2 * size_t(r)];
      rstd = stats[2 * size_t(r) + 1];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < kLnMaxJ; ++j)
        if (j < nj) {
          const int f0 = j * 256 + lane * 8;
          const float* pz = dz + size_t(r) * ld_dz + f0;
          const float* px = x + size_t(r) * ld_x + f0;
          const float4 z0 = *reinterpret_cast<const float4*>(pz), z1 = *reinterpret_cast<const float4*>(pz + 4);
          const float4 x0 = *reinterpret_cast<const float4*>(px), x1 = *reinterpret_cast<const float4*>(px + 4);
          const float zz[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
          const float xx[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float h = (xx[i] - mean) * rstd;
            const float d = zz[i] * __ldg(g + f0 + i);
            xh[j][i] = h, dxh[j][i] = d;
            s1 += d, s2 += d * h;
            adg[j][i] += zz[i] * h, adb[j][i] += zz[i];
          }
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      m1 = s1 / float(F), m2 = s2 / float(F);
    }
#pragma unroll
    for (int j = 0; j < kLnMaxJ; ++j)
      if (j < nj) {
        const int f0 = j * 256 + lane * 8;
        float y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = r < R ? rstd * (dxh[j][i] - m1 - xh[j][i] * m2) : 0.f;
        if (res && r < R) {
          const float* pr = res + size_t(r) * ld_res + f0;
          const float4 r0 = *reinterpret_cast<const float4*>(pr), r1 = *reinterpret_cast<const float4*>(pr + 4);
          y[0] += r0.x, y[1] += r0.y, y[2] += r0.z, y[3] += r0.w, y[4] += r1.x, y[5] += r1.y, y[6] += r1.z, y[7] += r1.w;
        }
        if (out_f32 && r < R) {
          float* po = out_f32 + size_t(r) * ld_out + f0;
          *reinterpret_cast<float4*>(po) = make_float4(y[0], y[1], y[2], y[3]);
          *reinterpret_cast<float4*>(po + 4) = make_float4(y[4], y[5], y[6], y[7]);
        }
        if (out_op) {
          uint8_t* img = out_op + (size_t(r >> 7) * FCo + (f0 >> 6)) * 2 * kImg;
          store_op8(img, uint32_t(r & 127), uint32_t(f0 & 63), y);
        }
      }
  }
#pragma unroll
  for (int j = 0; j < kLnMaxJ; ++j)
    if (j < nj) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        atomicAdd(&s_dg[j * 256 + lane * 8 + i], adg[j][i]);
        atomicAdd(&s_db[j * 256 + lane * 8 + i], adb[j][i]);
      }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < F; i += blockDim.x) {
    atomicAdd(dg + i, s_dg[i]);
    atomicAdd(db + i, s_db[i]);
  }
}

// ---------------------------------------------------------------------------------------------- LayerNorm, tiled fp32
// One CTA of 256 threads per 128-row tile of a tiled fp32 tensor [row tile][F / 8][128][8]: thread = (row, half).  Every
// global access of a warp is 1 KiB contiguous.  Pass 1: row statistics (shifted sums), pass 2: normalise + activation +
// bf16 hi/lo split, operand images assembled in shared memory and copied out coalesced (like the GEMM epilogue).
__global__ void __launch_bounds__(256) ln_fwd_tiled_kernel(const float* __restrict__ x, const float* __restrict__ sums, int R,
                                                           int F, const float* __restrict__ g, const float* __restrict__ b,
                                                           float eps, int act, float* __restrict__ stats,
                                                           uint8_t* __restrict__ out_op, int FCo) {
  extern __shared__ __align__(1024) uint8_t ln_smem[];
  uint8_t* stage = ln_smem;  // 32 KiB
  const int rt = blockIdx.x, t = threadIdx.x, r = t & 127, half = t >> 7;
  const int row = rt * 128 + r;
  const bool valid = row < R;
  const int G = F >> 3;
  const float* base = x + (size_t(rt) * G * 128 + r) * 8;  // group gi at + gi * 1024 floats
  float mean = 0.f, rstd = 0.f;
  if (valid) {
    const float t1 = sums[2 * size_t(row)] / float(F), t2 = sums[2 * size_t(row) + 1] / float(F);
    mean = t1;
    rstd = rsqrtf(fmaxf(t2 - t1 * t1, 0.f) + eps);
    if (half == 0 && blockIdx.y == 0) stats[2 * size_t(row)] = mean, stats[2 * size_t(row) + 1] = rstd;
  }
  const int ncc = F >> 6, per = (ncc + gridDim.y - 1) / gridDim.y;
  const int cc_end = min(ncc, int(blockIdx.y + 1) * per);
  for (int cc = blockIdx.y * per; cc < cc_end; ++cc) {
    const int n0 = cc * 64 + half * 32;
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) {
      float y[8];
      if (valid) {
        const float* p = base + size_t((n0 >> 3) + gq) * 1024;
        const float4 a0 = *reinterpret_cast<const float4*>(p), a1 = *reinterpret_cast<const float4*>(p + 4);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + n0 + gq * 8)), g1 = __ldg(reinterpret_cast<const float4*>(g + n0 + gq * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + n0 + gq * 8)), b1 = __ldg(reinterpret_cast<const float4*>(b + n0 + gq * 8 + 4));
        const float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = act_apply(act, (v[i] - mean) * rstd * gg[i] + bb[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = 0.f;
      }
      store_op8(stage, uint32_t(r), uint32_t(half * 32 + gq * 8), y);
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(out_op + (size_t(rt) * FCo + cc) * 2 * kImg);
    const uint4* src = reinterpret_cast<const uint4*>(stage);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[t + i * 256] = src[t + i * 256];
    __syncthreads();
  }
}

// dz (tiled) = gradient w.r.t. z = xhat g + b; dx = rstd (dz g - mean(dz g) - xhat mean(dz g xhat)) [+ res] -> fp32 (tiled,
// optional) and operand images; dg += sum_r dz xhat, db += sum_r dz (column sums over the tile through shared memory).
__global__ void __launch_bounds__(256) ln_bwd_tiled_kernel(const float* __restrict__ dz, const float* __restrict__ sums,
                                                           const float* __restrict__ x,
                                                           const float* __restrict__ stats, const float* __restrict__ g, int R,
                                                           int F, const float* __restrict__ res, float* __restrict__ out_f32,
                                                           uint8_t* __restrict__ out_op, int FCo, float* __restrict__ dg,
                                                           float* __restrict__ db) {
  extern __shared__ __align__(1024) uint8_t ln_smem[];
  uint8_t* stage = ln_smem;                                       // 32 KiB operand images of one chunk
  float* s_p = reinterpret_cast<float*>(ln_smem + 2 * kImg);      // [128][65] dz * xhat
  float* s_z = s_p + 128 * 65;                                    // [128][65] dz
  const int rt = blockIdx.x, t = threadIdx.x, r = t & 127, half = t >> 7;
  const int row = rt * 128 + r;
  const bool valid = row < R;
  const int G = F >> 3;
  const size_t tbase = (size_t(rt) * G * 128 + r) * 8;
  float mean = 0.f, rstd = 0.f, m1 = 0.f, m2 = 0.f;
  if (valid) {
    mean = stats[2 * size_t(row)], rstd = stats[2 * size_t(row) + 1];
    m1 = sums[2 * size_t(row)] / float(F), m2 = sums[2 * size_t(row) + 1] / float(F);
  }
  const int ncc = F >> 6, per = (ncc + gridDim.y - 1) / gridDim.y;
  const int cc_end = min(ncc, int(blockIdx.y + 1) * per);
  for (int cc = blockIdx.y * per; cc < cc_end; ++cc) {
    const int n0 = cc * 64 + half * 32;
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) {
      float y[8], pz8[8], zz8[8];
      const size_t off = tbase + size_t((n0 >> 3) + gq) * 1024;
      if (valid) {
        const float4 z0 = *reinterpret_cast<const float4*>(dz + off), z1 = *reinterpret_cast<const float4*>(dz + off + 4);
        const float4 x0 = *reinterpret_cast<const float4*>(x + off), x1 = *reinterpret_cast<const float4*>(x + off + 4);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + n0 + gq * 8)), g1 = __ldg(reinterpret_cast<const float4*>(g + n0 + gq * 8 + 4));
        const float zz[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
        const float xx[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float h = (xx[i] - mean) * rstd;
          y[i] = rstd * (zz[i] * gg[i] - m1 - h * m2);
          pz8[i] = zz[i] * h, zz8[i] = zz[i];
        }
        if (res) {
          const float4 r0 = *reinterpret_cast<const float4*>(res + off), r1 = *reinterpret_cast<const float4*>(res + off + 4);
          y[0] += r0.x, y[1] += r0.y, y[2] += r0.z, y[3] += r0.w, y[4] += r1.x, y[5] += r1.y, y[6] += r1.z, y[7] += r1.w;
        }
        if (out_f32) {
          *reinterpret_cast<float4*>(out_f32 + off) = make_float4(y[0], y[1], y[2], y[3]);
          *reinterpret_cast<float4*>(out_f32 + off + 4) = make_float4(y[4], y[5], y[6], y[7]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = 0.f, pz8[i] = 0.f, zz8[i] = 0.f;
      }
      store_op8(stage, uint32_t(r), uint32_t(half * 32 + gq * 8), y);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s_p[r * 65 + half * 32 + gq * 8 + i] = pz8[i];
        s_z[r * 65 + half * 32 + gq * 8 + i] = zz8[i];
      }
    }
    __syncthreads();
    {
      uint4* dst = reinterpret_cast<uint4*>(out_op + (size_t(rt) * FCo + cc) * 2 * kImg);
      const uint4* src = reinterpret_cast<const uint4*>(stage);
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[t + i * 256] = src[t + i * 256];
      if (t < 128) {  // column sums of this chunk: thread = (which array, column)
        const float* sp = (t < 64 ? s_p : s_z) + (t & 63);
        float acc = 0.f;
#pragma unroll 8
        for (int rr = 0; rr < 128; ++rr) acc += sp[rr * 65];
        atomicAdd((t < 64 ? dg : db) + cc * 64 + (t & 63), acc);
      }
    }
    __syncthreads();
  }
}

// ============================================================================================== launchers
static int g_ntile_cap = 256, g_max_stages = 0;
static bool g_fast_epilogue = true;
void set_row_gemm_fast_epilogue(bool on) { g_fast_epilogue = on; }
void set_row_gemm_tuning(int ntile_cap, int max_stages) {
  g_ntile_cap = (ntile_cap >= 64 && ntile_cap <= 256 && ntile_cap % 64 == 0) ? ntile_cap : 256;
  g_max_stages = max_stages;
}
int row_gemm_ntile_cap() { return g_ntile_cap; }
int row_gemm_ntile(int N) { return N >= g_ntile_cap ? g_ntile_cap : (N + 15) / 16 * 16; }

size_t packed_weight_bytes(int rows, int K, int NTILE) {
  const int NT = (rows + NTILE - 1) / NTILE, KC = (K + 63) / 64;
  return size_t(NT) * KC * 2 * NTILE * 128;
}

int launch_pack_weight(const float* W, int64_t s_row, int64_t s_col, int rows, int K, int NTILE, uint8_t* out, cudaStream_t st) {
  const int NT = (rows + NTILE - 1) / NTILE, KC = (K + 63) / 64;
  const long long total = (long long)NT * KC * NTILE * 8;
  pack_weight_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(W, s_row, s_col, rows, K, NTILE, NT, KC, out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPPO_OK : cuda_fail(e, "pack_weight_kernel launch");
}

int launch_pack_rows(const PackArgs& a, cudaStream_t st) {
  const long long total = (long long)((a.R + 127) / 128) * 128 * a.FCp * 8;
  if (total == 0) return DPPO_OK;
  pack_rows_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPPO_OK : cuda_fail(e, "pack_rows_kernel launch");
}

int launch_pack_weights(const PackWJob* d_jobs, int n_jobs, long long max_total, cudaStream_t st) {
  if (n_jobs <= 0) return DPPO_OK;
  long long bx = (max_total + 255) / 256;
  if (bx > 256) bx = 256;
  pack_weights_multi_kernel<<<dim3(unsigned(bx), unsigned(n_jobs)), 256, 0, st>>>(d_jobs);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPPO_OK : cuda_fail(e, "pack_weights_multi_kernel launch");
}

int launch_unpack_rows(const uint8_t* img, int FCp, int R, int F, float* out, int ld, cudaStream_t st) {
  const long long total = (long long)R * F;
  if (total == 0) return DPPO_OK;
  unpack_rows_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(img, FCp, R, F, out, ld);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPPO_OK : cuda_fail(e, "unpack_rows_kernel launch");
}

static uint64_t g_desc_mn = kDescSw128MnMajor;
void set_mn_desc_override(uint32_t lbo_bytes, uint32_t sbo_bytes) {
  g_desc_mn = (uint64_t(lbo_bytes >> 4) << 16) | (uint64_t(sbo_bytes >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}

constexpr size_t kUSmemBudget = 232448;

int launch_row_gemm(const RowGemmArgs& a0, int sm_count, cudaStream_t st) {
  RowGemmArgs a = a0;
  if (a.R <= 0) return DPPO_OK;
  if (a.NTP == 0) a.NTP = a.NTILE;
  if (a.NTILE < 16 || a.NTILE > 256 || a.NTILE % 16 || a.KC < 1 || a.NT < 1 || a.NTP % a.NTILE)
    return set_error("row gemm: bad tile geometry NTILE=%d (packed %d) KC=%d NT=%d", a.NTILE, a.NTP, a.KC, a.NT), DPPO_ERR_INVALID;
  if (a.out_op && ((a.op_col0 & 63) || (a.NT > 1 && (a.NTILE & 63)) || (a.op_col0 + a.NT * a.NTILE + 63) / 64 > a.FCo))
    return set_error("row gemm: operand output columns [%d, %d) must cover whole 64-feature chunks of %d", a.op_col0,
                     a.op_col0 + a.NT * a.NTILE, a.FCo), DPPO_ERR_INVALID;
  if ((a.pre && a.pre_mode && (a.ld_pre & 7)) || (a.res && a.res_mode && (a.ld_res & 7)) || (a.out_f32 && a.out_mode && (a.ld_out & 7)))
    return set_error("row gemm: tiled fp32 tensors need a feature count that is a multiple of 8"), DPPO_ERR_INVALID;
  // straight-line epilogue variants (see rows_epilogue_fast); anything else runs the generic epilogue
  int variant = kEpiGeneric;
  const bool fast_ok = a.NTILE % 64 == 0 && a.N % 64 == 0 && a.N == a.NT * a.NTILE && g_fast_epilogue &&
                       (!a.ln_stats || ((reinterpret_cast<uintptr_t>(a.ln_g) | reinterpret_cast<uintptr_t>(a.ln_b)) & 15) == 0) &&
                       (!a.pre || a.pre_mode == 1) && (!a.res || a.res_mode == 1) && (!a.out_f32 || a.out_mode == 1) &&
                       (!a.pre || a.ld_pre == a.N) && (!a.res || a.ld_res == a.N) && (!a.out_f32 || a.ld_out == a.N) &&
                       (!a.bias || (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0) && !(a.pre && a.mask_in) &&
                       (!(a.mask_in || a.mask_out) || a.mask_words * 32 == a.N);
  if (fast_ok) {
    const int pre = a.mask_in ? 1 : (a.pre ? (a.act_grad == kUActRelu ? 2 : (a.act_grad == kUActMish ? 3 : -1)) : 0);
    if (pre >= 0 && !(a.ln_stats && pre != 3))
      variant = epi_variant(a.bias != nullptr, pre, a.res != nullptr, a.out_f32 != nullptr, a.out_op ? a.act_out : 0, a.mask_out != nullptr,
                            a.out_op != nullptr, a.ln_stats != nullptr, a.stat_out != nullptr);
  }
  if (a.stat_out && (variant == kEpiGeneric || !((variant >> 10) & 1)))
    return set_error("row gemm: row statistics are only produced by the specialised epilogue variants"), DPPO_ERR_UNSUPPORTED;
  void (*kfn)(const RowGemmArgs) = nullptr;
  switch (variant) {
#define DPPO_EPI_CASE(...)                        \
  case epi_variant(__VA_ARGS__):                  \
    kfn = ugemm_rows_kernel<epi_variant(__VA_ARGS__)>; \
    break;
    //            bias   pre res    outf   act        mask-out
    DPPO_EPI_CASE(false, 0, false, true, kUActRelu, true)    // layer 0, ReLU actor
    DPPO_EPI_CASE(false, 0, false, true, kUActMish, false)   // layer 0, Mish actor
    DPPO_EPI_CASE(true, 0, false, true, kUActRelu, true)     // layer 0 with bias (critic), ReLU
    DPPO_EPI_CASE(true, 0, false, true, kUActMish, false)    // layer 0 with bias (critic), Mish; l1, Mish
    DPPO_EPI_CASE(true, 0, false, false, kUActRelu, true)    // l1, ReLU
    DPPO_EPI_CASE(true, 0, true, false, kUActNone, false)    // l2 of the last block
    DPPO_EPI_CASE(true, 0, true, true, kUActRelu, true)      // l2, ReLU
    DPPO_EPI_CASE(true, 0, true, true, kUActMish, false)     // l2, Mish
    DPPO_EPI_CASE(false, 0, false, true, kUActNone, false)   // dgrad of the output layer
    DPPO_EPI_CASE(false, 1, false, false, kUActNone, false)  // dgrad l2, ReLU mask
    DPPO_EPI_CASE(false, 3, false, false, kUActNone, false)  // dgrad l2, Mish
    DPPO_EPI_CASE(false, 1, true, true, kUActNone, false)    // dgrad l1, ReLU mask
    DPPO_EPI_CASE(false, 3, true, true, kUActNone, false)    // dgrad l1, Mish
    DPPO_EPI_CASE(false, 2, false, false, kUActNone, false)  // dgrad with relu'(fp32 pre)
    DPPO_EPI_CASE(false, 2, true, true, kUActNone, false)
    // LayerNorm nets: the GEMMs exchange fp32 tensors with the LayerNorm kernels, no operand output
    DPPO_EPI_CASE(false, 0, false, true, kUActNone, false, false, false, true)  // layer 0 -> h0 (+ row sums)
    DPPO_EPI_CASE(true, 0, false, true, kUActNone, false, false, false, true)   // l1 -> y
    DPPO_EPI_CASE(true, 0, true, true, kUActNone, false, false, false, true)    // l2 -> h + ...
    DPPO_EPI_CASE(false, 3, false, true, kUActNone, false, false, true, true)   // dgrad -> dz = (g W) mish'(LN(pre))
#undef DPPO_EPI_CASE
    default:
      variant = kEpiGeneric;
      kfn = ugemm_rows_kernel<kEpiGeneric>;
  }
  {
    static std::set<const void*> configured;
    if (!configured.count(reinterpret_cast<const void*>(kfn))) {
      cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kUSmemBudget));
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(ugemm_rows_kernel)");
      configured.insert(reinterpret_cast<const void*>(kfn));
    }
  }
  const size_t stage_bytes = 2 * size_t(kImg) + size_t(a.NTILE) * 256;
  int nstage = int((kUSmemBudget - 1024 - kUBarBytes - 2 * kImg) / stage_bytes);
  if (nstage > kUMaxStages) nstage = kUMaxStages;
  if (g_max_stages >= 1 && nstage > g_max_stages) nstage = g_max_stages;
  a.nstage = nstage;
  a.RT = (a.R + 127) / 128;
  const int n_tiles = a.RT * a.NT;
  const int grid = n_tiles < sm_count ? n_tiles : sm_count;
  const size_t smem = size_t(nstage) * stage_bytes + 2 * kImg + kUBarBytes + 1024;
  kfn<<<grid, kUThreads, smem, st>>>(a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPPO_OK : cuda_fail(e, "ugemm_rows_kernel launch");
}

int launch_wgrad(const WgradArgs& a0, int sm_count, cudaStream_t st) {
  WgradArgs a = a0;
  if (a.R <= 0) return DPPO_OK;
  a.desc_mn = g_desc_mn;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(ugemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kUSmemBudget));
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(ugemm_wgrad_kernel)");
    configured = true;
  }
  const int fcx_used = (a.K_in + 63) / 64;
  a.n_units = (a.R + 63) / 64;
  a.MT = (a.N_out + 127) / 128;
  a.NTn = (fcx_used + 3) / 4;
  a.NCH = (fcx_used + a.NTn - 1) / a.NTn;
  if (a.FCg < a.MT * 2 || a.FCx < fcx_used)
    return set_error("wgrad: operand images too narrow (FCg=%d MT=%d FCx=%d chunks=%d)", a.FCg, a.MT, a.FCx, fcx_used), DPPO_ERR_INVALID;
  const int per_split = a.MT * a.NTn;
  // one work item per CTA where possible: split the rows so that the items fill the SMs
  int S = sm_count / per_split;
  if (S < 1) S = 1;
  if (S > a.n_units) S = a.n_units;
  a.units_per_split = (a.n_units + S - 1) / S;
  a.S = (a.n_units + a.units_per_split - 1) / a.units_per_split;
  const size_t stage_bytes = 4 * size_t(kHalfImg) + size_t(a.NCH) * 2 * kHalfImg;
  int nstage = int((kUSmemBudget - 1024 - kUBarBytes - kHalfImg) / stage_bytes);
  if (nstage > kUMaxStages) nstage = kUMaxStages;
  a.nstage = nstage;
  const int n_items = per_split * a.S;
  const int grid = n_items < sm_count ? n_items : sm_count;
  const size_t smem = size_t(nstage) * stage_bytes + kHalfImg + kUBarBytes + 1024;
  ugemm_wgrad_kernel<<<grid, kUThreads, smem, st>>>(a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPPO_OK : cuda_fail(e, "ugemm_wgrad_kernel launch");
}

int launch_ln_fwd(const float* x, int ld, int R, int F, const float* g, const float* b, float eps, int act, float* stats,
                  uint8_t* out_op, int FCo, cudaStream_t st) {
  if (F % 256 || F > 1024 || (ld & 3)) return set_error("ln_fwd: F=%d ld=%d unsupported", F, ld), DPPO_ERR_INVALID;
  const int rows = (R + 127) / 128 * 128;
  ln_fwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, ld, R, F, g, b, eps, act, stats, out_op, FCo);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPPO_OK : cuda_fail(e, "ln_fwd_kernel launch");
}

// Column ranges per row tile of the LayerNorm kernels (grid.y): 4 for wide rows (more CTAs, more bytes in flight), and more
// when a minibatch shard has too few row tiles to fill the device (2200 rows = 18 row tiles: 72 CTAs at split 4 took as
// long as 8x the rows).
static int ln_sm_count() {
  static int n = 0;
  if (n <= 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 1;
  }
  return n;
}
static int ln_column_split(int R, int F) {
  const int ncc = F >> 6, rt = (R + 127) / 128;
  int split = ncc >= 8 ? 4 : (ncc >= 2 ? 2 : 1);
  while (rt * split < 2 * ln_sm_count() && split * 2 <= ncc && ncc % (split * 2) == 0) split *= 2;
  return split;
}

int launch_ln_fwd_tiled(const float* x, const float* sums, int R, int F, const float* g, const float* b, float eps, int act,
                        float* stats, uint8_t* out_op, int FCo, cudaStream_t st) {
  if (F % 64 || F < 64 || ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(b)) & 15))
    return set_error("ln_fwd_tiled: F=%d unsupported or unaligned parameters", F), DPPO_ERR_INVALID;
  if (R <= 0) return DPPO_OK;
  const int split = ln_column_split(R, F);
  ln_fwd_tiled_kernel<<<dim3((R + 127) / 128, split), 256, 2 * kImg, st>>>(x, sums, R, F, g, b, eps, act, stats, out_op, FCo);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPPO_OK : cuda_fail(e, "ln_fwd_tiled_kernel launch");
}

int launch_ln_bwd_tiled(const float* dz, const float* sums, const float* x, const float* stats, const float* g, int R, int F,
                        const float* res, float* out_f32, uint8_t* out_op, int FCo, float* dg, float* db, cudaStream_t st) {
  if (F % 64 || F < 64 || (reinterpret_cast<uintptr_t>(g) & 15))
    return set_error("ln_bwd_tiled: F=%d unsupported or unaligned parameters", F), DPPO_ERR_INVALID;
  if (R <= 0) return DPPO_OK;
  const size_t smem = 2 * kImg + 2 * 128 * 65 * sizeof(float);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(ln_bwd_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(ln_bwd_tiled_kernel)");
    configured = true;
  }
  const int split = ln_column_split(R, F);
  ln_bwd_tiled_kernel<<<dim3((R + 127) / 128, split), 256, smem, st>>>(dz, sums, x, stats, g, R, F, res, out_f32, out_op, FCo, dg, db);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPPO_OK : cuda_fail(e, "ln_bwd_tiled_kernel launch");
}

int launch_ln_bwd(const float* dz, int ld_dz, const float* x, int ld_x, const float* stats, const float* g, int R, int F,
                  const float* res, int ld_res, float* out_f32, int ld_out, uint8_t* out_op, int FCo, float* dg, float* db,
                  int sm_count, cudaStream_t st) {
  if (F % 256 || F > 1024 || (ld_dz & 3) || (ld_x & 3) || (res && (ld_res & 3)) || (out_f32 && (ld_out & 3)))
    return set_error("ln_bwd: F=%d unsupported", F), DPPO_ERR_INVALID;
  const int rows = (R + 127) / 128 * 128;
  int blocks = (rows + 7) / 8;
  if (blocks > sm_count * 2) blocks = sm_count * 2;
  ln_bwd_kernel<<<blocks, 256, 0, st>>>(dz, ld_dz, x, ld_x, stats, g, R, F, res, ld_res, out_f32, ld_out, out_op, FCo, dg, db);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPPO_OK : cuda_fail(e, "ln_bwd_kernel launch");
}

}  // namespace dppo
