// K2: the MLP contractions of prop_net / nerf_net (model.py:43-53, :131-158) on 5th-generation tensor
// cores.  bf16 operands are staged by TMA into 128-byte-swizzled shared memory, tcgen05.mma
// (M=128, N<=256, K=16 per instruction, issued by one thread) accumulates fp32 in tensor memory, and
// four epilogue warps read the accumulator back with tcgen05.ld and apply the fused epilogue.
//
//   linear_kernel<BN, EPI>   persistent, warp-specialised (TMA warp / MMA warp / 4 epilogue warps),
//                            4-6 stage smem ring, double-buffered TMEM accumulator:
//        EPI_FWD    Y = act(X W^T + b)            (bias + ReLU/Sigmoid, bf16 and/or fp32 head output)
//        EPI_DGRAD  dX = (dY Wt^T) .* act'(Yprev) (activation derivative from the saved output)
//   wgrad_kernel<BN>         dW = dY^T X with both operands MN-major (the reduction dimension is the
//                            slow one in memory), split over rays, fp32 vector atomics; the bias
//                            gradient rides along as one extra N=16 MMA against a tile of ones.
#include "common.cuh"
#include "tcgen05.cuh"

namespace mip360 {

using namespace ptx;

constexpr int BM = 128;   // UMMA M
constexpr int BK = 64;    // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue

enum { EPI_FWD = 0, EPI_DGRAD = 1, EPI_FWD_HEAD = 2 /* forward with the head folded into the epilogue */ };
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2, ACT_SIGMOID_FAST = 3 /* internal: bf16-only outputs */ };

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x BN tile;
// each CTA stages its own 128 rows of A and HALF of the B tile, so a stage is 32 KB instead of 48 KB (6 stages
// instead of 4 in the same shared memory) and B is fetched from L2 once per pair.
// OCC = 2 ("short-K" configuration): BN = 128 tiles, 2 smem stages, two CTAs co-resident per SM.  Layers with a
// short reduction (K <= 256: the 58-wide input layers, the 256-wide proposal net, the head dgrad) are bound by
// the epilogue, not the MMAs; two resident CTAs double the number of epilogue warps per SM.
template <int BN, int CG = 1, int OCC = 1>
struct LinearCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / CG) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = OCC == 2 ? 2 : (BN >= 256) ? (CG == 1 ? 4 : 5) : (BN >= 128 ? 5 : 6);
  // epilogue boxes per warp: 3 when shared memory allows (the saved-activation box of dgrad is then
  // prefetched a whole chunk ahead), else 2
  static constexpr int NBOX = ((BN >= 256 && CG == 1) || OCC == 2) ? 2 : 3;
  static constexpr int TMEM_COLS = 2 * BN;  // double-buffered accumulator (power of two >= 32)
  // epilogue staging: per epilogue warp one 4 KB [32 rows x 64 cols] bf16 box for the TMA store, and a second
  // region of the same size that holds the saved-activation box (dgrad) or, shared by all warps, the bias
  // tile of the two accumulator stages (fwd)
  static constexpr int EPI_BOX_BYTES = 32 * 64 * 2;
  static constexpr int EPI_BYTES = NBOX * 4 * EPI_BOX_BYTES + 1024 /* bias tile: up to 256 floats */;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024 /*align slack*/;  static_assert(SMEM_BYTES <= 232448 / OCC, "exceeds the shared memory available per CTA at this occupancy");
  // fused head (EPI_FWD_HEAD): the BN x 4 fp32 head weights of the CTA's column tile, after the barriers, when they fit
  static constexpr int HEAD_BYTES = BN * 16;
  static constexpr bool HEAD_SMEM = SMEM_BYTES + HEAD_BYTES <= 232448 / OCC;
  static_assert(TMEM_COLS * OCC <= 512, "tensor memory oversubscribed");
};

struct LinearParams {
  const float* bias;       // [N] (EPI_FWD) or null
  uint16_t* out_bf16;      // [M, N] or null (then only out_f32 is written)
  float* out_f32;          // [M, n_valid] or null
  int M, N, K, act, n_valid;
  int packed_epi;          // packed fp32x2 / bf16x2 epilogue arithmetic (same results, about half the instructions)
  // Head fused into this (the last trunk) layer, model.py:150-158,180-181: head_out[row, 0..3] += sum over this tile's
  // columns of act(...)[row, col] * head_w4[col, 0..3] (fp32 [N, 4], column-interleaved head weights, bias excluded).
  // The 64-column head GEMM and its 2 GB read of the trunk output disappear; with out_bf16 == null (inference) the
  // trunk output is never written either.
  const float* head_w4;
  float* head_out;
};

// 32 accumulator columns of one row: + bias (broadcast reads from shared memory), activation
template <int ACT>
__device__ __forceinline__ void fwd_math(const uint32_t (&v)[32], uint32_t bias_addr, float (&x)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint4 bb;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(bb.x), "=r"(bb.y), "=r"(bb.z), "=r"(bb.w) : "r"(bias_addr + 16u * i));
    const float z[4] = {__uint_as_float(v[4 * i + 0]) + __uint_as_float(bb.x), __uint_as_float(v[4 * i + 1]) + __uint_as_float(bb.y),
                        __uint_as_float(v[4 * i + 2]) + __uint_as_float(bb.z), __uint_as_float(v[4 * i + 3]) + __uint_as_float(bb.w)};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float r;
      if (ACT == ACT_RELU) {
        r = fmaxf(z[e], 0.f);
      } else if (ACT == ACT_SIGMOID) {
        r = __fdividef(1.f, 1.f + __expf(-z[e]));
      } else if (ACT == ACT_SIGMOID_FAST) {
        // sigmoid(z) = 0.5 tanh(z/2) + 0.5: one MUFU op; |error| ~ 2.4e-4, an eighth of the bf16 rounding of the output
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z[e]));
        r = fmaf(0.5f, t, 0.5f);
      } else {
        r = z[e];
      }
      x[4 * i + e] = r;
    }
  }
}
// dX = acc .* act'(y) with y the saved bf16 output of the layer (16 packed pairs)
template <int ACT>
__device__ __forceinline__ void bwd_math(const uint32_t (&v)[32], const uint32_t* yw, float (&x)[32]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float y_lo = __uint_as_float(yw[i] << 16), y_hi = __uint_as_float(yw[i] & 0xffff0000u);
    float d_lo = 1.f, d_hi = 1.f;
    if (ACT == ACT_RELU) {
      d_lo = y_lo > 0.f ? 1.f : 0.f;
      d_hi = y_hi > 0.f ? 1.f : 0.f;
    } else if (ACT == ACT_SIGMOID) {
      d_lo = y_lo * (1.f - y_lo);
      d_hi = y_hi * (1.f - y_hi);
    }
    x[2 * i] = __uint_as_float(v[2 * i]) * d_lo;
    x[2 * i + 1] = __uint_as_float(v[2 * i + 1]) * d_hi;
  }
}

// The two ReLU epilogues in packed arithmetic.  The board is power-capped, so epilogue instructions are paid for in
// clock rate: both forms below produce exactly the bf16 values of the scalar forms with about half the instructions.
//   forward:  bias add as fp32x2, round to bf16x2, ReLU as one bf16x2 max   (round(max(z,0)) == max(round(z),0))
__device__ __forceinline__ void fwd_relu_packed(const uint32_t (&v)[32], uint32_t bias_addr, uint32_t* pk) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint4 bb;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(bb.x), "=r"(bb.y), "=r"(bb.z), "=r"(bb.w) : "r"(bias_addr + 16u * i));
    uint32_t s0, s1, s2, s3;
    asm("{\n"
        ".reg .b64 a, b, c;\n"
        "mov.b64 a, {%4, %5};\n"
        "mov.b64 b, {%8, %9};\n"
        "add.rn.f32x2 c, a, b;\n"
        "mov.b64 {%0, %1}, c;\n"
        "mov.b64 a, {%6, %7};\n"
        "mov.b64 b, {%10, %11};\n"
        "add.rn.f32x2 c, a, b;\n"
        "mov.b64 {%2, %3}, c;\n"
        "}\n"
        : "=r"(s0), "=r"(s1), "=r"(s2), "=r"(s3)
        : "r"(v[4 * i]), "r"(v[4 * i + 1]), "r"(v[4 * i + 2]), "r"(v[4 * i + 3]), "r"(bb.x), "r"(bb.y), "r"(bb.z), "r"(bb.w));
    uint32_t p0 = pack_bf16x2(__uint_as_float(s0), __uint_as_float(s1));
    uint32_t p1 = pack_bf16x2(__uint_as_float(s2), __uint_as_float(s3));
    asm("max.bf16x2 %0, %0, %1;" : "+r"(p0) : "r"(0u));
    asm("max.bf16x2 %0, %0, %1;" : "+r"(p1) : "r"(0u));
    pk[2 * i] = p0;
    pk[2 * i + 1] = p1;
  }
}
//   trunk Sigmoid (bf16-only output): the scalar tanh form with its adds / multiplies / FMAs issued as fp32x2
__device__ __forceinline__ void fwd_sigmoid_fast_packed(const uint32_t (&v)[32], uint32_t bias_addr, uint32_t* pk) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint4 bb;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(bb.x), "=r"(bb.y), "=r"(bb.z), "=r"(bb.w) : "r"(bias_addr + 16u * i));
    uint32_t r0, r1, r2, r3;
    asm("{\n"
        ".reg .b64 a, b, c, hh;\n"
        ".reg .f32 t0, t1;\n"
        "mov.b64 hh, {%12, %12};\n"
        "mov.b64 a, {%4, %5};\n"
        "mov.b64 b, {%8, %9};\n"
        "add.rn.f32x2 c, a, b;\n"
        "mul.rn.f32x2 c, c, hh;\n"
        "mov.b64 {t0, t1}, c;\n"
        "tanh.approx.f32 t0, t0;\n"
        "tanh.approx.f32 t1, t1;\n"
        "mov.b64 c, {t0, t1};\n"
        "fma.rn.f32x2 c, c, hh, hh;\n"
        "mov.b64 {%0, %1}, c;\n"
        "mov.b64 a, {%6, %7};\n"
        "mov.b64 b, {%10, %11};\n"
        "add.rn.f32x2 c, a, b;\n"
        "mul.rn.f32x2 c, c, hh;\n"
        "mov.b64 {t0, t1}, c;\n"
        "tanh.approx.f32 t0, t0;\n"
        "tanh.approx.f32 t1, t1;\n"
        "mov.b64 c, {t0, t1};\n"
        "fma.rn.f32x2 c, c, hh, hh;\n"
        "mov.b64 {%2, %3}, c;\n"
        "}\n"
        : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
        : "r"(v[4 * i]), "r"(v[4 * i + 1]), "r"(v[4 * i + 2]), "r"(v[4 * i + 3]), "r"(bb.x), "r"(bb.y), "r"(bb.z), "r"(bb.w),
          "r"(0x3f000000u));
    pk[2 * i] = pack_bf16x2(__uint_as_float(r0), __uint_as_float(r1));
    pk[2 * i + 1] = pack_bf16x2(__uint_as_float(r2), __uint_as_float(r3));
  }
}
//   trunk Sigmoid + fused head: the same packed Sigmoid, then per column two fp32x2 FMAs of (y, y) against the column's four
//   head weights (one 16-byte shared-memory broadcast load) into the row's two accumulator pairs
__device__ __forceinline__ void fwd_sigmoid_fast_head_packed(const uint32_t (&v)[32], uint32_t bias_addr, uint32_t w4_addr,
                                                             uint32_t* pk, uint64_t& acc01, uint64_t& acc23) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint4 bb;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(bb.x), "=r"(bb.y), "=r"(bb.z), "=r"(bb.w) : "r"(bias_addr + 16u * i));
    uint32_t r0, r1, r2, r3;
    asm("{\n"
        ".reg .b64 a, b, c, hh;\n"
        ".reg .f32 t0, t1;\n"
        "mov.b64 hh, {%12, %12};\n"
        "mov.b64 a, {%4, %5};\n"
        "mov.b64 b, {%8, %9};\n"
        "add.rn.f32x2 c, a, b;\n"
        "mul.rn.f32x2 c, c, hh;\n"
        "mov.b64 {t0, t1}, c;\n"
        "tanh.approx.f32 t0, t0;\n"
        "tanh.approx.f32 t1, t1;\n"
        "mov.b64 c, {t0, t1};\n"
        "fma.rn.f32x2 c, c, hh, hh;\n"
        "mov.b64 {%0, %1}, c;\n"
        "mov.b64 a, {%6, %7};\n"
        "mov.b64 b, {%10, %11};\n"
        "add.rn.f32x2 c, a, b;\n"
        "mul.rn.f32x2 c, c, hh;\n"
        "mov.b64 {t0, t1}, c;\n"
        "tanh.approx.f32 t0, t0;\n"
        "tanh.approx.f32 t1, t1;\n"
        "mov.b64 c, {t0, t1};\n"
        "fma.rn.f32x2 c, c, hh, hh;\n"
        "mov.b64 {%2, %3}, c;\n"
        "}\n"
        : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
        : "r"(v[4 * i]), "r"(v[4 * i + 1]), "r"(v[4 * i + 2]), "r"(v[4 * i + 3]), "r"(bb.x), "r"(bb.y), "r"(bb.z), "r"(bb.w),
          "r"(0x3f000000u));
    pk[2 * i] = pack_bf16x2(__uint_as_float(r0), __uint_as_float(r1));
    pk[2 * i + 1] = pack_bf16x2(__uint_as_float(r2), __uint_as_float(r3));
    const uint32_t rr[4] = {r0, r1, r2, r3};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      uint4 w;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "r"(w4_addr + 16u * (4 * i + e)));
      asm("{\n"
          ".reg .b64 y, wa, wb;\n"
          "mov.b64 y, {%2, %2};\n"
          "mov.b64 wa, {%3, %4};\n"
          "mov.b64 wb, {%5, %6};\n"
          "fma.rn.f32x2 %0, y, wa, %0;\n"
          "fma.rn.f32x2 %1, y, wb, %1;\n"
          "}\n"
          : "+l"(acc01), "+l"(acc23)
          : "r"(rr[e]), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w));
    }
  }
}

//   dgrad:  round the accumulator to bf16x2, AND with the per-half mask (saved output > 0)
__device__ __forceinline__ void bwd_relu_packed(const uint32_t (&v)[32], const uint32_t* yw, uint32_t* pk) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const __nv_bfloat162 y = *reinterpret_cast<const __nv_bfloat162*>(&yw[i]);
    const unsigned m = __hgt2_mask(y, __float2bfloat162_rn(0.f));  // 0xFFFF per half where y > 0
    pk[i] = pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])) & m;
  }
}

//   dgrad through a Sigmoid:  acc * y (1 - y) with the subtraction and both multiplications as fp32x2
__device__ __forceinline__ void bwd_sigmoid_packed(const uint32_t (&v)[32], const uint32_t* yw, uint32_t* pk) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    uint32_t r0, r1;
    asm("{\n"
        ".reg .b64 y, one, d, a;\n"
        "mov.b64 y, {%4, %5};\n"
        "mov.b64 one, {%6, %6};\n"
        "sub.rn.f32x2 d, one, y;\n"
        "mul.rn.f32x2 d, y, d;\n"
        "mov.b64 a, {%2, %3};\n"
        "mul.rn.f32x2 a, a, d;\n"
        "mov.b64 {%0, %1}, a;\n"
        "}\n"
        : "=r"(r0), "=r"(r1)
        : "r"(v[2 * i]), "r"(v[2 * i + 1]), "r"(yw[i] << 16), "r"(yw[i] & 0xffff0000u), "r"(0x3f800000u));
    pk[i] = pack_bf16x2(__uint_as_float(r0), __uint_as_float(r1));
  }
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int BN, int EPI, int CG, int OCC>
__global__ void __launch_bounds__(GEMM_THREADS, OCC)
linear_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
              const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_y,
              const LinearParams p) {
  using Cfg = LinearCfg<BN, CG, OCC>;
  constexpr bool IS_FWD = EPI == EPI_FWD || EPI == EPI_FWD_HEAD;
  constexpr bool HEAD = EPI == EPI_FWD_HEAD;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;  // 1024-aligned
  const uint32_t bar_base = epi_base + Cfg::EPI_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + a); };
  auto y_bar = [&](int i) { return bar_base + 8u * (2 * Cfg::STAGES + 4 + i); };   // 2*6+4+12 = 28 barriers max
  const uint32_t tmem_slot = bar_base + 8u * 28;
  const uint32_t head_smem = bar_base + Cfg::BAR_BYTES;  // EPI_FWD_HEAD with Cfg::HEAD_SMEM: BN x 4 head weights

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // scheduling unit = CTA (CG 1) or CTA pair (CG 2); `rank` = position inside the pair
  const int rank = (CG == 2) ? (int)cluster_ctarank() : 0;
  const int unit = blockIdx.x / CG, nunits = gridDim.x / CG;
  const int tiles_m = (p.M + BM * CG - 1) / (BM * CG), tiles_n = p.N / BN;
  const int total_tiles = tiles_m * tiles_n;
  const int kblocks = p.K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4 * CG);  // CG 2: the leader's barrier collects both CTAs' epilogue warps
    }
    for (int i = 0; i < 4 * Cfg::NBOX; ++i) mbar_init(y_bar(i), 1);
    fence_barrier_init();
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_out);
    if (EPI == EPI_DGRAD) prefetch_tmap(&tmap_y);
  }
  __syncthreads();  // orders the barrier initialisation before tcgen05.alloc's write of the TMEM address (once per kernel)
  if (warp == 1) {
    tmem_alloc<CG>(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish<CG>();
  }
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();  // barrier inits visible to the peer before any remote arrive
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (one per CTA; CG 2: both CTAs signal the LEADER's full barrier) =====
      uint32_t stage = 0, phase = 0;
      for (int tile = unit; tile < total_tiles; tile += nunits) {
        const int m0 = (tile / tiles_n) * (BM * CG) + rank * BM, n0 = (tile % tiles_n) * BN + rank * (BN / CG) * (CG - 1);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          if (CG == 1) {
            mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
            tma_load_2d(sa, &tmap_a, full_bar(stage), kb * BK, m0);
            tma_load_2d(sa + Cfg::A_BYTES, &tmap_b, full_bar(stage), kb * BK, n0);
          } else {
            if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
            tma_load_2d_pair(sa, &tmap_a, full_bar(stage), kb * BK, m0);
            tma_load_2d_pair(sa + Cfg::A_BYTES, &tmap_b, full_bar(stage), kb * BK, n0);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===== MMA issuer (CG 2: the leader CTA issues M = 256 MMAs for the pair) =====
      constexpr uint32_t idesc = make_idesc_bf16(BN, 0, 0, BM * CG);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = unit; tile < total_tiles; tile += nunits) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t adesc = make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(sa + Cfg::A_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 bf16 = 32 B along K inside the swizzle row: +2 in the (addr >> 4) field
            umma_bf16<CG>(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit<CG>(empty_bar(stage));  // frees the smem slot (in both CTAs) when these MMAs have read it
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit<CG>(tfull_bar(acc));  // accumulator complete (signalled to both CTAs' epilogues)
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===== epilogue warps: TMEM lanes [32q, 32q+32) belong to the warp with (warp % 4) == q =====
    // accumulator -> registers -> (bias, activation | activation derivative) -> bf16 -> swizzled smem box
    // [32 rows x 64 cols] -> TMA store.  Each warp owns two boxes and alternates between them, so a store is
    // still reading one box while the next chunk is being produced; only __syncwarp is needed.
    // dgrad: the box first receives the saved-activation tile by TMA (prefetched one chunk ahead), each lane
    // reads its own row, and the result is written back in place before the box is stored.
    const int q = warp & 3;
    constexpr int NBOX = Cfg::NBOX;
    const uint32_t box0 = epi_base + q * NBOX * Cfg::EPI_BOX_BYTES;
    const uint32_t bias_smem = epi_base + 4 * NBOX * Cfg::EPI_BOX_BYTES;  // fwd: BN floats, shared by the 4 warps
    const uint32_t row_off = (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);
    const bool tma_out = p.out_bf16 != nullptr;
    uint32_t acc = 0, acc_phase = 0;
    uint32_t bi = 0, bphase = 0;  // current box of this warp and the parity of its saved-activation barrier
    int bias_n0 = -1;
    // issue the TMA load of the saved-activation box for column chunk (nt, nj) into box `nb`
    auto load_y = [&](int nt, int nj, uint32_t nb) {
      if (nt < total_tiles) {
        const int nm0 = (nt / tiles_n) * (BM * CG) + rank * BM, nn0 = (nt % tiles_n) * BN;
        mbar_arrive_expect_tx(y_bar(NBOX * q + nb), Cfg::EPI_BOX_BYTES);
        tma_load_2d(box0 + nb * Cfg::EPI_BOX_BYTES, &tmap_y, y_bar(NBOX * q + nb), nn0 + nj * 64, nm0 + q * 32);
      }
    };
    if (EPI == EPI_DGRAD && lane == 0) load_y(unit, 0, 0);
    for (int tile = unit; tile < total_tiles; tile += nunits) {
      const int m0 = (tile / tiles_n) * (BM * CG) + rank * BM, n0 = (tile % tiles_n) * BN;
      const int row = m0 + q * 32 + lane;
      if (IS_FWD && n0 != bias_n0) {
        // bias tile into shared memory (read back as broadcasts).  With gridDim.x a multiple of the number
        // of column tiles every CTA keeps the same column block, so this runs once per kernel.
        if (bias_n0 >= 0) epi_bar_sync();  // everyone has finished reading the previous tile's bias
        const int t = threadIdx.x - 64;    // 0..127
        for (int i = t; i < BN; i += 128)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_smem + i * 4u), "f"(__ldg(p.bias + n0 + i)) : "memory");
        if (HEAD && Cfg::HEAD_SMEM) {
          for (int i = t; i < BN; i += 128) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(p.head_w4) + n0 + i);
            st_shared_v4(head_smem + i * 16u, __float_as_uint(w.x), __float_as_uint(w.y), __float_as_uint(w.z), __float_as_uint(w.w));
          }
        }
        epi_bar_sync();
        bias_n0 = n0;
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      float hacc[4] = {0.f, 0.f, 0.f, 0.f};  // fused head: this row's partial dot products over the tile's columns
      uint64_t hacc01 = 0ull, hacc23 = 0ull;  // ... the same as fp32 pairs (packed path)
      constexpr bool fused_head = HEAD;
      if (IS_FWD && !tma_out && !fused_head) {
        // fp32-only output = an MLP head: n_valid <= 8 real columns in the first column tile; nothing else of the
        // accumulator is read, no activation is evaluated on padding
        if (n0 == 0) {
          uint32_t v8[8];
          tmem_ld_32x8(t_row, v8);
          tmem_ld_wait();
          if (row < p.M) {
            float* o = p.out_f32 + (size_t)row * p.n_valid;
            float r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float bv;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(bv) : "r"(bias_smem + 4u * i));
              const float z = __uint_as_float(v8[i]) + bv;
              r[i] = p.act == ACT_SIGMOID ? __fdividef(1.f, 1.f + __expf(-z)) : (p.act == ACT_RELU ? fmaxf(z, 0.f) : z);
            }
            if (p.n_valid == 4) {
              *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (i < p.n_valid) o[i] = r[i];
            }
          }
        }
      } else
#pragma unroll 1
      for (int jj = 0; jj < BN / 64; ++jj) {
        const uint32_t box = box0 + bi * Cfg::EPI_BOX_BYTES;
        const uint32_t nbi = (bi + 1 == NBOX) ? 0u : bi + 1;
        int nt = tile, nj = jj + 1;  // the chunk after this one
        if (nj == BN / 64) { nj = 0; nt = tile + nunits; }
        uint32_t packed[32];  // 64 bf16 of this lane's row
        uint4 yv[8];
        if (EPI == EPI_DGRAD) {
          if (NBOX >= 3 && lane == 0) {
            // three boxes: the next box was last read by the store issued two chunks ago -> prefetch now, a
            // whole chunk ahead of its use
            tma_store_wait_read<1>();
            load_y(nt, nj, nbi);
          }
          mbar_wait(y_bar(NBOX * q + bi), bphase);
#pragma unroll
          for (int c = 0; c < 8; ++c) yv[c] = ld_shared_v4(box + row_off + ((c ^ sw) << 4));
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld_32x32(t_row + jj * 64 + h * 32, v);
          tmem_ld_wait();
          if (fused_head && Cfg::HEAD_SMEM && p.act == ACT_SIGMOID_FAST) {
            fwd_sigmoid_fast_head_packed(v, bias_smem + (jj * 64 + h * 32) * 4u, head_smem + (jj * 64 + h * 32) * 16u,
                                         &packed[16 * h], hacc01, hacc23);
            continue;
          }
          if (fused_head) {
            // activation in fp32, then 4 FMAs per column against the head weights (uniform 16-byte loads: every lane
            // of the warp reads the same address, one L1 transaction per load)
            float x[32];
            const uint32_t bsm = bias_smem + (jj * 64 + h * 32) * 4u;
            if (p.act == ACT_RELU) fwd_math<ACT_RELU>(v, bsm, x);
            else if (p.act == ACT_SIGMOID || p.act == ACT_SIGMOID_FAST) fwd_math<ACT_SIGMOID_FAST>(v, bsm, x);
            else fwd_math<ACT_NONE>(v, bsm, x);
            const float4* w4 = reinterpret_cast<const float4*>(p.head_w4) + n0 + jj * 64 + h * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float4 w = __ldg(w4 + i);
              hacc[0] = fmaf(x[i], w.x, hacc[0]);
              hacc[1] = fmaf(x[i], w.y, hacc[1]);
              hacc[2] = fmaf(x[i], w.z, hacc[2]);
              hacc[3] = fmaf(x[i], w.w, hacc[3]);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) packed[16 * h + i] = pack_bf16x2(x[2 * i], x[2 * i + 1]);
            continue;
          }
          if (IS_FWD && p.act == ACT_RELU && !p.out_f32 && p.packed_epi) {
            fwd_relu_packed(v, bias_smem + (jj * 64 + h * 32) * 4u, &packed[16 * h]);
            continue;
          }
          if (IS_FWD && p.act == ACT_SIGMOID_FAST && p.packed_epi) {  // (bf16-only output by construction)
            fwd_sigmoid_fast_packed(v, bias_smem + (jj * 64 + h * 32) * 4u, &packed[16 * h]);
            continue;
          }
          if (EPI == EPI_DGRAD && p.act == ACT_RELU && p.packed_epi) {
            bwd_relu_packed(v, reinterpret_cast<const uint32_t*>(&yv[4 * h]), &packed[16 * h]);
            continue;
          }
          if (EPI == EPI_DGRAD && p.act == ACT_SIGMOID && p.packed_epi) {
            bwd_sigmoid_packed(v, reinterpret_cast<const uint32_t*>(&yv[4 * h]), &packed[16 * h]);
            continue;
          }
          float x[32];
          if (IS_FWD) {
            const uint32_t bsm = bias_smem + (jj * 64 + h * 32) * 4u;
            if (p.act == ACT_RELU) fwd_math<ACT_RELU>(v, bsm, x);
            else if (p.act == ACT_SIGMOID) fwd_math<ACT_SIGMOID>(v, bsm, x);
            else if (p.act == ACT_SIGMOID_FAST) fwd_math<ACT_SIGMOID_FAST>(v, bsm, x);
            else fwd_math<ACT_NONE>(v, bsm, x);
            if (p.out_f32 && jj == 0 && h == 0 && n0 == 0 && row < p.M) {
              float* o = p.out_f32 + (size_t)row * p.n_valid;
              if (p.n_valid == 4) {
                *reinterpret_cast<float4*>(o) = make_float4(x[0], x[1], x[2], x[3]);
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  if (i < p.n_valid) o[i] = x[i];
              }
            }
          } else {
            const uint32_t* yw = reinterpret_cast<const uint32_t*>(&yv[4 * h]);
            if (p.act == ACT_RELU) bwd_math<ACT_RELU>(v, yw, x);
            else if (p.act == ACT_SIGMOID) bwd_math<ACT_SIGMOID>(v, yw, x);
            else bwd_math<ACT_NONE>(v, yw, x);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) packed[16 * h + i] = pack_bf16x2(x[2 * i], x[2 * i + 1]);
        }
        if (EPI == EPI_DGRAD) {
          if (NBOX == 2 && lane == 0) {
            // two boxes: the store that used the other box (previous chunk) must have finished reading it before
            // the next saved-activation box can be fetched into it
            tma_store_wait_read<0>();
            load_y(nt, nj, nbi);
          }
          __syncwarp();
        } else if (tma_out) {
          if (lane == 0) tma_store_wait_read<NBOX - 1>();  // the store that last used this box has read it
          __syncwarp();
        }
        if (tma_out) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            st_shared_v4(box + row_off + ((c ^ sw) << 4), packed[4 * c], packed[4 * c + 1], packed[4 * c + 2],
                         packed[4 * c + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_out, box, n0 + jj * 64, m0 + q * 32);
            tma_store_commit();
          }
        }
        bi = nbi;
        if (bi == 0) bphase ^= 1u;
      }
      if (fused_head && row < p.M) {
        hacc[0] += __uint_as_float((uint32_t)hacc01);
        hacc[1] += __uint_as_float((uint32_t)(hacc01 >> 32));
        hacc[2] += __uint_as_float((uint32_t)hacc23);
        hacc[3] += __uint_as_float((uint32_t)(hacc23 >> 32));
        red_add_v4(p.head_out + (size_t)row * 4, hacc[0], hacc[1], hacc[2], hacc[3]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2 && rank != 0) mbar_arrive_remote(tempty_bar(acc), 0);  // the MMA issuer lives in the leader CTA
        else mbar_arrive(tempty_bar(acc));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();  // the peer may still signal barriers in this CTA's smem
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// wgrad: dW[N_out, K_in] += dY[M, N_out]^T X[M, K_in], db[N_out] += colsum(dY).
// MMA M <- 128 output features, MMA N <- BN input features, reduction over rows of dY / X in blocks
// of 64.  Both operands are MN-major: a TMA box is [64 rows x 64 contiguous features] (8 KB, 128-byte
// swizzle): descriptor SBO = 1024 B between 8-row groups along the reduction, LBO = 8192 B between
// 64-feature chunks; one K=16 MMA consumes two 8-row groups (2048 B).
// ---------------------------------------------------------------------------------------------
// CG = 2: CTA pair, M = 256 output features (128 per CTA), each CTA stages half of the BN input features.
template <int BN, int CG = 1>
struct WgradCfg {
  static constexpr int BOX_BYTES = 64 * 64 * 2;  // 8 KB
  static constexpr int A_BYTES = 2 * BOX_BYTES;  // 128 output features per CTA
  static constexpr int B_BOXES = BN / CG / 64;
  static constexpr int B_BYTES = B_BOXES * BOX_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256 && CG == 1) ? 4 : 6;
  static constexpr int ONES_BYTES = 2048;
  static constexpr int TMEM_COLS = (BN >= 256) ? 512 : (BN >= 128 ? 256 : 128);  // BN accumulator columns + 16 (bias gradient)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + ONES_BYTES + 1024 + 256;
};

struct WgradParams {
  float* dW;  // [N_out, K_in] fp32, accumulated into
  float* db;  // [N_out] or null
  int M, N_out, K_in, splits;
};

template <int BN, int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
             const WgradParams p) {
  using Cfg = WgradCfg<BN, CG>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ones_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = ones_base + Cfg::ONES_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * Cfg::STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = p.K_in / BN;               // along input features
  // split-major: the CTAs of one split (same sample range, all output tiles) are adjacent in launch order, run
  // concurrently and share their dY / X rows through L2 instead of re-reading them from HBM
  const int rank = (CG == 2) ? (int)cluster_ctarank() : 0;
  const int unit = blockIdx.x / CG;
  const int num_tiles = ((p.N_out + BM * CG - 1) / (BM * CG)) * tiles_n;
  const int tile = unit % num_tiles, split = unit / num_tiles;
  const int f0 = (tile / tiles_n) * (BM * CG) + rank * BM;  // output-feature offset of THIS CTA (rows of dW)
  const int i0 = (tile % tiles_n) * BN;                      // input-feature offset (cols of dW)
  const int kb_total = (p.M + BK - 1) / BK;
  const int kb_begin = (int)((long long)kb_total * split / p.splits);
  const int kb_end = (int)((long long)kb_total * (split + 1) / p.splits);
  const int nkb = kb_end - kb_begin;
  const bool do_bias = (p.db != nullptr) && (i0 == 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
    prefetch_tmap(&tmap_dy);
    prefetch_tmap(&tmap_x);
  }
  // tile of bf16 ones (0x3F80) for the bias-gradient MMA
  for (int i = threadIdx.x; i < Cfg::ONES_BYTES / 4; i += GEMM_THREADS)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(ones_base + 4u * i), "r"(0x3F803F80u) : "memory");
  fence_proxy_async_smem();
  __syncthreads();  // orders the generic-proxy fill above before tcgen05.alloc's write of the TMEM address (once per kernel)
  if (warp == 1) {
    tmem_alloc<CG>(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish<CG>();
  }
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (nkb > 0) {
    if (warp == 0) {
      if (lane == 0) {
        uint32_t stage = 0, phase = 0;
        const int ib = i0 + rank * (BN / CG) * (CG - 1);  // this CTA's share of the input features
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          if (CG == 1) {
            mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
#pragma unroll
            for (int c = 0; c < 2; ++c)
              tma_load_2d(sa + c * Cfg::BOX_BYTES, &tmap_dy, full_bar(stage), f0 + c * 64, kb * BK);
#pragma unroll
            for (int c = 0; c < Cfg::B_BOXES; ++c)
              tma_load_2d(sa + Cfg::A_BYTES + c * Cfg::BOX_BYTES, &tmap_x, full_bar(stage), ib + c * 64, kb * BK);
          } else {
            if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
#pragma unroll
            for (int c = 0; c < 2; ++c)
              tma_load_2d_pair(sa + c * Cfg::BOX_BYTES, &tmap_dy, full_bar(stage), f0 + c * 64, kb * BK);
#pragma unroll
            for (int c = 0; c < Cfg::B_BOXES; ++c)
              tma_load_2d_pair(sa + Cfg::A_BYTES + c * Cfg::BOX_BYTES, &tmap_x, full_bar(stage), ib + c * 64, kb * BK);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0 && rank == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(BN, 1, 1, BM * CG);
        constexpr uint32_t idesc_ones = make_idesc_bf16(16, 1, 1, BM * CG);
        const uint64_t ones_desc = make_smem_desc_sw128(ones_base, Cfg::BOX_BYTES, 1024);
        uint32_t stage = 0, phase = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t adesc = make_smem_desc_sw128(sa, Cfg::BOX_BYTES, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(sa + Cfg::A_BYTES, Cfg::BOX_BYTES, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // 16 reduction rows = 2 x 1024 B groups: +128 in the (addr >> 4) field
            const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
            umma_bf16<CG>(tmem_base, adesc + 128u * k, bdesc + 128u * k, idesc, accum);
            if (do_bias) umma_bf16<CG>(tmem_base + BN, adesc + 128u * k, ones_desc, idesc_ones, accum);
          }
          umma_commit<CG>(empty_bar(stage));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit<CG>(tfull_bar);
      }
    } else {
      const int q = warp & 3;
      const int row = f0 + q * 32 + lane;  // output feature
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + c * 32, v);
        tmem_ld_wait();
        if (row < p.N_out) {
          float* dst = p.dW + (size_t)row * p.K_in + i0 + c * 32;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            red_add_v4(dst + 4 * i, __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                       __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
        }
      }
      if (do_bias) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + BN, v);  // columns BN..BN+15 hold the bias gradient (all equal)
        tmem_ld_wait();
        if (row < p.N_out) atomicAdd(p.db + row, __uint_as_float(v[0]));
      }
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, Cfg::TMEM_COLS);
  }
}

// fp32 [N,K] -> bf16 [Npad,Kpad] (+ transposed [Kpad,Npad]), zero padded.  32 x 32 tiles through shared memory so
// that the read, the straight copy and the transposed copy are all coalesced (Npad, Kpad are multiples of 32).
__global__ void __launch_bounds__(256)
cast_weight_kernel(const float* __restrict__ W, int N, int K, int Npad, int Kpad, uint16_t* __restrict__ Wb,
                   uint16_t* __restrict__ Wt) {
  __shared__ uint16_t tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads, 4 rows each
  const int n0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, k = k0 + tx;
    const float v = (n < N && k < K) ? W[(long long)n * K + k] : 0.f;
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    const uint16_t bits = *reinterpret_cast<const uint16_t*>(&b);
    if (Wb) Wb[(long long)n * Kpad + k] = bits;
    tile[r][tx] = bits;
  }
  __syncthreads();
  if (Wt) {
#pragma unroll
    for (int r = ty; r < 32; r += 8) Wt[(long long)(k0 + r) * Npad + n0 + tx] = tile[tx][r];
  }
}

// fused AdamW (torch.optim.AdamW semantics: decoupled decay, bias-corrected moments)
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             long long n, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    float pi = p[i] * (1.f - lr * wd);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// One launch for a whole net: [AdamW on the flat fp32 parameters ->] bf16 operand copies of every layer (straight and
// transposed, zero padded) and the padded fp32 bias vectors.  A block owns one 32 x 32 tile of one layer's padded weight
// matrix (tile table in `entries`); the first tile column of a tile row also carries that row range of the bias.
// AdamW arithmetic is adamw_kernel's, element for element.
struct AdamHyper {
  float lr, beta1, beta2, eps, wd, bc1, bc2_sqrt;
};
__device__ __forceinline__ float adam_update(float* p, float* g, float* m, float* v, long long i, const AdamHyper& h,
                                             int zero_grad) {
  const float gi = g[i];
  const float pi = p[i] * (1.f - h.lr * h.wd);
  const float mi = h.beta1 * m[i] + (1.f - h.beta1) * gi;
  const float vi = h.beta2 * v[i] + (1.f - h.beta2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / h.bc2_sqrt + h.eps;
  const float out = pi - (h.lr / h.bc1) * (mi / denom);
  p[i] = out;
  if (zero_grad) g[i] = 0.f;
  return out;
}

__global__ void __launch_bounds__(256)
adamw_pack_kernel(const mip360_pack_entry* __restrict__ entries, int n_entries, float* p_base, float* g_base,
                  float* m_base, float* v_base, AdamHyper h, const float* __restrict__ hyper_dev, int do_adam,
                  int zero_grad) {
  __shared__ uint16_t tile[32][33];
  __shared__ int s_entry;
  if (threadIdx.x == 0) {
    int e = 0;
    while (e + 1 < n_entries && (int)blockIdx.x >= entries[e + 1].tile_begin) ++e;
    s_entry = e;
  }
  __syncthreads();
  const mip360_pack_entry E = entries[s_entry];
  if (do_adam && hyper_dev) {  // learning rate and bias corrections of THIS step, written by the host before a graph replay
    h.lr = hyper_dev[0];
    h.bc1 = hyper_dev[1];
    h.bc2_sqrt = hyper_dev[2];
  }
  const int local = (int)blockIdx.x - E.tile_begin;
  const int tiles_k = E.k_pad / 32;
  const int n0 = (local / tiles_k) * 32, k0 = (local % tiles_k) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int rows_all = E.rows[0] + E.rows[1];
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, k = k0 + tx;
    float val = 0.f;
    if (n < rows_all && k < E.K) {
      const float* src = n < E.rows[0] ? E.w_src[0] + (long long)n * E.K + k : E.w_src[1] + (long long)(n - E.rows[0]) * E.K + k;
      if (do_adam) {
        const long long i = src - p_base;
        val = adam_update(p_base, g_base, m_base, v_base, i, h, zero_grad);
      } else {
        val = *src;
      }
    }
    const __nv_bfloat16 b = __float2bfloat16_rn(val);
    const uint16_t bits = *reinterpret_cast<const uint16_t*>(&b);
    E.Wb[(long long)n * E.k_pad + k] = bits;
    tile[r][tx] = bits;
    // head weights once more as fp32 [k_pad, 4] (the bf16-rounded values, column-interleaved) for the fused-head epilogue
    if (E.w4 && n < 4) E.w4[(long long)k * 4 + n] = __bfloat162float(b);
  }
  __syncthreads();
  if (E.Wt) {
#pragma unroll
    for (int r = ty; r < 32; r += 8) E.Wt[(long long)(k0 + r) * E.n_pad + n0 + tx] = tile[tx][r];
  }
  if (k0 == 0 && ty == 0) {
    const int n = n0 + tx;
    float val = 0.f;
    if (n < rows_all) {
      const float* src = n < E.rows[0] ? E.b_src[0] + n : E.b_src[1] + (n - E.rows[0]);
      if (do_adam) val = adam_update(p_base, g_base, m_base, v_base, src - p_base, h, zero_grad);
      else val = *src;
    }
    if (E.bias) E.bias[n] = val;
  }
}

// Backward of a fused (<= 4 wide) head in ONE pass over the saved trunk output Y [M, N] (bf16):
//   dZ[r, c]  = (sum_h g[r, h] * w4[c, h]) * act'(Y[r, c])      gradient entering the last trunk layer      (dgrad)
//   dWh[h, c] += sum_r g[r, h] * Y[r, c],   dbh[h] += sum_r g[r, h]                                        (wgrad)
// replaces head_grad_pack + the 64-column head wgrad and dgrad GEMMs, which read Y twice and move a [M, 64] bf16
// gradient besides.  HBM-bound: 2N bytes read + 2N written + 16 per row.  A thread owns 8 columns (16-byte loads and
// stores), a block walks rows in a grid-stride loop with the weight gradient in registers and flushes it once.
// fp32 pair helpers (one instruction for two lanes of arithmetic: the kernel below is bound by instruction issue at the
// power-capped clock, not by HBM, unless its FMAs are packed)
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

template <int ACT>
__global__ void __launch_bounds__(256, 2)
head_bwd_kernel(const float* __restrict__ g, const float* __restrict__ w4, const uint16_t* __restrict__ Y, long long M,
                int N, uint16_t* __restrict__ dZ, float* __restrict__ dWh, int ldw, float* __restrict__ dbh) {
  const int cols8 = N >> 3;                       // 16-byte column groups per row
  const int cg = threadIdx.x % cols8;             // this thread's column group (N <= 2048: cols8 <= 256)
  const int rows_per_pass = blockDim.x / cols8;   // rows a block covers per step
  const int rsub = threadIdx.x / cols8;
  const int c0 = cg * 8;
  // head weights of the thread's 8 columns as column PAIRS per head: wp[q][h] = (w[2q][h], w[2q+1][h])
  uint64_t wp[4][4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(w4) + c0 + 2 * q);
    const float4 b = __ldg(reinterpret_cast<const float4*>(w4) + c0 + 2 * q + 1);
    wp[q][0] = f2_pack(a.x, b.x); wp[q][1] = f2_pack(a.y, b.y); wp[q][2] = f2_pack(a.z, b.z); wp[q][3] = f2_pack(a.w, b.w);
  }
  uint64_t acc[4][4];  // weight-gradient partial sums, same pairing
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int h = 0; h < 4; ++h) acc[q][h] = 0ull;
  float gsum[4] = {0.f, 0.f, 0.f, 0.f};
  const uint64_t one2 = f2_pack(1.f, 1.f), neg2 = f2_pack(-1.f, -1.f);
  if (rsub < rows_per_pass) {
    const long long stride = (long long)gridDim.x * rows_per_pass;
    constexpr int U = 4;  // rows in flight per thread: all loads of a step are issued before the arithmetic
    for (long long r0 = (long long)blockIdx.x * rows_per_pass + rsub; r0 < M; r0 += U * stride) {
      float4 gr[U];
      uint4 yv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long r = r0 + u * stride;
        if (r < M) {
          gr[u] = __ldg(reinterpret_cast<const float4*>(g) + r);
          yv[u] = __ldg(reinterpret_cast<const uint4*>(Y + r * N + c0));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long r = r0 + u * stride;
        if (r >= M) break;
        const uint32_t yw[4] = {yv[u].x, yv[u].y, yv[u].z, yv[u].w};
        const uint64_t g0 = f2_pack(gr[u].x, gr[u].x), g1 = f2_pack(gr[u].y, gr[u].y), g2 = f2_pack(gr[u].z, gr[u].z),
                       g3 = f2_pack(gr[u].w, gr[u].w);
        uint32_t out[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint64_t y2 = f2_pack(__uint_as_float(yw[q] << 16), __uint_as_float(yw[q] & 0xffff0000u));
          // dz pair = sum_h g_h * (w[2q][h], w[2q+1][h])
          uint64_t dz = f2_mul(g0, wp[q][0]);
          dz = f2_fma(g1, wp[q][1], dz);
          dz = f2_fma(g2, wp[q][2], dz);
          dz = f2_fma(g3, wp[q][3], dz);
          // weight gradient: acc[q][h] += g_h * (y[2q], y[2q+1])
          acc[q][0] = f2_fma(g0, y2, acc[q][0]);
          acc[q][1] = f2_fma(g1, y2, acc[q][1]);
          acc[q][2] = f2_fma(g2, y2, acc[q][2]);
          acc[q][3] = f2_fma(g3, y2, acc[q][3]);
          float lo, hi;
          if (ACT == ACT_SIGMOID) {
            dz = f2_mul(dz, f2_mul(y2, f2_fma(y2, neg2, one2)));  // * y (1 - y)
            f2_unpack(dz, lo, hi);
          } else {
            f2_unpack(dz, lo, hi);
            if (ACT == ACT_RELU) {
              float ylo, yhi;
              f2_unpack(y2, ylo, yhi);
              lo = ylo > 0.f ? lo : 0.f;
              hi = yhi > 0.f ? hi : 0.f;
            }
          }
          out[q] = pack_bf16x2(lo, hi);
        }
        *reinterpret_cast<uint4*>(dZ + r * N + c0) = make_uint4(out[0], out[1], out[2], out[3]);
        if (cg == 0) { gsum[0] += gr[u].x; gsum[1] += gr[u].y; gsum[2] += gr[u].z; gsum[3] += gr[u].w; }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        float lo, hi;
        f2_unpack(acc[q][h], lo, hi);
        atomicAdd(dWh + (long long)h * ldw + c0 + 2 * q, lo);
        atomicAdd(dWh + (long long)h * ldw + c0 + 2 * q + 1, hi);
      }
    if (cg == 0 && dbh) {
#pragma unroll
      for (int h = 0; h < 4; ++h) atomicAdd(dbh + h, gsum[h]);
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Layer-fused forward of a narrow MLP (the proposal net, model.py:43-53,91): K0 -> W -> W -> W -> W -> head, W = 256.
// One persistent CTA per SM walks 128-row tiles; a tile's activations never leave the SM:
//   warp 0      TMA: the tile's input rows [128 x K0] and, layer by layer, the weights as [N x 64] K-chunks (from L2)
//   warp 1      one lane issues tcgen05.mma; the A operand of layer s >= 1 is the previous layer's output, written by
//               the epilogue warps into shared memory in the UMMA K-major 128-byte-swizzled layout ([128 x 64] chunks)
//   warps 2-9   epilogue: TMEM accumulator -> bias + ReLU / Sigmoid -> bf16 -> activation buffer (and, when the
//               activations are needed for the backward pass, a TMA store of the same 4 KB box to HBM).  A tile costs
//               52 N=256 MMAs (~4.5 us) but four 128 x 256 epilogues, which four warps need ~10 us for (each waits out
//               its own tcgen05.ld latencies): the kernel is epilogue-bound, so every TMEM lane quarter is served by
//               FOUR warps, one per 64-column chunk
// Pipelining inside a tile: layer s+1's k-chunk j only needs output columns [64j, 64j+64) of layer s, so its MMAs start
// as soon as the epilogue has written that chunk (per-chunk mbarriers); layers alternate between two TMEM accumulators
// and two activation buffers.  Arithmetic (MMA order, epilogue) is that of linear_kernel: results are bit-identical to
// the layer-by-layer path.
// ---------------------------------------------------------------------------------------------
constexpr int PF_W = 256;        // trunk width
constexpr int PF_TRUNK = 4;      // trunk layers
constexpr int PF_STAGES = 4;     // weight ring: 128 KB in flight hide the L2 latency of the 32 KB chunks
constexpr int PF_EPI_WARPS = 16; // four per TMEM lane quarter, one per 64-column chunk
constexpr int PF_THREADS = 64 + 32 * PF_EPI_WARPS;
struct PropFusedCfg {
  static constexpr int X_BYTES = BM * BK * 2;                    // 16 KB (K0 = 64)
  static constexpr int ACT_BYTES = BM * PF_W * 2;                // 64 KB: 4 chunks of [128 x 64]
  static constexpr int W_STAGE_BYTES = PF_W * BK * 2;            // 32 KB
  static constexpr int BIAS_BYTES = (PF_TRUNK * PF_W + 64) * 4;  // 4 trunk biases + padded head bias
  static constexpr int BAR_BYTES = 256;
  // ONE activation buffer: a layer's input is dead once its MMAs have completed, which is exactly when the epilogue starts
  // to write that layer's output — in place, chunk by chunk, as the next layer's A operand
  static constexpr int SMEM_BYTES = X_BYTES + ACT_BYTES + PF_STAGES * W_STAGE_BYTES + BIAS_BYTES + BAR_BYTES + 1024;
  static_assert(SMEM_BYTES <= 232448, "fused proposal MLP exceeds shared memory");
};
struct PropFusedParams {
  const float* bias[PF_TRUNK + 1];  // [256] x 4, head [64]
  int act[PF_TRUNK];                // ACT_RELU / ACT_SIGMOID_FAST
  float* out;                       // [M, n_valid] fp32 head outputs (no activation)
  int M, n_valid, save_acts;
};

__global__ void __launch_bounds__(PF_THREADS, 1)
prop_fused_fwd_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w0,
                      const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_w2,
                      const __grid_constant__ CUtensorMap tmap_w3, const __grid_constant__ CUtensorMap tmap_wh,
                      const __grid_constant__ CUtensorMap tmap_a0, const __grid_constant__ CUtensorMap tmap_a1,
                      const __grid_constant__ CUtensorMap tmap_a2, const __grid_constant__ CUtensorMap tmap_a3,
                      const PropFusedParams p) {
  using Cfg = PropFusedCfg;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t xbuf = smem_base;
  const uint32_t actbuf = xbuf + Cfg::X_BYTES;                       // 64 KB
  const uint32_t wring = actbuf + Cfg::ACT_BYTES;                    // PF_STAGES x 32 KB
  const uint32_t bias_smem = wring + PF_STAGES * Cfg::W_STAGE_BYTES;
  const uint32_t bar_base = bias_smem + Cfg::BIAS_BYTES;
  auto w_full = [&](int s) { return bar_base + 8u * s; };
  auto w_empty = [&](int s) { return bar_base + 8u * (PF_STAGES + s); };
  const uint32_t x_full = bar_base + 8u * (2 * PF_STAGES), x_empty = x_full + 8u;
  auto act_ready = [&](int j) { return bar_base + 8u * (2 * PF_STAGES + 2 + j); };
  auto acc_full = [&](int a) { return bar_base + 8u * (2 * PF_STAGES + 6 + a); };
  auto acc_empty = [&](int a) { return bar_base + 8u * (2 * PF_STAGES + 8 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * PF_STAGES + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = (p.M + BM - 1) / BM;
  const CUtensorMap* tmap_w[PF_TRUNK + 1] = {&tmap_w0, &tmap_w1, &tmap_w2, &tmap_w3, &tmap_wh};
  const CUtensorMap* tmap_a[PF_TRUNK] = {&tmap_a0, &tmap_a1, &tmap_a2, &tmap_a3};

  if (threadIdx.x == 0) {
    for (int s = 0; s < PF_STAGES; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    for (int j = 0; j < 4; ++j) mbar_init(act_ready(j), 4);  // one arrival per TMEM lane quarter
    for (int a = 0; a < 2; ++a) { mbar_init(acc_full(a), 1); mbar_init(acc_empty(a), PF_EPI_WARPS); }
    fence_barrier_init();
    prefetch_tmap(&tmap_x);
    for (int s = 0; s <= PF_TRUNK; ++s) prefetch_tmap(tmap_w[s]);
  }
  // biases of every layer into shared memory (read back as broadcasts by the epilogue)
  for (int i = threadIdx.x; i < PF_TRUNK * PF_W + 64; i += PF_THREADS) {
    const int l = i < PF_TRUNK * PF_W ? i / PF_W : PF_TRUNK;
    const int c = i - l * PF_W;
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_smem + 4u * i), "f"(__ldg(p.bias[l] + c)) : "memory");
  }
  if (warp == 1) {
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: input rows of the tile, then the weights of the five GEMM stages as K-chunks =====
      uint32_t ws = 0, wphase = 0, xphase = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        mbar_wait(x_empty, xphase ^ 1u);
        mbar_arrive_expect_tx(x_full, Cfg::X_BYTES);
        tma_load_2d(xbuf, &tmap_x, x_full, 0, tile * BM);
        xphase ^= 1u;
        for (int s = 0; s <= PF_TRUNK; ++s) {
          const int kc = s == 0 ? 1 : PF_W / BK;
          const uint32_t bytes = (uint32_t)(s < PF_TRUNK ? PF_W : 64) * BK * 2;
          for (int j = 0; j < kc; ++j) {
            mbar_wait(w_empty(ws), wphase ^ 1u);
            mbar_arrive_expect_tx(w_full(ws), bytes);
            tma_load_2d(wring + ws * Cfg::W_STAGE_BYTES, tmap_w[s], w_full(ws), j * BK, 0);
            if (++ws == PF_STAGES) { ws = 0; wphase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc_trunk = make_idesc_bf16(PF_W, 0, 0, BM);
      constexpr uint32_t idesc_head = make_idesc_bf16(64, 0, 0, BM);
      uint32_t ws = 0, wphase = 0, xphase = 0;
      uint32_t acc_uses[2] = {0, 0};   // completed uses of each accumulator (phase of acc_empty)
      uint32_t fills = 0;              // fills of the activation buffer consumed so far (phase of act_ready)
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (int s = 0; s <= PF_TRUNK; ++s) {
          const int ab = s & 1;
          mbar_wait(acc_empty(ab), (acc_uses[ab] & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + ab * PF_W;
          const int kc = s == 0 ? 1 : PF_W / BK;
          for (int j = 0; j < kc; ++j) {
            mbar_wait(w_full(ws), wphase);
            if (s == 0) mbar_wait(x_full, xphase);
            else mbar_wait(act_ready(j), fills & 1u);
            tc_fence_after();
            const uint32_t a_addr = s == 0 ? xbuf : actbuf + j * (BM * BK * 2);
            const uint64_t adesc = make_smem_desc_sw128(a_addr, 16, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(wring + ws * Cfg::W_STAGE_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_bf16<1>(d_tmem, adesc + 2u * k, bdesc + 2u * k, s < PF_TRUNK ? idesc_trunk : idesc_head,
                           (j | k) != 0 ? 1u : 0u);
            umma_commit<1>(w_empty(ws));
            if (++ws == PF_STAGES) { ws = 0; wphase ^= 1u; }
          }
          if (s == 0) {
            umma_commit<1>(x_empty);  // the input rows may be replaced by the next tile's
            xphase ^= 1u;
          } else {
            ++fills;
          }
          umma_commit<1>(acc_full(ab));
          ++acc_uses[ab];
        }
      }
    }
  } else {
    // ===== epilogue warps: TMEM lanes [32q, 32q+32), column chunks half and half + 2 =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t row_off = (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);
    uint32_t acc_seen[2] = {0, 0};
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int m0 = tile * BM;
      const int row = m0 + q * 32 + lane;
      for (int s = 0; s <= PF_TRUNK; ++s) {
        const int ab = s & 1;
        mbar_wait(acc_full(ab), acc_seen[ab] & 1u);
        ++acc_seen[ab];
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + ab * PF_W;
        if (s < PF_TRUNK) {
          const uint32_t bsm = bias_smem + 4u * (s * PF_W);
          if (p.save_acts && lane == 0) tma_store_wait_read<0>();  // the previous layer's boxes have been read by their stores
          __syncwarp();
#pragma unroll 1
          for (int jj = half; jj < PF_W / 64; jj += PF_EPI_WARPS / 4) {
            const uint32_t box = actbuf + jj * (BM * BK * 2) + q * (32 * 128);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t v[32], packed[16];
              tmem_ld_32x32(t_row + jj * 64 + h * 32, v);
              tmem_ld_wait();
              if (p.act[s] == ACT_RELU) fwd_relu_packed(v, bsm + (jj * 64 + h * 32) * 4u, packed);
              else fwd_sigmoid_fast_packed(v, bsm + (jj * 64 + h * 32) * 4u, packed);
#pragma unroll
              for (int c = 0; c < 4; ++c)
                st_shared_v4(box + row_off + (((4 * h + c) ^ sw) << 4), packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (p.save_acts) {
                tma_store_2d(tmap_a[s], box, jj * 64, m0 + q * 32);
                tma_store_commit();
              }
              mbar_arrive(act_ready(jj));
            }
          }
        } else if (half == 0) {
          // head: n_valid <= 8 real columns, no activation (model.py:91: softplus follows in the compositing kernel)
          uint32_t v8[8];
          tmem_ld_32x8(t_row, v8);
          tmem_ld_wait();
          if (row < p.M) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float bv;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(bv) : "r"(bias_smem + 4u * (PF_TRUNK * PF_W + i)));
              if (i < p.n_valid) p.out[(size_t)row * p.n_valid + i] = __uint_as_float(v8[i]) + bv;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(ab));
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// The same fused forward for CTA pairs, two row tiles in flight ("ping-pong").  The single-CTA kernel above leaves the
// tensor pipe idle while a tile's epilogue runs and the epilogue warps idle while its MMAs run (a tile's layers form a
// chain), and it pulls the full 0.5 MB of weights through L2 for every 128 rows.  Here
//   * a cluster of two CTAs (tcgen05 cta_group::2, UMMA M = 256) owns 256-row tiles: each CTA keeps its own 128 rows of
//     activations and stages HALF of every weight chunk (16 KB instead of 32 KB), so weight traffic per row halves;
//   * every pair works on TWO tiles ("slots"), one accumulator (256 TMEM columns) and one in-place activation buffer
//     (64 KB) each: the leader issues layer s of slot 0, layer s of slot 1, layer s+1 of slot 0, ... while the epilogue
//     warps of both CTAs follow one unit behind, so the MMAs of one slot cover the epilogue of the other.
// What bounds it: reading a 128 x 256 fp32 accumulator out of tensor memory (128 KB at 64 B/clk/SM) takes as long as the 16
// MMAs that produce it (K = 256), and a slot's MMA -> epilogue -> next layer's MMA chain is serial, so with two slots the
// tensor pipe reaches ~50 % (ncu).  Splitting a layer into two N = 128 column halves so that the first half is read while
// the second accumulates was built and measured slower (0.59 vs 0.46 ms per 1M rows): M = 128-per-CTA x N = 128 MMAs need
// 128 B/clk of shared-memory operand bandwidth, all there is.  Inference launches four epilogue warps per TMEM lane
// quarter, training (activations saved with TMA stores) two.
// Barriers that the MMA issuer waits on live in the leader CTA (the peer's TMA completes on them, the peer's epilogue warps
// arrive remotely with release.cluster after fence.proxy.async, so their activation stores are visible to the MMA);
// tcgen05.commit multicasts to both CTAs.  Arithmetic is again that of linear_kernel: bit-identical results.
// ---------------------------------------------------------------------------------------------
constexpr int PP_STAGES = 3;  // x 16 KB per CTA
struct PropPairCfg {
  static constexpr int X_BYTES = BM * BK * 2;             // 16 KB per slot
  static constexpr int ACT_BYTES = BM * PF_W * 2;         // 64 KB per slot
  static constexpr int W_STAGE_BYTES = (PF_W / 2) * BK * 2;  // 16 KB: this CTA's half of a [256 x 64] weight chunk
  static constexpr int BIAS_BYTES = (PF_TRUNK * PF_W + 64) * 4;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = 2 * X_BYTES + 2 * ACT_BYTES + PP_STAGES * W_STAGE_BYTES + BIAS_BYTES + BAR_BYTES + 1024;
  static_assert(SMEM_BYTES <= 232448, "pair-fused proposal MLP exceeds shared memory");
};

__global__ void __launch_bounds__(PF_THREADS, 1)
prop_fused_pair_fwd_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w0,
                           const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_w2,
                           const __grid_constant__ CUtensorMap tmap_w3, const __grid_constant__ CUtensorMap tmap_wh,
                           const __grid_constant__ CUtensorMap tmap_a0, const __grid_constant__ CUtensorMap tmap_a1,
                           const __grid_constant__ CUtensorMap tmap_a2, const __grid_constant__ CUtensorMap tmap_a3,
                           const PropFusedParams p) {
  using Cfg = PropPairCfg;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto xbuf = [&](int P) { return smem_base + (uint32_t)P * Cfg::X_BYTES; };
  auto actbuf = [&](int P) { return smem_base + 2u * Cfg::X_BYTES + (uint32_t)P * Cfg::ACT_BYTES; };
  const uint32_t wring = smem_base + 2u * Cfg::X_BYTES + 2u * Cfg::ACT_BYTES;
  const uint32_t bias_smem = wring + PP_STAGES * Cfg::W_STAGE_BYTES;
  const uint32_t bar_base = bias_smem + Cfg::BIAS_BYTES;
  auto w_full = [&](int s) { return bar_base + 8u * s; };
  auto w_empty = [&](int s) { return bar_base + 8u * (PP_STAGES + s); };
  auto x_full = [&](int P) { return bar_base + 8u * (2 * PP_STAGES + P); };
  auto x_empty = [&](int P) { return bar_base + 8u * (2 * PP_STAGES + 2 + P); };
  auto acc_full = [&](int P) { return bar_base + 8u * (2 * PP_STAGES + 4 + P); };
  auto epi_done = [&](int P) { return bar_base + 8u * (2 * PP_STAGES + 6 + P); };  // leader's: accumulator drained AND activations written
  const uint32_t tmem_slot = bar_base + 8u * (2 * PP_STAGES + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nepi = (int)(blockDim.x >> 5) - 2;   // epilogue warps: 8 or 16
  const int rank = (int)cluster_ctarank();
  const int unit = blockIdx.x >> 1, nunits = gridDim.x >> 1;
  const int tiles = (p.M + 2 * BM - 1) / (2 * BM);        // 256-row tiles
  const int iters = (tiles + 2 * nunits - 1) / (2 * nunits);
  auto tile_of = [&](int it, int P) { return (it * nunits + unit) * 2 + P; };
  const CUtensorMap* tmap_w[PF_TRUNK + 1] = {&tmap_w0, &tmap_w1, &tmap_w2, &tmap_w3, &tmap_wh};
  const CUtensorMap* tmap_a[PF_TRUNK] = {&tmap_a0, &tmap_a1, &tmap_a2, &tmap_a3};

  if (threadIdx.x == 0) {
    for (int s = 0; s < PP_STAGES; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int P = 0; P < 2; ++P) {
      mbar_init(x_full(P), 1);
      mbar_init(x_empty(P), 1);
      mbar_init(acc_full(P), 1);
      mbar_init(epi_done(P), 2 * nepi);  // both CTAs' epilogue warps
    }
    fence_barrier_init();
    prefetch_tmap(&tmap_x);
    for (int s = 0; s <= PF_TRUNK; ++s) prefetch_tmap(tmap_w[s]);
  }
  for (int i = threadIdx.x; i < PF_TRUNK * PF_W + 64; i += blockDim.x) {
    const int l = i < PF_TRUNK * PF_W ? i / PF_W : PF_TRUNK;
    const int c = i - l * PF_W;
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_smem + 4u * i), "f"(__ldg(p.bias[l] + c)) : "memory");
  }
  __syncthreads();
  if (warp == 1) {
    tmem_alloc<2>(tmem_slot, 512);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  cluster_sync();  // barrier inits visible to the peer before any remote arrive / multicast commit
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (one per CTA; completion bytes go to the LEADER's barriers) =====
      uint32_t ws = 0, wphase = 0, xuses[2] = {0, 0};
      auto load_x = [&](int it, int P) {
        const int t = tile_of(it, P);
        if (it >= iters || t >= tiles) return;
        mbar_wait(x_empty(P), (xuses[P] & 1u) ^ 1u);
        ++xuses[P];
        if (rank == 0) mbar_arrive_expect_tx(x_full(P), 2 * Cfg::X_BYTES);
        tma_load_2d_pair(xbuf(P), &tmap_x, x_full(P), 0, t * 2 * BM + rank * BM);
      };
      load_x(0, 0);
      load_x(0, 1);
      for (int it = 0; it < iters; ++it) {
        for (int s = 0; s <= PF_TRUNK; ++s) {
          const int kc = s == 0 ? 1 : PF_W / BK;
          const int rows = s < PF_TRUNK ? PF_W / 2 : 32;    // this CTA's half of the layer's output features
          for (int P = 0; P < 2; ++P) {
            if (tile_of(it, P) >= tiles) continue;
            for (int j = 0; j < kc; ++j) {
              mbar_wait(w_empty(ws), wphase ^ 1u);
              if (rank == 0) mbar_arrive_expect_tx(w_full(ws), 2u * rows * BK * 2);
              tma_load_2d_pair(wring + ws * Cfg::W_STAGE_BYTES, tmap_w[s], w_full(ws), j * BK, rank * rows);
              if (++ws == PP_STAGES) { ws = 0; wphase ^= 1u; }
            }
          }
          // the next iteration's input rows, a whole iteration ahead (their buffers were released by this iteration's layer 0)
          if (s == 1) { load_x(it + 1, 0); load_x(it + 1, 1); }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===== MMA issuer (leader CTA): M = 256 MMAs over both CTAs' rows =====
      constexpr uint32_t idesc_trunk = make_idesc_bf16(PF_W, 0, 0, 2 * BM);
      constexpr uint32_t idesc_head = make_idesc_bf16(64, 0, 0, 2 * BM);
      uint32_t ws = 0, wphase = 0, units[2] = {0, 0}, xuses[2] = {0, 0};
      for (int it = 0; it < iters; ++it) {
        for (int s = 0; s <= PF_TRUNK; ++s) {
          const int kc = s == 0 ? 1 : PF_W / BK;
          for (int P = 0; P < 2; ++P) {
            if (tile_of(it, P) >= tiles) continue;
            // the slot's previous unit has left the accumulator and (s >= 1) written this layer's input
            mbar_wait(epi_done(P), (units[P] & 1u) ^ 1u);
            ++units[P];
            if (s == 0) { mbar_wait(x_full(P), xuses[P] & 1u); ++xuses[P]; }
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)P * PF_W;
            for (int j = 0; j < kc; ++j) {
              mbar_wait(w_full(ws), wphase);
              tc_fence_after();
              const uint32_t a_addr = s == 0 ? xbuf(P) : actbuf(P) + j * (BM * BK * 2);
              const uint64_t adesc = make_smem_desc_sw128(a_addr, 16, 1024);
              const uint64_t bdesc = make_smem_desc_sw128(wring + ws * Cfg::W_STAGE_BYTES, 16, 1024);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)
                umma_bf16<2>(d_tmem, adesc + 2u * k, bdesc + 2u * k, s < PF_TRUNK ? idesc_trunk : idesc_head,
                             (j | k) != 0 ? 1u : 0u);
              umma_commit<2>(w_empty(ws));
              if (++ws == PP_STAGES) { ws = 0; wphase ^= 1u; }
            }
            if (s == 0) umma_commit<2>(x_empty(P));
            umma_commit<2>(acc_full(P));
          }
        }
      }
    }
  } else {
    // ===== epilogue warps (both CTAs): TMEM lanes [32q, 32q+32), column chunks half and half + 2 =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t row_off = (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);
    uint32_t units[2] = {0, 0};
    for (int it = 0; it < iters; ++it) {
      const bool both = tile_of(it, 1) < tiles;
      for (int s = 0; s <= PF_TRUNK; ++s) {
        for (int P = 0; P < 2; ++P) {
          if (tile_of(it, P) >= tiles) continue;
          const int m0 = tile_of(it, P) * 2 * BM + rank * BM;
          mbar_wait(acc_full(P), units[P] & 1u);
          ++units[P];
          tc_fence_after();
          const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)P * PF_W;
          if (s < PF_TRUNK) {
            const uint32_t bsm = bias_smem + 4u * (s * PF_W);
            if (p.save_acts) {
              // this slot's boxes were last read by the stores of its previous unit; the bulk group of the other slot's
              // unit in between (one group younger) may still be in flight
              if (lane == 0) {
                if (both) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
              }
              __syncwarp();
            }
#pragma unroll 1
            for (int jj = half; jj < PF_W / 64; jj += nepi / 4) {
              const uint32_t box = actbuf(P) + jj * (BM * BK * 2) + q * (32 * 128);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                uint32_t v[32], packed[16];
                tmem_ld_32x32(t_row + jj * 64 + h * 32, v);
                tmem_ld_wait();
                if (p.act[s] == ACT_RELU) fwd_relu_packed(v, bsm + (jj * 64 + h * 32) * 4u, packed);
                else fwd_sigmoid_fast_packed(v, bsm + (jj * 64 + h * 32) * 4u, packed);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  st_shared_v4(box + row_off + (((4 * h + c) ^ sw) << 4), packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (p.save_acts && lane == 0) tma_store_2d(tmap_a[s], box, jj * 64, m0 + q * 32);
            }
            if (p.save_acts && lane == 0) tma_store_commit();   // one bulk group per unit
          } else if (half == 0) {
            // head: n_valid <= 8 real columns, no activation (model.py:91: softplus follows in the compositing kernel)
            const int row = m0 + q * 32 + lane;
            uint32_t v8[8];
            tmem_ld_32x8(t_row, v8);
            tmem_ld_wait();
            if (row < p.M) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float bv;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(bv) : "r"(bias_smem + 4u * (PF_TRUNK * PF_W + i)));
                if (i < p.n_valid) p.out[(size_t)row * p.n_valid + i] = __uint_as_float(v8[i]) + bv;
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (rank != 0) mbar_arrive_remote(epi_done(P), 0);
            else mbar_arrive(epi_done(P));
          }
        }
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  cluster_sync();  // the peer may still signal barriers in this CTA's shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<2>(tmem_base, 512);
  }
}

// ---- host side: tensor maps ------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// bf16 row-major [rows, cols]; box = [box_rows, 64 cols], 128-byte swizzle, OOB reads give zeros
static int make_tmap(CUtensorMap* out, const void* ptr, long long rows, long long cols, int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return MIP360_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld box_rows=%d ptr=%p", (int)r, rows, cols, box_rows,
              ptr);
    return MIP360_ERR_CUDA;
  }
  return MIP360_OK;
}

template <int BN, int EPI, int CG, int OCC = 1>
static int launch_linear(const uint16_t* A, const uint16_t* Bw, const uint16_t* yprev, const LinearParams& p,
                         cudaStream_t stream) {
  using Cfg = LinearCfg<BN, CG, OCC>;
  static bool configured[MAX_DEVICES] = {};  // the attribute is per device (one process may drive several)
  const int dev = current_device();
  constexpr int SMEM = Cfg::SMEM_BYTES + ((EPI == EPI_FWD_HEAD && Cfg::HEAD_SMEM) ? Cfg::HEAD_BYTES : 0);
  if (!configured[dev]) {
    MIP_CUDA(cudaFuncSetAttribute(linear_kernel<BN, EPI, CG, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured[dev] = true;
  }
  CUtensorMap ta, tb, tout, ty;
  int rc;
  if ((rc = make_tmap(&ta, A, p.M, p.K, BM)) != MIP360_OK) return rc;
  if ((rc = make_tmap(&tb, Bw, p.N, p.K, BN / CG)) != MIP360_OK) return rc;
  tout = ta;
  ty = ta;
  if (p.out_bf16 && (rc = make_tmap(&tout, p.out_bf16, p.M, p.N, 32)) != MIP360_OK) return rc;
  if (EPI == EPI_DGRAD && (rc = make_tmap(&ty, yprev, p.M, p.N, 32)) != MIP360_OK) return rc;
  const int tiles = ((p.M + BM * CG - 1) / (BM * CG)) * (p.N / BN);
  const int units = sm_count() * OCC / CG;
  const int grid = (tiles < units ? tiles : units) * CG;
  if (CG == 1) {
    linear_kernel<BN, EPI, CG, OCC><<<grid, GEMM_THREADS, SMEM, stream>>>(ta, tb, tout, ty, p);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MIP_CUDA(cudaLaunchKernelEx(&cfg, linear_kernel<BN, EPI, CG, OCC>, ta, tb, tout, ty, p));
  }
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

template <int EPI>
static int dispatch_linear(const uint16_t* A, const uint16_t* Bw, const uint16_t* yprev, const LinearParams& p,
                           cudaStream_t stream) {
  // CTA pairs (cta_group::2) for the big layers; single CTAs when there are too few 256-row tiles to fill the pairs
  const bool pair_ok = option(OPT_CTA_PAIR);
  if (p.N % 256 == 0 && p.K >= 512 && pair_ok && p.M >= 256 * (sm_count() / 2))
    return launch_linear<256, EPI, 2>(A, Bw, yprev, p, stream);
  const bool short_ok = option(OPT_SHORT_K);
  if (p.N % 128 == 0 && p.K <= (EPI == EPI_DGRAD ? 256 : 128) && short_ok && p.M >= 128 * sm_count())
    return launch_linear<128, EPI, 1, 2>(A, Bw, yprev, p, stream);
  if (p.N % 256 == 0) return launch_linear<256, EPI, 1>(A, Bw, yprev, p, stream);
  if (p.N == 128) return launch_linear<128, EPI, 1>(A, Bw, yprev, p, stream);
  if (p.N == 64) return launch_linear<64, EPI, 1>(A, Bw, yprev, p, stream);
  set_error("linear: N=%d not supported (need 64, 128 or a multiple of 256)", p.N);
  return MIP360_ERR_UNSUPPORTED;
}

static int launch_prop_fused(const uint16_t* x, int M, const mip360_layer* trunk, const mip360_layer* head, int n_valid,
                             uint16_t* const* acts, float* out, cudaStream_t stream) {
  using Cfg = PropFusedCfg;
  static bool configured[MAX_DEVICES] = {};
  const int dev = current_device();
  if (!configured[dev]) {
    MIP_CUDA(cudaFuncSetAttribute(prop_fused_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured[dev] = true;
  }
  CUtensorMap tx, tw[PF_TRUNK + 1], ta[PF_TRUNK];
  int rc;
  if ((rc = make_tmap(&tx, x, M, BK, BM)) != MIP360_OK) return rc;
  for (int l = 0; l < PF_TRUNK; ++l)
    if ((rc = make_tmap(&tw[l], trunk[l].W, PF_W, trunk[l].k_pad, PF_W)) != MIP360_OK) return rc;
  if ((rc = make_tmap(&tw[PF_TRUNK], head->W, 64, PF_W, 64)) != MIP360_OK) return rc;
  for (int l = 0; l < PF_TRUNK; ++l) {
    ta[l] = tx;
    if (acts && (rc = make_tmap(&ta[l], acts[l], M, PF_W, 32)) != MIP360_OK) return rc;
  }
  PropFusedParams p{};
  for (int l = 0; l < PF_TRUNK; ++l) {
    p.bias[l] = trunk[l].bias;
    p.act[l] = trunk[l].act == ACT_SIGMOID ? ACT_SIGMOID_FAST : trunk[l].act;
  }
  p.bias[PF_TRUNK] = head->bias;
  p.out = out;
  p.M = M;
  p.n_valid = n_valid;
  p.save_acts = acts ? 1 : 0;
  const int pairs = sm_count() / 2;
  if (option(OPT_CTA_PAIR) && M >= 2 * BM * 2 * pairs) {
    // enough 256-row tiles to give every CTA pair two of them: the ping-pong pair kernel
    static bool configured2[MAX_DEVICES] = {};
    if (!configured2[dev]) {
      MIP_CUDA(cudaFuncSetAttribute(prop_fused_pair_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PropPairCfg::SMEM_BYTES));
      configured2[dev] = true;
    }
    for (int l = 0; l < PF_TRUNK; ++l)
      if ((rc = make_tmap(&tw[l], trunk[l].W, PF_W, trunk[l].k_pad, PF_W / 2)) != MIP360_OK) return rc;
    if ((rc = make_tmap(&tw[PF_TRUNK], head->W, 64, PF_W, 32)) != MIP360_OK) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    // Same-box A/B (1M rows): inference 0.45 ms with four epilogue warps per TMEM lane quarter against 0.48 with two;
    // training (activations saved by TMA stores, whose issue slots then matter less than their queueing) 0.58 ms with two
    // against 0.66 with four, and 0.66-0.68 ms with plain 16-byte stores from a row-major read-back of the boxes.
    cfg.blockDim = dim3(64 + 32 * (acts ? 8 : 16));
    cfg.dynamicSmemBytes = PropPairCfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MIP_CUDA(cudaLaunchKernelEx(&cfg, prop_fused_pair_fwd_kernel, tx, tw[0], tw[1], tw[2], tw[3], tw[4], ta[0], ta[1], ta[2],
                                ta[3], p));
    MIP_LAUNCH_CHECK();
    return MIP360_OK;
  }
  const int tiles = (M + BM - 1) / BM;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  prop_fused_fwd_kernel<<<grid, PF_THREADS, Cfg::SMEM_BYTES, stream>>>(tx, tw[0], tw[1], tw[2], tw[3], tw[4], ta[0], ta[1],
                                                                        ta[2], ta[3], p);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

template <int BN, int CG>
static int launch_wgrad(const uint16_t* dY, const uint16_t* X, WgradParams p, cudaStream_t stream) {
  using Cfg = WgradCfg<BN, CG>;
  static bool configured[MAX_DEVICES] = {};
  const int dev = current_device();
  if (!configured[dev]) {
    MIP_CUDA(cudaFuncSetAttribute(wgrad_kernel<BN, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured[dev] = true;
  }
  CUtensorMap tdy, tx;
  int rc;
  if ((rc = make_tmap(&tdy, dY, p.M, p.N_out, 64)) != MIP360_OK) return rc;
  if ((rc = make_tmap(&tx, X, p.M, p.K_in, 64)) != MIP360_OK) return rc;
  const int tiles = ((p.N_out + BM * CG - 1) / (BM * CG)) * (p.K_in / BN);
  const int kb_total = (p.M + BK - 1) / BK;
  // two waves of CTAs (CTA pairs) when there is enough reduction depth, at least 8 k-blocks per split
  int splits = (2 * (sm_count() / CG)) / tiles;
  if (splits < 1) splits = 1;
  if (splits > kb_total / 8) splits = kb_total / 8 > 0 ? kb_total / 8 : 1;
  p.splits = splits;
  if (CG == 1) {
    wgrad_kernel<BN, CG><<<tiles * splits, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tdy, tx, p);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(tiles * splits * CG);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MIP_CUDA(cudaLaunchKernelEx(&cfg, wgrad_kernel<BN, CG>, tdy, tx, p));
  }
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

}  // namespace mip360

using namespace mip360;

extern "C" {

int mip360_linear_fwd(const uint16_t* X, const uint16_t* W, const float* bias, int M, int N, int K, int act,
                      uint16_t* out_bf16, float* out_f32, int n_valid, mip360_stream_t stream) {
  MIP_REQUIRE(X && W && bias, "linear_fwd: null pointer");
  MIP_REQUIRE(out_bf16 || out_f32, "linear_fwd: no output");
  MIP_REQUIRE(M > 0 && N > 0 && K > 0 && K % BK == 0, "linear_fwd: bad shape M=%d N=%d K=%d (K %% 64 != 0?)", M, N, K);
  MIP_REQUIRE(!out_f32 || (n_valid >= 1 && n_valid <= 8), "linear_fwd: n_valid=%d outside [1,8]", n_valid);
  MIP_REQUIRE(act >= 0 && act <= 2, "linear_fwd: act=%d", act);
  // trunk Sigmoid with bf16-only output: single-MUFU tanh form; fp32 head outputs keep the exact form
  const int act_k = (act == ACT_SIGMOID && !out_f32) ? ACT_SIGMOID_FAST : act;
  LinearParams p{bias, out_bf16, out_f32, M, N, K, act_k, n_valid, option(OPT_PACKED_EPILOGUE) ? 1 : 0};
  return dispatch_linear<EPI_FWD>(X, W, nullptr, p, (cudaStream_t)stream);
}

int mip360_mlp_fwd_fused_narrow(const uint16_t* x, int M, const mip360_layer* trunk, int n_trunk, const mip360_layer* head,
                                int n_valid, uint16_t* const* acts, float* out, mip360_stream_t stream) {
  MIP_REQUIRE(trunk && head && out, "mlp_fwd_fused_narrow: null pointer");
  if (M <= 0) return MIP360_OK;
  MIP_REQUIRE(x, "mlp_fwd_fused_narrow: null input");
  bool ok = n_trunk == PF_TRUNK && head->n_pad == 64 && head->k_pad == PF_W && head->act == ACT_NONE && n_valid >= 1 &&
            n_valid <= 8 && option(OPT_FUSED_NARROW);
  for (int l = 0; ok && l < n_trunk; ++l)
    ok = trunk[l].n_pad == PF_W && trunk[l].k_pad == (l == 0 ? BK : PF_W) && (trunk[l].act == ACT_RELU || trunk[l].act == ACT_SIGMOID) &&
         trunk[l].W && trunk[l].bias;
  if (!ok) {
    set_error("mlp_fwd_fused_narrow: only 64 -> 256 x 4 (ReLU / Sigmoid) -> head without activation is fused");
    return MIP360_ERR_UNSUPPORTED;
  }
  if (acts)
    for (int l = 0; l < n_trunk; ++l) MIP_REQUIRE(acts[l], "mlp_fwd_fused_narrow: activation buffer %d is null", l);
  return launch_prop_fused(x, M, trunk, head, n_valid, acts, out, (cudaStream_t)stream);
}

int mip360_linear_fwd_head(const uint16_t* X, const uint16_t* W, const float* bias, int M, int N, int K, int act,
                           uint16_t* out_bf16, const float* head_w4, float* head_out, mip360_stream_t stream) {
  MIP_REQUIRE(X && W && bias && head_w4 && head_out, "linear_fwd_head: null pointer");
  MIP_REQUIRE(M > 0 && N > 0 && K > 0 && K % BK == 0, "linear_fwd_head: bad shape M=%d N=%d K=%d (K %% 64 != 0?)", M, N, K);
  MIP_REQUIRE(act >= 0 && act <= 2, "linear_fwd_head: act=%d", act);
  LinearParams p{bias, out_bf16, nullptr, M, N, K, act == ACT_SIGMOID ? ACT_SIGMOID_FAST : act, 0, 0, head_w4, head_out};
  return dispatch_linear<EPI_FWD_HEAD>(X, W, nullptr, p, (cudaStream_t)stream);
}

int mip360_head_bwd(const float* g, const float* head_w4, const uint16_t* Y, int M, int N, int act, uint16_t* dZ,
                    float* dWh, int ldw, float* dbh, mip360_stream_t stream) {
  MIP_REQUIRE(g && head_w4 && Y && dZ && dWh, "head_bwd: null pointer");
  MIP_REQUIRE(M > 0 && N >= 64 && N % 64 == 0 && N <= 2048 && ldw >= N, "head_bwd: bad shape M=%d N=%d ldw=%d", M, N, ldw);
  MIP_REQUIRE(act >= 0 && act <= 2, "head_bwd: act=%d", act);
  int grid = sm_count() * 4;
  const int rows_per_pass = 256 / (N / 8);
  const long long need = ((long long)M + rows_per_pass - 1) / rows_per_pass;
  if (grid > need) grid = (int)need;
  cudaStream_t st = (cudaStream_t)stream;
  if (act == ACT_RELU) head_bwd_kernel<ACT_RELU><<<grid, 256, 0, st>>>(g, head_w4, Y, M, N, dZ, dWh, ldw, dbh);
  else if (act == ACT_SIGMOID) head_bwd_kernel<ACT_SIGMOID><<<grid, 256, 0, st>>>(g, head_w4, Y, M, N, dZ, dWh, ldw, dbh);
  else head_bwd_kernel<ACT_NONE><<<grid, 256, 0, st>>>(g, head_w4, Y, M, N, dZ, dWh, ldw, dbh);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_linear_dgrad(const uint16_t* dY, const uint16_t* Wt, const uint16_t* Yprev, int M, int N, int K, int act,
                        uint16_t* dX, mip360_stream_t stream) {
  // GEMM view: out[M, K] = dY[M, N] * Wt[K, N]^T : reduction over N, output width K
  MIP_REQUIRE(dY && Wt && dX, "linear_dgrad: null pointer");
  MIP_REQUIRE(Yprev, "linear_dgrad: the saved activation output Yprev is required");
  MIP_REQUIRE(M > 0 && N > 0 && K > 0 && N % BK == 0, "linear_dgrad: bad shape M=%d N=%d K=%d (N %% 64 != 0?)", M, N, K);
  LinearParams p{nullptr, dX, nullptr, M, /*N=*/K, /*K=*/N, act, 0, option(OPT_PACKED_EPILOGUE) ? 1 : 0};
  return dispatch_linear<EPI_DGRAD>(dY, Wt, Yprev, p, (cudaStream_t)stream);
}

int mip360_linear_wgrad(const uint16_t* dY, const uint16_t* X, int M, int N, int K, float* dW, float* db,
                        mip360_stream_t stream) {
  MIP_REQUIRE(dY && X && dW, "linear_wgrad: null pointer");
  MIP_REQUIRE(M > 0 && N > 0 && K > 0 && N % 64 == 0 && K % 64 == 0, "linear_wgrad: bad shape M=%d N=%d K=%d", M, N, K);
  WgradParams p{dW, db, M, N, K, 1};
  const bool pair_ok = option(OPT_CTA_PAIR);
  if (K % 256 == 0 && N % 256 == 0 && N >= 512 && pair_ok) return launch_wgrad<256, 2>(dY, X, p, (cudaStream_t)stream);
  if (K % 256 == 0) return launch_wgrad<256, 1>(dY, X, p, (cudaStream_t)stream);
  if (K == 128) return launch_wgrad<128, 1>(dY, X, p, (cudaStream_t)stream);
  if (K == 64) return launch_wgrad<64, 1>(dY, X, p, (cudaStream_t)stream);
  set_error("linear_wgrad: K=%d not supported (need 64, 128 or a multiple of 256)", K);
  return MIP360_ERR_UNSUPPORTED;
}

int mip360_cast_weight(const float* W, int N, int K, int Npad, int Kpad, uint16_t* Wb, uint16_t* Wt,
                       mip360_stream_t stream) {
  MIP_REQUIRE(W && (Wb || Wt), "cast_weight: null pointer");
  MIP_REQUIRE(N > 0 && K > 0 && Npad >= N && Kpad >= K, "cast_weight: bad shape");
  MIP_REQUIRE(Npad % 32 == 0 && Kpad % 32 == 0, "cast_weight: padded shape [%d, %d] must be multiples of 32", Npad, Kpad);
  cast_weight_kernel<<<dim3(Kpad / 32, Npad / 32), 256, 0, (cudaStream_t)stream>>>(W, N, K, Npad, Kpad, Wb, Wt);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_adamw_pack(const mip360_pack_entry* entries, int n_entries, int total_tiles, float* p, float* g, float* m,
                      float* v, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                      const float* hyper_dev, int do_adam, int zero_grad, mip360_stream_t stream) {
  MIP_REQUIRE(entries && n_entries >= 1 && total_tiles >= 1, "adamw_pack: empty table");
  MIP_REQUIRE(!do_adam || (p && g && m && v && (hyper_dev || step >= 1)), "adamw_pack: AdamW needs p, g, m, v and a step");
  AdamHyper h{lr, beta1, beta2, eps, weight_decay, 1.f, 1.f};
  if (do_adam && !hyper_dev) {
    h.bc1 = 1.f - powf(beta1, (float)step);
    h.bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  }
  adamw_pack_kernel<<<total_tiles, 256, 0, (cudaStream_t)stream>>>(entries, n_entries, p, g, m, v, h, hyper_dev, do_adam,
                                                                   zero_grad);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                 float eps, float weight_decay, int step, mip360_stream_t stream) {
  MIP_REQUIRE(p && g && m && v && n > 0 && step >= 1, "adamw: bad arguments");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  int grid = (int)((n + 255) / 256);
  if (grid > sm_count() * 8) grid = sm_count() * 8;
  adamw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

}  // extern "C"
