// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA (128B swizzle) -> smem ring -> tcgen05.mma
// (cta_group::1, UMMA 128 x BN x 16, fp32 accumulators double-buffered in TMEM) -> tcgen05.ld epilogue
// with the Transfusion-specific fused epilogues.
//
//   D[m][n] = sum_k A(m,k) * B(n,k)
//   A "K-major":  stored row-major [M][K]   (activations as GEMM input, dY for dgrad)
//   A "MN-major": stored row-major [K][M]   (dY^T for wgrad: K = tokens)
//   B likewise over n.
//
// Roles (320 threads): warp 0 = TMA producer (1 lane), warp 1 = MMA issuer (1 lane) + TMEM owner,
// warps 2..9 = epilogue (warp%4 selects the TMEM lane quadrant, (warp-2)/4 the column half of the tile;
// thread <-> one accumulator row).  Two epilogue warps per SM sub-partition hide each other's latencies.
//
// Epilogue memory traffic is staged through a per-warp 32 x 128 B shared-memory tile (XOR-swizzled 16 B
// chunks) so that every global load/store instruction covers whole 64/128-byte row segments: the TMEM
// layout gives a thread one ROW (good for per-row math: qk-RMSNorm, RoPE, gate select), the staging turns
// that into coalesced accesses.  v0 wrote one 16 B piece per lane per row and capped every K=512 GEMM at
// ~390 TFLOP/s (profiles/r01_gemm_harness_v0.log).
#pragma once
#include "sm100_ptx.cuh"

namespace tfx {

#ifndef GEMM_ABLATE_HALF_B
#define GEMM_ABLATE_HALF_B 0
#endif
constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;     // 64 bf16 = 128 B = one swizzle atom row
constexpr int GEMM_UK = 16;     // UMMA K for 16-bit inputs

enum : int { EPI_STORE = 0, EPI_QKVG = 1, EPI_RESID = 2, EPI_GEGLU = 3 };

struct GemmParams {
  int M, N, K;                 // D is M x N, reduction K
  int k_splits;                // >1: split-K, fp32 atomic accumulate (EPI_STORE only)
  // ---- EPI_STORE: out = alpha*acc + bias[n]
  float* out_f32; long long ld_f32;
  __nv_bfloat16* out_bf16; long long ld_bf16;
  const float* bias;           // [N] or null
  const long long* row_off;    // optional per-output-row element offset into out_f32 (-1 = skip row); replaces m*ld_f32
  float alpha;
  int accumulate_f32;          // 1: out_f32 += (red.add)
  int K1;                      // A is the concatenation [A | A2] along K; A2 starts at k = K1 (K1 % 64 == 0); K1 = K when unused
  // ---- EPI_QKVG (N tile 128 = 2 heads: H/2 q tiles | H/2 k tiles | H/2 v tiles | 1 gate tile)
  int H;                       // heads (even), head dim 64
  __nv_bfloat16 *q, *k, *v;    // [M][H*64]
  float* gates;                // [M][H]   raw gate logits
  float* mix_pre;              // [M][H]   optional: columns [H, 2H) of the gate tile = pre-activation of the learned value-residual mix (T.py:956-960)
  float* qk_inv;               // [M][2H]  1/max(|x|,eps) for q heads then k heads
  const float *q_gamma, *k_gamma;   // [64]
  const int* rope_pos;         // [M]
  const float2* rope_cs;       // [32][rope_len] (cos, sin): transposed table, consecutive positions are contiguous
  int rope_len;
  const int* kv_rows;          // optional [M]: destination ROW of token m inside k / v (in-place kv-cache append: k, v then point at a
                               // layer's cache slabs and q stays dense); null = row m
  // ---- EPI_RESID: y = acc + bias; y_bf16 = y; x_out = x_res + y * scale(row, col)
  const float* x_res; float* x_out; __nv_bfloat16* x_out_bf16;     // [M][N]
  __nv_bfloat16* y_bf16;       // [M][N] optional (pre-scale branch output, saved for backward)
  const int* cond_row;         // [M]  >=0: modality token -> row of zgate; <0: text token
  const float* zgate;          // [n_cond][zgate_ld]  sigmoid(to_ada_ln_zero(cond))
  long long zgate_ld;
  const float* ls;             // [N] layerscale (scale = ls + 1 for text rows); null => scale = 1
  // ---- EPI_GEGLU (N tile 128 = [64 value cols | 64 gate cols], N = 2*inner_pad)
  __nv_bfloat16* vg;           // [M][N] pre-activation (value|gate interleaved per tile), saved for backward
  __nv_bfloat16* h;            // [M][N/2] gelu(gate)*value
};

// Epilogue organisation.  A tcgen05.ld gives a thread one accumulator ROW, and a warp may only touch the TMEM lane quadrant warp%4,
// so parallelism in the epilogue comes from several warps per quadrant splitting the tile's COLUMNS.  The fused epilogues (QKVG,
// RESID, GEGLU) are latency-bound with 8 warps (profiles/r01_ncu_gemm_epi_v2.txt: 10-45 % issue utilisation, 1 block / SM), so they
// run 16 epilogue warps (4 per quadrant, <= 112 registers / thread); the plain STORE epilogue (dgrad / wgrad) keeps 8.
template <int BN, int EPI = 0> struct GemmCfg {
  static constexpr int EW = (EPI == 0 || (EPI == 1 && BN == 128)) ? 8 : 16;                   // epilogue warps
  static constexpr int THREADS = 64 + 32 * EW;
  static constexpr int STAGES = (BN == 256 || EPI == 2) ? 4 : 6;
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // per-warp staging: 32 rows x 128 B (swizzled) for 8-warp kernels and RESID (fp32 rows; doubles as the residual prefetch buffer),
  // 32 rows x 64 B for the 16-warp bf16 epilogues of the 256-wide kernels (smem is needed for the 4-stage operand ring there)
  static constexpr int STG_WARP = EW == 8 ? 4096 : (EPI == 2 ? 4096 + 2048 : 2048);   // RESID: residual / fp32 tile (4 KB) + bf16 tile (2 KB)
  static constexpr int STAGING = EW * STG_WARP;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 2 * BN;   // double-buffered accumulator (power of two: 256 / 512)
};

// Phi(g) and phi(g) of the exact-erf GELU (T.py:831-834, F.gelu default) from ONE exponential: Abramowitz-Stegun 7.1.26,
// erf(x) = 1 - (a1 t + .. + a5 t^5) e^{-x^2}, t = 1/(1 + p x), |err| <= 1.5e-7 - far below the bf16 rounding of the outputs.
// ~14 FMA-pipe instructions + 2 MUFU instead of the ~50 of erff + expf: the GEGLU epilogues are instruction-bound.
__device__ __forceinline__ void gelu_parts(float g, float& cdf, float& pdf) {
  const float ax = fabsf(g) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, ax, 1.f));
  const float E = __expf(-ax * ax);
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float h = 0.5f * poly * t * E;          // 0.5 * (1 - erf(|g|/sqrt 2))
  cdf = g >= 0.f ? 1.f - h : h;
  pdf = 0.3989422804014327f * E;
}
__device__ __forceinline__ float gelu_erf(float x) { float c, d; gelu_parts(x, c, d); return x * c; }
// two elements at once on the packed fp32x2 pipe (FFMA2 / FMUL2 / FADD2): v * g * Phi(g).  Same formula as gelu_parts, the 0.5 folded into the
// polynomial; 21 instructions per PAIR instead of ~18 per element - the GEGLU epilogue is issue-bound (profiles/r02_gemm_pair_experiments.txt).
__device__ __forceinline__ float2 geglu_pair(float2 g, float2 v) {
  const float2 ax = make_float2(fabsf(g.x) * 0.70710678118654752f, fabsf(g.y) * 0.70710678118654752f);
  const float2 den = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), ax, make_float2(1.f, 1.f));
  const float2 t = make_float2(__fdividef(1.f, den.x), __fdividef(1.f, den.y));
  const float2 q = __fmul2_rn(__fmul2_rn(ax, ax), make_float2(-1.4426950408889634f, -1.4426950408889634f));
  const float2 E = make_float2(exp2f(q.x), exp2f(q.y));
  float2 poly = __ffma2_rn(make_float2(0.5f * 1.061405429f, 0.5f * 1.061405429f), t, make_float2(0.5f * -1.453152027f, 0.5f * -1.453152027f));
  poly = __ffma2_rn(poly, t, make_float2(0.5f * 1.421413741f, 0.5f * 1.421413741f));
  poly = __ffma2_rn(poly, t, make_float2(0.5f * -0.284496736f, 0.5f * -0.284496736f));
  poly = __ffma2_rn(poly, t, make_float2(0.5f * 0.254829592f, 0.5f * 0.254829592f));
  const float2 h = __fmul2_rn(__fmul2_rn(poly, t), E);                       // 0.5 * (1 - erf(|g| / sqrt 2))
  const float2 sg = make_float2(copysignf(1.f, g.x), copysignf(1.f, g.y));
  const float2 cdf = __ffma2_rn(sg, __fadd2_rn(make_float2(0.5f, 0.5f), make_float2(-h.x, -h.y)), make_float2(0.5f, 0.5f));      // g >= 0 ? 1 - h : h
  return __fmul2_rn(__fmul2_rn(g, cdf), v);
}

__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(sz) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack2_bf16_(uint32_t w) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&w);
  return __bfloat1622float2(t);
}

// ---- per-warp staging tile: 32 rows x 128 B, 16-byte chunk c of row r lives at r*128 + ((c ^ (r & 7)) << 4)
// lane == row when writing/reading "own row"; (row, chunk) = f(iteration, lane) when touching global memory.
template <int CH>   // chunks (16 B) per row actually used: 8 (32 fp32 / 64 bf16) or 4 (32 bf16)
__device__ __forceinline__ void stg_put(uint8_t* sw, int lane, const uint32_t* w) {
#pragma unroll
  for (int c = 0; c < CH; ++c)
    *reinterpret_cast<uint4*>(sw + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
}
template <int CH>
__device__ __forceinline__ void stg_get(const uint8_t* sw, int lane, uint32_t* w) {
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const uint4 t = *reinterpret_cast<const uint4*>(sw + lane * 128 + ((c ^ (lane & 7)) << 4));
    w[4 * c] = t.x; w[4 * c + 1] = t.y; w[4 * c + 2] = t.z; w[4 * c + 3] = t.w;
  }
}
// coalesced global store of the staged tile: g points at (tile row 0, first column); pitch in bytes
template <int CH>
__device__ __forceinline__ void stg_store(const uint8_t* sw, int lane, uint8_t* g, long long pitch, int rows_valid) {
  constexpr int RPI = 32 / CH;
#pragma unroll
  for (int it = 0; it < 32 / RPI; ++it) {
    const int row = it * RPI + lane / CH, ch = lane % CH;
    if (row < rows_valid)
      *reinterpret_cast<uint4*>(g + row * pitch + ch * 16) = *reinterpret_cast<const uint4*>(sw + row * 128 + ((ch ^ (row & 7)) << 4));
  }
}
template <int CH>
__device__ __forceinline__ void stg_load(uint8_t* sw, int lane, const uint8_t* g, long long pitch, int rows_valid) {
  constexpr int RPI = 32 / CH;
#pragma unroll
  for (int it = 0; it < 32 / RPI; ++it) {
    const int row = it * RPI + lane / CH, ch = lane % CH;
    uint4 t = make_uint4(0, 0, 0, 0);
    if (row < rows_valid) t = *reinterpret_cast<const uint4*>(g + row * pitch + ch * 16);
    *reinterpret_cast<uint4*>(sw + row * 128 + ((ch ^ (row & 7)) << 4)) = t;
  }
}

// 32 fp32 values of the lane's row -> bf16 -> staging tile (128 B pitch, 4 chunks), packed chunk by chunk
__device__ __forceinline__ void stg_put_pack(uint8_t* sw, int lane, const float* y) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(sw + lane * 128 + ((c ^ (lane & 7)) << 4)) =
        make_uint4(pack_bf16(y[8 * c], y[8 * c + 1]), pack_bf16(y[8 * c + 2], y[8 * c + 3]), pack_bf16(y[8 * c + 4], y[8 * c + 5]), pack_bf16(y[8 * c + 6], y[8 * c + 7]));
}
// compact variant for 32 bf16 per row: 32 rows x 64 B, chunk c of row r at r*64 + ((c ^ ((r >> 1) & 3)) << 4)  (conflict-free both ways)
__device__ __forceinline__ void stg64_put(uint8_t* sw, int lane, const uint32_t* w) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(sw + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
}
__device__ __forceinline__ void stg64_put_pack(uint8_t* sw, int lane, const float* y) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(sw + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) =
        make_uint4(pack_bf16(y[8 * c], y[8 * c + 1]), pack_bf16(y[8 * c + 2], y[8 * c + 3]), pack_bf16(y[8 * c + 4], y[8 * c + 5]), pack_bf16(y[8 * c + 6], y[8 * c + 7]));
}
__device__ __forceinline__ void stg64_store(const uint8_t* sw, int lane, uint8_t* g, long long pitch, int rows_valid) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int row = it * 8 + (lane >> 2), ch = lane & 3;
    if (row < rows_valid)
      *reinterpret_cast<uint4*>(g + row * pitch + ch * 16) = *reinterpret_cast<const uint4*>(sw + row * 64 + ((ch ^ ((row >> 1) & 3)) << 4));
  }
}

// row-mapped variants: slab row r of the warp goes to global row rows[r] (rows already offset to the warp's first row)
__device__ __forceinline__ void stg64_store_rows(const uint8_t* sw, int lane, uint8_t* g, long long pitch, int rows_valid, const int* __restrict__ rows) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int row = it * 8 + (lane >> 2), ch = lane & 3;
    if (row < rows_valid)
      *reinterpret_cast<uint4*>(g + (long long)rows[row] * pitch + ch * 16) = *reinterpret_cast<const uint4*>(sw + row * 64 + ((ch ^ ((row >> 1) & 3)) << 4));
  }
}
template <int CH>
__device__ __forceinline__ void stg_store_rows(const uint8_t* sw, int lane, uint8_t* g, long long pitch, int rows_valid, const int* __restrict__ rows) {
  constexpr int RPI = 32 / CH;
#pragma unroll
  for (int it = 0; it < 32 / RPI; ++it) {
    const int row = it * RPI + lane / CH, ch = lane % CH;
    if (row < rows_valid)
      *reinterpret_cast<uint4*>(g + (long long)rows[row] * pitch + ch * 16) = *reinterpret_cast<const uint4*>(sw + row * 128 + ((ch ^ (row & 7)) << 4));
  }
}

// CL = 2: a CTA PAIR (cluster of two, tcgen05 cta_group::2) computes a 256 x BN tile.  Each CTA loads its own 128 rows of A and only HALF of the B
// tile (BN / 2 rows); the leader's MMA issuer runs 256 x BN x 16 UMMAs that read both shared memories and write 128 x BN accumulators into each
// CTA's TMEM, so each CTA's epilogue is unchanged.  TMA completion bytes of both CTAs are counted on the leader's `full` barrier; commits are
// multicast to both CTAs' `empty` / `tfull` barriers; both epilogues release the accumulator on the leader's `tempty` barrier.
// A pair ingests 32 KB per k-block and SM instead of 48.  Measured (profiles/r02_gemm_pair_experiments.txt): correct, and NOT faster on this model's
// shapes - operand ingest is not what bounds them - so it is an option (tfx_gemm_set_cluster_mode), off by default.
template <int BN, bool A_MN, bool B_MN, int EPI, int CL = 1>
__global__ void __launch_bounds__((GemmCfg<BN, EPI>::THREADS), 1)
gemm_sm100_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN, EPI>;
  // a CTA pair keeps only half of the B tile per stage: the same ring memory holds more, smaller stages (BN 256: 6 x 32 KB instead of 4 x 48 KB)
  constexpr int STAGE_STRIDE = CL == 1 ? Cfg::STAGE_BYTES : Cfg::A_BYTES + Cfg::B_BYTES / 2;
  constexpr int STAGES = CL == 1 ? Cfg::STAGES : (Cfg::STAGES * Cfg::STAGE_BYTES) / STAGE_STRIDE;
  static_assert(STAGE_STRIDE % 1024 == 0 && 2 * STAGES + 5 <= 32, "stage alignment / barrier area");
  // declared 1024-byte aligned (SWIZZLE_128B tiles) and used directly: rounding the pointer up through an integer loses the shared address space
  // and turned every staging access of the epilogues into a generic LD / ST (profiles/r02_gemm_pair_experiments.txt: "ST.E.128 desc[..]")
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) { printf("tfx gemm: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
  constexpr int EW = Cfg::EW;
  uint8_t* staging = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING);
  uint64_t* full_bar = bars;                  // [STAGES]
  uint64_t* empty_bar = bars + STAGES;        // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;    // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int kb_total = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int kb_per_split = (kb_total + p.k_splits - 1) / p.k_splits;
  const int m_units = (m_tiles + CL - 1) / CL;            // a work item = CL vertically adjacent tiles, one per CTA of the cluster
  const int unit_items = m_units * n_tiles;
  const int num_items = unit_items * p.k_splits;
  const int cta_rank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int first_item = blockIdx.x / CL, item_stride = gridDim.x / CL;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], CL * EW); }      // (pairs: only the leader's full / tempty barriers are used)
    mbar_fence_init();
  }
  if (warp == 1) {
    if constexpr (CL == 1) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
    else { tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish_pair(); }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();               // the peer's barriers and TMEM are set up before anything is signalled at them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = first_item; item < num_items; item += item_stride) {
        const int split = item / unit_items;
        const int rem = item - split * unit_items;
        const int m_unit = rem / n_tiles, n_blk = rem - m_unit * n_tiles;
        const int m_blk = m_unit * CL + cta_rank;             // (past the last tile for an odd tile count: zero-filled by TMA, results discarded)
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * STAGE_STRIDE;
          uint8_t* sB = sA + Cfg::A_BYTES;
          if constexpr (CL == 1) {
#if GEMM_ABLATE_HALF_B      // timing experiment only (wrong results): the SM ingests half of the B tile, as a CTA pair does
            mbar_expect_tx(&full_bar[stage], Cfg::A_BYTES + Cfg::B_BYTES / 2);
#else
            mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
#endif
            if (!A_MN) {
              if (kb * GEMM_BK < p.K1) tma_load_2d(&tmA, &full_bar[stage], sA, kb * GEMM_BK, m_blk * GEMM_BM);
              else tma_load_2d(&tmA2, &full_bar[stage], sA, kb * GEMM_BK - p.K1, m_blk * GEMM_BM);
            } else {
#pragma unroll
              for (int a = 0; a < GEMM_BM / 64; ++a)
                tma_load_2d(&tmA, &full_bar[stage], sA + a * (GEMM_BK * 128), m_blk * GEMM_BM + a * 64, kb * GEMM_BK);
            }
            if (!B_MN) {
              tma_load_2d(&tmB, &full_bar[stage], sB, kb * GEMM_BK, n_blk * BN);
            } else {
#pragma unroll
              for (int a = 0; a < BN / (GEMM_ABLATE_HALF_B ? 128 : 64); ++a)
                tma_load_2d(&tmB, &full_bar[stage], sB + a * (GEMM_BK * 128), n_blk * BN + a * 64, kb * GEMM_BK);
            }
          } else {
            // pair: this CTA's A rows and its half of the B tile; all bytes are counted on the leader's barrier
            const uint32_t lead_full = cluster_map(&full_bar[stage], 0);
            if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * (Cfg::A_BYTES + Cfg::B_BYTES / 2));
            if (!A_MN) {
              if (kb * GEMM_BK < p.K1) tma_load_2d_pair(&tmA, lead_full, sA, kb * GEMM_BK, m_blk * GEMM_BM);
              else tma_load_2d_pair(&tmA2, lead_full, sA, kb * GEMM_BK - p.K1, m_blk * GEMM_BM);
            } else {
#pragma unroll
              for (int a = 0; a < GEMM_BM / 64; ++a)
                tma_load_2d_pair(&tmA, lead_full, sA + a * (GEMM_BK * 128), m_blk * GEMM_BM + a * 64, kb * GEMM_BK);
            }
            if (!B_MN) {
              tma_load_2d_pair(&tmB, lead_full, sB, kb * GEMM_BK, n_blk * BN + cta_rank * (BN / 2));        // tmB box: BN / 2 rows
            } else {
#pragma unroll
              for (int a = 0; a < BN / 128; ++a)
                tma_load_2d_pair(&tmB, lead_full, sB + a * (GEMM_BK * 128), n_blk * BN + (cta_rank * (BN / 128) + a) * 64, kb * GEMM_BK);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0 && cta_rank == 0) {          // (pairs: the leader issues for both SMs)
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM * CL, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0; uint32_t phase = 0;
      int local = 0;
      for (int item = first_item; item < num_items; item += item_stride, ++local) {
        const int split = item / unit_items;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, kb_total);
        const int buf = local & 1;
        const uint32_t bphase = (local >> 1) & 1;
        mbar_wait(&tempty_bar[buf], bphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * STAGE_STRIDE);
          const uint32_t sB = sA + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / GEMM_UK; ++k) {
            const uint64_t da = A_MN ? umma_smem_desc_sw128(sA + k * (GEMM_UK * 128), GEMM_BK * 128, 1024)
                                     : umma_smem_desc_sw128(sA + k * (GEMM_UK * 2), 0, 1024);
            const uint64_t db = B_MN ? umma_smem_desc_sw128(sB + k * (GEMM_UK * 128), GEMM_BK * 128, 1024)
                                     : umma_smem_desc_sw128(sB + k * (GEMM_UK * 2), 0, 1024);
            if constexpr (CL == 1) umma_bf16_ss(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_bf16_ss_pair(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if constexpr (CL == 1) umma_commit(&empty_bar[stage]);          // frees the smem slot when these MMAs retire
          else umma_commit_pair(&empty_bar[stage], (uint16_t)3);          // ... in both CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if constexpr (CL == 1) umma_commit(&tfull_bar[buf]);              // accumulator ready for the epilogue
        else umma_commit_pair(&tfull_bar[buf], (uint16_t)3);
      }
    }
  } else {
    // ===================================================== epilogue warps (2..9)
    const int quad = warp & 3;
    const int part = (warp - 2) >> 2;        // column slice of the tile handled by this warp (EW / 4 slices)
    const int half = part;                   // 8-warp kernels: two column halves
    uint8_t* sw = staging + (warp - 2) * Cfg::STG_WARP;
    // EPI_RESID: the 32 x 32 fp32 residual slab of a tile is fetched into the warp's 4 KB tile by cp.async one tile AHEAD (issued as soon
    // as the current slab has been pulled into registers), so its DRAM latency hides behind the current tile's math and stores
    auto resid_prefetch = [&](int it_) {
      if constexpr (EPI == EPI_RESID) {
        const int rem_ = it_ % unit_items;
        const int mb_ = (rem_ / n_tiles) * CL + cta_rank, nb_ = rem_ % n_tiles;
        const int wrow_ = mb_ * GEMM_BM + quad * 32, cb_ = nb_ * BN + part * 32;
        const int rv_ = min(32, p.M - wrow_);
        if (cb_ < p.N) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + (lane >> 3), ch = lane & 7;
            const bool ok = rr < rv_;
            cp_async16_zfill(sw + rr * 128 + ((ch ^ (rr & 7)) << 4), p.x_res + (long long)(ok ? wrow_ + rr : 0) * p.N + cb_ + ch * 4, ok);
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
    };
    int local = 0;
    for (int item = first_item; item < num_items; item += item_stride, ++local) {
      const int split = item / unit_items;
      const int rem = item - split * unit_items;
      const int m_unit = rem / n_tiles, n_blk = rem - m_unit * n_tiles;
      const int m_blk = m_unit * CL + cta_rank;               // a tile past the end has rows_valid <= 0: every store below is row-guarded
      const int buf = local & 1;
      const uint32_t bphase = (local >> 1) & 1;
      const int wrow0 = m_blk * GEMM_BM + quad * 32;          // first row of this warp's 32-row slab
      const int row = wrow0 + lane;
      const bool row_ok = row < p.M;
      const int rows_valid = min(32, p.M - wrow0);            // may be <= 0
      const int col0 = n_blk * BN;
      int qk_pos = 0;
      if constexpr (EPI == EPI_RESID) {
        if (local == 0) resid_prefetch(item);         // later tiles were prefetched while the previous tile was being processed
      }
      if constexpr (EPI == EPI_QKVG) { qk_pos = row_ok ? p.rope_pos[row] : 0; }
      if constexpr (EPI == EPI_RESID) { qk_pos = (row_ok && p.cond_row) ? p.cond_row[row] : -1; }      // (reused as the condition row)
      mbar_wait(&tfull_bar[buf], bphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + buf * BN;
      if constexpr (EPI == EPI_RESID) { asm volatile("cp.async.wait_group 0;" ::: "memory"); __syncwarp(); }

      if constexpr (EPI == EPI_STORE) {
        const bool f32_staged = p.out_f32 && (p.row_off || (p.ld_f32 & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out_f32) & 15) == 0);
        const bool bf16_staged = p.out_bf16 && (p.ld_bf16 & 7) == 0 && ((reinterpret_cast<uintptr_t>(p.out_bf16) & 15) == 0);
#pragma unroll 1
        for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
          const int cbase = col0 + c * 32;
          if (cbase >= p.N) break;
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + c * 32, r);
          tmem_ld_wait();
          const bool full = cbase + 32 <= p.N;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (p.alpha != 1.f) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
          }
          if (p.bias) {
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) { const float4 b = *reinterpret_cast<const float4*>(p.bias + cbase + j); v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w; }
            } else {
              for (int j = 0; j < 32; ++j) if (cbase + j < p.N) v[j] += p.bias[cbase + j];
            }
          }
          if (p.out_f32) {
            if (full && f32_staged) {
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(v[j]);
              stg_put<8>(sw, lane, r);
              __syncwarp();
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const int rr = it * 4 + (lane >> 3), ch = lane & 7;
                if (rr < rows_valid) {
                  const long long off = p.row_off ? p.row_off[wrow0 + rr] : (long long)(wrow0 + rr) * p.ld_f32;
                  if (off >= 0) {
                    float* dst = p.out_f32 + off + cbase + ch * 4;
                    const float4 t = *reinterpret_cast<const float4*>(sw + rr * 128 + ((ch ^ (rr & 7)) << 4));
                    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                      if (p.accumulate_f32) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
                      else *reinterpret_cast<float4*>(dst) = t;
                    } else {
                      if (p.accumulate_f32) { atomicAdd(dst, t.x); atomicAdd(dst + 1, t.y); atomicAdd(dst + 2, t.z); atomicAdd(dst + 3, t.w); }
                      else { dst[0] = t.x; dst[1] = t.y; dst[2] = t.z; dst[3] = t.w; }
                    }
                  }
                }
              }
              __syncwarp();
            } else if (row_ok) {
              const long long off = p.row_off ? p.row_off[row] : (long long)row * p.ld_f32;
              if (off >= 0) {
                float* dst = p.out_f32 + off + cbase;
                for (int j = 0; j < 32; ++j)
                  if (cbase + j < p.N) { if (p.accumulate_f32) atomicAdd(dst + j, v[j]); else dst[j] = v[j]; }
              }
            }
          }
          if (p.out_bf16) {
            if (full && bf16_staged) {
              uint32_t w[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) w[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
              stg_put<4>(sw, lane, w);
              __syncwarp();
              stg_store<4>(sw, lane, reinterpret_cast<uint8_t*>(p.out_bf16 + (long long)wrow0 * p.ld_bf16 + cbase), p.ld_bf16 * 2, rows_valid);
              __syncwarp();
            } else if (row_ok) {
              __nv_bfloat16* dst = p.out_bf16 + (long long)row * p.ld_bf16 + cbase;
              for (int j = 0; j < 32; ++j) if (cbase + j < p.N) dst[j] = __float2bfloat16(v[j]);
            }
          }
        }
      } else if constexpr (EPI == EPI_QKVG && BN == 256) {
        // 256-wide tile = 4 heads, one head (64 accumulator columns) per warp of the quadrant
        const int tps = p.H >> 2;             // tiles per section (H % 4 == 0)
        const int kind = n_blk / tps;         // 0 q, 1 k, 2 v, 3 gates
        const int tis = n_blk - kind * tps;
        const long long HI = (long long)p.H * 64;
        const uint32_t tcol = taddr + part * 64;
        if (kind <= 1) {
          const float* gamma = kind == 0 ? p.q_gamma : p.k_gamma;
          __nv_bfloat16* dstm = kind == 0 ? p.q : p.k;
          const float2* cs = p.rope_cs + qk_pos;        // entry i of this row's position: cs[i * rope_len]
          const int head = tis * 4 + part;
          float ss = 0.f;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {    // pass 1: |x|^2 over the head (TMEM reads are cheap: the row is re-read in pass 2)
            uint32_t r[32];
            tmem_ld_32x32b_x32(tcol + hf * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) { const float a = __uint_as_float(r[j]); ss += a * a; }
          }
          const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
          if (row_ok) p.qk_inv[(long long)row * 2 * p.H + kind * p.H + head] = inv;
          const float sc = inv * 8.f;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {    // pass 2: normalise, gamma, RoPE, bf16
            uint32_t r[32], outw[16];
            tmem_ld_32x32b_x32(tcol + hf * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float2 gm = *reinterpret_cast<const float2*>(gamma + hf * 32 + 2 * i);
              const float y0 = __uint_as_float(r[2 * i]) * sc * (gm.x + 1.f);
              const float y1 = __uint_as_float(r[2 * i + 1]) * sc * (gm.y + 1.f);
              const float2 cc = cs[(long long)(hf * 16 + i) * p.rope_len];
              outw[i] = pack_bf16(y0 * cc.x - y1 * cc.y, y1 * cc.x + y0 * cc.y);
            }
            stg64_put(sw, lane, outw);
            __syncwarp();
            if (kind == 1 && p.kv_rows) stg64_store_rows(sw, lane, reinterpret_cast<uint8_t*>(dstm + head * 64 + hf * 32), HI * 2, rows_valid, p.kv_rows + wrow0);
            else stg64_store(sw, lane, reinterpret_cast<uint8_t*>(dstm + (long long)wrow0 * HI + head * 64 + hf * 32), HI * 2, rows_valid);
            __syncwarp();
          }
        } else if (kind == 2) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t r[32], w[16];
            tmem_ld_32x32b_x32(tcol + hf * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) w[j] = pack_bf16(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
            stg64_put(sw, lane, w);
            __syncwarp();
            if (p.kv_rows) stg64_store_rows(sw, lane, reinterpret_cast<uint8_t*>(p.v + tis * 256 + part * 64 + hf * 32), HI * 2, rows_valid, p.kv_rows + wrow0);
            else stg64_store(sw, lane, reinterpret_cast<uint8_t*>(p.v + (long long)wrow0 * HI + tis * 256 + part * 64 + hf * 32), HI * 2, rows_valid);
            __syncwarp();
          }
        } else if (part == 0) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr, r);
          tmem_ld_wait();
          if (row_ok) {
            float* dst = p.gates + (long long)row * p.H;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < p.H) dst[j] = __uint_as_float(r[j]);
            if (p.mix_pre) {
              float* dm = p.mix_pre + (long long)row * p.H;
#pragma unroll
              for (int j = 0; j < 32; ++j) if (j >= p.H && j < 2 * p.H) dm[j - p.H] = __uint_as_float(r[j]);
            }
          }
        }
      } else if constexpr (EPI == EPI_QKVG) {
        static_assert(EPI != EPI_QKVG || BN == 128 || BN == 256, "QKVG epilogue expects 128- or 256-wide N tiles");
        const int tps = p.H >> 1;             // tiles per section
        const int kind = n_blk / tps;         // 0 q, 1 k, 2 v, 3 gates
        const int tis = n_blk - kind * tps;   // tile in section
        const long long HI = (long long)p.H * 64;
        if (kind <= 1) {
          const float* gamma = kind == 0 ? p.q_gamma : p.k_gamma;
          __nv_bfloat16* dstm = kind == 0 ? p.q : p.k;
          const float2* cs = p.rope_cs + qk_pos;        // entry i of this row's position: cs[i * rope_len]
          {
            const int hh = half;
            uint32_t r0[32], r1[32];
            tmem_ld_32x32b_x32(taddr + hh * 64, r0);
            tmem_ld_32x32b_x32(taddr + hh * 64 + 32, r1);
            tmem_ld_wait();
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) { float a = __uint_as_float(r0[j]), b = __uint_as_float(r1[j]); ss += a * a + b * b; }
            const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
            const int head = tis * 2 + hh;
            if (row_ok) p.qk_inv[(long long)row * 2 * p.H + kind * p.H + head] = inv;
            const float sc = inv * 8.f;
            uint32_t outw[32];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t* rr = half == 0 ? r0 : r1;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int d0 = half * 32 + 2 * i;
                const float2 gm = *reinterpret_cast<const float2*>(gamma + d0);
                const float y0 = __uint_as_float(rr[2 * i]) * sc * (gm.x + 1.f);
                const float y1 = __uint_as_float(rr[2 * i + 1]) * sc * (gm.y + 1.f);
                const float2 cc = cs[(long long)(half * 16 + i) * p.rope_len];
                outw[half * 16 + i] = pack_bf16(y0 * cc.x - y1 * cc.y, y1 * cc.x + y0 * cc.y);
              }
            }
            stg_put<8>(sw, lane, outw);
            __syncwarp();
            if (kind == 1 && p.kv_rows) stg_store_rows<8>(sw, lane, reinterpret_cast<uint8_t*>(dstm + head * 64), HI * 2, rows_valid, p.kv_rows + wrow0);
            else stg_store<8>(sw, lane, reinterpret_cast<uint8_t*>(dstm + (long long)wrow0 * HI + head * 64), HI * 2, rows_valid);
            __syncwarp();
          }
        } else if (kind == 2) {
          {
            const int c = half;
            uint32_t r0[32], r1[32], w[32];
            tmem_ld_32x32b_x32(taddr + c * 64, r0);
            tmem_ld_32x32b_x32(taddr + c * 64 + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              w[j] = pack_bf16(__uint_as_float(r0[2 * j]), __uint_as_float(r0[2 * j + 1]));
              w[16 + j] = pack_bf16(__uint_as_float(r1[2 * j]), __uint_as_float(r1[2 * j + 1]));
            }
            stg_put<8>(sw, lane, w);
            __syncwarp();
            if (p.kv_rows) stg_store_rows<8>(sw, lane, reinterpret_cast<uint8_t*>(p.v + tis * 128 + c * 64), HI * 2, rows_valid, p.kv_rows + wrow0);
            else stg_store<8>(sw, lane, reinterpret_cast<uint8_t*>(p.v + (long long)wrow0 * HI + tis * 128 + c * 64), HI * 2, rows_valid);
            __syncwarp();
          }
        } else if (half == 0) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr, r);
          tmem_ld_wait();
          if (row_ok) {
            float* dst = p.gates + (long long)row * p.H;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < p.H) dst[j] = __uint_as_float(r[j]);
            if (p.mix_pre) {
              float* dm = p.mix_pre + (long long)row * p.H;
#pragma unroll
              for (int j = 0; j < 32; ++j) if (j >= p.H && j < 2 * p.H) dm[j - p.H] = __uint_as_float(r[j]);
            }
          }
        }
      } else if constexpr (EPI == EPI_RESID) {
        static_assert(EPI != EPI_RESID || BN == 128, "RESID epilogue expects 128-wide N tiles (4 x 32-column slices)");
        const int crow = qk_pos;                        // fetched before the accumulator wait
        const float* zrow = (p.zgate && crow >= 0) ? p.zgate + (long long)crow * p.zgate_ld : nullptr;
        const int cbase = col0 + part * 32;
        if (cbase < p.N) {                              // N is a multiple of 32 for every RESID use
          uint32_t r[32];
          uint8_t* swb = sw + 4096;                     // 64-byte-pitch tile for the bf16 outputs
          tmem_ld_32x32b_x32(taddr + part * 32, r);
          tmem_ld_wait();
          float y[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = __uint_as_float(r[j]);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) { const float4 b = *reinterpret_cast<const float4*>(p.bias + cbase + j); y[j] += b.x; y[j + 1] += b.y; y[j + 2] += b.z; y[j + 3] += b.w; }
          }
          if (p.y_bf16) {
            stg64_put_pack(swb, lane, y);
            __syncwarp();
            stg64_store(swb, lane, reinterpret_cast<uint8_t*>(p.y_bf16 + (long long)wrow0 * p.N + cbase), (long long)p.N * 2, rows_valid);
            __syncwarp();
          }
          {
            uint32_t xrv[32];
            stg_get<8>(sw, lane, xrv);                  // residual row (prefetched by cp.async one tile ahead)
            __syncwarp();                               // every lane holds its row: the 4 KB tile is free for the NEXT tile's slab
            if (item + item_stride < num_items) resid_prefetch(item + item_stride);
            if (zrow) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 s4 = *reinterpret_cast<const float4*>(zrow + cbase + j);
                y[j] = __uint_as_float(xrv[j]) + y[j] * s4.x; y[j + 1] = __uint_as_float(xrv[j + 1]) + y[j + 1] * s4.y;
                y[j + 2] = __uint_as_float(xrv[j + 2]) + y[j + 2] * s4.z; y[j + 3] = __uint_as_float(xrv[j + 3]) + y[j + 3] * s4.w;
              }
            } else if (p.ls) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 s4 = *reinterpret_cast<const float4*>(p.ls + cbase + j);
                y[j] = __uint_as_float(xrv[j]) + y[j] * (s4.x + 1.f); y[j + 1] = __uint_as_float(xrv[j + 1]) + y[j + 1] * (s4.y + 1.f);
                y[j + 2] = __uint_as_float(xrv[j + 2]) + y[j + 2] * (s4.z + 1.f); y[j + 3] = __uint_as_float(xrv[j + 3]) + y[j + 3] * (s4.w + 1.f);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) y[j] += __uint_as_float(xrv[j]);
            }
          }
          // x_out (fp32, optional when only the bf16 copy is kept) leaves in two 16-column passes through the 64-byte-pitch tile
          if (p.x_out)
#pragma unroll
          for (int hfc = 0; hfc < 2; ++hfc) {
            uint32_t w16[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) w16[j] = __float_as_uint(y[hfc * 16 + j]);
            __syncwarp();
            stg64_put(swb, lane, w16);
            __syncwarp();
            stg64_store(swb, lane, reinterpret_cast<uint8_t*>(p.x_out + (long long)wrow0 * p.N + cbase + hfc * 16), (long long)p.N * 4, rows_valid);
          }
          if (p.x_out_bf16) {
            __syncwarp();
            stg64_put_pack(swb, lane, y);
            __syncwarp();
            stg64_store(swb, lane, reinterpret_cast<uint8_t*>(p.x_out_bf16 + (long long)wrow0 * p.N + cbase), (long long)p.N * 2, rows_valid);
          }
          __syncwarp();
        } else if (item + item_stride < num_items) {
          resid_prefetch(item + item_stride);             // keep the one-ahead commit-group bookkeeping uniform
        }
      } else if constexpr (EPI == EPI_GEGLU) {
        static_assert(EPI != EPI_GEGLU || BN == 256, "GEGLU epilogue expects 256-wide N tiles (2 x [64 value | 64 gate])");
        const int pair = part >> 1, c = part & 1;
        const int cv = col0 + pair * 128 + c * 32, cg = cv + 64;
        if (cv < p.N) {                                 // N is a multiple of 128: a pair is either complete or absent
          // Three short passes (value -> vg, gate -> vg, value x gelu(gate) -> h): re-reading the accumulator from TMEM is cheaper
          // than keeping 64 fp32 + 48 packed words live in a 96-register budget.
          const uint32_t tv = taddr + pair * 128 + c * 32, tg = tv + 64;
          uint32_t r[32];
          float g[32];
          tmem_ld_32x32b_x32(tv, r);
          tmem_ld_wait();
          {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; j += 2) { const float2 t2 = __fadd2_rn(make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), *reinterpret_cast<const float2*>(p.bias + cv + j)); v[j] = t2.x; v[j + 1] = t2.y; }
            stg64_put_pack(sw, lane, v);
          }
          __syncwarp();
          stg64_store(sw, lane, reinterpret_cast<uint8_t*>(p.vg + (long long)wrow0 * p.N + cv), (long long)p.N * 2, rows_valid);
          tmem_ld_32x32b_x32(tg, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 2) { const float2 t2 = __fadd2_rn(make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), *reinterpret_cast<const float2*>(p.bias + cg + j)); g[j] = t2.x; g[j + 1] = t2.y; }
          __syncwarp();
          stg64_put_pack(sw, lane, g);
          __syncwarp();
          stg64_store(sw, lane, reinterpret_cast<uint8_t*>(p.vg + (long long)wrow0 * p.N + cg), (long long)p.N * 2, rows_valid);
          tmem_ld_32x32b_x32(tv, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 bv = *reinterpret_cast<const float2*>(p.bias + cv + j);
            const float2 o = geglu_pair(make_float2(g[j], g[j + 1]), __fadd2_rn(make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), bv));
            g[j] = o.x; g[j + 1] = o.y;
          }
          __syncwarp();
          stg64_put_pack(sw, lane, g);
          __syncwarp();
          stg64_store(sw, lane, reinterpret_cast<uint8_t*>(p.h + (long long)wrow0 * (p.N / 2) + (n_blk * 2 + pair) * 64 + c * 32), (long long)p.N, rows_valid);
          __syncwarp();
        }
      }
      // release this accumulator buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CL == 1) mbar_arrive(&tempty_bar[buf]);
        else mbar_arrive_cluster(cluster_map(&tempty_bar[buf], 0));      // the leader's issuer waits for both epilogues
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();               // no CTA leaves while its peer can still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CL == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess || !f) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(f);
  }
  return fn;
}

// bf16 2-D tensor map: `inner` contiguous elements per row, `outer` rows, row pitch `ld` elements,
// box = 64 x box_rows, 128-byte swizzle, zero OOB fill.
inline int make_tmap_bf16(CUtensorMap* tm, const void* ptr, long long inner, long long outer, long long ld, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return -1;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}

struct GemmOperand {
  const void* ptr;
  long long ld;      // row pitch in elements of the stored matrix
  bool mn_major;     // false: stored [MN][K];  true: stored [K][MN]
  const void* ptr2 = nullptr;   // optional second K segment of A (K-major only), starting at k = GemmParams::K1
  long long ld2 = 0;
};

// CTA pairing (CL = 2): mode 3 (default) pairs the launches whose main loop dominates - at least 16 k-blocks (K >= 1024) per work item, two tiles per
// SM - mode 2 pairs every launch, mode 1 none (TFX_GEMM_CLUSTER in the environment, or tfx_gemm_set_cluster_mode).  Measured on the b128 step shapes
// (profiles/r02_gemm_pair_experiments.txt): K = 2816 dgrad / K = 131072 split-K wgrad gain 7 % (1530 TFLOP/s), the K = 512 launches are bound by
// their epilogues and gain nothing (-1 .. +3 %).
inline int& gemm_cluster_mode_ref() {
  static int mode = -1;
  if (mode < 0) { const char* e = getenv("TFX_GEMM_CLUSTER"); mode = e ? atoi(e) : 3; }
  return mode;
}
inline int gemm_cluster_mode() { return gemm_cluster_mode_ref(); }

template <int BN, bool A_MN, bool B_MN, int EPI>
int launch_gemm_t(const GemmOperand& A, const GemmOperand& B, const GemmParams& p_in, int num_sms, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, EPI>;
  GemmParams p = p_in;
  CUtensorMap tmA, tmA2, tmB;
  int rc;
  const bool two = (!A_MN) && A.ptr2 != nullptr;
  if (!two) p.K1 = p.K;
  if (!A_MN) rc = make_tmap_bf16(&tmA, A.ptr, two ? p.K1 : p.K, p.M, A.ld, GEMM_BM); else rc = make_tmap_bf16(&tmA, A.ptr, p.M, p.K, A.ld, GEMM_BK);
  if (rc) return rc;
  if (two) { rc = make_tmap_bf16(&tmA2, A.ptr2, p.K - p.K1, p.M, A.ld2, GEMM_BM); if (rc) return rc; } else tmA2 = tmA;
  {
    const int kbt = (p.K + GEMM_BK - 1) / GEMM_BK;
    if (p.k_splits < 1) p.k_splits = 1;
    if (p.k_splits > kbt) p.k_splits = kbt;
    const int per = (kbt + p.k_splits - 1) / p.k_splits;
    p.k_splits = (kbt + per - 1) / per;            // no empty split
  }
  const int m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM, n_tiles = (p.N + BN - 1) / BN;
  const int items = m_tiles * n_tiles * p.k_splits;
  if (items <= 0) return 0;
  const int mode = gemm_cluster_mode();
  const int kb_item = ((p.K + GEMM_BK - 1) / GEMM_BK + p.k_splits - 1) / p.k_splits;      // k-blocks per work item
  const bool paired = mode == 2 || (mode == 3 && m_tiles >= 2 && kb_item >= 16 && items >= 2 * num_sms);
  // paired CTAs each fetch half of the B tile: the box of tmB is BN / 2 rows (K-major; the MN-major boxes are 64 wide either way)
#if GEMM_ABLATE_HALF_B
  if (!B_MN) rc = make_tmap_bf16(&tmB, B.ptr, p.K, p.N, B.ld, BN / 2);
#else
  if (!B_MN) rc = make_tmap_bf16(&tmB, B.ptr, p.K, p.N, B.ld, paired ? BN / 2 : BN);
#endif
  else rc = make_tmap_bf16(&tmB, B.ptr, p.N, p.K, B.ld, GEMM_BK);
  if (rc) return rc;
  if (paired) {
    auto kern = gemm_sm100_kernel<BN, A_MN, B_MN, EPI, 2>;
    static int max_clusters = 0;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(Cfg::THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = stream; cfg.attrs = at; cfg.numAttrs = 1;
    if (!max_clusters) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess) return -2;
      cfg.gridDim = dim3(2 * (num_sms / 2));
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) != cudaSuccess || nc < 1) { cudaGetLastError(); nc = num_sms / 2; }
      max_clusters = nc < num_sms / 2 ? nc : num_sms / 2;
    }
    const int pair_items = ((m_tiles + 1) / 2) * n_tiles * p.k_splits;
    const int clusters = pair_items < max_clusters ? pair_items : max_clusters;
    cfg.gridDim = dim3(2 * clusters);
    return cudaLaunchKernelEx(&cfg, kern, tmA, tmA2, tmB, p) == cudaSuccess ? 0 : -3;
  }
  auto kern = gemm_sm100_kernel<BN, A_MN, B_MN, EPI, 1>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess) return -2;
    attr_set = true;
  }
  const int grid = items < num_sms ? items : num_sms;
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmA2, tmB, p);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

}  // namespace tfx
