// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA (128B swizzle) -> smem ring -> tcgen05.mma
// (cta_group::1, UMMA 128 x BN x 16, fp32 accumulators double-buffered in TMEM) -> tcgen05.ld epilogue
// with the Transfusion-specific fused epilogues.
//
//   D[m][n] = sum_k A(m,k) * B(n,k)
//   A "K-major":  stored row-major [M][K]   (activations as GEMM input, dY for dgrad)
//   A "MN-major": stored row-major [K][M]   (dY^T for wgrad: K = tokens)
//   B likewise over n.
//
// Roles (192 threads): warp 0 = TMA producer (1 lane), warp 1 = MMA issuer (1 lane) + TMEM owner,
// warps 2..5 = epilogue (warp%4 selects the TMEM lane quadrant; thread <-> one accumulator row).
#pragma once
#include "sm100_ptx.cuh"

namespace tfx {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;     // 64 bf16 = 128 B = one swizzle atom row
constexpr int GEMM_UK = 16;     // UMMA K for 16-bit inputs
constexpr int GEMM_THREADS = 192;

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
  float* qk_inv;               // [M][2H]  1/max(|x|,eps) for q heads then k heads
  const float *q_gamma, *k_gamma;   // [64]
  const int* rope_pos;         // [M]
  const float2* rope_cs;       // [max_pos][32] (cos, sin)
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

template <int BN> struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 2 * BN;   // double-buffered accumulator (power of two: 256 / 512)
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_sm100_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
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
  const int num_items = m_tiles * n_tiles * p.k_splits;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4); }
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int split = item / (m_tiles * n_tiles);
        const int rem = item - split * (m_tiles * n_tiles);
        const int m_blk = rem / n_tiles, n_blk = rem - m_blk * n_tiles;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sB = sA + Cfg::A_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
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
            for (int a = 0; a < BN / 64; ++a)
              tma_load_2d(&tmB, &full_bar[stage], sB + a * (GEMM_BK * 128), n_blk * BN + a * 64, kb * GEMM_BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0; uint32_t phase = 0;
      int local = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++local) {
        const int split = item / (m_tiles * n_tiles);
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
          const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sB = sA + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / GEMM_UK; ++k) {
            const uint64_t da = A_MN ? umma_smem_desc_sw128(sA + k * (GEMM_UK * 128), GEMM_BK * 128, 1024)
                                     : umma_smem_desc_sw128(sA + k * (GEMM_UK * 2), 0, 1024);
            const uint64_t db = B_MN ? umma_smem_desc_sw128(sB + k * (GEMM_UK * 128), GEMM_BK * 128, 1024)
                                     : umma_smem_desc_sw128(sB + k * (GEMM_UK * 2), 0, 1024);
            umma_bf16_ss(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);          // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[buf]);              // accumulator ready for the epilogue
      }
    }
  } else {
    // ===================================================== epilogue warps (2..5)
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    int local = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++local) {
      const int split = item / (m_tiles * n_tiles);
      const int rem = item - split * (m_tiles * n_tiles);
      const int m_blk = rem / n_tiles, n_blk = rem - m_blk * n_tiles;
      const int buf = local & 1;
      const uint32_t bphase = (local >> 1) & 1;
      mbar_wait(&tfull_bar[buf], bphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + buf * BN;
      const int row = m_blk * GEMM_BM + row_in_tile;
      const bool row_ok = row < p.M;
      const int col0 = n_blk * BN;

      if constexpr (EPI == EPI_STORE) {
        long long f32_base = -1;
        if (p.out_f32 && row_ok) f32_base = p.row_off ? p.row_off[row] : (long long)row * p.ld_f32;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + c * 32, r);
          tmem_ld_wait();
          const int cbase = col0 + c * 32;
          if (row_ok && cbase < p.N) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              v[j] = __uint_as_float(r[j]) * p.alpha;
              if (p.bias && cbase + j < p.N) v[j] += p.bias[cbase + j];
            }
            if (f32_base >= 0) {
              float* dst = p.out_f32 + f32_base + cbase;
              const bool vec = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && (cbase + 32 <= p.N);
              if (p.accumulate_f32) {
                if (vec) {
#pragma unroll
                  for (int j = 0; j < 32; j += 4)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
                } else {
                  for (int j = 0; j < 32; ++j) if (cbase + j < p.N) atomicAdd(dst + j, v[j]);
                }
              } else {
                if (vec) {
#pragma unroll
                  for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
                  for (int j = 0; j < 32; ++j) if (cbase + j < p.N) dst[j] = v[j];
                }
              }
            }
            if (p.out_bf16) {
              __nv_bfloat16* dst = p.out_bf16 + (long long)row * p.ld_bf16 + cbase;
              const bool vec = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && (cbase + 32 <= p.N);
              if (vec) {
#pragma unroll
                for (int j = 0; j < 32; j += 8)
                  *reinterpret_cast<uint4*>(dst + j) = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]), pack_bf16(v[j + 4], v[j + 5]), pack_bf16(v[j + 6], v[j + 7]));
              } else {
                for (int j = 0; j < 32; ++j) if (cbase + j < p.N) dst[j] = __float2bfloat16(v[j]);
              }
            }
          }
        }
      } else if constexpr (EPI == EPI_QKVG) {
        static_assert(EPI != EPI_QKVG || BN == 128, "QKVG epilogue expects 128-wide N tiles");
        const int tps = p.H >> 1;             // tiles per section
        const int kind = n_blk / tps;         // 0 q, 1 k, 2 v, 3 gates
        const int tis = n_blk - kind * tps;   // tile in section
        const long long HI = (long long)p.H * 64;
        if (kind <= 1) {
          const float* gamma = kind == 0 ? p.q_gamma : p.k_gamma;
          __nv_bfloat16* dstm = kind == 0 ? p.q : p.k;
          const int pos = row_ok ? p.rope_pos[row] : 0;
          const float2* cs = p.rope_cs + (long long)pos * 32;
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32b_x32(taddr + hh * 64, r0);
            tmem_ld_32x32b_x32(taddr + hh * 64 + 32, r1);
            tmem_ld_wait();
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) { float a = __uint_as_float(r0[j]), b = __uint_as_float(r1[j]); ss += a * a + b * b; }
            const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
            const int head = tis * 2 + hh;
            if (row_ok) {
              p.qk_inv[(long long)row * 2 * p.H + kind * p.H + head] = inv;
              __nv_bfloat16* dst = dstm + (long long)row * HI + head * 64;
              const float sc = inv * 8.f;
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                uint32_t* rr = half == 0 ? r0 : r1;
                uint32_t outw[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const int d0 = half * 32 + 2 * i;
                  const float y0 = __uint_as_float(rr[2 * i]) * sc * (gamma[d0] + 1.f);
                  const float y1 = __uint_as_float(rr[2 * i + 1]) * sc * (gamma[d0 + 1] + 1.f);
                  const float2 c = cs[half * 16 + i];
                  outw[i] = pack_bf16(y0 * c.x - y1 * c.y, y1 * c.x + y0 * c.y);
                }
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                  *reinterpret_cast<uint4*>(dst + half * 32 + i * 2) = make_uint4(outw[i], outw[i + 1], outw[i + 2], outw[i + 3]);
              }
            }
          }
        } else if (kind == 2) {
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(taddr + c * 32, r);
            tmem_ld_wait();
            if (row_ok) {
              __nv_bfloat16* dst = p.v + (long long)row * HI + tis * 128 + c * 32;
#pragma unroll
              for (int j = 0; j < 32; j += 8)
                *reinterpret_cast<uint4*>(dst + j) = make_uint4(pack_bf16(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), pack_bf16(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])),
                                                                 pack_bf16(__uint_as_float(r[j + 4]), __uint_as_float(r[j + 5])), pack_bf16(__uint_as_float(r[j + 6]), __uint_as_float(r[j + 7])));
            }
          }
        } else {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr, r);
          tmem_ld_wait();
          if (row_ok) {
            float* dst = p.gates + (long long)row * p.H;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < p.H) dst[j] = __uint_as_float(r[j]);
          }
        }
      } else if constexpr (EPI == EPI_RESID) {
        const int crow = (row_ok && p.cond_row) ? p.cond_row[row] : -1;
        const float* zrow = (p.zgate && crow >= 0) ? p.zgate + (long long)crow * p.zgate_ld : nullptr;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + c * 32, r);
          tmem_ld_wait();
          const int cbase = col0 + c * 32;
          if (row_ok && cbase < p.N) {    // N is a multiple of 32 for every RESID use (512)
            const long long off = (long long)row * p.N + cbase;
            float y[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) y[j] = __uint_as_float(r[j]) + (p.bias ? p.bias[cbase + j] : 0.f);
            if (p.y_bf16) {
#pragma unroll
              for (int j = 0; j < 32; j += 8)
                *reinterpret_cast<uint4*>(p.y_bf16 + off + j) = make_uint4(pack_bf16(y[j], y[j + 1]), pack_bf16(y[j + 2], y[j + 3]), pack_bf16(y[j + 4], y[j + 5]), pack_bf16(y[j + 6], y[j + 7]));
            }
            float o[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 xr = *reinterpret_cast<const float4*>(p.x_res + off + j);
              float4 s = make_float4(1.f, 1.f, 1.f, 1.f);
              if (zrow) s = *reinterpret_cast<const float4*>(zrow + cbase + j);
              else if (p.ls) { s = *reinterpret_cast<const float4*>(p.ls + cbase + j); s.x += 1.f; s.y += 1.f; s.z += 1.f; s.w += 1.f; }
              o[j] = xr.x + y[j] * s.x; o[j + 1] = xr.y + y[j + 1] * s.y; o[j + 2] = xr.z + y[j + 2] * s.z; o[j + 3] = xr.w + y[j + 3] * s.w;
              *reinterpret_cast<float4*>(p.x_out + off + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
            }
            if (p.x_out_bf16) {
#pragma unroll
              for (int j = 0; j < 32; j += 8)
                *reinterpret_cast<uint4*>(p.x_out_bf16 + off + j) = make_uint4(pack_bf16(o[j], o[j + 1]), pack_bf16(o[j + 2], o[j + 3]), pack_bf16(o[j + 4], o[j + 5]), pack_bf16(o[j + 6], o[j + 7]));
            }
          }
        }
      } else if constexpr (EPI == EPI_GEGLU) {
        static_assert(EPI != EPI_GEGLU || BN == 128, "GEGLU epilogue expects 128-wide N tiles");
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t rv[32], rg[32];
          tmem_ld_32x32b_x32(taddr + c * 32, rv);
          tmem_ld_32x32b_x32(taddr + 64 + c * 32, rg);
          tmem_ld_wait();
          if (row_ok) {
            const int cv = col0 + c * 32, cg = col0 + 64 + c * 32;
            float v[32], g[32], hh[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              v[j] = __uint_as_float(rv[j]) + p.bias[cv + j];
              g[j] = __uint_as_float(rg[j]) + p.bias[cg + j];
              hh[j] = gelu_erf(g[j]) * v[j];
            }
            __nv_bfloat16* dv = p.vg + (long long)row * p.N + cv;
            __nv_bfloat16* dg = p.vg + (long long)row * p.N + cg;
            __nv_bfloat16* dh = p.h + (long long)row * (p.N / 2) + n_blk * 64 + c * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              *reinterpret_cast<uint4*>(dv + j) = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]), pack_bf16(v[j + 4], v[j + 5]), pack_bf16(v[j + 6], v[j + 7]));
              *reinterpret_cast<uint4*>(dg + j) = make_uint4(pack_bf16(g[j], g[j + 1]), pack_bf16(g[j + 2], g[j + 3]), pack_bf16(g[j + 4], g[j + 5]), pack_bf16(g[j + 6], g[j + 7]));
              *reinterpret_cast<uint4*>(dh + j) = make_uint4(pack_bf16(hh[j], hh[j + 1]), pack_bf16(hh[j + 2], hh[j + 3]), pack_bf16(hh[j + 4], hh[j + 5]), pack_bf16(hh[j + 6], hh[j + 7]));
            }
          }
        }
      }
      // release this accumulator buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
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

template <int BN, bool A_MN, bool B_MN, int EPI>
int launch_gemm_t(const GemmOperand& A, const GemmOperand& B, const GemmParams& p_in, int num_sms, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
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
  if (!B_MN) rc = make_tmap_bf16(&tmB, B.ptr, p.K, p.N, B.ld, BN); else rc = make_tmap_bf16(&tmB, B.ptr, p.N, p.K, B.ld, GEMM_BK);
  if (rc) return rc;
  auto kern = gemm_sm100_kernel<BN, A_MN, B_MN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess) return -2;
    attr_set = true;
  }
  const int m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM, n_tiles = (p.N + BN - 1) / BN;
  const int items = m_tiles * n_tiles * p.k_splits;
  if (items <= 0) return 0;
  const int grid = items < num_sms ? items : num_sms;
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmA2, tmB, p);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

}  // namespace tfx
