// Persistent tcgen05 / TMEM / TMA forward of the span-masked, soft-capped attention (bounded-logit path; semantics and the
// polynomial-tanh / fixed-maximum softmax are those of attention_sm100.cu; reference transfusion.py:998-1027, mask :452-470).
//
// What changed against the round-1 forward (one CTA per (128-row tile, head), 2 CTAs / SM, K / V single-buffered, P through shared memory):
// the per-tile chain TMA -> S MMA -> softmax -> P (smem, proxy fence) -> PV MMA was latency bound (a K / V load was only issued after the
// previous tile's MMA had retired, ~1 us of exposed load latency per key tile).  Here
//   * ONE persistent CTA per SM walks a cost-sorted list of work items; an item = (PAIR of adjacent 128-row query tiles of one sequence, head):
//     both tiles stream the SAME K / V tiles through a 4-stage TMA ring (prefetch distance 3 tiles, also across items); Q is double-buffered
//     across items, so the next item's first S = Q K^T is issued while the current item drains;
//   * two softmax warpgroups (one per query tile of the pair) ping-pong on two S accumulators in TMEM: while one group exponentiates, the
//     tensor core computes the other group's S / PV.  The two groups are independent consumers of the shared tile stream and may drift apart by
//     up to ring-depth tiles (a tile pair rarely has the same number of visible key tiles);
//   * P never touches shared memory: the softmax threads write bf16 P straight into TMEM (tcgen05.st) and the PV product is a TS-form
//     tcgen05.mma (A = P from TMEM, B = V from smem).  TMEM: S0 S1 [0,256) | O0 O1 [256,384) | P0 P1 [384,512) - all 512 columns;
//   * the MMA lane is a small polling state machine over both groups' next S / PV operation, so neither group is blocked behind the other.
//
//   warp 0   : TMA producer (Q pair per item, K_j | V_j per key tile)
//   warp 1   : tcgen05.mma issuer + TMEM owner
//   warps 2-5: softmax group 0 (query tile A), thread <-> query row          warps 6-9: softmax group 1 (query tile B)
#include "sm100_ptx.cuh"
#include "common.cuh"
#include "gemm_sm100.cuh"
#include "../../include/tfx_b200.h"
#include <limits.h>

namespace tfx {

int num_sms();

constexpr int F2_THREADS = 640;
constexpr int F2_STAGES = 4;                       // K / V ring depth (32 KB per stage: K tile | V tile)
constexpr int F2_SMEM = 2 * 32768 + F2_STAGES * 32768 + 512 /*barriers*/ + 4096 /*softmax denominators*/;

// tanh(y) ~= y * (C0 + C1 u + C2 u^2 + C3 u^3 + C4 u^4), u = y^2, |y| <= 0.75 (same fit as attention_sm100.cu)
#define F2_C0 9.9999722832e-01f
#define F2_C1 -3.3323076483e-01f
#define F2_C2 1.3226091649e-01f
#define F2_C3 -4.9280448379e-02f
#define F2_C4 1.2318833231e-02f

__device__ __forceinline__ float f2_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// p & ((a - b) >> 31): keeps p iff a < b.  Opaque PTX: written as C the compiler turns it back into compare + select and parks the predicates
// of a whole unrolled chunk in a register bit mask (PLOP3 / LOP3 chains, measured in the SASS).
__device__ __forceinline__ float f2_keep_if_less(float p, int a, int b) {
  uint32_t r;
  asm("{\n\t.reg .s32 t;\n\tsub.s32 t, %2, %3;\n\tshr.s32 t, t, 31;\n\tand.b32 %0, %1, t;\n\t}" : "=r"(r) : "r"(__float_as_uint(p)), "r"(a), "r"(b));
  return __uint_as_float(r);
}

// 32 scores of one query row -> 16 bf16 pairs of p = 2^(x poly(x^2) - m2); packed fp32x2 FMAs (FFMA2): two scores per instruction.
// MASKED: key i of the chunk is visible iff i < nvis; the mask is an arithmetic AND on the bits of p (no predicates: a compare + select per
// score made ptxas park 32 predicates in a register bit mask - 4 LOP3 per score, measured in the round-1 SASS).
template <bool MASKED>
__device__ __forceinline__ void f2_softmax_chunk(const uint32_t (&r)[32], int nvis, float2 A0, float2 A1, float2 A2, float2 A3, float2 A4, float2 NM2, float2& l2, uint32_t* out) {
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    const float2 x = make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1]));
    const float2 X = __fmul2_rn(x, x);
    float2 gq = __ffma2_rn(A4, X, A3);
    gq = __ffma2_rn(gq, X, A2);
    gq = __ffma2_rn(gq, X, A1);
    gq = __ffma2_rn(gq, X, A0);
    const float2 e = __ffma2_rn(x, gq, NM2);
    float p0 = f2_ex2(e.x), p1 = f2_ex2(e.y);
    if (MASKED) {
      p0 = f2_keep_if_less(p0, i, nvis);
      p1 = f2_keep_if_less(p1, i + 1, nvis);
    }
    l2 = __fadd2_rn(l2, make_float2(p0, p1));
    out[i >> 1] = pack_bf16(p0, p1);
  }
}

struct F2Item { int q0[2], qend[2], n[2]; int kv0, nmax, head; };

// work item k of this CTA (static snake schedule over the cost-sorted pair list, heaviest first)
__device__ __forceinline__ bool f2_item(int k, int n_items, int H, const int* __restrict__ pairs, const int* __restrict__ t_q0, const int* __restrict__ t_qend,
                                        const int* __restrict__ t_kv0, const int* __restrict__ t_kvend, F2Item& it) {
  const int G = gridDim.x;
  const int pos = (k & 1) ? (G - 1 - (int)blockIdx.x) : (int)blockIdx.x;
  const int idx = k * G + pos;
  if (idx >= n_items) return false;
  const int pr = idx / H;
  it.head = idx - pr * H;
  const int code = pairs[pr];
  const int ta = code >> 1;
  it.kv0 = t_kv0[ta];
  it.q0[0] = t_q0[ta]; it.qend[0] = t_qend[ta];
  it.n[0] = (t_kvend[ta] - it.kv0 + 127) >> 7;
  if (code & 1) {
    it.q0[1] = t_q0[ta + 1]; it.qend[1] = t_qend[ta + 1];
    it.n[1] = (t_kvend[ta + 1] - it.kv0 + 127) >> 7;
  } else {
    it.q0[1] = it.q0[0]; it.qend[1] = it.q0[0]; it.n[1] = 0;
  }
  it.nmax = it.n[0] > it.n[1] ? it.n[0] : it.n[1];
  return true;
}

// 20 warps: warp 0 TMA producer, warps 1 / 2 MMA issuers of query tile A / B, warp 3 idle, warps 4..19 softmax
// (query tile, 64-key column half, TMEM lane quadrant).  65536 / 640 threads -> 96 registers per thread.
__global__ void __launch_bounds__(F2_THREADS, 1)
attn_fwd_ts_k(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
              const float* __restrict__ gates, int H, const int* __restrict__ kv_limit, const int* __restrict__ t_q0, const int* __restrict__ t_qend,
              const int* __restrict__ t_kv0, const int* __restrict__ t_kvend, const int* __restrict__ pairs, int n_items,
              __nv_bfloat16* __restrict__ o, long long ld_o, float* __restrict__ lse, int M, float scale, float cap, const float* __restrict__ fast) {
  if (fast[0] == 0.f) return;                       // precondition of this path does not hold: the general kernel does the work
  // declared 1024-byte aligned (SWIZZLE_128B tiles) and used directly, so that the compiler keeps the shared address space (LDS / STS, not generic LD / ST)
  extern __shared__ __align__(1024) uint8_t f2_smem[];
  uint8_t* sQ = f2_smem;                             // [2 item slots][tile A 16 KB | tile B 16 KB]
  uint8_t* sKV = f2_smem + 65536;                    // [F2_STAGES][K 16 KB | V 16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(f2_smem + 65536 + F2_STAGES * 32768);
  uint64_t *q_full = bars, *q_empty = bars + 2, *kv_full = bars + 4, *kv_empty = bars + 4 + F2_STAGES, *s_full = bars + 4 + 2 * F2_STAGES, *s_empty = s_full + 2,
           *p_full = s_full + 4, *p_empty = s_full + 6, *o_full = s_full + 8, *o_empty = s_full + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 12);
  float* sL = reinterpret_cast<float*>(s_full + 14);  // [2 item parities][2 tiles][2 column halves][128 rows] partial softmax denominators
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if (smem_u32(f2_smem) & 1023u) { printf("tfx: attn_fwd_ts dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&q_full[b], 1); mbar_init(&q_empty[b], 2);
      mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], 8); mbar_init(&p_full[b], 8); mbar_init(&p_empty[b], 1); mbar_init(&o_full[b], 1); mbar_init(&o_empty[b], 8);
    }
    for (int s = 0; s < F2_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 2); }
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      F2Item it;
      uint32_t t = 0;
      for (int k = 0; f2_item(k, n_items, H, pairs, t_q0, t_qend, t_kv0, t_kvend, it); ++k) {
        const int qs = k & 1;
        mbar_wait(&q_empty[qs], ((k >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[qs], it.n[1] > 0 ? 32768 : 16384);
        tma_load_2d(&tmQ, &q_full[qs], sQ + qs * 32768, it.head * 64, it.q0[0]);
        if (it.n[1] > 0) tma_load_2d(&tmQ, &q_full[qs], sQ + qs * 32768 + 16384, it.head * 64, it.q0[1]);
        for (int j = 0; j < it.nmax; ++j, ++t) {
          const int st = t & (F2_STAGES - 1);
          mbar_wait(&kv_empty[st], ((t / F2_STAGES) & 1) ^ 1);
          mbar_expect_tx(&kv_full[st], 32768);
          tma_load_2d(&tmK, &kv_full[st], sKV + st * 32768, it.head * 64, it.kv0 + j * 128);
          tma_load_2d(&tmV, &kv_full[st], sKV + st * 32768 + 16384, it.head * 64, it.kv0 + j * 128);
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ===================================================== MMA issuer of query tile w: S(0), then per key tile [S(j+1)] PV(j).  Blocking mbarrier waits
    // (hardware-suspended try_wait, ~60 cycles wake-up) - a single polling lane for both tiles with __nanosleep between sweeps slept ~1 us per hand-off.
    // Both issuers hand K / V stages and Q slots back through count-2 barriers; a tile / item the group does not need is acknowledged once it has LANDED
    // (an earlier plain arrive could complete the previous phase of the stage while the other group still reads it).
    if (lane == 0) {
      const int w = warp - 1;
      constexpr uint32_t idS = umma_idesc_bf16(128, 128, 0, 0);       // S = Q K^T : A, B K-major
      constexpr uint32_t idO = umma_idesc_bf16(128, 64, 0, 1);        // O += P V  : A = P (TMEM), B = V MN-major
      const uint32_t aQ = smem_u32(sQ), aKV = smem_u32(sKV);
      const uint32_t tS = tmem_base + w * 128, tO = tmem_base + 256 + w * 64, tP = tmem_base + 384 + w * 64;
      uint32_t gs = 0, gp = 0, oc = 0, tb = 0;
      F2Item it;
      for (int k = 0; f2_item(k, n_items, H, pairs, t_q0, t_qend, t_kv0, t_kvend, it); ++k) {
        const int qs = k & 1, n = it.n[w], nmax = it.nmax;
        mbar_wait(&q_full[qs], (k >> 1) & 1);
        auto issue_S = [&](int j) {
          const uint32_t t = tb + j;
          const int st = t & (F2_STAGES - 1);
          mbar_wait(&kv_full[st], (t / F2_STAGES) & 1);
          if (gs > 0) mbar_wait(&s_empty[w], (gs - 1) & 1);            // the softmax group has pulled the previous S out of TMEM
          tc_fence_after();
          const uint32_t a = aQ + qs * 32768 + w * 16384, b = aKV + st * 32768;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_ss(tS, umma_smem_desc_sw128(a + kk * 32, 0, 1024), umma_smem_desc_sw128(b + kk * 32, 0, 1024), idS, kk > 0 ? 1u : 0u);
          umma_commit(&s_full[w]);
          ++gs;
          if (j == n - 1) umma_commit(&q_empty[qs]);                    // last S product of the item: this group is done with the Q slot
        };
        if (n > 0) {
          issue_S(0);
          for (int j = 0; j < n; ++j) {
            if (j + 1 < n) issue_S(j + 1);                              // runs ahead: overlaps the softmax of tile j
            const uint32_t t = tb + j;
            const int st = t & (F2_STAGES - 1);
            mbar_wait(&p_full[w], gp & 1);
            if (j == 0 && oc > 0) mbar_wait(&o_empty[w], (oc - 1) & 1);  // the previous item's O has been read out
            tc_fence_after();
            const uint32_t bV = aKV + st * 32768 + 16384;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_bf16_ts(tO, tP + kk * 8, umma_smem_desc_sw128(bV + kk * 2048, 8192, 1024), idO, (j > 0 || kk > 0) ? 1u : 0u);
            umma_commit(&p_empty[w]);
            umma_commit(&kv_empty[st]);
            ++gp;
            if (j == n - 1) { umma_commit(&o_full[w]); ++oc; }
          }
        } else {
          mbar_arrive(&q_empty[qs]);
        }
        for (int j = n; j < nmax; ++j) {                                 // key tiles only the other group needs
          const uint32_t t = tb + j;
          const int st = t & (F2_STAGES - 1);
          mbar_wait(&kv_full[st], (t / F2_STAGES) & 1);
          mbar_arrive(&kv_empty[st]);
        }
        tb += nmax;
      }
    }
  } else if (warp >= 4) {
    // ===================================================== softmax warps (thread <-> query row, 64 of the tile's 128 keys)
    const int idx = warp - 4;
    const int w = idx >> 3, ch = (idx >> 2) & 1, quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = uint32_t(quad * 32) << 16;
    const uint32_t tS = tmem_base + w * 128 + ch * 64 + lane_addr, tO = tmem_base + 256 + w * 64 + ch * 32 + lane_addr, tP = tmem_base + 384 + w * 64 + ch * 32 + lane_addr;
    const float k1 = scale / cap;                    // y = x * k1
    const float KL = cap * 1.4426950408889634f;      // exponent (base 2) = KL * tanh(y) - m2
    const float m2 = fast[1] * 1.4426950408889634f;
    const float k2 = k1 * k1;
    const float a0 = KL * k1 * F2_C0, a1 = KL * k1 * k2 * F2_C1, a2 = KL * k1 * k2 * k2 * F2_C2, a3 = KL * k1 * k2 * k2 * k2 * F2_C3,
                a4 = KL * k1 * k2 * k2 * k2 * k2 * F2_C4;
    const float2 A0 = make_float2(a0, a0), A1 = make_float2(a1, a1), A2 = make_float2(a2, a2), A3 = make_float2(a3, a3), A4 = make_float2(a4, a4),
                 NM2 = make_float2(-m2, -m2);
    uint32_t g = 0, oc = 0;
    F2Item it;
    for (int k = 0; f2_item(k, n_items, H, pairs, t_q0, t_qend, t_kv0, t_kvend, it); ++k) {
      const int n = it.n[w];
      if (n == 0) continue;
      const int grow = it.q0[w] + row;
      const bool valid = grow < it.qend[w];
      const int lim = valid ? kv_limit[grow] : -1;
      const int wmin = __reduce_min_sync(0xffffffffu, valid ? lim : INT_MAX);   // key columns entirely below it need no mask (per warp)
      float gate = 1.f;
      if (valid && gates) gate = 1.f / (1.f + __expf(-gates[(long long)grow * H + it.head]));
      float2 l2 = make_float2(0.f, 0.f);
      for (int j = 0; j < n; ++j, ++g) {
        const int key0 = it.kv0 + j * 128 + ch * 64;                      // first key of this warp's column half
        const bool all_visible = key0 + 63 <= wmin;
        mbar_wait(&s_full[w], g & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32], pk[16];
          tmem_ld_32x32b_x32(tS + c * 32, r);
          tmem_ld_wait();
          if (c == 1) {                                // this warp's part of S is in registers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[w]);
          }
          // the mask costs 3 integer instructions per score and is only needed on diagonal / span-boundary tiles: two specialised code paths, chosen per warp
          if (all_visible) f2_softmax_chunk<false>(r, 0, A0, A1, A2, A3, A4, NM2, l2, pk);
          else f2_softmax_chunk<true>(r, lim - (key0 + c * 32) + 1, A0, A1, A2, A3, A4, NM2, l2, pk);
          if (c == 0 && g > 0) { mbar_wait(&p_empty[w], (g - 1) & 1); tc_fence_after(); }    // the previous PV product has consumed the P buffer
          tmem_st_32x32b_x16(tP + c * 16, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[w]);
      }
      // ---- epilogue of the item: the two column halves add their denominators, each writes 32 of the 64 output columns
      float* sl = sL + (oc & 1) * 512 + w * 256;
      sl[ch * 128 + row] = l2.x + l2.y;
      asm volatile("bar.sync %0, 256;" ::"r"(1 + w) : "memory");
      const float l = sl[row] + sl[128 + row];
      mbar_wait(&o_full[w], oc & 1);
      ++oc;
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32b_x32(tO, r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[w]);
      const float gsc = (l > 0.f ? 1.f / l : 0.f) * gate;
      if (valid) {
        __nv_bfloat16* dst = o + (long long)grow * ld_o + it.head * 64 + ch * 32;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          uint32_t wv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) wv[e] = pack_bf16(__uint_as_float(r[qd * 8 + 2 * e]) * gsc, __uint_as_float(r[qd * 8 + 2 * e + 1]) * gsc);
          *reinterpret_cast<uint4*>(dst + qd * 8) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        }
        if (lse && ch == 0) lse[(long long)it.head * M + grow] = fast[1] + logf(l);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace tfx

using namespace tfx;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int tfx_attn_fwd_ts(const void* q, const void* k, const void* v, long long ld_q, long long ld_k, long long ld_v, const float* gates, int H,
                    const int* kv_limit, const int* tile_q0, const int* tile_qend, const int* tile_kv0, const int* tile_kvend, int n_tiles,
                    const int* pairs, int n_pairs, void* o, long long ld_o, float* lse, int M, int M_kv, float scale, float softcap, const float* fast_params,
                    void* stream) {
  if (n_pairs <= 0 || n_tiles <= 0) return 0;
  if (M_kv <= 0) M_kv = M;
  TFX_REQUIRE(fast_params != nullptr && pairs != nullptr, "attn_fwd_ts: fast_params (from tfx_attn_fast_params) and the tile-pair list are required");
  TFX_REQUIRE(ld_q % 8 == 0 && ld_k % 8 == 0 && ld_v % 8 == 0 && ld_o % 8 == 0, "attn_fwd_ts: row pitches must be multiples of 8 bf16");
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_tmap_bf16(&tq, q, (long long)H * 64, M, ld_q, 128)) || (rc = make_tmap_bf16(&tk, k, (long long)H * 64, M_kv, ld_k, 128)) ||
      (rc = make_tmap_bf16(&tv, v, (long long)H * 64, M_kv, ld_v, 128))) {
    set_error("attn_fwd_ts: cuTensorMapEncodeTiled failed (%d)", rc);
    return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(attn_fwd_ts_k, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM) != cudaSuccess) { set_error("attn_fwd_ts: cannot raise dynamic smem"); return -2; }
    attr_set = true;
  }
  const int n_items = n_pairs * H;
  const int grid = n_items < num_sms() ? n_items : num_sms();          // persistent: one CTA per SM
  attn_fwd_ts_k<<<grid, F2_THREADS, F2_SMEM, ST(stream)>>>(tq, tk, tv, gates, H, kv_limit, tile_q0, tile_qend, tile_kv0, tile_kvend, pairs, n_items, (__nv_bfloat16*)o, ld_o, lse, M,
                                                          scale, softcap, fast_params);
  return check_launch("attn_fwd_ts");
}

}  // extern "C"
