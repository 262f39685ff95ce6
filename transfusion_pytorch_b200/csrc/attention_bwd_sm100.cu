// Persistent tcgen05 / TMEM / TMA backward of the span-masked, soft-capped attention - TRANSPOSED-SCORE formulation (bounded-logit path;
// math of attention_sm100.cu / attention.cu; reference transfusion.py:998-1027 under autograd).
//
// One CTA per SM walks (128-key tile, head) items (heaviest first, snake order) and sweeps the 128-row query tiles that can see those keys.
// Against the round-1 kernel (S = Q K^T with thread <-> query row, P and dS both through shared memory, Q / dO ring of 2):
//   * the score tile is computed TRANSPOSED:  S^T = K Q^T,  dP^T = V dO^T   (TMEM lanes = keys).  P^T and dS^T are then exactly the A operands
//     of  dV += P^T dO  and  dK += dS^T Q : the softmax threads write them back into TMEM (tcgen05.st, over the columns they just read) and both
//     products are TS-form tcgen05.mma - no shared-memory round trip, no proxy fence for P at all; only dS^T also goes to shared memory, where the
//     SAME tile is the (MN-major) A operand of  dQ = dS K;
//   * the shared memory this frees deepens the Q / dO TMA ring to 3 stages: a stage is held from the load until the step's gradient products
//     retire, and S^T runs one step ahead, so with 2 stages every step exposed a full TMA round trip (the dominant cost in round 1);
//   * per-QUERY softmax statistics (lse, D = rowsum(dO o), visibility limit) are per-COLUMN data here: staged once per step in shared memory
//     and read as warp-uniform (broadcast) vector loads;
//   * the exponentials only need S^T: they are computed while dP^T of the same step is still in flight (separate completion barriers).
//
//   warp 0   : TMA producer (K | V per item - K double-buffered, V single; Q | dO per step, ring of 3)
//   warp 1   : tcgen05.mma issuer + TMEM owner
//   warps 4-19: softmax / gradient warps; warp = (TMEM lane quadrant = 32 keys, query-column quarter = 32 queries): 4 warps per scheduler hide the
//              TMEM-load / MUFU / barrier latencies that left the 8-warp version at 0.32 IPC
//   TMEM: S^T [0,128) | dP^T [128,256) (dS^T bf16 written back over each warp's own columns) | dV [256,320) | dK [320,384) | dQ [384,448) | P^T bf16 [448,512)
#include "sm100_ptx.cuh"
#include "common.cuh"
#include "gemm_sm100.cuh"
#include "../../include/tfx_b200.h"
#include <limits.h>

// Ablation hooks (tools/bench_attn.py with TFX_LIB=<variant library>; compiled out of the product build): bit 0 no dQ path, bit 1 no dS math, bit 2 no exp math,
// bit 3 no dV / dK products, bit 4 no S^T / dP^T products
#ifndef B2_VARIANT
#define B2_VARIANT 0
#endif
#ifndef B2_PIPE
#define B2_PIPE 0      // 1: software-pipelined softmax / gradient warps (exponentials of step g + 1 in the shadow of step g's gradient products).  Correct
                       // (unit tests green) and measured EQUAL to the plain order - 282.6 vs 280.6 us at M = 32 k, 1055 vs 1066 us at M = 131 k
                       // (profiles/r02_attn_bwd_ablation.txt): the warps are issue-bound (IPC 0.48 per scheduler), not idle in the hand-off shadow.
#endif

namespace tfx {

int num_sms();

constexpr int B2_THREADS = 640;                     // warp 0 TMA, warp 1 MMA, warps 2-3 idle, warps 4..19 softmax / gradient (65536 / 640 -> 96 registers)
constexpr int B2_QST = 3;                           // Q / dO ring depth
constexpr int B2_OFF_K = 0, B2_OFF_V = 32768, B2_OFF_QDO = 49152, B2_OFF_DS = B2_OFF_QDO + B2_QST * 32768, B2_OFF_DQ = B2_OFF_DS + 32768,
              B2_OFF_META = B2_OFF_DQ + 32768, B2_OFF_BARS = B2_OFF_META + 2 * 2048;
constexpr int B2_SMEM = B2_OFF_BARS + 512;

#define B2_C0 9.9999722832e-01f
#define B2_C1 -3.3323076483e-01f
#define B2_C2 1.3226091649e-01f
#define B2_C3 -4.9280448379e-02f
#define B2_C4 1.2318833231e-02f

__device__ __forceinline__ float b2_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void b2_tma_reduce_add_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void b2_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void b2_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void b2_bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// p & ((a - b) >> 31): keeps p iff a < b.  Opaque PTX: written as C the compiler turns it back into compare + select and parks the predicates
// of a whole unrolled chunk in a register bit mask (PLOP3 / LOP3 chains, measured in the SASS).
__device__ __forceinline__ float b2_keep_if_less(float p, int a, int b) {
  uint32_t r;
  asm("{\n\t.reg .s32 t;\n\tsub.s32 t, %2, %3;\n\tshr.s32 t, t, 31;\n\tand.b32 %0, %1, t;\n\t}" : "=r"(r) : "r"(__float_as_uint(p)), "r"(a), "r"(b));
  return __uint_as_float(r);
}

// pipelined order: the same chunk, but scale * (1 - tanh^2) = fma(e^2, OC, SC) leaves as bf16 pairs (it waits in registers across a hand-off)
template <bool MASKED>
__device__ __forceinline__ void b2_exp_chunk_om(const uint32_t (&rs)[16], const float* __restrict__ nl, const int* __restrict__ lim1, int key, float2 A0, float2 A1, float2 A2,
                                                float2 A3, float2 A4, float2 OC, float2 SC, uint32_t* __restrict__ om, uint32_t* __restrict__ wp) {
#pragma unroll
  for (int e2 = 0; e2 < 16; e2 += 2) {
    const float2 x = make_float2(__uint_as_float(rs[e2]), __uint_as_float(rs[e2 + 1]));
    const float2 X = __fmul2_rn(x, x);
    float2 gp = __ffma2_rn(A4, X, A3);
    gp = __ffma2_rn(gp, X, A2);
    gp = __ffma2_rn(gp, X, A1);
    gp = __ffma2_rn(gp, X, A0);
    const float2 e = __fmul2_rn(x, gp);
    const float2 pe = __fadd2_rn(e, *reinterpret_cast<const float2*>(nl + e2));
    float p0 = b2_ex2(pe.x), p1 = b2_ex2(pe.y);
    if (MASKED) {
      const int2 lm = *reinterpret_cast<const int2*>(lim1 + e2);
      p0 = b2_keep_if_less(p0, key, lm.x);
      p1 = b2_keep_if_less(p1, key, lm.y);
    }
    const float2 oms = __ffma2_rn(__fmul2_rn(e, e), OC, SC);
    om[e2 >> 1] = pack_bf16(oms.x, oms.y);
    wp[e2 >> 1] = pack_bf16(p0, p1);
  }
}

// 32 transposed scores (one key row x 32 queries) -> e = cap log2e tanh(y) (kept for the 1 - tanh^2 factor) and bf16 pairs of p = 2^(e - lse2[q]).
// nl: -lse2 per query, lim1: visibility limit + 1 per query (warp-uniform shared-memory addresses: broadcast loads).
// MASKED: query q sees this key iff key < lim1[q]; arithmetic AND on the bits of p instead of compare + select (no predicate pressure).
template <bool MASKED>
__device__ __forceinline__ void b2_exp_chunk(const uint32_t (&rs)[16], const float* __restrict__ nl, const int* __restrict__ lim1, int key, float2 A0, float2 A1, float2 A2,
                                             float2 A3, float2 A4, float2 (&ee)[8], uint32_t (&wp)[8]) {
#pragma unroll
  for (int e2 = 0; e2 < 16; e2 += 2) {
    const float2 x = make_float2(__uint_as_float(rs[e2]), __uint_as_float(rs[e2 + 1]));
    const float2 X = __fmul2_rn(x, x);
    float2 gp = __ffma2_rn(A4, X, A3);
    gp = __ffma2_rn(gp, X, A2);
    gp = __ffma2_rn(gp, X, A1);
    gp = __ffma2_rn(gp, X, A0);
    const float2 e = __fmul2_rn(x, gp);
    const float2 pe = __fadd2_rn(e, *reinterpret_cast<const float2*>(nl + e2));
    float p0 = b2_ex2(pe.x), p1 = b2_ex2(pe.y);
    if (MASKED) {
      const int2 lm = *reinterpret_cast<const int2*>(lim1 + e2);
      p0 = b2_keep_if_less(p0, key, lm.x);
      p1 = b2_keep_if_less(p1, key, lm.y);
    }
    ee[e2 >> 1] = e;
    wp[e2 >> 1] = pack_bf16(p0, p1);
  }
}

struct B2Item { int kv0, kv_end, q_begin, q_end, n_q, head; };

__device__ __forceinline__ bool b2_item(int k, int n_items, int H, const int* __restrict__ order, const int* __restrict__ kt_kv0, const int* __restrict__ kt_kvend,
                                        const int* __restrict__ kt_q0, const int* __restrict__ kt_qend, B2Item& it) {
  const int G = gridDim.x;
  const int pos = (k & 1) ? (G - 1 - (int)blockIdx.x) : (int)blockIdx.x;      // snake: odd stripes run backwards
  const int idx = k * G + pos;
  if (idx >= n_items) return false;
  const int t = idx / H;
  const int tile = order ? order[t] : t;
  it.head = idx - t * H;
  it.kv0 = kt_kv0[tile]; it.kv_end = kt_kvend[tile]; it.q_begin = kt_q0[tile]; it.q_end = kt_qend[tile];
  it.n_q = (it.q_end - it.q_begin + 127) >> 7;
  return true;
}

__global__ void __launch_bounds__(B2_THREADS, 1)
attn_bwd_ts_k(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
              const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmDQ,
              const float* __restrict__ lse, const float* __restrict__ dsum, const int* __restrict__ kv_limit,
              const int* __restrict__ kt_kv0, const int* __restrict__ kt_kvend, const int* __restrict__ kt_q0, const int* __restrict__ kt_qend,
              const int* __restrict__ kt_order, int n_items,
              float* __restrict__ dk, __nv_bfloat16* __restrict__ dv, long long ld_dv, int M, int H, float scale, float cap, const float* __restrict__ fast) {
  if (fast[0] == 0.f) return;
  // declared 1024-byte aligned and used directly: the compiler keeps the shared address space (LDS / STS instead of generic LD / ST)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem + B2_OFF_K;                     // [2][16 KB]
  uint8_t* sV = smem + B2_OFF_V;                     // [16 KB]
  uint8_t* sQDO = smem + B2_OFF_QDO;                 // [B2_QST][Q 16 KB | dO 16 KB]
  uint8_t* sDS = smem + B2_OFF_DS;                   // dS^T: [2 query halves][128 key rows][128 B]
  uint8_t* sDQ = smem + B2_OFF_DQ;                   // [2 column halves][128 rows][128 B] fp32
  float* sMeta = reinterpret_cast<float*>(smem + B2_OFF_META);      // [2 buffers][lse2 128 | D 128 | lim 128 (int) | min lim]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B2_OFF_BARS);
  uint64_t *k_full = bars /*[2]*/, *k_empty = bars + 2 /*[2]*/, *v_full = bars + 4, *v_empty = bars + 5, *qdo_full = bars + 6 /*[3]*/, *qdo_empty = bars + 9 /*[3]*/,
           *s_full = bars + 12, *dp_full = bars + 13, *s_free = bars + 14, *dp_free = bars + 15, *pt_full = bars + 16, *ds_full = bars + 17, *grad_done = bars + 18,
           *dq_full = bars + 19, *dq_free = bars + 20, *dkv_full = bars + 21, *dkv_free = bars + 22, *meta_full = bars + 26 /*[2]*/, *dq_staged = bars + 28, *dq_slab_free = bars + 29;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) { printf("tfx: attn_bwd_ts dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmDQ);
    for (int b = 0; b < 2; ++b) { mbar_init(&k_full[b], 1); mbar_init(&k_empty[b], 1); }
    mbar_init(v_full, 1); mbar_init(v_empty, 1);
    for (int b = 0; b < B2_QST; ++b) { mbar_init(&qdo_full[b], 1); mbar_init(&qdo_empty[b], 1); }
    mbar_init(s_full, 1); mbar_init(dp_full, 1); mbar_init(s_free, 16); mbar_init(pt_full, 16); mbar_init(ds_full, 16); mbar_init(grad_done, 1);
    mbar_init(dq_full, 1); mbar_init(dq_free, 16); mbar_init(dkv_full, 1); mbar_init(dkv_free, 16); mbar_init(&meta_full[0], 4); mbar_init(&meta_full[1], 4);
    mbar_init(dq_staged, 16); mbar_init(dq_slab_free, 1);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 320, tDQ = tmem_base + 384, tPT = tmem_base + 448;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      B2Item it;
      uint32_t g = 0;
      for (int k = 0; b2_item(k, n_items, H, kt_order, kt_kv0, kt_kvend, kt_q0, kt_qend, it); ++k) {
        const int kb = k & 1;
        mbar_wait(&k_empty[kb], ((k >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[kb], 16384);
        tma_load_2d(&tmK, &k_full[kb], sK + kb * 16384, it.head * 64, it.kv0);
        mbar_wait(v_empty, (k & 1) ^ 1);
        mbar_expect_tx(v_full, 16384);
        tma_load_2d(&tmV, v_full, sV, it.head * 64, it.kv0);
        for (int i = 0; i < it.n_q; ++i, ++g) {
          const int b = g % B2_QST;
          mbar_wait(&qdo_empty[b], ((g / B2_QST) & 1) ^ 1);
          mbar_expect_tx(&qdo_full[b], 32768);
          tma_load_2d(&tmQ, &qdo_full[b], sQDO + b * 32768, it.head * 64, it.q_begin + i * 128);
          tma_load_2d(&tmDO, &qdo_full[b], sQDO + b * 32768 + 16384, it.head * 64, it.q_begin + i * 128);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer.  Per step g:  S^T(g) | dP^T(g)  ...softmax...  dV(g), dK(g), dQ(g).
    // S^T(g+1) is issued as soon as S^T(g) has been read (s_free); dP^T(g+1) once dP^T(g) has been read AND dS^T(g), which lives over it, has been
    // consumed by dK(g) - i.e. right behind the gradient products of step g (in-order execution of the tensor pipe covers the hazard).
    if (lane == 0) {
      constexpr uint32_t idS = umma_idesc_bf16(128, 128, 0, 0);       // S^T, dP^T: A (K / V), B (Q / dO) K-major
      constexpr uint32_t idT = umma_idesc_bf16(128, 64, 0, 1);        // dV, dK: A from TMEM, B (dO / Q) MN-major
      constexpr uint32_t idQ = umma_idesc_bf16(128, 64, 1, 1);        // dQ: A = dS^T tile read MN-major (M = queries), B = K MN-major
      const uint32_t aDS = smem_u32(sDS);
      B2Item ia, ib;
      int ka = 0, ia_i = 0, kb_ = 0, ib_i = 0;      // cursor A = (item, query tile) of the next S^T / dP^T; cursor B of the next gradient products
      uint32_t ga = 0, gb = 0;
      bool has_a = b2_item(0, n_items, H, kt_order, kt_kv0, kt_kvend, kt_q0, kt_qend, ia);
      ib = ia;
      bool has_b = has_a;
      auto issue_S = [&]() {                        // S^T(ga) = K Q^T
        const int b = ga % B2_QST, kvb = ka & 1;
        if (ia_i == 0) mbar_wait(&k_full[kvb], (ka >> 1) & 1);
        mbar_wait(&qdo_full[b], (ga / B2_QST) & 1);
        if (ga >= 1) mbar_wait(s_free, (ga - 1) & 1);
        tc_fence_after();
        const uint32_t aQ = smem_u32(sQDO + b * 32768), aK = smem_u32(sK + kvb * 16384);
        if (!(B2_VARIANT & 16))
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16_ss(tS, umma_smem_desc_sw128(aK + kk * 32, 0, 1024), umma_smem_desc_sw128(aQ + kk * 32, 0, 1024), idS, kk > 0 ? 1u : 0u);
        umma_commit(s_full);
      };
      auto issue_dP = [&]() {                       // dP^T(ga) = V dO^T, then advance cursor A
        const int b = ga % B2_QST;
        if (ia_i == 0) mbar_wait(v_full, ka & 1);
        tc_fence_after();
        const uint32_t aDO = smem_u32(sQDO + b * 32768 + 16384), aV = smem_u32(sV);
        if (!(B2_VARIANT & 16))
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16_ss(tDP, umma_smem_desc_sw128(aV + kk * 32, 0, 1024), umma_smem_desc_sw128(aDO + kk * 32, 0, 1024), idS, kk > 0 ? 1u : 0u);
        umma_commit(dp_full);
        ++ga;
        if (++ia_i == ia.n_q) {
          umma_commit(v_empty);                     // last dP^T of the item: V may be refilled with the next item's tile
          ia_i = 0; ++ka; has_a = b2_item(ka, n_items, H, kt_order, kt_kv0, kt_kvend, kt_q0, kt_qend, ia);
        }
      };
      if (has_a) { issue_S(); issue_dP(); }
      while (has_b) {
        if (has_a) issue_S();                       // S^T of the next step: overlaps this step's softmax
        const int b = gb % B2_QST, kvb = kb_ & 1;
        const uint32_t aQ = smem_u32(sQDO + b * 32768), aDO = aQ + 16384, aK = smem_u32(sK + kvb * 16384);
        // ---- dV(gb) += P^T dO
        mbar_wait(pt_full, gb & 1);
        if (ib_i == 0 && kb_ >= 1) mbar_wait(dkv_free, (kb_ - 1) & 1);     // the previous item's dK / dV have been read out of TMEM
        tc_fence_after();
        if (!(B2_VARIANT & 8))
#pragma unroll
        for (int kq = 0; kq < 8; ++kq)               // contraction over the 128 queries
          umma_bf16_ts(tDV, tPT + kq * 8, umma_smem_desc_sw128(aDO + kq * 2048, 8192, 1024), idT, (ib_i > 0 || kq > 0) ? 1u : 0u);
        // ---- dK(gb) += dS^T Q ;  dQ(gb) = dS K
        mbar_wait(ds_full, gb & 1);
        tc_fence_after();
        if (!(B2_VARIANT & 8))
#pragma unroll
        for (int kq = 0; kq < 8; ++kq)
          umma_bf16_ts(tDK, tDP + (kq >> 1) * 32 + (kq & 1) * 8, umma_smem_desc_sw128(aQ + kq * 2048, 8192, 1024), idT, (ib_i > 0 || kq > 0) ? 1u : 0u);      // dS^T of query quarter qc sits at dP^T + qc * 32 + [0, 16)
        if (!(B2_VARIANT & 1)) if (gb >= 1) { mbar_wait(dq_free, (gb - 1) & 1); tc_fence_after(); }
        if (!(B2_VARIANT & 1))
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)               // contraction over the 128 keys (rows of the dS^T tile)
          umma_bf16_ss(tDQ, umma_smem_desc_sw128(aDS + kk * 2048, 16384, 1024), umma_smem_desc_sw128(aK + kk * 2048, 8192, 1024), idQ, kk > 0 ? 1u : 0u);
        umma_commit(&qdo_empty[b]);
        umma_commit(grad_done);
        umma_commit(dq_full);
        ++gb;
        if (++ib_i == ib.n_q) {
          umma_commit(dkv_full);
          umma_commit(&k_empty[kvb]);
          ib_i = 0; ++kb_;
          has_b = b2_item(kb_, n_items, H, kt_order, kt_kv0, kt_kvend, kt_q0, kt_qend, ib);
        }
        if (has_a) issue_dP();                      // dP^T of the next step (its TMEM columns held dS^T of this one until dK above)
      }
    }
  } else if (warp == 2) {
    // ===================================================== dQ store lane: one TMA reduce-add pair per step, issued when all softmax warps have staged their columns
    if (lane == 0 && !(B2_VARIANT & 1)) {
      B2Item it;
      uint32_t g = 0;
      for (int k = 0; b2_item(k, n_items, H, kt_order, kt_kv0, kt_kvend, kt_q0, kt_qend, it); ++k) {
        for (int i = 0; i < it.n_q; ++i, ++g) {
          mbar_wait(dq_staged, g & 1);
          b2_tma_reduce_add_2d(&tmDQ, sDQ, it.head * 64, it.q_begin + i * 128);
          b2_tma_reduce_add_2d(&tmDQ, sDQ + 16384, it.head * 64 + 32, it.q_begin + i * 128);
          b2_bulk_commit();
          b2_bulk_wait_read0();
          mbar_arrive(dq_slab_free);
        }
      }
      b2_bulk_wait0();                              // all dQ reductions have been performed before the CTA retires
    }
  } else if (warp >= 4) {
    // ===================================================== softmax / gradient warps (thread <-> key row x 32 queries)
    const int quad = warp & 3;
    const int qc = (warp - 4) >> 2;                  // query-column quarter of S^T / dP^T; dQ / dK / dV column quarter in the read-outs
    const int row = quad * 32 + lane;                // key row of the tile (S^T, dP^T, dK, dV) / query row (dQ read-out)
    const int tid = threadIdx.x - 128;               // 0 .. 511
    const uint32_t lane_addr = uint32_t(quad * 32) << 16;
    const float k1 = scale / cap;
    const float KL = cap * 1.4426950408889634f;
    const float k2 = k1 * k1;
    const float a0 = KL * k1 * B2_C0, a1 = KL * k1 * k2 * B2_C1, a2 = KL * k1 * k2 * k2 * B2_C2, a3 = KL * k1 * k2 * k2 * k2 * B2_C3,
                a4 = KL * k1 * k2 * k2 * k2 * k2 * B2_C4;
    const float oms_c = -scale / (KL * KL);          // scale * (1 - tanh^2) = fma(e2^2, oms_c, scale)
    const float2 A0 = make_float2(a0, a0), A1 = make_float2(a1, a1), A2 = make_float2(a2, a2), A3 = make_float2(a3, a3), A4 = make_float2(a4, a4),
                 OC = make_float2(oms_c, oms_c), SC = make_float2(scale, scale);
    const int swz_row = (row >> 3) * 1024 + (row & 7) * 128;
    uint32_t g = 0;

    // dQ of step g_done: TMEM -> the shared staging tile; warp 2 (otherwise idle) issues the TMA reduce-add once all 16 warps have staged their columns.
    // Hand-offs are mbarriers (dq_slab_free / dq_staged): the 8-warp version's two CTA-wide named barriers per step were 20 % of its stall cycles, and
    // one small reduce per warp (no hand-off at all) was slower than two 16 KB reduces (295 vs 273 us).
    auto dq_readout = [&](uint32_t g_done, int qrow0, int hd) {
      mbar_wait(dq_full, g_done & 1);
      tc_fence_after();
      uint32_t r[16];
      tmem_ld_32x32b_x16(tDQ + lane_addr + qc * 16, r);
      if (g_done > 0) mbar_wait(dq_slab_free, (g_done - 1) & 1);      // the previous reduce has finished reading the staging tile
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
      uint8_t* dst = sDQ + (qc >> 1) * 16384 + swz_row;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        *reinterpret_cast<uint4*>(dst + ((((qc & 1) * 4 + ch) ^ (row & 7)) << 4)) = make_uint4(r[4 * ch], r[4 * ch + 1], r[4 * ch + 2], r[4 * ch + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_staged);
      (void)qrow0; (void)hd;
    };

    // per-query statistics of a step: fetched into registers of threads 0..127 (one query each) one step ahead, published in shared memory through
    // an mbarrier (meta_full): no CTA-wide barrier.  Reuse of a buffer two steps later is ordered by the ds_full -> grad_done chain (every warp
    // has finished reading step g - 1's statistics before any warp passes the grad_done wait of step g).
    float m_lse = 0.f, m_D = 0.f; int m_lim = -1;
    auto fetch_meta = [&](int q_begin, int q_end, int head, int i) {
      if (tid < 128) {
        const int gr = q_begin + i * 128 + tid;
        const bool ok = gr < q_end;
        m_lim = ok ? kv_limit[gr] : -1;
        m_lse = ok ? lse[(long long)head * M + gr] : 0.f;
        m_D = ok ? dsum[(long long)head * M + gr] : 0.f;
      }
    };
    auto stage_meta = [&](int buf) {                 // stored negated (the consumers only add) and as limit + 1 (the mask is `key < limit + 1`)
      if (tid < 128) {
        float* mb = sMeta + buf * 512;
        mb[tid] = -m_lse * 1.4426950408889634f; mb[128 + tid] = -m_D; reinterpret_cast<int*>(mb)[256 + tid] = m_lim + 1;
        const int wmin = __reduce_min_sync(0xffffffffu, m_lim);
        if (lane == 0) { reinterpret_cast<int*>(mb)[384 + (tid >> 5)] = wmin; }
        __syncwarp();
        if (lane == 0) mbar_arrive(&meta_full[buf]);
      }
    };

#if B2_PIPE
    // ---- software-pipelined order (B2_PIPE): the exponentials of step g + 1 need only S^T(g + 1), which the MMA warp produces while step g is still in
    // its dS^T phase - so they are computed right after dS^T(g) has been handed over, in the shadow of dK(g) / dQ(g) / dP^T(g + 1), and P^T and the
    // scale (1 - tanh^2) factor wait in registers as bf16 pairs until the gradient products of step g have released the P^T columns.  The MMA warp's
    // order is unchanged; s_free arrives one phase earlier than in the unpipelined order.
    auto advance = [&](B2Item& t, int& kk, int& ii) -> bool {      // next (key tile item, query tile) step of this CTA
      if (++ii < t.n_q) return true;
      ii = 0; ++kk;
      return b2_item(kk, n_items, H, kt_order, kt_kv0, kt_kvend, kt_q0, kt_qend, t);
    };
    uint32_t wp[16], om[16];                           // P^T and scale * (1 - tanh^2) of the step in flight: this thread's key row x 32 queries, bf16 pairs
    auto phase_A = [&](uint32_t ga, int kv0) {
      const float* mb = sMeta + (ga & 1) * 512;
      const int* mbi = reinterpret_cast<const int*>(mb);
      mbar_wait(&meta_full[ga & 1], (ga >> 1) & 1);
      const int min_lim = min(min(mbi[384], mbi[385]), min(mbi[386], mbi[387]));
      const bool all_visible = kv0 + 127 <= min_lim;
      const int key = kv0 + row;
      mbar_wait(s_full, ga & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col0 = qc * 32 + c * 16;
        uint32_t rs[16];
        tmem_ld_32x32b_x16(tS + lane_addr + col0, rs);
        tmem_ld_wait();
        if (c == 1) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(s_free); }       // S^T is in registers: S^T of the next step may be issued
        if (all_visible) b2_exp_chunk_om<false>(rs, mb + col0, mbi + 256 + col0, key, A0, A1, A2, A3, A4, OC, SC, om + c * 8, wp + c * 8);
        else b2_exp_chunk_om<true>(rs, mb + col0, mbi + 256 + col0, key, A0, A1, A2, A3, A4, OC, SC, om + c * 8, wp + c * 8);
      }
    };
    auto dkv_readout = [&](const B2Item& t, int kk) {   // dK (fp32) and dV (bf16) of a finished key tile: 16 of the 64 columns per warp
      mbar_wait(dkv_full, kk & 1);
      tc_fence_after();
      uint32_t r[16], r2[16];
      tmem_ld_32x32b_x16(tDK + lane_addr + qc * 16, r);
      tmem_ld_32x32b_x16(tDV + lane_addr + qc * 16, r2);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dkv_free);          // the accumulators may be overwritten by the next item
      const int key = t.kv0 + row;
      if (key < t.kv_end) {
        float* dst = dk + (long long)key * H * 64 + t.head * 64 + qc * 16;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) *reinterpret_cast<uint4*>(dst + ch * 4) = make_uint4(r[4 * ch], r[4 * ch + 1], r[4 * ch + 2], r[4 * ch + 3]);
        __nv_bfloat16* dst2 = dv + (long long)key * ld_dv + t.head * 64 + qc * 16;
#pragma unroll
        for (int qd = 0; qd < 2; ++qd) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) w[e] = pack_bf16(__uint_as_float(r2[qd * 8 + 2 * e]), __uint_as_float(r2[qd * 8 + 2 * e + 1]));
          *reinterpret_cast<uint4*>(dst2 + qd * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    };
    // cursors: M = the step whose statistics are in registers (one ahead of the exponentials), B = the step of the gradient phase
    B2Item itM, itB;
    int kM = 0, iM = 0, kB = 0, iB = 0;
    uint32_t gA = 0, gB = 0;
    bool hasM = b2_item(0, n_items, H, kt_order, kt_kv0, kt_kvend, kt_q0, kt_qend, itM);
    bool hasB = hasM, pending = false;
    itB = itM;
    if (hasM) {
      fetch_meta(itM.q_begin, itM.q_end, itM.head, 0); stage_meta(0);
      const int kv0 = itM.kv0;
      hasM = advance(itM, kM, iM);
      if (hasM) fetch_meta(itM.q_begin, itM.q_end, itM.head, iM);       // statistics of step 1 travel through registers
      phase_A(0, kv0);
    }
    while (hasB) {
      // ---- P^T(gB) into TMEM once dV / dK / dQ of the previous step have consumed P^T and dS^T (TMEM and smem)
      if (gB > 0) { mbar_wait(grad_done, (gB - 1) & 1); tc_fence_after(); }
      tmem_st_32x32b_x16(tPT + lane_addr + qc * 16, wp);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pt_full);
      if (pending) dq_readout(gB - 1, 0, 0);            // dQ of the previous step (complete together with grad_done) leaves while dP^T lands
      // ---- dS^T(gB) = P^T o (dP^T - D) o scale (1 - tanh^2)
      mbar_wait(dp_full, gB & 1);
      tc_fence_after();
      const float* mbB = sMeta + (gB & 1) * 512;
      uint32_t wd[16];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col0 = qc * 32 + c * 16;
        uint32_t rp[16];
        tmem_ld_32x32b_x16(tDP + lane_addr + col0, rp);
        tmem_ld_wait();
#pragma unroll
        for (int e2 = 0; e2 < 16; e2 += 2) {
          const int j = c * 8 + (e2 >> 1);
          const float2 pf = unpack2_bf16(wp[j]);                                                  // the SAME bf16 P^T the dV product sees
          const float2 oms = unpack2_bf16(om[j]);
          const float2 dpd = __fadd2_rn(make_float2(__uint_as_float(rp[e2]), __uint_as_float(rp[e2 + 1])), *reinterpret_cast<const float2*>(mbB + 128 + col0 + e2));   // + (-D)
          const float2 d = __fmul2_rn(__fmul2_rn(pf, dpd), oms);
          wd[j] = pack_bf16(d.x, d.y);
        }
      }
      tmem_st_32x32b_x16(tDP + lane_addr + qc * 32, wd);
      {
        uint8_t* db = sDS + (qc >> 1) * 16384 + swz_row;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
          *reinterpret_cast<uint4*>(db + ((((qc & 1) * 4 + ch) ^ (row & 7)) << 4)) = make_uint4(wd[4 * ch], wd[4 * ch + 1], wd[4 * ch + 2], wd[4 * ch + 3]);
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
      pending = true;
      // ---- exponentials of the next step, in the shadow of this step's gradient products
      if (hasM) {                                       // cursor M sits on step gA + 1: its statistics are in registers
        ++gA;
        stage_meta(gA & 1);                             // (the buffer was last read two steps back, before the grad_done wait above)
        const int kv0 = itM.kv0;
        hasM = advance(itM, kM, iM);
        if (hasM) fetch_meta(itM.q_begin, itM.q_end, itM.head, iM);
        phase_A(gA, kv0);
      }
      if (iB + 1 == itB.n_q) dkv_readout(itB, kB);      // last query tile of this key tile
      hasB = advance(itB, kB, iB);
      ++gB;
    }
    if (pending) dq_readout(gB - 1, 0, 0);
#else
    B2Item it, nx;
    bool has = b2_item(0, n_items, H, kt_order, kt_kv0, kt_kvend, kt_q0, kt_qend, it);
    if (has) { fetch_meta(it.q_begin, it.q_end, it.head, 0); stage_meta(0); }
    int prev_qrow0 = 0, prev_head = 0;
    bool pending = false;
    for (int k = 0; has; ++k) {
      const bool has_n = b2_item(k + 1, n_items, H, kt_order, kt_kv0, kt_kvend, kt_q0, kt_qend, nx);
      const int kv0 = it.kv0;
      const int key = kv0 + row;
      for (int i = 0; i < it.n_q; ++i, ++g) {
        const float* mb = sMeta + (g & 1) * 512;
        const int* mbi = reinterpret_cast<const int*>(mb);
        // statistics of the NEXT step travel through registers while this one is processed
        if (i + 1 < it.n_q) fetch_meta(it.q_begin, it.q_end, it.head, i + 1); else if (has_n) fetch_meta(nx.q_begin, nx.q_end, nx.head, 0);
        mbar_wait(&meta_full[g & 1], (g >> 1) & 1);
        const int min_lim = min(min(mbi[384], mbi[385]), min(mbi[386], mbi[387]));
        const bool all_visible = kv0 + 127 <= min_lim;      // every query of the tile sees every key of the tile: no mask
        // Per 16-query chunk: exponentials from S^T (dP^T of this step may still be in flight), P^T back to TMEM, then dS^T from dP^T.
        uint32_t wd[16];                                    // dS^T of this thread's key row x 32 queries, bf16 pairs
        mbar_wait(s_full, g & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col0 = qc * 32 + c * 16;
          uint32_t wp[8];                                   // P^T chunk, bf16 pairs
          float2 ee[8];                                     // cap * log2e * tanh(y) of the same scores (for the 1 - tanh^2 factor)
          {
            uint32_t rs[16];
            tmem_ld_32x32b_x16(tS + lane_addr + col0, rs);
            tmem_ld_wait();
            if (c == 1) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(s_free); }      // S^T is in registers: S^T of the next step may be issued
            // the visibility mask is only needed on diagonal / span-boundary tiles: two specialised code paths (see attention_fwd_sm100.cu)
            if (B2_VARIANT & 4) {
#pragma unroll
              for (int j = 0; j < 8; ++j) { wp[j] = rs[2 * j] ^ rs[2 * j + 1]; ee[j] = make_float2(__uint_as_float(rs[2 * j]), __uint_as_float(rs[2 * j + 1])); }
            } else if (all_visible) b2_exp_chunk<false>(rs, mb + col0, mbi + 256 + col0, key, A0, A1, A2, A3, A4, ee, wp);
            else b2_exp_chunk<true>(rs, mb + col0, mbi + 256 + col0, key, A0, A1, A2, A3, A4, ee, wp);
          }
          if (c == 0 && g > 0) { mbar_wait(grad_done, (g - 1) & 1); tc_fence_after(); }    // dV / dK / dQ of the previous step have consumed P^T, dS^T (TMEM and smem)
          tmem_st_32x32b_x8(tPT + lane_addr + qc * 16 + c * 8, wp);
          if (c == 1) {                                     // P^T complete: dV may start while dS^T is still being computed
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(pt_full);
          }
          if (c == 0) { mbar_wait(dp_full, g & 1); tc_fence_after(); }
          uint32_t rp[16];
          tmem_ld_32x32b_x16(tDP + lane_addr + col0, rp);
          tmem_ld_wait();
          if (B2_VARIANT & 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) wd[c * 8 + j] = rp[2 * j] ^ rp[2 * j + 1] ^ wp[j] ^ __float_as_uint(ee[j].x);
          } else
#pragma unroll
          for (int e2 = 0; e2 < 16; e2 += 2) {
            const int j = e2 >> 1;
            const float2 e = ee[j];
            const float2 pf = unpack2_bf16(wp[j]);                                                // the SAME bf16 P^T the dV product sees
            const float2 oms = __ffma2_rn(__fmul2_rn(e, e), OC, SC);
            const float2 dpd = __fadd2_rn(make_float2(__uint_as_float(rp[e2]), __uint_as_float(rp[e2 + 1])), *reinterpret_cast<const float2*>(mb + 128 + col0 + e2));   // + (-D)
            const float2 d = __fmul2_rn(__fmul2_rn(pf, dpd), oms);
            wd[c * 8 + j] = pack_bf16(d.x, d.y);
          }
        }
        // dS^T over this warp's own dP^T columns (TS operand of dK) and into the shared tile (MN-major A operand of dQ)
        tmem_st_32x32b_x16(tDP + lane_addr + qc * 32, wd);
        {
          uint8_t* db = sDS + (qc >> 1) * 16384 + swz_row;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            *reinterpret_cast<uint4*>(db + ((((qc & 1) * 4 + ch) ^ (row & 7)) << 4)) = make_uint4(wd[4 * ch], wd[4 * ch + 1], wd[4 * ch + 2], wd[4 * ch + 3]);
        }
        tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ds_full);
        if (pending && !(B2_VARIANT & 1)) dq_readout(g - 1, prev_qrow0, prev_head);
        prev_qrow0 = it.q_begin + i * 128; prev_head = it.head; pending = true;
        stage_meta((g + 1) & 1);                             // statistics of the next step (fetched above); the buffer was last read in step g - 1
      }
      // ---- dK (fp32) and dV (bf16) of this key tile: 16 of the 64 columns per warp
      mbar_wait(dkv_full, k & 1);
      tc_fence_after();
      {
        uint32_t r[16], r2[16];
        tmem_ld_32x32b_x16(tDK + lane_addr + qc * 16, r);
        tmem_ld_32x32b_x16(tDV + lane_addr + qc * 16, r2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dkv_free);        // the accumulators may be overwritten by the next item
        if (key < it.kv_end) {
          float* dst = dk + (long long)key * H * 64 + it.head * 64 + qc * 16;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) *reinterpret_cast<uint4*>(dst + ch * 4) = make_uint4(r[4 * ch], r[4 * ch + 1], r[4 * ch + 2], r[4 * ch + 3]);
          __nv_bfloat16* dst2 = dv + (long long)key * ld_dv + it.head * 64 + qc * 16;
#pragma unroll
          for (int qd = 0; qd < 2; ++qd) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) w[e] = pack_bf16(__uint_as_float(r2[qd * 8 + 2 * e]), __uint_as_float(r2[qd * 8 + 2 * e + 1]));
            *reinterpret_cast<uint4*>(dst2 + qd * 8) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      it = nx; has = has_n;
    }
    if (pending && !(B2_VARIANT & 1)) dq_readout(g - 1, prev_qrow0, prev_head);
  #endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace tfx

namespace tfx {
// fp32 2-D tensor map with a 32-float (128 B, swizzled) x box_rows box - the destination of the dQ TMA reduce-add
static int b2_make_tmap_f32_sw128(CUtensorMap* tm, const void* ptr, long long inner, long long outer, long long ld, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return -1;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}
}  // namespace tfx

using namespace tfx;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int tfx_attn_bwd_ts(const void* q, const void* k, const void* v, const void* do_pre, long long ld_q, long long ld_k, long long ld_v, long long ld_do,
                    const float* lse, const float* dsum_hm, const int* kv_limit, const int* kt_kv0, const int* kt_kvend, const int* kt_q0, const int* kt_qend,
                    const int* kt_order, int n_kv_tiles, float* dq, float* dk, void* dv, long long ld_dv, int M, int H, float scale, float softcap, const float* fast_params,
                    void* stream) {
  if (n_kv_tiles <= 0) return 0;
  TFX_REQUIRE(fast_params != nullptr, "attn_bwd_ts: fast_params (from tfx_attn_fast_params) is required");
  TFX_REQUIRE(ld_q % 8 == 0 && ld_k % 8 == 0 && ld_v % 8 == 0 && ld_do % 8 == 0 && ld_dv % 8 == 0, "attn_bwd_ts: row pitches must be multiples of 8 bf16");
  CUtensorMap tq, tk, tv, tdo, tdq;
  int rc;
  if ((rc = make_tmap_bf16(&tq, q, (long long)H * 64, M, ld_q, 128)) || (rc = make_tmap_bf16(&tk, k, (long long)H * 64, M, ld_k, 128)) ||
      (rc = make_tmap_bf16(&tv, v, (long long)H * 64, M, ld_v, 128)) || (rc = make_tmap_bf16(&tdo, do_pre, (long long)H * 64, M, ld_do, 128)) ||
      (rc = b2_make_tmap_f32_sw128(&tdq, dq, (long long)H * 64, M, (long long)H * 64, 128))) {
    set_error("attn_bwd_ts: cuTensorMapEncodeTiled failed (%d)", rc);
    return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(attn_bwd_ts_k, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_SMEM) != cudaSuccess) { set_error("attn_bwd_ts: cannot raise dynamic smem"); return -2; }
    attr_set = true;
  }
  const int n_items = n_kv_tiles * H;
  const int grid = n_items < num_sms() ? n_items : num_sms();          // persistent: one CTA per SM
  attn_bwd_ts_k<<<grid, B2_THREADS, B2_SMEM, ST(stream)>>>(tq, tk, tv, tdo, tdq, lse, dsum_hm, kv_limit, kt_kv0, kt_kvend, kt_q0, kt_qend, kt_order, n_items, dk,
                                                          (__nv_bfloat16*)dv, ld_dv, M, H, scale, softcap, fast_params);
  return check_launch("attn_bwd_ts");
}

}  // extern "C"
