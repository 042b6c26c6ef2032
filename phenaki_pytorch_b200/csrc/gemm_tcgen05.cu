// bf16 GEMM on the 5th-generation tensor cores (sm_100a only):
//   C[map(m), n] = sum_k A[m,k] * W[n,k]  (+bias) (+residual)        nn.Linear semantics
// A, W bf16, K-major (row-major [rows, K]); fp32 accumulation in TMEM.
//
// Three kernels behind phk_gemm_bf16 / phk_gemm_bf16_x2 (dispatch rule and measurements at the host entry points):
//   gemm_bf16_kernel<EPI, DUAL>            one CTA per 128 x 128 tile, tcgen05.mma.cta_group::1
//   gemm_bf16_pair_kernel<EPI, BN, DUAL>   clusters of two CTAs per 256 x BN tile (BN = 128 / 256), cta_group::2
//   DUAL                                    two independent problems in one persistent launch
// All are persistent and warp-specialised (one CTA per SM, static round-robin tile scheduler, m-fastest so concurrent
// CTAs share the W tile in L2 while the A panel stays L2-resident):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d (SWIZZLE_128B) stages 128x64 A and W tiles into a 4..6-deep
//               shared-memory ring; out-of-bounds rows / the K tail are zero-filled by the TMA unit.
//   warp 1      TMEM allocator + MMA issuer: one thread issues tcgen05.mma.kind::f16 (K = 16), four per k-block;
//               tcgen05.commit frees the smem slot and publishes the accumulator.  Two accumulators in TMEM, so tile
//               i+1 is multiplied while tile i drains.
//   warps 2..17 epilogue (four per TMEM lane group, each drains a quarter of the columns).  Sixteen warps because the
//               epilogue math is a per-warp dependency chain: with eight the MMA warp mostly waited for a drained
//               accumulator.
// Epilogues: 0 fp32 (+bias, +residual, row map) -- through the TMA unit when the row map is the identity: the
//              residual tile is bulk-loaded into SWIZZLE_128B staging boxes during the main loop and the result tile
//              bulk-stored from them; otherwise staged in a padded smem tile and written with coalesced rows
//            1 bf16 (+bias)
//            2 GEGLU (attention.py:40-43) on W rows packed [64 value rows | 64 gate rows] per 128-column tile ->
//              bf16 [M, N/2]: accumulator read out of TMEM and released at once, fitted sigmoid-form GELU, swizzled
//              bf16 staging tile, coalesced 16-byte stores.
#include "phk_common.cuh"
#include "phk_sm100.cuh"
#include <cstdlib>
#include <mutex>
#include <unordered_map>

namespace phk {

constexpr int GM = 128;       // BLOCK_M = UMMA_M
constexpr int GN = 128;       // BLOCK_N = UMMA_N
constexpr int GK = 64;        // BLOCK_K: 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int GSTAGES = 4;
constexpr int EPI_WARPS = 16;       // epilogue warps 2..17: four per TMEM lane group, each drains a quarter of the columns
constexpr int EPI_PARTS = EPI_WARPS / 4;            // column parts per lane group
constexpr int GTHREADS = 64 + EPI_WARPS * 32;       // 576
constexpr int STAGE_BYTES = GM * GK * 2;          // 16 KB per operand per stage
constexpr int CPAD = 132;                          // fp32 staging row stride (floats): conflict-free 128-bit rows
constexpr int CSTAGE_BYTES = GM * CPAD * 4;        // 67.6 KB
constexpr int RING_BYTES = GSTAGES * 2 * STAGE_BYTES;
constexpr int SMEM_TOTAL = RING_BYTES + CSTAGE_BYTES + 256 /*barriers*/ + 1024 /*manual 1024-B alignment*/;
// epilogue 4: per-part row sums [4][128] float2 + the cluster exchange [2 buffers][8 ranks][128 rows] float2
constexpr int LN_MAX_CLUSTER = 8;
constexpr int LN_ROWSTAT_BYTES = EPI_PARTS * GM * 8, LN_XSTAT_BYTES = 2 * LN_MAX_CLUSTER * GM * 8;
constexpr int SMEM_TOTAL_LN = SMEM_TOTAL + LN_ROWSTAT_BYTES + LN_XSTAT_BYTES;

struct EpiParams {
  void* C; int64_t ldc; int64_t M; int N; int K;
  const float* bias; const float* residual;
  int64_t seg_len, seg_stride, seg_off;
  int m_tiles, n_tiles;
  long long* trace;  // debug: per-CTA clock64 stamps of the first tile (NULL in production)
  // fp32 epilogue through the TMA unit (identity row map, C 16-B aligned, residual == C or none): C as a tensor map
  // with [128 rows x 32 floats] SWIZZLE_128B boxes; the residual tile is bulk-LOADED into the staging boxes while the
  // main loop runs and the finished tile is bulk-STORED from them, so the epilogue warps issue no global accesses
  int tma_epi;
  // epilogue 3 (self-attention operands, attention.py:146-157): bf16 output; columns [0, norm_cols) are l2-normalised
  // per 64-column head (F.normalize, eps 1e-12), multiplied by nscale[col % 64] (q_scale / k_scale) and by nmul (the
  // fixed similarity scale 8 folded into q); columns >= norm_cols (the value half of to_kv) are only converted
  const float* nscale; int norm_cols; float nmul;
  // epilogue 4 (residual + the NEXT LayerNorm, attention.py:311-332): C = A W^T + C in place (fp32, bulk-stored) and, from
  // the same registers, LayerNorm(C) * ln_g + ln_b -> bf16 ln_out [M, ln_ld] (+ optionally the un-normalised row as bf16
  // in raw_out: the k,v projection input of attention.py:140-144).  One CTA holds 128 of the N columns of a row, so the
  // row statistics are summed over the cluster of n_tiles CTAs that share the m-tile (distributed shared memory).
  const float* ln_g; const float* ln_b; void* ln_out; void* raw_out; int64_t ln_ld; float ln_eps;
  // epilogue 5: the same result without a cluster -- the n_tiles CTAs of a group (consecutive block indices, all resident:
  // the grid is at most one CTA per SM) exchange their 128-column row statistics through global memory: xstat_g
  // [2][gridDim][128] float2 slots (double-buffered by tile parity) and one arrival counter per group (ln_counter[group],
  // zero when the kernel starts; release / acquire at gpu scope).
  float2* xstat_g; unsigned int* ln_counter;
  // the W operand is not written by the kernels that precede this one in the stream (inference weights): its first tiles
  // are requested BEFORE griddepcontrol.wait, while the previous kernel drains
  int w_static;
  alignas(64) CUtensorMap tmC;
};


__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory"); }

// gelu_erf(g) * v for the bf16 GEGLU epilogue (attention.py:40-43).  Phi(g) = 0.5 (1 + erf(g / sqrt2)) is evaluated
// as sigmoid(2 g (a + b g^2 + c g^4)) = 1 / (1 + 2^(g (a' + b' g^2 + c' g^4))) with (a, b, c) fitted to the erf form
// (max |error| of the GELU 2.5e-5 over all g -- two orders below the bf16 rounding of the result -- and, unlike
// 0.5 (1 + tanh), no cancellation for negative g).  10 issue slots per output (two of them MUFU) instead of ~19 for
// the erf polynomial: the epilogue was issue-bound (tools/gemm_trace.py).  g^2 is clamped to 64: beyond |g| = 8 the
// fitted quintic would turn over, while the sigmoid is already saturated.
__device__ __forceinline__ float geglu_fast(float g, float v) {
  const float g2 = fminf(g * g, 64.0f);
#ifndef PHK_GEGLU_EX2_RCP
  // sigmoid(z) = 0.5 + 0.5 tanh(z / 2) on ONE MUFU op (tanh.approx.f32, relative error 2^-11) instead of ex2 + rcp: the same
  // fitted quintic with its coefficients times -ln2 / 2.  The epilogue's math phase was MUFU-paced (2 ops x 32 outputs x 16
  // warps per tile = 2048 of its ~4200 cycles): FF1 + GEGLU 15.6 -> 14.7 us, encode 0.675 -> 0.666 ms
  // (profiles/r02/ab_geglu_tanh_c20.txt).  Phi is good to 2.4e-4 absolute -- below the bf16 rounding of the hidden
  // activations; the at-size parity figures did not move (cfg3 logits mean |err| 0.001215 in both builds, 83 vs 87 flipped
  // LFQ bits).  -DPHK_GEGLU_EX2_RCP restores the two-op form (2.5e-5, and relative accuracy in the negative tail).
  float u = fmaf(g2, -0.00035151678851506f, 0.037005646021930225f);
  u = fmaf(u, g2, 0.7975078842858219f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(g * u));
  const float h = 0.5f * (g * v);
  return fmaf(h, t, h);
#else
  float u = fmaf(g2, 0.0010142630551597833f, -0.10677572400146226f);   // -2 log2(e) * {c, b, a}
  u = fmaf(u, g2, -2.301121339458009f);
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(g * u));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return (g * v) * r;
#endif
}

// ---------------------------------------------------------------------------------------------------
// Epilogue pieces shared by the one-CTA and the CTA-pair kernels.  A "chunk" is 128 rows x 128 accumulator columns
// of this CTA: TMEM lanes 0..127, columns [tcol, tcol + 128).
// ---------------------------------------------------------------------------------------------------

// Residual prefetch: while the main loop of the tile runs, pull the residual chunk into the staging buffer with
// coalesced 512-B row loads (the previous chunk's write-out finished at its closing barrier).
template <int EPI>
__device__ __forceinline__ bool epi_residual_prefetch(const EpiParams& p, float* cstage, int64_t m0, int n0, int ew,
                                                      int lane) {
  const bool res_vec = EPI == 0 && p.residual && (p.ldc % 4 == 0) && (p.N % 4 == 0) &&
                       ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0);
  if (res_vec) {
    const int col = n0 + lane * 4;
    const uint32_t seg_len = (uint32_t)p.seg_len;
#pragma unroll 1
    for (int rb = 0; rb < GM / EPI_WARPS; rb += 8) {
      float4 rv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = (rb + u) * EPI_WARPS + ew;
        const uint32_t m = (uint32_t)m0 + r;
        uint32_t orow = m;
        if (seg_len > 0) { const uint32_t q = m / seg_len; orow = q * (uint32_t)p.seg_stride + (uint32_t)p.seg_off + (m - q * seg_len); }
        rv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < (uint32_t)p.M && col < p.N) rv[u] = *reinterpret_cast<const float4*>(p.residual + (int64_t)orow * p.ldc + col);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        *reinterpret_cast<float4*>(cstage + ((rb + u) * EPI_WARPS + ew) * CPAD + lane * 4) = rv[u];
    }
    epi_bar_sync();  // residual chunk complete before the row-per-thread accumulate
  }
  return res_vec;
}

// GEGLU epilogue of a whole tile of NCH 128-column chunks, each packed [64 value | 64 gate] (attention.py:40-43).
// All accumulator columns this thread needs (16 value + 16 gate per chunk of ONE row) are pulled out of TMEM first and
// the accumulator is released at once, so the MMA warp never waits for epilogue math.  The bf16 results (NCH x 128 B
// per row) go through a swizzled shared-memory tile and leave as fully coalesced 16-B-per-lane rows: one row per
// thread straight from registers made every warp store touch 32 different lines (~8 clk per line in the LSU).
template <int NCH, typename Release, typename Stamp>
__device__ __forceinline__ void epi_geglu_tile(const EpiParams& p, void* stage, uint32_t tmem_tile, int64_t m0, int n0,
                                               int ew, int lg, int part, int lane, Release&& release, Stamp&& stamp) {
  constexpr int W = 64 / EPI_PARTS;
  static_assert(W == 16, "two 16-B units per thread and chunk");
  constexpr int U = NCH * 8;  // 16-B units per staged row
  const uint32_t trow = tmem_tile + ((uint32_t)(lg * 32) << 16);
  uint32_t val[NCH][W], gate[NCH][W];
#pragma unroll
  for (int h = 0; h < NCH; ++h) {
    tmem_ld16(trow + h * 128 + part * W, val[h]);
    tmem_ld16(trow + h * 128 + 64 + part * W, gate[h]);
  }
  tmem_ld_wait();
  stamp(10);  // accumulator columns in registers
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  release();
  const uint32_t seg_len = (uint32_t)p.seg_len;
  const bool coalesced = (p.ldc % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
  const int r = lg * 32 + lane;  // this thread's row of the tile
  uint4* srow = reinterpret_cast<uint4*>(stage) + r * U;
#pragma unroll
  for (int h = 0; h < NCH; ++h) {
    uint32_t pk[W / 2];
#pragma unroll
    for (int j = 0; j < W; j += 2)
      pk[j / 2] = pack_bf16x2(geglu_fast(__uint_as_float(gate[h][j]), __uint_as_float(val[h][j])),
                              geglu_fast(__uint_as_float(gate[h][j + 1]), __uint_as_float(val[h][j + 1])));
    if (coalesced) {
      // logical unit u of row r lives at (u & ~7) | ((u ^ r) & 7): conflict-free for the row-per-lane stores here
      // and for the unit-per-lane loads below
      const int u0 = h * 8 + part * 2;
      srow[(u0 & ~7) | ((u0 ^ r) & 7)] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      srow[((u0 + 1) & ~7) | (((u0 + 1) ^ r) & 7)] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    } else {  // odd leading dimension / unaligned C: 4-byte stores from registers
      const uint32_t m = (uint32_t)m0 + r;
      const int col = (n0 + h * 128) / 2 + part * W;
      if (m < (uint32_t)p.M && col < p.N / 2) {
        int64_t orow = m;
        if (seg_len > 0) { const uint32_t q = m / seg_len; orow = q * (uint32_t)p.seg_stride + (uint32_t)p.seg_off + (m - q * seg_len); }
        __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(p.C) + orow * p.ldc + col;
#pragma unroll
        for (int j = 0; j < W / 2; ++j) *reinterpret_cast<uint32_t*>(crow + 2 * j) = pk[j];
      }
    }
  }
  if (!coalesced) return;  // uniform over the CTA
  stamp(11);               // this thread's outputs computed and staged
  epi_bar_sync();          // tile staged
  stamp(7);
  constexpr int RPI = 32 / U;                 // rows per warp instruction (4 or 2)
  const int rl = lane / U, u = lane % U;      // this lane's row within the instruction and its 16-B unit
  const bool chunk_ok = (n0 + (u >> 3) * 128) < p.N;
#pragma unroll
  for (int i = 0; i < GM / (EPI_WARPS * RPI); ++i) {
    const int rr = (i * EPI_WARPS + ew) * RPI + rl;
    const uint32_t m = (uint32_t)m0 + rr;
    const uint4 v = reinterpret_cast<const uint4*>(stage)[rr * U + ((u & ~7) | ((u ^ rr) & 7))];
    if (m < (uint32_t)p.M && chunk_ok) {
      int64_t orow = m;
      if (seg_len > 0) { const uint32_t q = m / seg_len; orow = q * (uint32_t)p.seg_stride + (uint32_t)p.seg_off + (m - q * seg_len); }
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + orow * p.ldc + n0 / 2 + u * 8) = v;
    }
  }
  epi_bar_sync();          // staging tile free for the next accumulator
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// fp32 chunk (128 rows x 128 columns) through the TMA unit.  Staging = four [128 rows x 128 B] SWIZZLE_128B boxes (one
// per 32-column part): thread (row r, part c) owns exactly one 128-byte swizzled row of box c, so its eight 16-byte
// accesses are conflict-free and need no other thread's data.  Order per chunk:
//   store thread: wait until the previous chunk's bulk stores have READ the boxes; barrier;
//   store thread: bulk-load the residual tile (C itself: x = f(x) + x) into the boxes (mbarrier `bar_res`);
//   all: accumulator columns TMEM -> registers, + residual (own swizzled row) + bias, back into the boxes; release the
//        accumulator; fence.proxy.async; barrier;  store thread: four bulk stores + commit.
// part 1, BEFORE the accumulator-ready wait (overlaps the main loop)
__device__ __forceinline__ void epi_tma_begin(const EpiParams& p, uint32_t stage_s, uint32_t bar_res, int m0, int n0,
                                              int ew, int lane) {
  const bool store_thread = ew == 0 && lane == 0;
  if (store_thread) tma_store_wait_read();
  epi_bar_sync();  // the boxes are free (the previous chunk's bulk stores have read them)
  if (p.residual && store_thread) {
    mbar_expect_tx(bar_res, 4 * GM * 128);
#pragma unroll
    for (int c = 0; c < 4; ++c) tma_load_2d(&p.tmC, bar_res, stage_s + c * (GM * 128), n0 + 32 * c, m0);
  }
}
// part 2, after the accumulator-ready wait
template <typename Release>
__device__ __forceinline__ void epi_tma_finish(const EpiParams& p, uint8_t* stage, uint32_t stage_s, uint32_t bar_res,
                                               uint32_t& res_phase, uint32_t tmem_chunk, int m0, int n0, int ew, int lg,
                                               int part, int lane, Release&& release) {
  const uint32_t trow = tmem_chunk + ((uint32_t)(lg * 32) << 16);
  uint32_t v[32];
  tmem_ld32(trow + part * 32, v);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  release();
  const int r = lg * 32 + lane;
  float4* box_row = reinterpret_cast<float4*>(stage + part * (GM * 128) + r * 128);
  if (p.residual) { mbar_wait(bar_res, res_phase); res_phase ^= 1; }
  const int col0 = n0 + part * 32;
  const bool bias_vec = p.bias && (col0 + 31 < p.N) && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                           __uint_as_float(v[4 * j + 3]));
    const int pos = j ^ (r & 7);
    if (p.residual) { const float4 rr = box_row[pos]; o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w; }
    if (bias_vec) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
      o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
    } else if (p.bias) {
      const int c = col0 + 4 * j;
      if (c < p.N) o.x += __ldg(p.bias + c);
      if (c + 1 < p.N) o.y += __ldg(p.bias + c + 1);
      if (c + 2 < p.N) o.z += __ldg(p.bias + c + 2);
      if (c + 3 < p.N) o.w += __ldg(p.bias + c + 3);
    }
    box_row[pos] = o;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA unit
  epi_bar_sync();  // whole chunk staged
  if (ew == 0 && lane == 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (n0 + 32 * c < p.N) tma_store_2d(&p.tmC, stage_s + c * (GM * 128), n0 + 32 * c, m0);  // OOB rows / columns are clipped
    tma_store_commit();
  }
}

// Drains one chunk and writes it out.  `release()` is called as soon as the TMEM columns have been read (the MMA
// warp may then overwrite them), before the slower global write-out.
template <int EPI, typename Release>
__device__ __forceinline__ void epi_chunk(const EpiParams& p, float* cstage, uint32_t tmem_chunk, int64_t m0, int n0,
                                          bool res_vec, int ew, int lg, int part, int lane, Release&& release) {
  const uint32_t trow = tmem_chunk + ((uint32_t)(lg * 32) << 16);
  float* srow = cstage + (lg * 32 + lane) * CPAD;
  const uint32_t seg_len = (uint32_t)p.seg_len;
  {
    const int c = part;  // 32 of the 128 columns
    uint32_t v[32];
    tmem_ld32(trow + c * 32, v);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                             __uint_as_float(v[j + 3]));
      if (res_vec) {  // accumulate onto the prefetched residual
        const float4 rr = *reinterpret_cast<const float4*>(srow + c * 32 + j);
        o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
      }
      *reinterpret_cast<float4*>(srow + c * 32 + j) = o;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  release();
  epi_bar_sync();  // whole 128 x 128 chunk staged

  // ---- coalesced write-out: one row per warp instruction ----
  constexpr int RPW = GM / EPI_WARPS;  // rows per warp
  const int ncol0 = n0;
  const int nlim = p.N;
  if (EPI == 0) {
    const int col = ncol0 + lane * 4;
    const bool vec = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && (col + 3 < nlim);
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) {
      if (col < nlim) bv.x = __ldg(p.bias + col);
      if (col + 1 < nlim) bv.y = __ldg(p.bias + col + 1);
      if (col + 2 < nlim) bv.z = __ldg(p.bias + col + 2);
      if (col + 3 < nlim) bv.w = __ldg(p.bias + col + 3);
    }
    // fast path (whole chunk uniform): vector rows, residual already folded in (or absent), full M tile
    const bool fast = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && (p.N % 4 == 0) &&
                      (res_vec || !p.residual) && (m0 + GM <= p.M) && seg_len == 0;
    if (fast) {
      if (col < nlim) {
        float* cbase = reinterpret_cast<float*>(p.C) + (m0 + ew) * p.ldc + col;
        const float* sbase = cstage + ew * CPAD + lane * 4;
#pragma unroll 1
        for (int rb = 0; rb < RPW; rb += 8) {
          float4 o[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) o[u] = *reinterpret_cast<const float4*>(sbase + (rb + u) * EPI_WARPS * CPAD);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            o[u].x += bv.x; o[u].y += bv.y; o[u].z += bv.z; o[u].w += bv.w;
            *reinterpret_cast<float4*>(cbase + (int64_t)(rb + u) * EPI_WARPS * p.ldc) = o[u];
          }
        }
      }
    } else
    // general path, 8 rows per batch: all residual loads of a batch are issued before the first store (C may
    // alias the residual -- in-place x = f(x) + x -- so the compiler cannot reorder them itself)
#pragma unroll 1
    for (int rb = 0; rb < RPW; rb += 8) {
      int64_t off[8];
      float4 rv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = (rb + u) * EPI_WARPS + ew;
        const uint32_t m = (uint32_t)m0 + r;
        uint32_t orow = m;
        if (seg_len > 0) { const uint32_t q = m / seg_len; orow = q * (uint32_t)p.seg_stride + (uint32_t)p.seg_off + (m - q * seg_len); }
        off[u] = (m < (uint32_t)p.M && col < nlim) ? (int64_t)orow * p.ldc + col : -1;
        rv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!res_vec && vec && p.residual && off[u] >= 0) rv[u] = *reinterpret_cast<const float4*>(p.residual + off[u]);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (off[u] < 0) continue;
        const int r = (rb + u) * EPI_WARPS + ew;
        float4 o = *reinterpret_cast<const float4*>(cstage + r * CPAD + lane * 4);
        o.x += bv.x + rv[u].x; o.y += bv.y + rv[u].y; o.z += bv.z + rv[u].z; o.w += bv.w + rv[u].w;
        float* crow = reinterpret_cast<float*>(p.C) + off[u];
        if (vec) {
          *reinterpret_cast<float4*>(crow) = o;
        } else {  // ragged N / unaligned C: scalar tail
          const float ov[4] = {o.x, o.y, o.z, o.w};
          for (int j = 0; j < 4; ++j)
            if (col + j < nlim) crow[j] = ov[j] + ((p.residual && !res_vec) ? p.residual[off[u] + j] : 0.f);
        }
      }
    }
  } else if (EPI == 3) {
    // bf16 attention operands: lane owns columns [4 lane, 4 lane + 4) of the chunk; a head is 64 columns = 16 lanes, so
    // the squared norm of a (row, head) is a 4-step xor-shuffle over the half warp
    const int col = ncol0 + lane * 4;
    const bool norm = ncol0 < p.norm_cols;   // chunks are 128-aligned and norm_cols % 128 == 0 (checked on the host)
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
    if (norm) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p.nscale + ((lane * 4) & 63)));
      sc = make_float4(t.x * p.nmul, t.y * p.nmul, t.z * p.nmul, t.w * p.nmul);
    }
#pragma unroll 4
    for (int rr = 0; rr < RPW; ++rr) {
      const int r = rr * EPI_WARPS + ew;
      const uint32_t m = (uint32_t)m0 + r;
      float4 o = *reinterpret_cast<const float4*>(cstage + r * CPAD + lane * 4);
      if (norm) {
        float ss = o.x * o.x + o.y * o.y + o.z * o.z + o.w * o.w;
#pragma unroll
        for (int d = 8; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
        const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
        o.x = (o.x * inv) * sc.x; o.y = (o.y * inv) * sc.y; o.z = (o.z * inv) * sc.z; o.w = (o.w * inv) * sc.w;
      }
      if (m < (uint32_t)p.M && col < nlim)
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.C) + (int64_t)m * p.ldc + col) =
            make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    }
  } else {
    // EPI 1: bf16 outputs, 128 columns (4 per lane)
    constexpr int CPL = 4;
    const int col = ncol0 + lane * CPL;
    float bv[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) bv[j] = (p.bias && col + j < nlim) ? __ldg(p.bias + col + j) : 0.f;
    const bool vec = (p.ldc % CPL == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && (col + CPL - 1 < nlim);
#pragma unroll 4
    for (int rr = 0; rr < RPW; ++rr) {
      const int r = rr * EPI_WARPS + ew;
      const uint32_t m = (uint32_t)m0 + r;
      if (m >= (uint32_t)p.M) break;
      int64_t orow = m;
      if (seg_len > 0) { const uint32_t q = m / seg_len; orow = q * (uint32_t)p.seg_stride + (uint32_t)p.seg_off + (m - q * seg_len); }
      float o[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) o[j] = cstage[r * CPAD + lane * CPL + j] + bv[j];
      __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(p.C) + orow * p.ldc + col;
      if (vec) {
        *reinterpret_cast<uint2*>(crow) = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
      } else {
#pragma unroll
        for (int j = 0; j < CPL; ++j)
          if (col + j < nlim) crow[j] = __float2bfloat16_rn(o[j]);
      }
    }
  }
  epi_bar_sync();  // staging buffer free for the next chunk
}

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {  // shared::cta -> shared::cluster
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// "Accumulator drained" signal of the pair kernel's epilogue warps to the leader's MMA warp.  Nothing the generic proxy wrote has
// to become visible with it -- the TMEM reads are complete (tcgen05.wait::ld) and fenced (tcgen05.fence::before_thread_sync) --
// so the arrive carries the default CTA scope (the form CUTLASS uses for this barrier).  With .release.cluster the compiler
// emitted an ERRBAR in front of it that waited for the thread's outstanding global stores of the PREVIOUS tile: 10 % of the
// FF1 + GEGLU kernel's stall samples sat on that one instruction (ncu source page, call 21).
__device__ __forceinline__ void mbar_arrive_remote_acc(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {  // acquire at cluster scope
  long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if ((spin & 1023u) == 1023u) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000LL) __trap();
    }
  }
}
__device__ __forceinline__ uint32_t cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void st_cluster_f2(uint32_t cluster_addr, float a, float b) {  // distributed shared memory store
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}

// Epilogue 4, after the accumulator-ready wait: x = acc + x (bulk-stored like epi_tma_finish) and LayerNorm(x) of the
// same registers -> bf16.  Row statistics: every thread sums its 32 columns, the four column parts meet in `rowstat`,
// the n_tiles CTAs of the cluster (same m-tile, adjacent 128-column slices) exchange their 128-column partials through
// distributed shared memory: thread (row r, part 0) of CTA c stores its partial into xstat[buf][c][r] of EVERY CTA of the
// cluster and arrives on that CTA's `bar_stat` (128 * cluster_size arrivals per tile, release / acquire at cluster scope).
// Two exchange buffers: a CTA can be at most one tile ahead of a peer (it needs the peer's arrival to get further).
template <bool GLOBAL, typename Release>
__device__ __forceinline__ void epi_tma_finish_ln(const EpiParams& p, uint8_t* stage, uint32_t stage_s, uint32_t bar_res,
                                                  uint32_t& res_phase, uint32_t tmem_chunk, int m0, int n0, int ew, int lg,
                                                  int part, int lane, float2* rowstat, const float2* xstat, uint32_t xstat_s,
                                                  uint32_t bar_stat,
                                                  uint32_t& stat_phase, uint32_t& stat_buf, uint32_t my_rank, uint32_t csize,
                                                  int iter, int group, Release&& release) {
  const uint32_t trow = tmem_chunk + ((uint32_t)(lg * 32) << 16);
  uint32_t v[32];
  tmem_ld32(trow + part * 32, v);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  release();
  const int r = lg * 32 + lane;
  float4* box_row = reinterpret_cast<float4*>(stage + part * (GM * 128) + r * 128);
  if (p.residual) { mbar_wait(bar_res, res_phase); res_phase ^= 1; }
  const int col0 = n0 + part * 32;
  float4 o[8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                       __uint_as_float(v[4 * j + 3]));
    const int pos = j ^ (r & 7);
    if (p.residual) { const float4 rr = box_row[pos]; o[j].x += rr.x; o[j].y += rr.y; o[j].z += rr.z; o[j].w += rr.w; }
    if (p.bias) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
      o[j].x += bv.x; o[j].y += bv.y; o[j].z += bv.z; o[j].w += bv.w;
    }
    box_row[pos] = o[j];
    s1 += (o[j].x + o[j].y) + (o[j].z + o[j].w);
    s2 += (o[j].x * o[j].x + o[j].y * o[j].y) + (o[j].z * o[j].z + o[j].w * o[j].w);
  }
  rowstat[part * GM + r] = make_float2(s1, s2);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA unit
  epi_bar_sync();  // whole chunk staged, the four parts of every row are in rowstat
  if (ew == 0 && lane == 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (n0 + 32 * c < p.N) tma_store_2d(&p.tmC, stage_s + c * (GM * 128), n0 + 32 * c, m0);
    tma_store_commit();
  }
  float t1 = 0.f, t2 = 0.f;
  if (GLOBAL) {
    // global-memory exchange: slot [buf][block][row]; one arrival per CTA and tile on the group's counter
    float2* mine = p.xstat_g + ((size_t)stat_buf * gridDim.x + blockIdx.x) * GM;
    if (part == 0) {
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int q = 0; q < EPI_PARTS; ++q) { const float2 e = rowstat[q * GM + r]; a1 += e.x; a2 += e.y; }
      __stcg(mine + r, make_float2(a1, a2));
      __threadfence();
    }
    epi_bar_sync();  // the 128 partials of this CTA are written and fenced
    if (ew == 0 && lane == 0) {
      unsigned int* ctr = p.ln_counter + group;
      const unsigned int target = (unsigned int)(iter + 1) * csize;
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
      long long t0 = 0;
      for (uint32_t spin = 0;; ++spin) {
        unsigned int seen;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
        if (seen >= target) break;
        if ((spin & 255u) == 255u) {
          if (t0 == 0) t0 = clock64();
          else if (clock64() - t0 > 4000000000LL) __trap();  // a peer CTA never arrived
        }
      }
    }
    epi_bar_sync();  // every peer's partials are visible
    const float2* grp = p.xstat_g + ((size_t)stat_buf * gridDim.x + (size_t)group * csize) * GM;
    for (uint32_t c = 0; c < csize; ++c) { const float2 e = __ldcg(grp + (size_t)c * GM + r); t1 += e.x; t2 += e.y; }
    stat_buf ^= 1;
  } else {
    if (part == 0) {  // 128 threads, one per row: this CTA's 128-column partial to every CTA of the cluster
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int q = 0; q < EPI_PARTS; ++q) { const float2 e = rowstat[q * GM + r]; a1 += e.x; a2 += e.y; }
      const uint32_t slot = xstat_s + ((stat_buf * LN_MAX_CLUSTER + my_rank) * GM + r) * 8;
      for (uint32_t dst = 0; dst < csize; ++dst) {
        st_cluster_f2(map_to_cta(slot, dst), a1, a2);
        mbar_arrive_remote(map_to_cta(bar_stat, dst));
      }
    }
    mbar_wait_cluster(bar_stat, stat_phase);
    stat_phase ^= 1;
    for (uint32_t c = 0; c < csize; ++c) { const float2 e = xstat[(stat_buf * LN_MAX_CLUSTER + c) * GM + r]; t1 += e.x; t2 += e.y; }
    stat_buf ^= 1;
  }
  const float inv_n = 1.0f / (float)p.N;
  const float mean = t1 * inv_n;
  const float rstd = rsqrtf(fmaxf(t2 * inv_n - mean * mean, 0.f) + p.ln_eps);
  if ((uint32_t)(m0 + r) < (uint32_t)p.M) {
    __nv_bfloat16* lrow = reinterpret_cast<__nv_bfloat16*>(p.ln_out) + (int64_t)(m0 + r) * p.ln_ld + col0;
    __nv_bfloat16* rrow = p.raw_out ? reinterpret_cast<__nv_bfloat16*>(p.raw_out) + (int64_t)(m0 + r) * p.ln_ld + col0 : nullptr;
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ln_g + col0) + j), g1 = __ldg(reinterpret_cast<const float4*>(p.ln_g + col0) + j + 1);
      float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
      if (p.ln_b) { b0 = __ldg(reinterpret_cast<const float4*>(p.ln_b + col0) + j); b1 = __ldg(reinterpret_cast<const float4*>(p.ln_b + col0) + j + 1); }
      const float4 a = o[j], c = o[j + 1];
      *reinterpret_cast<uint4*>(lrow + 4 * j) = make_uint4(
          pack_bf16x2(fmaf((a.x - mean) * rstd, g0.x, b0.x), fmaf((a.y - mean) * rstd, g0.y, b0.y)),
          pack_bf16x2(fmaf((a.z - mean) * rstd, g0.z, b0.z), fmaf((a.w - mean) * rstd, g0.w, b0.w)),
          pack_bf16x2(fmaf((c.x - mean) * rstd, g1.x, b1.x), fmaf((c.y - mean) * rstd, g1.y, b1.y)),
          pack_bf16x2(fmaf((c.z - mean) * rstd, g1.z, b1.z), fmaf((c.w - mean) * rstd, g1.w, b1.w)));
      if (rrow)
        *reinterpret_cast<uint4*>(rrow + 4 * j) = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w),
                                                             pack_bf16x2(c.x, c.y), pack_bf16x2(c.z, c.w));
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// One-CTA kernel: 128 x 128 tiles, tcgen05.mma.cta_group::1 (small problems: fewer than 256 rows)
// ---------------------------------------------------------------------------------------------------
// DUAL: two independent problems (own operands, output, N and K) in ONE launch -- the q and k,v projections of a
// self-attention block, which read different inputs (LayerNorm(x) vs raw x, attention.py:140-144) and are each a
// single wave of tiles: together they make ~3 tiles per CTA, so prologue, epilogue and main loop overlap across tiles
// instead of being paid twice.
template <int EPI, bool DUAL>
__global__ void __launch_bounds__(GTHREADS, 1) gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const __grid_constant__ EpiParams p,
                                                                const __grid_constant__ CUtensorMap tmA2,
                                                                const __grid_constant__ CUtensorMap tmB2,
                                                                const __grid_constant__ EpiParams p2) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-B alignment
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base, sB = base + GSTAGES * STAGE_BYTES;
  float* cstage = reinterpret_cast<float*>(base_ptr + RING_BYTES);
  const uint32_t bars = base + RING_BYTES + CSTAGE_BYTES;
  // full[s] @ +8s ; empty[s] @ +8(S+s) ; tmem_full[a] @ +16S+8a ; tmem_empty[a] @ +16S+16+8a ; tmem slot @ +16S+32
  const uint32_t bar_tfull = bars + 16 * GSTAGES, bar_tempty = bar_tfull + 16, tmem_slot = bar_tfull + 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles1 = p.m_tiles * p.n_tiles;
  const int num_tiles = tiles1 + (DUAL ? p2.m_tiles * p2.n_tiles : 0);
  // epilogue 4: clusters of n_tiles CTAs own one m-tile at a time (CTA rank = n-tile), round-robin over the m-tiles
  constexpr bool LN_EPI = EPI == 4 || EPI == 5;
  const uint32_t csize = EPI == 4 ? cluster_size() : EPI == 5 ? (uint32_t)p.n_tiles : 1u;
  const uint32_t crank = EPI == 4 ? cluster_rank() : EPI == 5 ? blockIdx.x % csize : 0u;
  const int cluster_id = (int)blockIdx.x / (int)csize, n_clusters = (int)gridDim.x / (int)csize;
  // tile schedule: round-robin over all tiles (problem 1 first), m-fastest; returns true for a tile of problem 2
  const int my_tiles = LN_EPI ? (cluster_id < p.m_tiles ? (p.m_tiles - 1 - cluster_id) / n_clusters + 1 : 0)
                                : ((int)blockIdx.x < num_tiles ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0);
  auto tile_of = [&](int i, int& m0, int& n0) -> bool {
    if (LN_EPI) { m0 = (cluster_id + i * n_clusters) * GM; n0 = (int)crank * GN; return false; }
    int tile = (int)blockIdx.x + i * (int)gridDim.x;
    const bool second = DUAL && tile >= tiles1;
    if (second) tile -= tiles1;
    const int mt = second ? p2.m_tiles : p.m_tiles;
    m0 = (tile % mt) * GM;
    n0 = (tile / mt) * GN;
    return second;
  };
  auto kblocks = [&](bool second) { return ((second ? p2.K : p.K) + GK - 1) / GK; };
  const long long t_start = clock64();
#define PHK_STAMP(slot) do { if (p.trace) p.trace[blockIdx.x * 16 + (slot)] = clock64() - t_start; } while (0)

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (DUAL) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB2) : "memory");
    }
    for (int s = 0; s < GSTAGES; ++s) {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 8 * (GSTAGES + s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, EPI_WARPS);  // one arrival per epilogue warp
    }
    mbar_init(tmem_slot + 8, 1);  // residual tile landed (TMA epilogue)
    if (EPI == 4) mbar_init(tmem_slot + 16, GM * csize);  // row statistics of the whole cluster landed (epilogue 4)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: two 128-column fp32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(2 * GN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (EPI == 4) { __syncwarp(); cluster_sync_all(); }  // every CTA of the cluster runs and has its barriers initialised
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  int pre_kb = 0;  // k-blocks of the first tile whose W half is already in flight (producer thread only)
  if (threadIdx.x == 0 && my_tiles > 0 && p.w_static) {
    int m0, n0;
    const bool second = tile_of(0, m0, n0);
    const int num_kb = kblocks(second);
    pre_kb = num_kb < GSTAGES ? num_kb : GSTAGES;
    for (int kb = 0; kb < pre_kb; ++kb) {
      mbar_expect_tx(bars + 8 * kb, 2 * STAGE_BYTES);  // both halves; the A half is requested after the wait
      tma_load_2d(second ? &tmB2 : &tmB, bars + 8 * kb, sB + kb * STAGE_BYTES, kb * GK, n0);
    }
  }
  pdl_wait();  // everything above (barriers, tensor-map prefetch, TMEM allocation, W prefetch) overlapped the previous kernel
  if (threadIdx.x == 0) PHK_STAMP(0);  // setup done

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        int m0, n0;
        const bool second = tile_of(it, m0, n0);
        const int num_kb = kblocks(second);
        const CUtensorMap* ma = second ? &tmA2 : &tmA;
        const CUtensorMap* mb = second ? &tmB2 : &tmB;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bars + 8 * (GSTAGES + stage), phase ^ 1);  // slot free (passes immediately on the first lap)
          const uint32_t full = bars + 8 * stage;
          if (it == 0 && kb < pre_kb) {  // W half requested before the wait (stage == kb on the first lap)
            tma_load_2d(ma, full, sA + stage * STAGE_BYTES, kb * GK, m0);
          } else {
            mbar_expect_tx(full, 2 * STAGE_BYTES);
            tma_load_2d(ma, full, sA + stage * STAGE_BYTES, kb * GK, m0);
            tma_load_2d(mb, full, sB + stage * STAGE_BYTES, kb * GK, n0);
          }
          if (it == 0 && kb == 0) PHK_STAMP(1);            // first TMA issued
          if (it == 0 && kb == num_kb - 1) PHK_STAMP(2);   // last TMA of the first tile issued
          if (++stage == GSTAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // cute::UMMA::InstrDescriptor: c=F32 (bit 4), a=b=BF16 (bits 7,10), K-major both, N>>3 @17, M>>4 @24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(GN >> 3) << 17) |
                             ((uint32_t)(GM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int acc = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(bar_tempty + 8 * acc, (use & 1) ^ 1);  // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + acc * GN;
        int m0u, n0u;
        const int num_kb = kblocks(tile_of(it, m0u, n0u));
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bars + 8 * stage, phase);
          if (it == 0 && kb == 0) PHK_STAMP(3);            // first operands landed
          if (it == 0 && kb == num_kb - 1) PHK_STAMP(4);   // last operands landed
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = umma_desc(sA + stage * STAGE_BYTES);
          const uint64_t db = umma_desc(sB + stage * STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < GK / 16; ++k)  // +32 B per UMMA_K inside the 128-B swizzle row => +2 in the address field
            umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          umma_commit(bars + 8 * (GSTAGES + stage));  // frees the smem slot once these MMAs have read it
          if (++stage == GSTAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(bar_tfull + 8 * acc);  // accumulator complete
        if (it == 0) PHK_STAMP(5);         // all MMAs of the first tile issued
      }
    }
  } else {
    // ===================== epilogue: warps 2..17, four per TMEM lane group =====================
    const int ew = warp - 2;        // 0..15: rows ew, ew+16, ... in the coalesced write-out
    const int lg = warp & 3;        // TMEM lane group this warp may access (rows lg*32 .. +31 of the tile)
    const int part = ew >> 2;       // which quarter of the accumulator columns this warp drains
    const uint32_t bar_res = tmem_slot + 8, bar_stat = tmem_slot + 16;
    uint32_t res_phase = 0, stat_phase = 0, stat_buf = 0;
    float2* rowstat = reinterpret_cast<float2*>(base_ptr + RING_BYTES + CSTAGE_BYTES + 256);
    const float2* xstat = rowstat + EPI_PARTS * GM;
    const uint32_t xstat_s = base + RING_BYTES + CSTAGE_BYTES + 256 + LN_ROWSTAT_BYTES;
    for (int it = 0; it < my_tiles; ++it) {
      const int acc = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      int m0, n0;
      const EpiParams& pp = tile_of(it, m0, n0) ? p2 : p;
      const bool tma = (EPI == 0 && pp.tma_epi) || LN_EPI;
      bool res_vec = false;
      if (tma) epi_tma_begin(pp, base + RING_BYTES, bar_res, m0, n0, ew, lane);
      else res_vec = epi_residual_prefetch<EPI>(pp, cstage, m0, n0, ew, lane);
      mbar_wait(bar_tfull + 8 * acc, use & 1);
      if (it == 0 && threadIdx.x == 64) PHK_STAMP(6);      // accumulator ready
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (LN_EPI)
        epi_tma_finish_ln<EPI == 5>(pp, reinterpret_cast<uint8_t*>(cstage), base + RING_BYTES, bar_res, res_phase,
                          tmem_base + acc * GN, m0, n0, ew, lg, part, lane, rowstat, xstat, xstat_s, bar_stat, stat_phase,
                          stat_buf, crank, csize, it, cluster_id, [&]() { if (lane == 0) mbar_arrive(bar_tempty + 8 * acc); });
      else if (tma)
        epi_tma_finish(pp, reinterpret_cast<uint8_t*>(cstage), base + RING_BYTES, bar_res, res_phase,
                       tmem_base + acc * GN, m0, n0, ew, lg, part, lane,
                       [&]() { if (lane == 0) mbar_arrive(bar_tempty + 8 * acc); });
      else if (EPI == 2)
        epi_geglu_tile<1>(pp, cstage, tmem_base + acc * GN, m0, n0, ew, lg, part, lane,
                          [&]() { if (lane == 0) mbar_arrive(bar_tempty + 8 * acc); },
                          [&](int slot) { if (it == 0 && threadIdx.x == 64) PHK_STAMP(slot); });
      else
        epi_chunk<EPI>(pp, cstage, tmem_base + acc * GN, m0, n0, res_vec, ew, lg, part, lane,
                       [&]() { if (lane == 0) mbar_arrive(bar_tempty + 8 * acc); });
      if (it == 0 && threadIdx.x == 64) PHK_STAMP(8);      // tile written out
    }
    if ((EPI == 0 || LN_EPI) && ew == 0 && lane == 0) tma_store_wait_read();  // staging read out before the CTA exits (the writes complete with the grid)
  }
  __syncthreads();
  if (threadIdx.x == 0) PHK_STAMP(9);  // CTA done
  if (EPI == 4) { __syncwarp(); cluster_sync_all(); }  // no CTA exits while a peer may still write its exchange buffer
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * GN) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// CTA-pair kernel: clusters of two CTAs (one TPC) compute 256 x BN tiles with tcgen05.mma.cta_group::2.
// Each CTA stages its own 128 rows of A and HALF of the W tile (BN/2 rows); the tensor cores read the other half
// from the peer's shared memory, so the L2 -> SM operand traffic per FLOP is half that of a 128 x 128 tile at
// BN = 256 (the 128 x 128 kernel ran at the ~6.3 kB/clk L2 cap on the large GEMMs).
//   * both CTAs: warp 0 = TMA producer (loads signal the LEADER's full barrier, .cta_group::2),
//                warps 2..17 = epilogue for the CTA's own 128 accumulator rows (its own TMEM);
//   * leader (cluster rank 0) warp 1 = the only MMA issuer; tcgen05.commit multicasts the "slot free" and
//     "accumulator ready" arrivals to the barriers at the same offset in both CTAs;
//   * the peer's epilogue warps release an accumulator with a remote arrive on the leader's barrier.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t leader_bar, uint32_t dst, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {  // arrives on `bar` in BOTH CTAs of the pair
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}

__host__ __device__ constexpr int pair_stages(int epi) { return epi == 2 ? 6 : 4; }
__host__ __device__ constexpr int pair_cstage_bytes(int epi) { return epi == 2 ? GM * 256 /* bf16 tile */ : CSTAGE_BYTES; }
__host__ __device__ constexpr int pair_smem_bytes(int epi) {
  return pair_stages(epi) * 2 * STAGE_BYTES + pair_cstage_bytes(epi) + 256 + 1024;
}

template <int EPI, int BN, bool DUAL>
__global__ void __launch_bounds__(GTHREADS, 1) gemm_bf16_pair_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                     const __grid_constant__ CUtensorMap tmB,
                                                                     const __grid_constant__ EpiParams p,
                                                                     const __grid_constant__ CUtensorMap tmA2,
                                                                     const __grid_constant__ CUtensorMap tmB2,
                                                                     const __grid_constant__ EpiParams p2) {
  static_assert(BN == 128 || BN == 256, "pair tile is 256 x 128 or 256 x 256");
  constexpr int B_BYTES = (BN / 2) * GK * 2;       // this CTA's half of the W tile per stage
  constexpr int TMEM_COLS = 2 * BN;                // two accumulators of BN fp32 columns
  // the remote signalling round trip (commit -> peer's empty barrier -> peer's TMA -> leader's full barrier) makes a
  // deeper ring worthwhile; the GEGLU epilogue stores from registers and needs no staging buffer
  constexpr int NS = pair_stages(EPI);
  constexpr int RING = NS * 2 * STAGE_BYTES;
  constexpr int CST = pair_cstage_bytes(EPI);
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base, sB = base + NS * STAGE_BYTES;
  float* cstage = reinterpret_cast<float*>(base_ptr + RING);
  const uint32_t bars = base + RING + CST;
  const uint32_t bar_tfull = bars + 16 * NS, bar_tempty = bar_tfull + 16, tmem_slot = bar_tfull + 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  const int tiles1 = p.m_tiles * p.n_tiles;         // pair tiles: m_tiles = ceil(M / 256), n_tiles = ceil(N / BN)
  const int num_tiles = tiles1 + (DUAL ? p2.m_tiles * p2.n_tiles : 0);
  const int pair = (int)blockIdx.x >> 1, num_pairs = (int)gridDim.x >> 1;
  const int my_tiles = pair < num_tiles ? (num_tiles - 1 - pair) / num_pairs + 1 : 0;
  // m0: this CTA's 128 rows; n0: start of the BN-wide tile; returns true for a tile of the second problem (DUAL)
  auto tile_of = [&](int i, int& m0, int& n0) -> bool {
    int tile = pair + i * num_pairs;
    const bool second = DUAL && tile >= tiles1;
    if (second) tile -= tiles1;
    const int mt = second ? p2.m_tiles : p.m_tiles;
    m0 = (tile % mt) * (2 * GM) + (int)rank * GM;
    n0 = (tile / mt) * BN;
    return second;
  };
  auto kblocks = [&](bool second) { return ((second ? p2.K : p.K) + GK - 1) / GK; };
  const long long t_start = clock64();

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (DUAL) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB2) : "memory");
    }
    for (int s = 0; s < NS; ++s) {
      mbar_init(bars + 8 * s, 1);                   // full: the leader's expect_tx arrival (used in the leader only)
      mbar_init(bars + 8 * (NS + s), 1);       // empty: one multicast commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);              // one multicast commit
      mbar_init(bar_tempty + 8 * a, 2 * EPI_WARPS); // every epilogue warp of BOTH CTAs (used in the leader only)
    }
    mbar_init(tmem_slot + 8, 1);                    // residual tile landed (TMA epilogue)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // the same warp of both CTAs allocates the pair's TMEM columns
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  __syncwarp();
  cluster_sync_all();  // the peer's barriers are initialised before any remote arrive / complete_tx targets them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();
  if (threadIdx.x == 0) PHK_STAMP(0);  // setup done

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        int m0, n0;
        const bool second = tile_of(it, m0, n0);
        const int num_kb = kblocks(second);
        const CUtensorMap* ma = second ? &tmA2 : &tmA;
        const CUtensorMap* mb = second ? &tmB2 : &tmB;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bars + 8 * (NS + stage), phase ^ 1);  // slot free in THIS CTA (multicast commit)
          if (leader) mbar_expect_tx(bars + 8 * stage, 2 * (STAGE_BYTES + B_BYTES));  // bytes of both CTAs
          const uint32_t full = map_to_cta(bars + 8 * stage, 0);                     // the leader's full barrier
          tma_load_2d_pair(ma, full, sA + stage * STAGE_BYTES, kb * GK, m0);
          tma_load_2d_pair(mb, full, sB + stage * STAGE_BYTES, kb * GK, n0 + (int)rank * (BN / 2));
          if (it == 0 && kb == 0) PHK_STAMP(1);
          if (it == 0 && kb == num_kb - 1) PHK_STAMP(2);
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
      }
      // tail: every multicast "slot free" arrival aimed at this CTA has landed before it may exit
      for (int s = 0; s < NS; ++s) {
        mbar_wait(bars + 8 * (NS + stage), phase ^ 1);
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader only) =====================
    if (leader && lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)((2 * GM) >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int acc = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait_cluster(bar_tempty + 8 * acc, (use & 1) ^ 1);  // both CTAs have drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + acc * BN;
        int m0u, n0u;
        const int num_kb = kblocks(tile_of(it, m0u, n0u));
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_cluster(bars + 8 * stage, phase);            // both CTAs' operands have landed
          if (it == 0 && kb == 0) PHK_STAMP(3);
          if (it == 0 && kb == num_kb - 1) PHK_STAMP(4);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = umma_desc(sA + stage * STAGE_BYTES);
          const uint64_t db = umma_desc(sB + stage * STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < GK / 16; ++k)
            umma_f16_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          umma_commit_pair(bars + 8 * (NS + stage));        // slot free in both CTAs
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(bar_tfull + 8 * acc);                   // accumulator complete, both CTAs
        if (it == 0) PHK_STAMP(5);
      }
    }
  } else {
    // ===================== epilogue (both CTAs): own 128 rows x BN columns =====================
    const int ew = warp - 2, lg = warp & 3, part = ew >> 2;
    const uint32_t tempty_leader = map_to_cta(bar_tempty, 0);
    const uint32_t bar_res = tmem_slot + 8, stage_s = base + RING;
    uint32_t res_phase = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const int acc = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      int m0, n0;
      const EpiParams& pp = tile_of(it, m0, n0) ? p2 : p;
      const bool tma = EPI == 0 && pp.tma_epi;
      bool res_vec = false;
      if (tma) epi_tma_begin(pp, stage_s, bar_res, m0, n0, ew, lane);
      else res_vec = epi_residual_prefetch<EPI>(pp, cstage, m0, n0, ew, lane);
      mbar_wait(bar_tfull + 8 * acc, use & 1);
      if (it == 0 && threadIdx.x == 64) PHK_STAMP(6);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (tma) {
#pragma unroll
        for (int h = 0; h < BN / 128; ++h) {
          if (h > 0) epi_tma_begin(pp, stage_s, bar_res, m0, n0 + h * 128, ew, lane);
          const bool last = h == BN / 128 - 1;
          epi_tma_finish(pp, reinterpret_cast<uint8_t*>(cstage), stage_s, bar_res, res_phase,
                         tmem_base + acc * BN + h * 128, m0, n0 + h * 128, ew, lg, part, lane,
                         [&]() { if (last && lane == 0) mbar_arrive_remote_acc(tempty_leader + 8 * acc); });
        }
      } else if (EPI == 2) {
        epi_geglu_tile<BN / 128>(pp, cstage, tmem_base + acc * BN, m0, n0, ew, lg, part, lane,
                                 [&]() { if (lane == 0) mbar_arrive_remote_acc(tempty_leader + 8 * acc); },
                                 [&](int slot) { if (it == 0 && threadIdx.x == 64) PHK_STAMP(slot); });
      } else {
#pragma unroll
        for (int h = 0; h < BN / 128; ++h) {
          if (h > 0) res_vec = epi_residual_prefetch<EPI>(pp, cstage, m0, n0 + h * 128, ew, lane);
          const bool last = h == BN / 128 - 1;
          epi_chunk<EPI>(pp, cstage, tmem_base + acc * BN + h * 128, m0, n0 + h * 128, res_vec, ew, lg, part, lane,
                         [&]() { if (last && lane == 0) mbar_arrive_remote_acc(tempty_leader + 8 * acc); });
        }
      }
      if (it == 0 && threadIdx.x == 64) PHK_STAMP(8);
    }
    if (EPI == 0 && ew == 0 && lane == 0) tma_store_wait_read();  // staging read out before the CTA exits (the writes complete with the grid)
  }
  __syncthreads();
  if (threadIdx.x == 0) PHK_STAMP(9);
  __syncwarp();
  cluster_sync_all();  // no CTA of the pair exits (or frees TMEM) while the other may still signal it
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// host side: tensor maps (driver entry point resolved at run time: libphk.so has no link-time libcuda dependency,
// so it also loads on a CPU-only box for the ABI tests)
// ---------------------------------------------------------------------------------------------------

struct MapKey {
  const void* ptr; int64_t rows, cols, ld; int box_rows;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    h = h * 1000003u ^ std::hash<int64_t>()(k.rows);
    h = h * 1000003u ^ std::hash<int64_t>()(k.cols);
    h = h * 1000003u ^ std::hash<int64_t>()(k.ld * 131 + k.box_rows);
    return h;
  }
};

// [rows, cols] bf16, row pitch ld elements; box = [box_rows, 64 cols], SWIZZLE_128B, zero OOB fill
static int get_tensor_map(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, CUtensorMap* out) {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static std::mutex mu;
  const MapKey key{ptr, rows, cols, ld, box_rows};
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return 0; }
  EncodeTiledFn fn = encode_fn();
  PHK_REQUIRE(fn, PHK_E_UNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)GK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  const CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PHK_REQUIRE(r == CUDA_SUCCESS, PHK_E_ARG, "cuTensorMapEncodeTiled rejected the operand (alignment / pitch)");
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return 0;
}

// fp32 [rows, cols] output, row pitch ld floats; box = [128 rows, 32 floats = 128 B], SWIZZLE_128B (TMA epilogue)
static int get_c_map(const void* ptr, int64_t rows, int64_t cols, int64_t ld, CUtensorMap* out) {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static std::mutex mu;
  const MapKey key{ptr, rows, cols, ld, -32};
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return 0; }
  EncodeTiledFn fn = encode_fn();
  PHK_REQUIRE(fn, PHK_E_UNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)GM};
  const cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  const CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PHK_REQUIRE(r == CUDA_SUCCESS, PHK_E_ARG, "cuTensorMapEncodeTiled rejected the output (alignment / pitch)");
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return 0;
}

// the TMA epilogue applies to plain fp32 outputs: identity row map, 16-byte aligned rows, residual == C (or none)
static int maybe_tma_epilogue(EpiParams& p, int epilogue) {
  static int on = -1;
  if (on < 0) { const char* e = std::getenv("PHK_GEMM_TMA_EPI"); on = (e && e[0] == '0') ? 0 : 1; }
  p.tma_epi = 0;
  if (!on || epilogue != 0 || p.seg_len != 0 || p.ldc % 4 != 0 || (reinterpret_cast<uintptr_t>(p.C) & 15) != 0 ||
      (p.residual && p.residual != p.C))
    return 0;
  PHK_TRY(get_c_map(p.C, p.M, p.N, p.ldc, &p.tmC));
  p.tma_epi = 1;
  return 0;
}

static long long* g_gemm_trace = nullptr;

// set by the inference drivers (api.cu) around their GEMM calls: the W operands are weights nothing in the stream rewrites
static thread_local int t_static_weights = 0;
void gemm_static_weights(bool on) { t_static_weights = on ? 1 : 0; }

template <int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const EpiParams& p_in, cudaStream_t st) {
  EpiParams p = p_in;
  p.w_static = t_static_weights;
  static unsigned long long configured_mask = 0;
  const bool configured = device_configured(&configured_mask);
  if (!configured) {
    PHK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<EPI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    mark_configured(&configured_mask);
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
  PHK_CUDA(launch_pdl(gemm_bf16_kernel<EPI, false>, dim3(grid), dim3(GTHREADS), (size_t)(SMEM_TOTAL), st, ta, tb, p, ta, tb, p));
  PHK_LAUNCH_CHECK();
  return 0;
}

template <int EPI>
static int launch_gemm_dual(const CUtensorMap& ta, const CUtensorMap& tb, const EpiParams& p_in, const CUtensorMap& ta2,
                            const CUtensorMap& tb2, const EpiParams& p2, cudaStream_t st) {
  EpiParams p = p_in;
  p.w_static = t_static_weights;  // (read from the first problem's parameters for both)
  static unsigned long long configured_mask = 0;
  const bool configured = device_configured(&configured_mask);
  if (!configured) {
    PHK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<EPI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    mark_configured(&configured_mask);
  }
  const int tiles = p.m_tiles * p.n_tiles + p2.m_tiles * p2.n_tiles;
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
  PHK_CUDA(launch_pdl(gemm_bf16_kernel<EPI, true>, dim3(grid), dim3(GTHREADS), (size_t)(SMEM_TOTAL), st, ta, tb, p, ta2, tb2, p2));
  PHK_LAUNCH_CHECK();
  return 0;
}

// clusters of two CTAs (+ programmatic dependent launch); one pair per TPC, persistent over the pair tiles
template <int EPI, int BN, bool DUAL>
static int launch_gemm_pair(const CUtensorMap& ta, const CUtensorMap& tb, const EpiParams& p, const CUtensorMap& ta2,
                            const CUtensorMap& tb2, const EpiParams& p2, cudaStream_t st) {
  static unsigned long long configured_mask = 0;
  const bool configured = device_configured(&configured_mask);
  if (!configured) {
    PHK_CUDA(cudaFuncSetAttribute(gemm_bf16_pair_kernel<EPI, BN, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, pair_smem_bytes(EPI)));
    mark_configured(&configured_mask);
  }
  const int tiles = p.m_tiles * p.n_tiles + (DUAL ? p2.m_tiles * p2.n_tiles : 0), max_pairs = kNumSMs / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(GTHREADS); cfg.dynamicSmemBytes = pair_smem_bytes(EPI); cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 2;
  PHK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_pair_kernel<EPI, BN, DUAL>, ta, tb, p, ta2, tb2, p2));
  PHK_LAUNCH_CHECK();
  return 0;
}

template <int BN>
static int launch_gemm_pair_epi(int epilogue, const CUtensorMap& ta, const CUtensorMap& tb, const EpiParams& p,
                                cudaStream_t st) {
  if (epilogue == 2) return launch_gemm_pair<2, BN, false>(ta, tb, p, ta, tb, p, st);
  if (epilogue == 1) return launch_gemm_pair<1, BN, false>(ta, tb, p, ta, tb, p, st);
  return launch_gemm_pair<0, BN, false>(ta, tb, p, ta, tb, p, st);
}

// 0: automatic; 1: always the one-CTA kernel; 2 / 3: CTA pairs with BN = 128 / 256 whenever M > 128 (A/B measurements)
static int g_gemm_mode_override = -1;
static int gemm_mode() {
  if (g_gemm_mode_override >= 0) return g_gemm_mode_override;
  static int mode = -1;
  if (mode < 0) {
    const char* e = std::getenv("PHK_GEMM_MODE");
    mode = (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 0;
  }
  return mode;
}

}  // namespace phk

using namespace phk;

extern "C" int phk_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int64_t M,
                             int32_t N, int32_t K, const float* bias, const float* residual, int64_t seg_len,
                             int64_t seg_stride, int64_t seg_off, int32_t epilogue, phk_stream_t s) {
  Prof prof_(FAM_GEMM_BF16, s, 2.0 * (double)M * N * K);
  PHK_REQUIRE(A && W && C, PHK_E_ARG, "phk_gemm_bf16: null pointer");
  PHK_REQUIRE(M >= 0 && N > 0 && K > 0 && lda >= K && ldw >= K, PHK_E_ARG, "phk_gemm_bf16: bad size");
  PHK_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(W) & 15) == 0,
              PHK_E_ARG, "phk_gemm_bf16: operands must be 16-byte aligned with leading dimensions multiple of 8 (TMA)");
  PHK_REQUIRE(epilogue >= 0 && epilogue <= 2, PHK_E_ARG, "phk_gemm_bf16: unknown epilogue");
  PHK_REQUIRE(epilogue != 2 || (N % 128 == 0 && !bias && !residual && ldc % 2 == 0 &&
                                (reinterpret_cast<uintptr_t>(C) & 3) == 0),
              PHK_E_ARG, "phk_gemm_bf16: GEGLU epilogue needs N % 128 == 0, an even ldc and no bias/residual");
  PHK_REQUIRE(M < (1LL << 31) - GM, PHK_E_UNSUPPORTED, "phk_gemm_bf16: M too large");
  if (M == 0) return 0;
  CUtensorMap ta, tb;
  PHK_TRY(get_tensor_map(A, M, K, lda, GM, &ta));
  cudaStream_t st = to_stream(s);
  const int mode = gemm_mode();
  // Kernel choice (measured, profiles/r01_gemm_modes.txt).  CTA pairs with 256 x 256 tiles halve the L2 -> SM operand
  // bytes per FLOP and win whenever there are enough tiles for more than one wave of the 74 pairs (FF1, to_pixels,
  // logits head: 1.4x).  One-wave problems (N = 512 / 1024 at K = 512..1408) are latency chains of a single tile per
  // CTA: there the one-CTA kernel's cheaper prologue (no cluster sync, no paired TMEM allocation) wins.  Long-K
  // single-wave problems (patch embedding, K = 6144) take 256 x 128 pair tiles (25 % fewer operand bytes).
  const int m_pairs = (int)((M + 2 * GM - 1) / (2 * GM));
  const bool many = (int64_t)m_pairs * ((N + 255) / 256) >= 100;
  const bool pair = M > GM && (mode == 2 || mode == 3 || (mode == 0 && (many || K >= 2048)));
  if (pair) {
    const bool wide = mode == 3 || (mode == 0 && many);
    const int bn = wide ? 256 : 128;
    PHK_TRY(get_tensor_map(W, N, K, ldw, bn / 2, &tb));
    EpiParams p{C, ldc, M, N, K, bias, residual, seg_len, seg_stride, seg_off, m_pairs, (N + bn - 1) / bn, g_gemm_trace};
    PHK_REQUIRE((int64_t)p.m_tiles * p.n_tiles < (1LL << 31), PHK_E_UNSUPPORTED, "phk_gemm_bf16: too many tiles");
    PHK_TRY(maybe_tma_epilogue(p, epilogue));
    return wide ? launch_gemm_pair_epi<256>(epilogue, ta, tb, p, st) : launch_gemm_pair_epi<128>(epilogue, ta, tb, p, st);
  }
  PHK_TRY(get_tensor_map(W, N, K, ldw, GN, &tb));
  EpiParams p{C, ldc, M, N, K, bias, residual, seg_len, seg_stride, seg_off, (int)((M + GM - 1) / GM), (N + GN - 1) / GN, g_gemm_trace};
  PHK_REQUIRE((int64_t)p.m_tiles * p.n_tiles < (1LL << 31), PHK_E_UNSUPPORTED, "phk_gemm_bf16: too many tiles");
  PHK_TRY(maybe_tma_epilogue(p, epilogue));
  if (epilogue == 2) return launch_gemm<2>(ta, tb, p, st);
  if (epilogue == 1) return launch_gemm<1>(ta, tb, p, st);
  return launch_gemm<0>(ta, tb, p, st);
}

// Two independent products C1 = A1 W1^T (+bias1) [M1,N1] and C2 = A2 W2^T (+bias2) [M2,N2] (fp32 outputs) in one
// launch: the q and k,v projections of a self-attention block (attention.py:140-146), or the first-frame and
// rest-frames patch embeddings (cvivit.py:542-549: 16 + 128 tiles, neither of which fills the machine alone).
extern "C" int phk_gemm_bf16_x2(const void* A1, int64_t lda1, const void* W1, int64_t ldw1, float* C1, int64_t ldc1,
                                int64_t M1, int32_t N1, int32_t K1, const float* bias1, const void* A2, int64_t lda2,
                                const void* W2, int64_t ldw2, float* C2, int64_t ldc2, int64_t M2, int32_t N2,
                                int32_t K2, const float* bias2, phk_stream_t s) {
  Prof prof_(FAM_GEMM_BF16, s, 2.0 * ((double)M1 * N1 * K1 + (double)M2 * N2 * K2));
  PHK_REQUIRE(A1 && W1 && C1 && A2 && W2 && C2, PHK_E_ARG, "phk_gemm_bf16_x2: null pointer");
  PHK_REQUIRE(M1 > 0 && M2 > 0 && N1 > 0 && K1 > 0 && N2 > 0 && K2 > 0 && lda1 >= K1 && ldw1 >= K1 && lda2 >= K2 &&
                  ldw2 >= K2, PHK_E_ARG, "phk_gemm_bf16_x2: bad size");
  PHK_REQUIRE(lda1 % 8 == 0 && ldw1 % 8 == 0 && lda2 % 8 == 0 && ldw2 % 8 == 0 &&
                  ((reinterpret_cast<uintptr_t>(A1) | reinterpret_cast<uintptr_t>(W1) | reinterpret_cast<uintptr_t>(A2) |
                    reinterpret_cast<uintptr_t>(W2)) & 15) == 0,
              PHK_E_ARG, "phk_gemm_bf16_x2: operands must be 16-byte aligned with leading dimensions multiple of 8 (TMA)");
  PHK_REQUIRE(M1 < (1LL << 31) - 2 * GM && M2 < (1LL << 31) - 2 * GM, PHK_E_UNSUPPORTED, "phk_gemm_bf16_x2: M too large");
  CUtensorMap ta, tb, ta2, tb2;
  PHK_TRY(get_tensor_map(A1, M1, K1, lda1, GM, &ta));
  PHK_TRY(get_tensor_map(A2, M2, K2, lda2, GM, &ta2));
  const int mode = gemm_mode();
  // measured (profiles/r01_gemm_modes.txt): with fp32 outputs the 256 x 256 pair tile's two-chunk epilogue costs more
  // than the halved operand traffic saves at K = 512, so the pair variant is only taken when forced (tests, A/B runs)
  const bool pair = M1 > GM && M2 > GM && mode >= 2;
  if (pair) {
    PHK_TRY(get_tensor_map(W1, N1, K1, ldw1, 128, &tb));
    PHK_TRY(get_tensor_map(W2, N2, K2, ldw2, 128, &tb2));
    EpiParams p{C1, ldc1, M1, N1, K1, bias1, nullptr, 0, 0, 0, (int)((M1 + 2 * GM - 1) / (2 * GM)), (N1 + 255) / 256, nullptr};
    EpiParams p2{C2, ldc2, M2, N2, K2, bias2, nullptr, 0, 0, 0, (int)((M2 + 2 * GM - 1) / (2 * GM)), (N2 + 255) / 256, nullptr};
    PHK_TRY(maybe_tma_epilogue(p, 0));
    PHK_TRY(maybe_tma_epilogue(p2, 0));
    return launch_gemm_pair<0, 256, true>(ta, tb, p, ta2, tb2, p2, to_stream(s));
  }
  PHK_TRY(get_tensor_map(W1, N1, K1, ldw1, GN, &tb));
  PHK_TRY(get_tensor_map(W2, N2, K2, ldw2, GN, &tb2));
  EpiParams p{C1, ldc1, M1, N1, K1, bias1, nullptr, 0, 0, 0, (int)((M1 + GM - 1) / GM), (N1 + GN - 1) / GN, nullptr};
  EpiParams p2{C2, ldc2, M2, N2, K2, bias2, nullptr, 0, 0, 0, (int)((M2 + GM - 1) / GM), (N2 + GN - 1) / GN, nullptr};
  PHK_REQUIRE((int64_t)p.m_tiles * p.n_tiles + (int64_t)p2.m_tiles * p2.n_tiles < (1LL << 31), PHK_E_UNSUPPORTED,
              "phk_gemm_bf16_x2: too many tiles");
  // (measured: the bulk-store epilogue does not pay for the two-problem launch -- 16.2 vs 15.5 us -- so it stays off)
  return launch_gemm_dual<0>(ta, tb, p, ta2, tb2, p2, to_stream(s));
}

// x = A W^T + x in place (fp32 residual stream) AND the LayerNorm the next sub-block applies to it (attention.py:311-332:
// x = attn(x) + x is followed by ff's / the cross-attention's LayerNorm, x = ff(x) + x by the next layer's), written as
// the bf16 operand of the next GEMM from the epilogue's registers: ln_out[M, ln_ld] = LayerNorm(x) * ln_g (+ ln_b), and
// optionally raw_out = bf16(x) (the k,v projection reads the un-normalised rows, attention.py:140-144).  N = the model
// width: N / 128 in {1, 2, 4, 8} CTAs form a cluster per 128-row tile and share the row statistics.
static int gemm_bf16_ln_impl(const void* A, int64_t lda, const void* W, int64_t ldw, float* C, int64_t ldc, int64_t M,
                             int32_t N, int32_t K, const float* bias, const float* ln_g, const float* ln_b, float ln_eps,
                             void* ln_out, void* raw_out, int64_t ln_ld, void* stat_ws, uint32_t* counters, phk_stream_t s) {
  Prof prof_(FAM_GEMM_BF16, s, 2.0 * (double)M * N * K);
  PHK_REQUIRE(A && W && C && ln_g && ln_out, PHK_E_ARG, "phk_gemm_bf16_ln: null pointer");
  PHK_REQUIRE(M > 0 && N > 0 && K > 0 && lda >= K && ldw >= K && ldc >= N && ln_ld >= N, PHK_E_ARG, "phk_gemm_bf16_ln: bad size");
  const int nt = N / GN;
  PHK_REQUIRE(N % GN == 0 && (nt == 1 || nt == 2 || nt == 4 || nt == 8), PHK_E_UNSUPPORTED,
              "phk_gemm_bf16_ln: the row width must be 128, 256, 512 or 1024");
  PHK_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && ldc % 4 == 0 && ln_ld % 8 == 0 &&
                  ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(C) |
                    reinterpret_cast<uintptr_t>(ln_g) | reinterpret_cast<uintptr_t>(ln_b) | reinterpret_cast<uintptr_t>(bias) |
                    reinterpret_cast<uintptr_t>(ln_out) | reinterpret_cast<uintptr_t>(raw_out)) & 15) == 0,
              PHK_E_ARG, "phk_gemm_bf16_ln: operands must be 16-byte aligned with leading dimensions multiple of 8 (TMA)");
  PHK_REQUIRE(M < (1LL << 31) - GM, PHK_E_UNSUPPORTED, "phk_gemm_bf16_ln: M too large");
  CUtensorMap ta, tb;
  PHK_TRY(get_tensor_map(A, M, K, lda, GM, &ta));
  PHK_TRY(get_tensor_map(W, N, K, ldw, GN, &tb));
  EpiParams p{C, ldc, M, N, K, bias, C, 0, 0, 0, (int)((M + GM - 1) / GM), nt, nullptr};
  PHK_TRY(get_c_map(C, M, N, ldc, &p.tmC));
  p.tma_epi = 1;
  p.ln_g = ln_g; p.ln_b = ln_b; p.ln_out = ln_out; p.raw_out = raw_out; p.ln_ld = ln_ld; p.ln_eps = ln_eps;
  p.w_static = t_static_weights;
  static unsigned long long configured_mask = 0;
  const bool configured = device_configured(&configured_mask);
  if (!configured) {
    PHK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL_LN));
    PHK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<5, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL_LN));
    mark_configured(&configured_mask);
  }
  const int max_clusters = kNumSMs / nt;
  const int clusters = p.m_tiles < max_clusters ? p.m_tiles : max_clusters;
  if (stat_ws) {  // groups of nt CTAs without a cluster: statistics through global memory (epilogue 5)
    p.xstat_g = reinterpret_cast<float2*>(stat_ws);
    p.ln_counter = counters;
    PHK_CUDA(launch_pdl(gemm_bf16_kernel<5, false>, dim3(clusters * nt), dim3(GTHREADS), (size_t)(SMEM_TOTAL_LN), to_stream(s),
                        ta, tb, p, ta, tb, p));
    PHK_LAUNCH_CHECK();
    return 0;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * nt); cfg.blockDim = dim3(GTHREADS); cfg.dynamicSmemBytes = SMEM_TOTAL_LN; cfg.stream = to_stream(s);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = nt; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 2;
  PHK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<4, false>, ta, tb, p, ta, tb, p));
  PHK_LAUNCH_CHECK();
  return 0;
}

extern "C" int phk_gemm_bf16_ln(const void* A, int64_t lda, const void* W, int64_t ldw, float* C, int64_t ldc, int64_t M,
                                int32_t N, int32_t K, const float* bias, const float* ln_g, const float* ln_b, float ln_eps,
                                void* ln_out, void* raw_out, int64_t ln_ld, phk_stream_t s) {
  return gemm_bf16_ln_impl(A, lda, W, ldw, C, ldc, M, N, K, bias, ln_g, ln_b, ln_eps, ln_out, raw_out, ln_ld, nullptr, nullptr, s);
}

// The same product with the row statistics exchanged through global memory instead of a cluster's distributed shared
// memory: the N / 128 CTAs of a 128-row tile are ordinary CTAs of a grid of at most one CTA per SM (all resident), so
// the GPC-local cluster placement that limits phk_gemm_bf16_ln to ~32 resident row tiles at N = 512 does not apply.
//   stat_ws : PHK_LN_STAT_BYTES bytes of scratch (any content);
//   counters: PHK_LN_COUNTERS zero-initialised 32-bit words, consumed by the call (left non-zero): give every call of a
//             stream-ordered sequence its own words and clear them together once per sequence.
extern "C" int phk_gemm_bf16_ln_ws(const void* A, int64_t lda, const void* W, int64_t ldw, float* C, int64_t ldc, int64_t M,
                                   int32_t N, int32_t K, const float* bias, const float* ln_g, const float* ln_b,
                                   float ln_eps, void* ln_out, void* raw_out, int64_t ln_ld, void* stat_ws,
                                   uint32_t* counters, phk_stream_t s) {
  PHK_REQUIRE(stat_ws && counters && (reinterpret_cast<uintptr_t>(stat_ws) & 15) == 0, PHK_E_ARG,
              "phk_gemm_bf16_ln_ws: statistics scratch / counters missing or misaligned");
  return gemm_bf16_ln_impl(A, lda, W, ldw, C, ldc, M, N, K, bias, ln_g, ln_b, ln_eps, ln_out, raw_out, ln_ld, stat_ws, counters, s);
}

// The q and k,v projections of a self-attention block (attention.py:140-157) in one launch, written as the bf16 operands
// of the attention core: Qn[M, I] = normalize_per_head(xn Wq^T) * q_scale * sim_scale, KVn[M, 2I] = [normalize_per_head(
// xraw Wk^T) * k_scale | xraw Wv^T].  dim_head 64, I % 128 == 0.  Replaces fp32 q / kv round trips + a separate
// normalisation pass.
extern "C" int phk_gemm_bf16_qkv(const void* xn, const void* xraw, int64_t lda, const void* Wq, const void* Wkv, int64_t ldw,
                                 void* Qn, void* KVn, int64_t M, int32_t I, int32_t K, const float* q_scale,
                                 const float* k_scale, float sim_scale, phk_stream_t s) {
  Prof prof_(FAM_GEMM_BF16, s, 2.0 * (double)M * 3.0 * I * K);
  PHK_REQUIRE(xn && xraw && Wq && Wkv && Qn && KVn && q_scale && k_scale, PHK_E_ARG, "phk_gemm_bf16_qkv: null pointer");
  PHK_REQUIRE(M > 0 && I > 0 && I % 128 == 0 && K > 0 && lda >= K && ldw >= K, PHK_E_ARG,
              "phk_gemm_bf16_qkv: heads * 64 must be a multiple of 128");
  PHK_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 &&
                  ((reinterpret_cast<uintptr_t>(xn) | reinterpret_cast<uintptr_t>(xraw) | reinterpret_cast<uintptr_t>(Wq) |
                    reinterpret_cast<uintptr_t>(Wkv) | reinterpret_cast<uintptr_t>(q_scale) |
                    reinterpret_cast<uintptr_t>(k_scale)) & 15) == 0 &&
                  ((reinterpret_cast<uintptr_t>(Qn) | reinterpret_cast<uintptr_t>(KVn)) & 7) == 0,
              PHK_E_ARG, "phk_gemm_bf16_qkv: operands must be 16-byte aligned with leading dimensions multiple of 8 (TMA)");
  PHK_REQUIRE(M < (1LL << 31) - 2 * GM, PHK_E_UNSUPPORTED, "phk_gemm_bf16_qkv: M too large");
  CUtensorMap ta, tb, ta2, tb2;
  PHK_TRY(get_tensor_map(xn, M, K, lda, GM, &ta));
  PHK_TRY(get_tensor_map(xraw, M, K, lda, GM, &ta2));
  PHK_TRY(get_tensor_map(Wq, I, K, ldw, GN, &tb));
  PHK_TRY(get_tensor_map(Wkv, 2 * I, K, ldw, GN, &tb2));
  EpiParams p{Qn, I, M, I, K, nullptr, nullptr, 0, 0, 0, (int)((M + GM - 1) / GM), I / GN, nullptr};
  EpiParams p2{KVn, 2 * (int64_t)I, M, 2 * I, K, nullptr, nullptr, 0, 0, 0, (int)((M + GM - 1) / GM), 2 * I / GN, nullptr};
  p.tma_epi = 0; p.nscale = q_scale; p.norm_cols = I; p.nmul = sim_scale;
  p2.tma_epi = 0; p2.nscale = k_scale; p2.norm_cols = I; p2.nmul = 1.0f;
  PHK_REQUIRE((int64_t)p.m_tiles * p.n_tiles + (int64_t)p2.m_tiles * p2.n_tiles < (1LL << 31), PHK_E_UNSUPPORTED,
              "phk_gemm_bf16_qkv: too many tiles");
  return launch_gemm_dual<3>(ta, tb, p, ta2, tb2, p2, to_stream(s));
}

// The q projection of a cross-attention block (attention.py:139, 153-157) written as the bf16 operand of the attention
// core: Qn[M, I] = normalize_per_head(xn Wq^T) * q_scale * sim_scale (epilogue 3 on one problem).  dim_head 64, I % 128 == 0.
extern "C" int phk_gemm_bf16_qnorm(const void* xn, int64_t lda, const void* Wq, int64_t ldw, void* Qn, int64_t M, int32_t I,
                                   int32_t K, const float* q_scale, float sim_scale, phk_stream_t s) {
  Prof prof_(FAM_GEMM_BF16, s, 2.0 * (double)M * I * K);
  PHK_REQUIRE(xn && Wq && Qn && q_scale, PHK_E_ARG, "phk_gemm_bf16_qnorm: null pointer");
  PHK_REQUIRE(M > 0 && I > 0 && I % 128 == 0 && K > 0 && lda >= K && ldw >= K, PHK_E_ARG,
              "phk_gemm_bf16_qnorm: heads * 64 must be a multiple of 128");
  PHK_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 &&
                  ((reinterpret_cast<uintptr_t>(xn) | reinterpret_cast<uintptr_t>(Wq) | reinterpret_cast<uintptr_t>(q_scale)) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(Qn) & 7) == 0,
              PHK_E_ARG, "phk_gemm_bf16_qnorm: operands must be 16-byte aligned with leading dimensions multiple of 8 (TMA)");
  PHK_REQUIRE(M < (1LL << 31) - 2 * GM, PHK_E_UNSUPPORTED, "phk_gemm_bf16_qnorm: M too large");
  CUtensorMap ta, tb;
  PHK_TRY(get_tensor_map(xn, M, K, lda, GM, &ta));
  PHK_TRY(get_tensor_map(Wq, I, K, ldw, GN, &tb));
  EpiParams p{Qn, I, M, I, K, nullptr, nullptr, 0, 0, 0, (int)((M + GM - 1) / GM), I / GN, nullptr};
  p.tma_epi = 0; p.nscale = q_scale; p.norm_cols = I; p.nmul = sim_scale;
  return launch_gemm<3>(ta, tb, p, to_stream(s));
}

// debug / tests: force the kernel choice (0 automatic, 1 one-CTA, 2 CTA pairs 256x128, 3 CTA pairs 256x256; < 0 returns
// to the PHK_GEMM_MODE environment default)
extern "C" int phk_debug_gemm_mode(int32_t mode) {
  PHK_REQUIRE(mode <= 3, PHK_E_ARG, "phk_debug_gemm_mode: mode must be <= 3");
  g_gemm_mode_override = mode;
  return 0;
}

extern "C" int phk_debug_static_weights(int32_t on) {
  gemm_static_weights(on != 0);
  return 0;
}

// debug: device buffer of 16 x int64 per CTA receiving clock64 stamps (relative to CTA start) of the first tile
extern "C" int phk_debug_gemm_trace(long long* device_buffer) { g_gemm_trace = device_buffer; return 0; }
