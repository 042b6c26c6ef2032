// bf16 GEMM on the 5th-generation tensor cores (sm_100a only):
//   C[map(m), n] = sum_k A[m,k] * W[n,k]  (+bias) (+residual)        nn.Linear semantics
// A, W bf16, K-major (row-major [rows, K]); fp32 accumulation in TMEM.
//
// Persistent, warp-specialised (one CTA per SM, static round-robin tile scheduler):
//   warp 0     TMA producer: cp.async.bulk.tensor.2d (SWIZZLE_128B) stages 128x64 A and 128x64 W tiles into a
//              4-deep shared-memory ring; out-of-bounds rows / the K tail are zero-filled by the TMA unit.
//   warp 1     TMEM allocator + MMA issuer: one thread issues tcgen05.mma.cta_group::1.kind::f16
//              (M=128, N=128, K=16), four per k-block; tcgen05.commit frees the smem slot and publishes the
//              accumulator.  Two 128-column accumulators in TMEM, so tile i+1 is multiplied while tile i drains.
//   warps 2..9 epilogue (two per TMEM lane group): residual-tile prefetch during the main loop; tcgen05.ld (32x32b)
//              -> registers -> padded smem tile; then fully coalesced row-wise write-out (512 B per warp
//              instruction) with bias / row map / bf16 pack / GEGLU.  EPI 3 (fused sampling head) adds warps
//              10..17 to the vocabulary-wide reduction.
// Tiles are walked m-fastest so the CTAs running concurrently share the same W tile (L2) while the A panel
// stays L2-resident.
// Epilogues: 0 fp32 (+bias,+residual)   1 bf16 (+bias)   2 GEGLU (attention.py:40-43) on W rows packed as
// [64 value rows | 64 gate rows] per 128-column tile -> bf16 [M, N/2].
#include "phk_common.cuh"
#include <cuda.h>
#include <mutex>
#include <unordered_map>

namespace phk {

constexpr int GM = 128;       // BLOCK_M = UMMA_M
constexpr int GN = 128;       // BLOCK_N = UMMA_N
constexpr int GK = 64;        // BLOCK_K: 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int GSTAGES = 4;
constexpr int EPI_WARPS = 8;        // epilogue warps 2..9: two per TMEM lane group, each drains half the columns
constexpr int HEAD_WARPS = 16;      // fused sampling head: warps 2..17 share the vocabulary-wide epilogue math
constexpr int HEAD_COLS = GN / (HEAD_WARPS / 2);  // columns per thread per tile (64 tokens x HEAD_WARPS/2 groups)
constexpr int GTHREADS = 64 + EPI_WARPS * 32;       // 320
constexpr int GTHREADS_HEAD = 64 + HEAD_WARPS * 32; // 576
constexpr int STAGE_BYTES = GM * GK * 2;          // 16 KB per operand per stage
constexpr int CPAD = 132;                          // fp32 staging row stride (floats): conflict-free 128-bit rows
constexpr int CSTAGE_BYTES = GM * CPAD * 4;        // 67.6 KB
constexpr int RING_BYTES = GSTAGES * 2 * STAGE_BYTES;
constexpr int SMEM_TOTAL = RING_BYTES + CSTAGE_BYTES + 256 /*barriers*/ + 1024 /*manual 1024-B alignment*/;

struct EpiParams {
  void* C; int64_t ldc; int64_t M; int N; int K;
  const float* bias; const float* residual;
  int64_t seg_len, seg_stride, seg_off;
  int m_tiles, n_tiles;
  long long* trace;  // debug: per-CTA clock64 stamps of the first tile (NULL in production)
  // EPI 3 (fused logits head + CFG + gumbel argmax + online softmax, phenaki_pytorch.py:161,83-93,547-550)
  int n_splits, tiles_per_split, n_tokens;
  float cond_scale, inv_T;
  unsigned long long seed, offset;
  float4* part_f;  // [n_tokens, n_splits] {best_y, l_at_best, max_l, sum_exp}
  int* part_i;     // [n_tokens, n_splits] argmax index
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must surface as a trap (launch failure), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
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
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// K-major, SWIZZLE_128B smem operand descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO=1 | SBO=1024B>>4 |
// version=1 (bit 46) | layout SWIZZLE_128B=2 (bits 61..63)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t* out) {  // identical to rowops.cu
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory"); }

// erf with |err| <= 1.5e-7 (Abramowitz-Stegun 7.1.26): one ex2 + one rcp instead of erff's long polynomial; used
// only where the result is rounded to bf16 anyway (GEGLU epilogue of the bf16 mode)
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = 1.0f - poly * t * __expf(-z * z);  // erf(|x|/sqrt2)
  return 0.5f * x * (1.0f + copysignf(e, x));
}

template <int EPI>
__global__ void __launch_bounds__(GTHREADS_HEAD, 1) gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                     const __grid_constant__ CUtensorMap tmB,
                                                                     EpiParams p) {
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
  const int num_kb = (p.K + GK - 1) / GK;
  const int num_tiles = p.m_tiles * p.n_tiles;
  // tile schedule.  EPI 0..2: round-robin over all tiles, m-fastest.  EPI 3: this CTA owns ONE 128-row tile
  // (64 tokens x {cond, null}) and walks a contiguous range of vocabulary tiles, keeping its reductions in registers.
  const int my_tiles = EPI == 3
      ? max(0, min(p.tiles_per_split, p.n_tiles - (int)(blockIdx.x % p.n_splits) * p.tiles_per_split))
      : ((int)blockIdx.x < num_tiles ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0);
  auto tile_of = [&](int i, int& m0, int& n0) {
    if (EPI == 3) {
      m0 = (int)(blockIdx.x / p.n_splits) * GM;
      n0 = ((int)(blockIdx.x % p.n_splits) * p.tiles_per_split + i) * GN;
    } else {
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      m0 = (tile % p.m_tiles) * GM;
      n0 = (tile / p.m_tiles) * GN;
    }
  };
  const long long t_start = clock64();
#define PHK_STAMP(slot) do { if (p.trace) p.trace[blockIdx.x * 16 + (slot)] = clock64() - t_start; } while (0)

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < GSTAGES; ++s) {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 8 * (GSTAGES + s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, EPI_WARPS);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: two 128-column fp32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(2 * GN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();  // everything above (barriers, tensor-map prefetch, TMEM allocation) overlapped the previous kernel
  if (threadIdx.x == 0) PHK_STAMP(0);  // setup done

  // ---- EPI 3: per-thread running reductions over the vocabulary (token t = te & 63, column group = te >> 6) ----
  float sm_best = -FLT_MAX, sm_lbest = 0.f, sm_max = -FLT_MAX, sm_sum = 0.f;
  int sm_idx = 0x7fffffff;
  auto head_bar = [&]() { asm volatile("bar.sync 2, %0;" ::"n"(HEAD_WARPS * 32) : "memory"); };
  // rows t (cond) and 64+t (null) of the staged tile belong to the same token; HEAD_WARPS*32 threads x HEAD_COLS columns
  auto head_reduce = [&](int m0, int n0) {
    const int te = (int)threadIdx.x - 64, t = te & 63, grp = te >> 6;
    const int tok = (m0 / GM) * 64 + t;
    if (tok >= p.n_tokens) return;
    const float* crow = cstage + t * CPAD + grp * HEAD_COLS;
    const float* nrow = cstage + (64 + t) * CPAD + grp * HEAD_COLS;
    const unsigned long long ctr0 = p.offset + (unsigned long long)tok * (unsigned long long)((p.N + 3) / 4);
#pragma unroll
    for (int c = 0; c < HEAD_COLS; c += 4) {
      const int v0 = n0 + grp * HEAD_COLS + c;
      if (v0 >= p.N) break;
      const float4 cv = *reinterpret_cast<const float4*>(crow + c);
      const float4 nv = *reinterpret_cast<const float4*>(nrow + c);
      const float cc[4] = {cv.x, cv.y, cv.z, cv.w}, nn[4] = {nv.x, nv.y, nv.z, nv.w};
      float bb[4] = {0.f, 0.f, 0.f, 0.f};
      if (p.bias) {
        if (v0 + 3 < p.N) { const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + v0)); bb[0] = b4.x; bb[1] = b4.y; bb[2] = b4.z; bb[3] = b4.w; }
        else { for (int j = 0; j < 4; ++j) if (v0 + j < p.N) bb[j] = __ldg(p.bias + v0 + j); }
      }
      uint32_t rnd[4];
      const unsigned long long ctr = ctr0 + (unsigned long long)(v0 >> 2);
      philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), rnd);
      float l4[4];
      float gm = -FLT_MAX;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float cb = cc[j] + bb[j], nb = nn[j] + bb[j];
        l4[j] = (v0 + j < p.N) ? fmaf(cb - nb, p.cond_scale, nb) : -FLT_MAX;
        gm = fmaxf(gm, l4[j]);
        const float u = (float)(rnd[j] >> 8) * (1.0f / 16777216.0f);
        const float g = -__logf(-__logf(u + 1e-10f) + 1e-10f);
        const float y = fmaf(l4[j], p.inv_T, g);
        if (v0 + j < p.N && y > sm_best) { sm_best = y; sm_idx = v0 + j; sm_lbest = l4[j]; }
      }
      if (gm > sm_max) { sm_sum *= __expf(sm_max - gm); sm_max = gm; }  // one rescale per 4 logits
#pragma unroll
      for (int j = 0; j < 4; ++j) sm_sum += __expf(l4[j] - sm_max);
    }
  };
  // combine the column groups of every token through smem, then one partial per (token, vocabulary split)
  auto head_publish = [&]() {
    constexpr int NG = HEAD_WARPS / 2;  // column groups per token
    const int te = (int)threadIdx.x - 64, t = te & 63, grp = te >> 6;
    float* ex = cstage;  // [64][NG-1][6]
    if (grp > 0) {
      float* e = ex + (t * (NG - 1) + grp - 1) * 6;
      e[0] = sm_best; e[1] = sm_lbest; e[2] = sm_max; e[3] = sm_sum; e[4] = __int_as_float(sm_idx);
    }
    head_bar();
    const int tok = (int)(blockIdx.x / p.n_splits) * 64 + t;
    if (grp == 0 && tok < p.n_tokens) {
      for (int qq = 0; qq < NG - 1; ++qq) {
        const float* e = ex + (t * (NG - 1) + qq) * 6;
        const float oy = e[0], ol = e[1], om = e[2], os = e[3];
        const int oi = __float_as_int(e[4]);
        if (oy > sm_best || (oy == sm_best && oi < sm_idx)) { sm_best = oy; sm_idx = oi; sm_lbest = ol; }
        const float nm = fmaxf(sm_max, om);
        sm_sum = sm_sum * __expf(sm_max - nm) + os * __expf(om - nm);
        sm_max = nm;
      }
      const int64_t slot = (int64_t)tok * p.n_splits + (blockIdx.x % p.n_splits);
      p.part_f[slot] = make_float4(sm_best, sm_lbest, sm_max, sm_sum);
      p.part_i[slot] = sm_idx;
    }
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        int m0, n0;
        tile_of(it, m0, n0);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bars + 8 * (GSTAGES + stage), phase ^ 1);  // slot free (passes immediately on the first lap)
          const uint32_t full = bars + 8 * stage;
          mbar_expect_tx(full, 2 * STAGE_BYTES);
          tma_load_2d(&tmA, full, sA + stage * STAGE_BYTES, kb * GK, m0);
          tma_load_2d(&tmB, full, sB + stage * STAGE_BYTES, kb * GK, n0);
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
  } else if (warp >= 2 + EPI_WARPS) {
    // ===================== extra math warps of the fused sampling head =====================
    if (EPI == 3) {
      for (int it = 0; it < my_tiles; ++it) {
        int m0, n0;
        tile_of(it, m0, n0);
        head_bar();            // tile staged by the epilogue warps
        head_reduce(m0, n0);
        head_bar();            // tile consumed
      }
      head_publish();
    }
  } else {
    // ===================== epilogue: warps 2..9, two per TMEM lane group =====================
    const int ew = warp - 2;        // 0..7: rows ew, ew+8, ... in the coalesced write-out
    const int lg = warp & 3;        // TMEM lane group this warp may access (rows lg*32 .. +31 of the tile)
    const int chalf = ew >> 2;      // which half of the accumulator columns this warp drains
    for (int it = 0; it < my_tiles; ++it) {
      const int acc = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      int m0i, n0;
      tile_of(it, m0i, n0);
      const int64_t m0 = m0i;
      // residual prefetch: while the main loop of this tile runs, pull the residual tile into the staging buffer
      // with coalesced 512-B row loads (the previous tile's write-out finished at the barrier below)
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
        epi_bar_sync();  // residual tile complete before the row-per-thread accumulate below
      }
      mbar_wait(bar_tfull + 8 * acc, use & 1);
      if (it == 0 && threadIdx.x == 64) PHK_STAMP(6);      // accumulator ready
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t trow = tmem_base + acc * GN + ((uint32_t)(lg * 32) << 16);
      float* srow = cstage + (lg * 32 + lane) * CPAD;
      if (EPI == 2) {
        // [64 value | 64 gate] -> 64 outputs gelu(gate) * value, staged as fp32 in columns 0..63
        const int c = chalf;
        uint32_t val[32], gate[32];
        tmem_ld32(trow + c * 32, val);
        tmem_ld32(trow + 64 + c * 32, gate);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o;
          o.x = gelu_erf_fast(__uint_as_float(gate[j])) * __uint_as_float(val[j]);
          o.y = gelu_erf_fast(__uint_as_float(gate[j + 1])) * __uint_as_float(val[j + 1]);
          o.z = gelu_erf_fast(__uint_as_float(gate[j + 2])) * __uint_as_float(val[j + 2]);
          o.w = gelu_erf_fast(__uint_as_float(gate[j + 3])) * __uint_as_float(val[j + 3]);
          *reinterpret_cast<float4*>(srow + c * 32 + j) = o;
        }
      } else {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = chalf * 2 + cc;
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
      }
      // accumulator drained: hand it back to the MMA warp before the (slower) global write-out
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      if (EPI == 3) {
        head_bar();                      // all epilogue + math warps: the staged tile is complete
        head_reduce((int)m0, n0);
        head_bar();                      // staging tile free for the next accumulator
        continue;
      }
      epi_bar_sync();  // whole 128 x 128 tile staged
      if (it == 0 && threadIdx.x == 64) PHK_STAMP(7);      // tile staged in smem

      // ---- coalesced write-out: one row per warp instruction ----
      constexpr int RPW = GM / EPI_WARPS;  // rows per warp
      const int ncol0 = EPI == 2 ? n0 / 2 : n0;
      const int nlim = EPI == 2 ? p.N / 2 : p.N;
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
        const uint32_t seg_len = (uint32_t)p.seg_len;
        // fast path (whole tile uniform): vector rows, residual already folded in (or absent), full M tile
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
      } else {
        // bf16 outputs: EPI 1 -> 128 columns (4 per lane), EPI 2 -> 64 columns (2 per lane)
        constexpr int CPL = EPI == 2 ? 2 : 4;
        const int col = ncol0 + lane * CPL;
        float bv[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) bv[j] = (p.bias && col + j < nlim) ? __ldg(p.bias + col + j) : 0.f;
        const bool vec = (p.ldc % CPL == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && (col + CPL - 1 < nlim);
        const uint32_t seg_len = (uint32_t)p.seg_len;
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
            if (CPL == 4) *reinterpret_cast<uint2*>(crow) = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
            else *reinterpret_cast<uint32_t*>(crow) = pack_bf16x2(o[0], o[1]);
          } else {
#pragma unroll
            for (int j = 0; j < CPL; ++j)
              if (col + j < nlim) crow[j] = __float2bfloat16_rn(o[j]);
          }
        }
      }
      epi_bar_sync();  // staging tile free for the next accumulator
      if (it == 0 && threadIdx.x == 64) PHK_STAMP(8);      // tile written out
    }
    if (EPI == 3) head_publish();
  }
  __syncthreads();
  if (threadIdx.x == 0) PHK_STAMP(9);  // CTA done
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * GN) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// host side: tensor maps (driver entry point resolved at run time: libphk.so has no link-time libcuda dependency,
// so it also loads on a CPU-only box for the ABI tests)
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

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

static long long* g_gemm_trace = nullptr;

template <int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const EpiParams& p, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    PHK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    configured = true;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = EPI == 3 ? p.m_tiles * p.n_splits : (tiles < kNumSMs ? tiles : kNumSMs);
  PHK_CUDA(launch_pdl(gemm_bf16_kernel<EPI>, dim3(grid), dim3(EPI == 3 ? GTHREADS_HEAD : GTHREADS), (size_t)(SMEM_TOTAL), st, ta, tb, p));
  PHK_LAUNCH_CHECK();
  return 0;
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
  PHK_REQUIRE(epilogue != 2 || (N % 128 == 0 && !bias && !residual), PHK_E_ARG,
              "phk_gemm_bf16: GEGLU epilogue needs N % 128 == 0 and no bias/residual");
  PHK_REQUIRE(M < (1LL << 31) - GM, PHK_E_UNSUPPORTED, "phk_gemm_bf16: M too large");
  if (M == 0) return 0;
  CUtensorMap ta, tb;
  PHK_TRY(get_tensor_map(A, M, K, lda, GM, &ta));
  PHK_TRY(get_tensor_map(W, N, K, ldw, GN, &tb));
  EpiParams p{C, ldc, M, N, K, bias, residual, seg_len, seg_stride, seg_off, (int)((M + GM - 1) / GM), (N + GN - 1) / GN, g_gemm_trace, 1, 0, 0, 1.f, 1.f, 0ull, 0ull, nullptr, nullptr};
  PHK_REQUIRE((int64_t)p.m_tiles * p.n_tiles < (1LL << 31), PHK_E_UNSUPPORTED, "phk_gemm_bf16: too many tiles");
  cudaStream_t st = to_stream(s);
  if (epilogue == 2) return launch_gemm<2>(ta, tb, p, st);
  if (epilogue == 1) return launch_gemm<1>(ta, tb, p, st);
  return launch_gemm<0>(ta, tb, p, st);
}

// debug: device buffer of 16 x int64 per CTA receiving clock64 stamps (relative to CTA start) of the first tile
extern "C" int phk_debug_gemm_trace(long long* device_buffer) { g_gemm_trace = device_buffer; return 0; }
