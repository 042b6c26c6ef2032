// bf16 GEMM on the 5th-generation tensor cores (sm_100a only):
//   C[map(m), n] = sum_k A[m,k] * W[n,k]  (+bias) (+residual)        nn.Linear semantics
// A, W bf16, K-major (row-major [rows, K]); fp32 accumulation in TMEM.
//
//   * TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) stages 128x64 A and BLOCK_Nx64 W tiles into a
//     3-deep shared-memory ring; out-of-bounds rows / K tail are zero-filled by the TMA unit.
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BLOCK_N, K=16), four per
//     64-wide k-block; tcgen05.commit releases the smem slot / publishes the accumulator via mbarriers.
//   * warp specialisation: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 =
//     epilogue (tcgen05.ld 32x32b -> registers -> fused bias / residual / row-map / GEGLU -> global).
//   * 96 KB of smem and 128 TMEM columns per CTA -> two CTAs per SM, so one CTA's epilogue overlaps the
//     other's main loop.
// Epilogues: 0 fp32 (+bias,+residual)   1 bf16 (+bias)   2 GEGLU (attention.py:40-43) on W rows packed as
// [64 value rows | 64 gate rows] per 128-column tile -> bf16 [M, N/2].
#include "phk_common.cuh"
#include <cuda.h>
#include <mutex>
#include <unordered_map>

namespace phk {

constexpr int GM = 128;       // BLOCK_M = UMMA_M
constexpr int GK = 64;        // BLOCK_K: 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int GSTAGES = 3;
constexpr int GTHREADS = 192;
constexpr int A_STAGE_BYTES = GM * GK * 2;  // 16 KB

struct EpiParams {
  void* C; int64_t ldc; int64_t M; int N; int K;
  const float* bias; const float* residual;
  int64_t seg_len, seg_stride, seg_off;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must surface as a trap (launch failure), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (true) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000LL) __trap();
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

template <int BLOCK_N>
struct GemmSmem {
  static constexpr int B_STAGE_BYTES = BLOCK_N * GK * 2;
  static constexpr int TILE_BYTES = GSTAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
  static constexpr int TOTAL = TILE_BYTES + 128 /*barriers*/ + 1024 /*manual 1024-B alignment*/;
};

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(GTHREADS, 1) gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                EpiParams p) {
  extern __shared__ uint8_t smem_raw[];
  using SM = GemmSmem<BLOCK_N>;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-B alignment
  const uint32_t sA = base, sB = base + GSTAGES * A_STAGE_BYTES;
  const uint32_t bars = base + SM::TILE_BYTES;
  // full[s] @ bars + 8s ; empty[s] @ bars + 8(S+s) ; tmem_full @ bars + 16S ; tmem ptr slot @ bars + 16S + 8
  const uint32_t bar_tmem_full = bars + 16 * GSTAGES;
  const uint32_t tmem_slot = bar_tmem_full + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = (int64_t)blockIdx.y * GM;
  const int n0 = blockIdx.x * BLOCK_N;
  const int num_kb = (p.K + GK - 1) / GK;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < GSTAGES; ++s) {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 8 * (GSTAGES + s), 1);
    }
    mbar_init(bar_tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: BLOCK_N fp32 accumulator columns (power of two >= 32)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(BLOCK_N)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(bars + 8 * (GSTAGES + stage), phase ^ 1);  // slot free (passes immediately on the first lap)
        const uint32_t full = bars + 8 * stage;
        mbar_expect_tx(full, A_STAGE_BYTES + SM::B_STAGE_BYTES);
        tma_load_2d(&tmA, full, sA + stage * A_STAGE_BYTES, kb * GK, (int)m0);
        tma_load_2d(&tmB, full, sB + stage * SM::B_STAGE_BYTES, kb * GK, n0);
        if (++stage == GSTAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // cute::UMMA::InstrDescriptor: c=F32 (bit 4), a=b=BF16 (bits 7,10), K-major both, N>>3 @17, M>>4 @24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) |
                             ((uint32_t)(GM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(bars + 8 * stage, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t da = umma_desc(sA + stage * A_STAGE_BYTES);
        const uint64_t db = umma_desc(sB + stage * SM::B_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < GK / 16; ++k)  // +32 B per UMMA_K inside the 128-B swizzle row => +2 in the address field
          umma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
        umma_commit(bars + 8 * (GSTAGES + stage));  // frees the smem slot once these MMAs have read it
        if (++stage == GSTAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(bar_tmem_full);  // accumulator complete
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int lg = warp & 3;  // TMEM lane group this warp may access
    mbar_wait(bar_tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int64_t m = m0 + lg * 32 + lane;
    const bool row_ok = m < p.M;
    int64_t orow = m;
    if (row_ok && p.seg_len > 0) {
      const int64_t q = m / p.seg_len;
      orow = q * p.seg_stride + p.seg_off + (m - q * p.seg_len);
    }
    const uint32_t trow = tmem_base + ((uint32_t)(lg * 32) << 16);
    if (EPI == 2) {
      // [64 value | 64 gate] per 128-wide tile -> out[m, n0/2 + c] = gelu(gate_c) * value_c
      static_assert(EPI != 2 || BLOCK_N == 128, "GEGLU epilogue needs 128-wide tiles");
      __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.C) + orow * p.ldc + n0 / 2;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t val[32], gate[32];
        tmem_ld32(trow + c * 32, val);
        tmem_ld32(trow + 64 + c * 32, gate);
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = gelu_erf(__uint_as_float(gate[j + 2 * e])) * __uint_as_float(val[j + 2 * e]);
              const float b = gelu_erf(__uint_as_float(gate[j + 2 * e + 1])) * __uint_as_float(val[j + 2 * e + 1]);
              pk[e] = pack_bf16x2(a, b);
            }
            *reinterpret_cast<uint4*>(out + c * 32 + j) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
    } else {
      const bool vec_ok = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(trow + c * 32, v);
        if (!row_ok) continue;
        const int nb = n0 + c * 32;
        if (nb >= p.N) continue;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        const bool full = nb + 32 <= p.N;
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (full || nb + j < p.N) f[j] += __ldg(p.bias + nb + j);
        }
        if (EPI == 0) {
          float* crow = reinterpret_cast<float*>(p.C) + orow * p.ldc + nb;
          if (full && vec_ok) {
            if (p.residual) {
              const float4* rr = reinterpret_cast<const float4*>(p.residual + orow * p.ldc + nb);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 r = rr[j];
                f[4 * j] += r.x; f[4 * j + 1] += r.y; f[4 * j + 2] += r.z; f[4 * j + 3] += r.w;
              }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              reinterpret_cast<float4*>(crow)[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          } else {
            for (int j = 0; j < 32 && nb + j < p.N; ++j) {
              float o = f[j];
              if (p.residual) o += p.residual[orow * p.ldc + nb + j];
              crow[j] = o;
            }
          }
        } else {
          __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(p.C) + orow * p.ldc + nb;
          if (full && (p.ldc % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              reinterpret_cast<uint4*>(crow)[j] =
                  make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                             pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
          } else {
            for (int j = 0; j < 32 && nb + j < p.N; ++j) crow[j] = __float2bfloat16_rn(f[j]);
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BLOCK_N) : "memory");
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

template <int BLOCK_N, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const EpiParams& p, cudaStream_t st) {
  using SM = GemmSmem<BLOCK_N>;
  static bool configured = false;
  if (!configured) {
    PHK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BLOCK_N, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    configured = true;
  }
  dim3 grid((unsigned)((p.N + BLOCK_N - 1) / BLOCK_N), (unsigned)((p.M + GM - 1) / GM));
  gemm_bf16_kernel<BLOCK_N, EPI><<<grid, GTHREADS, SM::TOTAL, st>>>(ta, tb, p);
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
  PHK_REQUIRE(epilogue != 2 || (N % 128 == 0 && !bias && !residual && ldc % 8 == 0 &&
                                (reinterpret_cast<uintptr_t>(C) & 15) == 0),
              PHK_E_ARG, "phk_gemm_bf16: GEGLU epilogue needs N % 128 == 0, aligned bf16 output, no bias/residual");
  PHK_REQUIRE((M + GM - 1) / GM <= 65535, PHK_E_UNSUPPORTED, "phk_gemm_bf16: M too large");
  if (M == 0) return 0;
  const int block_n = (epilogue == 2 || N > 64) ? 128 : 64;
  CUtensorMap ta, tb;
  PHK_TRY(get_tensor_map(A, M, K, lda, GM, &ta));
  PHK_TRY(get_tensor_map(W, N, K, ldw, block_n, &tb));
  EpiParams p{C, ldc, M, N, K, bias, residual, seg_len, seg_stride, seg_off};
  cudaStream_t st = to_stream(s);
  if (epilogue == 2) return launch_gemm<128, 2>(ta, tb, p, st);
  if (epilogue == 1) return block_n == 128 ? launch_gemm<128, 1>(ta, tb, p, st) : launch_gemm<64, 1>(ta, tb, p, st);
  return block_n == 128 ? launch_gemm<128, 0>(ta, tb, p, st) : launch_gemm<64, 0>(ta, tb, p, st);
}
