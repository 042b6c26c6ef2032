// TEST INFRASTRUCTURE: fiber scheduler of cuda_emu.h plus CPU statements of the forward building blocks that
// phk_maskgit_train_step calls (same contracts as include/phk.h; the CUDA versions are validated on the GPU by the
// -m gpu suite, these stand in for them when the driver runs under the emulator).
#include "cuda_emu.h"

namespace emu {
State S;
namespace {
struct Fiber {
  ucontext_t ctx;
  bool done = false;
  unsigned tid = 0;
};
constexpr size_t kStack = 128 * 1024;
std::vector<Fiber> fibers;
std::vector<char> stacks;
ucontext_t sched_ctx;
int cur = -1;
Group g_block;
std::vector<Group> g_warps;
std::vector<float> xchg;
const std::function<void()>* g_body = nullptr;
std::vector<char> dyn;

unsigned long events = 0;  // barrier releases + fiber exits: no event in a whole scheduler round = deadlock
void on_exit_group(Group& g) {
  g.alive -= 1;
  if (g.alive > 0 && g.count == g.alive) { g.count = 0; g.gen += 1; }
}
void trampoline() {
  (*g_body)();
  Fiber& f = fibers[cur];
  f.done = true;
  events += 1;
  on_exit_group(g_block);
  on_exit_group(g_warps[f.tid >> 5]);
  swapcontext(&f.ctx, &sched_ctx);
}
}  // namespace

void yield() { swapcontext(&fibers[cur].ctx, &sched_ctx); }
void barrier(Group& g) {
  const unsigned gen = g.gen;
  g.count += 1;
  if (g.count == g.alive) { g.count = 0; g.gen += 1; events += 1; return; }
  while (g.gen == gen) yield();
}
Group& block_group() { return g_block; }
Group& warp_group() { return g_warps[fibers[cur].tid >> 5]; }
float* warp_xchg() { return xchg.data() + (size_t)(fibers[cur].tid >> 5) * 32; }

// PHK_EMU_SHUFFLE=<seed>: blocks run in a random order and, inside a block, the runnable threads are resumed in a fresh
// random order every scheduler round.  The default (in-order) schedule would hide a missing __syncthreads() or an
// assumption about block order; the tests run both.
static uint64_t g_rng = 0;
static bool g_shuffle = false;
static uint64_t next_rand() {
  g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17;
  return g_rng;
}
static void set_shuffle(uint64_t seed) {
  g_shuffle = seed != 0;
  g_rng = 0x9E3779B97F4A7C15ull ^ (seed * 0x100000001b3ull);
  if (!g_rng) g_rng = 1;
}
static void init_schedule() {
  static bool done = false;
  if (done) return;
  done = true;
  const char* e = getenv("PHK_EMU_SHUFFLE");
  if (e && *e) set_shuffle((uint64_t)strtoull(e, nullptr, 10));
}
template <typename T> static void shuffle(std::vector<T>& v) {
  for (size_t i = v.size(); i > 1; --i) std::swap(v[i - 1], v[next_rand() % i]);
}

void set_schedule_seed(uint64_t seed) { init_schedule(); set_shuffle(seed); }

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  init_schedule();
  const unsigned nt = block.x * block.y * block.z;
  if (block.y != 1 || block.z != 1) { fprintf(stderr, "cuda_emu: only 1-D blocks\n"); abort(); }
  if (fibers.size() < nt) { fibers.resize(nt); stacks.resize((size_t)nt * kStack); }
  dyn.assign(smem + 16, 0);
  const unsigned nw = (nt + 31) / 32;
  std::vector<uint64_t> order((size_t)grid.x * grid.y * grid.z);
  for (size_t i = 0; i < order.size(); ++i) order[i] = i;
  if (g_shuffle) shuffle(order);
  std::vector<unsigned> tids(nt);
  for (uint64_t lin : order) {
        const unsigned bx = (unsigned)(lin % grid.x), by = (unsigned)((lin / grid.x) % grid.y), bz = (unsigned)(lin / ((uint64_t)grid.x * grid.y));
        g_body = &body;
        g_block = Group{(int)nt, 0, 0};
        g_warps.assign(nw, Group{});
        for (unsigned w = 0; w < nw; ++w) g_warps[w].alive = (int)((w + 1) * 32 <= nt ? 32 : nt - w * 32);
        xchg.assign((size_t)nw * 32, 0.f);
        for (unsigned t = 0; t < nt; ++t) {
          Fiber& f = fibers[t];
          f.done = false; f.tid = t;
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = stacks.data() + (size_t)t * kStack;
          f.ctx.uc_stack.ss_size = kStack;
          f.ctx.uc_link = nullptr;
          makecontext(&f.ctx, trampoline, 0);
        }
        unsigned remaining = nt;
        while (remaining) {
          const unsigned long before = events;
          for (unsigned t = 0; t < nt; ++t) tids[t] = t;
          if (g_shuffle) shuffle(tids);
          for (unsigned ti = 0; ti < nt; ++ti) {
            const unsigned t = tids[ti];
            Fiber& f = fibers[t];
            if (f.done) continue;
            cur = (int)t;
            S.t_idx = dim3(t, 0, 0); S.b_idx = dim3(bx, by, bz); S.b_dim = block; S.g_dim = grid;
            S.dyn_smem = dyn.data();
            swapcontext(&sched_ctx, &f.ctx);
            if (f.done) { remaining -= 1; }
          }
          if (events == before) { fprintf(stderr, "cuda_emu: deadlock (a barrier some threads never reach)\n"); abort(); }
        }
      }
}
}  // namespace emu

// ---------------------------------------------------------------------------------------------------------------------
// forward building blocks (include/phk.h contracts), straight loops
// ---------------------------------------------------------------------------------------------------------------------
namespace phk {
static char g_err[256];
void set_error(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }
}  // namespace phk
extern "C" const char* phk_last_error(void) { return phk::g_err; }
// test hook: 0 = in-order schedule, otherwise the seed of the random block / thread order
extern "C" void phk_emu_set_shuffle(uint64_t seed) { emu::set_schedule_seed(seed); }

extern "C" int phk_layernorm(const float* x, const float* gamma, const float* beta, void* out, void* raw_bf16, int64_t rows,
                             int32_t dim, int32_t out_bf16, int64_t seg_len, int64_t, int64_t, phk_stream_t) {
  if (out_bf16 || raw_bf16 || seg_len != 0) return PHK_E_UNSUPPORTED;
  float* o = (float*)out;
  for (int64_t r = 0; r < rows; ++r) {
    float mean = 0.f, var = 0.f;
    for (int c = 0; c < dim; ++c) mean += x[r * dim + c];
    mean /= dim;
    for (int c = 0; c < dim; ++c) { const float d = x[r * dim + c] - mean; var += d * d; }
    const float rstd = 1.0f / sqrtf(var / dim + 1e-5f);
    for (int c = 0; c < dim; ++c) o[r * dim + c] = (x[r * dim + c] - mean) * rstd * gamma[c] + beta[c];
  }
  return 0;
}

extern "C" int phk_gemm_f32(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc, int64_t M,
                            int32_t N, int32_t K, const float* bias, const float* residual, int64_t seg_len, int64_t,
                            int64_t, phk_stream_t) {
  if (seg_len > 0) return PHK_E_UNSUPPORTED;
  for (int64_t m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float a = 0.f;
      for (int k = 0; k < K; ++k) a += A[m * lda + k] * W[(int64_t)n * ldw + k];
      if (bias) a += bias[n];
      if (residual) a += residual[m * ldc + n];
      C[m * ldc + n] = a;
    }
  return 0;
}

// tcgen05 GEMM contract (include/phk.h): bf16 operands, fp32 accumulate, epilogue 0 (fp32 out + bias + residual)
extern "C" int phk_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int64_t M,
                             int32_t N, int32_t K, const float* bias, const float* residual, int64_t seg_len, int64_t,
                             int64_t, int32_t epilogue, phk_stream_t) {
  if (seg_len > 0 || epilogue != 0 || lda % 8 || ldw % 8 || lda < K || ldw < K) return PHK_E_ARG;
  const __nv_bfloat16* a = (const __nv_bfloat16*)A;
  const __nv_bfloat16* w = (const __nv_bfloat16*)W;
  float* c = (float*)C;
  for (int64_t m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k) acc += __bfloat162float(a[m * lda + k]) * __bfloat162float(w[(int64_t)n * ldw + k]);
      if (bias) acc += bias[n];
      if (residual) acc += residual[m * ldc + n];
      c[m * ldc + n] = acc;
    }
  return 0;
}

extern "C" int phk_geglu(const float* h, float* out, int64_t rows, int32_t inner, phk_stream_t) {
  for (int64_t r = 0; r < rows; ++r)
    for (int j = 0; j < inner; ++j) {
      const float val = h[r * 2 * inner + j], gate = h[r * 2 * inner + inner + j];
      out[r * inner + j] = 0.5f * gate * (1.0f + erff(gate * 0.70710678118654752440f)) * val;
    }
  return 0;
}

extern "C" int phk_token_embed(const int64_t* ids, const float* tok, const float* pos, float* out, int32_t b, int32_t n,
                               int32_t dim, int32_t, float alpha, int32_t replicas, phk_stream_t) {
  if (replicas != 1) return PHK_E_UNSUPPORTED;
  for (int64_t r = 0; r < (int64_t)b * n; ++r)
    for (int c = 0; c < dim; ++c) {
      float x = pos[(r % n) * dim + c] + tok[ids[r] * dim + c];
      if (alpha >= 0.f) x = x * alpha + x * (1.0f - alpha);
      out[r * dim + c] = x;
    }
  return 0;
}

extern "C" int phk_peg3d(const float* x, const float* w, const float* b, float* y, int32_t B, int32_t T, int32_t H, int32_t W,
                         int32_t D, int32_t causal, int32_t layout, phk_stream_t) {
  if (layout != 0) return PHK_E_UNSUPPORTED;
  const int pad = causal ? 2 : 1;
  for (int bi = 0; bi < B; ++bi)
    for (int t = 0; t < T; ++t)
      for (int h = 0; h < H; ++h)
        for (int ww = 0; ww < W; ++ww) {
          const int64_t o = (((int64_t)bi * T + t) * H + h) * W + ww;
          for (int d = 0; d < D; ++d) {
            float a = b[d] + x[o * D + d];
            for (int kt = 0; kt < 3; ++kt)
              for (int kh = 0; kh < 3; ++kh)
                for (int kw = 0; kw < 3; ++kw) {
                  const int ts = t + kt - pad, hs = h + kh - 1, ws = ww + kw - 1;
                  if (ts < 0 || ts >= T || hs < 0 || hs >= H || ws < 0 || ws >= W) continue;
                  a += w[((kt * 3 + kh) * 3 + kw) * D + d] * x[((((int64_t)bi * T + ts) * H + hs) * W + ws) * D + d];
                }
            y[o * D + d] = a;
          }
        }
  return 0;
}

extern "C" int64_t phk_cpb_scratch_floats(const phk_cpb_t* c, int32_t d0, int32_t d1, int32_t d2) {
  return (int64_t)(2 * d0 - 1) * (2 * d1 - 1) * (2 * d2 - 1) * c->heads;
}

extern "C" int phk_cpb_bias(const phk_cpb_t* c, int32_t d0, int32_t d1, int32_t d2, float*, float* out, phk_stream_t) {
  const int n = d0 * d1 * d2, hid = c->hidden;
  std::vector<float> h1(hid), h2(hid);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      const int di[3] = {i / (d1 * d2) - j / (d1 * d2), (i / d2) % d1 - (j / d2) % d1, i % d2 - j % d2};
      float in[3];
      for (int k = 0; k < 3; ++k) {
        const float sg = di[k] > 0 ? 1.f : (di[k] < 0 ? -1.f : 0.f);
        in[k] = sg * logf((float)abs(di[k]) + 1.f);
      }
      for (int a = 0; a < hid; ++a) {
        float v = c->b0[a];
        for (int k = 0; k < c->num_dims; ++k) v += in[k] * c->w0[a * c->num_dims + k];
        h1[a] = v > 0.f ? v : 0.1f * v;
      }
      for (int a = 0; a < hid; ++a) {
        float v = c->b1[a];
        for (int k = 0; k < hid; ++k) v += h1[k] * c->w1[(int64_t)a * hid + k];
        h2[a] = v > 0.f ? v : 0.1f * v;
      }
      for (int h = 0; h < c->heads; ++h) {
        float v = c->b2[h];
        for (int k = 0; k < hid; ++k) v += h2[k] * c->w2[(int64_t)h * hid + k];
        out[((int64_t)h * n + i) * n + j] = v;
      }
    }
  return 0;
}

// attention.py:146-181 for the (outer, inner, token) geometry of include/phk.h (fp32 out, no causal path)
extern "C" int phk_attention(const float* q, const float* kv, const float* null_kv, const float* q_scale,
                             const float* k_scale, const float* bias, const uint8_t* key_mask, const float*, void* out,
                             const phk_attn_geom_t* g, phk_stream_t) {
  if (g->causal || g->out_bf16) return PHK_E_UNSUPPORTED;
  const int dh = g->dim_head, I = g->heads * dh, nn = g->num_null_kv, nkt = nn + g->n_k;
  std::vector<float> kh((size_t)nkt * dh), vv((size_t)nkt * dh), s(nkt), qh(dh);
  float* o = (float*)out;
  for (int so = 0; so < g->n_outer; ++so)
    for (int si = 0; si < g->n_inner; ++si)
      for (int h = 0; h < g->heads; ++h) {
        const int kso = g->kv_outer_mod > 0 ? so % g->kv_outer_mod : so;
        const int mrow = g->mask_outer_mod > 0 ? so % g->mask_outer_mod : so;
        const bool dropped = g->mask_off_from >= 0 && so >= g->mask_off_from;
        for (int j = 0; j < nkt; ++j) {
          const float* kp = j < nn ? null_kv + ((int64_t)h * 2 * nn + 2 * j) * dh
                                   : kv + (int64_t)kso * g->k_outer + (int64_t)si * g->k_inner + (int64_t)(j - nn) * g->k_tok + (int64_t)h * dh;
          const float* vp = j < nn ? kp + dh : kp + I;
          float ss = 0.f;
          for (int d = 0; d < dh; ++d) ss += kp[d] * kp[d];
          const float nrm = fmaxf(sqrtf(ss), 1e-12f);
          for (int d = 0; d < dh; ++d) { kh[(size_t)j * dh + d] = kp[d] / nrm * k_scale[d]; vv[(size_t)j * dh + d] = vp[d]; }
        }
        for (int i = 0; i < g->n_q; ++i) {
          const float* qp = q + (int64_t)so * g->q_outer + (int64_t)si * g->q_inner + (int64_t)i * g->q_tok + (int64_t)h * dh;
          float ss = 0.f;
          for (int d = 0; d < dh; ++d) ss += qp[d] * qp[d];
          const float nrm = fmaxf(sqrtf(ss), 1e-12f);
          for (int d = 0; d < dh; ++d) qh[d] = qp[d] / nrm * q_scale[d];
          float mx = -FLT_MAX;
          for (int j = 0; j < nkt; ++j) {
            float a = 0.f;
            for (int d = 0; d < dh; ++d) a += qh[d] * kh[(size_t)j * dh + d];
            a *= g->scale;
            if (bias && j >= nn) a += bias[((int64_t)h * g->n_q + i) * g->n_k + (j - nn)];
            if (key_mask && j >= nn && (dropped || !key_mask[(int64_t)mrow * g->n_k + (j - nn)])) a = -FLT_MAX;
            s[j] = a;
            mx = fmaxf(mx, a);
          }
          float sum = 0.f;
          for (int j = 0; j < nkt; ++j) { s[j] = expf(s[j] - mx); sum += s[j]; }
          float* op = o + (int64_t)so * g->o_outer + (int64_t)si * g->o_inner + (int64_t)i * g->o_tok + (int64_t)h * dh;
          for (int d = 0; d < dh; ++d) {
            float a = 0.f;
            for (int j = 0; j < nkt; ++j) a += s[j] / sum * vv[(size_t)j * dh + d];
            op[d] = a;
          }
        }
      }
  return 0;
}
