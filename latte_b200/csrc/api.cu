// extern "C" surface of liblatte_b200.so (declared in include/latte_b200.h) and the forward orchestration.
#include "common.h"

#include <mutex>
#include <vector>

namespace b200 {
namespace {

// ---- optional per-kernel-class timing (bench.py roofline): CUDA events recorded on the launching stream
// around every launch of b200_latte_forward while enabled. Off by default: zero cost, no events.
enum { PROF_GEMM = 0, PROF_ATTN = 1, PROF_LN = 2, PROF_OTHER = 3, PROF_CLASSES = 4 };
struct ProfRec { int cls; cudaEvent_t e0, e1; };
struct Profiler {
  std::mutex mu;
  bool on = false;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
} g_prof;

struct ProfScope {
  cudaStream_t s;
  cudaEvent_t e1 = nullptr;
  ProfScope(int cls, cudaStream_t stream) : s(stream) {
    if (!g_prof.on) return;
    std::lock_guard<std::mutex> lk(g_prof.mu);
    ProfRec r{cls, g_prof.get(), g_prof.get()};
    cudaEventRecord(r.e0, s);
    e1 = r.e1;
    g_prof.recs.push_back(r);
  }
  ~ProfScope() { if (e1) cudaEventRecord(e1, s); }
};
#define B200_PROF(cls, expr) do { ProfScope _ps(cls, stream); B200_TRY(expr); } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Workspace {
  float* x;          // [T, D]   fp32 residual stream, rows (b, f, n)
  uint16_t* h;       // [T, D]   16-bit: LN+modulate output, then attention output
  uint16_t* qkv;     // [T, 3D]  16-bit
  uint16_t* g;       // [T, 4D]  16-bit MLP hidden
  float* tfreq;      // [B, 256]
  float* th;         // [B, D]   SiLU(Linear(256, D))
  float* c;          // [B, D]   t_emb (+ y_emb)
  float* mod;        // [B, depth*6D + 2D]
  unsigned long long* sk_flags;   // [B200_GEMM_SK_FLAGS] stream-K ordering flags (zeroed at the start of every forward)
  float* head;       // [T, 32]  fp32 output of the head GEMM (zeroed, then reduce-added into), followed by 32 ones (its "gate")
  size_t bytes;
};

// ---- output head (latte.py:197-201,297-310,374-376): LayerNorm + modulate -> Linear(D, p*p*C_out) -> unpatchify.
// With a 16-bit weight copy and p*p*C_out a multiple of 32 the Linear runs on the tensor cores: the GEMM's gated-residual
// epilogue on a zeroed fp32 buffer with gate = 1 IS "fp32 out = acc + bias".  Otherwise the fp32 CUDA-core kernel.
int output_head(const float* x, uint16_t* h, float* head, const float* shift, const float* scale, long long mod_bs,
                const float* w32, const void* w16, const float* bias, float* out, int batch, int F, int grid, int patch,
                int out_ch, int D, int bf16, int channels_first, unsigned long long* sk_flags, cudaStream_t stream);

int shape_ok(const B200LatteShape* s, int batch) {
  B200_REQUIRE(s != nullptr, B200_ERR_SHAPE, "shape is NULL");
  B200_REQUIRE(batch > 0, B200_ERR_SHAPE, "batch %d must be positive", batch);
  B200_REQUIRE(s->depth > 0 && s->depth % 2 == 0, B200_ERR_SHAPE, "depth %d must be even (spatial/temporal pairs)", s->depth);
  B200_REQUIRE(s->heads > 0 && s->hidden % s->heads == 0, B200_ERR_SHAPE, "hidden %d not divisible by heads %d", s->hidden, s->heads);
  const int hd = s->hidden / s->heads;
  B200_REQUIRE(hd == 64 || hd == 72 || hd == 80, B200_ERR_UNSUPPORTED, "head_dim %d unsupported", hd);
  B200_REQUIRE(s->hidden % 64 == 0 && s->mlp_hidden % 64 == 0, B200_ERR_UNSUPPORTED,
               "hidden %d and mlp_hidden %d must be multiples of 64 (GEMM K tile)", s->hidden, s->mlp_hidden);
  B200_REQUIRE(s->patch == 2, B200_ERR_UNSUPPORTED, "patch size %d not built (only 2)", s->patch);
  B200_REQUIRE(s->input_size % s->patch == 0, B200_ERR_SHAPE, "input_size %d not divisible by patch", s->input_size);
  B200_REQUIRE(s->dtype == B200_FP16 || s->dtype == B200_BF16, B200_ERR_DTYPE, "dtype %d unknown", s->dtype);
  B200_REQUIRE(s->out_channels * s->patch * s->patch <= 32, B200_ERR_UNSUPPORTED, "p*p*out_channels > 32");
  return B200_OK;
}

void carve(const B200LatteShape* s, int batch, void* base, Workspace* ws) {
  const size_t grid = s->input_size / s->patch;
  const size_t T = static_cast<size_t>(batch) * s->frames * grid * grid;
  const size_t D = s->hidden;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<uint8_t*>(base) + off : nullptr;
    off += align_up(bytes, 1024);
    return p;
  };
  ws->x = static_cast<float*>(take(T * D * 4));
  ws->h = static_cast<uint16_t*>(take(T * D * 2));
  ws->qkv = static_cast<uint16_t*>(take(T * 3 * D * 2));
  ws->g = static_cast<uint16_t*>(take(T * static_cast<size_t>(s->mlp_hidden) * 2));
  ws->tfreq = static_cast<float*>(take(static_cast<size_t>(batch) * 256 * 4));
  ws->th = static_cast<float*>(take(static_cast<size_t>(batch) * D * 4));
  ws->c = static_cast<float*>(take(static_cast<size_t>(batch) * D * 4));
  ws->mod = static_cast<float*>(take(static_cast<size_t>(batch) * (static_cast<size_t>(s->depth) * 6 * D + 2 * D) * 4));
  ws->sk_flags = static_cast<unsigned long long*>(take(static_cast<size_t>(B200_GEMM_SK_FLAGS) * 8));
  ws->head = static_cast<float*>(take((T + 1) * 32 * 4));
  ws->bytes = off;
}

// ---- conditioning, once per SAMPLE (latte.py:332-339): c = t_embedder(t) (+ y_embedder(y)); mod = adaLN(SiLU(c)) for all
// blocks and the final layer.  `n` rows; scratch tfreq [n,256], th [n,D], c [n,D]; mod [n, depth*6D + 2D].
int conditioning(const B200LatteShape* s, const B200LatteWeights* w, const int64_t* t, const int64_t* y, int n, float* tfreq,
                 float* th, float* c, float* mod, cudaStream_t stream) {
  const int D = s->hidden;
  const int bf16 = s->dtype == B200_BF16;
  const long long mod_bs = static_cast<long long>(s->depth) * 6 * D + 2 * D;
  B200_PROF(PROF_OTHER, launch_timestep_freq(reinterpret_cast<const long long*>(t), tfreq, n, stream));
  B200_PROF(PROF_OTHER, launch_gemv(w->t_w0, 32, 0, w->t_b0, tfreq, th, n, D, 256, 0, 1, nullptr, nullptr, 0, stream));
  B200_PROF(PROF_OTHER, launch_gemv(w->t_w2, 32, 0, w->t_b2, th, c, n, D, D, 0, 0, s->num_embed > 0 ? w->y_table : nullptr,
                       reinterpret_cast<const long long*>(y), s->num_embed, stream));
  B200_PROF(PROF_OTHER, launch_gemv(w->ada_w16, 16, bf16, w->ada_b, c, mod, n, static_cast<int>(mod_bs), D, 1, 0, nullptr,
                       nullptr, 0, stream));
  return B200_OK;
}

size_t conditioning_scratch_bytes(const B200LatteShape* s, int n) {
  return align_up(static_cast<size_t>(n) * 256 * 4, 1024) + 2 * align_up(static_cast<size_t>(n) * s->hidden * 4, 1024);
}

// premod != NULL: the caller already holds this batch's conditioning rows (b200_latte_conditioning, e.g. for a whole
// sampling trajectory at once -- SURVEY.md 8f rank 2); t is then unused and the adaLN weights are not read.
int forward(const B200LatteShape* s, const B200LatteWeights* w, const float* x, const int64_t* t, const int64_t* y,
            const float* premod, int batch, int use_cfg, float cfg_scale, float* out, void* workspace,
            size_t workspace_bytes, cudaStream_t stream) {
  B200_TRY(shape_ok(s, batch));
  B200_REQUIRE(w && x && (t || premod) && out && workspace, B200_ERR_SHAPE, "NULL argument");
  B200_REQUIRE(premod || (s->num_embed > 0) == (y != nullptr && w->y_table != nullptr), B200_ERR_SHAPE,
               "labels y and y_table must be given iff num_embed > 0 (extras == 2)");
  B200_REQUIRE(!premod || (reinterpret_cast<uintptr_t>(premod) & 15) == 0, B200_ERR_ALIGN, "conditioning rows must be 16-byte aligned");
  B200_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, B200_ERR_ALIGN, "workspace must be 1024-byte aligned");
  const bool cfg = use_cfg != 0;
  B200_REQUIRE(!cfg || batch % 2 == 0, B200_ERR_SHAPE, "classifier-free guidance needs an even batch (got %d)", batch);
  B200_TRY(check_arch());
  Workspace ws;
  carve(s, batch, workspace, &ws);
  B200_REQUIRE(ws.bytes <= workspace_bytes, B200_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", ws.bytes,
               workspace_bytes);

  const int D = s->hidden, H = s->heads, hd = D / H, F = s->frames, depth = s->depth;
  const int grid = s->input_size / s->patch, N = grid * grid;
  const int T = batch * F * N;
  const int rows_per_batch = F * N;
  const int bf16 = s->dtype == B200_BF16;
  const long long mod_bs = static_cast<long long>(depth) * 6 * D + 2 * D;
  const int HID = s->mlp_hidden;

  // stream-K ordering flags: zero at the start of the step (every stream-K GEMM leaves them zero again; this memset only
  // makes the step independent of whatever the workspace held before -- a fresh allocation, an aborted run)
  B200_CHECK_CUDA(cudaMemsetAsync(ws.sk_flags, 0, static_cast<size_t>(B200_GEMM_SK_FLAGS) * 8, stream));
  if (premod) ws.mod = const_cast<float*>(premod);
  else B200_TRY(conditioning(s, w, t, y, batch, ws.tfreq, ws.th, ws.c, ws.mod, stream));

  // ---- patch embedding + pos_embed -> fp32 residual stream (latte.py:330-331)
  B200_PROF(PROF_OTHER, launch_patch_embed(x, cfg ? batch / 2 : batch, w->patch_w, w->patch_b, w->pos_embed, ws.x, batch, F,
                              s->in_channels, s->input_size, s->patch, D, 0, stream));

  // ---- blocks (latte.py:345-368); rows stay in (b, f, n) order for all of them
  for (int i = 0; i < depth; ++i) {
    const float* m = ws.mod + static_cast<size_t>(i) * 6 * D;  // [shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp]
    const uint16_t* qkv_w = static_cast<const uint16_t*>(w->qkv_w16) + static_cast<size_t>(i) * 3 * D * D;
    const uint16_t* proj_w = static_cast<const uint16_t*>(w->proj_w16) + static_cast<size_t>(i) * D * D;
    const uint16_t* fc1_w = static_cast<const uint16_t*>(w->fc1_w16) + static_cast<size_t>(i) * HID * D;
    const uint16_t* fc2_w = static_cast<const uint16_t*>(w->fc2_w16) + static_cast<size_t>(i) * D * HID;

    B200_PROF(PROF_LN, launch_ln_modulate(ws.x, m + 0 * D, m + 1 * D, mod_bs, rows_per_batch, ws.h, T, D, bf16, stream));
    GemmArgs ga{};
    ga.A = ws.h; ga.W = qkv_w; ga.bias = w->qkv_b + static_cast<size_t>(i) * 3 * D;
    ga.M = T; ga.N = 3 * D; ga.K = D; ga.bf16 = bf16; ga.epilogue = B200_EPI_BIAS; ga.out16 = ws.qkv; ga.w_const = 1;
    B200_PROF(PROF_GEMM, launch_gemm(ga, stream));

    AttnArgs aa{};
    aa.qkv = ws.qkv; aa.out = ws.h; aa.batch = batch; aa.frames = F; aa.tokens = N; aa.heads = H; aa.head_dim = hd;
    aa.bf16 = bf16; aa.temporal = i & 1;
    B200_PROF(PROF_ATTN, launch_attention(aa, stream));

    GemmArgs gp{};
    gp.A = ws.h; gp.W = proj_w; gp.bias = w->proj_b + static_cast<size_t>(i) * D;
    gp.M = T; gp.N = D; gp.K = D; gp.bf16 = bf16; gp.epilogue = B200_EPI_GATE_RESIDUAL; gp.resid = ws.x; gp.w_const = 1;
    gp.gate = m + 2 * D; gp.gate_batch_stride = mod_bs; gp.rows_per_batch = rows_per_batch; gp.sk_flags = ws.sk_flags;
    B200_PROF(PROF_GEMM, launch_gemm(gp, stream));

    B200_PROF(PROF_LN, launch_ln_modulate(ws.x, m + 3 * D, m + 4 * D, mod_bs, rows_per_batch, ws.h, T, D, bf16, stream));
    GemmArgs g1{};
    g1.A = ws.h; g1.W = fc1_w; g1.bias = w->fc1_b + static_cast<size_t>(i) * HID;
    g1.M = T; g1.N = HID; g1.K = D; g1.bf16 = bf16; g1.epilogue = B200_EPI_BIAS_GELU; g1.out16 = ws.g; g1.w_const = 1;
    B200_PROF(PROF_GEMM, launch_gemm(g1, stream));

    GemmArgs g2{};
    g2.A = ws.g; g2.W = fc2_w; g2.bias = w->fc2_b + static_cast<size_t>(i) * D;
    g2.M = T; g2.N = D; g2.K = HID; g2.bf16 = bf16; g2.epilogue = B200_EPI_GATE_RESIDUAL; g2.resid = ws.x; g2.w_const = 1;
    g2.gate = m + 5 * D; g2.gate_batch_stride = mod_bs; g2.rows_per_batch = rows_per_batch; g2.sk_flags = ws.sk_flags;
    if (i == 0) {  // x = x + temp_embed before the first temporal block (latte.py:357-358), folded into block 0's last epilogue
      g2.row_add = w->temp_embed; g2.row_add_div = N; g2.row_add_period = F;
    }
    B200_PROF(PROF_GEMM, launch_gemm(g2, stream));
  }

  // ---- final layer + unpatchify (latte.py:374-376), then guidance (latte.py:394-398)
  const float* mf = ws.mod + static_cast<size_t>(depth) * 6 * D;  // [shift, scale]
  B200_TRY(output_head(ws.x, ws.h, ws.head, mf, mf + D, mod_bs, w->final_w, w->final_w16, w->final_b, out, batch, F, grid, s->patch,
                       s->out_channels, D, bf16, 0, ws.sk_flags, stream));
  if (cfg) {
    const long long per_sample = static_cast<long long>(F) * s->out_channels * s->input_size * s->input_size;
    B200_PROF(PROF_OTHER, launch_cfg_combine(out, batch, per_sample, F, s->out_channels, s->in_channels, s->input_size * s->input_size,
                                cfg_scale, stream));
  }
  return B200_OK;
}


// ====================================================================================================== LatteT2V
struct T2VWorkspace {
  float* x; uint16_t* h; uint16_t* qkv; uint16_t* g;
  uint16_t* text16; uint16_t* cap_h; uint16_t* cap_o; uint16_t* kv_all;
  float* ones; float* tfreq; float* th; float* emb; float* ts; float* mod;
  unsigned long long* sk_flags;
  float* head;
  size_t bytes;
};

int t2v_shape_ok(const B200T2VShape* s, int batch, int text_len) {
  B200_REQUIRE(s != nullptr && batch > 0, B200_ERR_SHAPE, "t2v: bad shape/batch");
  B200_REQUIRE(s->layers > 0 && s->heads > 0 && s->hidden % s->heads == 0, B200_ERR_SHAPE, "t2v: hidden %d / heads %d", s->hidden, s->heads);
  const int hd = s->hidden / s->heads;
  B200_REQUIRE(hd == 64 || hd == 72 || hd == 80, B200_ERR_UNSUPPORTED, "t2v: head_dim %d unsupported", hd);
  B200_REQUIRE(s->hidden % 64 == 0 && s->mlp_hidden % 64 == 0 && s->caption_channels % 64 == 0, B200_ERR_UNSUPPORTED,
               "t2v: hidden, mlp_hidden, caption_channels must be multiples of 64");
  B200_REQUIRE(s->patch == 2 && s->input_size % 2 == 0, B200_ERR_UNSUPPORTED, "t2v: patch size %d not built", s->patch);
  B200_REQUIRE(text_len >= 1 && text_len <= 128, B200_ERR_UNSUPPORTED, "t2v: text length %d (1..128 built)", text_len);
  B200_REQUIRE(s->dtype == B200_FP16 || s->dtype == B200_BF16, B200_ERR_DTYPE, "t2v: dtype %d unknown", s->dtype);
  B200_REQUIRE(s->out_channels * 4 <= 32, B200_ERR_UNSUPPORTED, "t2v: p*p*out_channels > 32");
  const int grid = s->input_size / 2;
  B200_REQUIRE((s->frames * grid * grid) % 128 == 0, B200_ERR_UNSUPPORTED, "t2v: tokens per sample must be a multiple of 128");
  return B200_OK;
}

void t2v_carve(const B200T2VShape* s, int batch, int text_len, void* base, T2VWorkspace* ws) {
  const size_t grid = s->input_size / s->patch;
  const size_t T = static_cast<size_t>(batch) * s->frames * grid * grid;
  const size_t D = s->hidden, R = static_cast<size_t>(batch) * text_len;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<uint8_t*>(base) + off : nullptr;
    off += align_up(bytes, 1024);
    return p;
  };
  ws->x = static_cast<float*>(take(T * D * 4));
  ws->h = static_cast<uint16_t*>(take(T * D * 2));
  ws->qkv = static_cast<uint16_t*>(take(T * 3 * D * 2));
  ws->g = static_cast<uint16_t*>(take(T * static_cast<size_t>(s->mlp_hidden) * 2));
  ws->text16 = static_cast<uint16_t*>(take(R * s->caption_channels * 2));
  ws->cap_h = static_cast<uint16_t*>(take(R * D * 2));
  ws->cap_o = static_cast<uint16_t*>(take(R * D * 2));
  ws->kv_all = static_cast<uint16_t*>(take(R * static_cast<size_t>(s->layers) * 2 * D * 2));
  ws->ones = static_cast<float*>(take(D * 4));
  ws->tfreq = static_cast<float*>(take(static_cast<size_t>(batch) * 256 * 4));
  ws->th = static_cast<float*>(take(static_cast<size_t>(batch) * D * 4));
  ws->emb = static_cast<float*>(take(static_cast<size_t>(batch) * D * 4));
  ws->ts = static_cast<float*>(take(static_cast<size_t>(batch) * 6 * D * 4));
  ws->mod = static_cast<float*>(take(static_cast<size_t>(batch) * (static_cast<size_t>(s->layers) * 2 * 6 * D + 2 * D) * 4));
  ws->sk_flags = static_cast<unsigned long long*>(take(static_cast<size_t>(B200_GEMM_SK_FLAGS) * 8));
  ws->head = static_cast<float*>(take((T + 1) * 32 * 4));
  ws->bytes = off;
}

int t2v_forward(const B200T2VShape* s, const B200T2VWeights* w, const float* x, const int64_t* t, const float* text,
                const float* text_bias, int batch, int text_len, int enable_temporal, float* out, void* workspace,
                size_t workspace_bytes, cudaStream_t stream) {
  B200_TRY(t2v_shape_ok(s, batch, text_len));
  B200_REQUIRE(w && x && t && text && out && workspace, B200_ERR_SHAPE, "t2v: NULL argument");
  B200_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, B200_ERR_ALIGN, "t2v: workspace must be 1024-byte aligned");
  B200_TRY(check_arch());
  T2VWorkspace ws;
  t2v_carve(s, batch, text_len, workspace, &ws);
  B200_REQUIRE(ws.bytes <= workspace_bytes, B200_ERR_WORKSPACE, "t2v: workspace too small: need %zu bytes, got %zu", ws.bytes, workspace_bytes);

  const int D = s->hidden, H = s->heads, hd = D / H, F = s->frames, L = s->layers, HID = s->mlp_hidden;
  const int grid = s->input_size / s->patch, N = grid * grid;
  const int T = batch * F * N, R = batch * text_len;
  const int rows_per_batch = F * N;
  const int bf16 = s->dtype == B200_BF16;
  const long long mod_bs = static_cast<long long>(L) * 2 * 6 * D + 2 * D;

  B200_CHECK_CUDA(cudaMemsetAsync(ws.sk_flags, 0, static_cast<size_t>(B200_GEMM_SK_FLAGS) * 8, stream));
  // ---- conditioning (latte_t2v.py:782-784): emb = TimestepEmbedding(sincos(t)); ts = Linear(SiLU(emb)); tables + ts
  B200_PROF(PROF_OTHER, launch_timestep_freq(reinterpret_cast<const long long*>(t), ws.tfreq, batch, stream));
  B200_PROF(PROF_OTHER, launch_gemv(w->t_w0, 32, 0, w->t_b0, ws.tfreq, ws.th, batch, D, 256, 0, 1, nullptr, nullptr, 0, stream));
  B200_PROF(PROF_OTHER, launch_gemv(w->t_w2, 32, 0, w->t_b2, ws.th, ws.emb, batch, D, D, 0, 0, nullptr, nullptr, 0, stream));
  B200_PROF(PROF_OTHER, launch_gemv(w->ada_w16, 16, bf16, w->ada_b, ws.emb, ws.ts, batch, 6 * D, D, 1, 0, nullptr, nullptr, 0, stream));
  B200_PROF(PROF_OTHER, launch_t2v_mod(w->tables, ws.ts, w->final_table, ws.emb, ws.mod, batch, 2 * L, D, stream));
  B200_PROF(PROF_OTHER, launch_fill(ws.ones, 1.0f, D, stream));

  // ---- text: caption projection once per sample (latte_t2v.py:789), then K/V of EVERY layer's cross-attention in one GEMM
  B200_PROF(PROF_OTHER, launch_cast16(text, ws.text16, static_cast<long long>(R) * s->caption_channels, bf16, stream));
  {
    GemmArgs a{};
    a.A = ws.text16; a.W = w->cap_w1_16; a.bias = w->cap_b1; a.M = R; a.N = D; a.K = s->caption_channels; a.bf16 = bf16;
    a.epilogue = B200_EPI_BIAS_GELU; a.out16 = ws.cap_h;
    B200_PROF(PROF_GEMM, launch_gemm(a, stream));
    GemmArgs b{};
    b.A = ws.cap_h; b.W = w->cap_w2_16; b.bias = w->cap_b2; b.M = R; b.N = D; b.K = D; b.bf16 = bf16;
    b.epilogue = B200_EPI_BIAS; b.out16 = ws.cap_o;
    B200_PROF(PROF_GEMM, launch_gemm(b, stream));
    GemmArgs c{};
    c.A = ws.cap_o; c.W = w->c_kv_w16; c.bias = w->c_kv_b; c.M = R; c.N = L * 2 * D; c.K = D; c.bf16 = bf16;
    c.epilogue = B200_EPI_BIAS; c.out16 = ws.kv_all;
    B200_PROF(PROF_GEMM, launch_gemm(c, stream));
  }

  // ---- patch embedding + pos_embed (latte_t2v.py:731,773); x arrives as (b c f h w)
  B200_PROF(PROF_OTHER, launch_patch_embed(x, batch, w->patch_w, w->patch_b, w->pos_embed, ws.x, batch, F, s->in_channels,
                                           s->input_size, s->patch, D, 1, stream));

  auto linear16 = [&](const void* A, const void* W, const float* bias, int M, int Nn, int K, int epi, void* o16) {
    GemmArgs a{};
    a.A = A; a.W = W; a.bias = bias; a.M = M; a.N = Nn; a.K = K; a.bf16 = bf16; a.epilogue = epi; a.out16 = o16; a.w_const = 1;
    return launch_gemm(a, stream);
  };
  auto linear_resid = [&](const void* A, const void* W, const float* bias, int K, const float* gate, long long gate_bs,
                          const float* row_add) {
    GemmArgs a{};
    a.A = A; a.W = W; a.bias = bias; a.M = T; a.N = D; a.K = K; a.bf16 = bf16; a.epilogue = B200_EPI_GATE_RESIDUAL; a.w_const = 1;
    a.resid = ws.x; a.gate = gate; a.gate_batch_stride = gate_bs; a.rows_per_batch = rows_per_batch; a.sk_flags = ws.sk_flags;
    if (row_add) { a.row_add = row_add; a.row_add_div = N; a.row_add_period = F; }
    return launch_gemm(a, stream);
  };
  const size_t DD = static_cast<size_t>(D) * D;

  for (int l = 0; l < L; ++l) {
    // ------------------------------------------------ spatial block (diffusers BasicTransformerBlock; latte_t2v.py:862-870)
    const float* m = ws.mod + static_cast<size_t>(2 * l) * 6 * D;
    B200_PROF(PROF_LN, launch_ln_modulate(ws.x, m + 0 * D, m + 1 * D, mod_bs, rows_per_batch, ws.h, T, D, bf16, stream));
    B200_PROF(PROF_GEMM, linear16(ws.h, static_cast<const uint16_t*>(w->s_qkv_w16) + l * 3 * DD, w->s_qkv_b + static_cast<size_t>(l) * 3 * D,
                                  T, 3 * D, D, B200_EPI_BIAS, ws.qkv));
    AttnArgs aa{};
    aa.qkv = ws.qkv; aa.out = ws.h; aa.batch = batch; aa.frames = F; aa.tokens = N; aa.heads = H; aa.head_dim = hd; aa.bf16 = bf16;
    aa.temporal = 0;
    B200_PROF(PROF_ATTN, launch_attention(aa, stream));
    B200_PROF(PROF_GEMM, linear_resid(ws.h, static_cast<const uint16_t*>(w->s_out_w16) + l * DD, w->s_out_b + static_cast<size_t>(l) * D, D,
                                      m + 2 * D, mod_bs, nullptr));
    // cross-attention on the UN-normalised stream (no norm2 before attn2 in ada_norm_single mode), residual without gate
    B200_PROF(PROF_OTHER, launch_cast16(ws.x, ws.h, static_cast<long long>(T) * D, bf16, stream));
    B200_PROF(PROF_GEMM, linear16(ws.h, static_cast<const uint16_t*>(w->c_q_w16) + l * DD, w->c_q_b + static_cast<size_t>(l) * D, T, D, D,
                                  B200_EPI_BIAS, ws.qkv));
    CrossAttnArgs ca{};
    ca.q = ws.qkv; ca.kv = ws.kv_all + static_cast<size_t>(l) * 2 * D; ca.out = ws.h; ca.batch = batch; ca.q_rows_per_batch = rows_per_batch;
    ca.kv_len = text_len; ca.q_row_stride = D; ca.kv_row_stride = L * 2 * D; ca.heads = H; ca.head_dim = hd; ca.bf16 = bf16;
    ca.key_bias = text_bias;     // padded prompts: (1 - mask) * -10000 per text token (latte_t2v.py:766-771), or NULL
    B200_PROF(PROF_ATTN, launch_cross_attention(ca, stream));
    B200_PROF(PROF_GEMM, linear_resid(ws.h, static_cast<const uint16_t*>(w->c_out_w16) + l * DD, w->c_out_b + static_cast<size_t>(l) * D, D,
                                      ws.ones, 0, nullptr));
    B200_PROF(PROF_LN, launch_ln_modulate(ws.x, m + 3 * D, m + 4 * D, mod_bs, rows_per_batch, ws.h, T, D, bf16, stream));
    B200_PROF(PROF_GEMM, linear16(ws.h, static_cast<const uint16_t*>(w->s_fc1_w16) + static_cast<size_t>(l) * HID * D,
                                  w->s_fc1_b + static_cast<size_t>(l) * HID, T, HID, D, B200_EPI_BIAS_GELU, ws.g));
    // + temp_pos_embed before the first temporal block (latte_t2v.py:894-895), folded into this epilogue
    B200_PROF(PROF_GEMM, linear_resid(ws.g, static_cast<const uint16_t*>(w->s_fc2_w16) + static_cast<size_t>(l) * D * HID,
                                      w->s_fc2_b + static_cast<size_t>(l) * D, HID, m + 5 * D, mod_bs,
                                      (l == 0 && enable_temporal && F > 1) ? w->temp_embed : nullptr));
    if (!enable_temporal) continue;
    // ------------------------------------------------ temporal block (BasicTransformerBlock_, latte_t2v.py:897-905)
    const float* mt = ws.mod + static_cast<size_t>(2 * l + 1) * 6 * D;
    B200_PROF(PROF_LN, launch_ln_modulate(ws.x, mt + 0 * D, mt + 1 * D, mod_bs, rows_per_batch, ws.h, T, D, bf16, stream));
    B200_PROF(PROF_GEMM, linear16(ws.h, static_cast<const uint16_t*>(w->t_qkv_w16) + l * 3 * DD, w->t_qkv_b + static_cast<size_t>(l) * 3 * D,
                                  T, 3 * D, D, B200_EPI_BIAS, ws.qkv));
    aa.temporal = 1;
    B200_PROF(PROF_ATTN, launch_attention(aa, stream));
    B200_PROF(PROF_GEMM, linear_resid(ws.h, static_cast<const uint16_t*>(w->t_out_w16) + l * DD, w->t_out_b + static_cast<size_t>(l) * D, D,
                                      mt + 2 * D, mod_bs, nullptr));
    B200_PROF(PROF_LN, launch_ln_modulate(ws.x, mt + 3 * D, mt + 4 * D, mod_bs, rows_per_batch, ws.h, T, D, bf16, stream));
    B200_PROF(PROF_GEMM, linear16(ws.h, static_cast<const uint16_t*>(w->t_fc1_w16) + static_cast<size_t>(l) * HID * D,
                                  w->t_fc1_b + static_cast<size_t>(l) * HID, T, HID, D, B200_EPI_BIAS_GELU, ws.g));
    B200_PROF(PROF_GEMM, linear_resid(ws.g, static_cast<const uint16_t*>(w->t_fc2_w16) + static_cast<size_t>(l) * D * HID,
                                      w->t_fc2_b + static_cast<size_t>(l) * D, HID, mt + 5 * D, mod_bs, nullptr));
  }

  // ---- output head (latte_t2v.py:918-936): table + embedded_timestep -> shift, scale; LN; modulate; proj_out; unpatchify to (b c f h w)
  const float* mf = ws.mod + static_cast<size_t>(2 * L) * 6 * D;
  B200_TRY(output_head(ws.x, ws.h, ws.head, mf, mf + D, mod_bs, w->final_w, w->final_w16, w->final_b, out, batch, F, grid, s->patch,
                       s->out_channels, D, bf16, 1, ws.sk_flags, stream));
  return B200_OK;
}

int output_head(const float* x, uint16_t* h, float* head, const float* shift, const float* scale, long long mod_bs,
                const float* w32, const void* w16, const float* bias, float* out, int batch, int F, int grid, int patch,
                int out_ch, int D, int bf16, int channels_first, unsigned long long* sk_flags, cudaStream_t stream) {
  const int n_out = patch * patch * out_ch;
  const int T = batch * F * grid * grid;
  if (w16 == nullptr || n_out != 32) {
    B200_PROF(PROF_OTHER, launch_final_layer(x, shift, scale, mod_bs, w32, bias, out, batch, F, grid, patch, out_ch, D, channels_first, stream));
    return B200_OK;
  }
  float* ones = head + static_cast<size_t>(T) * 32;
  B200_CHECK_CUDA(cudaMemsetAsync(head, 0, static_cast<size_t>(T) * 32 * 4, stream));
  B200_PROF(PROF_OTHER, launch_fill(ones, 1.0f, 32, stream));
  B200_PROF(PROF_LN, launch_ln_modulate(x, shift, scale, mod_bs, F * grid * grid, h, T, D, bf16, stream));
  GemmArgs g{};
  g.A = h; g.W = w16; g.bias = bias; g.M = T; g.N = n_out; g.K = D; g.bf16 = bf16; g.epilogue = B200_EPI_GATE_RESIDUAL; g.w_const = 1;
  g.resid = head; g.gate = ones; g.gate_batch_stride = 0; g.rows_per_batch = T; g.sk_flags = sk_flags;
  B200_PROF(PROF_GEMM, launch_gemm(g, stream));
  B200_PROF(PROF_OTHER, launch_unpatchify(head, out, batch, F, grid, patch, out_ch, channels_first, stream));
  return B200_OK;
}

// ====================================================================================================== T5 encoder
struct T5Workspace {
  float* x; uint16_t* h; uint16_t* qkv; uint16_t* att; uint16_t* g0; uint16_t* g; float* ones; unsigned long long* sk_flags;
  size_t bytes;
};

int t5_shape_ok(const B200T5Shape* s, int batch) {
  B200_REQUIRE(s != nullptr && batch > 0, B200_ERR_SHAPE, "t5: bad shape/batch");
  B200_REQUIRE(s->layers > 0 && s->heads > 0 && s->vocab > 0, B200_ERR_SHAPE, "t5: layers/heads/vocab must be positive");
  B200_REQUIRE(s->d_model % 64 == 0 && s->d_ff % 64 == 0, B200_ERR_UNSUPPORTED, "t5: d_model %d and d_ff %d must be multiples of 64", s->d_model, s->d_ff);
  B200_REQUIRE(s->dtype == B200_FP16 || s->dtype == B200_BF16, B200_ERR_DTYPE, "t5: dtype %d unknown", s->dtype);
  return B200_OK;
}

void t5_carve(const B200T5Shape* s, int batch, void* base, T5Workspace* ws) {
  const size_t R = static_cast<size_t>(batch) * 128, D = s->d_model, I = static_cast<size_t>(s->heads) * 64, FF = s->d_ff;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<uint8_t*>(base) + off : nullptr;
    off += align_up(bytes, 1024);
    return p;
  };
  ws->x = static_cast<float*>(take(R * D * 4));
  ws->h = static_cast<uint16_t*>(take(R * D * 2));
  ws->qkv = static_cast<uint16_t*>(take(R * 3 * I * 2));
  ws->att = static_cast<uint16_t*>(take(R * I * 2));
  ws->g0 = static_cast<uint16_t*>(take(R * FF * 2));
  ws->g = static_cast<uint16_t*>(take(R * FF * 2));
  ws->ones = static_cast<float*>(take(D * 4));
  ws->sk_flags = static_cast<unsigned long long*>(take(static_cast<size_t>(B200_GEMM_SK_FLAGS) * 8));
  ws->bytes = off;
}

// T5Stack.forward (encoder): embed -> [T5LayerSelfAttention, T5LayerFF] x layers -> final_layer_norm (transformers
// modeling_t5.py; reached from sample/pipeline_latte.py:214).  Sequences are padded to 128 rows, so one attention tile is one
// (sample, head); masked / padding keys carry a large negative key_bias.
int t5_encode(const B200T5Shape* s, const B200T5Weights* w, const int64_t* ids, const float* key_bias, const float* pos_bias,
              int batch, float* out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  B200_TRY(t5_shape_ok(s, batch));
  B200_REQUIRE(w && ids && pos_bias && out && workspace, B200_ERR_SHAPE, "t5: NULL argument");
  B200_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, B200_ERR_ALIGN, "t5: workspace must be 1024-byte aligned");
  B200_TRY(check_arch());
  T5Workspace ws;
  t5_carve(s, batch, workspace, &ws);
  B200_REQUIRE(ws.bytes <= workspace_bytes, B200_ERR_WORKSPACE, "t5: workspace too small: need %zu bytes, got %zu", ws.bytes, workspace_bytes);
  const int R = batch * 128, D = s->d_model, H = s->heads, I = H * 64, FF = s->d_ff;
  const int bf16 = s->dtype == B200_BF16;
  B200_CHECK_CUDA(cudaMemsetAsync(ws.sk_flags, 0, static_cast<size_t>(B200_GEMM_SK_FLAGS) * 8, stream));
  B200_TRY(launch_fill(ws.ones, 1.0f, D, stream));
  B200_TRY(launch_embed(reinterpret_cast<const long long*>(ids), w->embed16, ws.x, R, D, s->vocab, bf16, stream));
  auto linear16 = [&](const void* A, const void* W, int M, int N, int K, int epi, void* o16, const void* other) {
    GemmArgs a{};
    a.A = A; a.W = W; a.bias = nullptr; a.M = M; a.N = N; a.K = K; a.bf16 = bf16; a.epilogue = epi; a.out16 = o16; a.add16 = other; a.w_const = 1;
    return launch_gemm(a, stream);
  };
  auto linear_resid = [&](const void* A, const void* W, int K) {
    GemmArgs a{};
    a.A = A; a.W = W; a.bias = nullptr; a.M = R; a.N = D; a.K = K; a.bf16 = bf16; a.epilogue = B200_EPI_GATE_RESIDUAL; a.w_const = 1;
    a.resid = ws.x; a.gate = ws.ones; a.gate_batch_stride = 0; a.rows_per_batch = R; a.sk_flags = ws.sk_flags;
    return launch_gemm(a, stream);
  };
  for (int l = 0; l < s->layers; ++l) {
    const uint16_t* qkv_w = static_cast<const uint16_t*>(w->qkv_w16) + static_cast<size_t>(l) * 3 * I * D;
    const uint16_t* o_w = static_cast<const uint16_t*>(w->o_w16) + static_cast<size_t>(l) * D * I;
    const uint16_t* wi0 = static_cast<const uint16_t*>(w->wi0_w16) + static_cast<size_t>(l) * FF * D;
    const uint16_t* wi1 = static_cast<const uint16_t*>(w->wi1_w16) + static_cast<size_t>(l) * FF * D;
    const uint16_t* wo = static_cast<const uint16_t*>(w->wo_w16) + static_cast<size_t>(l) * D * FF;
    // ---- T5LayerSelfAttention: x += o(attention(q, k, v of T5LayerNorm(x)) with position bias + mask, NO 1/sqrt(d) scale)
    B200_TRY(launch_rms_norm(ws.x, w->ln0_w + static_cast<size_t>(l) * D, ws.h, nullptr, R, D, s->eps, bf16, stream));
    B200_TRY(linear16(ws.h, qkv_w, R, 3 * I, D, B200_EPI_BIAS, ws.qkv, nullptr));
    CrossAttnArgs ca{};
    ca.q = ws.qkv; ca.kv = ws.qkv + I; ca.out = ws.att; ca.batch = batch; ca.q_rows_per_batch = 128; ca.kv_len = 128;
    ca.kv_batch_rows = 128; ca.q_row_stride = 3 * I; ca.kv_row_stride = 3 * I; ca.heads = H; ca.head_dim = 64; ca.bf16 = bf16;
    ca.key_bias = key_bias; ca.pos_bias = pos_bias; ca.scale = 1.0f;
    B200_TRY(launch_cross_attention(ca, stream));
    B200_TRY(linear_resid(ws.att, o_w, I));
    // ---- T5LayerFF (gated-gelu): x += wo(gelu_new(wi_0 h) * wi_1 h), h = T5LayerNorm(x)
    B200_TRY(launch_rms_norm(ws.x, w->ln1_w + static_cast<size_t>(l) * D, ws.h, nullptr, R, D, s->eps, bf16, stream));
    B200_TRY(linear16(ws.h, wi0, R, FF, D, B200_EPI_BIAS_GELU, ws.g0, nullptr));
    B200_TRY(linear16(ws.h, wi1, R, FF, D, B200_EPI_BIAS_MUL16, ws.g, ws.g0));
    B200_TRY(linear_resid(ws.g, wo, FF));
  }
  B200_TRY(launch_rms_norm(ws.x, w->final_w, nullptr, out, R, D, s->eps, bf16, stream));
  return B200_OK;
}

}  // namespace
}  // namespace b200

extern "C" {

B200_API int b200_frames_to_uint8(const void* video, int dtype, int n, int c, int h, int w, int mode, uint8_t* out, void* stream) {
  B200_TRY(b200::check_arch());
  return b200::launch_frames_to_uint8(video, dtype, n, c, h, w, mode, out, static_cast<cudaStream_t>(stream));
}

B200_API size_t b200_t5_workspace_bytes(const B200T5Shape* shape, int batch) {
  if (b200::t5_shape_ok(shape, batch) != B200_OK) return 0;
  b200::T5Workspace ws;
  b200::t5_carve(shape, batch, nullptr, &ws);
  return ws.bytes;
}

B200_API int b200_t5_encode(const B200T5Shape* shape, const B200T5Weights* w, const int64_t* ids, const float* key_bias,
                            const float* pos_bias, int batch, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  return b200::t5_encode(shape, w, ids, key_bias, pos_bias, batch, out, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

B200_API int b200_sampler_step(const B200SamplerTables* tables, int method, int clip_denoised, const int64_t* t,
                               const float* x, const void* model_out, int model_out_dtype, const float* noise, int batch,
                               int frames, int channels, int hw, float* x_prev, float* pred_xstart, float* mean,
                               float* log_variance, void* stream) {
  return b200::launch_sampler_step(tables, method, clip_denoised, reinterpret_cast<const long long*>(t), x, model_out,
                                   model_out_dtype, noise, batch, frames, channels, hw, x_prev, pred_xstart, mean,
                                   log_variance, static_cast<cudaStream_t>(stream));
}

B200_API int b200_training_loss(const B200SamplerTables* tables, const int64_t* t, const float* x0, const float* xt, const float* noise,
                                const float* model_out, int batch, int frames, int channels, int hw, float* sums, float* dmo, void* stream) {
  return b200::launch_training_loss(tables, reinterpret_cast<const long long*>(t), x0, xt, noise, model_out, batch, frames, channels, hw,
                                    sums, dmo, static_cast<cudaStream_t>(stream));
}

B200_API int b200_wgrad_schedule(int rows, int n_out, int n_in, int num_sms, int* block_n_out, int* pairs_out, int* streamk_out,
                                 int32_t* segments, int max_segments) {
  return b200::gemm_schedule(n_out, n_in, rows, B200_EPI_GATE_RESIDUAL, 0, num_sms, block_n_out, pairs_out, streamk_out, segments, max_segments, 1);
}

B200_API int b200_gemm_schedule(int M, int N, int K, int epilogue, int block_n, int num_sms, int* block_n_out, int* pairs_out,
                                int* streamk_out, int32_t* segments, int max_segments) {
  return b200::gemm_schedule(M, N, K, epilogue, block_n, num_sms, block_n_out, pairs_out, streamk_out, segments, max_segments);
}

B200_API void b200_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(b200::g_prof.mu);
  b200::g_prof.on = on != 0;
}

B200_API int b200_profile_collect(double* ms_per_class, int* launches_per_class, int n_classes) {
  // synchronises on the recorded events, sums elapsed time per class, then clears the records
  std::lock_guard<std::mutex> lk(b200::g_prof.mu);
  for (int i = 0; i < n_classes; ++i) { ms_per_class[i] = 0.0; launches_per_class[i] = 0; }
  for (auto& r : b200::g_prof.recs) {
    B200_CHECK_CUDA(cudaEventSynchronize(r.e1));
    float ms = 0.f;
    B200_CHECK_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
    if (r.cls < n_classes) { ms_per_class[r.cls] += ms; launches_per_class[r.cls] += 1; }
    b200::g_prof.pool.push_back(r.e0);
    b200::g_prof.pool.push_back(r.e1);
  }
  b200::g_prof.recs.clear();
  return B200_OK;
}

B200_API int b200_set_attention_impl(int impl) { return b200::set_attention_impl(impl); }

B200_API const char* b200_last_error(void) { return b200::get_error(); }
B200_API int b200_abi_version(void) { return B200_ABI_VERSION; }

B200_API size_t b200_latte_workspace_bytes(const B200LatteShape* shape, int batch) {
  if (b200::shape_ok(shape, batch) != B200_OK) return 0;
  b200::Workspace ws;
  b200::carve(shape, batch, nullptr, &ws);
  return ws.bytes;
}

B200_API int b200_latte_forward(const B200LatteShape* shape, const B200LatteWeights* w, const float* x, const int64_t* t,
                       const int64_t* y, int batch, int use_cfg, float cfg_scale, float* out, void* workspace,
                       size_t workspace_bytes, void* stream) {
  return b200::forward(shape, w, x, t, y, nullptr, batch, use_cfg, cfg_scale, out, workspace, workspace_bytes,
                       static_cast<cudaStream_t>(stream));
}

B200_API size_t b200_latte_conditioning_bytes(const B200LatteShape* shape, int n) {
  if (b200::shape_ok(shape, n) != B200_OK) return 0;
  return static_cast<size_t>(n) * (static_cast<size_t>(shape->depth) * 6 * shape->hidden + 2 * shape->hidden) * sizeof(float);
}

B200_API size_t b200_latte_conditioning_workspace_bytes(const B200LatteShape* shape, int n) {
  if (b200::shape_ok(shape, n) != B200_OK) return 0;
  return b200::conditioning_scratch_bytes(shape, n);
}

B200_API int b200_latte_conditioning(const B200LatteShape* shape, const B200LatteWeights* w, const int64_t* t, const int64_t* y,
                                     int n, float* mod_out, void* workspace, size_t workspace_bytes, void* stream) {
  B200_TRY(b200::shape_ok(shape, n));
  B200_REQUIRE(w && t && mod_out && workspace, B200_ERR_SHAPE, "NULL argument");
  B200_REQUIRE((shape->num_embed > 0) == (y != nullptr && w->y_table != nullptr), B200_ERR_SHAPE,
               "labels y and y_table must be given iff num_embed > 0 (extras == 2)");
  B200_REQUIRE(((reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(mod_out)) & 1023) == 0, B200_ERR_ALIGN,
               "workspace and mod_out must be 1024-byte aligned");
  B200_REQUIRE(workspace_bytes >= b200::conditioning_scratch_bytes(shape, n), B200_ERR_WORKSPACE, "conditioning workspace too small");
  B200_TRY(b200::check_arch());
  uint8_t* base = static_cast<uint8_t*>(workspace);
  float* tfreq = reinterpret_cast<float*>(base);
  float* th = reinterpret_cast<float*>(base + b200::align_up(static_cast<size_t>(n) * 256 * 4, 1024));
  float* c = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(th) + b200::align_up(static_cast<size_t>(n) * shape->hidden * 4, 1024));
  return b200::conditioning(shape, w, t, y, n, tfreq, th, c, mod_out, static_cast<cudaStream_t>(stream));
}

B200_API int b200_latte_forward_conditioned(const B200LatteShape* shape, const B200LatteWeights* w, const float* x,
                                            const float* mod, int batch, int use_cfg, float cfg_scale, float* out,
                                            void* workspace, size_t workspace_bytes, void* stream) {
  B200_REQUIRE(mod != nullptr, B200_ERR_SHAPE, "conditioning rows are NULL");
  return b200::forward(shape, w, x, nullptr, nullptr, mod, batch, use_cfg, cfg_scale, out, workspace, workspace_bytes,
                       static_cast<cudaStream_t>(stream));
}

B200_API int b200_linear(const void* A, const void* W, const float* bias, int M, int N, int K, int dtype, int epilogue,
                void* out16, float* resid, const float* gate, int64_t gate_batch_stride, int rows_per_batch,
                int block_n, void* sk_flags, void* stream) {
  B200_REQUIRE(dtype == B200_FP16 || dtype == B200_BF16, B200_ERR_DTYPE, "dtype %d unknown", dtype);
  B200_REQUIRE((reinterpret_cast<uintptr_t>(sk_flags) & 7) == 0, B200_ERR_ALIGN, "sk_flags must be 8-byte aligned");
  b200::GemmArgs a{};
  a.sk_flags = static_cast<unsigned long long*>(sk_flags);
  a.A = A; a.W = W; a.bias = bias; a.M = M; a.N = N; a.K = K; a.bf16 = dtype == B200_BF16; a.epilogue = epilogue;
  a.out16 = out16; a.resid = resid; a.gate = gate; a.gate_batch_stride = gate_batch_stride;
  a.rows_per_batch = rows_per_batch; a.block_n = block_n;
  return b200::launch_gemm(a, static_cast<cudaStream_t>(stream));
}

B200_API int b200_attention(const void* qkv, void* out, int batch, int frames, int tokens, int heads, int head_dim, int dtype,
                   int temporal, void* stream) {
  B200_REQUIRE(dtype == B200_FP16 || dtype == B200_BF16, B200_ERR_DTYPE, "dtype %d unknown", dtype);
  b200::AttnArgs a{};
  a.qkv = qkv; a.out = out; a.batch = batch; a.frames = frames; a.tokens = tokens; a.heads = heads;
  a.head_dim = head_dim; a.bf16 = dtype == B200_BF16; a.temporal = temporal;
  return b200::launch_attention(a, static_cast<cudaStream_t>(stream));
}

B200_API size_t b200_t2v_workspace_bytes(const B200T2VShape* shape, int batch, int text_len) {
  if (b200::t2v_shape_ok(shape, batch, text_len) != B200_OK) return 0;
  b200::T2VWorkspace ws;
  b200::t2v_carve(shape, batch, text_len, nullptr, &ws);
  return ws.bytes;
}

B200_API int b200_t2v_forward(const B200T2VShape* shape, const B200T2VWeights* w, const float* x, const int64_t* t,
                              const float* text, const float* text_bias, int batch, int text_len, int enable_temporal,
                              float* out, void* workspace, size_t workspace_bytes, void* stream) {
  B200_REQUIRE(!text_bias || (reinterpret_cast<uintptr_t>(text_bias) & 15) == 0, B200_ERR_ALIGN, "t2v: text_bias must be 16-byte aligned");
  return b200::t2v_forward(shape, w, x, t, text, text_bias, batch, text_len, enable_temporal, out, workspace, workspace_bytes,
                           static_cast<cudaStream_t>(stream));
}

B200_API int b200_cross_attention(const void* q, const void* kv, const float* key_bias, void* out, int batch, int q_rows_per_batch,
                                  int kv_len, int q_row_stride, int kv_row_stride, int heads, int head_dim, int dtype, void* stream) {
  B200_REQUIRE(dtype == B200_FP16 || dtype == B200_BF16, B200_ERR_DTYPE, "dtype %d unknown", dtype);
  b200::CrossAttnArgs a{};
  a.q = q; a.kv = kv; a.out = out; a.batch = batch; a.q_rows_per_batch = q_rows_per_batch; a.kv_len = kv_len;
  a.key_bias = key_bias;
  a.q_row_stride = q_row_stride; a.kv_row_stride = kv_row_stride; a.heads = heads; a.head_dim = head_dim;
  a.bf16 = dtype == B200_BF16;
  return b200::launch_cross_attention(a, static_cast<cudaStream_t>(stream));
}

B200_API int b200_ln_modulate(const float* x, const float* shift, const float* scale, int64_t mod_batch_stride,
                     int rows_per_batch, void* out16, int rows, int dim, int dtype, void* stream) {
  B200_REQUIRE(dtype == B200_FP16 || dtype == B200_BF16, B200_ERR_DTYPE, "dtype %d unknown", dtype);
  return b200::launch_ln_modulate(x, shift, scale, mod_batch_stride, rows_per_batch, out16, rows, dim,
                                  dtype == B200_BF16, static_cast<cudaStream_t>(stream));
}

// ---- training-step passes (train.cu) ----
#define B200_DT(dtype) B200_REQUIRE(dtype == B200_FP16 || dtype == B200_BF16, B200_ERR_DTYPE, "dtype %d unknown", dtype)
B200_API int b200_wgrad(const void* dy16, const void* x16, const float* col_scale, float* dW, int rows, int n_out, int n_in, int dtype,
                        void* sk_flags, void* stream) {
  B200_DT(dtype);
  B200_REQUIRE((reinterpret_cast<uintptr_t>(sk_flags) & 7) == 0, B200_ERR_ALIGN, "sk_flags must be 8-byte aligned");
  B200_REQUIRE(col_scale != nullptr, B200_ERR_SHAPE, "wgrad: col_scale (n_in floats, 1.0 for a plain gradient) is required");
  b200::GemmArgs a{};
  a.A = dy16; a.W = x16; a.M = n_out; a.N = n_in; a.K = rows; a.bf16 = dtype == B200_BF16; a.epilogue = B200_EPI_GATE_RESIDUAL;
  a.resid = dW; a.gate = col_scale; a.gate_batch_stride = 0; a.rows_per_batch = n_out; a.mn_major = 3;
  a.sk_flags = static_cast<unsigned long long*>(sk_flags);
  return b200::launch_gemm(a, static_cast<cudaStream_t>(stream));
}
B200_API int b200_dgrad(const void* dy16, const void* w16, const void* gelu_u16, void* dx16, int rows, int n_out, int n_in, int dtype,
                        void* stream) {
  B200_DT(dtype);
  b200::GemmArgs a{};
  a.A = dy16; a.W = w16; a.M = rows; a.N = n_in; a.K = n_out; a.bf16 = dtype == B200_BF16;
  a.epilogue = gelu_u16 ? B200_EPI_MUL_GELUGRAD16 : B200_EPI_BIAS;
  a.add16 = gelu_u16; a.out16 = dx16; a.mn_major = 2;
  return b200::launch_gemm(a, static_cast<cudaStream_t>(stream));
}
B200_API int b200_linear_gelu_both(const void* A, const void* W, const float* bias, int M, int N, int K, int dtype, void* u16, void* a16,
                                   void* stream) {
  B200_DT(dtype);
  b200::GemmArgs a{};
  a.A = A; a.W = W; a.bias = bias; a.M = M; a.N = N; a.K = K; a.bf16 = dtype == B200_BF16; a.epilogue = B200_EPI_BIAS_GELU_BOTH;
  a.out16 = u16; a.out16b = a16; a.w_const = 1;
  return b200::launch_gemm(a, static_cast<cudaStream_t>(stream));
}
B200_API int b200_transpose16(const void* in16, void* out16, int rows, int cols, void* stream) {
  return b200::launch_transpose16(in16, out16, rows, cols, static_cast<cudaStream_t>(stream));
}
B200_API int b200_cast_transpose(const float* in, void* out16, void* out16_t, int rows, int cols, int dtype, void* stream) {
  B200_DT(dtype);
  return b200::launch_cast_transpose(in, out16, out16_t, rows, cols, dtype == B200_BF16, static_cast<cudaStream_t>(stream));
}
B200_API int b200_multi_cast(const void* table, int n_entries, int64_t total_chunks, int dtype, void* stream) {
  B200_DT(dtype);
  return b200::launch_multi_cast(table, n_entries, total_chunks, dtype == B200_BF16, static_cast<cudaStream_t>(stream));
}
B200_API int b200_multi_tensor(const void* table, int n_entries, int64_t total_chunks, int op, float a, float b, const float* scalar,
                               double* accum, void* stream) {
  return b200::launch_multi_tensor(table, n_entries, total_chunks, op, a, b, scalar, accum, static_cast<cudaStream_t>(stream));
}
B200_API int b200_cast16(const float* in, void* out16, int64_t n, int dtype, void* stream) {
  B200_DT(dtype);
  B200_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out16) & 7) == 0, B200_ERR_ALIGN, "cast16: misaligned pointer");
  return b200::launch_cast16(in, out16, n, dtype == B200_BF16, static_cast<cudaStream_t>(stream));
}
B200_API int b200_gate_residual(const float* x, const void* m16, const float* gate, int64_t gate_batch_stride, int rows_per_batch,
                                const float* row_add, int tokens, int frames, float* out, int rows, int dim, int dtype, void* stream) {
  B200_DT(dtype);
  return b200::launch_gate_residual(x, m16, gate, gate_batch_stride, rows_per_batch, row_add, tokens, frames, out, rows, dim,
                                    dtype == B200_BF16, static_cast<cudaStream_t>(stream));
}
B200_API int b200_gate_residual_ln(const float* x, const void* m16, const float* gate, int64_t gate_batch_stride, const float* shift,
                                   const float* scale, int64_t mod_batch_stride, int rows_per_batch, const float* row_add, int tokens, int frames,
                                   float* x_out, void* h16, int rows, int dim, int dtype, void* stream) {
  B200_DT(dtype);
  return b200::launch_gate_residual_ln(x, m16, gate, gate_batch_stride, shift, scale, mod_batch_stride, rows_per_batch, row_add, tokens, frames,
                                       x_out, h16, rows, dim, dtype == B200_BF16, static_cast<cudaStream_t>(stream));
}
B200_API int b200_gelu(const void* u16, void* a16, int64_t n, int dtype, void* stream) {
  B200_DT(dtype);
  return b200::launch_gelu_fwd(u16, a16, n, dtype == B200_BF16, static_cast<cudaStream_t>(stream));
}
B200_API int b200_gelu_bwd(const void* da16, const void* u16, void* du16, float* dbias, int rows, int dim, int dtype, void* stream) {
  B200_DT(dtype);
  return b200::launch_gelu_bwd(da16, u16, du16, dbias, rows, dim, dtype == B200_BF16, static_cast<cudaStream_t>(stream));
}
B200_API int b200_gate_bwd(const float* dx, const void* m16, const float* gate, int64_t gate_batch_stride, int rows_per_batch,
                           void* dm16, float* dgate, int64_t dgate_batch_stride, float* dbias, int rows, int dim, int dtype,
                           void* stream) {
  B200_DT(dtype);
  return b200::launch_gate_bwd(dx, m16, gate, gate_batch_stride, rows_per_batch, dm16, dgate, dgate_batch_stride, dbias, rows, dim,
                               dtype == B200_BF16, static_cast<cudaStream_t>(stream));
}
B200_API int b200_colsum(const void* a, int a_dtype, float* out, int rows, int dim, void* stream) {
  return b200::launch_colsum(a, a_dtype, out, rows, dim, static_cast<cudaStream_t>(stream));
}
B200_API int b200_ln_modulate_bwd(const void* dh16, const float* x, const float* scale, int64_t mod_batch_stride, int rows_per_batch,
                                  float* dx, float* dshift, float* dscale, int64_t dmod_batch_stride, int rows, int dim, int dtype,
                                  void* stream) {
  B200_DT(dtype);
  return b200::launch_ln_modulate_bwd(dh16, x, scale, mod_batch_stride, rows_per_batch, dx, dshift, dscale, dmod_batch_stride, rows,
                                      dim, dtype == B200_BF16, static_cast<cudaStream_t>(stream));
}
B200_API int b200_attention_bwd(const void* qkv16, const void* o16, const void* do16, void* dqkv16, float* stats, int batch,
                                int frames, int tokens, int heads, int head_dim, int dtype, int temporal, void* stream) {
  B200_DT(dtype);
  B200_TRY(b200::check_arch());
  return b200::launch_attention_bwd(qkv16, o16, do16, dqkv16, stats, batch, frames, tokens, heads, head_dim, dtype == B200_BF16,
                                    temporal, static_cast<cudaStream_t>(stream));
}
B200_API int b200_ada_outer(const float* dmod, int64_t dmod_batch_stride, const void* sc16, float* dW, int batch, int NA, int dim,
                            int dtype, void* stream) {
  B200_DT(dtype);
  return b200::launch_ada_outer(dmod, dmod_batch_stride, sc16, dW, batch, NA, dim, dtype == B200_BF16, static_cast<cudaStream_t>(stream));
}
B200_API int b200_ada_dsc(const float* dmod, int64_t dmod_batch_stride, const void* w16, float* dsc, int batch, int NA, int dim,
                          int dtype, void* stream) {
  B200_DT(dtype);
  return b200::launch_ada_dsc(dmod, dmod_batch_stride, w16, dsc, batch, NA, dim, dtype == B200_BF16, static_cast<cudaStream_t>(stream));
}
#undef B200_DT

}  // extern "C"
