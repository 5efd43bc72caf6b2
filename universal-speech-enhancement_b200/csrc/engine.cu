// Host runtime of the score network: weight packing, workspace planning, the launch program of one
// NCSN++ evaluation and the reverse-diffusion loop.  Mirrors the module list the reference builds in
// NCSNpp.__init__ (backbones/ncsnpp.py:186-316) and the dataflow of NCSNpp.forward (:324-501); every
// tensor is NHWC, ResBlocks follow layerspp.py:282-314.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/use_b200.h"
#include "kernels.h"

namespace use {

static thread_local char g_err[1024] = "";
static int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}
static int cuda_check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// host-side number formats
// ---------------------------------------------------------------------------------------------
static inline uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float f32_to_tf32(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) != 0x7f800000u) {
    u += 0xfffu + ((u >> 13) & 1u);
    u &= ~0x1fffu;
  }
  float r;
  memcpy(&r, &u, 4);
  return r;
}

// lo = false: the operand rounding of the act dtype; lo = true (fp32 only): the second TF32 term w - tf32(w) of the
// 3xTF32 split (USE_DTYPE_F32X3)
static void pack_conv_weight(int dt, const float* w, int O, int I, int ks, void* out, bool lo = false) {
  const int taps = ks * ks;
  for (int tap = 0; tap < taps; ++tap)
    for (int o = 0; o < O; ++o)
      for (int i = 0; i < I; ++i) {
        const float v = w[((size_t)o * I + i) * taps + tap];
        const size_t idx = ((size_t)tap * O + o) * I + i;
        if (dt == kBF16) ((uint16_t*)out)[idx] = f32_to_bf16(v);
        else ((float*)out)[idx] = lo ? f32_to_tf32(v - f32_to_tf32(v)) : f32_to_tf32(v);
      }
}

// pyramid-head weights [pc][C][3][3] fp32 -> act dtype [48][C], row tap * pc + co (rows >= 9 * pc are zero)
static void pack_head_weight(int dt, const float* w, int pc, int C, void* out) {
  memset(out, 0, (size_t)48 * C * act_size(dt));
  for (int tap = 0; tap < 9; ++tap)
    for (int co = 0; co < pc; ++co)
      for (int c = 0; c < C; ++c) {
        const float v = w[((size_t)co * C + c) * 9 + tap];
        const size_t idx = (size_t)(tap * pc + co) * C + c;
        if (dt == kBF16) ((uint16_t*)out)[idx] = f32_to_bf16(v);
        else ((float*)out)[idx] = f32_to_tf32(v);
      }
}

// ---------------------------------------------------------------------------------------------
// architecture plan (same order as the reference's all_modules list)
// ---------------------------------------------------------------------------------------------
enum Kind { K_GFP, K_LINEAR, K_CONV3, K_RB, K_COMBINE, K_ATTN, K_GN };
struct Mod {
  Kind kind;
  int cin = 0, cout = 0;
  bool up = false, down = false;
};

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
};

struct RbW {  // device offsets (bytes into the weight blob)
  size_t gn0_g, gn0_b, gn1_g, gn1_b, w0, w1, w2, bias1;
  int dense_off;  // row offset inside the stacked Dense_0 matrix
  bool has_conv2;
};

struct Arena {
  // first-fit offset allocator with coalescing; used both for the dry run (peak) and the real build
  std::vector<std::pair<size_t, size_t>> free_;  // (offset, size)
  std::map<size_t, size_t> live_;
  size_t top = 0, peak = 0;
  static size_t up(size_t n) { return (n + 1023) & ~size_t(1023); }
  size_t alloc(size_t n) {
    n = up(n ? n : 1);
    for (size_t i = 0; i < free_.size(); ++i) {
      if (free_[i].second >= n) {
        size_t off = free_[i].first;
        if (free_[i].second == n) free_.erase(free_.begin() + i);
        else { free_[i].first += n; free_[i].second -= n; }
        live_[off] = n;
        return off;
      }
    }
    size_t off = top;
    top += n;
    if (top > peak) peak = top;
    live_[off] = n;
    return off;
  }
  void release(size_t off) {
    auto it = live_.find(off);
    if (it == live_.end()) return;
    size_t n = it->second;
    live_.erase(it);
    if (off + n == top) {
      top = off;
      // absorb trailing free blocks
      bool again = true;
      while (again) {
        again = false;
        for (size_t i = 0; i < free_.size(); ++i)
          if (free_[i].first + free_[i].second == top) { top = free_[i].first; free_.erase(free_.begin() + i); again = true; break; }
      }
      return;
    }
    free_.push_back({off, n});
    // coalesce neighbours
    bool merged = true;
    while (merged) {
      merged = false;
      for (size_t i = 0; i < free_.size() && !merged; ++i)
        for (size_t j = 0; j < free_.size() && !merged; ++j)
          if (i != j && free_[i].first + free_[i].second == free_[j].first) {
            free_[i].second += free_[j].second;
            free_.erase(free_.begin() + j);
            merged = true;
          }
    }
  }
};

struct Act {  // activation tensor of the current program (batch implied)
  size_t off = 0;
  int C = 0, H = 0, W = 0;
  size_t stats_off = (size_t)-1;  // offset into the stats region or -1
  bool valid = false;
};

enum OpTag { TAG_CONV_TC = 0, TAG_GN_STATS, TAG_GN_APPLY, TAG_SMALL_CONV, TAG_ATTN, TAG_OTHER, TAG_COUNT };
static const char* kTagNames[TAG_COUNT] = {"conv_tc", "gn_stats", "gn_apply", "small_conv", "attn", "other"};

struct Op {
  std::function<void(cudaStream_t)> fn;
  int tag;
  int launches;
  double flops;  // algorithmic FLOPs of this op
  double bytes;  // algorithmic HBM bytes of this op (each operand / result once)
};

struct Program {
  int B = 0, F = 0, T = 0;
  std::vector<Op> ops;
  std::vector<TcConvPlan*> plans;
  std::vector<HeadPlan*> heads;
  size_t stats_bytes = 0;
  size_t pyramid_off = 0;   // final fp32 pyramid [B][F][T][4 or 2]
  size_t pyramid2_off = 0;  // 6-channel networks: channels 4, 5 as a second fp32 [B][F][T][2] tensor
  char* base = nullptr;
  // the launch sequence of one evaluation as a CUDA graph (captured on first use, replayed every step)
  cudaGraphExec_t graph = nullptr;
  bool graph_failed = false;
  long long graph_launches = 0;
  unsigned long long last_use = 0;  // LRU stamp of the program cache
  ~Program() {
    if (graph) cudaGraphExecDestroy(graph);
    for (auto* p : plans) tc_conv_plan_destroy(p);
    for (auto* p : heads) head_tc_plan_destroy(p);
  }
};

}  // namespace use

using namespace use;

struct use_engine {
  use_config cfg;
  int dt;
  std::vector<Mod> mods;
  std::map<std::string, HostTensor> host_w;
  // packed blob
  std::vector<uint8_t> blob;
  std::map<int, RbW> rbw;                         // module index -> ResBlock weight offsets
  std::map<std::string, size_t> off;              // misc named offsets
  int dense_rows = 0;
  int in_conv_idx = 3;  // index of the input convolution in all_modules (3 with the time-embedding MLP, 1 without)
  char* dev_w = nullptr;
  int num_sms = 148;
  // keyed by "B,F,T,base".  shared_ptr: a call holds ("pins") its programs for its whole duration, and eviction only
  // ever drops entries nobody holds -- a second lookup inside the same call can never free the first one's program
  std::map<std::string, std::shared_ptr<Program>> programs;
  unsigned long long program_tick = 0;
  // fixed head of the workspace (byte offsets)
  struct Head { size_t xr, xpad, t, gfp, sched, temb, temb_steps, dense, score, red, stats, arena; } head;
  // concurrency: a batch is split into `groups` halves that run on their own streams, so the HBM-bound GroupNorm
  // kernels of one half overlap the tensor-bound convolutions of the other (they fit beside the persistent conv CTA)
  int groups = 2;
  // GroupNorm + SiLU applied inside the convolution kernel's operand path (no normalised tensor in HBM); off: the
  // separate gn_apply kernel feeds the same convolutions (A/B testing: both give bit-identical results)
  bool fuse_gn = true;
  // the same for the pyramid heads (head_tc.cuh, round 2): the head reads the raw ResBlock output; off: gn_apply + plain head
  bool fuse_head = true;
  // latency mode (split-K clusters at the low-resolution levels, conv_tc_ks.cuh): 0 = never, 1 = always, 2 = auto: calls
  // with at most two clips.  A mode of the whole PROGRAM (part of its cache key): inside a mode per-clip results do not
  // depend on the batch; between the modes they differ in the last bits (same tolerance to the reference).
  int ksplit = 2;
  // scale / shift tables of the fused GroupNorms computed inside the consumer kernels (no gn_affine_kernel launches: 40 %
  // fewer launches per evaluation); off (default): one gn_affine_kernel launch per GroupNorm.  Bit-identical results.
  // Measured (bf16, CUDA graphs): batch 1 157.8 vs 157.1 ms per clip, batch 4 387.6 vs 377.5 ms per step -- inside a graph
  // the 98 tiny launches cost nothing, while the in-kernel table sits on the consumers' critical path; hence off.
  bool inline_gn = false;
  // USE_DTYPE_F32X3 (parity mode): dt = fp32 and every tensor-core convolution is the 3xTF32 split (three launches)
  bool x3 = false;
  // replay the ~230 launches of an evaluation as one CUDA graph (small batches are launch-latency bound: the deep levels
  // run dozens of kernels of a few microseconds each)
  bool use_graphs = true;
  cudaStream_t gstream[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  // instrumentation
  long long launches = 0;
  bool profiling = false;
  double prof_ms[8] = {0}, prof_flops[8] = {0}, prof_bytes[8] = {0}, prof_top_flops = 0, prof_top_ms = 0;
  long long prof_launches[8] = {0};
  struct ProfOp { int tag; float ms; double flops, bytes; };
  std::vector<ProfOp> prof_ops;
};

namespace use {

static void build_mods(use_engine* e) {
  const use_config& c = e->cfg;
  auto& m = e->mods;
  m.clear();
  const int nf = c.nf, L = c.num_levels;
  m.push_back({K_GFP});
  if (c.conditional) {
    m.push_back({K_LINEAR, 2 * nf, 4 * nf});
    m.push_back({K_LINEAR, 4 * nf, 4 * nf});
  }
  e->in_conv_idx = (int)m.size();
  m.push_back({K_CONV3, c.input_channels, nf});
  std::vector<int> hs{nf};
  int in_ch = nf;
  for (int l = 0; l < L; ++l) {
    for (int r = 0; r < c.num_res_blocks; ++r) {
      int out_ch = nf * c.ch_mult[l];
      m.push_back({K_RB, in_ch, out_ch});
      in_ch = out_ch;
      hs.push_back(in_ch);
    }
    if (l != L - 1) {
      Mod d{K_RB, in_ch, in_ch};
      d.down = true;
      m.push_back(d);
      m.push_back({K_COMBINE, c.input_channels, in_ch});
      hs.push_back(in_ch);
    }
  }
  in_ch = hs.back();
  m.push_back({K_RB, in_ch, in_ch});
  m.push_back({K_ATTN, in_ch, in_ch});
  m.push_back({K_RB, in_ch, in_ch});
  for (int l = L - 1; l >= 0; --l) {
    for (int r = 0; r < c.num_res_blocks + 1; ++r) {
      int out_ch = nf * c.ch_mult[l];
      m.push_back({K_RB, in_ch + hs.back(), out_ch});
      hs.pop_back();
      in_ch = out_ch;
    }
    m.push_back({K_GN, in_ch, in_ch});
    m.push_back({K_CONV3, in_ch, c.input_channels});
    if (l != 0) {
      Mod u{K_RB, in_ch, in_ch};
      u.up = true;
      m.push_back(u);
    }
  }
}

static const HostTensor* getw(use_engine* e, const std::string& name, std::vector<int64_t> shape) {
  auto it = e->host_w.find(name);
  if (it == e->host_w.end()) { fail("missing weight '%s'", name.c_str()); return nullptr; }
  if (it->second.shape != shape) {
    std::string s;
    for (auto d : it->second.shape) s += std::to_string(d) + ",";
    fail("weight '%s' has shape [%s], which does not match the architecture", name.c_str(), s.c_str());
    return nullptr;
  }
  return &it->second;
}

struct BlobWriter {
  std::vector<uint8_t>& b;
  size_t put(const void* p, size_t n) {
    size_t off = (b.size() + 255) & ~size_t(255);
    b.resize(off + n);
    memcpy(b.data() + off, p, n);
    return off;
  }
  size_t reserve(size_t n) {
    size_t off = (b.size() + 255) & ~size_t(255);
    b.resize(off + n);
    return off;
  }
};

static int pack_all(use_engine* e) {
  const use_config& c = e->cfg;
  const int dt = e->dt, nf = c.nf, D = 4 * nf;
  const size_t es = act_size(dt);
  e->blob.clear();
  e->rbw.clear();
  e->off.clear();
  BlobWriter bw{e->blob};
  auto P = [&](int i) { return "all_modules." + std::to_string(i); };
  auto putf = [&](const std::string& key, const std::string& name, std::vector<int64_t> shape) -> int {
    const HostTensor* t = getw(e, name, shape);
    if (!t) return 1;
    e->off[key] = bw.put(t->data.data(), t->data.size() * 4);
    return 0;
  };
  auto put_conv_tc = [&](const std::string& name, int O, int I, int ks, size_t* off) -> int {
    const HostTensor* t = getw(e, name, {O, I, ks, ks});
    if (!t) return 1;
    const size_t nb = (size_t)ks * ks * O * I * es;
    *off = bw.reserve(e->x3 ? 2 * nb : nb);  // 3xTF32: the lo term follows the hi term
    pack_conv_weight(dt, t->data.data(), O, I, ks, e->blob.data() + *off);
    if (e->x3) pack_conv_weight(dt, t->data.data(), O, I, ks, e->blob.data() + *off + nb, true);
    return 0;
  };
  {
    std::vector<float> z(1024, 0.f);
    e->off["zeros"] = bw.put(z.data(), z.size() * 4);
  }
  if (putf("gfp.W", P(0) + ".W", {nf})) return 1;
  if (c.conditional) {
    if (putf("l1.w", P(1) + ".weight", {D, 2 * nf}) || putf("l1.b", P(1) + ".bias", {D})) return 1;
    if (putf("l2.w", P(2) + ".weight", {D, D}) || putf("l2.b", P(2) + ".bias", {D})) return 1;
  }
  if (putf("out.w", "output_layer.weight", {2, c.input_channels, 1, 1}) || putf("out.b", "output_layer.bias", {2})) return 1;
  // stacked Dense_0 matrix + base bias (conv0 bias + dense bias)
  int rows = 0;
  for (auto& m : e->mods) if (m.kind == K_RB) rows += m.cout;
  e->dense_rows = rows;
  std::vector<float> dW(c.conditional ? (size_t)rows * D : 0), dbase(rows);
  int row = 0;
  for (size_t i = 0; i < e->mods.size(); ++i) {
    const Mod& m = e->mods[i];
    const std::string p = P((int)i);
    switch (m.kind) {
      case K_GFP: case K_LINEAR: break;
      case K_CONV3: {
        if (putf(p + ".w", p + ".weight", {m.cout, m.cin, 3, 3}) || putf(p + ".b", p + ".bias", {m.cout})) return 1;
        if (m.cin == c.input_channels && tc_conv_supported(dt, m.cout) && m.cout != 32 && !e->x3) {
          // input conv 4 -> nf: tcgen05 layout with the input channels zero-padded to one 128-byte chunk
          const int ck = 128 / (int)es;
          const HostTensor* t = getw(e, p + ".weight", {m.cout, m.cin, 3, 3});
          std::vector<float> wp((size_t)m.cout * ck * 9, 0.f);
          for (int o = 0; o < m.cout; ++o)
            for (int ci = 0; ci < m.cin; ++ci)
              for (int tap = 0; tap < 9; ++tap) wp[((size_t)o * ck + ci) * 9 + tap] = t->data[((size_t)o * m.cin + ci) * 9 + tap];
          size_t off = bw.reserve((size_t)9 * m.cout * ck * es);
          pack_conv_weight(dt, wp.data(), m.cout, ck, 3, e->blob.data() + off);
          e->off[p + ".wtc"] = off;
        }
        if (m.cout == c.input_channels && head_tc_supported(dt, m.cin, m.cout == 6 ? 4 : m.cout) && !e->x3) {
          // pyramid head C -> pc: the nine taps folded into the MMA's N dimension (head_tc.cuh), rows tap * pc + co.
          // A 6-channel head (condition="both") is a 4-channel head (rows 0..3) + a 2-channel head (rows 4, 5)
          const HostTensor* t = getw(e, p + ".weight", {m.cout, m.cin, 3, 3});
          const int pc0 = m.cout == 6 ? 4 : m.cout;
          size_t off = bw.reserve((size_t)48 * m.cin * es);
          pack_head_weight(dt, t->data.data(), pc0, m.cin, e->blob.data() + off);
          e->off[p + ".whd"] = off;
          if (m.cout == 6) {
            size_t off2 = bw.reserve((size_t)48 * m.cin * es);
            pack_head_weight(dt, t->data.data() + (size_t)4 * m.cin * 9, 2, m.cin, e->blob.data() + off2);
            e->off[p + ".whd2"] = off2;
          }
        }
        break;
      }
      case K_GN: {
        if (putf(p + ".g", p + ".weight", {m.cin}) || putf(p + ".b", p + ".bias", {m.cin})) return 1;
        break;
      }
      case K_COMBINE: {
        if (putf(p + ".w", p + ".Conv_0.weight", {m.cout, m.cin, 1, 1}) || putf(p + ".b", p + ".Conv_0.bias", {m.cout})) return 1;
        if (m.cin == 6) {  // 6 input channels = a [C][4] and a [C][2] matrix for the 4- / 2-channel Combine kernels
          const HostTensor* t = getw(e, p + ".Conv_0.weight", {m.cout, m.cin, 1, 1});
          std::vector<float> w4((size_t)m.cout * 4), w2((size_t)m.cout * 2);
          for (int o = 0; o < m.cout; ++o) {
            for (int k = 0; k < 4; ++k) w4[(size_t)o * 4 + k] = t->data[(size_t)o * 6 + k];
            for (int k = 0; k < 2; ++k) w2[(size_t)o * 2 + k] = t->data[(size_t)o * 6 + 4 + k];
          }
          e->off[p + ".w4"] = bw.put(w4.data(), w4.size() * 4);
          e->off[p + ".w2"] = bw.put(w2.data(), w2.size() * 4);
        }
        break;
      }
      case K_ATTN: {
        if (putf(p + ".g", p + ".GroupNorm_0.weight", {m.cin}) || putf(p + ".gb", p + ".GroupNorm_0.bias", {m.cin})) return 1;
        for (int j = 0; j < 4; ++j) {
          const std::string n = p + ".NIN_" + std::to_string(j);
          if (putf(n + ".W", n + ".W", {m.cin, m.cin}) || putf(n + ".b", n + ".b", {m.cin})) return 1;
        }
        break;
      }
      case K_RB: {
        RbW r{};
        r.has_conv2 = (m.cin != m.cout) || m.up || m.down;
        const HostTensor *g0 = getw(e, p + ".GroupNorm_0.weight", {m.cin}), *b0 = getw(e, p + ".GroupNorm_0.bias", {m.cin});
        const HostTensor *g1 = getw(e, p + ".GroupNorm_1.weight", {m.cout}), *b1 = getw(e, p + ".GroupNorm_1.bias", {m.cout});
        const HostTensor *c0b = getw(e, p + ".Conv_0.bias", {m.cout}), *c1b = getw(e, p + ".Conv_1.bias", {m.cout});
        // Dense_0 exists in the state dict either way (temb_dim is always passed) but is dead without a time embedding
        const HostTensor *dw = c.conditional ? getw(e, p + ".Dense_0.weight", {m.cout, D}) : nullptr;
        const HostTensor *db = c.conditional ? getw(e, p + ".Dense_0.bias", {m.cout}) : nullptr;
        if (!g0 || !b0 || !g1 || !b1 || !c0b || !c1b || (c.conditional && (!dw || !db))) return 1;
        r.gn0_g = bw.put(g0->data.data(), m.cin * 4);
        r.gn0_b = bw.put(b0->data.data(), m.cin * 4);
        r.gn1_g = bw.put(g1->data.data(), m.cout * 4);
        r.gn1_b = bw.put(b1->data.data(), m.cout * 4);
        if (put_conv_tc(p + ".Conv_0.weight", m.cout, m.cin, 3, &r.w0)) return 1;
        if (put_conv_tc(p + ".Conv_1.weight", m.cout, m.cout, 3, &r.w1)) return 1;
        std::vector<float> bias1(c1b->data);
        if (r.has_conv2) {
          if (put_conv_tc(p + ".Conv_2.weight", m.cout, m.cin, 1, &r.w2)) return 1;
          const HostTensor* c2b = getw(e, p + ".Conv_2.bias", {m.cout});
          if (!c2b) return 1;
          for (int k = 0; k < m.cout; ++k) bias1[k] += c2b->data[k];
        }
        r.bias1 = bw.put(bias1.data(), m.cout * 4);
        r.dense_off = row;
        if (c.conditional) memcpy(&dW[(size_t)row * D], dw->data.data(), (size_t)m.cout * D * 4);
        for (int k = 0; k < m.cout; ++k) dbase[row + k] = c0b->data[k] + (c.conditional ? db->data[k] : 0.f);
        row += m.cout;
        e->rbw[(int)i] = r;
        break;
      }
    }
  }
  if (c.conditional) e->off["dense.W"] = bw.put(dW.data(), dW.size() * 4);
  e->off["dense.base"] = bw.put(dbase.data(), dbase.size() * 4);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// program construction
// ---------------------------------------------------------------------------------------------
struct Builder {
  use_engine* e;
  Program* prog;   // nullptr on the dry run
  int B, F, T;
  Arena arena;
  size_t stats_top = 0;
  char* base;      // workspace base (nullptr on the dry run)
  bool dry;
  int err = 0;
  bool latency = false;  // latency-mode program: split-K clusters at the low-resolution levels (use_engine::ksplit)
  const float kInvSqrt2 = 0.70710678118654752440f;

  size_t es() const { return act_size(e->dt); }
  char* ws(size_t off) const { return base + e->head.arena + off; }
  char* wt(size_t off) const { return e->dev_w + off; }
  const float* wf(const std::string& k) const { return (const float*)(e->dev_w + e->off.at(k)); }
  long long* stats_ptr(size_t off) const { return (long long*)(base + e->head.stats + off); }

  Act new_act(int C, int H, int W) {
    Act a;
    a.C = C; a.H = H; a.W = W;
    a.off = arena.alloc((size_t)B * H * W * C * es());
    a.valid = true;
    return a;
  }
  size_t new_f32(size_t n) { return arena.alloc(n * 4); }
  void free_act(Act& a) {
    if (a.valid) arena.release(a.off);
    a.valid = false;
  }
  void emit(std::function<void(cudaStream_t)> f, int tag = TAG_OTHER, int launches = 1, double flops = 0, double bytes = 0) {
    if (!dry) prog->ops.push_back(Op{std::move(f), tag, launches, flops, bytes});
  }
  void ensure_stats(Act& a) {
    if (a.stats_off != (size_t)-1) return;
    a.stats_off = stats_top;
    stats_top += (size_t)B * a.C * 2 * sizeof(long long);
    if (dry) return;
    const int dt = e->dt, Bn = B, HW = a.H * a.W, C = a.C;
    const void* x = ws(a.off);
    long long* st = stats_ptr(a.stats_off);
    emit([=](cudaStream_t s) { launch_gn_stats(dt, x, st, Bn, HW, C, s); }, TAG_GN_STATS, 1, 3.0 * Bn * HW * C,
         (double)Bn * HW * C * es());
  }
  void gn_apply(Act& s0, Act* s1, size_t gamma_off, size_t beta_off, int fir, bool silu, bool operand, Act& out, Act* raw) {
    // the resampling kernels work tile by tile: give them the scale / shift table instead of the statistics
    size_t aff_off = (size_t)-1;
    if (fir != 0 && !s1 && !e->inline_gn) aff_off = gn_affine(s0, nullptr, gamma_off, beta_off);
    ensure_stats(s0);
    if (s1) ensure_stats(*s1);
    if (dry) {
      if (aff_off != (size_t)-1) arena.release(aff_off);
      return;
    }
    const float* affp = aff_off != (size_t)-1 ? (const float*)ws(aff_off) : nullptr;
    GnSrc a{ws(s0.off), stats_ptr(s0.stats_off), s0.C};
    GnSrc b{nullptr, nullptr, 0};
    if (s1) b = GnSrc{ws(s1->off), stats_ptr(s1->stats_off), s1->C};
    const float* g = (const float*)wt(gamma_off);
    const float* bt = (const float*)wt(beta_off);
    void* o = ws(out.off);
    void* r = raw ? ws(raw->off) : nullptr;
    const int dt = e->dt, Bn = B, H = s0.H, W = s0.W;
    {
      const double nin = (double)Bn * H * W * (s0.C + (s1 ? s1->C : 0));
      const double nout = fir == 1 ? nin / 4 : (fir == 2 ? nin * 4 : nin);
      emit([=](cudaStream_t s) { launch_gn_apply(dt, a, b, g, bt, 1e-6f, fir, silu, operand, o, r, Bn, H, W, s, affp); },
           TAG_GN_APPLY, 1, 8.0 * nout, (nin + nout * (raw ? 2 : 1)) * es());
    }
    if (aff_off != (size_t)-1) arena.release(aff_off);
  }
  // scale / shift table [B][2][C] of GroupNorm(cat[s0, s1]) for a fused conv operand; returns its arena offset
  size_t gn_affine(Act& s0, Act* s1, size_t gamma_off, size_t beta_off) {
    ensure_stats(s0);
    if (s1) ensure_stats(*s1);
    const int Ct = s0.C + (s1 ? s1->C : 0);
    const size_t off = new_f32((size_t)B * 2 * Ct);
    if (dry) return off;
    GnSrc a{nullptr, stats_ptr(s0.stats_off), s0.C};
    GnSrc b{nullptr, nullptr, 0};
    if (s1) b = GnSrc{nullptr, stats_ptr(s1->stats_off), s1->C};
    const float* g = (const float*)wt(gamma_off);
    const float* bt = (const float*)wt(beta_off);
    float* o = (float*)ws(off);
    const int Bn = B, HW = s0.H * s0.W;
    emit([=](cudaStream_t s) { launch_gn_affine(a, b, g, bt, 1e-6f, HW, o, Bn, s); }, TAG_GN_STATS, 1, 0, 0);
    return off;
  }
  // inline GroupNorm of a fused convolution: hand the statistics to the kernel instead of a table
  void gn_inline(TcConvDesc& d, Act& s0, Act* s1, size_t gamma_off, size_t beta_off) {
    ensure_stats(s0);
    if (s1) ensure_stats(*s1);
    if (dry) return;
    d.gn_st0 = stats_ptr(s0.stats_off);
    d.gn_C0 = s0.C;
    d.gn_st1 = s1 ? stats_ptr(s1->stats_off) : nullptr;
    d.gn_C1 = s1 ? s1->C : 0;
    d.gn_HW = s0.H * s0.W;
    d.gn_gamma = (const float*)wt(gamma_off);
    d.gn_beta = (const float*)wt(beta_off);
    d.gn_eps = 1e-6f;
  }
  // stat_target: the conv's output tensor when the epilogue should also produce its GroupNorm statistics
  // flops_alg: algorithmic FLOPs when they differ from 2 px N sum(taps C) (the input conv's K is zero-padded to one chunk)
  void conv_tc(TcConvDesc d, Act* stat_target = nullptr, double flops_alg = 0) {
    if (stat_target) {
      stat_target->stats_off = stats_top;
      stats_top += (size_t)B * d.N * 2 * sizeof(long long);
      if (!dry) d.stats_acc = stats_ptr(stat_target->stats_off);
    }
    if (e->x3) { conv_tc_x3(d); return; }
    if (dry) return;
    emit_conv_plan(d, flops_alg);
  }
  void emit_conv_plan(const TcConvDesc& d0, double flops_alg, bool count = true) {
    char msg[512];
    TcConvDesc d = d0;
    d.latency = latency ? 1 : 0;
    TcConvPlan* p = tc_conv_plan_create(e->dt, d, e->num_sms, msg, sizeof(msg));
    if (!p) { err = fail("%s", msg); return; }
    prog->plans.push_back(p);
    double k = 0, cin = 0;
    for (int i = 0; i < d.nseg; ++i) { k += (double)d.seg[i].taps * d.seg[i].C; cin += d.seg[i].C; }
    const double px = (double)d.B * d.H * d.W;
    const double fl = flops_alg > 0 ? flops_alg : 2.0 * px * d.N * k;
    emit([=](cudaStream_t s) { tc_conv_launch(p, s); }, TAG_CONV_TC, 1, count ? fl : 0.0,
         count ? (px * (cin + d.N * (d.res ? 2 : 1))) * es() : 0.0);
  }
  // 3xTF32 parity mode: out = ((x_hi w_hi + bias) + (x_lo w_hi + (x_hi w_lo + res))) * scale as three launches of the same
  // tcgen05 kernel (small terms first; the fp32 adds of the epilogue chain them through its residual input)
  void conv_tc_x3(const TcConvDesc& d) {
    const size_t px = (size_t)d.B * d.H * d.W;
    size_t hi[3], lo[3];
    for (int i = 0; i < d.nseg; ++i) { hi[i] = new_f32(px * d.seg[i].C); lo[i] = new_f32(px * d.seg[i].C); }
    const size_t t1 = new_f32(px * d.N), t2 = new_f32(px * d.N);
    if (!dry) {
      TcConvDesc p1 = d, p2 = d, p3 = d;
      for (int i = 0; i < d.nseg; ++i) {
        const TcSegDesc& sg = d.seg[i];
        float *h = (float*)ws(hi[i]), *l = (float*)ws(lo[i]);
        const float* src = (const float*)sg.act;
        const int Ct = sg.C_tensor, c0 = sg.c0, C = sg.C;
        emit([=](cudaStream_t s) { launch_split_tf32(src, Ct, c0, C, h, l, px, s); }, TAG_OTHER, 1, 0, 0);
        const char* w_lo = (const char*)sg.w + (size_t)sg.taps * d.N * sg.Cw_total * 4;
        TcSegDesc a = sg;
        a.C_tensor = C; a.c0 = 0; a.aff = nullptr; a.aff_C = 0; a.aff_c0 = 0;
        a.act = h; a.w = w_lo; p1.seg[i] = a;
        a.act = l; a.w = sg.w; p2.seg[i] = a;
        a.act = h; a.w = sg.w; p3.seg[i] = a;
      }
      const float* zeros = wf("zeros");
      p1.bias = zeros; p1.bias_bstride = 0; p1.scale = 1.0f; p1.out = ws(t1); p1.stats_acc = nullptr;  // res = d.res
      p2.bias = zeros; p2.bias_bstride = 0; p2.scale = 1.0f; p2.out = ws(t2); p2.stats_acc = nullptr; p2.res = ws(t1);
      p3.res = ws(t2);
      emit_conv_plan(p1, 0, false);
      if (!err) emit_conv_plan(p2, 0, false);
      if (!err) emit_conv_plan(p3, 0, true);
    }
    arena.release(t1); arena.release(t2);
    for (int i = 0; i < d.nseg; ++i) { arena.release(hi[i]); arena.release(lo[i]); }
  }

  // ResnetBlockBigGANpp.forward (layerspp.py:282-314).  x1 != nullptr: input is cat[x0, x1].
  Act resblock(int idx, Act& x0, Act* x1) {
    const Mod& m = e->mods[idx];
    const RbW& w = e->rbw.at(idx);
    const int Cin = m.cin, Cout = m.cout;
    const int fir = m.down ? 1 : (m.up ? 2 : 0);
    const int Ho = m.down ? x0.H / 2 : (m.up ? x0.H * 2 : x0.H);
    const int Wo = m.down ? x0.W / 2 : (m.up ? x0.W * 2 : x0.W);
    const bool fuse = e->fuse_gn && !e->x3;
    const bool opnd = !e->x3;  // 3xTF32: the normalised tensors stay full fp32 and are split in front of each convolution
    const bool fuse0 = fuse && fir == 0;  // FIR blocks resample between GroupNorm/SiLU and Conv_0: separate kernel
    Act a0, raw;
    size_t aff0 = (size_t)-1;
    const bool inl = e->inline_gn;
    if (fuse0) {
      if (!inl) aff0 = gn_affine(x0, x1, w.gn0_g, w.gn0_b);
    } else {
      a0 = new_act(Cin, Ho, Wo);
      if (fir) raw = new_act(Cin, Ho, Wo);
      gn_apply(x0, x1, w.gn0_g, w.gn0_b, fir, true, opnd, a0, fir ? &raw : nullptr);
    }
    Act h1 = new_act(Cout, Ho, Wo);
    {
      TcConvDesc d{};
      if (fuse0) {
        const float* ap = (dry || inl) ? (const float*)1 : (const float*)ws(aff0);
        if (inl) gn_inline(d, x0, x1, w.gn0_g, w.gn0_b);
        d.seg[0] = TcSegDesc{dry ? nullptr : ws(x0.off), x0.C, 0, x0.C, dry ? nullptr : wt(w.w0), Cin, 0, 9, ap, Cin, 0};
        d.nseg = 1;
        if (x1) {
          d.seg[1] = TcSegDesc{dry ? nullptr : ws(x1->off), x1->C, 0, x1->C, dry ? nullptr : wt(w.w0), Cin, x0.C, 9, ap, Cin, x0.C};
          d.nseg = 2;
        }
      } else {
        d.nseg = 1;
        d.seg[0] = TcSegDesc{dry ? nullptr : ws(a0.off), Cin, 0, Cin, dry ? nullptr : wt(w.w0), Cin, 0, 9};
      }
      d.B = B; d.H = Ho; d.W = Wo; d.N = Cout;
      d.out = dry ? nullptr : ws(h1.off);
      if (e->cfg.conditional) {  // conv bias + Dense_0(act(temb)), per sample
        d.bias = dry ? nullptr : (const float*)(base + e->head.dense) + w.dense_off;
        d.bias_bstride = e->dense_rows;
      } else {
        d.bias = dry ? nullptr : wf("dense.base") + w.dense_off;
        d.bias_bstride = 0;
      }
      d.res = nullptr;
      d.scale = 1.0f;
      conv_tc(d, &h1);
    }
    if (fuse0) { if (!inl) arena.release(aff0); }
    else free_act(a0);
    Act a1;
    size_t aff1 = (size_t)-1;
    if (fuse) {
      if (!inl) aff1 = gn_affine(h1, nullptr, w.gn1_g, w.gn1_b);
    } else {
      a1 = new_act(Cout, Ho, Wo);
      gn_apply(h1, nullptr, w.gn1_g, w.gn1_b, 0, true, opnd, a1, nullptr);
      free_act(h1);
    }
    Act out = new_act(Cout, Ho, Wo);
    {
      TcConvDesc d{};
      if (fuse) {
        d.seg[0] = TcSegDesc{dry ? nullptr : ws(h1.off), Cout, 0, Cout, dry ? nullptr : wt(w.w1), Cout, 0, 9,
                             (dry || inl) ? (const float*)1 : (const float*)ws(aff1), Cout, 0};
        if (inl) gn_inline(d, h1, nullptr, w.gn1_g, w.gn1_b);
      } else
        d.seg[0] = TcSegDesc{dry ? nullptr : ws(a1.off), Cout, 0, Cout, dry ? nullptr : wt(w.w1), Cout, 0, 9};
      d.nseg = 1;
      d.res = nullptr;
      if (w.has_conv2) {
        if (fir) {
          d.seg[1] = TcSegDesc{dry ? nullptr : ws(raw.off), Cin, 0, Cin, dry ? nullptr : wt(w.w2), Cin, 0, 1};
          d.nseg = 2;
        } else {
          d.seg[1] = TcSegDesc{dry ? nullptr : ws(x0.off), x0.C, 0, x0.C, dry ? nullptr : wt(w.w2), Cin, 0, 1};
          d.nseg = 2;
          if (x1) {
            d.seg[2] = TcSegDesc{dry ? nullptr : ws(x1->off), x1->C, 0, x1->C, dry ? nullptr : wt(w.w2), Cin, x0.C, 1};
            d.nseg = 3;
          }
        }
      } else {
        d.res = dry ? nullptr : ws(x0.off);
      }
      d.B = B; d.H = Ho; d.W = Wo; d.N = Cout;
      d.out = dry ? nullptr : ws(out.off);
      d.bias = dry ? nullptr : (const float*)wt(w.bias1);
      d.bias_bstride = 0;
      d.scale = kInvSqrt2;
      conv_tc(d, m.down ? nullptr : &out);  // a down block's output is modified by Combine before any GroupNorm
    }
    if (fuse) { if (!inl) arena.release(aff1); free_act(h1); }
    else free_act(a1);
    if (fir) free_act(raw);
    return out;
  }

  // AttnBlockpp.forward (layerspp.py:77-93), all fp32 internally
  Act attn(int idx, Act& x) {
    const std::string p = "all_modules." + std::to_string(idx);
    const int C = x.C, Pn = x.H * x.W, M = B * Pn;
    Act hn = new_act(C, x.H, x.W);
    gn_apply(x, nullptr, e->off.at(p + ".g"), e->off.at(p + ".gb"), 0, false, false, hn, nullptr);
    const size_t n = (size_t)M * C;
    size_t q = new_f32(n), k = new_f32(n), v = new_f32(n), att = new_f32(n);
    Act out = new_act(C, x.H, x.W);
    if (!dry) {
      const int dt = e->dt, Bn = B;
      const void* hnp = ws(hn.off);
      float *qp = (float*)ws(q), *kp = (float*)ws(k), *vp = (float*)ws(v), *ap = (float*)ws(att);
      const float *W0 = wf(p + ".NIN_0.W"), *b0 = wf(p + ".NIN_0.b"), *W1 = wf(p + ".NIN_1.W"), *b1 = wf(p + ".NIN_1.b");
      const float *W2 = wf(p + ".NIN_2.W"), *b2 = wf(p + ".NIN_2.b"), *W3 = wf(p + ".NIN_3.W"), *b3 = wf(p + ".NIN_3.b");
      const void* xp = ws(x.off);
      void* op = ws(out.off);
      const float sc = kInvSqrt2;
      emit([=](cudaStream_t s) {
        launch_nin_qkv(dt, hnp, W0, b0, qp, W1, b1, kp, W2, b2, vp, M, C, s);
        launch_attn_core(qp, kp, vp, ap, Bn, Pn, C, s);
        launch_nin_proj(dt, ap, W3, b3, xp, sc, op, M, C, s);
      }, TAG_ATTN, 3, 8.0 * M * C * C + 4.0 * Bn * Pn * Pn * C, 0);
    }
    free_act(hn);
    arena.release(q); arena.release(k); arena.release(v); arena.release(att);
    return out;
  }

  void build() {
    const use_config& c = e->cfg;
    const int L = c.num_levels;
    const int dt = e->dt, Bn = B;
    const int npc = c.input_channels;  // channels of the input / output pyramids
    int idx = e->in_conv_idx;
    // input conv (ncsnpp.py:381); the input pyramid level 0 is the packed network input itself
    std::vector<Act> hs;
    const float* xr = dry ? nullptr : (const float*)(base + e->head.xr);
    {
      Act h0 = new_act(c.nf, F, T);
      const std::string inp = "all_modules." + std::to_string(e->in_conv_idx);
      if (e->off.count(inp + ".wtc")) {
        const int ck = 128 / (int)es();
        TcConvDesc d{};
        d.nseg = 1;
        d.seg[0] = TcSegDesc{dry ? nullptr : base + e->head.xpad, ck, 0, ck, dry ? nullptr : wt(e->off.at(inp + ".wtc")), ck, 0, 9};
        d.B = B; d.H = F; d.W = T; d.N = c.nf;
        d.out = dry ? nullptr : ws(h0.off);
        d.bias = dry ? nullptr : wf(inp + ".b");
        d.bias_bstride = 0;
        d.scale = 1.0f;
        conv_tc(d, &h0, 2.0 * B * F * T * c.nf * 9.0 * c.input_channels);  // algorithmic K = 9 * input_channels, not the padded chunk
      } else if (!dry) {
        const float *w = wf(inp + ".w"), *b = wf(inp + ".b");
        void* o = ws(h0.off);
        const int H = F, W = T, N = c.nf;
        emit([=](cudaStream_t s) { launch_conv_in4(dt, xr, w, b, o, Bn, H, W, N, s); }, TAG_SMALL_CONV, 1,
             2.0 * Bn * H * W * N * 36, (double)Bn * H * W * (16 + N * es()));
      }
      hs.push_back(h0);
      idx = e->in_conv_idx + 1;
    }
    // A 6-channel pyramid (condition="both") is kept as a 4-channel + a 2-channel tensor: every pyramid op is channel-wise
    // (FIR, sums) or linear in the channels (Combine, heads, output layer), so the 4- / 2-channel kernels do the work
    const int nparts = npc == 6 ? 2 : 1;
    const int part_pc[2] = {npc == 6 ? 4 : npc, 2};
    const int part_c0[2] = {0, 4};
    size_t pyr_off[2] = {(size_t)-1, (size_t)-1};  // fp32 input pyramid of the current level (arena), level 0 = xr
    int pH = F, pW = T;
    for (int l = 0; l < L; ++l) {
      for (int r = 0; r < c.num_res_blocks; ++r) {
        Act h = resblock(idx++, hs.back(), nullptr);
        hs.push_back(h);
      }
      if (l != L - 1) {
        Act h = resblock(idx++, hs.back(), nullptr);
        // input_pyramid = FIR-down(input_pyramid); h = Conv1x1(input_pyramid) + h  (ncsnpp.py:404-406)
        size_t np[2] = {(size_t)-1, (size_t)-1};
        for (int k = 0; k < nparts; ++k) np[k] = new_f32((size_t)B * (pH / 2) * (pW / 2) * part_pc[k]);
        // the Combine kernel also produces the GroupNorm statistics of its output (the next ResBlock's GroupNorm_0)
        h.stats_off = stats_top;
        if (!dry) {
          const std::string p = "all_modules." + std::to_string(idx);
          void* hp = ws(h.off);
          const int HW = h.H * h.W, C = h.C, H = pH, W = pW;
          long long* hst = stats_ptr(h.stats_off);
          for (int k = 0; k < nparts; ++k) {
            const float* src = (pyr_off[k] == (size_t)-1) ? xr + (size_t)B * F * T * part_c0[k] : (const float*)ws(pyr_off[k]);
            float* dst = (float*)ws(np[k]);
            const int pc = part_pc[k];
            const float* cw = npc == 6 ? wf(k == 0 ? p + ".w4" : p + ".w2") : wf(p + ".w");
            const float* cb = k == 0 ? wf(p + ".b") : wf("zeros");      // (bias + w4 p4 + h) + w2 p2
            long long* st_k = k == nparts - 1 ? hst : nullptr;          // statistics of the FINAL sum only
            emit([=](cudaStream_t s) {
              launch_fir4_down(src, dst, Bn, H, W, pc, s);
              launch_combine(dt, hp, dst, cw, cb, hp, st_k, Bn, HW, C, pc, s);
            }, TAG_SMALL_CONV, 2, 2.0 * Bn * HW * C * pc, 2.0 * Bn * HW * C * es());
          }
        }
        stats_top += (size_t)B * h.C * 2 * sizeof(long long);
        for (int k = 0; k < nparts; ++k) {
          if (pyr_off[k] != (size_t)-1) arena.release(pyr_off[k]);
          pyr_off[k] = np[k];
        }
        pH /= 2; pW /= 2;
        idx++;
        hs.push_back(h);
      }
    }
    for (int k = 0; k < nparts; ++k)
      if (pyr_off[k] != (size_t)-1) arena.release(pyr_off[k]);
    // bottleneck
    Act h = resblock(idx++, hs.back(), nullptr);
    {
      Act a = attn(idx++, h);
      free_act(h);
      h = a;
    }
    {
      Act r = resblock(idx++, h, nullptr);
      free_act(h);
      h = r;
    }
    // up path
    size_t opyr[2] = {(size_t)-1, (size_t)-1};  // fp32 output pyramid [B][H][W][4 or 2] (+ [B][H][W][2] for 6 channels)
    for (int l = L - 1; l >= 0; --l) {
      for (int r = 0; r < c.num_res_blocks + 1; ++r) {
        Act skip = hs.back();
        hs.pop_back();
        Act o = resblock(idx++, h, &skip);
        free_act(h);
        free_act(skip);
        h = o;
      }
      // pyramid: GN -> SiLU -> conv3x3 C->pc (+ FIR-up of the previous pyramid)   (ncsnpp.py:440-461)
      {
        const std::string pg = "all_modules." + std::to_string(idx), pc = "all_modules." + std::to_string(idx + 1);
        const bool head_tc = e->off.count(pc + ".whd") != 0;
        // fused GroupNorm operand (head_tc.cuh): the head reads the RAW ResBlock output and normalises it in shared memory
        // (from 4 clips on: measured bf16, 16 clips 1404.6 -> 1399.7 ms per step, but batch 1 145.6 -> 146.3 ms per clip -- with
        // one 32 KB window in flight instead of two the fused head itself runs at 1.8 TB/s against 5.3-6.3 TB/s plain)
        const bool head_fuse = head_tc && e->fuse_gn && e->fuse_head && B >= 4;
        Act a;
        size_t haff = (size_t)-1;
        if (head_fuse) {
          haff = gn_affine(h, nullptr, e->off.at(pg + ".g"), e->off.at(pg + ".b"));
        } else {
          a = new_act(h.C, h.H, h.W);
          gn_apply(h, nullptr, e->off.at(pg + ".g"), e->off.at(pg + ".b"), 0, true, head_tc, a, nullptr);
        }
        size_t np[2] = {(size_t)-1, (size_t)-1};
        for (int k = 0; k < nparts; ++k) np[k] = new_f32((size_t)B * h.H * h.W * part_pc[k]);
        if (head_tc) {
          for (int k = 0; k < nparts && !dry; ++k) {
            char msg[512];
            HeadPlan* hp = head_tc_plan_create(e->dt, head_fuse ? ws(h.off) : ws(a.off),
                                               wt(e->off.at(k == 0 ? pc + ".whd" : pc + ".whd2")), wf(pc + ".b") + part_c0[k],
                                               opyr[k] == (size_t)-1 ? nullptr : (const float*)ws(opyr[k]), (float*)ws(np[k]), B,
                                               h.H, h.W, h.C, part_pc[k], e->num_sms, msg, sizeof(msg),
                                               head_fuse ? (const float*)ws(haff) : nullptr);
            if (!hp) { err = fail("%s", msg); return; }
            prog->heads.push_back(hp);
            const double px = (double)B * h.H * h.W;
            emit([=](cudaStream_t s) { head_tc_launch(hp, s); }, TAG_SMALL_CONV, 1, 2.0 * px * h.C * 9 * part_pc[k],
                 px * (h.C * es() + 4.0 * part_pc[k]));
          }
        } else if (npc != 4) {
          err = fail("pyramid head with %d channels needs the tensor-core head (C %% %d == 0, not the fp32x3 mode)", npc,
                     128 / (int)es());
        } else if (!dry) {
          const void* ap = ws(a.off);
          const float *w = wf(pc + ".w"), *b = wf(pc + ".b");
          const float* prev = (opyr[0] == (size_t)-1) ? nullptr : (const float*)ws(opyr[0]);
          float* o = (float*)ws(np[0]);
          const int H = h.H, W = h.W, C = h.C;
          emit([=](cudaStream_t s) { launch_conv_out4(dt, ap, w, b, prev, o, Bn, H, W, C, s); }, TAG_SMALL_CONV, 1,
               2.0 * Bn * H * W * C * 36, (double)Bn * H * W * (C * es() + 16));
        }
        if (head_fuse) arena.release(haff);
        else free_act(a);
        for (int k = 0; k < nparts; ++k) {
          if (opyr[k] != (size_t)-1) arena.release(opyr[k]);
          opyr[k] = np[k];
        }
        idx += 2;
      }
      if (l != 0) {
        Act u = resblock(idx++, h, nullptr);
        free_act(h);
        h = u;
      }
    }
    free_act(h);
    if (!dry) {
      prog->pyramid_off = e->head.arena + opyr[0];
      prog->pyramid2_off = nparts > 1 ? e->head.arena + opyr[1] : 0;
    }
    if (!dry) prog->stats_bytes = stats_top;
    if ((size_t)idx != e->mods.size() || !hs.empty()) err = fail("internal: module walk mismatch (%d of %zu)", idx, e->mods.size());
  }
};

static size_t align_up(size_t n, size_t a) { return (n + a - 1) / a * a; }
constexpr int kMaxSteps = 1024;  // reverse-diffusion steps one use_pc_sample call can schedule

// lays out the fixed head of the workspace and returns total bytes (dry run of the arena)
static int plan_workspace(use_engine* e, int B, int F, int T, size_t* total, size_t* stats_bytes) {
  const use_config& c = e->cfg;
  if (F % (1 << (c.num_levels - 1)) || T % (1 << (c.num_levels - 1)))
    return fail("spectrogram %dx%d is not divisible by 2^%d", F, T, c.num_levels - 1);
  Builder b{e, nullptr, B, F, T};
  b.base = nullptr;
  b.dry = true;
  // provisional head so pointer arithmetic on the dry run stays defined
  e->head = {};
  b.build();
  if (b.err) return 1;
  size_t off = 0;
  e->head.xr = off; off = align_up(off + (size_t)B * F * T * c.input_channels * 4, 1024);
  e->head.xpad = off; off = align_up(off + (size_t)B * F * T * 128, 1024);
  e->head.t = off; off = align_up(off + (size_t)B * 4, 1024);
  e->head.gfp = off; off = align_up(off + (size_t)B * 2 * c.nf * 4, 1024);
  e->head.sched = off; off = align_up(off + (size_t)kMaxSteps * (2 * c.nf + 1) * 4, 1024);  // per-step t_i and Fourier features
  e->head.temb = off; off = align_up(off + (size_t)B * 4 * c.nf * 4, 1024);
  e->head.temb_steps = off; off = align_up(off + (size_t)kMaxSteps * 4 * c.nf * 4, 1024);  // time embedding of every step
  e->head.dense = off; off = align_up(off + (size_t)B * e->dense_rows * 4, 1024);
  e->head.score = off; off = align_up(off + (size_t)B * F * T * 8, 1024);  // score of a corrector step (complex64)
  e->head.red = off; off = align_up(off + corrector_scratch_bytes(B), 1024);
  e->head.stats = off; off = align_up(off + b.stats_top, 1024);
  e->head.arena = off;
  *total = off + b.arena.peak;
  if (stats_bytes) *stats_bytes = b.stats_top;
  return 0;
}

constexpr size_t kMaxPrograms = 8;  // cached launch programs (distinct (B, F, T, workspace) combinations)

// Drop least-recently-used programs until at most kMaxPrograms - 1 remain, skipping every entry that is pinned by a
// running call (use_count > 1).  The destructor of a Program whose graph is still executing is safe:
// cudaGraphExecDestroy defers the release until the launch has completed.
static void evict_programs(use_engine* e) {
  while (e->programs.size() >= kMaxPrograms) {
    auto victim = e->programs.end();
    for (auto it = e->programs.begin(); it != e->programs.end(); ++it)
      if (it->second.use_count() == 1 && (victim == e->programs.end() || it->second->last_use < victim->second->last_use))
        victim = it;
    if (victim == e->programs.end()) return;  // everything is pinned: let the cache grow for this call
    e->programs.erase(victim);
  }
}

static bool latency_mode(const use_engine* e, int B_call) { return e->ksplit == 1 || (e->ksplit == 2 && B_call <= 2); }

// latency: build / fetch the latency-mode program (use_engine::ksplit; decided by the CALL's batch, not the group's)
static std::shared_ptr<Program> get_program(use_engine* e, int B, int F, int T, void* workspace, size_t workspace_bytes,
                                            bool latency) {
  char key[128];
  snprintf(key, sizeof(key), "%d,%d,%d,%p,%d", B, F, T, workspace, (int)latency);
  size_t need = 0;
  if (plan_workspace(e, B, F, T, &need, nullptr)) return nullptr;
  if (workspace_bytes < need) {
    fail("workspace too small: %zu bytes given, %zu needed for B=%d F=%d T=%d", workspace_bytes, need, B, F, T);
    return nullptr;
  }
  auto it = e->programs.find(key);
  if (it != e->programs.end()) {
    it->second->last_use = ++e->program_tick;
    return it->second;
  }
  if (!e->dev_w) { fail("weights not uploaded"); return nullptr; }
  evict_programs(e);
  std::shared_ptr<Program> p(new Program());
  p->B = B; p->F = F; p->T = T; p->base = (char*)workspace;
  Builder b{e, p.get(), B, F, T};
  b.base = (char*)workspace;
  b.dry = false;
  b.latency = latency;
  b.build();
  if (b.err) return nullptr;
  p->last_use = ++e->program_tick;
  e->programs[key] = p;
  return p;
}

// one network evaluation: t / gfp already in the workspace head; xr packed
// silu(Linear(silu(Linear(gfp)))) for `rows` Fourier-feature rows (ncsnpp.py:349-368)
static void run_temb_mlp(use_engine* e, const float* gfp, int gfp_bstride, float* temb, int rows, cudaStream_t st) {
  launch_temb_mlp(gfp, gfp_bstride, (const float*)(e->dev_w + e->off.at("l1.w")), (const float*)(e->dev_w + e->off.at("l1.b")),
                  (const float*)(e->dev_w + e->off.at("l2.w")), (const float*)(e->dev_w + e->off.at("l2.b")), temb, rows,
                  e->cfg.nf, st);
  e->launches += 1;
}

// temb_shared != nullptr: the time embedding of this evaluation was computed ahead of the loop (one row for the whole batch)
static void run_network(use_engine* e, Program* p, cudaStream_t st, const float* gfp, int gfp_bstride, bool allow_graph,
                        const float* temb_shared = nullptr) {
  char* base = p->base;
  const int nf = e->cfg.nf;
  cudaMemsetAsync(base + e->head.stats, 0, p->stats_bytes, st);  // fixed-point accumulators start at zero
  if (e->cfg.conditional) {
    if (!temb_shared) run_temb_mlp(e, gfp, gfp_bstride, (float*)(base + e->head.temb), p->B, st);
    launch_dense_all(temb_shared ? temb_shared : (const float*)(base + e->head.temb), temb_shared ? 0 : 4 * nf,
                     (const float*)(e->dev_w + e->off.at("dense.W")), (const float*)(e->dev_w + e->off.at("dense.base")),
                     (float*)(base + e->head.dense), p->B, e->dense_rows, 4 * nf, st);
    e->launches += 1;
  }
  if (!e->profiling) {
    if (allow_graph && e->use_graphs && !p->graph && !p->graph_failed) {
      // capture needs a non-legacy stream (use_pc_sample runs on the engine's own streams); anything that refuses to be
      // captured falls back to plain launches for good
      cudaGraph_t g = nullptr;
      if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        long long n = 0;
        for (auto& op : p->ops) { op.fn(st); n += op.launches; }
        if (cudaStreamEndCapture(st, &g) == cudaSuccess && g != nullptr &&
            cudaGraphInstantiate(&p->graph, g, 0) == cudaSuccess) {
          p->graph_launches = n;
        } else {
          p->graph = nullptr;
          p->graph_failed = true;
        }
        if (g) cudaGraphDestroy(g);
      } else {
        p->graph_failed = true;
      }
      cudaGetLastError();
    }
    if (allow_graph && p->graph) {
      cudaGraphLaunch(p->graph, st);
      e->launches += p->graph_launches;
      return;
    }
    for (auto& op : p->ops) { op.fn(st); e->launches += op.launches; }
    return;
  }
  // profiling pass: one event pair per op, on the launch stream (never used inside a timed bench step)
  std::vector<cudaEvent_t> ev(p->ops.size() + 1);
  for (auto& x : ev) cudaEventCreate(&x);
  for (size_t i = 0; i < p->ops.size(); ++i) {
    cudaEventRecord(ev[i], st);
    p->ops[i].fn(st);
    e->launches += p->ops[i].launches;
  }
  cudaEventRecord(ev.back(), st);
  cudaStreamSynchronize(st);
  for (size_t i = 0; i < p->ops.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
    const Op& o = p->ops[i];
    e->prof_ms[o.tag] += ms;
    e->prof_flops[o.tag] += o.flops;
    e->prof_bytes[o.tag] += o.bytes;
    e->prof_launches[o.tag] += o.launches;
    if (o.tag == TAG_CONV_TC && o.flops > e->prof_top_flops) { e->prof_top_flops = o.flops; e->prof_top_ms = ms; }
    if (e->prof_ops.size() < 4096) e->prof_ops.push_back({o.tag, ms, o.flops, o.bytes});
  }
  for (auto& x : ev) cudaEventDestroy(x);
}

}  // namespace use

// =================================================================================================
// C ABI
// =================================================================================================
static int g_op_latency = 0;  // use_op_set_latency (test hook)

extern "C" {

int use_abi_version(void) { return USE_B200_ABI_VERSION; }
const char* use_last_error(void) { return g_err; }

use_engine* use_engine_create(const use_config* cfg) {
  if (!cfg) { fail("null config"); return nullptr; }
  if (cfg->num_levels < 1 || cfg->num_levels > 8 || cfg->nf <= 0 || (cfg->input_channels != 4 && cfg->input_channels != 2 && cfg->input_channels != 6) ||
      (cfg->act_dtype != USE_DTYPE_F32 && cfg->act_dtype != USE_DTYPE_BF16 && cfg->act_dtype != USE_DTYPE_F32X3)) {
    fail("unsupported config (levels=%d nf=%d input_channels=%d dtype=%d)", cfg->num_levels, cfg->nf,
         cfg->input_channels, cfg->act_dtype);
    return nullptr;
  }
  use_engine* e = new use_engine();
  e->cfg = *cfg;
  e->x3 = cfg->act_dtype == USE_DTYPE_F32X3;
  e->dt = e->x3 ? (int)kF32 : cfg->act_dtype;
  if (e->x3 && cfg->input_channels == 6) {
    fail("the fp32x3 parity mode is not built for the 6-channel (condition=\"both\") network");
    delete e;
    return nullptr;
  }
  build_mods(e);
  for (auto& m : e->mods) {
    if (m.kind == K_RB) {
      const int ck = 128 / (int)act_size(e->dt);
      if (!tc_conv_supported(e->dt, m.cout) || m.cin % ck) {
        fail("architecture not supported by the tcgen05 conv path: ResBlock %d -> %d channels", m.cin, m.cout);
        delete e;
        return nullptr;
      }
    }
  }
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) e->num_sms = sms;
  }
  cudaGetLastError();
  if (const char* v = getenv("USE_B200_FUSE_GN")) e->fuse_gn = v[0] != '0';
  if (const char* v = getenv("USE_B200_FUSE_HEAD")) e->fuse_head = v[0] != '0';
  if (const char* v = getenv("USE_B200_GRAPHS")) e->use_graphs = v[0] != '0';
  if (const char* v = getenv("USE_B200_INLINE_GN")) e->inline_gn = v[0] != '0';
  if (const char* v = getenv("USE_B200_KSPLIT")) e->ksplit = (v[0] >= '0' && v[0] <= '2') ? v[0] - '0' : 2;
  return e;
}

void use_engine_destroy(use_engine* e) { delete e; }

int use_engine_set_weight(use_engine* e, const char* name, const float* host, const int64_t* shape, int ndim) {
  if (!e || !name || !host) return fail("null argument");
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); n *= (size_t)shape[i]; }
  t.data.assign(host, host + n);
  e->host_w[name] = std::move(t);
  return 0;
}

int use_engine_pack(use_engine* e, size_t* bytes) {
  if (!e) return fail("null engine");
  if (pack_all(e)) return 1;
  if (bytes) *bytes = e->blob.size();
  return 0;
}

int use_engine_upload(use_engine* e, void* dev_weights, size_t bytes, void* stream) {
  if (!e || !dev_weights) return fail("null argument");
  if (e->blob.empty()) return fail("use_engine_pack has not been called");
  if (bytes < e->blob.size()) return fail("weight buffer too small: %zu < %zu", bytes, e->blob.size());
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemcpyAsync(dev_weights, e->blob.data(), e->blob.size(), cudaMemcpyHostToDevice, st);
  cudaStreamSynchronize(st);
  e->dev_w = (char*)dev_weights;
  e->programs.clear();
  return cuda_check("use_engine_upload");
}

static int min_group_batch() {
  static const int v = getenv("USE_B200_GROUP_MIN_BATCH") ? atoi(getenv("USE_B200_GROUP_MIN_BATCH")) : 4;
  return v < 2 ? 2 : v;
}
static int group_count(const use_engine* e, int B) {
  return (e->groups >= 2 && !e->profiling && B >= min_group_batch() && B % 2 == 0) ? 2 : 1;
}

int use_engine_workspace_bytes(use_engine* e, int B, int F, int T, size_t* bytes) {
  if (!e || !bytes) return fail("null argument");
  if (e->dense_rows == 0) return fail("use_engine_pack has not been called");
  // large enough for one program over B and for two half-batch programs side by side
  size_t whole = 0, half = 0;
  if (plan_workspace(e, B, F, T, &whole, nullptr)) return 1;
  if (B >= min_group_batch() && B % 2 == 0) {
    if (plan_workspace(e, B / 2, F, T, &half, nullptr)) return 1;
    half = 2 * align_up(half, 4096);
  }
  *bytes = std::max(whole, half);
  return 0;
}

int use_engine_set_option(use_engine* e, const char* key, int value) {
  if (!e || !key) return fail("null argument");
  if (!strcmp(key, "overlap_groups")) {
    if (value != 1 && value != 2) return fail("overlap_groups must be 1 or 2");
    e->groups = value;
    return 0;
  }
  if (!strcmp(key, "use_graphs")) {
    e->use_graphs = value != 0;
    e->programs.clear();
    return 0;
  }
  if (!strcmp(key, "fuse_gn")) {
    e->fuse_gn = value != 0;
    e->programs.clear();
    return 0;
  }
  if (!strcmp(key, "ksplit")) {
    if (value < 0 || value > 2) return fail("ksplit must be 0 (off), 1 (on) or 2 (auto: calls with at most two clips)");
    e->ksplit = value;  // part of the program key: nothing to drop
    return 0;
  }
  if (!strcmp(key, "fuse_head")) {
    e->fuse_head = value != 0;
    e->programs.clear();
    return 0;
  }
  if (!strcmp(key, "inline_gn")) {
    e->inline_gn = value != 0;
    e->programs.clear();
    return 0;
  }
  return fail("unknown option '%s'", key);
}

// one evaluation of the network; sign = -1 gives the score (-net), +1 the raw network output
static int net_forward(use_engine* e, int B, int F, int T, const void* x, const void* Y, const float* t_host,
                       const float* gfp_host, void* out, float sign, void* workspace, size_t workspace_bytes, void* stream,
                       const void* Y2 = nullptr, const void* sde_y = nullptr, float drift_g = 0.f, float drift_pf = 1.f) {
  if (!e || !x || !out || !workspace) return fail("null argument");
  const bool cond = e->cfg.conditional != 0;
  if (e->cfg.input_channels >= 4 && !Y) return fail("the conditioning spectrogram Y is required (input_channels >= 4)");
  if (e->cfg.input_channels == 6 && !Y2) return fail("the second conditioning spectrogram is required (input_channels = 6)");
  if ((cond || e->cfg.scale_by_sigma) && (!t_host || (cond && !gfp_host))) return fail("time inputs are required");
  std::shared_ptr<Program> pin = get_program(e, B, F, T, workspace, workspace_bytes, latency_mode(e, B));
  if (!pin) return 1;
  Program* p = pin.get();
  cudaStream_t st = (cudaStream_t)stream;
  if (t_host) cudaMemcpyAsync(p->base + e->head.t, t_host, (size_t)B * 4, cudaMemcpyHostToDevice, st);
  if (cond) cudaMemcpyAsync(p->base + e->head.gfp, gfp_host, (size_t)B * 2 * e->cfg.nf * 4, cudaMemcpyHostToDevice, st);
  const size_t per = (size_t)F * T;
  launch_pack_input(e->dt, e->cfg.input_channels, (const float2*)x, (const float2*)Y, (const float2*)Y2,
                    (float*)(p->base + e->head.xr), p->base + e->head.xpad, per * B, st);
  run_network(e, p, st, (const float*)(p->base + e->head.gfp), 2 * e->cfg.nf, false);  // caller's stream: plain launches
  StepArgs a{};
  a.pyramid = (const float*)(p->base + p->pyramid_off);
  a.pyramid2 = (const float*)(p->base + p->pyramid2_off);
  a.pc = e->cfg.input_channels;
  a.out_sign = sign;
  a.t = e->cfg.scale_by_sigma ? (const float*)(p->base + e->head.t) : nullptr;
  a.t_bstride = 1;
  a.ow = (const float*)(e->dev_w + e->off.at("out.w"));
  a.ob = (const float*)(e->dev_w + e->off.at("out.b"));
  a.score = (float2*)out;
  a.x = nullptr;
  if (sde_y) {  // reverse-time drift instead of the score: out = theta (y - x) - g^2 score pf
    a.score = nullptr;
    a.x = (const float2*)x;
    a.Y = (const float2*)sde_y;
    a.x_mean = (float2*)out;
    a.x_next = (float2*)out;
    a.mode = kStepDrift;
    a.theta = e->cfg.theta;
    a.G = drift_g;
    a.pf = drift_pf;
  }
  a.B = B;
  a.per_clip = per;
  launch_final_step(a, st);
  e->launches += 2;
  return cuda_check("net_forward");
}

int use_score_forward(use_engine* e, int B, int F, int T, const void* x, const void* Y, const float* t_host,
                      const float* gfp_host, void* score, void* workspace, size_t workspace_bytes, void* stream) {
  return net_forward(e, B, F, T, x, Y, t_host, gfp_host, score, -1.0f, workspace, workspace_bytes, stream);
}

int use_score_forward2(use_engine* e, int B, int F, int T, const void* x, const void* Y, const void* Y2, const float* t_host,
                       const float* gfp_host, void* score, void* workspace, size_t workspace_bytes, void* stream) {
  return net_forward(e, B, F, T, x, Y, t_host, gfp_host, score, -1.0f, workspace, workspace_bytes, stream, Y2);
}

int use_reverse_drift(use_engine* e, int B, int F, int T, const void* x, const void* sde_y, const void* cond,
                      const void* cond2, const float* t_host, const float* gfp_host, float g, int probability_flow, void* drift,
                      void* workspace, size_t workspace_bytes, void* stream) {
  if (!sde_y) return fail("null argument");
  return net_forward(e, B, F, T, x, cond ? cond : sde_y, t_host, gfp_host, drift, -1.0f, workspace, workspace_bytes, stream, cond2,
                     sde_y, g, probability_flow ? 0.5f : 1.0f);
}

int use_net_forward(use_engine* e, int B, int F, int T, const void* x, const void* Y, const float* t_host,
                    const float* gfp_host, void* out, void* workspace, size_t workspace_bytes, void* stream) {
  return net_forward(e, B, F, T, x, Y, t_host, gfp_host, out, 1.0f, workspace, workspace_bytes, stream);
}

int use_train_forward(use_engine* e, int B, int F, int T, const void* X0, const void* Y, const float* t_host,
                      const float* gfp_host, const float* coef_host, const void* noise, uint64_t seed, uint32_t clip0,
                      int loss_type, void* x_t, float* loss, void* workspace, size_t workspace_bytes, void* stream) {
  if (!e || !X0 || !Y || !t_host || !gfp_host || !coef_host || !x_t || !loss || !workspace) return fail("null argument");
  if (loss_type != 0 && loss_type != 1) return fail("loss_type must be 0 (mse) or 1 (mae)");
  if (e->cfg.input_channels != 4 || !e->cfg.conditional || !e->cfg.scale_by_sigma)
    return fail("use_train_forward needs the 4-channel noise-conditional score network");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t per = (size_t)F * T;
  // plan first: the head offsets (score buffer, reduction scratch, per-sample coefficient slot) belong to this shape
  size_t need = 0;
  if (plan_workspace(e, B, F, T, &need, nullptr)) return 1;
  if (workspace_bytes < need) return fail("workspace too small: %zu bytes given, %zu needed", workspace_bytes, need);
  char* base = (char*)workspace;
  float* coef_dev = (float*)(base + e->head.gfp);  // [2][B] -- parked in the Fourier-feature slot until net_forward refills it
  if ((size_t)2 * B > (size_t)B * 2 * e->cfg.nf) return fail("internal: coefficient slot too small");
  cudaMemcpyAsync(coef_dev, coef_host, (size_t)2 * B * 4, cudaMemcpyHostToDevice, st);
  launch_perturb((const float2*)X0, (const float2*)Y, (const float2*)noise, coef_dev, (float2*)x_t, seed, clip0, B, per, st);
  float2* score = (float2*)(base + e->head.score);
  if (net_forward(e, B, F, T, x_t, Y, t_host, gfp_host, score, -1.0f, workspace, workspace_bytes, stream)) return 1;
  // net_forward refilled the gfp slot: upload the coefficients again (tiny) into the temb slot, free after the network ran
  float* coef2 = (float*)(base + e->head.temb);
  cudaMemcpyAsync(coef2, coef_host, (size_t)2 * B * 4, cudaMemcpyHostToDevice, st);
  launch_dsm_loss(score, (const float2*)noise, coef2, base + e->head.red, loss, loss_type, seed, clip0, B, per, st);
  e->launches += 3;
  return cuda_check("use_train_forward");
}

int use_pc_sample_ex(use_engine* e, int B, int F, int T, const void* Y, void* x_state, void* x_mean, int N,
                     const float* t_host, const float* G_host, const float* gfp_host, float prior_std, const void* noise,
                     uint64_t seed, uint32_t clip0, const use_sampler_opts* opts, void* workspace, size_t workspace_bytes,
                     void* stream) {
  if (!e || !Y || !x_state || !x_mean || !t_host || !G_host || !gfp_host || !workspace) return fail("null argument");
  if (N < 1 || N > kMaxSteps) return fail("N must be in [1, %d]", kMaxSteps);
  if ((e->cfg.input_channels != 4 && e->cfg.input_channels != 6) || !e->cfg.conditional || !e->cfg.scale_by_sigma)
    return fail("use_pc_sample needs the noise-conditional score network (input_channels=4 or 6, conditional, scale_by_sigma)");
  use_sampler_opts o{};
  o.predictor = USE_PRED_REVERSE_DIFFUSION;
  o.corrector = USE_CORR_NONE;
  o.denoise = 1;
  if (opts) o = *opts;
  if (o.predictor < 0 || o.predictor > USE_PRED_NONE) return fail("unknown predictor %d", o.predictor);
  if (o.corrector < 0 || o.corrector > USE_CORR_ALD) return fail("unknown corrector %d", o.corrector);
  const int cs = o.corrector == USE_CORR_NONE ? 0 : o.corrector_steps;  // NoneCorrector.n_steps = 0 (correctors.py:106-108)
  if (cs < 0) return fail("corrector_steps must be >= 0");
  if (o.predictor == USE_PRED_EULER_MARUYAMA && !o.g_host) return fail("euler_maruyama needs the diffusion table g_host[N]");
  if (o.corrector == USE_CORR_ALD && cs > 0 && !o.ald_step_host) return fail("ald needs the step-size table ald_step_host[N]");
  if (e->cfg.input_channels == 6 && !o.cond2) return fail("the 6-channel network needs the second conditioning spectrogram (cond2)");
  const int pe = o.predictor == USE_PRED_NONE ? 0 : 1;
  const int draws_per_step = cs + pe;  // normal draws per outer step, in the reference's order: corrector steps, predictor
  cudaStream_t st = (cudaStream_t)stream;
  // Langevin's step size is a batch mean (correctors.py:55-57): the batch must stay one group
  const int G = (o.corrector == USE_CORR_LANGEVIN && cs > 0) ? 1 : group_count(e, B), Bg = B / G;
  const size_t per = (size_t)F * T;
  const int nf2 = 2 * e->cfg.nf;
  size_t need = 0;
  if (plan_workspace(e, Bg, F, T, &need, nullptr)) return 1;
  const size_t slice = G > 1 ? align_up(need, 4096) : need;
  if (workspace_bytes < slice * G) return fail("workspace too small: %zu bytes given, %zu needed", workspace_bytes, slice * G);
  std::shared_ptr<Program> prog[2];  // pinned for the duration of the call (the cache never evicts a held program)
  cudaStream_t gs[2] = {st, st};
  for (int g = 0; g < G; ++g) {
    prog[g] = get_program(e, Bg, F, T, (char*)workspace + g * slice, slice, latency_mode(e, B));
    if (!prog[g]) return 1;
  }
  // the loop runs on the engine's own streams: two half-batches side by side, and (also for a single group) a stream
  // that CUDA graph capture accepts -- the caller's stream may be the legacy default stream
  const bool own_streams = G > 1 || e->use_graphs;
  if (own_streams) {
    for (int g = 0; g < G; ++g) {
      if (!e->gstream[g]) cudaStreamCreateWithFlags(&e->gstream[g], cudaStreamNonBlocking);
      if (!e->ev_join[g]) cudaEventCreateWithFlags(&e->ev_join[g], cudaEventDisableTiming);
      gs[g] = e->gstream[g];
    }
    if (!e->ev_fork) cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming);
    cudaEventRecord(e->ev_fork, st);
    for (int g = 0; g < G; ++g) cudaStreamWaitEvent(gs[g], e->ev_fork, 0);
  }
  // the float32 step schedule (t_i, Fourier features of log t_i) goes to the device once; vec_t = ones(B) * t_i is
  // batch-uniform, so the kernels read it with batch stride 0 and no host round trip happens inside the loop
  std::vector<float> sched((size_t)N * (nf2 + 1));
  memcpy(sched.data(), t_host, (size_t)N * 4);
  memcpy(sched.data() + N, gfp_host, (size_t)N * nf2 * 4);
  const float2* z = (const float2*)noise;
  for (int g = 0; g < G; ++g) {
    char* base = prog[g]->base;
    cudaMemcpyAsync(base + e->head.sched, sched.data(), sched.size() * 4, cudaMemcpyHostToDevice, gs[g]);
    // the time embedding depends on t_i only: all N of them in one launch ahead of the loop (it was a 2-GEMV chain in ONE
    // block per sample at the head of every evaluation: ~40 us of a 5.1 ms batch-1 step, B times the same numbers)
    run_temb_mlp(e, (const float*)(base + e->head.sched) + N, nf2, (float*)(base + e->head.temb_steps), N, gs[g]);
    const size_t off = (size_t)g * Bg * per;
    if (o.x_init) {
      if ((const float2*)o.x_init != (const float2*)x_state)
        cudaMemcpyAsync((float2*)x_state + off, (const float2*)o.x_init + off, per * Bg * sizeof(float2), cudaMemcpyDeviceToDevice,
                        gs[g]);
    } else {
      // x_0 = Y + z_0 * std(1)   (sdes.py:248-254)
      launch_prior((const float2*)Y + off, z ? z + off : nullptr, (float2*)x_state + off, prior_std, seed, clip0 + g * Bg, Bg,
                   per, gs[g]);
    }
  }
  for (int i = 0; i < N; ++i) {
    for (int g = 0; g < G; ++g) {  // interleaved enqueue: both streams always have work queued
      Program* p = prog[g].get();
      char* base = p->base;
      const size_t off = (size_t)g * Bg * per, n = per * Bg;
      const float* t_dev = (const float*)(base + e->head.sched) + i;
      const float* gfp_dev = (const float*)(base + e->head.sched) + N + (size_t)i * nf2;
      auto evaluate = [&](StepArgs& a) {  // one network evaluation at t_i from the current state + the fused tail
        launch_pack_input(e->dt, e->cfg.input_channels, (const float2*)x_state + off, (const float2*)(o.cond ? o.cond : Y) + off,
                          o.cond2 ? (const float2*)o.cond2 + off : nullptr, (float*)(base + e->head.xr), base + e->head.xpad,
                          n, gs[g]);
        run_network(e, p, gs[g], gfp_dev, 0, own_streams, (const float*)(base + e->head.temb_steps) + (size_t)i * 2 * nf2);
        a.pyramid = (const float*)(base + p->pyramid_off);
        a.pyramid2 = (const float*)(base + p->pyramid2_off);
        a.pc = e->cfg.input_channels;
        a.out_sign = -1.0f;
        a.t = t_dev;
        a.t_bstride = 0;
        a.ow = (const float*)(e->dev_w + e->off.at("out.w"));
        a.ob = (const float*)(e->dev_w + e->off.at("out.b"));
        a.B = Bg;
        a.per_clip = per;
        launch_final_step(a, gs[g]);
        e->launches += 2;
      };
      // ---- corrector: n_steps x (score evaluation, Langevin update)   (sampling/__init__.py:67, correctors.py:37-98)
      for (int j = 0; j < cs; ++j) {
        const unsigned draw = (unsigned)(i * draws_per_step + j);
        float2* grad = (float2*)(base + e->head.score);
        StepArgs a{};
        a.score = grad;
        evaluate(a);
        CorrectorArgs c{};
        c.x = (const float2*)x_state + off;
        c.grad = grad;
        c.z = z ? z + (size_t)(1 + draw) * per * B + off : nullptr;
        c.x_mean = (float2*)x_mean + off;
        c.x_next = (float2*)x_state + off;
        c.langevin = o.corrector == USE_CORR_LANGEVIN;
        c.snr = o.snr;
        c.step = o.corrector == USE_CORR_ALD ? o.ald_step_host[i] : 0.f;
        c.scratch = base + e->head.red;
        c.seed = seed;
        c.draw = draw;
        c.clip0 = clip0 + g * Bg;
        c.B = Bg;
        c.per_clip = per;
        launch_corrector_step(c, gs[g]);
        e->launches += c.langevin ? 3 : 1;
      }
      // ---- predictor   (sampling/__init__.py:68, predictors.py:40-68)
      if (pe) {
        const unsigned draw = (unsigned)(i * draws_per_step + cs);
        StepArgs a{};
        a.x = (const float2*)x_state + off;
        a.Y = (const float2*)Y + off;
        a.z = z ? z + (size_t)(1 + draw) * per * B + off : nullptr;
        a.x_mean = (float2*)x_mean + off;
        a.x_next = (float2*)x_state + off;
        a.theta = e->cfg.theta;
        const int Ndt = o.dt_steps > 0 ? o.dt_steps : N;
        a.dt = 1.0f / (float)Ndt;
        a.pf = o.probability_flow ? 0.5f : 1.0f;
        if (o.predictor == USE_PRED_EULER_MARUYAMA) {
          a.mode = kStepEulerMaruyama;
          a.G = o.g_host[i];
          a.Gz = o.probability_flow ? 0.f : o.g_host[i] * sqrtf(1.0f / (float)Ndt);
        } else {
          a.mode = kStepReverseDiffusion;
          a.G = G_host[i];
          a.Gz = o.probability_flow ? 0.f : G_host[i];
        }
        a.seed = seed;
        a.step = draw;
        a.clip0 = clip0 + g * Bg;
        evaluate(a);
      }
      if (o.trace)  // parity instrumentation: xt_mean after outer step i (NonePredictor returns (x, x): the state itself)
        cudaMemcpyAsync((float2*)o.trace + (size_t)i * per * B + off, (pe ? (const float2*)x_mean : (const float2*)x_state) + off,
                        n * sizeof(float2), cudaMemcpyDeviceToDevice, gs[g]);
    }
  }
  if (!o.denoise || (pe == 0 && o.denoise != 2)) {  // x_result = xt (sampling/__init__.py:69); NonePredictor returns (x, x) (predictors.py:78-79)
    for (int g = 0; g < G; ++g) {
      const size_t off = (size_t)g * Bg * per;
      cudaMemcpyAsync((float2*)x_mean + off, (const float2*)x_state + off, per * Bg * sizeof(float2), cudaMemcpyDeviceToDevice,
                      gs[g]);
    }
  }
  if (own_streams) {
    for (int g = 0; g < G; ++g) {
      cudaEventRecord(e->ev_join[g], gs[g]);
      cudaStreamWaitEvent(st, e->ev_join[g], 0);
    }
  }
  e->launches += G;
  return cuda_check("use_pc_sample");
}

int use_pc_sample(use_engine* e, int B, int F, int T, const void* Y, void* x_state, void* x_mean, int N,
                  const float* t_host, const float* G_host, const float* gfp_host, float prior_std, const void* noise,
                  uint64_t seed, uint32_t clip0, void* workspace, size_t workspace_bytes, void* stream) {
  return use_pc_sample_ex(e, B, F, T, Y, x_state, x_mean, N, t_host, G_host, gfp_host, prior_std, noise, seed, clip0, nullptr,
                          workspace, workspace_bytes, stream);
}

long long use_engine_launch_count(use_engine* e) { return e ? e->launches : -1; }

int use_engine_set_profiling(use_engine* e, int on) {
  if (!e) return fail("null engine");
  e->profiling = on != 0;
  for (int i = 0; i < 8; ++i) e->prof_ms[i] = e->prof_flops[i] = e->prof_bytes[i] = 0, e->prof_launches[i] = 0;
  e->prof_top_flops = e->prof_top_ms = 0;
  e->prof_ops.clear();
  return 0;
}

/* CSV "tag,ms,flops,bytes" per op of the profiled evaluations, in launch order */
int use_engine_get_profile_ops(use_engine* e, char* csv, size_t cap) {
  if (!e || !csv) return fail("null argument");
  size_t n = 0;
  for (auto& o : e->prof_ops) {
    if (n + 96 >= cap) break;
    n += snprintf(csv + n, cap - n, "%s,%.5f,%.4e,%.4e\n", kTagNames[o.tag], o.ms, o.flops, o.bytes);
  }
  if (n < cap) csv[n] = 0;
  return 0;
}

int use_engine_get_profile(use_engine* e, char* json, size_t cap) {
  if (!e || !json) return fail("null argument");
  size_t n = 0;
  n += snprintf(json + n, cap - n, "{");
  for (int i = 0; i < TAG_COUNT; ++i)
    n += snprintf(json + n, n < cap ? cap - n : 0, "%s\"%s\": {\"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e, \"launches\": %lld}",
                  i ? ", " : "", kTagNames[i], e->prof_ms[i], e->prof_flops[i], e->prof_bytes[i], e->prof_launches[i]);
  n += snprintf(json + n, n < cap ? cap - n : 0, ", \"top_conv\": {\"ms\": %.6f, \"flops\": %.6e}}", e->prof_top_ms, e->prof_top_flops);
  return n < cap ? 0 : fail("profile buffer too small");
}

int use_stft(use_engine* e, int B, int L, int Tp, const float* y, void* Y, const float* window, const float* twiddle,
             void* stream) {
  if (!e || !y || !Y || !window || !twiddle) return fail("null argument");
  const int T = 1 + L / e->cfg.hop;
  if (Tp < T) return fail("Tp=%d smaller than the frame count %d", Tp, T);
  if (L <= e->cfg.n_fft / 2) return fail("clip too short for reflect padding: L=%d", L);
  launch_stft(y, (float2*)Y, window, (const float2*)twiddle, B, L, e->cfg.n_fft, e->cfg.hop, T, Tp, e->cfg.spec_factor,
              e->cfg.spec_abs_exponent, (cudaStream_t)stream);
  return cuda_check("use_stft");
}

int use_istft(use_engine* e, int B, int L, int Tp, const void* X, float* y, float* frames_scratch, const float* window,
              const float* twiddle, const float* envelope, void* stream) {
  if (!e || !X || !y || !frames_scratch || !window || !twiddle || !envelope) return fail("null argument");
  if (L + e->cfg.n_fft / 2 > e->cfg.n_fft + e->cfg.hop * (Tp - 1)) return fail("requested length %d exceeds the signal", L);
  launch_istft((const float2*)X, frames_scratch, y, window, (const float2*)twiddle, envelope, B, L, e->cfg.n_fft,
               e->cfg.hop, Tp, e->cfg.spec_factor, e->cfg.spec_abs_exponent, (cudaStream_t)stream);
  return cuda_check("use_istft");
}

int use_upfirdn2d_f32(const float* in, float* out, int major, int in_h, int in_w, int minor, const float* kernel,
                      int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0,
                      int pad_y1, void* stream) {
  if (!in || !out || !kernel) return fail("null argument");
  if (up_x < 1 || up_y < 1 || down_x < 1 || down_y < 1 || kh < 1 || kw < 1) return fail("invalid upfirdn2d factors");
  launch_upfirdn2d(in, out, major, in_h, in_w, minor, kernel, kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0,
                   pad_y1, (cudaStream_t)stream);
  return cuda_check("use_upfirdn2d_f32");
}

// ---- predict-side audio preparation ---------------------------------------------------------------------------------
int use_resample_workspace_bytes(int B, int n_in, int n_out, size_t* bytes) {
  if (!bytes || B < 1 || n_in < 2 || n_out < 2) return fail("invalid resample shape (B=%d n_in=%d n_out=%d)", B, n_in, n_out);
  if (n_in > (1 << 27) || n_out > (1 << 27)) return fail("clip too long for the resampler");
  *bytes = resample_workspace_bytes(B, n_in, n_out);
  return 0;
}
int use_resample_fft_f32(const float* x, int B, int n_in, float* y, int n_out, int y_stride, void* work, size_t work_bytes,
                         void* stream) {
  if (!x || !y || !work) return fail("null argument");
  size_t need = 0;
  if (use_resample_workspace_bytes(B, n_in, n_out, &need)) return 1;
  if (work_bytes < need) return fail("resample workspace too small: %zu < %zu", work_bytes, need);
  if (y_stride < n_out) return fail("y_stride %d < n_out %d", y_stride, n_out);
  launch_resample_fft(x, y, B, n_in, n_out, y_stride, work, (cudaStream_t)stream);
  return cuda_check("use_resample_fft_f32");
}
int use_peak_normalize_pad_f32(float* y, const int* lengths_dev, int B, int stride, float target_peak, void* peaks_scratch,
                               void* stream) {
  if (!y || !lengths_dev || !peaks_scratch || B < 1 || stride < 1) return fail("invalid argument");
  launch_peak_normalize_pad(y, lengths_dev, B, stride, target_peak, (unsigned int*)peaks_scratch, (cudaStream_t)stream);
  return cuda_check("use_peak_normalize_pad_f32");
}

// ---- single-kernel exports ----------------------------------------------------------------------
int use_op_gn_stats(int dtype, const void* x, long long* stats, int B, int HW, int C, void* stream) {
  if (!x || !stats) return fail("null argument");
  launch_gn_stats(dtype, x, stats, B, HW, C, (cudaStream_t)stream);
  return cuda_check("use_op_gn_stats");
}
int use_op_gn_apply_aff(int dtype, const void* x0, const long long* stats0, int C0, const void* x1, const long long* stats1,
                        int C1, const float* gamma, const float* beta, float eps, int fir, int do_silu, int as_operand,
                        void* out_act, void* out_raw, int B, int Hin, int Win, const float* aff, void* stream) {
  if (aff != nullptr && (fir == 0 || C1 != 0)) return fail("aff is only used by the single-source resampling forms");
  launch_gn_apply(dtype, GnSrc{x0, stats0, C0}, GnSrc{x1, stats1, C1}, gamma, beta, eps, fir, do_silu != 0, as_operand != 0,
                  out_act, out_raw, B, Hin, Win, (cudaStream_t)stream, aff);
  return cuda_check("use_op_gn_apply");
}
int use_op_gn_apply(int dtype, const void* x0, const long long* stats0, int C0, const void* x1, const long long* stats1, int C1,
                    const float* gamma, const float* beta, float eps, int fir, int do_silu, int as_operand, void* out_act,
                    void* out_raw, int B, int Hin, int Win, void* stream) {
  return use_op_gn_apply_aff(dtype, x0, stats0, C0, x1, stats1, C1, gamma, beta, eps, fir, do_silu, as_operand, out_act,
                             out_raw, B, Hin, Win, nullptr, stream);
}
int use_op_gn_affine(const long long* stats0, int C0, const long long* stats1, int C1, const float* gamma,
                     const float* beta, float eps, int HW, float* aff, int B, void* stream) {
  if (!stats0 || !gamma || !beta || !aff) return fail("null argument");
  launch_gn_affine(GnSrc{nullptr, stats0, C0}, GnSrc{nullptr, stats1, stats1 ? C1 : 0}, gamma, beta, eps, HW, aff, B,
                   (cudaStream_t)stream);
  return cuda_check("use_op_gn_affine");
}
int use_op_conv_tc_gn(int dtype, int nseg, const void* const* seg_act, const int* seg_ctensor, const int* seg_c0,
                      const int* seg_c, const void* const* seg_w, const int* seg_cw, const int* seg_wc0, const int* seg_taps,
                      const float* const* seg_aff, const int* seg_aff_c, const int* seg_aff_c0, int B, int H, int W, int N,
                      const float* bias, int bias_bstride, const void* res, float scale, void* out, long long* stats,
                      void* stream) {
  if (nseg < 1 || nseg > 3) return fail("nseg must be 1..3");
  TcConvDesc d{};
  d.nseg = nseg;
  for (int i = 0; i < nseg; ++i)
    d.seg[i] = TcSegDesc{seg_act[i], seg_ctensor[i], seg_c0[i], seg_c[i], seg_w[i], seg_cw[i], seg_wc0[i], seg_taps[i],
                         seg_aff ? seg_aff[i] : nullptr, seg_aff ? seg_aff_c[i] : 0, seg_aff ? seg_aff_c0[i] : 0};
  d.B = B; d.H = H; d.W = W; d.N = N;
  d.out = out; d.bias = bias; d.bias_bstride = bias_bstride; d.res = res; d.scale = scale;
  d.stats_acc = stats;  // fixed-point accumulators, zeroed by the caller
  d.latency = g_op_latency;  // test hook (use_op_set_latency): the split-K cluster form for <= 10-tile images
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  char msg[512];
  TcConvPlan* p = tc_conv_plan_create(dtype, d, sms, msg, sizeof(msg));
  if (!p) return fail("%s", msg);
  tc_conv_launch(p, (cudaStream_t)stream);
  if (const char* reps_s = getenv("USE_B200_CONV_TIME")) {  // kernel-tuning hook (tools/conv_bench.py): time `reps` launches
    const int reps = atoi(reps_s) > 0 ? atoi(reps_s) : 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, (cudaStream_t)stream);
    for (int i = 0; i < reps; ++i) tc_conv_launch(p, (cudaStream_t)stream);
    cudaEventRecord(e1, (cudaStream_t)stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    fprintf(stderr, "USE_B200_CONV_TIME ms_per_launch=%.5f reps=%d\n", ms / reps, reps);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  cudaStreamSynchronize((cudaStream_t)stream);  // the plan (tensor maps live in kernel params) can go now
  tc_conv_plan_destroy(p);
  return cuda_check("use_op_conv_tc");
}
int use_op_conv_tc(int dtype, int nseg, const void* const* seg_act, const int* seg_ctensor, const int* seg_c0,
                   const int* seg_c, const void* const* seg_w, const int* seg_cw, const int* seg_wc0, const int* seg_taps,
                   int B, int H, int W, int N, const float* bias, int bias_bstride, const void* res, float scale, void* out,
                   long long* stats, void* stream) {
  return use_op_conv_tc_gn(dtype, nseg, seg_act, seg_ctensor, seg_c0, seg_c, seg_w, seg_cw, seg_wc0, seg_taps, nullptr,
                           nullptr, nullptr, B, H, W, N, bias, bias_bstride, res, scale, out, stats, stream);
}
int use_op_set_latency(int on) {
  g_op_latency = on != 0;
  return 0;
}
int use_op_conv_ref(int dtype, const void* x, const float* w, const float* bias, int bias_bstride, const void* res,
                    float scale, void* out, int B, int H, int W, int Cin, int Cout, int ksize, void* stream) {
  launch_conv_ref(dtype, x, w, bias, bias_bstride, res, scale, out, B, H, W, Cin, Cout, ksize, (cudaStream_t)stream);
  return cuda_check("use_op_conv_ref");
}
int use_op_conv_in4(int dtype, const float* x, const float* w, const float* bias, void* out, int B, int H, int W, int N,
                    void* stream) {
  launch_conv_in4(dtype, x, w, bias, out, B, H, W, N, (cudaStream_t)stream);
  return cuda_check("use_op_conv_in4");
}
int use_op_conv_out4(int dtype, const void* a, const float* w, const float* bias, const float* prev, float* out, int B,
                     int H, int W, int C, void* stream) {
  launch_conv_out4(dtype, a, w, bias, prev, out, B, H, W, C, (cudaStream_t)stream);
  return cuda_check("use_op_conv_out4");
}
static int op_head_tc(int dtype, const void* a, const float* aff, const float* w_oihw_host, const float* bias,
                      const float* prev, float* out, int B, int H, int W, int C, int pc, void* w_packed_dev, void* stream) {
  if (!a || !w_oihw_host || !bias || !out || !w_packed_dev) return fail("null argument");
  if (!head_tc_supported(dtype, C, pc)) return fail("pyramid head: unsupported C=%d pc=%d", C, pc);
  std::vector<uint8_t> packed((size_t)48 * C * act_size(dtype));
  pack_head_weight(dtype, w_oihw_host, pc, C, packed.data());
  cudaMemcpyAsync(w_packed_dev, packed.data(), packed.size(), cudaMemcpyHostToDevice, (cudaStream_t)stream);
  cudaStreamSynchronize((cudaStream_t)stream);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  char msg[512];
  HeadPlan* p = head_tc_plan_create(dtype, a, w_packed_dev, bias, prev, out, B, H, W, C, pc, sms, msg, sizeof(msg), aff);
  if (!p) return fail("%s", msg);
  head_tc_launch(p, (cudaStream_t)stream);
  if (const char* reps_s = getenv("USE_B200_CONV_TIME")) {
    const int reps = atoi(reps_s) > 0 ? atoi(reps_s) : 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, (cudaStream_t)stream);
    for (int i = 0; i < reps; ++i) head_tc_launch(p, (cudaStream_t)stream);
    cudaEventRecord(e1, (cudaStream_t)stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    fprintf(stderr, "USE_B200_CONV_TIME ms_per_launch=%.5f reps=%d\n", ms / reps, reps);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  cudaStreamSynchronize((cudaStream_t)stream);
  head_tc_plan_destroy(p);
  return cuda_check("use_op_head_tc");
}
int use_op_head_tc(int dtype, const void* a, const float* w_oihw_host, const float* bias, const float* prev, float* out,
                   int B, int H, int W, int C, int pc, void* w_packed_dev, void* stream) {
  return op_head_tc(dtype, a, nullptr, w_oihw_host, bias, prev, out, B, H, W, C, pc, w_packed_dev, stream);
}
int use_op_head_tc_gn(int dtype, const void* x, const float* aff, const float* w_oihw_host, const float* bias,
                      const float* prev, float* out, int B, int H, int W, int C, int pc, void* w_packed_dev, void* stream) {
  if (!aff) return fail("null argument");
  return op_head_tc(dtype, x, aff, w_oihw_host, bias, prev, out, B, H, W, C, pc, w_packed_dev, stream);
}
int use_op_combine(int dtype, const void* h, const float* pyr, const float* w, const float* bias, void* out, int B, int HW,
                   int C, int pc, void* stream) {
  if (pc != 2 && pc != 4) return fail("pc must be 2 or 4");
  launch_combine(dtype, h, pyr, w, bias, out, nullptr, B, HW, C, pc, (cudaStream_t)stream);
  return cuda_check("use_op_combine");
}
int use_op_combine_stats(int dtype, const void* h, const float* pyr, const float* w, const float* bias, void* out,
                         long long* stats, int B, int HW, int C, int pc, void* stream) {
  if (pc != 2 && pc != 4) return fail("pc must be 2 or 4");
  launch_combine(dtype, h, pyr, w, bias, out, stats, B, HW, C, pc, (cudaStream_t)stream);
  return cuda_check("use_op_combine_stats");
}
int use_op_fir4_down(const float* x, float* out, int B, int Hin, int Win, int pc, void* stream) {
  launch_fir4_down(x, out, B, Hin, Win, pc, (cudaStream_t)stream);
  return cuda_check("use_op_fir4_down");
}
int use_op_philox(void* z, uint64_t seed, uint32_t step, uint32_t clip0, int B, size_t per_clip, void* stream) {
  launch_philox_fill((float2*)z, seed, step, clip0, B, per_clip, (cudaStream_t)stream);
  return cuda_check("use_op_philox");
}
int use_pack_head_weight(int dtype, const float* w_oihw, int pc, int C, void* out) {
  if (!w_oihw || !out) return fail("null argument");
  if (pc != 2 && pc != 4) return fail("pc must be 2 or 4");
  pack_head_weight(dtype, w_oihw, pc, C, out);
  return 0;
}
int use_pack_conv_weight(int dtype, const float* w_oihw, int O, int I, int ksize, void* out) {
  if (!w_oihw || !out) return fail("null argument");
  pack_conv_weight(dtype, w_oihw, O, I, ksize, out);
  return 0;
}

}  // extern "C"
