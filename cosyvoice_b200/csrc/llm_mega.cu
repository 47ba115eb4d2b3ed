// Persistent decode kernel of the speech-token LM: ALL transformer layers of one decode step (cosyvoice/llm/llm.py:536-549
// loop body -> Qwen2Encoder.forward_one_step :242-254 -> transformers Qwen2 decoder layers) in ONE
// cooperative kernel instead of 7 PDL-chained kernels per layer.
//
// Why: the step is HBM-bound in principle (727.6 MB of bf16 weights per step shared by all rows) but the
// per-op chain was latency-bound: 171 kernels per step, each a launch + a few dependent memory round trips over <= 148 CTAs
// (977 us per step against a 0.126 ms floor, profiles/r01_ncu_summary.md §C).  Here
//   * one CTA per SM stays resident for the whole step; phases are separated by grid-wide barriers (one atomic + one polled
//     word in L2) instead of kernel boundaries;
//   * the weights are re-laid out once into PER-CTA STREAMS: the 16 KB SWIZZLE_128B operand blocks a CTA will consume, in
//     the order it consumes them, across phases and layers.  A dedicated producer warp walks that stream with plain bulk
//     copies into a 10-stage (160 KB) shared-memory ring and never waits for a phase boundary, so HBM keeps streaming
//     through the barriers and through the latency-bound phases (attention, split-K reductions);
//   * GEMMs are swap-AB tcgen05 (weights = the 128-row M side, the <= 64 batch rows = UMMA N) with the fp32 accumulator in
//     TMEM; activations (the B operand) arrive by TMA after the barrier;
//   * split-K partial sums go to a small fp32 scratch and are reduced in a fixed order (deterministic, no atomics) by the
//     attention unit (qkv) or by a per-row reduce + residual + RMSNorm phase (o_proj, down_proj).
// Phases per layer: qkv GEMM | bias + RoPE + cache append + attention | o GEMM | +res, RMSNorm | gate/up GEMM + SwiGLU |
// down GEMM | +res, RMSNorm.
#include "llm_decode_attn.cuh"
#include <algorithm>
#include <type_traits>

using namespace lm;

namespace {

constexpr int MG_THREADS = 512;          // warp 0: weight producer; warps 1..15: workers (1: MMA issue, 2: activation TMA, 4..7: epilogue)
constexpr int MG_WORKERS = MG_THREADS - 32;
constexpr int MG_NW = MG_WORKERS / 32;   // 15
constexpr int MG_BM = 128, MG_BK = 64;
constexpr uint32_t MG_BLK = MG_BM * MG_BK * 2;   // 16 KB weight block
constexpr int MG_MAX_STAGES = 10;
constexpr int MG_KCH = D / MG_BK;        // 14 K chunks of the 896-wide activations
constexpr int MG_PH = 4;                 // GEMM phases per layer: qkv, o, gate_up, down

struct MegaUnit { int nt, kc0, nblk, split; };
struct MegaLayerDev { const float* qkv_bias; const float* ln2; const float* next_gamma; };

struct MegaParams {
  const uint8_t* wstream;
  const unsigned long long* cta_off;   // [G] byte offset of the CTA's stream
  const int* cta_bpl;                  // [G] blocks per layer of the CTA
  const MegaUnit* units;               // [4][G]
  const MegaLayerDev* layers;          // [L]
  int num_layers, B, max_ctx;
  float* x; bf16* xn; bf16* att; bf16* ffa;
  float *part_qkv, *part_o, *part_down;
  int splits_qkv, splits_o, splits_down;
  bf16 *kcache, *vcache;
  unsigned long long kv_layer_stride;  // elements
  const int* ctx_len;
  const float* inv_freq;
  unsigned* bar;                       // [0] arrival counter (monotonic), [1] generation base of the next launch
  long long* tl;                       // optional phase timeline (CTA 0)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000ll) break;     // ~2 s: a protocol bug traps instead of hanging the GPU
  }
  __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// generic-proxy writes to global memory that a TMA (async proxy) of another CTA will read after the grid barrier, and the
// mirror fence on the reading side
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, %0;" ::"n"(MG_WORKERS) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// static schedule constants shared by the host-side table builder and the kernel (compile-time split counts let every
// partial-sum load of a reduction be issued before the first add)
constexpr int MG_CPU_QKV = 2, MG_CPU_O = 2, MG_CPU_DOWN = 8;        // K chunks per unit
constexpr int MG_SPL_QKV = MG_KCH / MG_CPU_QKV;                      // 7
constexpr int MG_SPL_O = MG_KCH / MG_CPU_O;                          // 7
constexpr int MG_SPL_DOWN = (DFF / MG_BK + MG_CPU_DOWN - 1) / MG_CPU_DOWN;   // 10
constexpr int MG_MAX_LAYERS = 32;

__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int BPAD>
__global__ void __launch_bounds__(MG_THREADS, 1)
lm_mega_kernel(const MegaParams p) {
  constexpr int NST = BPAD == 32 ? 10 : 7;
  constexpr uint32_t ACT_CHUNK = BPAD * 128;     // one 64-wide K chunk of the activations: [BPAD rows][128 B], SWIZZLE_128B
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BPAD >> 3) << 17) | ((uint32_t)(MG_BM >> 4) << 24);
  static_assert(decode_attn_smem_floats<MG_NW>() * 4 <= MG_KCH * ACT_CHUNK, "attention scratch must fit the activation buffer");
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MG_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[MG_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_acc;
  __shared__ uint32_t tmem_slot;
  __shared__ float red[16];
  __shared__ MegaUnit s_units[MG_PH];
  __shared__ MegaLayerDev s_layers[MG_MAX_LAYERS];

  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t ring = sbase, act = sbase + NST * MG_BLK;
  uint8_t* act_ptr = sptr + NST * MG_BLK;
  float* attn_sm = reinterpret_cast<float*>(act_ptr);       // aliases the activation buffer (disjoint phases)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, G = gridDim.x;
  const int B = p.B;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_acc), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + MG_PH) s_units[threadIdx.x - 64] = p.units[(threadIdx.x - 64) * G + cta];
  if (threadIdx.x >= 128 && threadIdx.x < 128 + p.num_layers) s_layers[threadIdx.x - 128] = p.layers[threadIdx.x - 128];
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((uint32_t)BPAD) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------------------------------- weight producer
    if (lane == 0) {
      const uint8_t* src = p.wstream + p.cta_off[cta];
      const int total = p.cta_bpl[cta] * p.num_layers;
      for (int it = 0; it < total; ++it) {
        const int s = it % NST;
        const uint32_t round = (uint32_t)(it / NST);
        if (it >= NST) mbar_wait(smem_u32(&bar_empty[s]), (round & 1u) ^ 1u);
        const uint32_t fb = smem_u32(&bar_full[s]);
        mbar_expect_tx(fb, MG_BLK);
        bulk_load(ring + s * MG_BLK, src + (size_t)it * MG_BLK, MG_BLK, fb);
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------------------------- workers
    const int wt = threadIdx.x - 32;
    const int ww = warp - 1;
    unsigned bar_target = __ldcg(p.bar + 1);
    const unsigned gen0 = bar_target;
    uint32_t acc_par = 0;
    int mma_it = 0, tl_i = 0;
    const bool stamp = p.tl != nullptr && cta == 0 && wt == 0;
    if (stamp) p.tl[tl_i++] = gtime();
    // fine-grained debug stamps of layer 1: CTA 0 (row 0 of the reduce phases, qkv / down GEMM unit) -> tl[512..], CTA 100 (o GEMM unit,
    // gate_up unit) -> tl[768..]
    int fs_i = 0;
    bool fs_on = false;
    long long* fs = p.tl ? p.tl + (cta == 0 ? 512 : 768) : nullptr;
    const bool fs_cta = p.tl != nullptr && (cta == 0 || cta == 100) && wt == 0;
#define FS() do { if (fs_cta && fs_on && fs_i < 250) fs[fs_i++] = gtime(); } while (0)

    // grid-wide barrier: every worker's global writes happen-before (bar.sync) the release-add of thread 0; the acquire-load that
    // sees the last arrival happens-before (bar.sync) every worker's following reads (which use ld.global.cg: L2 is the
    // coherence point, stale L1 lines of the previous layer's activations are never consulted)
    auto grid_sync = [&]() {
      bar_target += (unsigned)G;
      FS();
      worker_bar();
      FS();
      if (wt == 0) {
        red_release_add(p.bar, 1u);
        FS();
        const long long t0 = clock64();
        while ((int)(ld_acquire(p.bar) - bar_target) < 0) {
          if (clock64() - t0 > 4000000000ll) __trap();
        }
        if (stamp) p.tl[tl_i++] = gtime();
        FS();
      }
      worker_bar();
      FS();
    };

    // one GEMM phase: this CTA's unit = tile u.nt of the weight, K chunks [kc0, kc0 + nblk) of the activation matrix `actg`
    // [B][K] bf16.  mode 0: fp32 partial sums -> out_f32[(split * B + b) * N + n]; mode 1: SwiGLU on interleaved gate/up rows -> ffa
    auto gemm_phase = [&](int ph, const bf16* actg, int K, float* out_f32, int N, int mode) {
      const MegaUnit u = s_units[ph];
      if (u.nblk <= 0) return;
      {
        // activations -> shared memory in the K-major SWIZZLE_128B operand layout (what a TMA box {64, BPAD} would write): row b,
        // 16-byte chunk c of K chunk j at j*ACT_CHUNK + b*128 + ((c ^ (b & 7)) << 4).  All loads of a thread are issued before the
        // first store.  Rows >= B are left as they are: column b of the accumulator depends on row b only and is never stored.
        const int items = u.nblk * B * 8;
        constexpr int UN = 4;
        for (int i0 = wt; i0 < items; i0 += MG_WORKERS * UN) {
          uint4 v[UN];
#pragma unroll
          for (int q = 0; q < UN; ++q) {
            const int i = i0 + q * MG_WORKERS;
            if (i < items) {
              const int c = i & 7, b = (i >> 3) % B, j = (i >> 3) / B;
              v[q] = __ldcg(reinterpret_cast<const uint4*>(actg + (size_t)b * K + (size_t)(u.kc0 + j) * MG_BK) + c);
            }
          }
#pragma unroll
          for (int q = 0; q < UN; ++q) {
            const int i = i0 + q * MG_WORKERS;
            if (i < items) {
              const int c = i & 7, b = (i >> 3) % B, j = (i >> 3) / B;
              *reinterpret_cast<uint4*>(act_ptr + j * ACT_CHUNK + b * 128 + ((c ^ (b & 7)) << 4)) = v[q];
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the MMA (async proxy)
        worker_bar();
        FS();
      }
      if (warp == 1) {
        if (lane == 0) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          for (int j = 0; j < u.nblk; ++j) {
            const int it = mma_it + j;
            const int s = it % NST;
            mbar_wait(smem_u32(&bar_full[s]), (uint32_t)(it / NST) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = ring + s * MG_BLK, sb = act + j * ACT_CHUNK;
#pragma unroll
            for (int k = 0; k < MG_BK / 16; ++k) umma(tmem, desc_sw128(sa + k * 32), desc_sw128(sb + k * 32), IDESC, (j > 0 || k > 0) ? 1u : 0u);
            umma_commit(smem_u32(&bar_empty[s]));
          }
          umma_commit(smem_u32(&bar_acc));
        }
        __syncwarp();
      }
      if (warp >= 4 && warp < 8) {
        const int q = warp & 3;
        const int n = u.nt * MG_BM + q * 32 + lane;
        mbar_wait(smem_u32(&bar_acc), acc_par);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < BPAD; c += 16) {
          float acc[16];
          tmem_ld16(trow + (uint32_t)c, acc);
          if (mode == 1) {
            // rows of W are interleaved (2i = gate_i, 2i+1 = up_i): lanes pair up, the even lane emits silu(g) * u
            // (Qwen2 MLP act_fn(gate_proj(x)) * up_proj(x)) into column n/2
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float other = __shfl_xor_sync(0xffffffffu, acc[e], 1);
              const int b = c + e;
              if (!(lane & 1) && b < B) {
                const float g = acc[e];
                p.ffa[(size_t)b * DFF + (n >> 1)] = __float2bfloat16_rn(__fdividef(g, 1.f + fast_exp(-g)) * other);
              }
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int b = c + e;
              if (b < B) out_f32[((size_t)u.split * B + b) * N + n] = acc[e];
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
      acc_par ^= 1u;
      mma_it += u.nblk;
    };

    // x[b] += sum_s part[s][b]; xn[b] = bf16(gamma * rmsnorm(x[b]))   (one row per CTA; fixed summation order; every load of a
    // thread in flight at once)
    auto reduce_norm = [&](const float* part, auto splits_tag, const float* gamma) {
      constexpr int S = decltype(splits_tag)::value;
      for (int b = cta; b < B; b += G) {
        float ld[2][S + 1];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int n = wt + h * MG_WORKERS;
          const bool ok = n < D;
          ld[h][0] = ok ? __ldcg(p.x + (size_t)b * D + n) : 0.f;
#pragma unroll
          for (int s = 0; s < S; ++s) ld[h][s + 1] = ok ? __ldcg(part + ((size_t)s * B + b) * D + n) : 0.f;
        }
        float v[2], ss = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float a = 0.f;
#pragma unroll
          for (int s = 0; s < S; ++s) a += ld[h][s + 1];
          v[h] = ld[h][0] + a;
          const int n = wt + h * MG_WORKERS;
          if (n < D) {
            p.x[(size_t)b * D + n] = v[h];
            ss += v[h] * v[h];
          }
        }
        ss = warp_sum(ss);
        if (lane == 0) red[ww] = ss;
        FS();
        worker_bar();
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < MG_NW; ++i) tot += red[i];
        const float r = rsqrtf(tot / D + RMS_EPS);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int n = wt + h * MG_WORKERS;
          if (n < D) p.xn[(size_t)b * D + n] = __float2bfloat16_rn(gamma[n] * (v[h] * r));
        }
        worker_bar();
      }
    };

    for (int l = 0; l < p.num_layers; ++l) {
      const MegaLayerDev Lw = s_layers[l];
      fs_on = l == 1;
      bf16* kc_l = p.kcache + (size_t)l * p.kv_layer_stride;
      bf16* vc_l = p.vcache + (size_t)l * p.kv_layer_stride;
      if (cta < B * NKV && warp >= 8) {
        // this CTA's attention unit of the NEXT phase: pull its K / V rows (one 128-byte line per position) into L2 while the qkv
        // projection runs (the caches of 24 layers x 32 rows do not stay L2-resident between steps)
        const int b = cta / NKV, kvh = cta % NKV;
        const int L = min(p.ctx_len[b], p.max_ctx);
        const bf16* kb = kc_l + ((size_t)b * NKV + kvh) * p.max_ctx * HD;
        const bf16* vb = vc_l + ((size_t)b * NKV + kvh) * p.max_ctx * HD;
        for (int j = threadIdx.x - 256; j < L; j += MG_THREADS - 256) {
          prefetch_l2(kb + (size_t)j * HD);
          prefetch_l2(vb + (size_t)j * HD);
        }
      }
      gemm_phase(0, p.xn, D, p.part_qkv, QKV_N, 0);
      grid_sync();
      for (int u = cta; u < B * NKV; u += G) {
        const int b = u / NKV, kvh = u % NKV;
        bf16* kb = kc_l + ((size_t)b * NKV + kvh) * p.max_ctx * HD;
        bf16* vb = vc_l + ((size_t)b * NKV + kvh) * p.max_ctx * HD;
        decode_attn_unit<MG_NW, 1, MG_SPL_QKV>(attn_sm, wt, p.part_qkv, MG_SPL_QKV, B, b, kvh, Lw.qkv_bias, kb, vb, p.ctx_len[b], p.max_ctx,
                                               p.inv_freq, p.att + (size_t)b * D);
        worker_bar();
      }
      grid_sync();
      gemm_phase(1, p.att, D, p.part_o, D, 0);
      grid_sync();
      reduce_norm(p.part_o, std::integral_constant<int, MG_SPL_O>(), Lw.ln2);
      grid_sync();
      gemm_phase(2, p.xn, D, nullptr, 2 * DFF, 1);
      grid_sync();
      gemm_phase(3, p.ffa, DFF, p.part_down, D, 0);
      grid_sync();
      reduce_norm(p.part_down, std::integral_constant<int, MG_SPL_DOWN>(), Lw.next_gamma);
      if (l + 1 < p.num_layers) grid_sync();
    }
    if (cta == 0 && wt == 0) {
      p.bar[1] = gen0 + (unsigned)G * (unsigned)(7 * p.num_layers - 1);    // generation base of the next launch
      if (stamp) p.tl[tl_i++] = gtime();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)BPAD) : "memory");
  }
}

// dst[i] <- src[i], 16 KB blocks
__global__ void copy_blocks_kernel(const uint8_t* const* __restrict__ src, uint8_t* const* __restrict__ dst, int n) {
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    const uint4* s = reinterpret_cast<const uint4*>(src[i]);
    uint4* d = reinterpret_cast<uint4*>(dst[i]);
    for (int e = threadIdx.x; e < (int)(MG_BLK / 16); e += blockDim.x) d[e] = s[e];
  }
}

template <int BPAD> constexpr size_t mega_smem() { return (size_t)(BPAD == 32 ? 10 : 7) * MG_BLK + (size_t)MG_KCH * BPAD * 128 + 1024; }

}  // namespace

struct LmMega {
  int G = 0;
  int splits[MG_PH] = {0, 0, 0, 0};
  uint8_t* wstream = nullptr;
  unsigned long long* cta_off = nullptr;
  int* cta_bpl = nullptr;
  MegaUnit* units = nullptr;
  MegaLayerDev* layers = nullptr;
  size_t stream_bytes = 0;
};

struct MegaSession {
  float *part_qkv = nullptr, *part_o = nullptr, *part_down = nullptr;
  unsigned* bar = nullptr;
};

void lm_mega_build(cvk_ctx* ctx, LlmModel* m) {
  const int G = ctx->num_sms, L = m->num_layers;
  LmMega* mg = new LmMega();
  mg->G = G;
  // static schedule (identical for every layer): which (tile, K range) of each projection a CTA owns.  The HBM stream is
  // decoupled from the phases by the per-CTA prefetch ring, so the assignment optimises the critical path of a phase
  // (blocks per unit = serial MMAs; splits = partial sums the consumer must add) and only roughly balances bytes per CTA.
  struct Cfg { int tiles, kchunks, cpu /*chunks per unit*/, first; };
  const Cfg cfg[MG_PH] = {
      {QKV_N / MG_BM, D / MG_BK, MG_CPU_QKV, 0},        // qkv: 9 tiles x 7 splits = 63 units
      {D / MG_BM, D / MG_BK, MG_CPU_O, 63},             // o: 7 x 7 = 49 units
      {2 * DFF / MG_BM, D / MG_BK, D / MG_BK, G - 76},  // gate|up (interleaved rows): 76 tiles, whole K (SwiGLU in the epilogue)
      {D / MG_BM, DFF / MG_BK, MG_CPU_DOWN, 0},         // down: 7 tiles x 10 splits = 70 units
  };
  std::vector<MegaUnit> units((size_t)MG_PH * G, MegaUnit{0, 0, 0, 0});
  for (int ph = 0; ph < MG_PH; ++ph) {
    const int splits = ceil_div(cfg[ph].kchunks, cfg[ph].cpu);
    mg->splits[ph] = splits;
    CVK_REQUIRE(ph == 2 || splits == (ph == 0 ? MG_SPL_QKV : ph == 1 ? MG_SPL_O : MG_SPL_DOWN), "lm mega: split constants out of sync");
    CVK_REQUIRE(cfg[ph].tiles * splits <= G, "lm mega: more units than CTAs in a phase");
    for (int t = 0; t < cfg[ph].tiles; ++t)
      for (int s = 0; s < splits; ++s) {
        const int c = ((cfg[ph].first + t * splits + s) % G + G) % G;
        MegaUnit& u = units[(size_t)ph * G + c];
        CVK_REQUIRE(u.nblk == 0, "lm mega: two units on one CTA in a phase");
        u.nt = t;
        u.kc0 = s * cfg[ph].cpu;
        u.nblk = std::min(cfg[ph].cpu, cfg[ph].kchunks - u.kc0);
        u.split = s;
      }
  }
  std::vector<int> bpl(G, 0);
  std::vector<unsigned long long> off(G, 0);
  size_t total_blocks = 0;
  for (int c = 0; c < G; ++c) {
    for (int ph = 0; ph < MG_PH; ++ph) bpl[c] += units[(size_t)ph * G + c].nblk;
    off[c] = (unsigned long long)total_blocks * L * MG_BLK;
    total_blocks += bpl[c];
  }
  mg->stream_bytes = total_blocks * (size_t)L * MG_BLK;
  mg->wstream = (uint8_t*)ctx->dmalloc(mg->stream_bytes);
  // gather list: stream block <- block (tile, chunk) of the pre-tiled weight (gemm_skinny.cu tile_weights_kernel: the exact
  // SWIZZLE_128B shared-memory image of a 128 x 64 K-major operand tile)
  std::vector<const uint8_t*> src;
  std::vector<uint8_t*> dst;
  src.reserve(total_blocks * L);
  dst.reserve(total_blocks * L);
  for (int c = 0; c < G; ++c)
    for (int l = 0; l < L; ++l) {
      const LayerW& w = m->layers[l];
      const ConvW* Ws[MG_PH] = {&w.qkv, &w.o, &w.gate_up_il, &w.down};
      int j0 = 0;
      for (int ph = 0; ph < MG_PH; ++ph) {
        const MegaUnit& u = units[(size_t)ph * G + c];
        if (u.nblk == 0) continue;
        const uint8_t* tw = (const uint8_t*)skinny_tiled_weights(ctx, *Ws[ph]);
        for (int j = 0; j < u.nblk; ++j) {
          src.push_back(tw + ((size_t)u.nt * cfg[ph].kchunks + u.kc0 + j) * MG_BLK);
          dst.push_back(mg->wstream + off[c] + ((size_t)l * bpl[c] + j0 + j) * MG_BLK);
        }
        j0 += u.nblk;
      }
    }
  {
    const uint8_t** dsrc = nullptr;
    uint8_t** ddst = nullptr;
    CVK_CHECK_CUDA(cudaMalloc((void**)&dsrc, src.size() * sizeof(void*)));
    CVK_CHECK_CUDA(cudaMalloc((void**)&ddst, dst.size() * sizeof(void*)));
    CVK_CHECK_CUDA(cudaMemcpy(dsrc, src.data(), src.size() * sizeof(void*), cudaMemcpyHostToDevice));
    CVK_CHECK_CUDA(cudaMemcpy(ddst, dst.data(), dst.size() * sizeof(void*), cudaMemcpyHostToDevice));
    copy_blocks_kernel<<<148 * 8, 256>>>(dsrc, ddst, (int)src.size());
    CVK_LAUNCH_CHECK();
    CVK_CHECK_CUDA(cudaDeviceSynchronize());
    cudaFree(dsrc);
    cudaFree(ddst);
  }
  std::vector<MegaLayerDev> lay(L);
  for (int l = 0; l < L; ++l) {
    lay[l].qkv_bias = m->layers[l].qkv.bias;
    lay[l].ln2 = m->layers[l].ln2;
    lay[l].next_gamma = l + 1 < L ? m->layers[l + 1].ln1 : m->final_norm;
  }
  auto up = [&](const void* h, size_t bytes) {
    void* d = ctx->dmalloc(bytes);
    CVK_CHECK_CUDA(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice));
    return d;
  };
  mg->cta_off = (unsigned long long*)up(off.data(), off.size() * sizeof(unsigned long long));
  mg->cta_bpl = (int*)up(bpl.data(), bpl.size() * sizeof(int));
  mg->units = (MegaUnit*)up(units.data(), units.size() * sizeof(MegaUnit));
  mg->layers = (MegaLayerDev*)up(lay.data(), lay.size() * sizeof(MegaLayerDev));
  CVK_CHECK_CUDA(cudaFuncSetAttribute(lm_mega_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mega_smem<32>()));
  CVK_CHECK_CUDA(cudaFuncSetAttribute(lm_mega_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mega_smem<64>()));
  m->mega = mg;
}

bool lm_mega_usable(cvk_ctx* ctx, const cvk_lm_session* s, int B) {
  return ctx->lm_mega && ctx->llm && ctx->llm->mega && s->kv_dtype == DT_BF16 && ctx->use_tc && B >= 1 && B <= 64 && s->mega_state;
}

void lm_mega_session_init(cvk_ctx* ctx, cvk_lm_session* s) {
  const LmMega* mg = ctx->llm ? ctx->llm->mega : nullptr;
  if (!mg || s->kv_dtype != DT_BF16 || s->max_batch > 64) return;
  MegaSession* ms = new MegaSession();
  auto alloc = [&](size_t bytes) {
    void* p = nullptr;
    CVK_CHECK_CUDA(cudaMalloc(&p, bytes));
    CVK_CHECK_CUDA(cudaMemset(p, 0, bytes));
    s->owned.push_back(p);
    return p;
  };
  ms->part_qkv = (float*)alloc((size_t)mg->splits[0] * s->max_batch * QKV_N * sizeof(float));
  ms->part_o = (float*)alloc((size_t)mg->splits[1] * s->max_batch * D * sizeof(float));
  ms->part_down = (float*)alloc((size_t)mg->splits[3] * s->max_batch * D * sizeof(float));
  ms->bar = (unsigned*)alloc(64);
  s->mega_state = ms;
}

void lm_mega_session_free(cvk_lm_session* s) {
  delete (MegaSession*)s->mega_state;
  s->mega_state = nullptr;
}

void lm_mega_layers(cvk_ctx* ctx, cudaStream_t st, cvk_lm_session* s, int B) {
  const LlmModel* m = ctx->llm;
  const LmMega* mg = m->mega;
  const MegaSession* ms = (const MegaSession*)s->mega_state;
  CVK_REQUIRE(mg && ms && B >= 1 && B <= 64 && m->num_layers <= MG_MAX_LAYERS, "lm mega: not initialised");
  const int bpad = B <= 32 ? 32 : 64;
  MegaParams p;
  p.wstream = mg->wstream; p.cta_off = mg->cta_off; p.cta_bpl = mg->cta_bpl; p.units = mg->units; p.layers = mg->layers;
  p.num_layers = m->num_layers; p.B = B; p.max_ctx = s->max_ctx;
  p.x = s->x; p.xn = (bf16*)s->xn; p.att = (bf16*)s->att; p.ffa = (bf16*)s->ffa;
  p.part_qkv = ms->part_qkv; p.part_o = ms->part_o; p.part_down = ms->part_down;
  p.splits_qkv = mg->splits[0]; p.splits_o = mg->splits[1]; p.splits_down = mg->splits[3];
  p.kcache = (bf16*)s->kcache; p.vcache = (bf16*)s->vcache;
  p.kv_layer_stride = (unsigned long long)s->max_batch * NKV * s->max_ctx * HD;
  p.ctx_len = s->ctx_len; p.inv_freq = m->d_inv_freq;
  p.bar = ms->bar;
  p.tl = ctx->tl ? (long long*)ctx->tl + 2048 : nullptr;   // second half of the chain-timeline buffer
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(mg->G);
  cfg.blockDim = dim3(MG_THREADS);
  cfg.dynamicSmemBytes = bpad == 32 ? mega_smem<32>() : mega_smem<64>();
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;     // all CTAs co-resident: the grid barriers cannot deadlock against other work
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = ctx->mega_coop ? 1 : 0;
  if (bpad == 32) CVK_CHECK_CUDA(cudaLaunchKernelEx(&cfg, lm_mega_kernel<32>, p));
  else CVK_CHECK_CUDA(cudaLaunchKernelEx(&cfg, lm_mega_kernel<64>, p));
  ctx->launches++;
}
