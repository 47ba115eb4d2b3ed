// bf16 conv-GEMM on 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
//   out[r, n] = epilogue( sum_{j<taps} sum_{k<K} A[r + shift0 + j*dil, k] * W[n][j][k] ),  fp32 accumulate
//
// One CTA computes a 128 x BN output tile.  Warp 0 is the TMA producer, warp 1 allocates TMEM and issues the
// MMAs (single elected lane), warps 2-9 run the fused epilogue straight out of TMEM (one accumulator row per
// thread, two warps per 32-row lane quarter splitting the columns).  A convolution tap is just a row-shifted TMA box of the time-major activation matrix; rows outside the
// matrix and the K tail are zero-filled by the TMA unit, gap rows between ragged sequences hold zeros in memory.
// Two CTAs are co-resident per SM (3 stages x 32 KB each) so one CTA's epilogue overlaps the other's main loop.
//
// Reference ops served: every dense Linear / Conv1d of the flow estimator + encoder (flow/decoder.py,
// matcha transformer.py), the LM projections (transformers Qwen2, llm/llm.py:244-251) and the HiFT ResBlock /
// upsampling convolutions (hifigan/generator.py:110-117, 432-443).
#include "common.cuh"

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;   // 64 bf16 = 128 B = one SWIZZLE_128B atom row
constexpr int TC_STAGES = 3;
constexpr int TC_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two warps per TMEM lane quarter, half the columns each)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded spin: a broken pipeline traps (-> CUDA error in the host API) instead of hanging the GPU
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
    if (clock64() - t0 > 4000000000ll) break;   // ~2 s at 2 GHz
  }
  __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory operand descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// start>>4 | LBO(=1, unused for swizzled K-major)<<16 | SBO(8 rows * 128 B = 1024 B)>>4 <<32 | version 1 <<46 | layout 2 <<61
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
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

// 16 consecutive columns of one row through the fused epilogue (vector fast path + scalar tail).
__device__ __forceinline__ void epi_store16(const EpiDev& e, int r, int n0, int N, const float* acc) {
  if (n0 + 16 > N) {
    for (int i = 0; i < 16 && n0 + i < N; ++i) epi_store<true>(e, r, n0 + i, acc[i]);
    return;
  }
  int seq = 0;
  bool valid = true;
  if (e.row2seq) {
    seq = e.row2seq[r];
    valid = seq >= 0;
  }
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = acc[i];
  if (e.bias) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      float4 b = *reinterpret_cast<const float4*>(e.bias + n0 + i);
      v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
    }
  }
  if (e.rowvec && valid) {
    const float* rv = e.rowvec + (size_t)seq * e.rowvec_ld + n0;
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += rv[i];
  }
  if (e.act1 != ACT_NONE) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = apply_act_fast(e.act1, v[i], e.act1_param, e.alpha1 ? e.alpha1[n0 + i] : 1.f);
  }
  if (e.scale != 1.f) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] *= e.scale;
  }
  if (e.resid) {
    const float* rp = e.resid + (size_t)r * e.resid_ld + n0;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      float4 b = *reinterpret_cast<const float4*>(rp + i);
      v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
    }
  }
  if (!valid) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.f;
  }
  size_t o = (size_t)r * e.out_ld + n0;
  if (e.out_dtype == DT_F32) {
    float* op = (float*)e.out + o;
    if (e.accumulate) {
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        float4 b = *reinterpret_cast<const float4*>(op + i);
        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
      }
    }
#pragma unroll
    for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(op + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  } else {
    bf16* op = (bf16*)e.out + o;
    if (e.accumulate) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += __bfloat162float(op[i]);
    }
    __align__(16) bf16 t[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = __float2bfloat16_rn(v[i]);
    *reinterpret_cast<uint4*>(op) = *reinterpret_cast<uint4*>(t);
    *reinterpret_cast<uint4*>(op + 8) = *reinterpret_cast<uint4*>(t + 8);
  }
  if (e.out2) {
    float w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
      w[i] = valid ? apply_act_fast(e.act2, v[i], e.act2_param, e.alpha2 ? e.alpha2[n0 + i] : 1.f) : 0.f;
    size_t o2 = (size_t)r * e.out2_ld + n0;
    if (e.out2_dtype == DT_F32) {
      float* op = (float*)e.out2 + o2;
#pragma unroll
      for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(op + i) = make_float4(w[i], w[i + 1], w[i + 2], w[i + 3]);
    } else {
      bf16* op = (bf16*)e.out2 + o2;
      __align__(16) bf16 t[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) t[i] = __float2bfloat16_rn(w[i]);
      *reinterpret_cast<uint4*>(op) = *reinterpret_cast<uint4*>(t);
      *reinterpret_cast<uint4*>(op + 8) = *reinterpret_cast<uint4*>(t + 8);
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 2)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, int N, int K, int taps,
                    int dil, int shift0, int rowsOut, EpiDev ep) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[TC_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[TC_STAGES];
  __shared__ __align__(8) uint64_t bar_acc;
  __shared__ uint32_t tmem_base_slot;

  constexpr uint32_t A_BYTES = TC_BM * TC_BK * 2;
  constexpr uint32_t B_BYTES = BN * TC_BK * 2;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  // UMMA instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D=f32, A=B=bf16, K-major both,
  // N>>3 at bit 17, M>>4 at bit 24
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B needs 1024-B aligned tiles
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int r0 = blockIdx.y * TC_BM;
  const int kchunks = (K + TC_BK - 1) / TC_BK;
  const int iters = taps * kchunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_acc), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    // TMEM: BN fp32 accumulator columns (power of two >= 32)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % TC_STAGES;
        const uint32_t round = (uint32_t)(it / TC_STAGES);
        mbar_wait(smem_u32(&bar_empty[s]), (round & 1u) ^ 1u);
        const int j = it / kchunks, kc = it - j * kchunks;
        const uint32_t sa = smem_base + s * STAGE_BYTES;
        const uint32_t sb = sa + A_BYTES;
        const uint32_t fb = smem_u32(&bar_full[s]);
        mbar_expect_tx(fb, STAGE_BYTES);
        tma_load_2d(sa, &tmap_a, fb, kc * TC_BK, r0 + shift0 + j * dil);
        tma_load_3d(sb, &tmap_w, fb, kc * TC_BK, j, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % TC_STAGES;
        const uint32_t round = (uint32_t)(it / TC_STAGES);
        mbar_wait(smem_u32(&bar_full[s]), round & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_base + s * STAGE_BYTES;
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          // advance 16 bf16 = 32 B along K inside the 128-B swizzle atom
          umma_bf16(tmem_base, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), IDESC, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(smem_u32(&bar_empty[s]));   // frees the smem stage once these MMAs have read it
      }
      umma_commit(smem_u32(&bar_acc));          // accumulator complete
    }
  } else {
    // epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31 only
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    mbar_wait(smem_u32(&bar_acc), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int r = r0 + q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c = half * (BN / 2); c < (half + 1) * (BN / 2); c += 16) {
      if (n0 + c >= N) break;   // warp-uniform
      float acc[16];
      tmem_ld16(trow + (uint32_t)c, acc);
      if (r < rowsOut) epi_store16(ep, r, n0 + c, N, acc);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode(cvk_ctx* ctx) {
  if (!ctx->encode_tiled) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CVK_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    CVK_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    ctx->encode_tiled = fn;
  }
  return (EncodeTiledFn)ctx->encode_tiled;
}

template <int BN>
void launch_tc(cvk_ctx* ctx, cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tw, const ConvW& W, int rowsOut,
               const EpiDev& e) {
  constexpr size_t smem = (size_t)TC_STAGES * (TC_BM * TC_BK * 2 + BN * TC_BK * 2) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    CVK_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid(ceil_div(W.N, BN), ceil_div(rowsOut, TC_BM));
  conv_gemm_tc_kernel<BN><<<grid, TC_THREADS, smem, st>>>(ta, tw, W.N, W.K, W.taps, W.dil, W.shift0, rowsOut, e);
}

}  // namespace

void conv_gemm_tc(cvk_ctx* ctx, cudaStream_t st, const Mat& A, const ConvW& W, const Epilogue& ep) {
  CVK_REQUIRE(A.dtype == DT_BF16 && W.w16 != nullptr, "conv_gemm_tc: bf16 operands required");
  CVK_REQUIRE(A.cols >= W.K, "conv_gemm_tc: A has fewer columns than K");
  CVK_REQUIRE(W.K % 8 == 0 && A.ld % 8 == 0 && ((uintptr_t)A.p & 15) == 0, "conv_gemm_tc: operands must be 16-byte aligned");
  CVK_REQUIRE(ep.out.p != nullptr && ep.out.cols >= W.N, "conv_gemm_tc: bad output");
  EncodeTiledFn enc = get_encode(ctx);
  const int BN = ctx->tc_bn256 && W.N > 128 ? 256 : (W.N > 64 ? 128 : 64);
  CUtensorMap ta, tw;
  {
    cuuint64_t dims[2] = {(cuuint64_t)W.K, (cuuint64_t)A.rows};
    cuuint64_t strides[1] = {(cuuint64_t)A.ld * 2};
    cuuint32_t box[2] = {TC_BK, TC_BM};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A.p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CVK_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A) failed: " + std::to_string((int)r));
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)W.K, (cuuint64_t)W.taps, (cuuint64_t)W.N};
    cuuint64_t strides[2] = {(cuuint64_t)W.K * 2, (cuuint64_t)W.K * W.taps * 2};
    cuuint32_t box[3] = {TC_BK, 1, (cuuint32_t)BN};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, W.w16, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CVK_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(W) failed: " + std::to_string((int)r));
  }
  EpiDev e = to_dev(ep);
  if (!e.bias) e.bias = W.bias;
  int rowsOut = ep.out.rows;
  CVK_REQUIRE((ep.out.ld * ep.out.esize()) % 16 == 0 && ((uintptr_t)ep.out.p & 15) == 0, "conv_gemm_tc: output rows must be 16-byte aligned");
  CVK_REQUIRE(!ep.resid.p || (ep.resid.ld % 4 == 0 && ((uintptr_t)ep.resid.p & 15) == 0), "conv_gemm_tc: residual alignment");
  CVK_REQUIRE(!ep.out2.p || ((ep.out2.ld * ep.out2.esize()) % 16 == 0 && ((uintptr_t)ep.out2.p & 15) == 0), "conv_gemm_tc: out2 alignment");
  const double flops = 2.0 * rowsOut * (double)W.N * W.K * W.taps;
  const double bytes = (double)rowsOut * W.K * 2 + (double)W.N * W.K * W.taps * 2 + (double)rowsOut * W.N * ep.out.esize();
  ProfScope ps(ctx, st, FAM_GEMM_TC, flops, bytes);
  if (BN == 256) launch_tc<256>(ctx, st, ta, tw, W, rowsOut, e);
  else if (BN == 128) launch_tc<128>(ctx, st, ta, tw, W, rowsOut, e);
  else launch_tc<64>(ctx, st, ta, tw, W, rowsOut, e);
  ctx->launches++;
  CVK_LAUNCH_CHECK();
}
