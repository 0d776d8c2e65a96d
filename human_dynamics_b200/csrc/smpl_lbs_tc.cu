// Linear-blend skinning with the 6890x24 blend-weight x transform contraction on the tensor cores (sm_100a).
//
//   verts[n, v, :] = (sum_k W[v, k] A[n, k]) [v_posed[n, v]; 1]            (src/tf_smpl/batch_smpl.py:141-151)
//
// The reference tiles W to [N, 6890, 24] and runs a batched matmul; per (pose, vertex) that contraction is 24 x 12 MACs over
// operands that a CUDA-core kernel has to pull from shared memory (192 B per pose-vertex: the smem port, not HBM, bounds it).
// Here it is a GEMM:  T[v, (n, i)] = sum_k W[v, k] * A12[n, k, i],  M = 128 vertices (TMEM lane = vertex), N = 16 poses x 16
// columns (12 entries + 4 unused, so every pose starts on a 16-column boundary) = 256, K = 24 joints (32 with zero padding), fp16 head/remainder split x 3 MMAs (operands are O(1), so the
// remainder needs no scaling and all three products share one accumulator).  Each epilogue thread owns one vertex: per pose
// it reads the 12 entries of T from TMEM, applies them to v_posed (staged by cp.async as 1536-byte row segments) and leaves
// the result in the same staging slot; the CTA then writes 16 x 1536 contiguous bytes with 8-byte vector stores.
// HBM traffic = read v_posed + write verts; W tile and the A12 operand blocks (75 MB for 65536 poses) live in smem / L2.
//
// Warp roles (320 threads): warps 0-7 epilogue (thread = vertex row; warps 0-3 take poses 0-7 of a batch, warps 4-7 poses 8-15,
// two warps per scheduler hide each other's TMEM / shared-memory latency), warp 8 operand loader (cp.async, 128B-swizzled
// K-major tiles), warp 9 TMEM allocator + MMA issuer.  Persistent: CTA c walks a contiguous range of the (vertex tile, pose batch) list.
#include "conv_common.cuh"
#include "tc_ptx.cuh"

namespace hd {
namespace {
using namespace ptx;

constexpr int LP = 16;                          // poses per batch
constexpr int LN = LP * 16;                     // GEMM N = 256: 16 columns per pose, 12 used
constexpr int W_TILE = 128 * 128;               // 128 vertex rows x 128-byte swizzle row (first 64 B = 32 fp16 of K used)
constexpr int B_TILE = LN * 128;                // 256 rows (rows 12..15 of every pose are never written or read back)
constexpr int VP_ROW = 128 * 3 * 4;             // 1536 B: 128 vertices x xyz fp32
constexpr int VP_STAGE = LP * VP_ROW;           // 24 KiB
constexpr int OFF_W = 0;                        // W hi, W lo
constexpr int OFF_B = OFF_W + 2 * W_TILE;       // 2 stages x (hi, lo)
constexpr int OFF_VP = OFF_B + 2 * 2 * B_TILE;  // 2 stages
constexpr int OFF_BAR = OFF_VP + 2 * VP_STAGE;
constexpr int LBS_SMEM = OFF_BAR + 128 + 1024;
constexpr uint32_t LBS_IDESC = (1u << 4) | ((uint32_t)(LN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);     // f16 x f16 -> f32, M=128, N=256
constexpr int TMEM_COLS = 512;                  // 2 accumulators of 256 columns

__global__ void __launch_bounds__(320, 1)
smpl_lbs_tc_kernel(const __half *__restrict__ w_hi, const __half *__restrict__ w_lo,          // [Vpad, 32]
                   const __half *__restrict__ a_hi, const __half *__restrict__ a_lo,          // [N, 12, 32]
                   const float *__restrict__ v_posed, long long vp_ld, float *__restrict__ verts, int N, int V, int out_mul,
                   int out_off) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar = smem_base + OFF_BAR;
  auto full_bar = [&](int s) { return bar + 8u * s; };            // B operands of stage s landed (32 loader lanes)
  auto empty_bar = [&](int s) { return bar + 8u * (2 + s); };     // MMAs that read stage s (and the W tile) completed
  auto tfull_bar = [&](int s) { return bar + 8u * (4 + s); };     // accumulator s complete
  auto tempty_bar = [&](int s) { return bar + 8u * (6 + s); };    // accumulator s drained (4 epilogue warps)
  const uint32_t w_bar = bar + 8u * 8;                            // W tile landed
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + OFF_BAR + 8 * 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_batches = (N + LP - 1) / LP;
  const int n_vt = (V + 127) / 128;
  const long long total = (long long)n_vt * n_batches;
  const long long per = (total + gridDim.x - 1) / gridDim.x;
  const long long w0 = (long long)blockIdx.x * per;
  const long long w1 = (w0 + per < total) ? w0 + per : total;
  const int my = w1 > w0 ? (int)(w1 - w0) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(full_bar(s), 32);
      mbar_init(empty_bar(s), 1);
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 8);
    }
    mbar_init(w_bar, 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // =============================== operand loader ===============================
    int cur_vt = -1;
    for (int q = 0; q < my; ++q) {
      const long long w = w0 + q;
      const int vt = (int)(w / n_batches), pb = (int)(w % n_batches);
      const int s = q & 1;
      const uint32_t u = (uint32_t)(q >> 1);
      mbar_wait(empty_bar(s), (u & 1u) ^ 1u);                       // MMAs of batch q-2 have read stage s
      if (vt != cur_vt) {
        // new vertex tile: every earlier MMA (they read the W tile) must have completed before it is overwritten
        if (q >= 1) mbar_wait(empty_bar((q - 1) & 1), ((uint32_t)((q - 1) >> 1)) & 1u);
        cur_vt = vt;
        for (int c = lane; c < 128 * 4; c += 32) {                 // 128 rows x 4 chunks of 16 B (32 fp16 of K)
          const int r = c >> 2, j = c & 3;
          const int v = vt * 128 + r;
          const bool ok = v < V;
          const size_t e = (size_t)(ok ? v : 0) * 32 + j * 8;
          const uint32_t off = (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
          cp_async16(smem_base + OFF_W + off, w_hi + e, ok ? 16u : 0u);
          cp_async16(smem_base + OFF_W + W_TILE + off, w_lo + e, ok ? 16u : 0u);
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(w_bar) : "memory");
      }
      const uint32_t b_hi = smem_base + OFF_B + s * 2 * B_TILE, b_lo = b_hi + B_TILE;
      const int p0 = pb * LP;
      for (int c = lane; c < LP * 12 * 4; c += 32) {               // 16 poses x 12 entries x 4 chunks of 16 B
        const int g = c >> 2, j = c & 3;                           // g = pose * 12 + entry: row g of the contiguous global block
        const int p = g / 12, i = g - p * 12;
        const bool ok = p0 + p < N;
        const size_t e = ((size_t)(ok ? p0 : 0) * 12 + (ok ? g : 0)) * 32 + j * 8;
        const int r = p * 16 + i;                                  // smem row: 16 per pose
        const uint32_t off = (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
        cp_async16(b_hi + off, a_hi + e, ok ? 16u : 0u);
        cp_async16(b_lo + off, a_lo + e, ok ? 16u : 0u);
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full_bar(s)) : "memory");
    }
    cp_async_commit();
    cp_async_wait<0>();
  } else if (warp == 9) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      int cur_vt = -1;
      uint32_t w_phase = 0;
      for (int q = 0; q < my; ++q) {
        const long long w = w0 + q;
        const int vt = (int)(w / n_batches);
        const int s = q & 1;
        const uint32_t u = (uint32_t)(q >> 1);
        if (vt != cur_vt) {
          cur_vt = vt;
          mbar_wait(w_bar, w_phase);
          w_phase ^= 1u;
        }
        mbar_wait(tempty_bar(s), (u & 1u) ^ 1u);                    // accumulator s drained by the epilogue of batch q-2
        mbar_wait(full_bar(s), u & 1u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // cp.async (generic proxy) data -> UMMA (async proxy)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem + (uint32_t)(s * 256);
        const uint64_t da_hi = make_smem_desc(smem_base + OFF_W), da_lo = make_smem_desc(smem_base + OFF_W + W_TILE);
        const uint32_t b_hi = smem_base + OFF_B + s * 2 * B_TILE;
        const uint64_t db_hi = make_smem_desc(b_hi), db_lo = make_smem_desc(b_hi + B_TILE);
#pragma unroll
        for (int k = 0; k < 2; ++k) {                               // K = 32 fp16 = two UMMA K steps of 32 bytes
          const uint64_t adv = (uint64_t)((k * 32) >> 4);
          umma_f16(acc, da_hi + adv, db_hi + adv, LBS_IDESC, k != 0);
          umma_f16(acc, da_lo + adv, db_hi + adv, LBS_IDESC, 1u);
          umma_f16(acc, da_hi + adv, db_lo + adv, LBS_IDESC, 1u);
        }
        umma_commit(empty_bar(s));
        umma_commit(tfull_bar(s));
      }
    }
  } else {
    // =============================== epilogue: thread = vertex, two groups of 8 poses ===============================
    const int tid = threadIdx.x;                                    // 0..255
    const int t = tid & 127;                                        // TMEM lane = vertex row of the tile
    const int half = tid >> 7;                                      // poses half*8 .. half*8+7 of the batch
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const int vgrp = tid / 96, vj = tid - vgrp * 96;                // v_posed staging: threads 0..191 = 2 groups x 96 chunks of a pose row
    auto issue_vp = [&](int q) {                                    // v_posed rows of batch q -> staging buffer q & 1
      if (q < my && vgrp < 2) {
        const long long w = w0 + q;
        const int vt = (int)(w / n_batches), pb = (int)(w % n_batches);
        const uint32_t dst = smem_base + OFF_VP + (q & 1) * VP_STAGE;
        const int nv = (V - vt * 128) < 128 ? (V - vt * 128) : 128;
        const int chunks = (nv * 3 + 3) / 4;                        // 16-byte chunks per pose row (<= 96; reads stay inside the padded row)
        const float *src0 = v_posed + (size_t)vt * 384 + vj * 4;
#pragma unroll
        for (int i = 0; i < LP / 2; ++i) {
          const int p = vgrp * (LP / 2) + i;
          const int n = pb * LP + p;
          const bool ok = n < N && vj < chunks;
          cp_async16(dst + (uint32_t)(p * VP_ROW + vj * 16), ok ? src0 + (size_t)n * vp_ld : v_posed, ok ? 16u : 0u);
        }
      }
      cp_async_commit();                                            // one (possibly empty) group per call keeps the wait counts uniform
    };
    issue_vp(0);
    issue_vp(1);
    for (int q = 0; q < my; ++q) {
      const long long w = w0 + q;
      const int vt = (int)(w / n_batches), pb = (int)(w % n_batches);
      const int s = q & 1;
      const uint32_t u = (uint32_t)(q >> 1);
      cp_async_wait<1>();                                           // this thread's copies of batch q have landed (batch q+1 may be in flight)
      named_bar_sync(1, 256);                                       // ... and everybody else's
      mbar_wait(tfull_bar(s), u & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float *stage = reinterpret_cast<float *>(smem + OFF_VP + s * VP_STAGE);
      const int np = (N - pb * LP) < LP ? (N - pb * LP) : LP;
#pragma unroll
      for (int g = 0; g < 2; ++g) {                                 // 4 poses per TMEM round trip
        uint32_t T[4][16];
#pragma unroll
        for (int i = 0; i < 4; ++i) tmem_ld16_nowait(tmem + lane_off + (uint32_t)(s * 256 + (half * 8 + g * 4 + i) * 16), T[i]);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float *vp = stage + (half * 8 + g * 4 + i) * 384 + t * 3;
          const float x = vp[0], y = vp[1], z = vp[2];
          const float *Tf = reinterpret_cast<const float *>(T[i]);
          vp[0] = Tf[0] * x + Tf[1] * y + Tf[2] * z + Tf[3];
          vp[1] = Tf[4] * x + Tf[5] * y + Tf[6] * z + Tf[7];
          vp[2] = Tf[8] * x + Tf[9] * y + Tf[10] * z + Tf[11];
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(s));
      named_bar_sync(1, 256);                                       // all result rows of the batch are in the staging buffer
      // coalesced write-out: per pose 128 vertices x 12 B = 1536 contiguous bytes (8-byte aligned for every pose slot)
      const int nv = (V - vt * 128) < 128 ? (V - vt * 128) : 128;
      const int units = (nv * 3) / 2;                               // 8-byte units per pose row
      const bool odd = ((nv * 3) & 1) != 0;
      if (tid < 192) {
        float *dst = verts + ((size_t)(pb * LP) * out_mul + out_off) * ((size_t)V * 3) + (size_t)vt * 384 + 2 * tid;
        const size_t dstep = (size_t)out_mul * ((size_t)V * 3);
        const float *src = stage + 2 * tid;
        for (int p = 0; p < np; ++p) {
          if (tid < units) *reinterpret_cast<float2 *>(dst) = *reinterpret_cast<const float2 *>(src);
          else if (odd && tid == units) dst[0] = src[0];
          dst += dstep; src += 384;
        }
      }
      named_bar_sync(1, 256);                                       // staging buffer s is free again
      issue_vp(q + 2);
    }
    cp_async_wait<0>();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 9) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
  }
}

}  // namespace
}  // namespace hd

extern "C" int hd_smpl_lbs_tc(const void *w_hi, const void *w_lo, const void *a12t_hi, const void *a12t_lo, const float *v_posed,
                              long long vp_ld, float *verts, int N, int V, int out_mul, int out_off, void *stream) {
  HD_REQUIRE(w_hi && w_lo && a12t_hi && a12t_lo && v_posed && verts && N > 0 && V > 0 && out_mul >= 1 && out_off >= 0 && out_off < out_mul &&
                 vp_ld >= (long long)((V * 3 + 3) / 4 * 4) && vp_ld % 4 == 0 && hd::aligned16(w_hi) && hd::aligned16(w_lo) && hd::aligned16(a12t_hi) &&
                 hd::aligned16(a12t_lo) && hd::aligned16(v_posed) && (reinterpret_cast<uintptr_t>(verts) & 7u) == 0 && ((long long)V * 3 * 4) % 8 == 0,
             "hd_smpl_lbs_tc: bad arguments (aligned operands, vp_ld % 4 == 0 and >= roundup4(3V), 8-byte aligned vertex rows)");
  static bool configured[64] = {};
  static int num_sms[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) { hd::set_last_error_text("hd_smpl_lbs_tc: device ordinal out of range"); return HD_ERR_UNSUPPORTED; }
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(hd::smpl_lbs_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, hd::LBS_SMEM);
    if (e != cudaSuccess) { hd::set_last_error("smpl_lbs_tc attr", e); return HD_ERR_CUDA; }
    cudaDeviceGetAttribute(&num_sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (num_sms[dev] <= 0) num_sms[dev] = 148;
    configured[dev] = true;
  }
  const long long total = (long long)((V + 127) / 128) * ((N + hd::LP - 1) / hd::LP);
  const int grid = (int)(total < num_sms[dev] ? total : num_sms[dev]);
  hd::smpl_lbs_tc_kernel<<<grid, 320, hd::LBS_SMEM, (cudaStream_t)stream>>>(
      reinterpret_cast<const __half *>(w_hi), reinterpret_cast<const __half *>(w_lo), reinterpret_cast<const __half *>(a12t_hi),
      reinterpret_cast<const __half *>(a12t_lo), v_posed, vp_ld, verts, N, V, out_mul, out_off);
  return hd::check_launch("smpl_lbs_tc_kernel");
}
