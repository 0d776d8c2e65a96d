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
// HBM traffic = read v_posed + write verts (+ the A12 operand once: 100 MB for 65536 poses); the weight tiles live in L2.
//
// Warp roles (416 threads): warps 0-7 epilogue (thread = vertex row; warps 0-3 take poses 0-7 of a batch, warps 4-7 poses 8-15,
// two warps per scheduler hide each other's TMEM / shared-memory latency), warps 8-11 operand loaders (cp.async into
// 128B-swizzled K-major tiles, per-lane offsets precomputed), warp 12 TMEM allocator + MMA issuer.  Persistent (see WorkIter).
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
constexpr int OFF_W = 0;                        // 2 stages x (hi, lo): the vertex tile's weights change every item
constexpr int OFF_B = OFF_W + 2 * 2 * W_TILE;   // (hi, lo): the pose batch's transforms stay for all 54 vertex tiles
constexpr int OFF_VP = OFF_B + 2 * B_TILE;      // 2 stages
constexpr int OFF_BAR = OFF_VP + 2 * VP_STAGE;
constexpr int LBS_SMEM = OFF_BAR + 128 + 1024;
constexpr uint32_t LBS_IDESC = (1u << 4) | ((uint32_t)(LN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);     // f16 x f16 -> f32, M=128, N=256
constexpr int TMEM_COLS = 512;                  // 2 accumulators of 256 columns

// Work order.  An item is (pose batch, vertex tile); a run is one pose batch with all its vertex tiles, runs are dealt round-robin
// to the CTAs.  A CTA therefore reads / writes whole 82 KB pose rows front to back in 1536-byte steps (DRAM-page and TLB friendly;
// walking the pose dimension for a fixed vertex tile instead touched a new page every access), keeps the batch's A12 operand in
// shared memory for the whole run, and streams the 54 weight tiles (442 KB in all, L2-resident) through a 2-stage ring.
struct WorkIter {
  int n_batches, n_vt, run, vt;
  __device__ WorkIter(int N, int V, int first_run) {
    n_batches = (N + LP - 1) / LP;
    n_vt = (V + 127) / 128;
    run = first_run;
    vt = 0;
  }
  __device__ bool valid() const { return run < n_batches; }
  __device__ int pb() const { return run; }
  __device__ bool first_of_run() const { return vt == 0; }
  __device__ void next(int stride) {
    if (++vt == n_vt) { vt = 0; run += stride; }
  }
};

__global__ void __launch_bounds__(416, 1)
smpl_lbs_tc_kernel(const __half *__restrict__ w_hi, const __half *__restrict__ w_lo,          // [Vpad, 32]
                   const __half *__restrict__ a_hi, const __half *__restrict__ a_lo,          // [N, 12, 32]
                   const float *__restrict__ v_posed, long long vp_ld, float *__restrict__ verts, int N, int V, int out_mul,
                   int out_off) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar = smem_base + OFF_BAR;
  auto full_bar = [&](int s) { return bar + 8u * s; };            // weight tile of stage s landed (128 loader lanes)
  auto empty_bar = [&](int s) { return bar + 8u * (2 + s); };     // MMAs that read stage s (and the A12 tile) completed
  auto tfull_bar = [&](int s) { return bar + 8u * (4 + s); };     // accumulator s complete
  auto tempty_bar = [&](int s) { return bar + 8u * (6 + s); };    // accumulator s drained (8 epilogue warps)
  const uint32_t w_bar = bar + 8u * 8;                            // A12 tile of the run landed
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + OFF_BAR + 8 * 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_LOAD = 8, W_MMA = 12;                           // warps 0-7 epilogue, 8-11 loaders, 12 MMA + TMEM allocator

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(full_bar(s), 128);
      mbar_init(empty_bar(s), 1);
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 8);
    }
    mbar_init(w_bar, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp >= W_LOAD && warp < W_MMA) {
    // =============================== operand loaders (128 lanes) ===============================
    const int lt = threadIdx.x - W_LOAD * 32;
    // this lane's 6 chunk pairs of an A12 block (16 poses x 12 entries x 4 chunks of 16 B): constant per lane
    uint32_t b_soff[6], b_goff[6], b_pose[6];
#pragma unroll
    for (int it = 0; it < 6; ++it) {
      const int c = it * 128 + lt, g = c >> 2, j = c & 3;
      const int p = g / 12, i = g - p * 12, r = p * 16 + i;       // smem row: 16 per pose
      b_soff[it] = (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
      b_goff[it] = (uint32_t)(g * 32 + j * 8);                     // halves from the start of the batch's contiguous block
      b_pose[it] = (uint32_t)p;
    }
    int q = 0;
    for (WorkIter wi(N, V, blockIdx.x); wi.valid(); wi.next(gridDim.x), ++q) {
      const int s = q & 1;
      const uint32_t u = (uint32_t)(q >> 1);
      mbar_wait(empty_bar(s), (u & 1u) ^ 1u);                       // MMAs of item q-2 have read weight stage s
      if (wi.first_of_run()) {
        // new pose batch: every earlier MMA (they read the A12 tile) must have completed before it is overwritten
        if (q >= 1) mbar_wait(empty_bar((q - 1) & 1), ((uint32_t)((q - 1) >> 1)) & 1u);
        const int p0 = wi.pb() * LP;
        const __half *gh = a_hi + (size_t)p0 * 12 * 32, *gl = a_lo + (size_t)p0 * 12 * 32;
#pragma unroll
        for (int it = 0; it < 6; ++it) {
          const bool ok = p0 + (int)b_pose[it] < N;
          cp_async16(smem_base + OFF_B + b_soff[it], ok ? gh + b_goff[it] : a_hi, ok ? 16u : 0u);
          cp_async16(smem_base + OFF_B + B_TILE + b_soff[it], ok ? gl + b_goff[it] : a_lo, ok ? 16u : 0u);
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(w_bar) : "memory");
      }
      const uint32_t w_s = smem_base + OFF_W + s * 2 * W_TILE;
#pragma unroll
      for (int it = 0; it < 4; ++it) {                             // 128 rows x 4 chunks of 16 B (32 fp16 of K)
        const int c = it * 128 + lt, r = c >> 2, j = c & 3;
        const int v = wi.vt * 128 + r;
        const bool ok = v < V;
        const size_t e = (size_t)(ok ? v : 0) * 32 + j * 8;
        const uint32_t off = (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
        cp_async16(w_s + off, w_hi + e, ok ? 16u : 0u);
        cp_async16(w_s + W_TILE + off, w_lo + e, ok ? 16u : 0u);
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full_bar(s)) : "memory");
    }
    cp_async_commit();
    cp_async_wait<0>();
  } else if (warp == W_MMA) {
    // =============================== MMA issuer (whole warp in lock step, one elected lane issues: see conv_tc.cu) ==========
    {
      uint32_t w_phase = 0;
      int q = 0;
      for (WorkIter wi(N, V, blockIdx.x); wi.valid(); wi.next(gridDim.x), ++q) {
        const int s = q & 1;
        const uint32_t u = (uint32_t)(q >> 1);
        if (wi.first_of_run()) {
          mbar_wait(w_bar, w_phase);
          w_phase ^= 1u;
        }
        mbar_wait(tempty_bar(s), (u & 1u) ^ 1u);                    // accumulator s drained by the epilogue of batch q-2
        mbar_wait(full_bar(s), u & 1u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // cp.async (generic proxy) data -> UMMA (async proxy)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem + (uint32_t)(s * 256);
        const uint32_t w_s = smem_base + OFF_W + s * 2 * W_TILE;
        const uint64_t da_hi = make_smem_desc(w_s), da_lo = make_smem_desc(w_s + W_TILE);
        const uint64_t db_hi = make_smem_desc(smem_base + OFF_B), db_lo = make_smem_desc(smem_base + OFF_B + B_TILE);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {                             // K = 32 fp16 = two UMMA K steps of 32 bytes
            const uint64_t adv = (uint64_t)((k * 32) >> 4);
            umma_f16(acc, da_hi + adv, db_hi + adv, LBS_IDESC, k != 0);
            umma_f16(acc, da_lo + adv, db_hi + adv, LBS_IDESC, 1u);
            umma_f16(acc, da_hi + adv, db_lo + adv, LBS_IDESC, 1u);
          }
          umma_commit(empty_bar(s));
          umma_commit(tfull_bar(s));
        }
        __syncwarp();
      }
    }
  } else {
    // =============================== epilogue: thread = vertex, two groups of 8 poses ===============================
    const int tid = threadIdx.x;                                    // 0..255
    const int t = tid & 127;                                        // TMEM lane = vertex row of the tile
    const int half = tid >> 7;                                      // poses half*8 .. half*8+7 of the batch
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const int vgrp = tid / 96, vj = tid - vgrp * 96;                // v_posed staging: threads 0..191 = 2 groups x 96 chunks of a pose row
    auto issue_vp = [&](const WorkIter &wi, int q) {                // v_posed rows of this item -> staging buffer q & 1
      if (wi.valid() && vgrp < 2) {
        const int vt = wi.vt, pb = wi.pb();
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
    WorkIter pre(N, V, blockIdx.x);                                 // prefetch cursor: two items ahead of the compute cursor
    issue_vp(pre, 0);
    if (pre.valid()) pre.next(gridDim.x);
    issue_vp(pre, 1);
    if (pre.valid()) pre.next(gridDim.x);
    int q = 0;
    for (WorkIter wi(N, V, blockIdx.x); wi.valid(); wi.next(gridDim.x), ++q) {
      const int vt = wi.vt, pb = wi.pb();
      const int s = q & 1;
      const uint32_t u = (uint32_t)(q >> 1);
      cp_async_wait<1>();                                           // this thread's copies of item q have landed (item q+1 may be in flight)
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
      issue_vp(pre, q + 2);
      if (pre.valid()) pre.next(gridDim.x);
    }
    cp_async_wait<0>();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == W_MMA) {
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
  const long long runs = (N + hd::LP - 1) / hd::LP;                 // one run = one batch of 16 poses over all vertex tiles
  const int grid = (int)(runs < num_sms[dev] ? runs : num_sms[dev]);
  hd::smpl_lbs_tc_kernel<<<grid, 416, hd::LBS_SMEM, (cudaStream_t)stream>>>(
      reinterpret_cast<const __half *>(w_hi), reinterpret_cast<const __half *>(w_lo), reinterpret_cast<const __half *>(a12t_hi),
      reinterpret_cast<const __half *>(a12t_lo), v_posed, vp_ld, verts, N, V, out_mul, out_off);
  return hd::check_launch("smpl_lbs_tc_kernel");
}
