// tcgen05 implicit-GEMM convolution / FC for sm_100a with FP32-class accuracy via 3xTF32 error compensation.
//
//   D[128 x BN] (fp32)  =  sum over 32-wide K chunks of   A_hi*B_hi  +  (A_lo*B_hi + A_hi*B_lo)
//
// A (activations) never exists in HBM in im2col form: eight producer warps gather the 128 x 32 fp32 tile
// (zero padding, stride, per-channel BN/GN affine + ReLU prologue fused; three chunks of loads in flight per
// thread), split every value into its TF32-exact head `hi` (rounded to nearest; low 13 mantissa bits clear --
// all the tensor core reads) and the remainder `lo = x - hi`, and store both straight into the 128-byte-swizzled
// K-major layout the UMMA shared-memory descriptor expects.  B (weights, pre-split offline into hi/lo, K-major)
// arrives by TMA.  One elected thread issues tcgen05.mma.kind::tf32; accumulators live in TMEM.
//
// Accumulation precision: the tensor core TRUNCATES the fp32 accumulator on every MMA (measured: relative
// error ~2.7e-8 per accumulation, growing linearly with K; 1.4e-4 at K=16384), which is not FP32-class.
// So accumulation is two-level: the dominant A_hi*B_hi term is summed in TMEM for only PCH K-chunks
// (4*PCH MMAs) into one of two ping-pong accumulators, which four drain warps then add -- round-to-nearest,
// on the CUDA cores -- into per-thread fp32 running sums while the MMA warp fills the other accumulator.
// The two cross terms (2^-11 smaller) accumulate in their own TMEM region for the whole tile (ping-pong per tile).
//
// Persistent: grid = min(#tiles, #SMs); each CTA walks tiles blockIdx.x, +gridDim.x, ...  Barrier phases run
// continuously across tiles, the producers' prefetch runs across tile boundaries, and the drain warps run the
// fused epilogue (scale/shift, residual, ReLU; coalesced through a small smem transpose) from registers while
// the producers and the MMA warp are already working on the next tile.
//
// Warp roles (512 threads, register budgets re-balanced with setmaxnreg):
//   WG0 warps 0-3   drain + epilogue (TMEM lane quarter = warp id)        200 regs
//   WG1-2 warps 4-11 A producers                                          136 regs
//   WG3 warp 12 TMEM allocator + TMA producer for B, warp 13 MMA issuer    40 regs
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include "conv_common.cuh"
#include "tc_ptx.cuh"

namespace hd {
namespace {

constexpr int BM = 128;
constexpr int A_TILE_BYTES = BM * 128;      // 16 KiB: 128 rows x one 128-byte swizzle row (32 tf32 or 64 fp16 of K)

using namespace ptx;

// TEPI (TMA epilogue; K <= 256, i.e. the HBM-shaped 1x1 layers): one drain group whose threads own accumulator rows and
// only ever touch shared memory; the residual arrives and fp32 / fp16-pair outputs leave as 128-row x 32-column slabs
// moved by TMA (cp.async.bulk.tensor load / store) through a ring of three swizzled staging buffers.
template <int BN, bool HALF, bool DUAL = false, int TEPI = 0>
struct Cfg {
  // DUAL (epilogue-bound layers, K <= 256, one drain group per tile): two drain/epilogue warp groups take alternate tiles;
  // the main loop is short there, so 2 operand stages suffice and pay for the second set of staging tiles.
  // TEPI 1: full staging ring (fp32 residual / output + fp16 pair), 2 operand stages.  TEPI 2: layers that write ONLY the fp16 pair
  // and add no residual (conv1 / conv2 of a bottleneck: long K, tensor-bound): 16 KB slabs, one buffer per drain group, which leaves
  // room for the third operand stage their main loop wants.
  // 64-wide tiles (conv1 of the root, the 64-channel layers of block 1, the few-tile head GEMMs) have 48 KB stages: their main loops are
  // load-LATENCY-bound (per-role counters, profiles/r02_roles_plan_layers.txt: ~1450 cycles per chunk against 393 cycles of tensor
  // time and ~940 of shared-memory port time with 2-3 chunks in flight), so they trade staging buffers for operand stages:
  // TEPI 1: 3 stages + a ring of 2 (215 KB), TEPI 2: 4 stages + its ring of 2 (226 KB).
  static constexpr int STAGES = DUAL ? 2 : (TEPI == 1 ? (BN == 64 ? 3 : 2) : ((TEPI == 2 && BN == 64) ? 4 : 3));
  static constexpr int DW = (DUAL || TEPI) ? 8 : 4;             // drain + epilogue warps (DUAL: two groups on alternate tiles;
                                                                // TEPI: two groups on alternate 32-column slabs of the same tile)
  static constexpr int NUM_THREADS = (DW + 8 + 4) * 32;         // + 8 producer warps + {TMA, MMA, 2 idle}
  static constexpr int BKE = HALF ? 64 : 32;                    // K elements per chunk (one 128-byte row)
  static constexpr int B_TILE_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int SLAB = 32;                               // accumulator columns transposed per epilogue pass
  static constexpr int STG_LD = SLAB + 4;                       // padded row of the per-warp 32 x 32 staging tile (floats)
  static constexpr int STG_OFFSET = BAR_OFFSET + 128;
  static constexpr int NSTG = DUAL ? 2 : 1;                     // DUAL has the smem for double-buffered residual slabs
  static constexpr int STG_BYTES = DW * NSTG * 32 * STG_LD * 4; // 32 x 32 tile(s) per drain warp
  static constexpr int LUT_OFFSET = STG_OFFSET + STG_BYTES;      // GATHER: k -> (ky, kx, offset) table, 256 entries
  // TEPI staging ring: per buffer one fp32 slab [128][32] (128-byte rows, SWIZZLE_128B) + two fp16 slabs [128][32]
  // (64-byte rows, SWIZZLE_64B); 1024-byte aligned
  static constexpr int EPI_NB = (TEPI == 2 || (TEPI == 1 && BN == 64)) ? 2 : 3;
  static constexpr int EPI_F32_BYTES = TEPI == 2 ? 0 : 128 * 128, EPI_H_BYTES = 128 * 64;
  static constexpr int EPI_BUF_BYTES = EPI_F32_BYTES + 2 * EPI_H_BYTES;
  static constexpr int EPI_OFFSET = (BAR_OFFSET + 128 + 1023) / 1024 * 1024;
  static constexpr int SMEM_BYTES = TEPI ? EPI_OFFSET + EPI_NB * EPI_BUF_BYTES + 1024 : LUT_OFFSET + 1024 + 1024;   // + alignment slack
  static_assert(SMEM_BYTES <= 232448, "dynamic shared memory beyond the 227 KB a CTA can opt into");
  static constexpr int TMEM_COLS = 4 * BN;                      // 2 cross-term + 2 ping-pong accumulators (512 / 256)
  // D=f32, A/B K-major: c_format[4,6)=1, a_format[7,10), b_format[10,13) (2 = TF32, 0 = F16), N>>3 [17,23), M>>4 [24,29)
  static constexpr uint32_t FMT = HALF ? 0u : 2u;
  static constexpr uint32_t IDESC = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  static constexpr int PF = HALF ? 2 : 3;                       // producer prefetch ring depth (chunks in flight per thread)
  static constexpr int V = HALF ? 2 : 1;                        // float4 loads per row per chunk per thread
};

struct RowState {       // 4 output rows of one producer thread: image index and top-left input coordinate
  int n[4], iy[4], ix[4];
};

struct EpiMaps {      // TEPI only: activation tensor maps (fp32 residual / output, fp16 head / remainder outputs)
  CUtensorMap res, out, ohi, olo;
};

template <int BN, bool SPLIT, int PCH, bool HALF, bool GATHER, bool ASPLIT, bool DUAL, int TEPI>
__global__ void __launch_bounds__((Cfg<BN, HALF, DUAL, TEPI>::NUM_THREADS), 1)
conv_gemm_tc_kernel(const ConvParams p, const __grid_constant__ CUtensorMap tmap_hi, const __grid_constant__ CUtensorMap tmap_lo,
                    const __grid_constant__ EpiMaps em) {
  using C = Cfg<BN, HALF, DUAL, TEPI>;
  constexpr int BKE = C::BKE, PF = C::PF, V = C::V, STAGES = C::STAGES, DW = C::DW, NUM_THREADS = C::NUM_THREADS;
  constexpr int W_TMA = DW + 8, W_MMA = DW + 9;
  constexpr int W_RES = DW + 10, W_ST = DW + 11;      // TEPI: residual-load agent, store agent (idle warps otherwise)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + C::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto accf_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };        // ping-pong accumulator b ready to drain
  auto acce_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };    // accumulator b drained
  auto smallf_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 4 + b); };  // cross-term accumulator b drained
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + C::BAR_OFFSET + 8 * (2 * STAGES + 6));
  // TEPI staging ring (the gap up to EPI_OFFSET holds these): residual slab landed in buffer b / slab b computed and written by the
  // 128 drain threads / the TMA stores of buffer b have read it
  // The two drain groups use a buffer alternately (its uses u = 0, 1, 2, ... belong to groups g, 1-g, g, ...).  A parity wait is only
  // safe for a waiter that observes every phase of its barrier, so "residual landed" and "buffer free" exist twice per buffer, one
  // barrier per use parity: each is then waited on by ONE group (or by the in-order residual agent) in consecutive phases.
  auto res_bar = [&](int b, int u) { return bar_base + 8u * (2 * STAGES + 7 + b * 2 + (u & 1)); };          // phase u >> 1
  auto ready_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 13 + b); };                            // phase u (store agent, in order)
  auto free_bar = [&](int b, int u) { return bar_base + 8u * (2 * STAGES + 16 + b * 2 + (u & 1)); };        // enables use u >= 1
  auto wait_free = [&](int b, int u) {          // the stores of use u-1 of buffer b have read it
    if (u > 0) mbar_wait(free_bar(b, u), (uint32_t)((u & 1) ? (u >> 1) : (u >> 1) - 1) & 1u);
  };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = (GATHER ? p.K_pad : p.K) / BKE;   // GATHER: ragged Cin (conv1: K=147 zero-padded to 192)
  const int num_g = (num_k + PCH - 1) / PCH;          // drain groups per tile (DUAL: host guarantees 1)
  const int tiles_n = (p.Cout + BN - 1) / BN;
  const int num_tiles = ((p.M + BM - 1) / BM) * tiles_n;
  const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), ASPLIT ? 257 : 9);   // 8 producer warps (ASPLIT: 256 threads, hardware arrive per thread when its cp.asyncs land)
                                                  // + 1 arrive.expect_tx from the TMA lane
      mbar_init(empty_bar(s), 1);    // tcgen05.commit
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accf_bar(b), 1);     // tcgen05.commit
      mbar_init(acce_bar(b), TEPI ? 8 : 4);     // one lane per drain warp that reads the accumulator
      mbar_init(smallf_bar(b), TEPI ? 8 : 4);
    }
    if (TEPI)
      for (int b = 0; b < C::EPI_NB; ++b) {
        mbar_init(res_bar(b, 0), 1);
        mbar_init(res_bar(b, 1), 1);
        mbar_init(ready_bar(b), 128);
        mbar_init(free_bar(b, 0), 1);
        mbar_init(free_bar(b, 1), 1);
      }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_TMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_slot)), "n"(C::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;
  const int xmode = p.dbg ? (int)p.dbg[15] : 0;      // timing experiments (hd_conv_gemm_profile only; results invalid)
  // TMEM columns: [0,BN) cross terms 0, [BN,2BN) cross terms 1, [2BN,3BN) main 0, [3BN,4BN) main 1

  if (warp >= DW && warp < DW + 8) {
    // =============================== A producers (256 threads) ===============================
    if (DUAL || TEPI) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");
    const int t = threadIdx.x - DW * 32;  // 0..255
    const int j = t & 7;                  // 16-byte chunk within the 128-byte K row
    const int rb = t >> 3;                // rows rb + 32*i, i < 4
    const uint32_t sw_off = (uint32_t)((j ^ (rb & 7)) << 4);
    const bool prof = p.dbg != nullptr && blockIdx.x == 0 && t == 0;
    long long t_wait = 0, t_start = prof ? clock64() : 0;
    const int total = my_tiles * num_k;

    auto enter_tile = [&](int ti, RowState &rs) {
      const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
      const int m0 = (tile / tiles_n) * BM;
      const int hw = p.Ho * p.Wo;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + rb + 32 * i;
        if (m < p.M) {
          const int n = m / hw;
          const int r = m - n * hw;
          const int oy = r / p.Wo, ox = r - oy * p.Wo;
          rs.n[i] = n; rs.iy[i] = oy * p.stride - p.pad_t; rs.ix[i] = ox * p.stride - p.pad_l;
        } else {
          rs.n[i] = -1; rs.iy[i] = 0; rs.ix[i] = 0;
        }
      }
    };
    if (ASPLIT) {
      // ---- pre-split fp16 activations: cp.async straight into the swizzled tile, STAGES chunks in flight, no registers ----
      const __half *ihi = reinterpret_cast<const __half *>(p.in_hi);
      const __half *ilo = reinterpret_cast<const __half *>(p.in_lo);
      if (!p.planes && p.KH * p.KW <= 32 && (long long)p.n_img * p.H * p.W * p.in_ld < (1ll << 31)) {
        // Lean loop (these warps run on 32 registers and, on the long-K layers, pace the whole kernel: the general loop below spent
        // ~280 instructions per chunk on two integer divisions, 64-bit addressing and spilled row state).  Per tile: one 32-bit
        // element offset and one tap-validity bit mask per row.  Per chunk: (tap, channel) advance incrementally -- a 64-wide chunk
        // lies inside one tap because Cin % 64 == 0 -- and a row costs a shift, an add and two cp.asyncs.  Taken when the input's
        // element offsets fit 32 bits and the kernel window fits the 32-bit tap mask; anything else runs the general loop.
        const uint32_t off0 = (uint32_t)rb * 128u + sw_off;       // row rb + 32 i of the tile: + i * 4096
        int base[4];
        uint32_t mask[4];
        int kc = 0, ti = 0, tap = 0, ci0 = 0, kx = 0, ky = 0, tap_off = 0;
        for (int q = 0; q < total; ++q) {
          if (kc == 0) {
            const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
            const int m0 = (tile / tiles_n) * BM + rb;
            const int hw = p.Ho * p.Wo;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int m = m0 + 32 * i;
              uint32_t mk = 0;
              int b = 0;
              if (m < p.M) {
                const int n = m / hw;
                const int r = m - n * hw;
                const int oy = r / p.Wo, ox = r - oy * p.Wo;
                const int iy0 = oy * p.stride - p.pad_t, ix0 = ox * p.stride - p.pad_l;
                b = ((n * p.H + iy0) * p.W + ix0) * p.in_ld + j * 8;
                int t = 0;
                for (int yy = 0; yy < p.KH; ++yy)
                  for (int xx = 0; xx < p.KW; ++xx, ++t)
                    if ((unsigned)(iy0 + yy) < (unsigned)p.H && (unsigned)(ix0 + xx) < (unsigned)p.W) mk |= 1u << t;
              }
              base[i] = b;
              mask[i] = mk;
            }
            tap = 0; ci0 = 0; kx = 0; ky = 0; tap_off = 0;
          }
          const int s = q % STAGES;
          const uint32_t ph = (uint32_t)(q / STAGES) & 1u;
          long long tw0 = prof ? clock64() : 0;
          mbar_wait(empty_bar(s), ph ^ 1u);
          if (prof) t_wait += clock64() - tw0;
          const uint32_t a_hi = smem_base + s * C::STAGE_BYTES + off0;
          const int eo = tap_off + ci0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool ok = (mask[i] >> tap) & 1u;
            const int e = ok ? base[i] + eo : 0;
            cp_async16(a_hi + i * 4096, ihi + e, ok ? 16u : 0u);
            cp_async16(a_hi + A_TILE_BYTES + i * 4096, ilo + e, ok ? 16u : 0u);
          }
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full_bar(s)) : "memory");
          ci0 += BKE;
          if (ci0 == p.Cin) {
            ci0 = 0; ++tap;
            if (++kx == p.KW) { kx = 0; ++ky; }
            tap_off = (ky * p.W + kx) * p.in_ld;
          }
          if (++kc == num_k) { kc = 0; ++ti; }
        }
      } else if (p.planes && (long long)p.n_img * p.H * p.W * p.in_ld < (1ll << 31)) {
        // resnet conv1 over the padded RGBX fp16 planes, same lean scheme: Cin = 32 halves = one kernel row of 7(+1) pixels x 4, so a
        // 64-wide chunk holds TWO kernel rows (taps 2 kc and 2 kc + 1); this thread's 16-byte piece is pixels 2*(j&3), 2*(j&3)+1 of row
        // 2 kc + (j >> 2).  KH = 8: row 7 is a phantom (zero weights) and is zero-filled.  No padding tests: the planes are pre-padded.
        const uint32_t off0 = (uint32_t)rb * 128u + sw_off;
        const int jt = j >> 2;
        const int row_step = 2 * p.W * (int)p.in_ld;                // two kernel rows per chunk
        int base[4];
        int kc = 0, ti = 0;
        for (int q = 0; q < total; ++q) {
          if (kc == 0) {
            const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
            const int m0 = (tile / tiles_n) * BM + rb;
            const int hw = p.Ho * p.Wo;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int m = m0 + 32 * i;
              int b = -1;
              if (m < p.M) {
                const int n = m / hw;
                const int r = m - n * hw;
                const int oy = r / p.Wo, ox = r - oy * p.Wo;
                b = ((n * p.H + oy * p.stride + jt) * p.W + ox * p.stride) * (int)p.in_ld + (j & 3) * 8;
              }
              base[i] = b;
            }
          }
          const int s = q % STAGES;
          const uint32_t ph = (uint32_t)(q / STAGES) & 1u;
          long long tw0 = prof ? clock64() : 0;
          mbar_wait(empty_bar(s), ph ^ 1u);
          if (prof) t_wait += clock64() - tw0;
          const uint32_t a_hi = smem_base + s * C::STAGE_BYTES + off0;
          const bool tap_ok = 2 * kc + jt < 7;
          const int eo = kc * row_step;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool ok = tap_ok && base[i] >= 0;
            const int e = ok ? base[i] + eo : 0;
            // neighbouring output pixels read overlapping 64-byte windows (each 16-byte piece 4x): keep them in L1
            cp_async16_ca(a_hi + i * 4096, ihi + e, ok ? 16u : 0u);
            cp_async16_ca(a_hi + A_TILE_BYTES + i * 4096, ilo + e, ok ? 16u : 0u);
          }
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full_bar(s)) : "memory");
          if (++kc == num_k) { kc = 0; ++ti; }
        }
      } else {
      RowState rs;
      int kc = 0, ti = 0;
      for (int q = 0; q < total; ++q) {
        if (kc == 0) enter_tile(ti, rs);
        const int s = q % STAGES;
        const uint32_t ph = (uint32_t)(q / STAGES) & 1u;
        long long tw0 = prof ? clock64() : 0;
        mbar_wait(empty_bar(s), ph ^ 1u);
        if (prof) t_wait += clock64() - tw0;
        const int kb = kc * BKE;
        // planes (resnet conv1 over the padded RGBX fp16 planes, Cin = 32 halves = one kernel row of 7(+1) pixels x 4): a 64-wide
        // chunk holds TWO kernel rows; this thread's 16-byte piece is pixels 2*(j&3), 2*(j&3)+1 of row ky.  KH = 8: row 7 is a
        // phantom (zero weights) and is zero-filled.
        const int tap = p.planes ? 2 * kc + (j >> 2) : kb / p.Cin;
        const int ci = p.planes ? (j & 3) * 8 : kb - tap * p.Cin + j * 8;
        const int ky = tap / p.KW, kx = tap - ky * p.KW;
        const bool tap_ok = !p.planes || tap < 7;
        const uint32_t a_hi = smem_base + s * C::STAGE_BYTES, a_lo = a_hi + A_TILE_BYTES;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int iy = rs.iy[i] + ky, ix = rs.ix[i] + kx;
          const bool ok = tap_ok && rs.n[i] >= 0 && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
          const size_t e = ok ? ((size_t)((size_t)rs.n[i] * p.H + iy) * p.W + ix) * p.in_ld + ci : 0;
          const uint32_t off = (uint32_t)(rb + 32 * i) * 128u + sw_off;
          if (p.planes) {        // conv1: neighbouring output pixels read overlapping 64-byte windows (each 16-byte piece 4x): keep them in L1
            cp_async16_ca(a_hi + off, ihi + e, ok ? 16u : 0u);
            cp_async16_ca(a_lo + off, ilo + e, ok ? 16u : 0u);
          } else {
            cp_async16(a_hi + off, ihi + e, ok ? 16u : 0u);
            cp_async16(a_lo + off, ilo + e, ok ? 16u : 0u);
          }
        }
        // the barrier itself is told to arrive (without a pending-count increment) once this thread's copies have landed:
        // chunks are published the moment their data is in smem, with no producer thread in the loop, so up to STAGES
        // chunks are genuinely in flight.  The MMA thread issues the generic->async proxy fence after its wait.
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full_bar(s)) : "memory");
        if (++kc == num_k) { kc = 0; ++ti; }
      }
      }
      cp_async_commit();
      cp_async_wait<0>();
      if (prof) { p.dbg[0] = clock64() - t_start; p.dbg[1] = t_wait; }
    } else {
    RowState pf_rs, st_rs;                 // prefetch-side and store-side row state (may be one tile apart)
    float4 ring[PF][4 * V];
    uint32_t vmask[PF];
    int pf_kc = 0, pf_ti = 0;              // next chunk to prefetch
    auto prefetch = [&](float4 *dst, uint32_t &vm) {
      if (pf_kc == 0) enter_tile(pf_ti, pf_rs);
      if (GATHER) {
        // Ragged Cin (conv1, 7x7x3): K is laid out ky-major with each kernel row's KW*Cin contiguous input floats padded to a
        // multiple of 8 (21 -> 24), so a thread's 8 consecutive k belong to ONE kernel row and are 8 consecutive floats in
        // memory: one address per (row, chunk), element-wise bounds only for the left/right image edge.
        const int seg = p.KW * p.Cin, segp = (seg + 7) & ~7;
        const int g8 = (pf_kc * BKE + j * (4 * V)) / 8;          // 8-float group index along the padded K
        const int gpr = segp >> 3;                               // groups per kernel row
        const int ky = g8 / gpr, r0 = (g8 - ky * gpr) * 8;
        vm = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int iy = pf_rs.iy[i] + ky;
          const bool okr = pf_rs.n[i] >= 0 && ky < p.KH && iy >= 0 && iy < p.H;
          const int c0 = pf_rs.ix[i] * p.Cin + r0;               // float offset inside the input row (may be < 0 at the left edge)
          const float *src = p.in + ((size_t)((size_t)(okr ? pf_rs.n[i] : 0) * p.H + (okr ? iy : 0)) * p.W) * p.in_ld + c0;
          const int rowlen = p.W * p.Cin;
          float xs[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const bool ok = okr && (r0 + e) < seg && (c0 + e) >= 0 && (c0 + e) < rowlen;
            xs[e] = ok ? __ldg(src + e) : 0.f;
          }
          dst[i * V] = make_float4(xs[0], xs[1], xs[2], xs[3]);
          dst[i * V + V - 1] = make_float4(xs[4], xs[5], xs[6], xs[7]);
          if (pf_rs.n[i] >= 0) vm |= 1u << i;
        }
        if (++pf_kc == num_k) { pf_kc = 0; ++pf_ti; }
        return;
      }
      const int kb = pf_kc * BKE;
      const int tap = kb / p.Cin, ci = kb - tap * p.Cin + j * (4 * V);
      const int ky = tap / p.KW, kx = tap - ky * p.KW;
      vm = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int iy = pf_rs.iy[i] + ky, ix = pf_rs.ix[i] + kx;
        const bool ok = pf_rs.n[i] >= 0 && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
        if (ok && !(xmode & 4)) {
          const float4 *src = reinterpret_cast<const float4 *>(p.in + ((size_t)((size_t)pf_rs.n[i] * p.H + iy) * p.W + ix) * p.in_ld + ci);
#pragma unroll
          for (int v = 0; v < V; ++v) dst[i * V + v] = __ldg(src + v);
          vm |= 1u << i;
        } else {
#pragma unroll
          for (int v = 0; v < V; ++v) dst[i * V + v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (++pf_kc == num_k) { pf_kc = 0; ++pf_ti; }
    };
    int st_kc = 0, st_ti = 0;
    auto consume = [&](int q, float4 *cur, uint32_t vm) {
      if (st_kc == 0) enter_tile(st_ti, st_rs);
      const int s = q % STAGES;
      const uint32_t ph = (uint32_t)(q / STAGES) & 1u;
      long long tw0 = prof ? clock64() : 0;
      mbar_wait(empty_bar(s), ph ^ 1u);
      if (prof) t_wait += clock64() - tw0;
      uint8_t *a_hi = smem + s * C::STAGE_BYTES;
      uint8_t *a_lo = a_hi + A_TILE_BYTES;
      const int pci = p.pre_scale ? (st_kc * BKE) % p.Cin + j * (4 * V) : 0;
      // row by row (keeps the live set small): prologue affine (+ReLU) on real pixels only, then split every value into a
      // head and a remainder that the tensor core reads exactly (zero-mean rounding):
      //   TF32: hi = RN_tf32(x), lo = RN_tf32(x - hi);   FP16: hi = RN_f16(x), lo = RN_f16((x - hi) * 2^11)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float xs[4 * V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float4 x = cur[i * V + v];
          xs[4 * v] = x.x; xs[4 * v + 1] = x.y; xs[4 * v + 2] = x.z; xs[4 * v + 3] = x.w;
        }
        if (p.pre_scale && (vm & (1u << i))) {
          const size_t po = (size_t)(p.pre_img_stride != 0 ? st_rs.n[i] : 0) * p.pre_img_stride + pci;
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const float4 sc = __ldg(reinterpret_cast<const float4 *>(p.pre_scale + po) + v);
            const float4 sh = __ldg(reinterpret_cast<const float4 *>(p.pre_shift + po) + v);
            xs[4 * v] = xs[4 * v] * sc.x + sh.x; xs[4 * v + 1] = xs[4 * v + 1] * sc.y + sh.y;
            xs[4 * v + 2] = xs[4 * v + 2] * sc.z + sh.z; xs[4 * v + 3] = xs[4 * v + 3] * sc.w + sh.w;
          }
          if (p.pre_relu) {
#pragma unroll
            for (int e = 0; e < 4 * V; ++e) xs[e] = fmaxf(xs[e], 0.f);
          }
        }
        uint32_t h[4], l[4];
        if (!HALF) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float hf = rn_tf32(xs[e]);
            h[e] = __float_as_uint(hf);
            l[e] = __float_as_uint(rn_tf32(xs[e] - hf));
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) split_f16x2(xs[2 * e], xs[2 * e + 1], h[e], l[e]);
        }
        const uint32_t off = (uint32_t)(rb + 32 * i) * 128u + sw_off;
        if (xmode & 1) continue;
        *reinterpret_cast<uint4 *>(a_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
        if (SPLIT) *reinterpret_cast<uint4 *>(a_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the UMMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(s));
      if (++st_kc == num_k) { st_kc = 0; ++st_ti; }
    };
    // prime the ring with PF-1 chunks, then: issue chunk q+PF-1, consume chunk q
#pragma unroll
    for (int u = 0; u < PF - 1; ++u)
      if (u < total) prefetch(ring[u], vmask[u]);
    for (int q0 = 0; q0 < total; q0 += PF) {
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        const int q = q0 + u;
        if (q < total) {
          if (q + PF - 1 < total) prefetch(ring[(u + PF - 1) % PF], vmask[(u + PF - 1) % PF]);
          consume(q, ring[u], vmask[u]);
        }
      }
    }
    if (prof) { p.dbg[0] = clock64() - t_start; p.dbg[1] = t_wait; }
    }   // !ASPLIT
  } else if (warp < DW) {
    // =============================== drain + epilogue ===============================
    const int dgroup = warp >> 2, quarter = warp & 3;     // DUAL: group 0 takes even tiles, group 1 odd tiles
    if (DUAL || TEPI) asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const bool prof = p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    long long t_wait = 0, t_epi = 0, t_start = prof ? clock64() : 0;
    float *stg0 = reinterpret_cast<float *>(smem + C::STG_OFFSET) + warp * (C::NSTG * 32 * C::STG_LD);
    const int hw = p.Ho * p.Wo;
    // TEPI: both groups work on every tile; group g owns the 32-column slabs g, g+2, ... (sums[] holds only those columns)
    constexpr int SN = TEPI ? BN / 2 : BN;
    auto sum_col = [&](int i) { return TEPI ? ((i >> 5) * 2 + dgroup) * 32 + (i & 31) : i; };     // accumulator column of sums[i]
    for (int ti = (DUAL ? dgroup : 0); ti < my_tiles; ti += (DUAL ? 2 : 1)) {
      const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      float sums[SN];
#pragma unroll
      for (int i = 0; i < SN; ++i) sums[i] = 0.f;
      if (TEPI) {
      } else if (p.res && p.vec_out) {       // pull the residual rows of this group's NEXT tile towards L2 (a whole tile period of lead
                                      // time; the very first tile is prefetched on entry), so the epilogue's cp.asyncs hit L2
        const int tstep = DUAL ? 2 : 1;
        const int tin = (ti == dgroup) ? ti : ti + tstep;                 // first iteration: this tile
        for (int tn = tin; tn <= ti + tstep && tn < my_tiles; tn += tstep) {
          const int tilen = (int)blockIdx.x + tn * (int)gridDim.x;
          const int m = (tilen / tiles_n) * BM + quarter * 32 + lane, n0n = (tilen % tiles_n) * BN;
          if (m < p.M) {
            size_t rrow = (size_t)m;
            if (!(p.res_stride == 1 && p.res_H == p.Ho && p.res_W == p.Wo)) {
              const int n = m / hw;
              const int rr = m - n * hw;
              const int oy = rr / p.Wo, ox = rr - oy * p.Wo;
              rrow = ((size_t)n * p.res_H + (size_t)oy * p.res_stride) * p.res_W + (size_t)ox * p.res_stride;
            }
            const float *rp = p.res + rrow * p.res_ld + n0n;
            for (int c = 0; c < BN && n0n + c < p.Cout; c += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + c));
          }
        }
      }
      for (int g = 0; g < num_g; ++g) {
        const int G = ti * num_g + g;
        const int b = G & 1;
        long long tw0 = prof ? clock64() : 0;
        mbar_wait(accf_bar(b), (uint32_t)(G >> 1) & 1u);
        if (prof) t_wait += clock64() - tw0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int c0 = 0; c0 < SN; c0 += 16) {       // x16 loads: 16 live temporaries next to the running sums
          uint32_t v[16];
          tmem_ld16(tmem_d + lane_off + (uint32_t)(BN * (2 + b) + sum_col(c0)), v);
#pragma unroll
          for (int i = 0; i < 16; ++i) sums[c0 + i] += __uint_as_float(v[i]);     // round-to-nearest fp32 adds
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(acce_bar(b));
      }
      if (SPLIT) {      // the tile's last accf commit also covers every cross-term MMA of the tile
        const int sb = ti & 1;
#pragma unroll
        for (int c0 = 0; c0 < SN; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_d + lane_off + (uint32_t)(BN * sb + sum_col(c0)), v);
#pragma unroll
          for (int i = 0; i < 16; ++i) sums[c0 + i] += __uint_as_float(v[i]) * (HALF ? (1.0f / 2048.0f) : 1.0f);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(smallf_bar(sb));
      }
      long long te0 = prof ? clock64() : 0;
      if constexpr (!TEPI) {
      // Epilogue (overlaps the next tile's main loop).  Each lane owns one accumulator ROW (TMEM lane) but global memory
      // wants a warp to touch contiguous columns of a row.  Per 32-column slab: the residual slab is fetched by cp.async
      // straight into the warp's padded 32 x 32 smem tile (no registers), every lane adds its 32 accumulator values into its
      // row of the tile, then the warp walks the tile 4 rows x 128 contiguous bytes per access: shift (+ReLU), fp32 store
      // and/or the next layer's pre-activated fp16 head/remainder pair.  The loops are kept rolled on purpose: this code runs
      // once per tile on four warps, so its size (instruction-cache misses) is what it costs.
      const int rq = lane >> 3, c4 = (lane & 7) * 4;       // coalesced phase: row rq + 4*i of the warp's 32, columns c4..c4+3 of the slab
      const bool res_plain = p.res_stride == 1 && p.res_H == p.Ho && p.res_W == p.Wo;
      const bool res_smem = p.res != nullptr && p.vec_out;
      const int mrow0 = m0 + quarter * 32 + rq;             // this lane's first row of a slab; its rows are mrow0 + 4*i
      // residual slab `sl` -> staging buffer `buf` (cp.async, zero-fill outside the tensor); one commit group per slab
      auto fetch_res = [&](int sl, float *buf) {
        const int co = n0 + sl * C::SLAB + c4;
        const bool cvalid = co < p.Cout;
        const uint32_t tdst = smem_u32(buf + rq * C::STG_LD + c4);
        if (res_plain) {                                      // residual row == output row: step a pointer, no index math
          const float *rp = p.res + (size_t)mrow0 * p.res_ld + (cvalid ? co : 0);
          const size_t rstep = 4 * (size_t)p.res_ld;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const bool ok = cvalid && (mrow0 + 4 * i) < p.M;
            cp_async16(tdst + (uint32_t)(i * 4 * C::STG_LD * 4), ok ? (const void *)(rp + i * rstep) : (const void *)p.res,
                       (ok && !(xmode & 16)) ? 16u : 0u);
          }
        } else {
#pragma unroll 2
          for (int i = 0; i < 8; ++i) {
            const int m = mrow0 + 4 * i;
            const bool ok = m < p.M && cvalid;
            size_t rrow = 0;
            if (ok) {
              const int n = m / hw;
              const int rr = m - n * hw;
              const int oy = rr / p.Wo, ox = rr - oy * p.Wo;
              rrow = ((size_t)n * p.res_H + (size_t)oy * p.res_stride) * p.res_W + (size_t)ox * p.res_stride;
            }
            cp_async16(tdst + (uint32_t)(i * 4 * C::STG_LD * 4), p.res + rrow * p.res_ld + (ok ? co : 0), (ok && !(xmode & 16)) ? 16u : 0u);
          }
        }
        cp_async_commit();
      };
      constexpr int NSL = BN / C::SLAB;
      __syncwarp();
      if (res_smem && C::NSTG == 2) fetch_res(0, stg0);       // double-buffered: slab sl+1 is in flight while slab sl is processed
#pragma unroll
      for (int sl = 0; sl < NSL; ++sl) {
        const int co = n0 + sl * C::SLAB + c4;
        const bool cvalid = co < p.Cout;
        float *stg = stg0 + (C::NSTG == 2 ? (sl & 1) * (32 * C::STG_LD) : 0);
        __syncwarp();
        if (res_smem) {
          if (C::NSTG == 2) {
            if (sl + 1 < NSL) { fetch_res(sl + 1, stg0 + ((sl + 1) & 1) * (32 * C::STG_LD)); cp_async_wait<1>(); }
            else cp_async_wait<0>();
          } else {
            fetch_res(sl, stg);
            cp_async_wait<0>();
          }
          __syncwarp();
        }
        float4 psc = make_float4(1.f, 1.f, 1.f, 1.f), psh = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 qsc = psc, qsh = psh;
        if (p.vec_out && cvalid) {
          if (p.post_scale && !res_smem) psc = __ldg(reinterpret_cast<const float4 *>(p.post_scale + co));
          if (p.post_shift) psh = __ldg(reinterpret_cast<const float4 *>(p.post_shift + co));
          if (p.out_hi && p.post2_scale) qsc = __ldg(reinterpret_cast<const float4 *>(p.post2_scale + co));
          if (p.out_hi && p.post2_shift) qsh = __ldg(reinterpret_cast<const float4 *>(p.post2_shift + co));
        }
        // thread = row: accumulator slab (+ residual already sitting in the tile) -> tile.  sums[] needs compile-time indices.
#pragma unroll
        for (int c = 0; c < C::SLAB; c += 4) {
          float4 v = make_float4(sums[sl * C::SLAB + c], sums[sl * C::SLAB + c + 1], sums[sl * C::SLAB + c + 2], sums[sl * C::SLAB + c + 3]);
          float4 *tp = reinterpret_cast<float4 *>(stg + lane * C::STG_LD + c);
          if (res_smem) {
            if (p.post_scale) {      // (rare: scale and residual together) scale before the residual add, broadcast loads
              const int cc = n0 + sl * C::SLAB + c;
              if (cc < p.Cout) {
                const float4 sc = __ldg(reinterpret_cast<const float4 *>(p.post_scale + cc));
                v.x *= sc.x; v.y *= sc.y; v.z *= sc.z; v.w *= sc.w;
              }
            }
            const float4 r = *tp;
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
          }
          *tp = v;
        }
        __syncwarp();
        const float *tsrc = stg + rq * C::STG_LD + c4;
        float *op = p.out ? p.out + (size_t)mrow0 * p.out_ld + co : nullptr;
        __half *ohp = p.out_hi ? reinterpret_cast<__half *>(p.out_hi) + (size_t)mrow0 * p.out2_ld + co : nullptr;
        __half *olp = p.out_hi ? reinterpret_cast<__half *>(p.out_lo) + (size_t)mrow0 * p.out2_ld + co : nullptr;
        const size_t ostep = 4 * (size_t)p.out_ld, hstep = 4 * (size_t)p.out2_ld;
        const bool st32 = op != nullptr && !(xmode & 8), st16 = ohp != nullptr && !(xmode & 32);
#pragma unroll 4
        for (int i = 0; i < 8; ++i) {      // 4 independent rows in flight per lane (8 measured slower: code size); latency-, not issue-bound
          const int m = mrow0 + 4 * i;
          if (m >= p.M || !cvalid) continue;
          const float4 a = *reinterpret_cast<const float4 *>(tsrc + i * 4 * C::STG_LD);
          if (p.vec_out) {
            float x0 = a.x * psc.x + psh.x, x1 = a.y * psc.y + psh.y, x2 = a.z * psc.z + psh.z, x3 = a.w * psc.w + psh.w;
            if (p.post_relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f); }
            if (st32) *reinterpret_cast<float4 *>(op + i * ostep) = make_float4(x0, x1, x2, x3);
            if (st16) {          // the next layer's A operand: second affine (+ReLU) = its pre-activation, split into fp16 head/remainder
              float y0 = x0 * qsc.x + qsh.x, y1 = x1 * qsc.y + qsh.y, y2 = x2 * qsc.z + qsh.z, y3 = x3 * qsc.w + qsh.w;
              if (p.post2_relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); y2 = fmaxf(y2, 0.f); y3 = fmaxf(y3, 0.f); }
              uint32_t h0, h1, l0, l1;
              split_f16x2(y0, y1, h0, l0);
              split_f16x2(y2, y3, h1, l1);
              *reinterpret_cast<uint2 *>(ohp + i * hstep) = make_uint2(h0, h1);
              *reinterpret_cast<uint2 *>(olp + i * hstep) = make_uint2(l0, l1);
            }
          } else {             // ragged / unaligned outputs (IEF 85- and 72-wide heads): element-wise
            const float av[4] = {a.x, a.y, a.z, a.w};
            size_t res_row = 0;
            if (p.res) {
              const int n = m / hw;
              const int rr = m - n * hw;
              const int oy = rr / p.Wo, ox = rr - oy * p.Wo;
              res_row = ((size_t)n * p.res_H + (size_t)oy * p.res_stride) * p.res_W + (size_t)ox * p.res_stride;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (co + e >= p.Cout) continue;
              float y = av[e];
              if (p.post_scale) y *= __ldg(p.post_scale + co + e);
              if (p.post_shift) y += __ldg(p.post_shift + co + e);
              if (p.res) y += p.res[res_row * p.res_ld + co + e];
              if (p.post_relu) y = fmaxf(y, 0.f);
              p.out[(size_t)m * p.out_ld + co + e] = y;
            }
          }
        }
      }
      } else {
        // ---- TMA epilogue: thread = accumulator row (TMEM lane); per 32-column slab: wait for the residual slab (TMA load by the
        // residual agent, up to two slabs ahead), v = acc*scale + shift (+ residual) (ReLU) in place in the swizzled staging
        // buffer, the next layer's pre-activated fp16 head/remainder beside it, then arrive on the slab's `ready` barrier: the store
        // agent hands the three slabs to TMA stores and recycles the buffer.  No thread of this group ever waits for a store. ----
        constexpr int NSL = BN / 32, NB = C::EPI_NB;
        const int row = quarter * 32 + lane;
        const uint32_t sw128 = (uint32_t)(row & 7), sw64 = (uint32_t)((row >> 1) & 3);
        const bool has_res = p.res != nullptr, has_o32 = p.out != nullptr && !p.out_sub, has_o16 = p.out_hi != nullptr;
        // out_sub: this thread's row goes to the dense subsampled fp32 output iff its pixel is (oy % s == 0, ox % s == 0); one pair of
        // divisions per tile, then eight 16-byte stores per slab straight from the registers (a quarter of the threads, 128 contiguous
        // bytes each) -- the fp32 slab is neither staged for nor stored by TMA
        float *sub_row = nullptr;
        if (p.out_sub) {
          const int m = m0 + row;
          if (m < p.M) {
            const int n = m / hw, r = m - n * hw;
            const int oy = r / p.Wo, ox = r - oy * p.Wo;
            if (oy % p.out_sub == 0 && ox % p.out_sub == 0) {
              const int Hs = (p.Ho + p.out_sub - 1) / p.out_sub, Ws = (p.Wo + p.out_sub - 1) / p.out_sub;
              sub_row = p.out + ((size_t)((size_t)n * Hs + oy / p.out_sub) * Ws + ox / p.out_sub) * p.out_ld + n0;
            }
          }
        }
#pragma unroll
        for (int s2 = 0; s2 < NSL / 2; ++s2) {
          const int sl = s2 * 2 + dgroup;               // this group's slabs: dgroup, dgroup + 2
          const int gi = ti * NSL + sl;                 // running slab index of this CTA (the agents walk them in this order)
          const int b = gi % NB;
          uint8_t *buf = smem + C::EPI_OFFSET + b * C::EPI_BUF_BYTES;
          const int use = gi / NB;
          if (has_res) mbar_wait(res_bar(b, use), (uint32_t)(use >> 1) & 1u);        // residual landed (=> buffer was free)
          else wait_free(b, use);                                                    // stores of slab gi - NB have read the buffer
          uint8_t *frow = buf + row * 128;
          uint8_t *hrow = buf + C::EPI_F32_BYTES + row * 64;
          uint8_t *lrow = hrow + C::EPI_H_BYTES;
          const int cb = n0 + sl * 32;
          const bool slab_ok = cb < p.Cout;             // ragged last N tile (Cout % 32 == 0): whole slabs beyond Cout are skipped
#pragma unroll
          for (int h = 0; h < 4; ++h) {                 // 8 columns: two fp32 chunks, one fp16 chunk
            if (!slab_ok) break;
            float y[8];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int c = 2 * h + q;
              float4 v = make_float4(sums[s2 * 32 + 4 * c], sums[s2 * 32 + 4 * c + 1], sums[s2 * 32 + 4 * c + 2], sums[s2 * 32 + 4 * c + 3]);
              if (p.post_scale) {
                const float4 s = __ldg(reinterpret_cast<const float4 *>(p.post_scale + cb + 4 * c));
                v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w;
              }
              if (p.post_shift) {
                const float4 s = __ldg(reinterpret_cast<const float4 *>(p.post_shift + cb + 4 * c));
                v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
              }
              float4 *sp = reinterpret_cast<float4 *>(frow + (((uint32_t)c ^ sw128) << 4));
              if (has_res) {
                const float4 r = *sp;
                v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
              }
              if (p.post_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
              if (has_o32) *sp = v;
              if (sub_row) *reinterpret_cast<float4 *>(sub_row + sl * 32 + 4 * c) = v;
              y[4 * q] = v.x; y[4 * q + 1] = v.y; y[4 * q + 2] = v.z; y[4 * q + 3] = v.w;
            }
            if (has_o16) {       // the next layer's A operand: second affine (+ReLU) = its pre-activation, split into fp16 head/remainder
              if (p.post2_scale) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                  const float4 s = __ldg(reinterpret_cast<const float4 *>(p.post2_scale + cb + 8 * h + 4 * q));
                  y[4 * q] *= s.x; y[4 * q + 1] *= s.y; y[4 * q + 2] *= s.z; y[4 * q + 3] *= s.w;
                }
              }
              if (p.post2_shift) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                  const float4 s = __ldg(reinterpret_cast<const float4 *>(p.post2_shift + cb + 8 * h + 4 * q));
                  y[4 * q] += s.x; y[4 * q + 1] += s.y; y[4 * q + 2] += s.z; y[4 * q + 3] += s.w;
                }
              }
              uint32_t hh[4], ll[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float a0 = y[2 * e], a1 = y[2 * e + 1];
                if (p.post2_relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
                split_f16x2(a0, a1, hh[e], ll[e]);
              }
              *reinterpret_cast<uint4 *>(hrow + (((uint32_t)h ^ sw64) << 4)) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
              *reinterpret_cast<uint4 *>(lrow + (((uint32_t)h ^ sw64) << 4)) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the TMA store
          mbar_arrive(ready_bar(b));
        }
      }
      if (prof) t_epi += clock64() - te0;
    }
    if (prof) { p.dbg[2] = clock64() - t_start; p.dbg[3] = t_wait; p.dbg[4] = t_epi; }
  } else {
    // register pool = threads x launch allocation (512 x 128, or 640 x 96 for DUAL); the budgets below must fit in it or
    // setmaxnreg.inc never returns:  4-warp: 128*200 + 256*136 + 128*40 = 65536;  DUAL: 256*192 + 256*32 + 128*24 = 60416 <= 61440
    if (DUAL || TEPI) asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == W_TMA) {
      // =============================== B producer (TMA) ===============================
      if (lane == 0) {
        const bool prof = p.dbg != nullptr && blockIdx.x == 0;
        long long t_wait = 0, t_start = prof ? clock64() : 0;
        int q = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
          const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
          const int n0 = (tile % tiles_n) * BN;
          for (int kc = 0; kc < num_k; ++kc, ++q) {
            const int s = q % STAGES;
            const uint32_t ph = (uint32_t)(q / STAGES) & 1u;
            long long tw0 = prof ? clock64() : 0;
            mbar_wait(empty_bar(s), ph ^ 1u);
            if (prof) t_wait += clock64() - tw0;
            const uint32_t b_hi = smem_base + s * C::STAGE_BYTES + 2 * A_TILE_BYTES;
            if (xmode & 2) { mbar_arrive(full_bar(s)); continue; }
            mbar_arrive_expect_tx(full_bar(s), SPLIT ? 2 * C::B_TILE_BYTES : C::B_TILE_BYTES);
            tma_load_2d(b_hi, &tmap_hi, full_bar(s), kc * BKE, n0);
            if (SPLIT) tma_load_2d(b_hi + C::B_TILE_BYTES, &tmap_lo, full_bar(s), kc * BKE, n0);
          }
        }
        if (prof) { p.dbg[8] = clock64() - t_start; p.dbg[9] = t_wait; }
      }
    } else if (TEPI && warp == W_RES) {
      // =============================== TEPI: residual-load agent ===============================
      if (lane == 0 && p.res) {
        constexpr int NSL = BN / 32, NB = C::EPI_NB;
        const int total = my_tiles * NSL;
        for (int j = 0; j < total; ++j) {
          const int tj = j / NSL, sl = j - tj * NSL;
          const int tile = (int)blockIdx.x + tj * (int)gridDim.x;
          const int c0 = (tile % tiles_n) * BN + sl * 32, c1 = (tile / tiles_n) * BM;
          const int b = j % NB;
          if (tj + 1 < my_tiles) {                 // the same slab of the NEXT tile -> L2, a whole tile period ahead
            const int tilen = tile + (int)gridDim.x;
            tma_prefetch_2d(&em.res, (tilen % tiles_n) * BN + sl * 32, (tilen / tiles_n) * BM);
          }
          const int use = j / NB;
          wait_free(b, use);
          mbar_arrive_expect_tx(res_bar(b, use), C::EPI_F32_BYTES);
          tma_load_2d(smem_base + C::EPI_OFFSET + b * C::EPI_BUF_BYTES, &em.res, res_bar(b, use), c0, c1);
        }
      }
    } else if (TEPI && warp == W_ST) {
      // =============================== TEPI: store agent ===============================
      if (lane == 0) {
        constexpr int NSL = BN / 32, NB = C::EPI_NB;
        const int total = my_tiles * NSL;
        for (int j = 0; j < total; ++j) {
          const int tj = j / NSL, sl = j - tj * NSL;
          const int tile = (int)blockIdx.x + tj * (int)gridDim.x;
          const int c0 = (tile % tiles_n) * BN + sl * 32, c1 = (tile / tiles_n) * BM;
          const int b = j % NB;
          mbar_wait(ready_bar(b), (uint32_t)(j / NB) & 1u);
          const uint32_t sb = smem_base + C::EPI_OFFSET + b * C::EPI_BUF_BYTES;
          if (c0 < p.Cout) {
            if (p.out && !p.out_sub) tma_store_2d(&em.out, sb, c0, c1);
            if (p.out_hi) {
              tma_store_2d(&em.ohi, sb + C::EPI_F32_BYTES, c0, c1);
              tma_store_2d(&em.olo, sb + C::EPI_F32_BYTES + C::EPI_H_BYTES, c0, c1);
            }
          }
          bulk_commit();
          bulk_wait_read<0>();                     // only this agent waits for the TMA engine; then the buffer goes back into the ring
          mbar_arrive(free_bar(b, j / NB + 1));    // enables the buffer's next use
        }
      }
    } else if (warp == W_MMA) {
      // =============================== MMA issuer ===============================
      // The whole warp walks the loop in lock step (all lanes poll the barriers, all values are warp-uniform) and ONE elected lane issues
      // the MMAs and commits of a chunk: under a divergent `lane == 0` the compiler wraps every tcgen05 instruction in an
      // ELECT / BRA.U.ANY loop (~55 issue cycles per 64-cycle MMA, measured: the issue loop, not the tensor pipe, paced the kernel).
      {
        const bool prof = p.dbg != nullptr && blockIdx.x == 0;
        long long t_wfull = 0, t_wacc = 0, t_start = prof ? clock64() : 0;
        int q = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
          const int sb = ti & 1;
          const uint32_t tmem_small = tmem_d + (uint32_t)(BN * sb);
          if (SPLIT) {
            long long tw0 = prof ? clock64() : 0;
            mbar_wait(smallf_bar(sb), ((uint32_t)(ti >> 1) & 1u) ^ 1u);     // cross-term accumulator of tile ti-2 drained
            if (prof) t_wacc += clock64() - tw0;
          }
          for (int kc = 0; kc < num_k; ++kc, ++q) {
            const int s = q % STAGES;
            const uint32_t ph = (uint32_t)(q / STAGES) & 1u;
            const int G = ti * num_g + kc / PCH, b = G & 1;
            const bool group_start = (kc % PCH) == 0;
            long long tw0 = prof ? clock64() : 0;
            if (group_start) mbar_wait(acce_bar(b), ((uint32_t)(G >> 1) & 1u) ^ 1u);    // accumulator b drained
            long long tw1 = prof ? clock64() : 0;
            mbar_wait(full_bar(s), ph);
            if (prof) { t_wacc += tw1 - tw0; t_wfull += clock64() - tw1; }
            if (ASPLIT) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async (generic proxy) data -> UMMA (async proxy)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_big = tmem_d + (uint32_t)(BN * (2 + b));
            const uint32_t a_hi = smem_base + s * C::STAGE_BYTES;
            const uint32_t a_lo = a_hi + A_TILE_BYTES;
            const uint32_t b_hi = a_hi + 2 * A_TILE_BYTES;
            const uint32_t b_lo = b_hi + C::B_TILE_BYTES;
            const uint64_t da_hi = make_smem_desc(a_hi), da_lo = make_smem_desc(a_lo);
            const uint64_t db_hi = make_smem_desc(b_hi), db_lo = make_smem_desc(b_lo);
            if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {                // UMMA K = 8 tf32 / 16 fp16 = 32 bytes: advance inside the swizzle atom
              const uint64_t adv = (uint64_t)((k * 32) >> 4);
              if (HALF) {
                umma_f16(tmem_small, da_lo + adv, db_hi + adv, C::IDESC, (kc | k) != 0);
                umma_f16(tmem_small, da_hi + adv, db_lo + adv, C::IDESC, 1u);
                umma_f16(tmem_big, da_hi + adv, db_hi + adv, C::IDESC, !(group_start && k == 0));
              } else {
                if (SPLIT) {
                  umma_tf32(tmem_small, da_lo + adv, db_hi + adv, C::IDESC, (kc | k) != 0);
                  umma_tf32(tmem_small, da_hi + adv, db_lo + adv, C::IDESC, 1u);
                }
                umma_tf32(tmem_big, da_hi + adv, db_hi + adv, C::IDESC, !(group_start && k == 0));
              }
            }
            umma_commit(empty_bar(s));                   // frees the stage once these MMAs have read it
            if ((kc % PCH) == PCH - 1 || kc == num_k - 1) umma_commit(accf_bar(b));     // hand accumulator b to the drain warps
            }
            __syncwarp();
          }
        }
        if (prof && lane == 0) { p.dbg[5] = clock64() - t_start; p.dbg[6] = t_wfull; p.dbg[7] = t_wacc; }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == W_TMA) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(C::TMEM_COLS));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(ptr);
    else (void)cudaGetLastError();
  }
  return fn;
}

constexpr int kMaxDevices = 64;

template <int BN, bool SPLIT, int PCH, bool HALF, bool GATHER = false, bool ASPLIT = false, bool DUAL = false, int TEPI = 0>
int launch_tc(const ConvParams &p, const hd_conv_desc *d, cudaStream_t st) {
  using C = Cfg<BN, HALF, DUAL, TEPI>;
  // function attributes and the SM count are per device: a process may drive several GPUs through this library
  static bool configured[kMaxDevices] = {};
  static int num_sms[kMaxDevices] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDevices) { set_last_error_text("conv_gemm_tc: device ordinal out of range"); return HD_ERR_UNSUPPORTED; }
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_tc_kernel<BN, SPLIT, PCH, HALF, GATHER, ASPLIT, DUAL, TEPI>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) { set_last_error("conv_gemm_tc attr", e); return HD_ERR_CUDA; }
    cudaDeviceGetAttribute(&num_sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (num_sms[dev] <= 0) num_sms[dev] = 148;
    configured[dev] = true;
  }
  alignas(64) CUtensorMap thi, tlo;
  // weight maps whose box matches this tile width: the 64-row variants serve wide layers run with 64-wide tiles (few-tile GEMMs)
  const bool n64 = BN == 64 && p.Cout > 64;
  const void *mhi = n64 ? d->tmap_hi_n64 : d->tmap_hi, *mlo = n64 ? d->tmap_lo_n64 : d->tmap_lo;
  memcpy(&thi, mhi, sizeof(CUtensorMap));
  memcpy(&tlo, mlo ? mlo : mhi, sizeof(CUtensorMap));
  alignas(64) EpiMaps em;
  if (TEPI) {      // absent maps are never dereferenced by the kernel (guarded by the same null tests on p.res / p.out / p.out_hi)
    const void *any = d->tmap_out ? d->tmap_out : d->tmap_out_hi;
    memcpy(&em.res, d->tmap_res ? d->tmap_res : any, sizeof(CUtensorMap));
    memcpy(&em.out, d->tmap_out ? d->tmap_out : any, sizeof(CUtensorMap));
    memcpy(&em.ohi, d->tmap_out_hi ? d->tmap_out_hi : any, sizeof(CUtensorMap));
    memcpy(&em.olo, d->tmap_out_lo ? d->tmap_out_lo : any, sizeof(CUtensorMap));
  } else {
    memcpy(&em.res, &thi, sizeof(CUtensorMap)); memcpy(&em.out, &thi, sizeof(CUtensorMap));
    memcpy(&em.ohi, &thi, sizeof(CUtensorMap)); memcpy(&em.olo, &thi, sizeof(CUtensorMap));
  }
  const int num_tiles = ceil_div(p.M, BM) * ceil_div(p.Cout, BN);
  dim3 grid(num_tiles < num_sms[dev] ? num_tiles : num_sms[dev]);     // persistent: one CTA per SM walks the tile list
  conv_gemm_tc_kernel<BN, SPLIT, PCH, HALF, GATHER, ASPLIT, DUAL, TEPI><<<grid, C::NUM_THREADS, C::SMEM_BYTES, st>>>(p, thi, tlo, em);
  return check_launch("conv_gemm_tc_kernel");
}

}  // namespace

int launch_conv_tc(const ConvParams &p, const hd_conv_desc *d, cudaStream_t st) {
  if (!d->tmap_hi || (d->impl != HD_IMPL_TC_1XTF32 && !d->tmap_lo)) {
    set_last_error_text("hd_conv_gemm(tc): missing tensor maps");
    return HD_ERR_INVALID;
  }
  const bool half = d->impl == HD_IMPL_TC_3XF16;
  const int bke = half ? 64 : 32;
  if (p.out_sub && (!p.in_hi || p.planes)) {
    set_last_error_text("hd_conv_gemm: out_subsample is only implemented by the TMA epilogue (pre-split input, Cout % 32 == 0, activation maps given)");
    return HD_ERR_UNSUPPORTED;
  }
  if ((p.out_hi || p.in_hi) && (!half || !p.vec_out)) {
    set_last_error_text("hd_conv_gemm(tc): pre-split activations need impl 3 and 16-byte aligned, 4-column-multiple outputs");
    return HD_ERR_INVALID;
  }
  if (p.in_hi && p.planes) {           // resnet conv1 over padded RGBX fp16 planes (hd_pack_conv1_planes)
    if (p.Cin != 32 || p.KH != 8 || p.KW != 1 || p.stride != 2 || p.pad_t != 0 || p.pad_l != 0 || p.in_ld != 4 || p.Cout > 64 ||
        p.W % 2 != 0 || 2 * (p.Wo - 1) + 8 > p.W || 2 * (p.Ho - 1) + 7 > p.H || !aligned16(p.in_hi) || !aligned16(p.in_lo) || p.pre_scale) {
      set_last_error_text("hd_conv_gemm(tc planes): needs the conv1 plane geometry (Cin 32, KH 8, KW 1, stride 2, in_ld 4, even W)");
      return HD_ERR_INVALID;
    }
    const bool maps = p.out && d->tmap_out && !p.res && !p.out_hi;
    if (maps && p.Cout % 32 == 0 && !(d->flags & HD_CONV_NO_TMA_EPILOGUE)) return launch_tc<64, true, 4, true, false, true, false, 1>(p, d, st);
    return launch_tc<64, true, 4, true, false, true, true>(p, d, st);
  }
  if (p.in_hi) {                       // pre-split fp16 activations: cp.async producer
    if (p.Cin % 64 != 0 || p.K % 64 != 0 || p.in_ld % 8 != 0 || !aligned16(p.in_hi) || !aligned16(p.in_lo) || p.pre_scale) {
      set_last_error_text("hd_conv_gemm(tc split-A): needs Cin % 64 == 0, in_ld % 8 == 0, aligned in_hi/in_lo, no prologue");
      return HD_ERR_INVALID;
    }
    {
      // TMA epilogue (8 drain warps, staging ring, store / residual agents): every layer whose activation tensor maps were
      // supplied.  Its main loop runs on 2 operand stages and drains the hi*hi accumulator every 4 chunks (16 truncating
      // accumulations, ~4e-7 relative).  HD_TEPI_MAXK (env) restricts it to K <= that value (A/B switch; 256 = round-2 first cut).
      static const int tepi_maxk = [] { const char *e = getenv("HD_TEPI_MAXK"); return e ? atoi(e) : (1 << 30); }();
      const bool res_plain = !p.res || (p.res_stride == 1 && p.res_H == p.Ho && p.res_W == p.Wo);
      const bool maps = (!p.res || d->tmap_res) && (!p.out || p.out_sub || d->tmap_out) && (!p.out_hi || (d->tmap_out_hi && d->tmap_out_lo));
      if (maps && res_plain && p.Cout % 32 == 0 && !(d->flags & HD_CONV_NO_TMA_EPILOGUE) && (p.K <= 256 || p.K <= tepi_maxk))
      {
        // few-tile GEMMs (IEF FCs at 640 rows: 40 tiles of 128x128 for 148 SMs): 64-wide tiles double the CTA count
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int tiles128 = ceil_div(p.M, BM) * ceil_div(p.Cout, 128);
        const bool narrow = p.Cout <= 64 || (d->tmap_hi_n64 && d->tmap_lo_n64 && p.Cout % 64 == 0 && 2 * tiles128 <= sms);
        static const bool slim_ok = [] { const char *e = getenv("HD_TEPI_SLIM"); return !e || atoi(e) != 0; }();
        if (slim_ok && !p.res && !p.out && p.out_hi && p.K > 256)     // split-only, long K: 3 operand stages + slim ring
          return narrow ? launch_tc<64, true, 4, true, false, true, false, 2>(p, d, st)
                        : launch_tc<128, true, 4, true, false, true, false, 2>(p, d, st);
        return narrow ? launch_tc<64, true, 4, true, false, true, false, 1>(p, d, st)
                      : launch_tc<128, true, 4, true, false, true, false, 1>(p, d, st);
      }
    }
    if (p.out_sub) {
      set_last_error_text("hd_conv_gemm: out_subsample is only implemented by the TMA epilogue (pre-split input, Cout % 32 == 0, activation maps given)");
      return HD_ERR_UNSUPPORTED;
    }
    if (p.K <= 256)      // (strided-subsample residuals / no tensor maps) two drain/epilogue warp groups with per-thread global accesses
      return p.Cout <= 64 ? launch_tc<64, true, 4, true, false, true, true>(p, d, st) : launch_tc<128, true, 4, true, false, true, true>(p, d, st);
    return p.Cout <= 64 ? launch_tc<64, true, 2, true, false, true>(p, d, st) : launch_tc<128, true, 2, true, false, true>(p, d, st);
  }
  if (half && p.Cin % bke != 0) {      // ragged Cin (resnet conv1: 7x7x3): element-wise gather producer, K zero-padded
    const int segp = (p.KW * p.Cin + 7) & ~7;
    if (p.K_pad % 64 != 0 || p.K_pad < p.KH * segp || p.Cout > 64 || p.pre_scale || p.in_ld != p.Cin) {
      set_last_error_text("hd_conv_gemm(tc gather): needs K_pad % 64 == 0, K_pad >= KH*roundup8(KW*Cin), Cout <= 64, dense pixels, no prologue");
      return HD_ERR_INVALID;
    }
    return launch_tc<64, true, 2, true, true>(p, d, st);
  }
  if (!p.in || p.Cin % bke != 0 || p.in_ld % 4 != 0 || !aligned16(p.in) || p.K % bke != 0 ||
      (p.pre_scale && (!aligned16(p.pre_scale) || !aligned16(p.pre_shift) || p.pre_img_stride % 4 != 0))) {
    set_last_error_text("hd_conv_gemm(tc): needs Cin % 32 (tf32) / % 64 (fp16) == 0 and 16-byte aligned input / prologue vectors");
    return HD_ERR_INVALID;
  }
  const bool split = d->impl != HD_IMPL_TC_1XTF32;
  // split modes: drain every 2 chunks = 8 truncating accumulations (~2e-7 relative, below fp32 SIMT summation noise;
  // the drain warps are idle most of the time, so the extra drains are free);
  // 1xTF32 is ~1e-3 anyway: drain rarely
  if (half) return p.Cout <= 64 ? launch_tc<64, true, 2, true>(p, d, st) : launch_tc<128, true, 2, true>(p, d, st);
  if (p.Cout <= 64) return split ? launch_tc<64, true, 2, false>(p, d, st) : launch_tc<64, false, 8, false>(p, d, st);
  return split ? launch_tc<128, true, 2, false>(p, d, st) : launch_tc<128, false, 8, false>(p, d, st);
}

}  // namespace hd

// K-major weight matrix [rows, k_pad] (fp32 for the tf32 path, fp16 for the fp16 path) -> CUtensorMap with a
// {128 bytes x box_rows} box and 128-byte swizzle.  box_rows must equal the kernel's N tile: 64 when Cout <= 64, else 128.
extern "C" int hd_make_weight_tmap(const void *w_nk, int rows, int k_pad, int box_rows, int elem_bytes, void *tmap_out) {
  HD_REQUIRE(w_nk && tmap_out && rows > 0 && k_pad > 0 && (elem_bytes == 4 || elem_bytes == 2) && k_pad % (128 / elem_bytes) == 0 &&
                 (box_rows == 64 || box_rows == 128) && rows % box_rows == 0 && hd::aligned16(w_nk),
             "hd_make_weight_tmap: bad arguments");
  hd::EncodeTiledFn fn = hd::get_encode_fn();
  if (!fn) { hd::set_last_error_text("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return HD_ERR_UNSUPPORTED; }
  alignas(64) CUtensorMap tm;
  const cuuint64_t gdim[2] = {(cuuint64_t)k_pad, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)k_pad * (cuuint64_t)elem_bytes};
  const cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(&tm, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(w_nk),
                  gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[96];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    hd::set_last_error_text(msg);
    return HD_ERR_CUDA;
  }
  memcpy(tmap_out, &tm, sizeof(tm));
  return HD_OK;
}

// Row-major activation matrix [rows, cols] (leading dimension ld_elems; fp32 or fp16) -> CUtensorMap with a {32 columns x 128 rows}
// box: the slab the TMA epilogue loads (residual) / stores (outputs).  fp32 slabs use the 128-byte swizzle, fp16 slabs the 64-byte one.
extern "C" int hd_make_act_tmap(const void *base, long long rows, int cols, long long ld_elems, int elem_bytes, void *tmap_out) {
  HD_REQUIRE(base && tmap_out && rows > 0 && cols > 0 && ld_elems >= cols && (elem_bytes == 4 || elem_bytes == 2) && hd::aligned16(base) &&
                 (ld_elems * elem_bytes) % 16 == 0,
             "hd_make_act_tmap: bad arguments (16-byte aligned base and row pitch required)");
  hd::EncodeTiledFn fn = hd::get_encode_fn();
  if (!fn) { hd::set_last_error_text("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return HD_ERR_UNSUPPORTED; }
  alignas(64) CUtensorMap tm;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld_elems * (cuuint64_t)elem_bytes};
  const cuuint32_t box[2] = {32u, 128u};
  const cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(&tm, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), gdim,
                  gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, elem_bytes == 4 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[96];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled (activation) failed with CUresult %d", (int)r);
    hd::set_last_error_text(msg);
    return HD_ERR_CUDA;
  }
  memcpy(tmap_out, &tm, sizeof(tm));
  return HD_OK;
}
