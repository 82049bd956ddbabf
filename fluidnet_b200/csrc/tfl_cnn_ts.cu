// 3x3x3 convolution layer with the A operand in TENSOR MEMORY ("TS" tcgen05.mma): the second,
// faster tensor-core formulation of the projection network's conv layers (3xTF32 arithmetic,
// same numerics as the shared-memory-operand kernel in tfl_cnn_tc.cu).
//
// Why: with only 8 output channels the shared-memory-operand kernel is bound by the tensor core's
// operand fetch from shared memory (measured ~64 B/clk/SM: a 128 x 8 tf32 A tile costs ~80 cycles
// per MMA whatever N is).  Here the activations are written ONCE into TMEM and every filter tap
// reads them from there; only the (tiny) weight operand comes from shared memory:
//
//   * an M tile is one x-row of 128 voxels at fixed (y, z): TMEM lane = x, so a filter tap in
//     (dy, dz) is simply ANOTHER ROW = another TMEM column range -- no data movement per tap;
//     the three x-taps are folded into N exactly as in tfl_cnn_tc.cu and resolved in the epilogue
//     (lane shifts: warp shuffles + a 16-float exchange at the three warp boundaries);
//   * a CTA owns TY = 2 output rows and marches along z with a ring of 4 planes x 4 rows in TMEM
//     (16 columns per row: 8 channels "hi" + 8 channels "lo" of the 3xTF32 split) = 256 columns,
//     plus four 64-column accumulator buffers = 512 columns, the whole tensor memory of the SM;
//   * per (dy, dz) two MMAs: [hi x (w_hi | w_lo)] (N = 64: columns 0-23 hi.hi, 32-55 hi.lo) and
//     [lo x w_hi] (N = 32, accumulated onto columns 32-63); 18 MMAs per 128-voxel row;
//   * warp-specialised: 8 loader warps in two sets alternating planes (global -> registers -> hi/lo
//     split -> tcgen05.st, each set prefetching its next plane), 1 MMA-issuer warp, 16 epilogue warps (four warpgroups, one accumulator
//     buffer each -- the epilogue is instruction-latency bound per warp, so tiles in flight are what
//     buys throughput); mbarriers hand planes and accumulators back and forth.
//
// Covers grids with nx <= 128 (one M tile spans the row, so the x padding is the zero border);
// wider grids use tfl_cnn_tc.cu.
#include <cuda_runtime.h>
#include <stdint.h>

#include "tfl_cnn_tc.h"

namespace tfl {

namespace {

constexpr int kTY = 2;                     // output rows per CTA
constexpr int kRows = kTY + 2;             // rows per staged plane
constexpr int kRing = 4;                   // planes resident in TMEM
constexpr int kSlotCols = 16;              // hi (8) + lo (8)
constexpr int kACols = kRing * kRows * kSlotCols;   // 256
constexpr int kNDBuf = 4;                  // accumulator buffers (one per epilogue warpgroup)
constexpr int kDCols = 64;
constexpr int kGroups = 9;
constexpr int kNB = 64;                    // rows of one B block
constexpr int kBGroupBytes = 2 * kNB * 16;
constexpr int kLoadSets = 2;                // loader warp sets, alternating planes
constexpr int kLoadWarps = 4 * kLoadSets;
constexpr int kEpiWarps = 4 * kNDBuf;
constexpr int kIssuerWarp = kLoadWarps + kEpiWarps;
constexpr int kThreadsTS = (kIssuerWarp + 1) * 32;   // 8 loader + 16 epilogue + 1 issuer warps

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done) __nanosleep(40);      // do not starve the other warps of this SM sub-partition while polling
  } while (!done);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One lane of a fully active warp (warp-uniform control flow keeps the operands in uniform registers).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ uint32_t hi_bits(float v) { return __float_as_uint(v) & 0xFFFFE000u; }

// barriers: plane_full[4] (128 loader arrivals), plane_free[4] (1 commit), d_full[2] (1 commit),
// d_empty[2] (128 epilogue arrivals)
enum { kBarPlaneFull = 0, kBarPlaneFree = 4, kBarDFull = 8, kBarDEmpty = 8 + kNDBuf, kNumBars = 8 + 2 * kNDBuf };

__device__ long long* g_ts_dbg = nullptr;   // optional wait-time counters (tests/dbg only)

template <int IN_PLANES, bool FINAL>
__global__ void __launch_bounds__(kThreadsTS, 1)
k_conv3_ts(const float4* __restrict__ in, float4* __restrict__ out, float* __restrict__ p_net,
           const float* __restrict__ wB, const float* __restrict__ bias, const float* __restrict__ tail,
           ConvTcGeo g, int zchunk) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int B_BYTES = kGroups * kBGroupBytes;
  uint8_t* sB = smem;
  uint64_t* bars = (uint64_t*)(smem + B_BYTES);
  uint32_t* tmem_slot = (uint32_t*)(bars + kNumBars);
  float* sTail = (float*)(tmem_slot + 4);                 // bias[8] (+ w4[64] b4[8] w5[8] b5[1])
  float* sEdge = sTail + 96;                              // [2 buffers][4 wg][4 warps][2 sides][8]
  float* sZero = sEdge + 2 * kNDBuf * 64;                 // 16 zeros

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long* dbg = g_ts_dbg ? g_ts_dbg + 8 * (blockIdx.z * gridDim.y + blockIdx.y) : nullptr;
  const long long t_start = clock64();
  const int y0 = blockIdx.y * kTY;
  const int nchunks = (g.nz + zchunk - 1) / zchunk;
  const int b = blockIdx.z / nchunks;
  const int zc0 = (blockIdx.z % nchunks) * zchunk;
  const int nq = min(zchunk, g.nz - zc0);                 // output planes of this CTA
  const int nplanes = nq + 2;

  if (tid == 0) {
    for (int i = 0; i < 4; i++) mbar_init(smem_u32(&bars[kBarPlaneFull + i]), 128);
    for (int i = 0; i < 4; i++) mbar_init(smem_u32(&bars[kBarPlaneFree + i]), 1);
    for (int i = 0; i < kNDBuf; i++) mbar_init(smem_u32(&bars[kBarDFull + i]), 1);
    for (int i = 0; i < kNDBuf; i++) mbar_init(smem_u32(&bars[kBarDEmpty + i]), 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kIssuerWarp) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  for (int i = tid; i < B_BYTES / 16; i += kThreadsTS) ((float4*)sB)[i] = __ldg((const float4*)wB + i);
  const int n_tail = FINAL ? (8 + 64 + 8 + 8 + 1) : 8;
  for (int i = tid; i < n_tail; i += kThreadsTS) sTail[i] = (i < 8) ? bias[i] : tail[i - 8];
  if (tid < 16) sZero[tid] = 0.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const long long plane_g = (long long)(g.nz + 2) * g.py * g.px;     // float4 per global channel plane
  const long long batch_g = plane_g * 2;

  if (warp < kLoadWarps) {
    // ===== loaders: thread <-> x position.  Two warp sets alternate planes; a row's registers are
    // refilled with the set's NEXT plane right after the row has been written to TMEM, so the global
    // loads have a whole plane period to land and only one register set is live. =====
    const int lset = warp >> 2, qw = warp & 3;
    const int l = qw * 32 + lane;                         // 0..127 = unpadded x
    const bool xin = l < g.nx;
    const float4* inb = in + b * batch_g + (l + 1);       // padded x = l + 1
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 cur[kRows][IN_PLANES];
    auto fetch_row = [&](int p, int r) {
      const int zp = zc0 + p;                             // padded z of plane p (unpadded zc0 - 1 + p)
      const int yp = y0 + r;                              // padded y (unpadded y0 - 1 + r)
      const bool ok = xin && p < nplanes && yp < g.py && zp < g.nz + 2;
      const long long o = ((long long)(ok ? zp : 0) * g.py + (ok ? yp : 0)) * g.px;
#pragma unroll
      for (int h = 0; h < IN_PLANES; h++) cur[r][h] = ok ? __ldg(inb + h * plane_g + o) : z4;
    };
#pragma unroll
    for (int r = 0; r < kRows; r++) fetch_row(lset, r);
    for (int p = lset; p < nplanes; p += kLoadSets) {
      const int ring = p & 3;
      if (p >= kRing) {
        mbar_wait(smem_u32(&bars[kBarPlaneFree + ring]), (uint32_t)(((p >> 2) - 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
#pragma unroll
      for (int r = 0; r < kRows; r++) {
        uint32_t v[16];
        const float4 a0 = cur[r][0];
        const float4 a1 = IN_PLANES == 2 ? cur[r][IN_PLANES - 1] : z4;
        const float f[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int c = 0; c < 8; c++) {
          const uint32_t hb = hi_bits(f[c]);
          v[c] = hb;
          v[8 + c] = __float_as_uint(f[c] - __uint_as_float(hb));
        }
        const uint32_t taddr = tmem_base + (uint32_t)((ring * kRows + r) * kSlotCols) + ((uint32_t)(qw * 32) << 16);
        tmem_st16(taddr, v);
        fetch_row(p + kLoadSets, r);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(smem_u32(&bars[kBarPlaneFull + ring]));
    }
  } else if (warp == kIssuerWarp) {
    // ===== MMA issuer: the whole warp runs the (uniform) control flow, one elected lane issues =====
    {
      const uint32_t sB_u = smem_u32(sB);
      constexpr uint32_t IDESC_MAIN = make_idesc(128, 64);
      constexpr uint32_t IDESC_LO = make_idesc(128, 32);
      uint64_t db[kGroups];
#pragma unroll
      for (int gi = 0; gi < kGroups; gi++) db[gi] = make_desc(sB_u + gi * kBGroupBytes, kNB * 16, 128);
      int tile = 0;
      for (int q = 0; q < nq; q++) {
        // planes q, q+1, q+2 must be resident (q and q+1 were awaited for earlier output planes)
        const long long w0 = clock64();
        for (int p = (q == 0 ? 0 : q + 2); p <= q + 2; p++)
          mbar_wait(smem_u32(&bars[kBarPlaneFull + (p & 3)]), (uint32_t)((p >> 2) & 1));
        if (dbg && lane == 0) dbg[2] += clock64() - w0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int t = 0; t < kTY; t++, tile++) {
          const int buf = tile & (kNDBuf - 1);
          if (tile >= kNDBuf) {
            const long long w1 = clock64();
            mbar_wait(smem_u32(&bars[kBarDEmpty + buf]), (uint32_t)(((tile / kNDBuf) - 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (dbg && lane == 0) dbg[3] += clock64() - w1;
          }
          const uint32_t d_addr = tmem_base + (uint32_t)(kACols + buf * kDCols);
#pragma unroll
          for (int gi = 0; gi < kGroups; gi++) {
            const int dz = gi / 3, dy = gi % 3;
            const uint32_t a_hi = tmem_base + (uint32_t)((((q + dz) & 3) * kRows + (t + dy)) * kSlotCols);
            if (elect_one()) {
              umma_tf32_ts(d_addr, a_hi, db[gi], IDESC_MAIN, gi > 0 ? 1u : 0u);
              umma_tf32_ts(d_addr + 32, a_hi + 8, db[gi], IDESC_LO, 1u);
            }
          }
          if (elect_one()) umma_commit(smem_u32(&bars[kBarDFull + buf]));
          __syncwarp();
        }
        if (elect_one()) umma_commit(smem_u32(&bars[kBarPlaneFree + (q & 3)]));   // plane q is dead after output plane q
        __syncwarp();
      }
      if (dbg && lane == 0) dbg[4] = clock64() - t_start;
    }
  } else {
    // ===== epilogue: warps 4..19, warpgroup wg owns accumulator buffer wg =====
    const int ew = warp - kLoadWarps, wg = ew >> 2, qtr = ew & 3;
    const float* sBias = sTail;
    const int ntiles = nq * kTY;
    int it = 0;
    for (int tile = wg; tile < ntiles; tile += kNDBuf, it++) {
      const long long w2 = clock64();
      mbar_wait(smem_u32(&bars[kBarDFull + wg]), (uint32_t)((tile / kNDBuf) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (dbg && ew == 0 && lane == 0) dbg[5] += clock64() - w2;
      const uint32_t taddr = tmem_base + (uint32_t)(kACols + wg * kDCols) + ((uint32_t)(qtr * 32) << 16);
      uint32_t r0[24], r1[24];
      tmem_ld8(taddr + 0, r0);
      tmem_ld8(taddr + 8, r0 + 8);
      tmem_ld8(taddr + 16, r0 + 16);
      tmem_ld8(taddr + 32, r1);
      tmem_ld8(taddr + 40, r1 + 8);
      tmem_ld8(taddr + 48, r1 + 16);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(smem_u32(&bars[kBarDEmpty + wg]));
      const long long e1 = clock64();
      if (dbg && ew == 0 && lane == 0) dbg[7] += e1 - w2;          // (wait d_full +) TMEM loads
      float dm[8], d0[8], dp[8];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        dm[i] = __uint_as_float(r0[i]) + __uint_as_float(r1[i]);
        d0[i] = __uint_as_float(r0[8 + i]) + __uint_as_float(r1[8 + i]);
        dp[i] = __uint_as_float(r0[16 + i]) + __uint_as_float(r1[16 + i]);
      }
      // x shift across the three warp boundaries of the 128-lane tile
      float* edge = sEdge + (((it & 1) * kNDBuf + wg) * 4) * 16;
      if (lane == 31) {
#pragma unroll
        for (int o = 0; o < 8; o++) edge[qtr * 16 + o] = dm[o];
      }
      if (lane == 0) {
#pragma unroll
        for (int o = 0; o < 8; o++) edge[qtr * 16 + 8 + o] = dp[o];
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
      if (dbg && ew == 0 && lane == 0) dbg[0] += clock64() - e1;   // reuse slot 0: edge exchange + barrier
      // neighbours' edge values: broadcast shared-memory reads + selects (no divergent branches);
      // the outermost quarters read the zero block (x padding of the row)
      const float* eL = qtr > 0 ? edge + (qtr - 1) * 16 : sZero;
      const float* eR = qtr < 3 ? edge + (qtr + 1) * 16 + 8 : sZero;
      float h[8];
#pragma unroll
      for (int o = 0; o < 8; o++) {
        const float sa = __shfl_up_sync(0xffffffffu, dm[o], 1);
        const float sc = __shfl_down_sync(0xffffffffu, dp[o], 1);
        const float a = lane == 0 ? eL[o] : sa;
        const float c = lane == 31 ? eR[o] : sc;
        const float v = (a + d0[o]) + c + sBias[o];
        h[o] = v > 0.0f ? v : 0.0f;
      }
      const int xg = qtr * 32 + lane;
      const int yg = y0 + (tile % kTY);
      const int zg = zc0 + tile / kTY;
      const bool valid = xg < g.nx && yg < g.ny && zg < g.nz;
      if (!FINAL) {
        if (valid) {
          const long long o = b * batch_g + ((long long)(zg + 1) * g.py + (yg + 1)) * g.px + (xg + 1);
          out[o] = make_float4(h[0], h[1], h[2], h[3]);
          out[o + plane_g] = make_float4(h[4], h[5], h[6], h[7]);
        }
      } else {
        const float* w4 = sTail + 8;
        const float* b4 = w4 + 64;
        const float* w5 = b4 + 8;
        const float b5 = w5[8];
        float pacc = b5;
#pragma unroll
        for (int o = 0; o < 8; o++) {
          float a = b4[o];
#pragma unroll
          for (int c = 0; c < 8; c++) a = fmaf(h[c], w4[o * 8 + c], a);
          a = a > 0.0f ? a : 0.0f;
          pacc = fmaf(a, w5[o], pacc);
        }
        if (valid) p_net[(long long)b * g.nz * g.ny * g.nx + ((long long)zg * g.ny + yg) * g.nx + xg] = pacc;
      }
      if (dbg && ew == 0 && lane == 0) dbg[1] += clock64() - e1;   // reuse slot 1: everything after the loads
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (dbg && tid == 0) dbg[6] = clock64() - t_start;
  if (warp == kIssuerWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

template <int IN_PLANES, bool FINAL>
void launch_ts(const float4* in, float4* out, float* p_net, const float* wB, const float* bias,
               const float* tail, const ConvTcGeo& g, cudaStream_t st) {
  const size_t smem = (size_t)kGroups * kBGroupBytes + kNumBars * 8 + 16 + 96 * 4 + 2 * kNDBuf * 64 * 4 + 64 + 64;
  auto kern = k_conv3_ts<IN_PLANES, FINAL>;
  static unsigned long long configured = 0;       // per device (function attributes are)
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((configured >> (dev & 63)) & 1ULL)) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured |= 1ULL << (dev & 63);
  }
  const int nty = (g.ny + kTY - 1) / kTY;
  // one wave of persistent-style CTAs: as many z chunks as fit 148 SMs
  int chunks = 148 / (nty * g.nb);
  if (chunks < 1) chunks = 1;
  if (chunks > g.nz) chunks = g.nz;
  int zchunk = (g.nz + chunks - 1) / chunks;
  if (zchunk < 4 && g.nz >= 4) zchunk = 4;
  const int nchunks = (g.nz + zchunk - 1) / zchunk;
  dim3 grid(1, nty, nchunks * g.nb);
  kern<<<grid, kThreadsTS, smem, st>>>(in, out, p_net, wB, bias, tail, g, zchunk);
}

}  // namespace

void conv_ts_set_debug(long long* dev_buf) { cudaMemcpyToSymbol(g_ts_dbg, &dev_buf, sizeof(dev_buf)); }

int conv_ts_b_floats() { return kGroups * 2 * kNB * 4; }

// [cout=8][cin][3][3][3] -> per (dz, dy) block [kchunk 2][n 64][4]: rows 0-23 tf32 "hi" weights
// (n = kx*8 + o), rows 32-55 the fp32 residual "lo", other rows zero.
void conv_ts_pack_weights(const float* w, int cin, float* out) {
  for (int i = 0; i < conv_ts_b_floats(); i++) out[i] = 0.0f;
  for (int dz = 0; dz < 3; dz++)
    for (int dy = 0; dy < 3; dy++) {
      float* blk = out + (size_t)(dz * 3 + dy) * 2 * kNB * 4;
      for (int kx = 0; kx < 3; kx++)
        for (int o = 0; o < 8; o++)
          for (int c = 0; c < cin; c++) {
            const float v = w[((((size_t)o * cin + c) * 3 + dz) * 3 + dy) * 3 + kx];
            union { float f; uint32_t u; } hi;
            hi.f = v;
            hi.u &= 0xFFFFE000u;
            const int n = kx * 8 + o;
            blk[((c >> 2) * kNB + n) * 4 + (c & 3)] = hi.f;
            blk[((c >> 2) * kNB + 32 + n) * 4 + (c & 3)] = v - hi.f;
          }
    }
}

bool conv_ts_supported(const ConvTcGeo& g) { return g.nx <= 128 && g.nz >= 1; }

int launch_conv3_ts(const float* in, float* out, float* p_net, const float* wB, const float* bias,
                    const float* tail, int in_planes, int final_layer, const ConvTcGeo& g, cudaStream_t st) {
  const float4* i4 = (const float4*)in;
  float4* o4 = (float4*)out;
  if (in_planes == 1 && !final_layer) { launch_ts<1, false>(i4, o4, p_net, wB, bias, tail, g, st); return 1; }
  if (in_planes == 2 && !final_layer) { launch_ts<2, false>(i4, o4, p_net, wB, bias, tail, g, st); return 1; }
  if (in_planes == 2 && final_layer) { launch_ts<2, true>(i4, o4, p_net, wB, bias, tail, g, st); return 1; }
  return -1;
}

}  // namespace tfl
