// b2t_pcg_tm.cuh -- k_pcg_tm: the matrix-free PCG of k_pcg3 with the per-lane matrix data resident in TENSOR MEMORY (TMEM) instead of
// registers, so that TWO instances are resident per SM (16 warps instead of 8) and the FP64 pipe, the shared-memory pipe and the
// barrier chain of one instance overlap with those of the other.
//
// Why: k_pcg3 keeps Ab_k (6 x 5 per lane), D^-1 rows (3 x 12 per lane) and the Ghat factors of one instance in registers for all
// iterations: 254 registers x 256 threads = the whole register file for ONE instance, 2 warps per scheduler, and the kernel sits at
// 22 % of the FP64 peak because nothing hides its shared-memory / shuffle latencies (6 lanes per knot, D^-1 in shared
// memory and the column form were all measured slower).  Blackwell's tensor memory is a second 256 KB on-chip store per SM with its own
// load path (tcgen05.ld: measured here 350-430 B/clk/SM with 8 warps against 128 B/clk of shared memory, 12 cycles latency, and it
// overlaps with DFMA issue: scripts/ubench/tmem_bench.cu).  No tensor-core instruction is involved: TMEM is used as a per-thread
// matrix store.  Each thread owns a 32-bit x 128-column window (lane = its lane of the warp's TMEM sub-partition (warp % 4), columns
// (warp / 4) * 128 ...): 64 doubles = the 36 doubles of its D^-1 rows + 28 of its 30 Ab entries; the other 2 and the Ghat factors stay
// in registers.  The products stream the window back in chunks of <= 24 registers right before they are used, in k_pcg3's order of
// operations => bit-identical iterates and iteration counts (tests/test_gpu_variants.py).
//
// One persistent CTA of 512 threads per SM = two independent halves of 256 threads (named barriers 1 and 2); a half takes the next
// instance of the work list from a global ticket counter when its PCG ends (iteration counts differ per instance).
// Reference: GBD-PCG-Python/PCG.py:66-111 (same recurrence and exit test as k_pcg3).
#pragma once
#include <cstdint>

namespace b2t {

__device__ __forceinline__ void tm_ld_x2(uint32_t a, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];"
               : "=r"(v[0]), "=r"(v[1])
               : "r"(a));
}
__device__ __forceinline__ void tm_st_x2(uint32_t a, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};"
               :: "r"(a), "r"(v[0]), "r"(v[1]));
}
__device__ __forceinline__ void tm_ld_x4(uint32_t a, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(a));
}
__device__ __forceinline__ void tm_st_x4(uint32_t a, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               :: "r"(a), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]));
}
__device__ __forceinline__ void tm_ld_x8(uint32_t a, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(a));
}
__device__ __forceinline__ void tm_st_x8(uint32_t a, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "r"(a), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
}
__device__ __forceinline__ void tm_ld_x16(uint32_t a, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(a));
}
__device__ __forceinline__ void tm_st_x16(uint32_t a, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               :: "r"(a), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]));
}
__device__ __forceinline__ void tm_ld_x32(uint32_t a, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(a));
}
__device__ __forceinline__ void tm_st_x32(uint32_t a, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
               :: "r"(a), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]));
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// NW 32-bit columns (any even count) starting at `a` -> v / v -> TMEM, as the fewest power-of-two transfers
template <int NW>
__device__ __forceinline__ void tm_load(uint32_t a, uint32_t* v) {
  if constexpr (NW >= 32) { tm_ld_x32(a, v); tm_load<NW - 32>(a + 32, v + 32); }
  else if constexpr (NW >= 16) { tm_ld_x16(a, v); tm_load<NW - 16>(a + 16, v + 16); }
  else if constexpr (NW >= 8) { tm_ld_x8(a, v); tm_load<NW - 8>(a + 8, v + 8); }
  else if constexpr (NW >= 4) { tm_ld_x4(a, v); tm_load<NW - 4>(a + 4, v + 4); }
  else if constexpr (NW >= 2) { tm_ld_x2(a, v); tm_load<NW - 2>(a + 2, v + 2); }
}
template <int NW>
__device__ __forceinline__ void tm_store(uint32_t a, const uint32_t* v) {
  if constexpr (NW >= 32) { tm_st_x32(a, v); tm_store<NW - 32>(a + 32, v + 32); }
  else if constexpr (NW >= 16) { tm_st_x16(a, v); tm_store<NW - 16>(a + 16, v + 16); }
  else if constexpr (NW >= 8) { tm_st_x8(a, v); tm_store<NW - 8>(a + 8, v + 8); }
  else if constexpr (NW >= 4) { tm_st_x4(a, v); tm_store<NW - 4>(a + 4, v + 4); }
  else if constexpr (NW >= 2) { tm_st_x2(a, v); tm_store<NW - 2>(a + 2, v + 2); }
}
// ND doubles from TMEM columns [a, a + 2 ND): load, wait, and pin the registers behind the wait for the compiler
template <int ND>
__device__ __forceinline__ void tm_load_f64(uint32_t a, double* out) {
  uint32_t w[2 * ND];
  tm_load<2 * ND>(a, w);
  tm_wait_ld();
#pragma unroll
  for (int i = 0; i < 2 * ND; ++i) asm volatile("" : "+r"(w[i]));
#pragma unroll
  for (int i = 0; i < ND; ++i) out[i] = __hiloint2double((int)w[2 * i + 1], (int)w[2 * i]);
}
// the same in two halves: issue the loads (no wait), and later wait + pin + convert -- the first chunk of a product is fetched before the
// barrier the product waits on anyway, so its TMEM latency is hidden behind the barrier
template <int ND>
__device__ __forceinline__ void tm_issue_f64(uint32_t a, uint32_t* w) { tm_load<2 * ND>(a, w); }
template <int ND>
__device__ __forceinline__ void tm_complete_f64(uint32_t* w, double* out) {
  tm_wait_ld();
#pragma unroll
  for (int i = 0; i < 2 * ND; ++i) asm volatile("" : "+r"(w[i]));
#pragma unroll
  for (int i = 0; i < ND; ++i) out[i] = __hiloint2double((int)w[2 * i + 1], (int)w[2 * i]);
}
template <int ND>
__device__ __forceinline__ void tm_store_f64(uint32_t a, const double* in) {
  uint32_t w[2 * ND];
#pragma unroll
  for (int i = 0; i < ND; ++i) { w[2 * i] = (uint32_t)__double2loint(in[i]); w[2 * i + 1] = (uint32_t)__double2hiint(in[i]); }
  tm_store<2 * ND>(a, w);
}

// dinv u - hh h has two products and one subtraction: which product is fused with the subtraction decides the rounding.  Left to the
// compiler, the choice depends on predication and scheduling (it differs between instantiations of the same source), so the two forms
// that k_pcg3's binary uses are spelled out: form A everywhere except the W-only copy (u2) of all but the lane's last column.
__device__ __forceinline__ double ghat_a(double dinv, double u, double hh, double h) { return __fma_rn(u, dinv, -__dmul_rn(hh, h)); }
__device__ __forceinline__ double ghat_b(double dinv, double u, double hh, double h) { return __fma_rn(-h, hh, __dmul_rn(dinv, u)); }

constexpr int PCGTM_THREADS = 512;   // threads per CTA: HT = 256 -> two instances (N <= 64) side by side, HT = 512 -> one instance (N <= 128)
constexpr int PCGTM_CAP = 64;        // doubles per thread in TMEM: 512 columns / (4 warps per sub-partition) / 2
constexpr bool pcg_tm_eligible() { return NX % 4 == 0 && (NX / 4) * NX <= PCGTM_CAP; }

// NK: compile-time horizon (0: d.N at run time), INTEG: compile-time integrator type (-1: d.integrator at run time).  The specialised
// instantiation (N = 64 of BASELINE.json, Euler) turns the liveness tests, strides and integrator selects that the compiler otherwise
// re-derives every iteration (it has no registers to keep them) into immediates.
// ABC: Ab columns fetched per TMEM load (register pressure against the number of tcgen05.ld; 1 / 2 / 3 measured 688.8 / 683.1 / 681.6 ms
// of PCG per default step: within noise of each other, 2 kept).
// PRE: fetch the first TMEM chunk of every product ahead of the barrier in front of it (PCG 666 -> 650 ms per default step).
// Measured on top of this and NOT kept (all bit-identical): requesting chunk q + 1 right after chunk q has arrived, so that it
// loads behind the FMAs of chunk q (654 ms); issuing the shared-memory loads of the group's own block ahead of the barrier as well (653 ms,
// more spills); neighbour-warp mbarriers instead of the 4 exchange barriers per iteration (734 ms: an arrive + try_wait round trip and ~40
// instructions per exchange cost more than the skew of 8 warps at a block barrier); the step recovery dz = Ghat (g - C^T l) as an epilogue
// of this kernel instead of the prologue of k_linesearch (line search -6.5 ms, this kernel +6.5 ms and a worse register allocation of its loop).
template <typename T, int HT, int NK = 0, int INTEG = -1, int ABC = 2, bool PRE = true>
__global__ void __launch_bounds__(PCGTM_THREADS, 1) k_pcg_tm(Dev<T> d, const int* list, const int* count, int* ticket, int stair, T tol, int max_iter) {
  static_assert(sizeof(T) == 8, "k_pcg_tm: fp64 only (the fp32 solver keeps k_pcg3)");
  constexpr int LPK = 4;
  static_assert(HT == 256 || HT == 512, "one or two instances per CTA");
  constexpr int RPT = (NX % LPK == 0) ? NX / LPK : 1;   // owned rows per lane
  constexpr int MC = (NM + LPK - 1) / LPK;               // Ab columns per lane
  constexpr int NMS = PCG3_NMS;
  constexpr int PD_D = RPT * NX;                         // doubles of D^-1 rows per lane (all in TMEM)
  constexpr int AB_D = NJ * MC;                          // doubles of Ab per lane, column-major f = i * NJ + a
  constexpr int AB_T = AB_D < PCGTM_CAP - PD_D ? AB_D : PCGTM_CAP - PD_D;   // ... of which in TMEM
  constexpr int AB_R = AB_D - AB_T;                      // ... and in registers
  constexpr int PD_CH = 4 * RPT;                         // doubles per D^-1 chunk: 4 columns x RPT rows
  using T2 = double2;
  __shared__ uint32_t tm_base_s;
  __shared__ int slot_s[2];
  extern __shared__ unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "l"((uint64_t)__cvta_generic_to_shared(&tm_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm_base = tm_base_s;
  // this thread's window: TMEM lane = 32 (warp % 4) + lane id (implied by the .32x32b shape), columns (warp / 4) * 128 ...
  const uint32_t tmw = tm_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * (2 * PCGTM_CAP));
  const uint32_t tm_pd = tmw, tm_ab = tmw + 2 * PD_D;
  const int h = threadIdx.x / HT, tid = threadIdx.x % HT;
  const int N = NK > 0 ? NK : d.N;
  const size_t K = d.K;
  const size_t half_T = (size_t)2 * (N + 1) * NX + (size_t)2 * N * NMS + 64;
  T* V = reinterpret_cast<T*>(smem_raw) + (size_t)h * half_T;   // [(N+1)][NX]  p / r / O y; block N stays zero
  T* V2 = V + (N + 1) * NX;                   // [(N+1)][NX]  y; block N stays zero
  T* W = V2 + (N + 1) * NX;                   // [N][NMS]     w or q'
  T* Wq = W + N * NMS;                        // [N][NMS]     q
  T* red = Wq + N * NMS;                      // 2 x 32: double-buffered warp partial sums
  auto hbar = [&]() { asm volatile("bar.sync %0, %1;" :: "r"(h + 1), "n"(HT) : "memory"); };
  const bool live = (NK > 0 && LPK * NK >= HT) ? true : tid < LPK * N;
  const int k = live ? tid / LPK : 0;         // knot
  const int g = tid % LPK;                    // lane within the knot group
  const bool has_next = live && (k < N - 1);
  const int jo = (k + 1 == N) ? 0 : k + 1;    // owned block row
  const int c0 = g * MC, i0 = g * RPT;
  const int integ = INTEG >= 0 ? INTEG : d.integrator;
  const T dte = integ == 0 ? d.dt : T(0);
  const T tau = integ == 0 ? T(0) : d.dt;
  const bool top = i0 < NJ;
  const bool odd = (i0 % NJ) != 0;
  const int ownV = jo * NX, ownW = jo * NMS + i0, kV = k * NX, nV = (k + 1) * NX, kW = k * NMS;
  // per-column constants (instance-independent)
  bool cvalid[MC], sval[MC];
  int eidx[MC], sidx[MC];
#pragma unroll
  for (int i = 0; i < MC; ++i) {
    const int c = c0 + i;
    cvalid[i] = live && c < NM;
    eidx[i] = (c < NJ) ? c : ((c < NX) ? c - NJ : 0);
    sval[i] = cvalid[i] && c < NX;
    sidx[i] = (c < NX) ? c : 0;
  }

  for (;;) {
    if (tid == 0) slot_s[h] = atomicAdd(ticket, 1);
    hbar();
    const int slot = slot_s[h];
    if (slot >= *count) break;
    const int b = list[slot];
    const size_t tk = (size_t)b * N + k, tj = (size_t)b * N + jo;
    // ---- resident data: Ghat factors and the Ab tail in registers, D^-1 rows and Ab in this thread's TMEM window
    T dinv[MC], hh[MC], emul[MC], abr[AB_R > 0 ? AB_R : 1];
#pragma unroll
    for (int i = 0; i < MC; ++i) {
      const int c = c0 + i;
      dinv[i] = cvalid[i] ? d.Gh[(size_t)c * K + tk] : T(0);
      hh[i] = cvalid[i] ? d.Gh[(size_t)(NM + c) * K + tk] : T(0);
      emul[i] = (!has_next || !cvalid[i] || c >= NX) ? T(0) : (c < NJ ? T(1) : dte);
    }
    static_for<0, (MC + ABC - 1) / ABC>([&](auto qc) {
      constexpr int I0 = ABC * decltype(qc)::value;
      constexpr int CNT = (I0 + ABC <= MC) ? ABC : MC - I0;
      T m[CNT * NJ];
#pragma unroll
      for (int ii = 0; ii < CNT; ++ii) {
        const int c = c0 + I0 + ii;
#pragma unroll
        for (int a = 0; a < NJ; ++a)
          m[ii * NJ + a] = (cvalid[I0 + ii] && has_next) ? d.dt * d.dyn[(size_t)(a * 3 * NJ + c) * K + tk] + ((c == NJ + a) ? T(1) : T(0)) : T(0);
      }
      constexpr int f0 = I0 * NJ, f1 = (I0 + CNT) * NJ;
      constexpr int t1 = f1 < AB_T ? f1 : AB_T;
      if constexpr (t1 > f0) tm_store_f64<t1 - f0>(tm_ab + 2 * f0, m);
#pragma unroll
      for (int f = (f0 > AB_T ? f0 : AB_T); f < f1; ++f) abr[f - AB_T] = m[f - f0];
    });
    static_for<0, NX / 4>([&](auto qc) {
      constexpr int q = decltype(qc)::value;
      T m[PD_CH];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc)
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          const int c = 4 * q + 2 * cc;
          m[(cc * RPT + r) * 2] = live ? d.Pd[(size_t)((i0 + r) * NX + c) * K + tj] : T(0);
          m[(cc * RPT + r) * 2 + 1] = live ? d.Pd[(size_t)((i0 + r) * NX + c + 1) * K + tj] : T(0);
        }
      tm_store_f64<PD_CH>(tm_pd + 2 * q * PD_CH, m);
    });
    tm_wait_st();
    const T sS = live ? d.Gh[(size_t)(2 * NM) * K + tk] : T(0);
    bool hnz = false;
#pragma unroll
    for (int i = 0; i < MC; ++i) hnz = hnz || (hh[i] != T(0));
    const bool rank1 = d.lim.any != 0 && __any_sync(0xffffffffu, hnz);
    for (int idx = tid; idx < (int)half_T; idx += HT) V[idx] = T(0);
    hbar();

    // Ab columns [I0, I0 + CNT) of this lane -> m (column-major), from TMEM and the register tail
    constexpr int AB0 = ((ABC < MC ? ABC : MC) * NJ) < AB_T ? ((ABC < MC ? ABC : MC) * NJ) : AB_T;      // doubles of the first Ab chunk that live in TMEM
    uint32_t wpre[2 * (AB0 > PD_CH ? AB0 : PD_CH)];          // words of a prefetched first chunk (Ab or D^-1)
    auto ab_prefetch = [&]() { if constexpr (PRE && AB0 > 0) tm_issue_f64<AB0>(tm_ab, wpre); };
    auto pd_prefetch = [&]() { if constexpr (PRE) tm_issue_f64<PD_CH>(tm_pd, wpre); };
    auto ab_get = [&](auto i0c, auto cntc, T* m) {
      constexpr int I0 = decltype(i0c)::value, CNT = decltype(cntc)::value;
      constexpr int f0 = I0 * NJ, f1 = (I0 + CNT) * NJ;
      constexpr int t1 = f1 < AB_T ? f1 : AB_T;
      if constexpr (PRE && I0 == 0 && AB0 > 0) tm_complete_f64<AB0>(wpre, m);
      else if constexpr (t1 > f0) tm_load_f64<t1 - f0>(tm_ab + 2 * f0, m);
#pragma unroll
      for (int f = (f0 > AB_T ? f0 : AB_T); f < f1; ++f) m[f - f0] = abr[f - AB_T];
    };
    auto quad = [&](T v) -> T {
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      return v;
    };
    int red_sel = 0;
    auto bsum = [&](T v) -> T {               // same tree as k_pcg3<T, HT, ...>::bsum
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      T* rb = red + 32 * red_sel;
      red_sel ^= 1;
      if ((tid & 31) == 0) rb[tid >> 5] = v;
      hbar();
      constexpr int NW = HT / 32;
      T t[NW];
#pragma unroll
      for (int i = 0; i < NW; i += 2) { const T2 p2 = *reinterpret_cast<const T2*>(rb + i); t[i] = p2.x; t[i + 1] = p2.y; }
#pragma unroll
      for (int st = 1; st < NW; st *= 2)
#pragma unroll
        for (int i = 0; i + st < NW; i += 2 * st) t[i] += t[i + st];
      return t[0];
    };
    auto publish = [&](T* buf, const T* val) {
      if (live) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) buf[ownV + i0 + r] = val[r];
      }
    };
    // tc_i = (AB_k^T z)_c for my columns, z = block k+1 of `buf`
    auto abt = [&](const T* buf, T* tc) {
      T pi[NJ];
#pragma unroll
      for (int a = 0; a < NJ; a += 2) {
        const T2 zv = *reinterpret_cast<const T2*>(buf + nV + NJ + a);
        if constexpr (INTEG == 0) {           // tau = 0: fma(0, z_q, z_v) = z_v, and the z_q half of the block is not loaded at all
          pi[a] = zv.x;
          if (a + 1 < NJ) pi[a + 1] = zv.y;
        } else {
          const T2 zq = *reinterpret_cast<const T2*>(buf + nV + a);
          pi[a] = tau * zq.x + zv.x;
          if (a + 1 < NJ) pi[a + 1] = tau * zq.y + zv.y;
        }
      }
      static_for<0, (MC + ABC - 1) / ABC>([&](auto qc) {
        constexpr int I0 = ABC * decltype(qc)::value;
        constexpr int CNT = (I0 + ABC <= MC) ? ABC : MC - I0;
        T m[CNT * NJ];
        ab_get(std::integral_constant<int, I0>{}, std::integral_constant<int, CNT>{}, m);
#pragma unroll
        for (int ii = 0; ii < CNT; ++ii) {
          const int i = I0 + ii;
          T acc0 = emul[i] * buf[nV + eidx[i]], acc1 = T(0);
#pragma unroll
          for (int a = 0; a < NJ; a += 2) {
            acc0 += m[ii * NJ + a] * pi[a];
            if (a + 1 < NJ) acc1 += m[ii * NJ + a + 1] * pi[a + 1];
          }
          tc[i] = acc0 + acc1;
        }
      });
    };
    // out_r (owned rows of block jo) = (AB_k zz)_row + sign * Wn[jo][row]
    auto abmul = [&](const T* zc, const T* full, const T* Wn, T sign, T* out) {
      T pb[NJ];
#pragma unroll
      for (int a = 0; a < NJ; ++a) pb[a] = T(0);
      static_for<0, (MC + ABC - 1) / ABC>([&](auto qc) {
        constexpr int I0 = ABC * decltype(qc)::value;
        constexpr int CNT = (I0 + ABC <= MC) ? ABC : MC - I0;
        T m[CNT * NJ];
        ab_get(std::integral_constant<int, I0>{}, std::integral_constant<int, CNT>{}, m);
#pragma unroll
        for (int ii = 0; ii < CNT; ++ii)
#pragma unroll
          for (int a = 0; a < NJ; ++a) pb[a] += m[ii * NJ + a] * zc[I0 + ii];
      });
      T bot[RPT];
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const T send = odd ? pb[r] : pb[RPT + r];
        T keep = odd ? pb[RPT + r] : pb[r];
        keep += __shfl_xor_sync(0xffffffffu, send, 1);
        keep += __shfl_xor_sync(0xffffffffu, keep, 2);
        bot[r] = keep;
      }
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const T zq = full[kW + (top ? i0 + r : 0)], zd = full[kW + NJ + (top ? i0 + r : 0)];
        const T val = top ? (INTEG == 0 ? (zq + dte * zd) : (zq + dte * zd + tau * bot[r])) : bot[r];      // INTEG == 0: tau = 0
        // branch-free form of  has_next ? val + sign Wn : (live ? sign Wn : 0)  (sign = +-1: the product is exact, adding +0 changes nothing)
        const T wn = live ? Wn[ownW + r] : T(0);
        out[r] = (has_next ? val : T(0)) + sign * wn;
      }
    };
    // out = Pd_jo * buf[jo]
    auto pd_mul = [&](const T* buf, T* out) {
      T o0[RPT], o1[RPT];
#pragma unroll
      for (int r = 0; r < RPT; ++r) { o0[r] = T(0); o1[r] = T(0); }
      static_for<0, NX / 4>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        T m[PD_CH];
        if constexpr (PRE && q == 0) tm_complete_f64<PD_CH>(wpre, m);
        else tm_load_f64<PD_CH>(tm_pd + 2 * q * PD_CH, m);
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const T2 v = *reinterpret_cast<const T2*>(buf + ownV + 4 * q + 2 * cc);
#pragma unroll
          for (int r = 0; r < RPT; ++r) { o0[r] += m[(cc * RPT + r) * 2] * v.x; o1[r] += m[(cc * RPT + r) * 2 + 1] * v.y; }
        }
      });
#pragma unroll
      for (int r = 0; r < RPT; ++r) out[r] = o0[r] + o1[r];
    };
    T rr[RPT], xx[RPT], pp[RPT], rt[RPT], yv[RPT], tmp[RPT];
    auto precond = [&]() {      // the caller has issued pd_prefetch()
      publish(V, rr);
      __syncwarp();
      pd_mul(V, yv);
      if (!stair) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) rt[r] = yv[r];
        return;
      }
      publish(V2, yv);
      ab_prefetch();
      hbar();
      T u1[MC], u2[MC], h1 = T(0), h2 = T(0);
      abt(V2, u2);
#pragma unroll
      for (int i = 0; i < MC; ++i) {
        u1[i] = sval[i] ? V2[kV + sidx[i]] : T(0);
        h1 += hh[i] * u1[i];
        h2 += hh[i] * u2[i];
      }
      if (rank1) { h1 = sS * quad(h1); h2 = sS * quad(h2); } else { h1 = T(0); h2 = T(0); }
#pragma unroll
      for (int i = 0; i < MC; ++i) {
        u1[i] = ghat_a(dinv[i], u1[i], hh[i], h1);
        u2[i] = (i < MC - 1) ? ghat_b(dinv[i], u2[i], hh[i], h2) : ghat_a(dinv[i], u2[i], hh[i], h2);
        if (cvalid[i]) { Wq[kW + c0 + i] = u1[i]; W[kW + c0 + i] = u2[i]; }
      }
      ab_prefetch();
      hbar();
      abmul(u1, Wq, W, T(1), tmp);
      publish(V, tmp);
      pd_prefetch();
      __syncwarp();
      pd_mul(V, tmp);
#pragma unroll
      for (int r = 0; r < RPT; ++r) rt[r] = yv[r] - tmp[r];
    };
#pragma unroll
    for (int r = 0; r < RPT; ++r) { rr[r] = live ? d.gam[(size_t)(i0 + r) * K + tj] : T(0); xx[r] = T(0); }
    pd_prefetch();
    precond();
    T part = T(0);
#pragma unroll
    for (int r = 0; r < RPT; ++r) { pp[r] = rt[r]; part += rr[r] * rt[r]; }
    T nu = bsum(part);
    if (tid == 0 && d.nu_trace) d.nu_trace[(size_t)b * NU_TRACE_LEN] = fabs(nu);
    int iters = 0;
    for (int it = 0; it < max_iter; ++it) {
      const T inv_nu = T(1) / nu;
      publish(V, pp);
      ab_prefetch();
      hbar();
      T uc[MC], hu = T(0), ap[RPT];
      abt(V, uc);
#pragma unroll
      for (int i = 0; i < MC; ++i) {
        uc[i] = (sval[i] ? V[kV + sidx[i]] : T(0)) - uc[i];
        hu += hh[i] * uc[i];
      }
      hu = rank1 ? sS * quad(hu) : T(0);
#pragma unroll
      for (int i = 0; i < MC; ++i) {
        uc[i] = ghat_a(dinv[i], uc[i], hh[i], hu);
        if (cvalid[i]) W[kW + c0 + i] = uc[i];
      }
      ab_prefetch();
      hbar();
      abmul(uc, W, W, T(-1), ap);
      part = T(0);
#pragma unroll
      for (int r = 0; r < RPT; ++r) part += pp[r] * ap[r];
      const T pAp = bsum(part);
      pd_prefetch();                              // behind the division
      const T alpha = nu / pAp;
#pragma unroll
      for (int r = 0; r < RPT; ++r) { rr[r] -= ap[r] * alpha; xx[r] += pp[r] * alpha; }
      precond();
      part = T(0);
#pragma unroll
      for (int r = 0; r < RPT; ++r) part += rr[r] * rt[r];
      const T nu_prime = bsum(part);
      iters = it + 1;
      if (tid == 0 && d.nu_trace && iters < NU_TRACE_LEN) d.nu_trace[(size_t)b * NU_TRACE_LEN + iters] = fabs(nu_prime);
      if (fabs(nu_prime) < tol) break;
      const T beta = nu_prime * inv_nu;
#pragma unroll
      for (int r = 0; r < RPT; ++r) pp[r] = rt[r] + pp[r] * beta;
      nu = nu_prime;
    }
    if (live) {
#pragma unroll
      for (int r = 0; r < RPT; ++r) d.l[(size_t)(i0 + r) * K + tj] = xx[r];
    }
    if (tid == 0) {
      d.pcg_iters[b] = iters;
      d.tot_pcg[b] += iters;
      d.tot_qp[b] += 1;
    }
    hbar();      // every read of this instance's shared vectors is complete before the next instance zeroes them
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tm_base));
}

}  // namespace b2t
