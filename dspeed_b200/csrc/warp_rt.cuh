// dspeed_b200 -- device routines of the warp-per-row chain kernels (dspeed_b200/warpchain.py).
//
// Short waveforms (<= 2048 samples, the SiPM / LAr chains of BASELINE.json config 4): ONE WARP owns
// one waveform.  Lane l holds samples [CH*l, CH*l + CH) in registers (CH = 8..64, a power of two),
// recursive filters are lane-local running sums + one warp scan, halos travel by shuffles, and the
// only shared memory is (a) the raw row of the NEXT waveform, staged while this one is processed, and
// (b) one float copy of the wave the peak finder walks -- aliased with (a) when both exist, the next
// row then travels through registers (RowPieces).  No block-wide barrier anywhere: warps run rows
// independently, up to 24 warps per SM at CH = 64.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "dspeed_b200.h"

namespace wrt {

constexpr unsigned FULL = 0xffffffffu;

// sample i of a wave slot: lane chunks are CH + 4 words apart (128-bit accesses of the 32 chunks
// and unit-stride accesses inside a chunk are both conflict-free)
template <int CH>
__device__ __forceinline__ int widx(int i) {
  return (i / CH) * (CH + 4) + (i % CH);
}

// ---------------------------------------------------------------------------------------
// raw row staging: 16-byte cp.async pieces, chunk c of the row at byte c * (2 CH + 16)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// n % 8 == 0 and 16-byte aligned rows (checked by the launcher)
template <int CH>
__device__ __forceinline__ void stage_row_16(unsigned char* raw, const uint16_t* g, int n, int lane) {
  constexpr int PPC = CH / 8;  // 16-byte pieces per chunk
  const int np = n >> 3;
  for (int q = lane; q < np; q += 32) cp_async16(raw + (q / PPC) * (2 * CH + 16) + (q % PPC) * 16, g + 8 * q);
  cp_async_commit();
}
// The same pieces through registers, for kernels whose staging buffer is aliased with the wave copy of the peak
// finder: the next row is REQUESTED before the walk (the 64-sample register chunk is dead by then, so the pieces cost
// no extra registers) and WRITTEN to the staging buffer after it.  N = samples per row (compile time: the loops unroll).
template <int N>
struct RowPieces {
  static constexpr int NP = N >> 3, K = (NP + 31) / 32;
  uint4 q[K];
};
template <int N>
__device__ __forceinline__ void fetch_row_16(RowPieces<N>& r, const uint16_t* g, int lane, bool on) {
#pragma unroll
  for (int k = 0; k < RowPieces<N>::K; k++) {
    const int q = lane + 32 * k;
    r.q[k] = make_uint4(0u, 0u, 0u, 0u);
    if (on && q < RowPieces<N>::NP)
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(r.q[k].x), "=r"(r.q[k].y), "=r"(r.q[k].z), "=r"(r.q[k].w)
                   : "l"(reinterpret_cast<const uint4*>(g) + q));
  }
}
template <int CH, int N>
__device__ __forceinline__ void stage_pieces_16(unsigned char* raw, const RowPieces<N>& r, int lane) {
  constexpr int PPC = CH / 8;
#pragma unroll
  for (int k = 0; k < RowPieces<N>::K; k++) {
    const int q = lane + 32 * k;
    if (q < RowPieces<N>::NP) *reinterpret_cast<uint4*>(raw + (q / PPC) * (2 * CH + 16) + (q % PPC) * 16) = r.q[k];
  }
}
template <bool SIGNED>
__device__ __forceinline__ float cvt16(unsigned h) {
  return SIGNED ? (float)(short)(unsigned short)h : (float)h;
}
// own chunk of the staged row -> registers (positions >= n hold finite garbage, masked by the users)
template <int CH, bool SIGNED>
__device__ __forceinline__ void read_chunk_16(const unsigned char* raw, int lane, float (&x)[CH]) {
  const uint4* p = reinterpret_cast<const uint4*>(raw + lane * (2 * CH + 16));
#pragma unroll
  for (int k = 0; k < CH / 8; k++) {
    const uint4 q = p[k];
    x[8 * k + 0] = cvt16<SIGNED>(q.x & 0xffffu); x[8 * k + 1] = cvt16<SIGNED>(q.x >> 16);
    x[8 * k + 2] = cvt16<SIGNED>(q.y & 0xffffu); x[8 * k + 3] = cvt16<SIGNED>(q.y >> 16);
    x[8 * k + 4] = cvt16<SIGNED>(q.z & 0xffffu); x[8 * k + 5] = cvt16<SIGNED>(q.z >> 16);
    x[8 * k + 6] = cvt16<SIGNED>(q.w & 0xffffu); x[8 * k + 7] = cvt16<SIGNED>(q.w >> 16);
  }
}

// own chunk <-> wave slot (128-bit, conflict-free)
template <int CH>
__device__ __forceinline__ void st_chunk(float* S, int lane, const float (&x)[CH]) {
  float4* p = reinterpret_cast<float4*>(S + lane * (CH + 4));
#pragma unroll
  for (int k = 0; k < CH / 4; k++) p[k] = make_float4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
}
// own chunk -> a global row (waveform-valued outputs)
template <int CH>
__device__ __forceinline__ void stg_chunk(float* g, int lane, int n, const float (&x)[CH], bool nan) {
#pragma unroll
  for (int k = 0; k < CH / 4; k++) {
    const int i = CH * lane + 4 * k;
    if (i + 3 < n && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) {
      reinterpret_cast<float4*>(g + i)[0] = nan ? make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F)
                                                : make_float4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
    } else {
#pragma unroll
      for (int e = 0; e < 4; e++)
        if (i + e < n) g[i + e] = nan ? CUDART_NAN_F : x[4 * k + e];
    }
  }
}

// ---------------------------------------------------------------------------------------
// warp scans
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float wscan_excl_add(float v, int lane) {
  float s = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float t = __shfl_up_sync(FULL, s, d);
    if (lane >= d) s += t;
  }
  const float e = __shfl_up_sync(FULL, s, 1);
  return lane == 0 ? 0.f : e;
}
__device__ __forceinline__ float wscan_excl_add_rev(float v, int lane) {  // sum over the lanes ABOVE this one
  float s = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float t = __shfl_down_sync(FULL, s, d);
    if (lane + d < 32) s += t;
  }
  const float e = __shfl_down_sync(FULL, s, 1);
  return lane == 31 ? 0.f : e;
}
__device__ __forceinline__ float wscan_incl_max(float v, int lane) {
  // shfl.up hands lanes below `d` their own value back and max is idempotent: no lane predicate
  float s = v;
  (void)lane;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) s = fmaxf(s, __shfl_up_sync(FULL, s, d));
  return s;
}
__device__ __forceinline__ int wscan_incl_add_i(int v, int lane) {
  int s = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(FULL, s, d);
    if (lane >= d) s += t;
  }
  return s;
}

// ---------------------------------------------------------------------------------------
// bl_subtract.py:11-46
// ---------------------------------------------------------------------------------------
template <int CH>
__device__ __forceinline__ void bl_sub(const float (&x)[CH], float b, float (&y)[CH]) {
#pragma unroll
  for (int j = 0; j < CH; j++) y[j] = x[j] - b;
}

// ---------------------------------------------------------------------------------------
// Invariant of every wave in registers: positions >= n repeat sample n - 1 ("right-edge extended").
// It costs one pass when a wave is created with a new length and makes the window filters and the
// chunk summaries branch-free (the reference's edge padding x[min(i + L, n - 1)] comes for free).
// ---------------------------------------------------------------------------------------
template <int CH, int N>
__device__ __forceinline__ void edge_extend(float (&x)[CH], int lane) {
  if (N >= 32 * CH) return;
  constexpr int LAST = (N - 1) / CH, JL = (N - 1) % CH;   // lane / register of sample N - 1
  const float e = __shfl_sync(FULL, x[JL], LAST);
  const bool above = lane > LAST, from = lane >= LAST;    // one select per register, no branch
#pragma unroll
  for (int j = 0; j < CH; j++) x[j] = (j > JL ? from : above) ? e : x[j];
}

// ---------------------------------------------------------------------------------------
// moving_windows.py:12-114 : out[0] = x[0]; out[i] = out[i-1] + (x[i] - x[max(i-L, 0)]) / L and
// its mirror image.  Lane-local running sum of the increments + one warp scan; the L samples of
// the neighbouring chunk arrive by shuffles (L <= CH, compile time).  Lane 0 sees x[0] in its left
// halo and the right halo is the edge extension, so no position is special.
// ---------------------------------------------------------------------------------------
template <int CH, int L, int N>
__device__ __forceinline__ void mw_left(const float (&x)[CH], float (&y)[CH], float il, int lane) {
  static_assert(L >= 1 && L <= CH, "window longer than a lane chunk");
  const float e0 = __shfl_sync(FULL, x[0], 0);
  float h[L];
#pragma unroll
  for (int j = 0; j < L; j++) {
    const float t = __shfl_up_sync(FULL, x[CH - L + j], 1);
    h[j] = lane == 0 ? e0 : t;
  }
  float run = 0.f;
#pragma unroll
  for (int j = 0; j < CH; j++) {
    const float prev = j >= L ? x[j >= L ? j - L : 0] : h[j < L ? j : 0];
    run += (x[j] - prev) * il;
    y[j] = run;
  }
  const float off = wscan_excl_add(run, lane) + e0;
#pragma unroll
  for (int j = 0; j < CH; j++) y[j] += off;
  edge_extend<CH, N>(y, lane);   // the increments beyond n - 1 are not zero: restore the invariant
}
template <int CH, int L, int N>
__device__ __forceinline__ void mw_right(const float (&x)[CH], float (&y)[CH], float il, int lane) {
  static_assert(L >= 1 && L <= CH, "window longer than a lane chunk");
  const float e0 = __shfl_sync(FULL, x[(N - 1) % CH], (N - 1) / CH);
  float h[L];
#pragma unroll
  for (int j = 0; j < L; j++) {
    const float t = __shfl_down_sync(FULL, x[j], 1);
    h[j] = lane == 31 ? e0 : t;
  }
  float run = 0.f;
#pragma unroll
  for (int j = CH - 1; j >= 0; j--) {
    const float next = j + L < CH ? x[j + L < CH ? j + L : 0] : h[j + L >= CH ? j + L - CH : 0];
    run += (x[j] - next) * il;
    y[j] = run;
  }
  const float off = wscan_excl_add_rev(run, lane) + e0;
#pragma unroll
  for (int j = 0; j < CH; j++) y[j] += off;
}

// avg_current (moving_windows.py:206-249): out[i] = (x[i + L] - x[i]) / L, n - L samples
template <int CH, int L>
__device__ __forceinline__ void avg_current(const float (&x)[CH], float (&y)[CH], float il, int lane) {
  static_assert(L >= 1 && L <= CH, "window longer than a lane chunk");
  float h[L];
#pragma unroll
  for (int j = 0; j < L; j++) h[j] = __shfl_down_sync(FULL, x[j], 1);
#pragma unroll
  for (int j = 0; j < CH; j++) {
    const float next = j + L < CH ? x[j + L < CH ? j + L : 0] : h[j + L >= CH ? j + L - CH : 0];
    y[j] = (next - x[j]) * il;
  }
}

// ---------------------------------------------------------------------------------------
// min_max.py:11-82 over [lo, hi): first arg-min / arg-max and the values
// ---------------------------------------------------------------------------------------
template <int CH>
__device__ __forceinline__ void min_max(const float (&x)[CH], int lo, int hi, int lane, float& t_min, float& t_max,
                                        float& a_min, float& a_max) {
  float vmin = CUDART_INF_F, vmax = -CUDART_INF_F;
  int imin = 0x7fffffff, imax = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < CH; j++) {
    const int i = CH * lane + j;
    const bool in = i >= lo && i < hi;
    if (in && x[j] < vmin) { vmin = x[j]; imin = i; }
    if (in && x[j] > vmax) { vmax = x[j]; imax = i; }
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    const float om = __shfl_xor_sync(FULL, vmin, d), oM = __shfl_xor_sync(FULL, vmax, d);
    const int oi = __shfl_xor_sync(FULL, imin, d), oI = __shfl_xor_sync(FULL, imax, d);
    if (om < vmin || (om == vmin && oi < imin)) { vmin = om; imin = oi; }
    if (oM > vmax || (oM == vmax && oI < imax)) { vmax = oM; imax = oI; }
  }
  t_min = (float)(imin - lo); t_max = (float)(imax - lo); a_min = vmin; a_max = vmax;
}

// ---------------------------------------------------------------------------------------
// get_multi_local_extrema.py:12-306 -- the peak-detection state machine, walked by the whole
// warp.  The machine only changes state at an *event* (a sample that drops a_delta below the
// running maximum, or rises a_delta above the running minimum), so the warp jumps from event to
// event instead of visiting samples:
//   * lane chunks carry (max, min) summaries; with the running extreme carried in by a prefix-max
//     scan over the chunks, "no sample of this chunk can fire" is decided for 32 chunks at once
//     (min < fl(max(carry, chunk max) - delta) is necessary for an event inside the chunk);
//   * the first chunk that may fire is examined exactly, 32 samples at a time: prefix max over the
//     lanes, the reference's float32 comparison v < fl(M - delta) per sample, ballot -> first event.
// The find-min state is the find-max state on negated samples (negation is exact), the
// right-to-left walk is the same walk on mirrored positions.  Extrema are recorded as bits of the
// lane that owns the sample (bit k of lane l <-> sample CH*l + k), so merging the two search
// directions (aggressive search = union) and sorting are free.
// ---------------------------------------------------------------------------------------
// per-window summaries (a window = 32 samples, or the whole lane chunk when it is shorter): extreme
// values, the largest fall (v[i] - v[j], i < j) and the largest rise (v[j] - v[i], i < j) inside it
struct WinSumm {
  float vmax, vmin, fall, rise;
};
template <int CH>
struct ChunkSumm {
  static constexpr int WS = CH < 32 ? CH : 32;   // window size
  static constexpr int NWL = CH / WS;            // windows per lane
  WinSumm w[NWL];
};
template <int CH>
__device__ __forceinline__ ChunkSumm<CH> chunk_summary(const float (&x)[CH]) {
  // (edge-extended waves: the repeated last sample changes none of the four)
  ChunkSumm<CH> s;
#pragma unroll
  for (int h = 0; h < ChunkSumm<CH>::NWL; h++) {
    WinSumm a{-CUDART_INF_F, CUDART_INF_F, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < ChunkSumm<CH>::WS; k++) {
      const float v = x[h * ChunkSumm<CH>::WS + k];
      a.vmax = fmaxf(a.vmax, v);
      a.vmin = fminf(a.vmin, v);
      a.fall = fmaxf(a.fall, a.vmax - v);
      a.rise = fmaxf(a.rise, v - a.vmin);
    }
    s.w[h] = a;
  }
  return s;
}

template <int CH, bool BWD>
__device__ __forceinline__ void peak_walk(const float* S, int n, float d_max, float d_min, float a_max, float a_min, int m,
                                            const ChunkSumm<CH>& cs, unsigned long long& bmax, unsigned long long& bmin,
                                            int& n_found_max, int& n_found_min) {
  constexpr int NTOT = 32 * CH;
  constexpr int WS = ChunkSumm<CH>::WS, NWL = ChunkSumm<CH>::NWL;
  const int lane = threadIdx.x & 31;
  // summaries in walk order: logical window lane * NWL + h.  Walking backward, the largest drop below a
  // running maximum that starts INSIDE a window is a rise of the samples (and a fall for the find-min state).
  float lmax[NWL], lmin[NWL], ldrop[NWL], lclimb[NWL], slack[NWL];
#pragma unroll
  for (int h = 0; h < NWL; h++) {
    const WinSumm& w = cs.w[BWD ? NWL - 1 - h : h];
    lmax[h] = BWD ? __shfl_sync(FULL, w.vmax, 31 - lane) : w.vmax;
    lmin[h] = BWD ? __shfl_sync(FULL, w.vmin, 31 - lane) : w.vmin;
    ldrop[h] = BWD ? __shfl_sync(FULL, w.rise, 31 - lane) : w.fall;
    lclimb[h] = BWD ? __shfl_sync(FULL, w.fall, 31 - lane) : w.rise;
    // float32 slack of "v < fl(M - delta)" against "fl(M - v) > delta" (only used to SKIP windows)
    slack[h] = 0x1p-20f * (fabsf(lmax[h]) + fabsf(lmin[h]) + d_max + d_min);
  }
  int p = BWD ? NTOT - n : 0;  // walk position; sample index = BWD ? NTOT - 1 - p : p
  const int p_begin = p;
  const int p_end = BWD ? NTOT : n;
#define WRT_IDX(q) (BWD ? NTOT - 1 - (q) : (q))
  bool mode_max = true;
  int ei = WRT_IDX(p);
  float ev = S[widx<CH>(ei)];
  p += 1;
  // Everything that depends on the state (find-max / find-min) is kept as a loop-carried "current" and "other" copy
  // and SWAPPED when an event flips the state, instead of being selected in every iteration: the sign, the deltas,
  // the absolute thresholds, the list counters and the signed window summaries (max' = -min, min' = -max).
  float sg = 1.f, dl = d_max, dl_o = d_min, ab = a_max, ab_o = -a_min;
  int cc = 0, cc_o = 0;                       // events recorded in the current / the other state
  float smx[NWL], smn[NWL], drp[NWL], drp_o[NWL], thr[NWL], thr_o[NWL];
#pragma unroll
  for (int h = 0; h < NWL; h++) {
    smx[h] = lmax[h];
    smn[h] = lmin[h];
    drp[h] = ldrop[h];
    drp_o[h] = lclimb[h];
    thr[h] = d_max - slack[h];
    thr_o[h] = d_min - slack[h];
  }
  while (p < p_end) {
    if (cc >= m) break;  // the list is full: this state never fires again
    const int W = p / WS;
    // ---- exact examination of the rest of window W ------------------------------------------------
    {
      const int q0 = W * WS;
      const int q = q0 + lane;
      const bool valid = q >= p && lane < WS && q < p_end;
      const float v = valid ? sg * S[widx<CH>(WRT_IDX(q))] : -CUDART_INF_F;
      const float pm = wscan_incl_max(v, lane);
      const float M = fmaxf(ev, pm);
      const bool cross = valid && (v < M - dl) && (M > ab);
      const unsigned b = __ballot_sync(FULL, cross);
      if (b) {
        const int ls = __ffs(b) - 1;
        const float Ms = __shfl_sync(FULL, M, ls);
        if (Ms > ev) {  // the running extreme lies in this window: its first occurrence
          const unsigned e = __ballot_sync(FULL, valid && v == Ms);
          ei = WRT_IDX(q0 + __ffs(e) - 1);
        }
        if (lane == ei / CH) {
          if (mode_max) bmax |= 1ull << (ei % CH);
          else bmin |= 1ull << (ei % CH);
        }
        ev = -__shfl_sync(FULL, v, ls);  // the event sample starts the opposite search
        ei = WRT_IDX(q0 + ls);
        p = q0 + ls + 1;
        // flip the state: swap the current and the other copies
        mode_max = !mode_max;
        sg = -sg;
        { const int t = cc + 1; cc = cc_o; cc_o = t; }
        { const float t = dl; dl = dl_o; dl_o = t; }
        { const float t = ab; ab = ab_o; ab_o = t; }
#pragma unroll
        for (int h = 0; h < NWL; h++) {
          const float t = smx[h];
          smx[h] = -smn[h];
          smn[h] = -t;
          const float u = drp[h]; drp[h] = drp_o[h]; drp_o[h] = u;
          const float w = thr[h]; thr[h] = thr_o[h]; thr_o[h] = w;
        }
        continue;
      }
      const float wm = __shfl_sync(FULL, pm, 31);
      if (wm > ev) {
        const unsigned e = __ballot_sync(FULL, valid && v == wm);
        ei = WRT_IDX(q0 + __ffs(e) - 1);
        ev = wm;
      }
    }
    // ---- next window that may fire (all windows at once) --------------------------------------------
    float pmw[NWL];
    bool act[NWL];
    float tot = -CUDART_INF_F;
#pragma unroll
    for (int h = 0; h < NWL; h++) {
      act[h] = lane * NWL + h > W;
      pmw[h] = act[h] ? smx[h] : -CUDART_INF_F;
      tot = fmaxf(tot, pmw[h]);
    }
    const float ex = __shfl_up_sync(FULL, wscan_incl_max(tot, lane), 1);
    float m_in[NWL];
    int first = 0x7fffffff;
    m_in[0] = fmaxf(ev, lane == 0 ? -CUDART_INF_F : ex);  // running extreme on entry of the lane's first window
#pragma unroll
    for (int h = 0; h < NWL; h++) {
      if (h > 0) m_in[h] = fmaxf(m_in[h - 1], pmw[h - 1]);
      const float m_full = fmaxf(m_in[h], smx[h]);
      // an event inside the window needs a sample below fl(carry - delta), or a drop of (almost) delta below
      // a maximum of the window itself
      const unsigned b = __ballot_sync(FULL, act[h] && (m_full > ab) && ((smn[h] < m_in[h] - dl) || (drp[h] > thr[h])));
      if (b) first = min(first, (__ffs(b) - 1) * NWL + h);
    }
    if (first == 0x7fffffff) break;
    const int W2 = first;
    float nev = __shfl_sync(FULL, m_in[0], W2 / NWL);
#pragma unroll
    for (int h = 1; h < NWL; h++) {
      const float t = __shfl_sync(FULL, m_in[h], W2 / NWL);
      if (W2 % NWL == h) nev = t;
    }
    if (nev > ev) {  // the extreme moved into one of the skipped windows: first occurrence of it
      int k = 0x7fffffff;
#pragma unroll
      for (int h = 0; h < NWL; h++) {
        const unsigned kb = __ballot_sync(FULL, act[h] && lane * NWL + h < W2 && smx[h] == nev);
        if (kb) k = min(k, (__ffs(kb) - 1) * NWL + h);
      }
      const int q = k * WS + lane;
      const bool valid = lane < WS && q < p_end && q >= p_begin;
      const float v = valid ? sg * S[widx<CH>(WRT_IDX(q))] : -CUDART_INF_F;
      const unsigned e = __ballot_sync(FULL, valid && v == nev);
      ei = WRT_IDX(k * WS + __ffs(e) - 1);
      ev = nev;
    }
    p = W2 * WS;
  }
#undef WRT_IDX
  n_found_max = mode_max ? cc : cc_o;
  n_found_min = mode_max ? cc_o : cc;
}

// extrema bit sets -> one list element per lane (NaN-padded, m <= 32), ascending or descending
// positions; `buf`: 32 floats of shared memory of this warp.  Returns the number of entries.
template <int CH>
__device__ __forceinline__ int emit_list(unsigned long long bits, int m, bool descending, float* buf, int lane,
                                         float& elem) {
  const int cnt = __popcll(bits);
  const int incl = wscan_incl_add_i(cnt, lane);
  const int total = __shfl_sync(FULL, incl, 31);
  int pos = incl - cnt;
  buf[lane] = CUDART_NAN_F;
  __syncwarp();
  while (bits) {
    const int k = __ffsll((long long)bits) - 1;
    bits &= bits - 1;
    const int o = descending ? total - 1 - pos : pos;
    if (o >= 0 && o < m) buf[o] = (float)(CH * lane + k);
    pos++;
  }
  __syncwarp();
  elem = buf[lane];
  __syncwarp();
  return min(total, m);
}

}  // namespace wrt
