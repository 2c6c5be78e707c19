// fold_kernel2.cuh — v2 batch kernel: one CTA per sequence, warp ROLES pipelined across anti-diagonals
// (see fold_phases.cuh for the phases and their dependencies).  Storage modes:
//   MODE_SMEM    matrices, bit matrix and cell lists resident in shared memory   (tRNA .. ~150 nt)
//   MODE_GLOBAL  the same working set in an HBM/L2 workspace slot of the CTA     (Rfam-length batches)
// Sequences longer than 1024 nt use the cooperative multi-CTA wavefront of fold_kernel.cuh.
#pragma once
#include "fold_kernel.cuh"
#include "fold_phases.cuh"

// prefetch depths of the shared-memory mode's chains whose operands live in the CTA's L2 slot (split points / terms
// in flight ahead of the fold)
#ifndef RNA_Z_PF
#define RNA_Z_PF 2
#endif
#ifndef RNA_ML_PF
#define RNA_ML_PF 2
#endif
#ifndef RNA_Y_CH
#define RNA_Y_CH 1
#endif

namespace rna {

template <int MODE> struct PIdxOf { typedef uint16_t type; };
template <> struct PIdxOf<MODE_SMEM> { typedef uint8_t type; };   // SMEM mode is only used for L <= 255

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) / 16 * 16; }

// bytes of the table / sequence part that always lives in shared memory
template <bool CONTRA>
__host__ __device__ inline size_t fold2_fixed_bytes(int Lcap) {
  return 128 + align16(sizeof(typename Model2<CONTRA>::Small)) + align16((size_t)Lcap + 8);
}
// upper bound of the number of lane groups of a sequence of length <= Lcap: sum over diagonals of ceil(cells/32)
__host__ __device__ inline size_t fold2_ngcap(int Lcap) {
  size_t n = 0;
  for (int c = 1; c <= Lcap; c++) n += (size_t)(c + 31) / 32;
  return n;
}
// bytes of one sequence's working set (SeqViewT) for capacity Lcap
// nmat = matrices inside this region: 2 (C, log P: shared-memory mode, the dense matrices live in the CTA's
// HBM/L2 slot) or 5 (everything in one space, log P aliasing E)
__host__ __device__ inline size_t fold2_seq_bytes(int Lcap, size_t pidx_size, int nmat) {
  const size_t T = (size_t)Lcap * ((size_t)Lcap + 1) / 2;
  const size_t W2 = ((size_t)Lcap + 31) / 32 + 2;
  size_t b = (size_t)nmat * T * 4;
  b += (3 + 2) * (size_t)Lcap * 4;              // Mroll, E0, EL
  b += align16(2 * ((size_t)Lcap + 2) * 4);     // traceback stack
  b += align16((size_t)Lcap * W2 * 4);          // closable bit matrix
  b += align16((size_t)Lcap * 4 + 4);           // pcnt, RR, LL
  b += align16(4 * ((size_t)Lcap / 2 + 3) * 4);    // gcumI, gcumO, ccumI, ccumO
  b += align16(2 * (fold2_ngcap(Lcap) + 2) * 4);   // gbin, gbout
  b += align16(2 * (fold2_ngcap(Lcap) + 2) * 2);   // gstepI, gstepO
  b += align16(T * pidx_size);                  // closable-cell lists
  return align16(b);
}

// bytes of one CTA's term-stream slot
__host__ __device__ inline size_t fold2_stream_bytes(int Lcap, uint32_t tcap) {
  const size_t Tc = (size_t)Lcap * ((size_t)Lcap + 1) / 2;
  (void)Tc;
  return align16(2 * (size_t)tcap * 8);
}

struct Roles { int nX, nY, nZ; };   // warps per role
// X (the latency-critical two-loop chains) gets a lane for every closable cell of a step (two diagonals, ~0.37 of
// the cells each); the dense chains (Z) and the sparse rightmost-pair chains (Y, CONTRAfold) share the rest and
// stride over their cells when they have fewer lanes than cells.
__host__ __device__ inline int fold2_x_warps(int Lcap) { return (3 * Lcap / 4 + 31) / 32; }
__host__ __device__ inline int fold2_min_warps(int Lcap, bool contra) { return fold2_x_warps(Lcap) + (contra ? 4 : 2); }
__host__ __device__ inline Roles fold2_roles(int Lcap, bool contra, int max_warps) {
  Roles r;
  const int full = (Lcap + 31) / 32;
  r.nX = fold2_x_warps(Lcap);
  if (r.nX > max_warps - (contra ? 2 : 1)) r.nX = max_warps - (contra ? 2 : 1) > 1 ? max_warps - (contra ? 2 : 1) : 1;
  int rest = max_warps - r.nX;
  if (rest < (contra ? 2 : 1)) rest = contra ? 2 : 1;
  if (contra) {
    r.nZ = (rest + 1) / 2 < full ? (rest + 1) / 2 : full;
    r.nY = rest - r.nZ < full ? rest - r.nZ : full;
    if (r.nY < 1) r.nY = 1;
  } else {
    r.nZ = rest < full ? rest : full;
    r.nY = 0;
  }
  return r;
}

// Roles of the HBM-resident one-CTA mode (sequences beyond the shared-memory mode, up to 1024 nt).  There the dense
// chains dominate: per sequence the three chains of role Z walk L^3/6 split points, the sparse rightmost-pair chains
// of role Y (CONTRAfold) ~0.12 L^3 terms, the two-loop chains of role X ~37 L^2 terms (scored on the fly), so the
// lanes are split in proportion to that work instead of "a lane per closable cell for X".
__host__ __device__ inline Roles fold2_roles_global(int Lcap, bool contra, int max_warps) {
  const double L = Lcap;
  // (fitted to the role timers with the shared-memory operand rings: 6/4/6 warps at 460 nt)
  const double wx = 21500.0 * L * L, wz = 50.0 * L * L * L, wy = contra ? 30.0 * L * L * L : 0.0;
  const double tot = wx + wy + wz;
  Roles r;
  r.nY = contra ? (int)(max_warps * wy / tot + 0.5) : 0;
  if (contra && r.nY < 1) r.nY = 1;
  r.nX = (int)(max_warps * wx / tot + 0.5);
  if (r.nX < 2) r.nX = 2;
  r.nZ = max_warps - r.nX - r.nY;
  if (r.nZ < 2) { r.nZ = 2; r.nX = max_warps - r.nY - r.nZ; }
  return r;
}
// X warps of the OUTSIDE pass in the HBM-resident mode.  Phase 1 runs X (exterior + enclosing two-loops, scored on the
// fly: ~L^2) beside Y (probs_multibranch(2): ~L^3) on the other warps; phase 2 is X alone (the multiloop chain: ~L^3).
__host__ __device__ inline int fold2_xwarps_outside_global(int Lcap, int max_warps) {
  const double L = Lcap;
  // (X's two-loop chains stop scaling once every closable cell of a step has a lane, so the fit is deliberately biased
  // towards Y: 8 of 16 warps at 460 nt, 9 at 300 nt measured best; 9-12 X warps at 460 nt are 1-5 % slower)
  const double wx = 3.08e-3 * L * L, wy = 7.49e-6 * L * L * L, wm = 3.88e-6 * L * L * L;
  int best = 2;
  double bt = 1e300;
  for (int nx = 2; nx <= max_warps - 2; nx++) {
    const double p1 = wx / nx > wy / (max_warps - nx) ? wx / nx : wy / (max_warps - nx);
    const double t = p1 + wm / nx;
    if (t < bt) { bt = t; best = nx; }
  }
  return best;
}

// SUMS: the build that also exports the FoldSums / FoldScores planes (rna_fold_sums_batch) and can stop after the inside
// pass; the default build carries none of it (the batch kernel sits at its 64-register cap: measured 10 % slower with
// the export code compiled in).
template <bool CONTRA, int MODE, bool SUMS = false>
__global__ void __launch_bounds__(512, 2) fold_kernel2(const FoldArgs a) {
  typedef typename Model2<CONTRA>::Dev Dev;
  typedef typename Model2<CONTRA>::Small Small;
  typedef typename Model2<CONTRA>::View View;
  typedef typename PIdxOf<MODE>::type PIdx;
  typedef SeqViewT<PIdx> SV;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* lut = reinterpret_cast<float4*>(smem_raw);
  Small* small = reinterpret_cast<Small*>(smem_raw + 128);
  uint8_t* sseq = smem_raw + 128 + align16(sizeof(Small));
  unsigned char* sregion = smem_raw + fold2_fixed_bytes<CONTRA>(a.Lcap);
  __shared__ int s_work, s_fill_next;

  const Dev* dev = reinterpret_cast<const Dev*>(a.tables);
  const int tid = threadIdx.x, nt = blockDim.x;
  load_lse_lut(lut);
  {
    const Small* gs = dev_small<CONTRA>(dev);
    for (int x = tid; x < (int)(sizeof(Small) / 4); x += nt)
      reinterpret_cast<float*>(small)[x] = reinterpret_cast<const float*>(gs)[x];
  }
  View T;
  T.g = dev;
  T.sm = small;
  ModelParams P;
  P.MINSPAN = dev->min_span;
  if constexpr (CONTRA) P.MAX2 = dev->max_loop_len; else P.MAX2 = dev->max_2loop_len;
  P.allows_short = a.allows_short;
  const float NEG = RNA_NEG_INF;
  const int warp = tid >> 5;
  const int nXl = a.nXw * 32, nYl = a.nYw * 32, nZl = a.nZw * 32;
  __syncthreads();

  for (;;) {
    __syncthreads();
    if (tid == 0) s_work = atomicAdd(a.work_counter, 1);
    __syncthreads();
    const uint32_t w = (uint32_t)s_work;
    if (w >= a.n_launch) break;
    const uint32_t sidx = a.order ? a.order[w] : w;
    const uint32_t sbeg = a.offsets[sidx];
    const int L = (int)(a.offsets[sidx + 1] - sbeg);
    const int TRI = L * (L + 1) / 2;
    const int ngcap = (int)fold2_ngcap(a.Lcap);

    // ---- carve the working set (sizes by L: every pointer stays inside the Lcap-sized region) -----------
    unsigned char* base = (MODE == MODE_SMEM) ? sregion
                                              : reinterpret_cast<unsigned char*>(a.workspace) + (size_t)blockIdx.x * a.ws_stride * 4;
    SV v;
    v.L = L;
    v.W2 = (L + 31) / 32 + 2;
    uint8_t* s = sseq + 4;
    v.s = s;
    float* f = reinterpret_cast<float*>(base);
    v.C = f; f += TRI;
    if (MODE == MODE_SMEM) {
      // on chip: sums_close and log P (gathered by the latency-critical chains); in the CTA's HBM/L2 slot: the
      // dense matrices, read with affine addresses that the phases prefetch
      v.Pm = f; f += TRI;
      float* g = a.workspace + (size_t)blockIdx.x * a.ws_stride;
      v.R = g; v.X = g + TRI; v.E = g + 2 * (size_t)TRI; v.M1 = g + 3 * (size_t)TRI;
    } else {
      v.R = f; f += TRI;
      v.X = f; f += TRI;
      v.E = f; f += TRI;
      v.M1 = f; f += TRI;
      v.Pm = v.E;
    }
    v.Mroll = f; f += 3 * L;
    v.E0 = f; f += L;
    v.EL = f; f += L;
    int* tstack = reinterpret_cast<int*>(f);
    v.mask = reinterpret_cast<uint32_t*>(tstack + 2 * (L + 2));
    v.gcumI = v.mask + L * v.W2;                                 // 4-byte items first, then 2-byte, then bytes
    v.gcumO = v.gcumI + L / 2 + 3;
    v.ccumI = v.gcumO + L / 2 + 3;
    v.ccumO = v.ccumI + L / 2 + 3;
    v.gbin = v.ccumO + L / 2 + 3;
    v.gbout = v.gbin + ngcap + 2;
    v.gstepI = reinterpret_cast<uint16_t*>(v.gbout + ngcap + 2);
    v.gstepO = v.gstepI + ngcap + 2;
    v.pcnt = v.gstepO + ngcap + 2;
    v.din0 = CONTRA ? 0 : (P.MINSPAN - 1);
    v.dout0 = CONTRA ? (a.allows_short ? 1 : P.MINSPAN - 1) : (P.MINSPAN - 1);
    v.plist = reinterpret_cast<PIdx*>(v.pcnt + ((L + 1) & ~1));   // (u8 or u16: 2-byte aligned)
    v.RR = reinterpret_cast<uint8_t*>(v.plist + TRI);
    v.LL = v.RR + L;
    v.tin = nullptr; v.tout = nullptr; v.ccnt = nullptr; v.tcap = a.tcap;
    float* sums_out = nullptr;   // rna_fold_sums_batch planes of this sequence
    if constexpr (SUMS) sums_out = a.sums ? a.sums + a.sums_offsets[sidx] : nullptr;
    v.Mfull = sums_out ? sums_out + (size_t)5 * TRI : nullptr;
    v.M1rm = nullptr; v.MB = nullptr;

    for (int x = tid; x < L; x += nt) s[x] = a.bases[sbeg + x];
    if (tid < 4) { sseq[tid] = 0; s[L + tid] = 0; }
    const bool dbg_on = a.dbg && w == 3u * gridDim.x;   // a steady-state sequence (CTAs have desynchronised by then)
    long long tc0 = dbg_on ? clock64() : 0;
    const long long tseq0 = tc0;
    __syncthreads();
    for (int x = tid; x < L * v.W2; x += nt) setup_mask_word<CONTRA>(v, P, x);
    for (int x = tid; x < L; x += nt) setup_codes(v, x);
    __syncthreads();
    for (int d = tid; d < L; d += nt) setup_list_diag(v, d);
    __syncthreads();
    if (dbg_on && tid == 0) { a.dbg[2047 * 16 + 0] = clock64() - tc0; tc0 = clock64(); }
    // ---- two-loop term streams: count, group maxima, scan, fill (all threads; fold_phases.cuh "term streams")
    if (a.stream_ws) {
      if (tid == 0) { setup_groups(v); s_fill_next = 0; }
      v.ccnt = reinterpret_cast<uint16_t*>(v.C);   // scratch: 2 x TRI u16 = the not-yet-initialised C matrix
      __syncthreads();
      stream_count(v, P, tid, nt);
      __syncthreads();
      stream_groupmax(v, tid, nt);
      __syncthreads();
      if (tid == 0) stream_scan(v);
      __syncthreads();
      if (dbg_on && tid == 0) { a.dbg[2047 * 16 + 1] = clock64() - tc0; tc0 = clock64(); }
      const uint32_t NGI = v.gcumI[max(num_steps_inside(v), 0)], NGO = v.gcumO[max(num_steps_outside(v), 0)];
      if (v.gbin[NGI] <= a.tcap && v.gbout[NGO] <= a.tcap) {   // else: does not fit its slot, score on the fly
        unsigned char* sw = a.stream_ws + (size_t)blockIdx.x * a.stream_stride;
        v.tin = reinterpret_cast<uint2*>(sw);
        v.tout = v.tin + a.tcap;
        // one group per warp at a time, handed out dynamically, longest partner lists first
        const uint32_t ntask = stream_num_tasks(v);
        for (;;) {
          uint32_t tau = 0;
          if ((tid & 31) == 0) tau = (uint32_t)atomicAdd(&s_fill_next, 1);
          tau = __shfl_sync(0xffffffffu, tau, 0);
          if (tau >= ntask) break;
          stream_fill_task<CONTRA>(v, T, P, tau, tid & 31);
          __syncwarp();
        }
      }
      __syncthreads();
      if (dbg_on && tid == 0) { a.dbg[2047 * 16 + 3] = clock64() - tc0; a.dbg[2047 * 16 + 4] = v.gbin[NGI]; }
    }
    for (int x = tid; x < TRI; x += nt) { v.C[x] = NEG; v.R[x] = NEG; v.X[x] = NEG; v.E[x] = 0.f; v.M1[x] = NEG; }
    for (int x = tid; x < 3 * L; x += nt) v.Mroll[x] = NEG;
    __syncthreads();

    // ================================ inside, pair steps ================================================
    // phase 1: X = two-loop parts of sums_close(t), (t+1)  |  Z(t-2), Y(t-1); bar(Y,Z); Z(t-1), Y(t)
    // phase 2: X = closing multibranch terms of (t), (t+1)
    const int d_in0 = v.din0;
    const int nYZl = nYl + nZl;
    const bool helper = warp >= a.nXw + a.nYw + a.nZw;   // extra warps: only the all-thread phases (setup, fill, output)
    for (int st = 0; d_in0 + 2 * st <= L + 1; st++) {
      const int t = d_in0 + 2 * st;
      const long long c0 = dbg_on ? clock64() : 0;
      if (warp < a.nXw) {
        inside_X<CONTRA>(v, T, lut, P, st, tid, nXl);
      } else if (!helper) {
        const bool isY = warp < a.nXw + a.nYw;
        if constexpr (MODE == MODE_GLOBAL) {
          // HBM-resident mode: shared memory is free, the chains keep their operands in flight in it (ring variants of
          // fold_phases.cuh; inside pass: the Z lanes' ring columns first, then the Y lanes')
          float4* const zring = a.ring_bytes ? reinterpret_cast<float4*>(sregion) : nullptr;
          float4* const yring = zring ? zring + RNA_Z_RING * nZl : nullptr;
          for (int half = 0; half < 2; half++) {
            const int dy = t - 1 + half, dz = t - 2 + half;
            if (isY) {
              if constexpr (CONTRA) {
                if (dy >= d_in0 && dy < L) { if (yring) inside_Y_contra_ring(v, T, lut, dy, tid - nXl, nYl, 0, yring); else inside_Y_contra<4>(v, T, lut, dy, tid - nXl, nYl); }
              }
            } else if (dz >= d_in0 && dz < L) {
              if (zring) inside_Z_ring<CONTRA, SV, SUMS>(v, T, lut, dz, tid - nXl - nYl, nZl, zring);
              else inside_Z<CONTRA, 4, SV, SUMS>(v, T, lut, dz, tid - nXl - nYl, nZl);
            }
            if (half == 0) asm volatile("bar.sync 1, %0;" ::"r"(nYZl) : "memory");
          }
        } else {
        if (isY) { if constexpr (CONTRA) { if (t - 1 >= d_in0 && t - 1 < L) inside_Y_contra<(MODE == MODE_SMEM ? RNA_Y_CH : 4)>(v, T, lut, t - 1, tid - nXl, nYl); } }
        else if (t - 2 >= d_in0 && t - 2 < L) inside_Z<CONTRA, (MODE == MODE_SMEM ? RNA_Z_PF : 4), SV, SUMS>(v, T, lut, t - 2, tid - nXl - nYl, nZl);
        asm volatile("bar.sync 1, %0;" ::"r"(nYZl) : "memory");
        if (isY) { if constexpr (CONTRA) { if (t < L) inside_Y_contra<(MODE == MODE_SMEM ? RNA_Y_CH : 4)>(v, T, lut, t, tid - nXl, nYl); } }
        else if (t - 1 >= d_in0 && t - 1 < L) inside_Z<CONTRA, (MODE == MODE_SMEM ? RNA_Z_PF : 4), SV, SUMS>(v, T, lut, t - 1, tid - nXl - nYl, nZl);
        }
      }
      if (dbg_on && (tid & 31) == 0 && !helper) a.dbg[(size_t)t * 16 + warp] = clock64() - c0;
      __syncthreads();
      if (warp < a.nXw) inside_X_fin<CONTRA>(v, T, lut, st, tid, nXl);
      __syncthreads();
    }

    // ================================ outside ===========================================================
    for (int x = tid; x < L; x += nt) {
      v.E0[x] = v.E[doff(x, L)];
      v.EL[x] = v.E[doff(L - 1 - x, L) + x];
    }
    if constexpr (SUMS) { if (sums_out) export_fold_sums<CONTRA>(v, T, P, sums_out, tid, nt); }   // FoldSums / FoldScores (N3)
    __syncthreads();
    const float Z = v.E0[L - 1];
    if (tid == 0 && a.out_logz) a.out_logz[sidx] = Z;
    if constexpr (SUMS) { if (a.inside_only) continue; }
    for (int x = tid; x < TRI; x += nt) { v.Pm[x] = NEG; v.R[x] = NEG; v.X[x] = NEG; }
    __syncthreads();
    const int d_out0 = v.dout0;
    // pair steps.  phase 1: X = exterior + two-loop parts of log P(d), (d-1)  |  Y = probs_multibranch(2) of d+1, d
    //             phase 2: X = multiloop parts
    if constexpr (MODE == MODE_GLOBAL) {
      // HBM-resident mode: its own split of the role warps for this pass (X also folds the multiloop chain here, Y takes
      // all the other role warps) and the ring variants of the chains
      float4* const zring = a.ring_bytes ? reinterpret_cast<float4*>(sregion) : nullptr;
      const int nXo = a.nXw_out > 0 ? a.nXw_out : a.nXw, nXol = nXo * 32, nYZol = nXl + nYZl - nXol;
      for (int st = 0; L - 1 - 2 * st >= d_out0; st++) {
        const int d = L - 1 - 2 * st;
        const long long c0 = dbg_on ? clock64() : 0;
        if (warp < nXo) {
          outside_X<CONTRA>(v, T, lut, P, Z, st, tid, nXol);
        } else if (!helper) {
          for (int dd = d + 1; dd >= d; dd--) {
            if (dd >= L) continue;
            if (zring) outside_Y_ring<CONTRA>(v, T, lut, dd, tid - nXol, nYZol, zring);
            else outside_Y<CONTRA, 4>(v, T, lut, dd, tid - nXol, nYZol);
          }
        }
        if (dbg_on && (tid & 31) == 0 && !helper) a.dbg[(size_t)(1024 + d) * 16 + warp] = clock64() - c0;
        __syncthreads();
        const long long c1 = dbg_on ? clock64() : 0;
        if (warp < nXo) {
          if (zring) outside_X_ml_ring<CONTRA>(v, T, lut, st, tid, nXol, zring);
          else outside_X_ml<CONTRA, 3>(v, T, lut, st, tid, nXol);
        }
        if (dbg_on && (tid & 31) == 0 && warp < nXo) a.dbg[(size_t)(1024 + d) * 16 + 8 + warp] = clock64() - c1;
        __syncthreads();
      }
    } else {
    for (int st = 0; L - 1 - 2 * st >= d_out0; st++) {
      const int d = L - 1 - 2 * st;
      const long long c0 = dbg_on ? clock64() : 0;
      if (warp < a.nXw) {
        outside_X<CONTRA>(v, T, lut, P, Z, st, tid, nXl);
      } else if (!helper) {
        if (d + 1 < L) outside_Y<CONTRA, (MODE == MODE_SMEM ? RNA_Y_CH : 4)>(v, T, lut, d + 1, tid - nXl, nYZl);
        outside_Y<CONTRA, (MODE == MODE_SMEM ? RNA_Y_CH : 4)>(v, T, lut, d, tid - nXl, nYZl);
      }
      if (dbg_on && (tid & 31) == 0 && !helper) a.dbg[(size_t)(1024 + d) * 16 + warp] = clock64() - c0;
      __syncthreads();
      const long long c1 = dbg_on ? clock64() : 0;
      if (warp < a.nXw) outside_X_ml<CONTRA, (MODE == MODE_SMEM ? RNA_ML_PF : 3)>(v, T, lut, st, tid, nXl);
      if (dbg_on && (tid & 31) == 0 && warp < a.nXw) a.dbg[(size_t)(1024 + d) * 16 + 8 + warp] = clock64() - c1;
      __syncthreads();
    }

    }

    const long long tpost0 = dbg_on ? clock64() : 0;
    // ================================ BPP = expf(P) =====================================================
    for (int x = tid; x < TRI; x += nt) {
      const float val = v.Pm[x];
      v.Pm[x] = (val > NEG) ? approx_expf(val) : -1.0f;
    }
    __syncthreads();
    if (a.out_bpp) {
      float* ob = a.out_bpp + a.bpp_offsets[sidx];
      for (int i = 0; i < L - 1; i++) {
        const size_t rowoff = (size_t)i * (size_t)(2 * L - i - 1) / 2;
        for (int x = tid; x < L - 1 - i; x += nt) ob[rowoff + x] = v.Pm[doff(x + 1, L) + i];
      }
    }
    // ================================ centroid (src/centroid_fold.rs:25-105) ============================
    {
      const float* Pm = v.Pm;
      auto getp = [=](int d, int i) -> float { return Pm[doff(d, L) + i]; };
      centroid_run<MODE>(a, sidx, sbeg, L, v.C, tstack, getp);   // W reuses sums_close (dead after the outside pass)
    }
    if (dbg_on && tid == 0) { a.dbg[2047 * 16 + 5] = clock64() - tpost0; a.dbg[2047 * 16 + 6] = clock64() - tseq0; a.dbg[2047 * 16 + 7] = L; }
  }
}

}  // namespace rna

namespace rna {

// =========================================================================================================
// fold_kernel2_coop — one LONG sequence at a time on the whole GPU (cooperative launch, HBM-resident working set).
// Same phases and roles as fold_kernel2; a role's lanes are spread over the grid (consecutive warps of a role sit
// on different SMs: the chains are latency-bound, so an SM per warp is the fastest placement), a step is ONE
// diagonal (whole folds, no partial sums across steps) and the barrier is grid-wide.  grid.sync() orders memory
// for every thread of the grid, so plain loads see what other CTAs wrote in earlier steps.
// =========================================================================================================
// =========================================================================================================
// Producer / consumer split of the multiloop part of log P (cooperative kernel, outside pass).  That chain is three
// dependent logsumexp's per split point and it is the critical path of the whole pass; a warp issues at most one
// instruction every other cycle, so every load, address step and operand add that sits in the chain's warp costs two
// cycles of critical path.  Here a PRODUCER warp (same CTA, another scheduler) walks the row-major matrices, forms
// the three operands of each split point and leaves them in a shared-memory double buffer; the CONSUMER warp's loop
// is nothing but LDS + logsumexp.  One named barrier (64 threads) per batch of RNA_ML_BATCH split points hands a
// buffer over; buffers alternate on a running batch counter that both warps advance identically, across cell groups
// and diagonals, so the producer is always exactly one batch ahead and never touches the buffer being read.
// =========================================================================================================
#define RNA_ML_BATCH 32
#define RNA_ML_RING_FLOATS (2 * RNA_ML_BATCH * 3 * 32)
// The consumer's folds of one batch, as a function of its own: inside the big kernel ptxas has two predicate registers
// left for the seven breakpoint compares (the rest hold long-lived flags), so it interleaves the four coefficient select
// trees level by level and re-materialises their sixteen register-side constants for every fold — 45 instructions and
// ~142 cycles per fold.  A separate function starts with all predicates and registers free.  The loop is NOT unrolled
// beyond four folds: the body stays inside the L0 instruction cache.
__device__ __noinline__ float ml_consume_batch(const float4* buf, float sm, const float4* lut) {
  float4 nx = buf[0];
#pragma unroll 1
  for (int q4 = 0; q4 < 3 * RNA_ML_BATCH / 4; q4++) {
    const float4 x = nx;
    if (q4 + 1 < 3 * RNA_ML_BATCH / 4) nx = buf[(q4 + 1) * 32];
    sm = RNA_COOP_LSE_NN(sm, x.x, lut);
    sm = RNA_COOP_LSE_NN(sm, x.y, lut);
    sm = RNA_COOP_LSE_NN(sm, x.z, lut);
    sm = RNA_COOP_LSE_NN(sm, x.w, lut);
  }
  return sm;
}

// Called through a pointer read from device memory: an indirect call obeys the full ABI, so the callee is compiled with
// every predicate and register at its disposal instead of the few the big kernel leaves over.
typedef float (*MlConsumeFn)(const float4*, float, const float4*);
__device__ MlConsumeFn g_ml_consume = ml_consume_batch;

template <bool CONTRA, class SV>
__device__ __forceinline__ void outside_X_diag_split(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut,
                                                     const ModelParams& P, float Z, int d, int lane, int nl, bool producer,
                                                     float* ring, unsigned& gbatch, long long* tsplit) {
  constexpr int B = RNA_ML_BATCH;
  const MlConsumeFn consume = *reinterpret_cast<MlConsumeFn volatile*>(&g_ml_consume);
  const int L = v.L;
  if (d < v.dout0) return;
  const float NEG = RNA_NEG_INF;
  const int st = (L - 1 - d) >> 1;
  const StepCells sc = step_cells<false>(v, st);
  const int tot = sc.cA + sc.cB, x0 = (d == sc.dA) ? 0 : sc.cA, cnt = v.pcnt[d], od = doff(d, L);
  const int lane32 = lane & 31;
  const typename Model2<CONTRA>::Dev* dev = T.g;
  for (int r0 = lane - lane32; r0 < cnt; r0 += nl) {   // warp-uniform: both warps of a pair see the same groups
    const int r = r0 + lane32;
    bool act = r < cnt;
    int i = 0, j = 0;
    float Cij = NEG;
    if (act) { i = v.plist[od + r]; j = i + d; Cij = v.C[od + i]; act = Cij > NEG; }
    const int ie = act ? i : 0;                                      // split points of this lane
    const int nb = (__reduce_max_sync(0xffffffffu, ie) + B - 1) / B;   // batches of this group (warp-uniform)
    if (producer) {
      float sa = NEG, unp = 0.f;
      if (act) {
        const float Aij = __fadd_rn(Cij, v2_acc<CONTRA>(T, v.s, L, i, j));
        if constexpr (CONTRA) sa = __fadd_rn(Aij, dev->mb_bp); else sa = __fadd_rn(Aij, dev->coeff_num_branches);
      }
      if constexpr (CONTRA) unp = dev->mb_unpair;
      const float* pX = v.X + j;                // probs_multibranch2[k][j], row-major, row k = 0
      const float* pR = v.R + j;                // probs_multibranch[k][j]
      const float* pM = v.M1rm + (L + i - 2);   // sums_1ormore[k+1][i-1], row 1
      int sQ = L - 1, sM = L - 2;
      for (int b = 0; b < nb; b++) {
        float p2[B], y[B], x1[B];
#pragma unroll
        for (int u = 0; u < B; u++) {           // all loads of the batch in flight before the first use
          const int kk = b * B + u;
          p2[u] = NEG; y[u] = NEG; x1[u] = NEG;
#ifndef RNA_ML_FAKE_PRODUCER
          if (kk < ie) { p2[u] = *pX; y[u] = *pR; if (kk < ie - 1) x1[u] = *pM; }
#endif
          pX += sQ; pR += sQ; sQ--;
          pM += sM; sM--;
        }
        // operands q = 3 u + {0,1,2} in fold order, four to a 16-byte shared-memory word per lane
        float4* buf = reinterpret_cast<float4*>(ring) + ((gbatch + (unsigned)b) & 1u) * (B * 3 / 4 * 32) + lane32;
        float o[3 * B];
#pragma unroll
        for (int u = 0; u < B; u++)
          ml_operands<CONTRA>(sa, unp, ie - 1 - (b * B + u), p2[u], y[u], x1[u], o[3 * u], o[3 * u + 1], o[3 * u + 2]);
#pragma unroll
        for (int q4 = 0; q4 < 3 * B / 4; q4++) buf[q4 * 32] = make_float4(o[4 * q4], o[4 * q4 + 1], o[4 * q4 + 2], o[4 * q4 + 3]);
        asm volatile("bar.sync 1, 64;" ::: "memory");
      }
    } else {
      const long long c0 = tsplit ? clock64() : 0;
      float sm = NEG;
      if (act) sm = outside_cell_partial<CONTRA>(v, T, lut, P, Z, st, tot, x0 + r, i, j, Cij);
      if (tsplit) tsplit[0] += clock64() - c0;
      for (int b = 0; b < nb; b++) {
        asm volatile("bar.sync 1, 64;" ::: "memory");
        // LDS.128: ptxas sinks each shared-memory load in front of its first use (in-order issue puts that latency on
        // the chain), so the operands come four at a time
        const float4* buf = reinterpret_cast<const float4*>(ring) + ((gbatch + (unsigned)b) & 1u) * (B * 3 / 4 * 32) + lane32;
        // NOT unrolled: the body (four folds, ~170 instructions) must stay inside the ~6 KB L0 instruction cache — a
        // lone warp that streams a 16 KB straight-line body pays an instruction-fetch miss per cache line, measured
        // 142 cycles per fold against 95 for a compact loop
        sm = consume(buf, sm, lut);
      }
      if (act) v.Pm[od + i] = sm;
    }
    gbatch += (unsigned)nb;
  }
}

template <bool CONTRA>
__global__ void __launch_bounds__(256, 1) fold_kernel2_coop(const FoldArgs a) {
  typedef typename Model2<CONTRA>::Dev Dev;
  typedef typename Model2<CONTRA>::Small Small;
  typedef typename Model2<CONTRA>::View View;
  typedef CoopView SV;
  unsigned* const bar_ctr = reinterpret_cast<unsigned*>(a.work_counter);
  unsigned bar_target = 0;
  auto grid_sync = [&]() { coop_barrier(bar_ctr, bar_target); };

  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* lut = reinterpret_cast<float4*>(smem_raw);
  Small* small = reinterpret_cast<Small*>(smem_raw + 128);
  uint8_t* sseq = smem_raw + 128 + align16(sizeof(Small));
  float* ml_ring = reinterpret_cast<float*>(smem_raw + fold2_fixed_bytes<CONTRA>(a.Lcap));   // RNA_ML_RING_FLOATS
  unsigned ml_batches = 0;

  const Dev* dev = reinterpret_cast<const Dev*>(a.tables);
  const int tid = threadIdx.x;
  load_lse_lut(lut);
  {
    const Small* gs = dev_small<CONTRA>(dev);
    for (int x = tid; x < (int)(sizeof(Small) / 4); x += blockDim.x)
      reinterpret_cast<float*>(small)[x] = reinterpret_cast<const float*>(gs)[x];
  }
  View T;
  T.g = dev;
  T.sm = small;
  ModelParams P;
  P.MINSPAN = dev->min_span;
  if constexpr (CONTRA) P.MAX2 = dev->max_loop_len; else P.MAX2 = dev->max_2loop_len;
  P.allows_short = a.allows_short;
  const float NEG = RNA_NEG_INF;
  const int gtid = (int)(blockIdx.x * blockDim.x + threadIdx.x), gnt = (int)(gridDim.x * blockDim.x);
  // global warp id with consecutive ids on different SMs; role lanes are counted from the role's first warp
  const int gw = (tid >> 5) * (int)gridDim.x + (int)blockIdx.x;
  const int nXl = a.nXw * 32;
  const int lnX = gw * 32 + (tid & 31), lnY = (gw - a.nXw) * 32 + (tid & 31);
  const bool isX = gw < a.nXw, isY = !isX && gw < a.nXw + a.nYw, isZ = !isX && !isY && gw < a.nXw + a.nYw + a.nZw;
  __syncthreads();

  for (uint32_t w = 0; w < a.n_launch; w++) {
    const uint32_t sidx = a.order ? a.order[w] : w;
    const uint32_t sbeg = a.offsets[sidx];
    const int L = (int)(a.offsets[sidx + 1] - sbeg);
    const size_t TRI = (size_t)L * (L + 1) / 2;
    size_t ngcap = 0;
    for (int c = 1; c <= L; c += 32) ngcap += (size_t)min(32, L - c + 1) * ((c + 31) / 32);   // = fold2_ngcap(L)

    SV v;
    v.L = L;
    v.W2 = (L + 31) / 32 + 2;
    uint8_t* s = sseq + 4;
    v.s = s;
    float* f = a.workspace;
    v.C = f; f += TRI;
    v.R = f; f += TRI;
    v.X = f; f += TRI;
    v.E = f; f += TRI;
    v.M1 = f; f += TRI;
    v.Pm = v.E;
    v.Mroll = f; f += 3 * L;
    v.E0 = f; f += L;
    v.EL = f; f += L;
    int* tstack = reinterpret_cast<int*>(f);
    int* fill_ctr = tstack + 2 * (L + 2);
    v.mask = reinterpret_cast<uint32_t*>(fill_ctr + 2);
    v.gcumI = v.mask + (size_t)L * v.W2;
    v.gcumO = v.gcumI + L / 2 + 3;
    v.ccumI = v.gcumO + L / 2 + 3;
    v.ccumO = v.ccumI + L / 2 + 3;
    v.gbin = v.ccumO + L / 2 + 3;
    v.gbout = v.gbin + ngcap + 2;
    v.gstepI = reinterpret_cast<uint16_t*>(v.gbout + ngcap + 2);
    v.gstepO = v.gstepI + ngcap + 2;
    v.pcnt = v.gstepO + ngcap + 2;
    v.din0 = CONTRA ? 0 : (P.MINSPAN - 1);
    v.dout0 = CONTRA ? (a.allows_short ? 1 : P.MINSPAN - 1) : (P.MINSPAN - 1);
    v.plist = v.pcnt + ((L + 1) & ~1);
    v.RR = reinterpret_cast<uint8_t*>(v.plist + TRI);
    v.LL = v.RR + L;
    v.tin = nullptr; v.tout = nullptr; v.ccnt = nullptr; v.tcap = a.tcap;
    v.M1rm = a.workspace + fold2_seq_bytes(L, 2, 5) / 4 + 16;   // two more triangular matrices behind the region
    v.MB = v.M1rm + TRI;
    float* sums_out = a.sums ? a.sums + a.sums_offsets[sidx] : nullptr;
    v.Mfull = sums_out ? sums_out + (size_t)5 * TRI : nullptr;

    long long tk = (a.dbg && gtid == 0) ? clock64() : 0;
    auto mark = [&](int slot) { if (a.dbg && gtid == 0) { const long long now = clock64(); a.dbg[slot] = now - tk; tk = now; } };
    __syncthreads();
    for (int x = tid; x < L; x += blockDim.x) s[x] = a.bases[sbeg + x];   // every CTA keeps its own copy of the bases
    if (tid < 4) { sseq[tid] = 0; s[L + tid] = 0; }
    __syncthreads();
    for (int x = gtid; x < L * v.W2; x += gnt) setup_mask_word<CONTRA>(v, P, x);
    for (int x = gtid; x < L; x += gnt) setup_codes(v, x);
    grid_sync();
    for (int d = gtid; d < L; d += gnt) setup_list_diag(v, d);
    grid_sync();
    if (a.stream_ws) {
      if (gtid == 0) { setup_groups(v); *fill_ctr = 0; }
      v.ccnt = reinterpret_cast<uint16_t*>(v.C);
      grid_sync();
      stream_count(v, P, gtid, gnt);
      grid_sync();
      stream_groupmax(v, gtid, gnt);
      grid_sync();
      if (gtid == 0) stream_scan(v);
      grid_sync();
      const uint32_t NGI = v.gcumI[max(num_steps_inside(v), 0)], NGO = v.gcumO[max(num_steps_outside(v), 0)];
      if (v.gbin[NGI] <= a.tcap && v.gbout[NGO] <= a.tcap) {
        v.tin = reinterpret_cast<uint2*>(a.stream_ws);
        v.tout = v.tin + a.tcap;
        const uint32_t ntask = stream_num_tasks(v);
        for (;;) {
          uint32_t tau = 0;
          if ((tid & 31) == 0) tau = (uint32_t)atomicAdd(fill_ctr, 1);
          tau = __shfl_sync(0xffffffffu, tau, 0);
          if (tau >= ntask) break;
          stream_fill_task<CONTRA>(v, T, P, tau, tid & 31);
          __syncwarp();
        }
      }
      grid_sync();
    }
    if constexpr (CONTRA) score_table_acc(v, T, gtid, gnt);   // accessible scores for the dense Y chains (in v.MB)
    for (size_t x = gtid; x < TRI; x += gnt) { v.C[x] = NEG; v.R[x] = NEG; v.X[x] = NEG; v.E[x] = 0.f; v.M1[x] = NEG; }
    for (int x = gtid; x < 3 * L; x += gnt) v.Mroll[x] = NEG;
    grid_sync();
    mark(0);   // setup + term streams

    // ---- inside, pair steps (fold_phases.cuh "PAIR-STEP schedule"): phase A = X two-loop parts of (t, t+1) |
    //      Y partial rightmost-pair sums of (t, t+1) | Z one dense chain per lane for (t-2, t-1); phase B = finish ----
    const int d_in0 = v.din0;
    const int wZ = gw - a.nXw - a.nYw;
    long long cyc[4] = {0, 0, 0, 0};
    const bool timed = a.dbg && (tid & 31) == 0;
    for (int t = d_in0; t <= L + 1; t += 2) {
      const long long c0 = timed ? clock64() : 0;
      if (isX) inside_X<CONTRA>(v, T, lut, P, (t - d_in0) >> 1, lnX, nXl);
      else if (isY) {
        if constexpr (CONTRA) inside_Y_dense_pair<4>(v, T, lut, t, gw - a.nXw, a.nYw, tid & 31);
      } else if (isZ) inside_chain_pair<CONTRA, 8>(v, T, lut, t, wZ, a.nZw, tid & 31);
      const long long c1 = timed ? clock64() : 0;
      unsigned long long* gslot = reinterpret_cast<unsigned long long*>(a.dbg) + 64 + (size_t)((t - d_in0) >> 1) * 8;
      if (timed) atomicMax(gslot + 5, global_ns());   // last arrival at the phase-A barrier
      grid_sync();
      const long long c2 = timed ? clock64() : 0;
      if (timed) atomicMax(gslot + 6, global_ns());   // last release
      inside_fin_pair<CONTRA>(v, T, lut, t, gtid, gnt);
      grid_sync();
      if (timed) atomicMax(gslot + 7, global_ns());   // end of the step
      if (timed) {
        cyc[0] += c1 - c0; cyc[1] += c2 - c1; cyc[2] += clock64() - c2;
        const int r = isX ? 0 : isY ? 1 : isZ ? 2 + wZ % 3 : -1;   // slowest warp of a role (chain kind) per step
        if (r >= 0) atomicMax(reinterpret_cast<unsigned long long*>(a.dbg) + 64 + (size_t)((t - d_in0) >> 1) * 8 + r, (unsigned long long)(c1 - c0));
      }
    }
    if (timed) {   // work / wait cycles of the first warp of each role, and of phase B (with its barrier)
      const int slot = (isX && lnX == 0) ? 8 : (isY && lnY == 0) ? 12 : (isZ && wZ == 0) ? 16 : -1;
      if (slot >= 0) { a.dbg[slot] = cyc[0]; a.dbg[slot + 1] = cyc[1]; a.dbg[slot + 2] = cyc[2]; }
    }
    for (int x = gtid; x < L; x += gnt) {
      v.E0[x] = v.E[doff(x, L)];
      v.EL[x] = v.E[doff(L - 1 - x, L) + x];
    }
    if (sums_out) export_fold_sums<CONTRA>(v, T, P, sums_out, gtid, gnt);   // FoldSums / FoldScores (N3)
    grid_sync();
    mark(1);   // inside
    const float Z = v.E0[L - 1];
    if (gtid == 0 && a.out_logz) a.out_logz[sidx] = Z;
    if (a.inside_only) continue;
    outside_prep<CONTRA>(v, T, gtid, gnt);   // row-major sums_1ormore, table of multibranch closing scores
    for (size_t x = gtid; x < TRI; x += gnt) { v.Pm[x] = NEG; v.R[x] = NEG; v.X[x] = NEG; }
    grid_sync();
    // ---- outside, one diagonal per step: X(d) | Y(d) --------------------------------------------------------------
    const int d_out0 = v.dout0;
    cyc[0] = cyc[1] = 0;
    // the producer/consumer split needs every X warp to be warp 0 of its CTA and 8 warps per CTA
    const bool ml_split = a.nXw <= (int)gridDim.x && blockDim.x == 256 && !a.no_ml_split;
    for (int d = L - 1; d >= d_out0; d--) {
      const long long c0 = timed ? clock64() : 0;
      long long tsp[2] = {0, 0};
      if (ml_split) {
        // consumers: warp 0 of the first nXw CTAs (the X lanes as before); producers: warp 1 of the same CTAs;
        // probs_multibranch(2): every warp from index 2 up
        const int wi = tid >> 5;
        if (isX) outside_X_diag_split<CONTRA>(v, T, lut, P, Z, d, lnX, nXl, false, ml_ring, ml_batches, timed ? tsp : nullptr);
        else if (wi == 1 && (int)blockIdx.x < a.nXw) outside_X_diag_split<CONTRA>(v, T, lut, P, Z, d, (int)blockIdx.x * 32 + (tid & 31), nXl, true, ml_ring, ml_batches, nullptr);
        else if (wi >= 2) outside_Y_dense<CONTRA, 4>(v, T, lut, d, (wi - 2) * (int)gridDim.x + (int)blockIdx.x, 6 * (int)gridDim.x, tid & 31);
      } else if (isX) outside_X_diag_rm<CONTRA, 4>(v, T, lut, P, Z, d, lnX, nXl, timed ? tsp : nullptr);
      else if (isY || isZ) outside_Y_dense<CONTRA, 4>(v, T, lut, d, gw - a.nXw, a.nYw + a.nZw, tid & 31);
      const long long c1 = timed ? clock64() : 0;
      grid_sync();
      if (timed) {
        cyc[0] += c1 - c0; cyc[1] += clock64() - c1;
        if (gw < a.nXw + a.nYw + a.nZw) {
          // X: (total << 0) in slot 0; the stream part of the slowest warp rides along in slot 2 (packed with the total)
          unsigned long long* sl = reinterpret_cast<unsigned long long*>(a.dbg) + (1 << 19) + (size_t)d * 4;
          atomicMax(sl + (isX ? 0 : 1), (unsigned long long)(c1 - c0));
          if (isX) atomicMax(sl + 2, ((unsigned long long)(c1 - c0) << 24) | (unsigned long long)min((long long)0xffffff, tsp[0] >> 4));
        }
      }
    }
    if (timed) {
      const int slot = (isX && lnX == 0) ? 20 : (!isX && lnY == 0) ? 24 : -1;
      if (slot >= 0) { a.dbg[slot] = cyc[0]; a.dbg[slot + 1] = cyc[1]; }
    }
    mark(2);   // outside
    for (size_t x = gtid; x < TRI; x += gnt) {
      const float val = v.Pm[x];
      v.Pm[x] = (val > NEG) ? approx_expf(val) : -1.0f;
    }
    grid_sync();
    if (a.out_bpp) {
      float* ob = a.out_bpp + a.bpp_offsets[sidx];
      for (int i = (int)blockIdx.x; i < L - 1; i += (int)gridDim.x) {
        const size_t rowoff = (size_t)i * (size_t)(2 * L - i - 1) / 2;
        for (int x = tid; x < L - 1 - i; x += blockDim.x) ob[rowoff + x] = v.Pm[doff(x + 1, L) + i];
      }
    }
    {
      const float* Pm = v.Pm;
      auto getp = [=](int d, int i) -> float { return __ldcg(&Pm[doff(d, L) + i]); };
      // every threshold's W matrix behind the two extra matrices of the outside pass, then the traceback stacks
      float* Wall = v.MB + TRI;
      int* tstacks = reinterpret_cast<int*>(Wall + (size_t)a.n_gammas * TRI);
      centroid_coop_all(a, sidx, sbeg, L, Wall, TRI, tstacks, getp, grid_sync);
    }
    mark(3);   // BPP + centroid
  }
}

}  // namespace rna
