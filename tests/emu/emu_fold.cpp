// emu_fold.cpp — HOST emulator of the v2 fold kernel (rna_algos_b200/csrc/fold_kernel2.cuh).
// TEST INFRASTRUCTURE ONLY: it compiles the product's phase functions (fold_phases.cuh) with g++ and runs
// them role by role, barrier by barrier, exactly as the CUDA kernel schedules them — with a configurable
// lane / role order — so that the parallel decomposition (who computes what, in which step, from which
// inputs) can be checked bit-for-bit against the oracle on a machine without a GPU.  It is never loaded by
// the product package.
//   g++ -O1 -std=c++17 -ffp-contract=off -fno-fast-math -shared -fPIC -o _emu.so emu_fold.cpp
#include <stdlib.h>

#include <vector>

#include "../../rna_algos_b200/csrc/fold_phases.cuh"
#include "../../rna_algos_b200/csrc/table_pack.h"

using namespace rna;

namespace {
template <bool CONTRA>
int run(const uint8_t* seq, int L, int allows_short, const typename Model2<CONTRA>::Dev* dev, int nX, int nY, int nZ,
        int order, int tcap, float* out_bpp, float* out_logz) {
  typedef typename Model2<CONTRA>::View View;
  View T;
  T.g = dev;
  T.sm = dev_small<CONTRA>(dev);
  const float4* lut = kLnExp1pCoef;
  ModelParams P;
  P.MINSPAN = dev->min_span;
  if constexpr (CONTRA) P.MAX2 = dev->max_loop_len; else P.MAX2 = dev->max_2loop_len;
  P.allows_short = allows_short;
  const float NEG = RNA_NEG_INF;
  const int TRI = L * (L + 1) / 2;
  std::vector<uint8_t> sbuf(L + 8, 0);
  for (int x = 0; x < L; x++) sbuf[4 + x] = seq[x];
  SeqViewT<uint16_t> v;
  v.L = L;
  v.W2 = (L + 31) / 32 + 2;
  v.s = sbuf.data() + 4;
  std::vector<uint32_t> mask((size_t)L * v.W2);
  std::vector<uint16_t> plist(TRI), pcnt(L);
  std::vector<uint8_t> RR(L), LL(L);
  std::vector<float> C(TRI, NEG), R(TRI, NEG), X(TRI, NEG), E(TRI, 0.f), M1(TRI, NEG), Mroll(3 * L, NEG), E0(L), EL(L);
  v.mask = mask.data(); v.plist = plist.data(); v.pcnt = pcnt.data(); v.RR = RR.data(); v.LL = LL.data();
  v.C = C.data(); v.R = R.data(); v.X = X.data(); v.E = E.data(); v.M1 = M1.data(); v.Mroll = Mroll.data();
  v.E0 = E0.data(); v.EL = EL.data();
  std::vector<float> M1rm(TRI, NEG), MB(TRI, 0.f);
  v.M1rm = M1rm.data(); v.MB = MB.data(); v.Mfull = nullptr;
  // log P either aliases sums_external (all-in-one-space modes) or is a matrix of its own (the shared-memory mode
  // keeps C and log P on chip and the dense matrices in its HBM/L2 slot)
  std::vector<float> Pown(TRI, NEG);
  const bool separateP = (order == 1);
  v.Pm = separateP ? Pown.data() : E.data();
  // setup (two barriers in the kernel: masks, then lists)
  for (int x = 0; x < L * v.W2; x++) setup_mask_word<CONTRA>(v, P, x);
  for (int p = 0; p < L; p++) setup_codes(v, p);
  for (int d = 0; d < L; d++) setup_list_diag(v, d);
  // term streams (tcap: -1 unbounded, 0 = score on the fly (the kernel's fallback), n = streams if they fit)
  const int nt = nX + nY + nZ, nwarps = (nt + 31) / 32;
  const int d_in0 = CONTRA ? 0 : (P.MINSPAN - 1);
  const int d_out0 = CONTRA ? (allows_short ? 1 : P.MINSPAN - 1) : (P.MINSPAN - 1);
  v.din0 = d_in0; v.dout0 = d_out0;
  const int nsmax = L / 2 + 2;
  size_t ngcap = 0;
  for (int c = 1; c <= L; c += 1) ngcap += (size_t)(c + 31) / 32;
  std::vector<uint32_t> gcumI(nsmax + 1, 0), gcumO(nsmax + 1, 0), ccumI(nsmax + 1, 0), ccumO(nsmax + 1, 0), gbin(ngcap + 2, 0), gbout(ngcap + 2, 0);
  std::vector<uint16_t> gstepI(ngcap + 1, 0), gstepO(ngcap + 1, 0);
  v.gcumI = gcumI.data(); v.gcumO = gcumO.data(); v.ccumI = ccumI.data(); v.ccumO = ccumO.data(); v.gstepI = gstepI.data(); v.gstepO = gstepO.data();
  v.gbin = gbin.data(); v.gbout = gbout.data();
  setup_groups(v);
  const uint32_t NGI = gcumI[std::max(num_steps_inside(v), 0)], NGO = gcumO[std::max(num_steps_outside(v), 0)];
  if (NGI > ngcap || NGO > ngcap) return 3;
  std::vector<uint16_t> ccnt(2 * (size_t)TRI + 2, 0);
  std::vector<uint2> tin, tout;
  v.tin = nullptr; v.tout = nullptr; v.ccnt = ccnt.data(); v.tcap = 0;
  if (tcap != 0) {
    for (int l = 0; l < nt; l++) stream_count(v, P, l, nt);
    for (int l = nt - 1; l >= 0; l--) stream_groupmax(v, l, nt);
    stream_scan(v);
    if (tcap < 0 || (gbin[NGI] <= (uint32_t)tcap && gbout[NGO] <= (uint32_t)tcap)) {
      tin.assign(gbin[NGI] + 1, make_uint2(0xdeadbeefu, 0x7fffffffu));    // poison: every element must be written
      tout.assign(gbout[NGO] + 1, make_uint2(0xdeadbeefu, 0x7fffffffu));
      v.tin = tin.data(); v.tout = tout.data();
      const uint32_t ntask = stream_num_tasks(v);
      for (uint32_t tau = 0; tau < ntask; tau++) {     // (the kernel hands tasks to warps dynamically)
        const uint32_t tk = (order == 0) ? tau : ntask - 1 - tau;
        for (int ln = 31; ln >= 0; ln--) stream_fill_task<CONTRA>(v, T, P, tk, ln);
      }
      (void)nwarps;
    }
  }
  auto lanes = [&](int n, auto&& fn) {
    if (order == 0) for (int l = 0; l < n; l++) fn(l);
    else for (int l = n - 1; l >= 0; l--) fn(l);
  };
  auto validZ = [&](int d) { return d >= d_in0 && d < L; };
  const bool pairsplit = (order == 3);   // the cooperative kernel's inside pass: pair steps, one chain per lane
  const bool single = (order == 2) || pairsplit;   // one diagonal per step: the cooperative kernel's outside pass
  CoopView cv;
  static_cast<SeqViewT<uint16_t>&>(cv) = v;
  if (pairsplit) {
    const int nZw = (nZ + 31) / 32;
    if constexpr (CONTRA) for (int l = 0; l < nt; l++) score_table_acc(cv, T, l, nt);
    for (int st = 0; d_in0 + 2 * st <= L + 1; st++) {
      const int t = d_in0 + 2 * st;
      // phase A (roles in an arbitrary order; none reads what another writes in this phase)
      for (int w = nZw - 1; w >= 0; w--)
        for (int ln = 0; ln < 32; ln++) inside_chain_pair<CONTRA, 3>(cv, T, lut, t, w, nZw, ln);
      if constexpr (CONTRA) {
        if (nX & 1) {   // (the sparse partial sums, still a valid Y of this schedule)
          if (t + 1 < L) for (int l = 0; l < nY; l++) inside_Y_contra<3>(cv, T, lut, t + 1, l, nY, 1);
          if (t < L) for (int l = nY - 1; l >= 0; l--) inside_Y_contra<1>(cv, T, lut, t, l, nY, 0);
        } else {
          const int nYw = (nY + 31) / 32;
          for (int w = 0; w < nYw; w++)
            for (int ln = 31; ln >= 0; ln--) inside_Y_dense_pair<3>(cv, T, lut, t, w, nYw, ln);
        }
      }
      for (int l = nX - 1; l >= 0; l--) inside_X<CONTRA>(cv, T, lut, P, st, l, nX);
      // phase B
      for (int l = nt - 1; l >= 0; l--) inside_fin_pair<CONTRA>(cv, T, lut, t, l, nt);
    }
  } else if (single) {
    // step t: X(t) whole fold | Y(t) | Z(t-1)
    for (int t = d_in0; t <= L; t++) {
      if (validZ(t - 1)) for (int l = nZ - 1; l >= 0; l--) inside_Z<CONTRA, 3>(v, T, lut, t - 1, l, nZ);
      if constexpr (CONTRA) { if (t < L) for (int l = 0; l < nY; l++) inside_Y_contra<3>(v, T, lut, t, l, nY); }
      if (t < L) for (int l = nX - 1; l >= 0; l--) inside_X_diag<CONTRA>(v, T, lut, P, t, l, nX);
    }
  } else {
  // inside step st (diagonals t = d_in0 + 2 st, t+1): phase 1 = X two-loop parts | [Z(t-2), Y(t-1)] bar [Z(t-1), Y(t)];
  // phase 2 = X closing multibranch terms
  for (int st = 0; d_in0 + 2 * st <= L + 1; st++) {
    const int t = d_in0 + 2 * st;
    auto rX = [&] { lanes(nX, [&](int l) { inside_X<CONTRA>(v, T, lut, P, st, l, nX); }); };
    auto rZ = [&](int d) { if (validZ(d)) lanes(nZ, [&](int l) { inside_Z<CONTRA, 2>(v, T, lut, d, l, nZ); }); };
    auto rY = [&](int d) { if constexpr (CONTRA) { if (validZ(d)) lanes(nY, [&](int l) { inside_Y_contra<1>(v, T, lut, d, l, nY); }); } };
    auto rYZ = [&] {
      if (order == 0) { rZ(t - 2); rY(t - 1); } else { rY(t - 1); rZ(t - 2); }
      if (order == 0) { rY(t); rZ(t - 1); } else { rZ(t - 1); rY(t); }
    };
    if (order == 0) { rX(); rYZ(); } else { rYZ(); rX(); }
    lanes(nX, [&](int l) { inside_X_fin<CONTRA>(v, T, lut, st, l, nX); });
  }
  }
  for (int x = 0; x < L; x++) { E0[x] = E[doff(x, L)]; EL[x] = E[doff(L - 1 - x, L) + x]; }
  const float Z = E0[L - 1];
  for (int x = 0; x < TRI; x++) { v.Pm[x] = NEG; R[x] = NEG; X[x] = NEG; }
  if (out_logz) *out_logz = Z;
  if (pairsplit) {
    // the cooperative kernel's outside pass: row-major probs_multibranch(2), dense Y chains on two lanes per cell
    for (int l = nt - 1; l >= 0; l--) outside_prep<CONTRA>(cv, T, l, nt);
    const int nYw = (nY + nZ + 31) / 32;
    for (int d = L - 1; d >= d_out0; d--) {
      for (int w = 0; w < nYw; w++)
        for (int ln = 31; ln >= 0; ln--) outside_Y_dense<CONTRA, 3>(cv, T, lut, d, w, nYw, ln);
      for (int l = 0; l < nX; l++) outside_X_diag_rm<CONTRA, 3>(cv, T, lut, P, Z, d, l, nX);
    }
  } else if (single) {
    for (int d = L - 1; d >= d_out0; d--) {   // step d: X(d) whole fold | Y(d)
      const int nl = nY + nZ;
      for (int l = nX - 1; l >= 0; l--) outside_X_diag<CONTRA, 4>(v, T, lut, P, Z, d, l, nX);
      for (int l = 0; l < nl; l++) outside_Y<CONTRA, 2>(v, T, lut, d, l, nl);
    }
  }
  for (int st = 0; !single && L - 1 - 2 * st >= d_out0; st++) {
    const int d = L - 1 - 2 * st;
    auto rX = [&] { lanes(nX, [&](int l) { outside_X<CONTRA>(v, T, lut, P, Z, st, l, nX); }); };
    auto rY = [&] {
      const int nl = nY + nZ;
      if (d + 1 < L) lanes(nl, [&](int l) { outside_Y<CONTRA, 1>(v, T, lut, d + 1, l, nl); });
      lanes(nl, [&](int l) { outside_Y<CONTRA, 2>(v, T, lut, d, l, nl); });
    };
    if (order == 0) { rX(); rY(); } else { rY(); rX(); }
    lanes(nX, [&](int l) { outside_X_ml<CONTRA, 2>(v, T, lut, st, l, nX); });
  }
  if (out_bpp) {
    for (int i = 0; i < L - 1; i++) {
      const size_t rowoff = (size_t)i * (size_t)(2 * L - i - 1) / 2;
      for (int x = 0; x < L - 1 - i; x++) {
        const float val = v.Pm[doff(x + 1, L) + i];
        out_bpp[rowoff + x] = (val > NEG) ? approx_expf(val) : -1.0f;
      }
    }
  }
  return 0;
}
}  // namespace

extern "C" int emu_fold(const uint8_t* seq, int L, int contra, int allows_short, const RnaTurnerTables* tt,
                        const RnaContraTables* ct, int nX, int nY, int nZ, int order, int tcap, float* out_bpp,
                        float* out_logz) {
  std::string err;
  if (contra) {
    static DevContra d;   // big struct: keep off the stack
    if (pack_contra(ct, &d, &err)) return 1;
    return run<true>(seq, L, allows_short, &d, nX, nY, nZ, order, tcap, out_bpp, out_logz);
  }
  static DevTurner d;
  std::vector<float> hp;
  if (pack_turner(tt, &d, &hp, &err)) return 1;
  d.hairpin_init_ext = hp.data();
  d.int11 = &tt->interior_scores_1x1[0][0][0][0][0][0];
  d.int12 = &tt->interior_scores_1x2[0][0][0][0][0][0][0];
  d.int22 = &tt->interior_scores_2x2[0][0][0][0][0][0][0][0];
  return run<false>(seq, L, allows_short, &d, nX, nY, nZ, order, tcap, out_bpp, out_logz);
}
