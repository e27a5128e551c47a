// spl_advect_tma.cuh -- tile-staged semi-Lagrangian advection (fp32, nx % 4 = 0, sm_100a)
//
// Convection.apply + Grid.update_velocities (Convection.fs:21-66, Simulator.fs:82) with the
// intended semantics of spl_advect.cuh (per-face RK2 backtrace + trilinear sample), and the same
// trace from the cell centre for cell-centred scalar fields (the reference's Cell carries no
// scalar, Cell.fs:11-16: north_star's "velocity and scalar fields", SURVEY 8d "+8 B").
//
// One CTA owns a 32 x 2NW tile of the x-y plane and marches along z.  The whole sampling
// footprint of the tile -- offsets are limited to ADV_MAXOFF < 3 cells, so corners span
// [-3, +3] in every axis -- lives in shared memory: a ring of 8 z-planes (k-3 .. k+4) of the
// raw (32 + 8) x (2NW + 6) tile of u, v, w (and the scalar), filled by a producer warp with one
// cp.async.bulk.tensor per field and plane.  Cells outside the box and planes outside the
// allocation are zero-filled by the TMA unit, which *is* "an absent cell contributes 0"
// (Convection.fs:36-38): box-shell cells need no special case and no second pass.
//
// Instruction diet (the L1-gather kernel of spl_advect.cuh is issue-bound at 744 instructions
// per cell):
//   * a thread owns TWO cells, rows ty and ty + NW of the tile, and every floating-point
//     operation of the pair is one packed instruction (add / sub / mul / fma .f32x2: FADD2,
//     FMUL2, FFMA2) -- the 8 corner loads of the two cells land in the halves of 64-bit
//     register pairs, so packing costs no moves;
//   * the 96 + 18 gathers of a cell are LDS at immediate offsets from one 32-bit address per
//     sample (conflict-free across a warp wherever the floor of the offsets is uniform);
//   * floor / fraction without the conversion pipe: t = o + 1.5 * 2^23 rounded DOWN leaves
//     floor(o) in the low mantissa bits, usable as an integer in the address arithmetic with
//     the exponent bits folded into a per-thread constant;
//   * slot 8 of the ring duplicates slot 0, so plane k+1 of a corner pair is always one plane
//     stride after plane k (no second wrap).
// Only cells whose offsets exceed ADV_MAXOFF go to the worklist of k_advect_rest.
#pragma once
#include <cuda.h>

#include "spl_advect.cuh"
#include "spl_common.cuh"
#include "spl_pcg_tma.cuh"

namespace spl {

constexpr int ADVT_ROW = 40;            // floats per raw row: cells i0-4 .. i0+35
constexpr int ADVT_SLOTS = 8;           // ring planes; slot ADVT_SLOTS mirrors slot 0
constexpr float ADVT_MAGIC = 12582912.0f;           // 1.5 * 2^23
constexpr unsigned ADVT_MAGIC_I = 0x4B400000u;      // its bit pattern

template <int NW>
struct AdvT {
  static constexpr int TY = 2 * NW;                                 // rows per tile
  static constexpr int ROWS = TY + 6;                               // j0-3 .. j0+TY+2
  static constexpr int PL = ((ROWS * ADVT_ROW + 31) / 32) * 32;     // plane stride (128 B aligned)
  static constexpr unsigned BOX_BYTES = ROWS * ADVT_ROW * 4;        // bytes one TMA box delivers
  static constexpr int DB = NW * ADVT_ROW;                          // second cell of a thread
};

template <int NW, int NF>
struct alignas(128) AdvTSmem {
  float ring[NF][ADVT_SLOTS + 1][AdvT<NW>::PL];
  unsigned long long full[ADVT_SLOTS];   // TMA completion per slot
  unsigned long long empty[2];           // "every row warp is done with plane k" (k parity)
};

// ---- packed pairs: .x = first cell of the thread, .y = second ------------------------------
typedef unsigned long long pk2;
__device__ __forceinline__ pk2 pk(float a, float b) {
  pk2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float pk_lo(pk2 v) {
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
  return a;
}
__device__ __forceinline__ float pk_hi(pk2 v) {
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
  return b;
}
__device__ __forceinline__ pk2 pk_add(pk2 a, pk2 b) {
  pk2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ pk2 pk_add_rm(pk2 a, pk2 b) {   // round towards -infinity
  pk2 d;
  asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ pk2 pk_sub(pk2 a, pk2 b) {
  pk2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ pk2 pk_mul(pk2 a, pk2 b) {
  pk2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ pk2 pk_fma(pk2 a, pk2 b, pk2 c) {
  pk2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ pk2 pk_mix(pk2 a, pk2 b, pk2 w) {   // a + w (b - a)
  return pk_fma(w, pk_sub(b, a), a);
}

struct AdvAx2 {      // floor / fraction of one offset, both cells
  pk2 t;             // bit patterns ADVT_MAGIC_I + floor(o)
  pk2 w;             // o - floor(o)
};
__device__ __forceinline__ AdvAx2 adv_split2(pk2 o, pk2 magic) {
  AdvAx2 r;
  r.t = pk_add_rm(o, magic);
  r.w = pk_sub(o, pk_sub(r.t, magic));
  return r;
}

// trilinear samples of one staged field at (cell + offset) for the two cells of a thread.
// F = ring of the field, tbm = index of the first cell in the raw tile with the magic constants
// of the x / y terms folded in (second cell: + DB), kcm likewise for the plane -> slot term.
// The index is clamped (unsigned min) so that a wild offset -- NaN, Inf, or beyond ADV_MAXOFF:
// cells that are flagged and recomputed by k_advect_rest -- still reads inside the ring; there
// is no branch on the flags, so the samples of the three faces interleave freely.
template <int NW>
__device__ __forceinline__ pk2 adv_tri2(const float* __restrict__ F, unsigned tbm, unsigned kcm,
                                        const AdvAx2& X, const AdvAx2& Y, const AdvAx2& Z) {
  constexpr int PL = AdvT<NW>::PL;
  constexpr int ROW = ADVT_ROW;
  constexpr unsigned LIMIT = (unsigned)(ADVT_SLOTS * PL - ROW - 2);
  const unsigned xa = (unsigned)X.t, xb = (unsigned)(X.t >> 32);
  const unsigned ya = (unsigned)Y.t, yb = (unsigned)(Y.t >> 32);
  const unsigned za = (unsigned)Z.t, zb = (unsigned)(Z.t >> 32);
  const unsigned sa = (za + kcm) & (ADVT_SLOTS - 1);
  const unsigned sb = (zb + kcm) & (ADVT_SLOTS - 1);
  const float* __restrict__ pa = F + min(sa * PL + (ya * ROW + xa) + tbm, LIMIT);
  const float* __restrict__ pb = F + min(sb * PL + (yb * ROW + xb) + (tbm + AdvT<NW>::DB), LIMIT);
  const pk2 c000 = pk(pa[0], pb[0]), c100 = pk(pa[1], pb[1]);
  const pk2 c010 = pk(pa[ROW], pb[ROW]), c110 = pk(pa[ROW + 1], pb[ROW + 1]);
  const pk2 c001 = pk(pa[PL], pb[PL]), c101 = pk(pa[PL + 1], pb[PL + 1]);
  const pk2 c011 = pk(pa[PL + ROW], pb[PL + ROW]), c111 = pk(pa[PL + ROW + 1], pb[PL + ROW + 1]);
  const pk2 A = pk_mix(pk_mix(c000, c100, X.w), pk_mix(c010, c110, X.w), Y.w);
  const pk2 B = pk_mix(pk_mix(c001, c101, X.w), pk_mix(c011, c111, X.w), Y.w);
  return pk_mix(A, B, Z.w);
}

// some |offset| of a cell > lim ?  (bit 0: first cell, bit 1: second)
__device__ __forceinline__ unsigned adv_bad2(pk2 a, pk2 b, pk2 c, float lim) {
  // comparisons, not fmaxf: a NaN offset must count as bad
  const bool oka = (fabsf(pk_lo(a)) <= lim) && (fabsf(pk_lo(b)) <= lim) && (fabsf(pk_lo(c)) <= lim);
  const bool okb = (fabsf(pk_hi(a)) <= lim) && (fabsf(pk_hi(b)) <= lim) && (fabsf(pk_hi(c)) <= lim);
  return (oka ? 0u : 1u) | (okb ? 0u : 2u);
}

// RK2 trace from a face centre (FACE = 0, 1, 2: the +x, +y, +z face) or the cell centre
// (FACE = 3) and the sample of field RF at the departure point, for both cells of a thread.
// (mx, my, mz) = half-step offsets from stage 0.  Sampling component b from a point of type a
// shifts axis a by +1/2 (a face only) and axis b by -1/2 unless a = b (SURVEY A.2 intended).
template <int NW, int FACE>
__device__ __forceinline__ pk2 adv_trace2(const float* __restrict__ RU, const float* __restrict__ RV,
                                          const float* __restrict__ RW, const float* __restrict__ RF,
                                          unsigned tbm, unsigned kcm, pk2 mx, pk2 my, pk2 mz,
                                          pk2 cc, pk2 magic, unsigned& bad) {
  const pk2 ph = pk(0.5f, 0.5f), nh = pk(-0.5f, -0.5f);
  bad |= adv_bad2(mx, my, mz, ADV_MAXOFF - 0.5f);
  const AdvAx2 P0 = adv_split2(mx, magic), P1 = adv_split2(my, magic), P2 = adv_split2(mz, magic);
  const AdvAx2 Q0 = adv_split2(pk_add(mx, FACE == 0 ? ph : nh), magic);
  const AdvAx2 Q1 = adv_split2(pk_add(my, FACE == 1 ? ph : nh), magic);
  const AdvAx2 Q2 = adv_split2(pk_add(mz, FACE == 2 ? ph : nh), magic);
#define ADV_VAR(comp, axis) \
  (((FACE < 3 && (axis) == FACE && (comp) != FACE) || ((axis) == (comp) && (comp) != FACE)))
  const pk2 u1 = adv_tri2<NW>(RU, tbm, kcm, ADV_VAR(0, 0) ? Q0 : P0, ADV_VAR(0, 1) ? Q1 : P1,
                              ADV_VAR(0, 2) ? Q2 : P2);
  const pk2 v1 = adv_tri2<NW>(RV, tbm, kcm, ADV_VAR(1, 0) ? Q0 : P0, ADV_VAR(1, 1) ? Q1 : P1,
                              ADV_VAR(1, 2) ? Q2 : P2);
  const pk2 w1 = adv_tri2<NW>(RW, tbm, kcm, ADV_VAR(2, 0) ? Q0 : P0, ADV_VAR(2, 1) ? Q1 : P1,
                              ADV_VAR(2, 2) ? Q2 : P2);
#undef ADV_VAR
  const pk2 dx = pk_mul(cc, u1), dy = pk_mul(cc, v1), dz = pk_mul(cc, w1);
  bad |= adv_bad2(dx, dy, dz, ADV_MAXOFF);
  return adv_tri2<NW>(RF, tbm, kcm, adv_split2(dx, magic), adv_split2(dy, magic),
                      adv_split2(dz, magic));
}

// VEL = 1: u, v, w -> o0, o1, o2.  NSC = 1: one cell-centred scalar (fourth staged field -> os).
// Block = 32 x (NW + 1): NW row warps (rows ty and ty + NW of the tile) + the producer warp.
template <int NW, int VEL, int NSC>
__global__ void __launch_bounds__(32 * (NW + 1), (NW <= 8 && NSC == 0) ? 2 : 1)
k_advect_tma(const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_v,
             const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_s,
             Geo g, const uint8_t* __restrict__ media, float* __restrict__ o0,
             float* __restrict__ o1, float* __restrict__ o2, float* __restrict__ os, float c,
             int* __restrict__ worklist, Scal* __restrict__ sc, int zchunk) {
  constexpr int NF = 3 + NSC;
  constexpr int PL = AdvT<NW>::PL;
  constexpr int ROW = ADVT_ROW;
  constexpr int DB = AdvT<NW>::DB;
  using Smem = AdvTSmem<NW, NF>;
  extern __shared__ __align__(1024) unsigned char smem_adv[];
  Smem& sm_ = *reinterpret_cast<Smem*>(smem_adv);
  const int lane = threadIdx.x, ty = threadIdx.y;
  if (lane == 0 && ty == 0) {
#pragma unroll
    for (int s = 0; s < ADVT_SLOTS; ++s) mbar_init(&sm_.full[s], 1);
    mbar_init(&sm_.empty[0], NW);
    mbar_init(&sm_.empty[1], NW);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();

  const int i0 = blockIdx.x * 32, j0 = blockIdx.y * AdvT<NW>::TY;
  const int kb = blockIdx.z * zchunk;
  const int ke = min(kb + zchunk, g.nz);
  const int pfirst = kb - 3;                 // first staged plane
  const int plast = ke - 1 + 3;              // last staged plane
  auto slot_of = [](int p) { return (p + 64) & (ADVT_SLOTS - 1); };

  if (ty == NW) {
    // ===================== producer warp =====================================================
    if (lane == 0) {
      auto issue = [&](int p) {
        const int s = slot_of(p);
        unsigned long long* fb = &sm_.full[s];
        const int z = p + g.hz;              // tensor z coordinate (allocation plane)
        const int copies = (s == 0) ? 2 : 1;
        mbar_expect_tx(fb, (unsigned)(NF * copies) * AdvT<NW>::BOX_BYTES);
        for (int cpy = 0; cpy < copies; ++cpy) {
          const int sd = cpy ? ADVT_SLOTS : s;
          tma_load_3d(&sm_.ring[0][sd][0], &tm_u, i0 - 4, j0 - 3, z, fb);
          tma_load_3d(&sm_.ring[1][sd][0], &tm_v, i0 - 4, j0 - 3, z, fb);
          tma_load_3d(&sm_.ring[2][sd][0], &tm_w, i0 - 4, j0 - 3, z, fb);
          if (NSC) tma_load_3d(&sm_.ring[NF - 1][sd][0], &tm_s, i0 - 4, j0 - 3, z, fb);
        }
      };
      for (int p = pfirst; p < pfirst + ADVT_SLOTS && p <= plast; ++p) issue(p);
      for (int k = kb; k < ke; ++k) {
        const int p = k + 5;                 // goes into the slot of plane k-3
        if (p > plast) break;
        mbar_wait_bounded(&sm_.empty[(k - kb) & 1], (unsigned)((k - kb) >> 1) & 1u);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        issue(p);
      }
    }
    return;
  }

  // ===================== row warps ===========================================================
  const int i = i0 + lane, ja = j0 + ty, jb = ja + NW;
  const bool act_a = (i < g.nx) && (ja < g.ny);
  const bool act_b = (i < g.nx) && (jb < g.ny);
  const float* __restrict__ RU = &sm_.ring[0][0][0];
  const float* __restrict__ RV = &sm_.ring[1][0][0];
  const float* __restrict__ RW = &sm_.ring[2][0][0];
  const float* __restrict__ RS = &sm_.ring[NF - 1][0][0];
  const int tb = (ty + 3) * ROW + lane + 4;                          // first cell in the raw tile
  const unsigned tbm = (unsigned)tb - ADVT_MAGIC_I * (unsigned)(ROW + 1);
  const pk2 magic = pk(ADVT_MAGIC, ADVT_MAGIC);
  const pk2 hh = pk(0.5f, 0.5f), qd = pk(0.25f, 0.25f);
  const pk2 cc = pk(c, c), hc = pk(0.5f * c, 0.5f * c);
  const long long dbg = (long long)NW * g.nx;                        // second cell in global memory
  long long idx = ((long long)kb * g.ny + ja) * g.nx + i;
#define LD2(R, a, off) pk((R)[(a) + (off)], (R)[(a) + DB + (off)])

  // stage-0 values that a plane hands to the next one (plane k+1 of this iteration is plane k
  // of the next): primed here for k = kb
  for (int p = pfirst; p <= kb + 2; ++p)
    mbar_wait_bounded(&sm_.full[slot_of(p)], 0u);
  pk2 u_c, u_xm, v_c, v_ym, w_zm, w_zm_xp, w_zm_yp;
  {
    const int a0 = slot_of(kb) * PL + tb, am = slot_of(kb - 1) * PL + tb;
    u_c = LD2(RU, a0, 0); u_xm = LD2(RU, a0, -1);
    v_c = LD2(RV, a0, 0); v_ym = LD2(RV, a0, -ROW);
    w_zm = LD2(RW, am, 0); w_zm_xp = LD2(RW, am, 1); w_zm_yp = LD2(RW, am, ROW);
  }
  unsigned med_a = act_a ? media[idx] : (unsigned)SOLID;
  unsigned med_b = act_b ? media[idx + dbg] : (unsigned)SOLID;

  for (int k = kb; k < ke; ++k, idx += g.plane) {
    mbar_wait_bounded(&sm_.full[slot_of(k + 3)], (unsigned)((k + 3 - pfirst) >> 3) & 1u);
    const bool fl_a = med_a == FLUID, fl_b = med_b == FLUID;
    if (k + 1 < ke) {                        // labels of the next plane: a global-memory latency
      if (act_a) med_a = media[idx + g.plane];
      if (act_b) med_b = media[idx + g.plane + dbg];
    }
    const unsigned kcm = (unsigned)(k + 64) - ADVT_MAGIC_I;
    const int a0 = slot_of(k) * PL + tb;
    const int ap = slot_of(k + 1) * PL + tb;
    const pk2 u00 = u_c, v00 = v_c, u_m = u_xm, v_m = v_ym;
    const pk2 w00 = LD2(RW, a0, 0), w_xp = LD2(RW, a0, 1), w_yp = LD2(RW, a0, ROW);
    const pk2 u_zp = LD2(RU, ap, 0), u_zp_xm = LD2(RU, ap, -1);
    const pk2 v_zp = LD2(RV, ap, 0), v_zp_ym = LD2(RV, ap, -ROW);
    pk2 s00 = 0;
    if (NSC) s00 = LD2(RS, a0, 0);
    pk2 nu = u00, nv = v00, nw = w00, ns = s00;
    unsigned bad = 0;                        // cells that go to the worklist
    // warp-uniform: lanes without a Fluid cell compute on staged data and discard the result
    if (__any_sync(0xffffffffu, fl_a || fl_b)) {
      if (VEL) {
        {   // U face at (i + 1/2, j, k): closed-form stage 0 (v, w averaged over 4 faces)
          const pk2 v0 = pk_mul(qd, pk_add(pk_add(v_m, v00),
                                           pk_add(LD2(RV, a0, -ROW + 1), LD2(RV, a0, 1))));
          const pk2 w0 = pk_mul(qd, pk_add(pk_add(w_zm, w00), pk_add(w_zm_xp, w_xp)));
          nu = adv_trace2<NW, 0>(RU, RV, RW, RU, tbm, kcm, pk_mul(hc, u00), pk_mul(hc, v0),
                                 pk_mul(hc, w0), cc, magic, bad);
        }
        {   // V face at (i, j + 1/2, k)
          const pk2 u0 = pk_mul(qd, pk_add(pk_add(u_m, LD2(RU, a0, ROW - 1)),
                                           pk_add(u00, LD2(RU, a0, ROW))));
          const pk2 w0 = pk_mul(qd, pk_add(pk_add(w_zm, w00), pk_add(w_zm_yp, w_yp)));
          nv = adv_trace2<NW, 1>(RU, RV, RW, RV, tbm, kcm, pk_mul(hc, u0), pk_mul(hc, v00),
                                 pk_mul(hc, w0), cc, magic, bad);
        }
        {   // W face at (i, j, k + 1/2)
          const pk2 u0 = pk_mul(qd, pk_add(pk_add(u_m, u_zp_xm), pk_add(u00, u_zp)));
          const pk2 v0 = pk_mul(qd, pk_add(pk_add(v_m, v_zp_ym), pk_add(v00, v_zp)));
          nw = adv_trace2<NW, 2>(RU, RV, RW, RW, tbm, kcm, pk_mul(hc, u0), pk_mul(hc, v0),
                                 pk_mul(hc, w00), cc, magic, bad);
        }
      }
      if (NSC) {   // cell-centred scalar: velocity at the centre = mean of the two faces
        const pk2 u0 = pk_mul(hh, pk_add(u_m, u00));
        const pk2 v0 = pk_mul(hh, pk_add(v_m, v00));
        const pk2 w0 = pk_mul(hh, pk_add(w_zm, w00));
        ns = adv_trace2<NW, 3>(RU, RV, RW, RS, tbm, kcm, pk_mul(hc, u0), pk_mul(hc, v0),
                               pk_mul(hc, w0), cc, magic, bad);
      }
    }
    u_c = u_zp; u_xm = u_zp_xm; v_c = v_zp; v_ym = v_zp_ym;
    w_zm = w00; w_zm_xp = w_xp; w_zm_yp = w_yp;
    // non-Fluid cells keep their values (Convection.fs:60-61 maps over markers); flagged cells
    // keep them until k_advect_rest has recomputed them
    const bool take_a = fl_a && !(bad & 1u), take_b = fl_b && !(bad & 2u);
    if (fl_a && (bad & 1u)) worklist[atomicAdd(&sc->wl_count, 1u)] = (int)idx;
    if (fl_b && (bad & 2u)) worklist[atomicAdd(&sc->wl_count, 1u)] = (int)(idx + dbg);
    if (act_a) {
      if (VEL) {
        o0[idx] = take_a ? pk_lo(nu) : pk_lo(u00);
        o1[idx] = take_a ? pk_lo(nv) : pk_lo(v00);
        o2[idx] = take_a ? pk_lo(nw) : pk_lo(w00);
      }
      if (NSC) os[idx] = take_a ? pk_lo(ns) : pk_lo(s00);
    }
    if (act_b) {
      if (VEL) {
        o0[idx + dbg] = take_b ? pk_hi(nu) : pk_hi(u00);
        o1[idx + dbg] = take_b ? pk_hi(nv) : pk_hi(v00);
        o2[idx + dbg] = take_b ? pk_hi(nw) : pk_hi(w00);
      }
      if (NSC) os[idx + dbg] = take_b ? pk_hi(ns) : pk_hi(s00);
    }
    // this warp is done with plane k: its oldest staged plane (k-3) may be overwritten
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm_.empty[(k - kb) & 1]);
  }
#undef LD2
}

}  // namespace spl
