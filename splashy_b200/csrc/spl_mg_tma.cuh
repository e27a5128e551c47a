// spl_mg_tma.cuh -- the fused multigrid passes (spl_mg_march.cuh) with a TMA producer warp
//
// Same modes, epilogues and arithmetic as k_mg_march<MODE, OUT, DOT, RUPD>; the load path
// is the one of k_phaseA_tma (spl_pcg_tma.cuh): a ninth warp issues one cp.async.bulk.tensor
// per field and plane for the whole raw tile of the CTA -- (128 + 8) x (8 + 2) cells with the
// halo rows and x-neighbour columns, box edges zero-filled by the hardware; for MG_PROLONG
// also the (64 + 8) x (4 + 2) tile of the coarse correction -- and combines the halo rows and
// the x-edge cells; the eight row warps issue no global loads.  The right-hand side of the
// own cells travels from the combine step to the epilogue in registers, so a ring stage is
// only read while its plane is combined and can be refilled as in k_phaseA_tma.
// Used on levels with nx % 16 = 0 (and an even coarse extent for MG_PROLONG).
#pragma once
#include <cuda.h>

#include "spl_common.cuh"
#include "spl_mg.cuh"
#include "spl_mg_march.cuh"
#include "spl_pcg_tma.cuh"

namespace spl {

constexpr int MGT_RING = 3;            // stages (ring depth 3 and 4 measured equal for these passes)
constexpr int MGT_EX = 64 + 8;         // coarse floats per raw row: cols i0/2-4 .. i0/2+67
constexpr int MGT_EY = TILE_Y / 2 + 2; // coarse rows j0/2-1 .. j0/2+4
constexpr unsigned MGT_BYTES_E = MGT_EX * MGT_EY * 4;

template <int MODE, int RUPD>
struct alignas(128) MgtStage {
  alignas(128) float a[TMA_BY][TMA_BX];                                  // b (FIRST2) or u
  alignas(128) float b[(MODE == MG_FIRST2 && !RUPD) ? 1 : TMA_BY][TMA_BX];  // rhs, or q (RUPD)
  alignas(128) uint8_t code[TMA_BY][TMA_CX];
  alignas(128) float ec[MODE == MG_PROLONG ? MGT_EY : 1][MGT_EX];
};

template <int MODE, int OUT, int RUPD>
struct MgtSmem {
  MgtStage<MODE, RUPD> st[MGT_RING];
  float4 tile[3][TILE_Y + 2][32];
  float edge[3][2 * TILE_Y];
  float2 rtile[OUT == MG_RESTRICT ? TILE_Y : 1][32];
  double red[32];
  float wl[8], wv[8];
  double alpha;
  unsigned long long bar;
  unsigned long long full[MGT_RING];
};

struct MgTmaMaps {
  CUtensorMap a, b, code, ec;     // b: rhs of the level, or q for RUPD; ec: coarse correction
};

template <int MODE, int OUT, int DOT, int RUPD = 0>
__global__ void __launch_bounds__(32 * (TILE_Y + 1), 3)
k_mg_tma(const __grid_constant__ MgTmaMaps tm, const MgMarchArgs p) {
  static_assert(!RUPD || (MODE == MG_FIRST2 && OUT == MG_SMOOTH && !DOT), "RUPD is a FIRST2 option");
  extern __shared__ __align__(1024) unsigned char smem_mgt[];
  using Smem = MgtSmem<MODE, OUT, RUPD>;
  using Stage = MgtStage<MODE, RUPD>;
  Smem& sm_ = *reinterpret_cast<Smem*>(smem_mgt);
  if (p.sc->done) return;
  constexpr int NTHR = 32 * (TILE_Y + 1);
  constexpr bool HAS_B = !(MODE == MG_FIRST2 && !RUPD);   // second float tile in use
  const Geo& g = p.g;
  const int tid = threadIdx.x + threadIdx.y * 32;
  if (tid < 8) {
    sm_.wl[tid] = (tid >= 1 && tid <= 6) ? p.omega / float(tid) : 0.f;
    sm_.wv[tid] = (tid >= 1 && tid <= 6) ? p.omega_v / float(tid) : 0.f;
  }
  if (tid == 0) {
    if (RUPD) {   // alpha of this PCG iteration, exactly as k_phaseB forms it
      const double rho1 = rho_of(p.sc, p.par ^ 1);
      const double sq = sum_slots(p.sc->gA, p.sc->world);
      const double alpha = rho1 / sq;
      if (linear_block() == 0) p.sc->alpha_hist[(p.sc->count + 1) % p.nsbuf] = alpha;
      sm_.alpha = alpha;
    }
    mbar_init(&sm_.bar, NTHR);
#pragma unroll
    for (int s = 0; s < MGT_RING; ++s) mbar_init(&sm_.full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();
  const float alpha = RUPD ? (float)sm_.alpha : 0.f;
  const float* wl = sm_.wl;
  const float* wv = sm_.wv;
  const bool sumform = p.sumform != 0;
  const float one_m_omega = 1.f - p.omega;
  const float inv_scale = p.inv_scale;

  const int lane = threadIdx.x, ty = threadIdx.y;
  const int i0 = (blockIdx.x * 32 + lane) * 4;
  const int j0 = blockIdx.y * TILE_Y;
  const int kb = blockIdx.z * p.zchunk;
  const int ke = min(kb + p.zchunk, g.nz);
  const long long row = g.nx, plane = g.plane;

  auto vcell = [&](float a, unsigned code, float pe) {
    if (MODE == MG_FIRST2) return wv[__popc(code & 63u)] * (a * inv_scale);
    if (MODE == MG_PROLONG) return (code & (unsigned)B_FLUID) ? a + pe : 0.f;
    return a;
  };
  auto v4 = [&](float4 a, unsigned c, float2 pe) {
    float4 r;
    r.x = vcell(a.x, c & 0xffu, pe.x);
    r.y = vcell(a.y, (c >> 8) & 0xffu, pe.x);
    r.z = vcell(a.z, (c >> 16) & 0xffu, pe.y);
    r.w = vcell(a.w, c >> 24, pe.y);
    return r;
  };
  auto upd4 = [&](float4 a, float4 q) {
    a.x -= alpha * q.x; a.y -= alpha * q.y; a.z -= alpha * q.z; a.w -= alpha * q.w;
    return a;
  };
  auto next_tile = [](int t) { return (t + 1 == 3) ? 0 : t + 1; };
  auto slot_of = [&](int pl) { return (pl - kb) % MGT_RING; };
  auto wait_full = [&](int pl) {
    mbar_wait_bounded(&sm_.full[slot_of(pl)], (unsigned)((pl - kb) / MGT_RING) & 1u);
  };
  // raw row r (0 = halo row j0-1, 1..8 = own rows, 9 = halo row j0+8) of plane pl, the 4 cells
  // of this lane: the operand the stencil value is formed from (r' for RUPD), its code word,
  // and v
  auto row_v = [&](int pl, int r, float4& opnd, unsigned& code) {
    const Stage& st = sm_.st[slot_of(pl)];
    opnd = *reinterpret_cast<const float4*>(&st.a[r][4 + 4 * lane]);
    if (RUPD) opnd = upd4(opnd, *reinterpret_cast<const float4*>(&st.b[r][4 + 4 * lane]));
    code = *reinterpret_cast<const unsigned*>(&st.code[r][16 + 4 * lane]);
    float2 pe = make_float2(0.f, 0.f);
    if (MODE == MG_PROLONG) {
      const int er = (r == 0) ? 0 : ((r == TILE_Y + 1) ? MGT_EY - 1 : ((r - 1) >> 1) + 1);
      pe = *reinterpret_cast<const float2*>(&st.ec[er][4 + 2 * lane]);
    }
    return v4(opnd, code, pe);
  };
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);

  double acc = 0.0;
  float rmax = 0.f;
  int rbad = 0;
  if (ty < TILE_Y) {
    // ===================== row warps ======================================================
    const int j = j0 + ty;
    const bool active = (i0 < g.nx) && (j < g.ny);
    const long long idx0 = ((long long)kb * g.ny + j) * row + i0;
    const long long c_own = (long long)(j >> 1) * p.cnx + (i0 >> 1);
    float4 vm = z4, vcur = z4, vp = z4, bcur = z4, bnext = z4;
    if (active) {                                     // plane kb-1 through registers
      const long long idm = idx0 - plane;
      float4 a = *reinterpret_cast<const float4*>(p.a + idm);
      if (RUPD) a = upd4(a, *reinterpret_cast<const float4*>(p.q + idm));
      unsigned c = 0u;
      float2 pe = make_float2(0.f, 0.f);
      if (MODE != MG_PLAIN) c = *reinterpret_cast<const unsigned*>(p.codes + idm);
      if (MODE == MG_PROLONG)
        pe = *reinterpret_cast<const float2*>(p.ec + (long long)((kb - 1) >> 1) * p.cplane + c_own);
      vm = v4(a, c, pe);
    }
    // right-hand side of the own cells on plane pl (FIRST2: the operand itself)
    auto own_b = [&](int pl, float4 opnd) {
      if (MODE == MG_FIRST2) return opnd;
      return *reinterpret_cast<const float4*>(&sm_.st[slot_of(pl)].b[ty + 1][4 + 4 * lane]);
    };
    wait_full(kb);
    unsigned codec = 0;
    int tcur = 0;
    {
      float4 opnd;
      vcur = row_v(kb, ty + 1, opnd, codec);
      bcur = own_b(kb, opnd);
      sm_.tile[0][ty + 1][lane] = vcur;
    }
    mbar_arrive(&sm_.bar);                            // phase 0: tile(kb)
    float zacc0 = 0.f, zacc1 = 0.f;
    unsigned ccode = 0;
    long long idx = idx0;
    for (int c = kb; c < ke; ++c, idx += plane) {
      if (OUT == MG_RESTRICT) {
        if (active && !(c & 1) && !(j & 1)) {
          const uint8_t* cc = p.codes_c + (long long)(c >> 1) * p.cplane + c_own;
          ccode = *reinterpret_cast<const unsigned short*>(cc);
        }
      }
      wait_full(c + 1);
      mbar_wait(&sm_.bar, (unsigned)(c - kb) & 1u);
      const int tnext = next_tile(tcur);
      unsigned coden = 0;
      {
        float4 opnd;
        vp = row_v(c + 1, ty + 1, opnd, coden);
        bnext = own_b(c + 1, opnd);
      }
      if (c + 1 < ke) sm_.tile[tnext][ty + 1][lane] = vp;
      mbar_arrive(&sm_.bar);
      const float lsh = __shfl_up_sync(0xffffffffu, vcur.w, 1);
      const float rsh = __shfl_down_sync(0xffffffffu, vcur.x, 1);
      const float left = (lane != 0) ? lsh : sm_.edge[tcur][2 * ty];
      const float right = (lane != 31) ? rsh : sm_.edge[tcur][2 * ty + 1];
      if (active) {
        const float4 sym = sm_.tile[tcur][ty][lane];
        const float4 syp = sm_.tile[tcur][ty + 2][lane];
        const float4 b4 = bcur;
        if (RUPD) {   // r' of the own cells: stored, and watched for max |r| / non-finite
          *reinterpret_cast<float4*>(p.r_out + idx) = b4;
          const float m4 = fmaxf(fmaxf(fabsf(b4.x), fabsf(b4.y)), fmaxf(fabsf(b4.z), fabsf(b4.w)));
          rbad |= !(fabsf(b4.x) <= 3.402823466e38f) | !(fabsf(b4.y) <= 3.402823466e38f) |
                  !(fabsf(b4.z) <= 3.402823466e38f) | !(fabsf(b4.w) <= 3.402823466e38f);
          rmax = fmaxf(rmax, m4);
        }
        const float s[4] = {vcur.x, vcur.y, vcur.z, vcur.w};
        const float ym[4] = {sym.x, sym.y, sym.z, sym.w};
        const float yp[4] = {syp.x, syp.y, syp.z, syp.w};
        const float zm[4] = {vm.x, vm.y, vm.z, vm.w};
        const float zp[4] = {vp.x, vp.y, vp.z, vp.w};
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
        float o[4];
        float part = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const unsigned code = (codec >> (8 * e)) & 0xffu;
          const float xm = (e > 0) ? s[e > 0 ? e - 1 : 0] : left;
          const float xp = (e < 3) ? s[e < 3 ? e + 1 : 0] : right;
          const int n = __popc(code & 63u);
          if (OUT == MG_SMOOTH) {
            const float w = wl[n];
            float r;
            if (sumform) {
              const float sum = ((xm + xp) + (ym[e] + yp[e])) + (zm[e] + zp[e]);
              r = fmaf(w, fmaf(bb[e], inv_scale, sum), one_m_omega * s[e]);
            } else {
              const float lap = lap_cell<float>(code, s[e], xm, xp, ym[e], yp[e], zm[e], zp[e]);
              r = s[e] + w * (bb[e] * inv_scale - lap);
            }
            o[e] = (w != 0.f) ? r : 0.f;
            if (DOT) part += bb[e] * o[e];
          } else {
            float lap;
            if (sumform) {
              const float sum = ((xm + xp) + (ym[e] + yp[e])) + (zm[e] + zp[e]);
              lap = fmaf(float(n), s[e], -sum);
            } else {
              lap = lap_cell<float>(code, s[e], xm, xp, ym[e], yp[e], zm[e], zp[e]);
            }
            o[e] = (code & (unsigned)B_FLUID) ? bb[e] - p.scale * lap : 0.f;
          }
        }
        if (OUT == MG_SMOOTH) {
          if (DOT) acc += (double)part;
          *reinterpret_cast<float4*>(p.out + idx) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
          zacc0 += o[0] + o[1];
          zacc1 += o[2] + o[3];
        }
      }
      if (OUT == MG_RESTRICT) {
        if ((c & 1) || c == g.nz - 1) {               // the z pair is complete (uniform per CTA)
          sm_.rtile[ty][lane] = make_float2(zacc0, zacc1);
          zacc0 = zacc1 = 0.f;
          // the eight row warps only (the producer warp never touches rtile)
          asm volatile("bar.sync 1, %0;\n" ::"n"(32 * TILE_Y) : "memory");
          if (active && !(ty & 1)) {
            float2 t = sm_.rtile[ty][lane];
            if (j + 1 < g.ny) {
              const float2 u = sm_.rtile[ty + 1][lane];
              t.x += u.x;
              t.y += u.y;
            }
            const float rw = (float)mg_restrict_weight(g);
            float2 r;
            r.x = (ccode & (unsigned)B_FLUID) ? rw * t.x : 0.f;
            r.y = ((ccode >> 8) & (unsigned)B_FLUID) ? rw * t.y : 0.f;
            *reinterpret_cast<float2*>(p.bc + (long long)(c >> 1) * p.cplane + c_own) = r;
          }
        }
      }
      vm = vcur;
      vcur = vp;
      bcur = bnext;
      codec = coden;
      tcur = tnext;
    }
  } else {
    // ===================== producer / halo warp =============================================
    const int x0f = blockIdx.x * 128 - 4, x0c = blockIdx.x * 128 - 16, y0 = j0 - 1;
    const int x0e = blockIdx.x * 64 - 4, y0e = (j0 >> 1) - 1;
    auto issue = [&](int pl) {                          // lane 0 only
      Stage& st = sm_.st[slot_of(pl)];
      unsigned long long* fb = &sm_.full[slot_of(pl)];
      unsigned bytes = TMA_BYTES_F + TMA_BYTES_C;
      if (HAS_B) bytes += TMA_BYTES_F;
      if (MODE == MG_PROLONG) bytes += MGT_BYTES_E;
      mbar_expect_tx(fb, bytes);
      tma_load_3d(&st.a[0][0], &tm.a, x0f, y0, pl + g.hz, fb);
      if (HAS_B) tma_load_3d(&st.b[0][0], &tm.b, x0f, y0, pl + g.hz, fb);
      tma_load_3d(&st.code[0][0], &tm.code, x0c, y0, pl + g.hz, fb);
      if (MODE == MG_PROLONG) tma_load_3d(&st.ec[0][0], &tm.ec, x0e, y0e, (pl >> 1) + 1, fb);
    };
    const int er = lane >> 1, eside = lane & 1;
    auto publish = [&](int pl, int t) {
      const Stage& st = sm_.st[slot_of(pl)];
      float4 opnd;
      unsigned code;
      sm_.tile[t][0][lane] = row_v(pl, 0, opnd, code);
      sm_.tile[t][TILE_Y + 1][lane] = row_v(pl, TILE_Y + 1, opnd, code);
      if (lane < 2 * TILE_Y) {
        const int xf = eside ? 4 + 128 : 3, xc = eside ? 16 + 128 : 15;
        float a = st.a[er + 1][xf];
        if (RUPD) a -= alpha * st.b[er + 1][xf];
        float pe = 0.f;
        if (MODE == MG_PROLONG) pe = st.ec[(er >> 1) + 1][eside ? 4 + 64 : 3];
        sm_.edge[t][lane] = vcell(a, st.code[er + 1][xc], pe);
      }
    };
    if (lane == 0)
      for (int pl = kb; pl < kb + MGT_RING - 1; ++pl)
        if (pl <= ke) issue(pl);
    wait_full(kb);
    publish(kb, 0);
    mbar_arrive(&sm_.bar);
    int tcur = 0;
    for (int c = kb; c < ke; ++c) {
      if (lane == 0 && c + MGT_RING - 1 <= ke) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        issue(c + MGT_RING - 1);
      }
      wait_full(c + 1);
      mbar_wait(&sm_.bar, (unsigned)(c - kb) & 1u);
      const int tnext = next_tile(tcur);
      if (c + 1 < ke) publish(c + 1, tnext);
      mbar_arrive(&sm_.bar);
      tcur = tnext;
    }
  }

  if (RUPD) {   // max |r| of the iteration into the slot k_phaseB would have filled
    const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
    const int anybad = __syncthreads_or(rbad);
    const double bm = block_max((double)rmax, sm_.red);
    const unsigned nb = num_blocks();
    if (tid == 0) p.partials[MAX_PARTIALS + linear_block()] = anybad ? kNaN : bm;
    if (last_block(&p.sc->ticketB, nb)) {
      double m = 0.0;
      int nanf = 0;
      if (tid < 32 * TILE_Y)
        for (unsigned b = tid; b < nb; b += 32 * TILE_Y) {
          const double v = __ldcg(p.partials + MAX_PARTIALS + b);
          if (v != v) nanf = 1; else m = fmax(m, v);
        }
      nanf = __syncthreads_or(nanf);
      m = block_max(m, sm_.red);
      if (tid == 0) {
        p.sc->gBM[p.par][p.sc->rank][1] = nanf ? kNaN : m;
        p.sc->count = p.sc->count + 1;
      }
    }
  }
  if (DOT) {
    const double bs = block_sum(acc, sm_.red);
    const unsigned nb = num_blocks();
    if (tid == 0) p.partials[linear_block()] = bs;
    if (last_block(&p.sc->ticketC, nb)) {
      double t = 0.0;
      if (tid < 32 * TILE_Y)
        for (unsigned b = tid; b < nb; b += 32 * TILE_Y) t += __ldcg(p.partials + b);
      t = block_sum(t, sm_.red);
      if (tid == 0) p.sc->gBM[p.par][p.sc->rank][0] = t;
    }
  }
}

}  // namespace spl
