// spl_mg_march.cuh -- fused, HBM-streaming kernels of the multigrid V-cycle (fp32, nx % 4 = 0)
//
// One z-marching engine (the k_phaseA_ring scheme of spl_pcg.cuh: CTA = 32 x 8 threads,
// a thread owns 4 consecutive x cells of one row, raw operands of the next planes are
// streamed into thread-private slots of a shared-memory ring with cp.async, +-x
// neighbours by warp shuffle, +-y through a triple-buffered tile, +-z in registers)
// with the value the stencil is applied to, v, formed on the fly from the raw operands:
//
//   MG_FIRST2  : v = omega/N * b/scale                 (first Jacobi sweep from u = 0)
//   MG_PLAIN   : v = u
//   MG_PROLONG : v = u + e_coarse[parent] on Fluid cells (piecewise-constant prolongation)
//
// and one of two epilogues:
//
//   MG_SMOOTH   : out = v + omega/N (b/scale - L v)     [+ partial r.out when DOT]
//   MG_RESTRICT : b_coarse[parent] = 1/8 sum over the 8 children of (b - scale L v)
//
// so the fine-level work of a V(2,2) cycle is four passes -- FIRST2 (9 B/cell), RESTRICT
// (9 + 0.5), PROLONG+smooth (13 + 0.5), smooth+dot (13) -- 45 B/cell instead of the 74 of
// the seven one-thread-per-cell kernels in spl_mg.cuh (which stay as the fp64 / ragged /
// tiny-level path and define the arithmetic).  Halo planes k = -1 and k = nz are read
// through the same path as interior planes: they hold zeros (codes 0) on a single GPU.
#pragma once
#include "spl_common.cuh"
#include "spl_pcg.cuh"
#include "spl_mg.cuh"

namespace spl {

#ifndef SPL_MG_RING_D
#define SPL_MG_RING_D 4
#endif
#ifndef SPL_MG_MINB
#define SPL_MG_MINB 3
#endif
constexpr int MG_RING_D = SPL_MG_RING_D;

enum { MG_FIRST2 = 0, MG_PLAIN = 1, MG_PROLONG = 2 };
enum { MG_SMOOTH = 0, MG_RESTRICT = 1 };

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem) : "memory");
}

struct MgMarchArgs {
  Geo g;
  const uint8_t* codes;
  const float* a;          // the field v is formed from: b (FIRST2) or u
  const float* b;          // right-hand side of the level (own cells only)
  float* out;              // smoothed field (MG_SMOOTH)
  float inv_scale, scale;
  float omega;             // weight of the sweep the epilogue performs
  float omega_v;           // MG_FIRST2: weight of the first sweep (forms v)
  int sumform;             // Laplacian as N s - sum (fewer instructions) instead of differences
  // coarse level (MG_PROLONG reads ec, MG_RESTRICT writes bc)
  const float* ec;
  const uint8_t* codes_c;
  float* bc;
  int cnx, cny;
  long long cplane;
  Scal* sc;
  double* partials;
  int par;                 // rho slot (DOT)
  int zchunk;              // even
  // RUPD (level 0, MG_FIRST2): the PCG residual update r -= alpha q is folded into this
  // pass (it replaces k_phaseB): a = r is read together with q, r' is written back
  const float* q;
  float* r_out;
  int nsbuf;
};

template <int MODE, int RUPD = 0>
struct alignas(16) MgStage {
  float4 a[256];
  float4 q[RUPD ? 256 : 1];                    // RUPD: A s of the PCG iteration
  float4 hq[RUPD ? 64 : 1];
  float eq[RUPD ? 16 : 1];
  float4 b[MODE == MG_FIRST2 ? 1 : 256];
  float4 ha[64];                               // CTA-edge rows (warp 0: row j-1, warp 7: row j+8)
  unsigned code[256];
  unsigned hcode[MODE == MG_PLAIN ? 1 : 64];
  float ea[16];                                // CTA-edge lanes
  unsigned ecode[MODE == MG_PLAIN ? 1 : 16];
  float2 pe[MODE == MG_PROLONG ? 256 : 1];     // e_coarse of the two parents of the 4 cells
  float2 hpe[MODE == MG_PROLONG ? 64 : 1];
  float epe[MODE == MG_PROLONG ? 16 : 1];
};

template <int MODE, int OUT, int RUPD = 0>
struct MgSmem {
  MgStage<MODE, RUPD> st[MG_RING_D];
  float4 tile[3][TILE_Y + 2][32];
  float2 rtile[OUT == MG_RESTRICT ? TILE_Y : 1][32];
  double red[32];
  float wl[8];      // omega / N
  float wv[8];      // omega_v / N
  double alpha;     // RUPD
  unsigned long long bar;   // split-phase barrier of the tile hand-over (spl_pcg.cuh)
};

template <int MODE, int OUT, int DOT, int RUPD = 0>
__global__ void __launch_bounds__(32 * TILE_Y, SPL_MG_MINB)
k_mg_march(const MgMarchArgs p) {
  static_assert(!RUPD || (MODE == MG_FIRST2 && OUT == MG_SMOOTH && !DOT), "RUPD is a FIRST2 option");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Smem = MgSmem<MODE, OUT, RUPD>;
  using Stage = MgStage<MODE, RUPD>;
  Smem& sm_ = *reinterpret_cast<Smem*>(smem_raw);
  if (p.sc->done) return;
  const Geo& g = p.g;
  const int tid = threadIdx.x + threadIdx.y * 32;
  if (tid < 8) {
    sm_.wl[tid] = (tid >= 1 && tid <= 6) ? p.omega / float(tid) : 0.f;
    sm_.wv[tid] = (tid >= 1 && tid <= 6) ? p.omega_v / float(tid) : 0.f;
  }
  if (RUPD && tid == 0) {   // alpha of this PCG iteration, exactly as k_phaseB forms it
    const double rho1 = rho_of(p.sc, p.par ^ 1);
    const double sq = sum_slots(p.sc->gA, p.sc->world);
    const double alpha = rho1 / sq;
    if (linear_block() == 0) p.sc->alpha_hist[(p.sc->count + 1) % p.nsbuf] = alpha;
    sm_.alpha = alpha;
  }
  if (tid == 0) mbar_init(&sm_.bar, 32 * TILE_Y);
  __syncthreads();
  const float alpha = RUPD ? (float)sm_.alpha : 0.f;
  const float* wl = sm_.wl;
  const float* wv = sm_.wv;
  const bool sumform = p.sumform != 0;
  const float one_m_omega = 1.f - p.omega;
  const float inv_scale = p.inv_scale;

  const int lane = threadIdx.x, ty = threadIdx.y;
  const int i0 = (blockIdx.x * 32 + lane) * 4;
  const int j = blockIdx.y * TILE_Y + ty;
  const bool active = (i0 < g.nx) && (j < g.ny);
  const int kb = blockIdx.z * p.zchunk;
  const int ke = min(kb + p.zchunk, g.nz);
  const long long row = g.nx, plane = g.plane;
  const int hrow_j = (ty == 0) ? j - 1 : j + 1;
  const bool has_hrow = active && (ty == 0 || ty == TILE_Y - 1) && hrow_j >= 0 && hrow_j < g.ny;
  const long long hoff = (ty == 0) ? -row : row;
  const int hslot = (ty == 0) ? 0 : TILE_Y + 1;
  const int hidx = ((ty == 0) ? 0 : 32) + lane;
  {
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      if (!active) sm_.tile[b][ty + 1][lane] = z4;
      if (ty == 0 && !has_hrow) sm_.tile[b][0][lane] = z4;
      if (ty == TILE_Y - 1 && !has_hrow) sm_.tile[b][TILE_Y + 1][lane] = z4;
    }
  }
  const bool has_e = active && ((lane == 0 && i0 > 0) || (lane == 31 && i0 + 4 < g.nx));
  const long long eoff = (lane == 0) ? -1 : 4;
  const int eidx = ty * 2 + ((lane == 0) ? 0 : 1);
  const int eshift = (lane == 0) ? 24 : 0;
  const long long idx0 = ((long long)kb * g.ny + j) * row + i0;
  // coarse-level offsets of the parents (row part; the plane part moves with k >> 1)
  const long long c_own = (long long)(j >> 1) * p.cnx + (i0 >> 1);
  const long long c_hrow = (long long)(hrow_j >> 1) * p.cnx + (i0 >> 1);
  const long long c_edge = (long long)(j >> 1) * p.cnx + ((i0 + (int)eoff) >> 1);

  auto request = [&](int pl, int slot, long long off) {
    if (active && pl <= ke) {
      Stage& st = sm_.st[slot];
      const bool centre = pl < ke;
      cp_async16(&st.a[tid], p.a + off);
      if (RUPD) cp_async16(&st.q[tid], p.q + off);
      if (MODE != MG_PLAIN || centre) cp_async4(&st.code[tid], p.codes + off);
      long long cpl = 0;
      if (MODE == MG_PROLONG) {
        cpl = (long long)(pl >> 1) * p.cplane;
        cp_async8(&st.pe[tid], p.ec + cpl + c_own);
      }
      if (centre) {
        if (MODE != MG_FIRST2) cp_async16(&st.b[tid], p.b + off);
        if (has_hrow) {
          cp_async16(&st.ha[hidx], p.a + off + hoff);
          if (RUPD) cp_async16(&st.hq[hidx], p.q + off + hoff);
          if (MODE != MG_PLAIN) cp_async4(&st.hcode[hidx], p.codes + off + hoff);
          if (MODE == MG_PROLONG) cp_async8(&st.hpe[hidx], p.ec + cpl + c_hrow);
        }
        if (has_e) {
          cp_async4(&st.ea[eidx], p.a + off + eoff);
          if (RUPD) cp_async4(&st.eq[eidx], p.q + off + eoff);
          if (MODE != MG_PLAIN) cp_async4(&st.ecode[eidx], p.codes + ((off + eoff) & ~3LL));
          if (MODE == MG_PROLONG) cp_async4(&st.epe[eidx], p.ec + cpl + c_edge);
        }
      }
    }
    cp_async_commit();
  };
  // v of one cell from its raw operand, code byte and parent correction
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
  // RUPD: the operand is the updated residual r - alpha q (same expression as k_phaseB)
  auto upd4 = [&](float4 a, float4 q) {
    a.x -= alpha * q.x; a.y -= alpha * q.y; a.z -= alpha * q.z; a.w -= alpha * q.w;
    return a;
  };
  auto combine_own = [&](int slot, unsigned& code_out) {
    const Stage& st = sm_.st[slot];
    code_out = st.code[tid];
    const float2 pe = (MODE == MG_PROLONG) ? st.pe[tid] : make_float2(0.f, 0.f);
    return v4(RUPD ? upd4(st.a[tid], st.q[tid]) : st.a[tid], code_out, pe);
  };
  auto combine_hrow = [&](int slot) {
    const Stage& st = sm_.st[slot];
    const unsigned c = (MODE != MG_PLAIN) ? st.hcode[hidx] : 0u;
    const float2 pe = (MODE == MG_PROLONG) ? st.hpe[hidx] : make_float2(0.f, 0.f);
    return v4(RUPD ? upd4(st.ha[hidx], st.hq[hidx]) : st.ha[hidx], c, pe);
  };
  auto combine_edge = [&](int slot) {
    const Stage& st = sm_.st[slot];
    const unsigned c = (MODE != MG_PLAIN) ? (st.ecode[eidx] >> eshift) & 0xffu : 0u;
    const float pe = (MODE == MG_PROLONG) ? st.epe[eidx] : 0.f;
    float a = st.ea[eidx];
    if (RUPD) a -= alpha * st.eq[eidx];
    return vcell(a, c, pe);
  };
  auto next_slot = [](int s) { return (s + 1 == MG_RING_D) ? 0 : s + 1; };
  auto next_tile = [](int t) { return (t + 1 == 3) ? 0 : t + 1; };

  // ---- prologue -----------------------------------------------------------------
  float4 vm = make_float4(0.f, 0.f, 0.f, 0.f), vcur = vm, vp = vm;
  float leftc = 0.f, rightc = 0.f;
  long long roff = idx0;
  int rslot = 0;
  for (int pl = kb; pl < kb + MG_RING_D - 1; ++pl) {
    request(pl, rslot, roff);
    roff += plane;
    rslot = next_slot(rslot);
  }
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
  cp_async_wait<MG_RING_D - 2>();
  unsigned codec = 0;
  int tcur = 0;
  if (active) {
    vcur = combine_own(0, codec);
    sm_.tile[0][ty + 1][lane] = vcur;
    if (has_hrow) sm_.tile[0][hslot][lane] = combine_hrow(0);
    if (has_e) {
      const float e = combine_edge(0);
      if (lane == 0) leftc = e; else rightc = e;
    }
  }

#if SPL_SPLIT_BARRIER
  mbar_arrive(&sm_.bar);                            // phase 0: tile(kb); see k_phaseA_ring
#endif
  double acc = 0.0;
  float rmax = 0.f;                                 // RUPD
  int rbad = 0;
  float zacc0 = 0.f, zacc1 = 0.f;                   // MG_RESTRICT: x-pair sums of the open z pair
  unsigned ccode = 0;                               // codes of the two parents (bytes 0, 1)
  long long idx = idx0;
  int cslot = 0, nslot = 1;
  for (int c = kb; c < ke; ++c, idx += plane) {
    request(c + MG_RING_D - 1, rslot, roff);
    roff += plane;
    rslot = next_slot(rslot);
    if (OUT == MG_RESTRICT) {
      if (active && !(c & 1) && !(j & 1)) {
        const uint8_t* cc = p.codes_c + (long long)(c >> 1) * p.cplane + c_own;
        ccode = *reinterpret_cast<const unsigned short*>(cc);
      }
    }
    cp_async_wait<MG_RING_D - 2>();
#if SPL_SPLIT_BARRIER
    mbar_wait(&sm_.bar, (unsigned)(c - kb) & 1u);
#endif
    const int tnext = next_tile(tcur);
    float leftn = 0.f, rightn = 0.f;
    unsigned coden = 0;
    if (active) {
      vp = combine_own(nslot, coden);
      if (c + 1 < ke) {
        sm_.tile[tnext][ty + 1][lane] = vp;
        if (has_hrow) sm_.tile[tnext][hslot][lane] = combine_hrow(nslot);
        if (has_e) {
          const float e = combine_edge(nslot);
          if (lane == 0) leftn = e; else rightn = e;
        }
      }
    }
#if SPL_SPLIT_BARRIER
    mbar_arrive(&sm_.bar);
#else
    __syncthreads();
#endif
    const float lsh = __shfl_up_sync(0xffffffffu, vcur.w, 1);
    const float rsh = __shfl_down_sync(0xffffffffu, vcur.x, 1);
    const float left = (lane != 0) ? lsh : leftc;
    const float right = (lane != 31) ? rsh : rightc;
    if (active) {
      const float4 sym = sm_.tile[tcur][ty][lane];
      const float4 syp = sm_.tile[tcur][ty + 2][lane];
      float4 b4 = (MODE == MG_FIRST2) ? sm_.st[cslot].a[tid] : sm_.st[cslot].b[tid];
      if (RUPD) {   // r' of the own cells: stored, and watched for max |r| / non-finite
        b4 = upd4(b4, sm_.st[cslot].q[tid]);
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
          if (sumform) {   // v + w (b' - N v + sum) = (1 - omega) v + w (b' + sum)
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
        __syncthreads();
        if (active && !(ty & 1)) {
          float2 t = sm_.rtile[ty][lane];
          if (j + 1 < g.ny) {
            const float2 u = sm_.rtile[ty + 1][lane];
            t.x += u.x;
            t.y += u.y;
          }
          float2 r;
          const float rw = (float)mg_restrict_weight(g);
          r.x = (ccode & (unsigned)B_FLUID) ? rw * t.x : 0.f;
          r.y = ((ccode >> 8) & (unsigned)B_FLUID) ? rw * t.y : 0.f;
          *reinterpret_cast<float2*>(p.bc + (long long)(c >> 1) * p.cplane + c_own) = r;
        }
      }
    }
    vm = vcur;
    vcur = vp;
    codec = coden;
    leftc = leftn;
    rightc = rightn;
    tcur = tnext;
    cslot = nslot;
    nslot = next_slot(nslot);
  }
  cp_async_wait<0>();

  if (RUPD) {   // max |r| of the iteration into the slot k_phaseB would have filled
    const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
    const int anybad = __syncthreads_or(rbad);
    const double bm = block_max((double)rmax, sm_.red);
    const unsigned nb = num_blocks();
    if (tid == 0) p.partials[MAX_PARTIALS + linear_block()] = anybad ? kNaN : bm;
    if (last_block(&p.sc->ticketB, nb)) {
      double m = 0.0;
      int nanf = 0;
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
      for (unsigned b = tid; b < nb; b += 32 * TILE_Y) t += __ldcg(p.partials + b);
      t = block_sum(t, sm_.red);
      if (tid == 0) p.sc->gBM[p.par][p.sc->rank][0] = t;
    }
  }
}

}  // namespace spl
