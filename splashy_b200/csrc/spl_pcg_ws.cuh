// spl_pcg_ws.cuh -- phase A of the PCG iteration, warp-specialised ring kernel (fp32, nx % 4 = 0)
//
// Same arithmetic, ring and tile scheme as k_phaseA_ring (spl_pcg.cuh), with the CTA-edge
// work moved out of the eight row warps: a ninth warp streams and combines the two halo
// rows (j0-1 and j0+8) and the 16 x-neighbour cells that belong to the adjacent CTAs, and
// hands them over through the shared-memory tile / a small edge array.  In k_phaseA_ring
// warps 0 and 7 carried ~15 % more instructions per plane than the other six (which then
// waited for them at the tile hand-over every plane, profiles/README.md) and every row
// warp issued the predicated edge-lane path; here the eight row warps run identical,
// shorter code and the halo warp (no stencil, no stores) is never the slowest.
#pragma once
#include "spl_common.cuh"
#include "spl_pcg.cuh"

namespace spl {

struct RingSmemWS {
  RingStage st[RING_D];
  float4 tile[3][TILE_Y + 2][32];
  float edge[3][2 * TILE_Y];      // [row * 2 + side]: x neighbour owned by the adjacent CTA
  double red[32];
  float lut[8];
  double beta;
  unsigned long long bar;
  int stop;
};

template <int PC>
__global__ void __launch_bounds__(32 * (TILE_Y + 1), SPL_RING_MINB)
k_phaseA_ws(Geo g, const uint8_t* __restrict__ codes, const float* __restrict__ zr,
            const float* __restrict__ sold, float* __restrict__ snew, float* __restrict__ q,
            Scal* __restrict__ sc, double* __restrict__ partials, int par, int zchunk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RingSmemWS& sm_ = *reinterpret_cast<RingSmemWS*>(smem_raw);
  constexpr int NTHR = 32 * (TILE_Y + 1);
  const int tid = threadIdx.x + threadIdx.y * 32;
  fill_inv_lut<float>(sm_.lut);
  if (tid == 0) {
    phaseA_scalars(sc, par, &sm_.stop, &sm_.beta);
    mbar_init(&sm_.bar, NTHR);
  }
  if (tid < 3 * 2 * TILE_Y) (&sm_.edge[0][0])[tid] = 0.f;
  __syncthreads();
  if (sm_.stop) return;

  const float beta = (float)sm_.beta;
  const float* lut = sm_.lut;
  const int lane = threadIdx.x, ty = threadIdx.y;
  const int i0 = (blockIdx.x * 32 + lane) * 4;
  const int j0 = blockIdx.y * TILE_Y;
  const int kb = blockIdx.z * zchunk;
  const int ke = min(kb + zchunk, g.nz);
  const int nz = g.nz;
  const long long row = g.nx, plane = g.plane;

  auto dirf = [&](float a, float o, unsigned code) {
    const float zz = (PC == PC_JACOBI) ? a * lut[lut_index(code)] : a;
    return zz + beta * o;
  };
  auto dir4 = [&](float4 a, float4 o, unsigned c) {
    float4 r;
    r.x = dirf(a.x, o.x, c & 0xffu);
    r.y = dirf(a.y, o.y, (c >> 8) & 0xffu);
    r.z = dirf(a.z, o.z, (c >> 16) & 0xffu);
    r.w = dirf(a.w, o.w, c >> 24);
    return r;
  };
  auto next_slot = [](int s) { return (s + 1 == RING_D) ? 0 : s + 1; };
  auto next_tile = [](int t) { return (t + 1 == 3) ? 0 : t + 1; };
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);

  double acc = 0.0;
  if (ty < TILE_Y) {
    // ===================== row warps: one row of 128 cells each ===========================
    const int j = j0 + ty;
    const bool active = (i0 < g.nx) && (j < g.ny);
    if (!active) {
#pragma unroll
      for (int b = 0; b < 3; ++b) sm_.tile[b][ty + 1][lane] = z4;
    }
    const long long idx0 = ((long long)kb * g.ny + j) * row + i0;
    auto request = [&](int p, int slot, long long off) {
      if (active && p <= ke) {
        RingStage& st = sm_.st[slot];
        if (p < 0 || p >= nz) {
          cp_async16(&st.a[tid], snew + off);        // halo of s' already holds the new direction
        } else {
          cp_async16(&st.a[tid], zr + off);
          cp_async16(&st.old[tid], sold + off);
          cp_async4(&st.code[tid], codes + off);
        }
      }
      cp_async_commit();
    };
    auto combine_own = [&](int p, int slot, unsigned& code_out) {
      const RingStage& st = sm_.st[slot];
      const float4 a = st.a[tid];
      code_out = 0u;
      if (p < 0 || p >= nz) return a;
      code_out = st.code[tid];
      return dir4(a, st.old[tid], code_out);
    };
    float4 smv = z4, scur = z4, sp = z4;
    long long roff = idx0;
    int rslot = 0;
    for (int p = kb; p < kb + RING_D - 1; ++p) {
      request(p, rslot, roff);
      roff += plane;
      rslot = next_slot(rslot);
    }
    if (active) {                                     // plane kb-1: registers
      const long long idm = idx0 - plane;
      if (kb - 1 < 0) {
        smv = *reinterpret_cast<const float4*>(snew + idm);
      } else {
        smv = dir4(*reinterpret_cast<const float4*>(zr + idm),
                   *reinterpret_cast<const float4*>(sold + idm),
                   *reinterpret_cast<const unsigned*>(codes + idm));
      }
    }
    cp_async_wait<RING_D - 2>();
    unsigned codec = 0;
    int tcur = 0;
    if (active) {
      scur = combine_own(kb, 0, codec);
      sm_.tile[0][ty + 1][lane] = scur;
    }
    mbar_arrive(&sm_.bar);                            // phase 0: tile(kb)
    long long idx = idx0;
    int nslot = 1;
    for (int c = kb; c < ke; ++c, idx += plane) {
      request(c + RING_D - 1, rslot, roff);
      roff += plane;
      rslot = next_slot(rslot);
      cp_async_wait<RING_D - 2>();
      mbar_wait(&sm_.bar, (unsigned)(c - kb) & 1u);   // tile(c) and edge(c) are complete
      const int tnext = next_tile(tcur);
      unsigned coden = 0;
      if (active) {
        sp = combine_own(c + 1, nslot, coden);
        if (c + 1 < ke) sm_.tile[tnext][ty + 1][lane] = sp;
      }
      nslot = next_slot(nslot);
      mbar_arrive(&sm_.bar);
      const float lsh = __shfl_up_sync(0xffffffffu, scur.w, 1);
      const float rsh = __shfl_down_sync(0xffffffffu, scur.x, 1);
      const float left = (lane != 0) ? lsh : sm_.edge[tcur][2 * ty];
      const float right = (lane != 31) ? rsh : sm_.edge[tcur][2 * ty + 1];
      if (active) {
        const float4 sym = sm_.tile[tcur][ty][lane];
        const float4 syp = sm_.tile[tcur][ty + 2][lane];
        const float s[4] = {scur.x, scur.y, scur.z, scur.w};
        const float ym[4] = {sym.x, sym.y, sym.z, sym.w};
        const float yp[4] = {syp.x, syp.y, syp.z, syp.w};
        const float zm[4] = {smv.x, smv.y, smv.z, smv.w};
        const float zp[4] = {sp.x, sp.y, sp.z, sp.w};
        float qa[4];
        float part = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float xm = (e > 0) ? s[e > 0 ? e - 1 : 0] : left;
          const float xp = (e < 3) ? s[e < 3 ? e + 1 : 0] : right;
          qa[e] = lap_cell<float>((codec >> (8 * e)) & 0xffu, s[e], xm, xp, ym[e], yp[e], zm[e],
                                  zp[e]);
          part += s[e] * qa[e];
        }
        acc += (double)part;
        *reinterpret_cast<float4*>(q + idx) = make_float4(qa[0], qa[1], qa[2], qa[3]);
        *reinterpret_cast<float4*>(snew + idx) = scur;
      }
      smv = scur;
      scur = sp;
      codec = coden;
      tcur = tnext;
    }
    cp_async_wait<0>();
  } else {
    // ===================== halo warp: rows j0-1, j0+TILE_Y and the x-edge cells ===========
    const int jA = j0 - 1, jB = j0 + TILE_Y;
    const bool okx = i0 < g.nx;
    const bool hasA = okx && jA >= 0 && j0 < g.ny;
    const bool hasB = okx && jB < g.ny;
    // edge item of this lane: row r of the tile, side 0 = cell left of the CTA, 1 = right
    const int er = lane >> 1, eside = lane & 1;
    const int ei = eside ? (blockIdx.x * 32 + 32) * 4 : blockIdx.x * 32 * 4 - 1;
    const bool hasE = lane < 2 * TILE_Y && (j0 + er) < g.ny && ei >= 0 && ei < g.nx;
    const int eshift = 8 * (ei & 3);
    if (!hasA || !hasB) {
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        if (!hasA) sm_.tile[b][0][lane] = z4;
        if (!hasB) sm_.tile[b][TILE_Y + 1][lane] = z4;
      }
    }
    const long long offA0 = ((long long)kb * g.ny + jA) * row + i0;
    const long long offB0 = ((long long)kb * g.ny + jB) * row + i0;
    const long long offE0 = ((long long)kb * g.ny + (j0 + er)) * row + ei;
    auto request = [&](int p, int slot, long long dp) {   // dp = (p - kb) * plane
      if (p < ke) {
        RingStage& st = sm_.st[slot];
        if (hasA) {
          cp_async16(&st.ha[lane], zr + offA0 + dp);
          cp_async16(&st.hold[lane], sold + offA0 + dp);
          cp_async4(&st.hcode[lane], codes + offA0 + dp);
        }
        if (hasB) {
          cp_async16(&st.ha[32 + lane], zr + offB0 + dp);
          cp_async16(&st.hold[32 + lane], sold + offB0 + dp);
          cp_async4(&st.hcode[32 + lane], codes + offB0 + dp);
        }
        if (hasE) {
          cp_async4(&st.ea[lane], zr + offE0 + dp);
          cp_async4(&st.eold[lane], sold + offE0 + dp);
          cp_async4(&st.ecode[lane], codes + ((offE0 + dp) & ~3LL));
        }
      }
      cp_async_commit();
    };
    auto publish = [&](int slot, int t) {               // combine one plane into tile t
      const RingStage& st = sm_.st[slot];
      if (hasA) sm_.tile[t][0][lane] = dir4(st.ha[lane], st.hold[lane], st.hcode[lane]);
      if (hasB)
        sm_.tile[t][TILE_Y + 1][lane] = dir4(st.ha[32 + lane], st.hold[32 + lane],
                                             st.hcode[32 + lane]);
      if (hasE)
        sm_.edge[t][lane] = dirf(st.ea[lane], st.eold[lane], (st.ecode[lane] >> eshift) & 0xffu);
    };
    long long rdp = 0;
    int rslot = 0;
    for (int p = kb; p < kb + RING_D - 1; ++p) {
      request(p, rslot, rdp);
      rdp += plane;
      rslot = next_slot(rslot);
    }
    cp_async_wait<RING_D - 2>();
    publish(0, 0);
    mbar_arrive(&sm_.bar);                            // phase 0
    int tcur = 0, nslot = 1;
    for (int c = kb; c < ke; ++c) {
      request(c + RING_D - 1, rslot, rdp);
      rdp += plane;
      rslot = next_slot(rslot);
      cp_async_wait<RING_D - 2>();
      mbar_wait(&sm_.bar, (unsigned)(c - kb) & 1u);   // every row warp is done with tile(c-2)
      const int tnext = next_tile(tcur);
      if (c + 1 < ke) publish(nslot, tnext);
      nslot = next_slot(nslot);
      mbar_arrive(&sm_.bar);
      tcur = tnext;
    }
    cp_async_wait<0>();
  }

  const double bs = block_sum(acc, sm_.red);
  const unsigned nb = num_blocks();
  if (tid == 0) partials[linear_block()] = bs;
  if (last_block(&sc->ticketA, nb)) {
    double t = 0.0;
    // same summation order as k_phaseA_ring (stride 32 * TILE_Y): bit-identical s.q
    if (tid < 32 * TILE_Y)
      for (unsigned b = tid; b < nb; b += 32 * TILE_Y) t += __ldcg(partials + b);
    t = block_sum(t, sm_.red);
    if (tid == 0) sc->gA[sc->rank] = t;
  }
}

}  // namespace spl
