// spl_pcg_tma.cuh -- phase A of the PCG iteration with a TMA producer warp (fp32, nx % 16 = 0)
//
// k_phaseA_ws (spl_pcg_ws.cuh) with the operand streaming moved from per-thread cp.async to
// the Tensor Memory Accelerator: the ninth warp -- which in k_phaseA_ws idles at the tile
// hand-over most of the time -- issues one cp.async.bulk.tensor per field and plane for the
// whole raw tile of the CTA, (128 + 8) x (8 + 2) cells including the two halo rows and the
// x-neighbour columns; cells outside the box are zero-filled by the hardware, so every
// box-edge predicate of the load path disappears and the eight row warps no longer spend
// instructions on address arithmetic, cp.async issue and wait_group.  Completion is tracked
// per ring stage with an mbarrier (expect_tx); a stage is refilled one plane after the tile
// hand-over barrier has shown that every thread is done with it.
// Arithmetic, tile scheme, reduction order: identical to k_phaseA_ws / k_phaseA_ring.
#pragma once
#include <cuda.h>

#include "spl_common.cuh"
#include "spl_pcg.cuh"

namespace spl {

constexpr int TMA_BX = 128 + 8;        // floats per raw row: cells i0-4 .. i0+131
constexpr int TMA_BY = TILE_Y + 2;     // rows j0-1 .. j0+TILE_Y
constexpr int TMA_CX = 128 + 32;       // code bytes per raw row: cells i0-16 .. i0+143
constexpr unsigned TMA_BYTES_F = TMA_BX * TMA_BY * 4;
constexpr unsigned TMA_BYTES_C = TMA_CX * TMA_BY;

struct alignas(128) TmaStage {
  alignas(128) float a[TMA_BY][TMA_BX];       // z or r (or the s' halo plane of a remote slab)
  alignas(128) float old[TMA_BY][TMA_BX];     // previous search direction
  alignas(128) uint8_t code[TMA_BY][TMA_CX];
};

struct TmaSmem {
  TmaStage st[RING_D];
  float4 tile[3][TILE_Y + 2][32];
  float edge[3][2 * TILE_Y];
  double red[32];
  float lut[8];
  double beta;
  unsigned long long bar;             // tile hand-over (split phase, all threads)
  unsigned long long full[RING_D];    // TMA completion of a stage (1 arrival + tx bytes)
  int stop;
};

__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(a),
               "r"(bytes)
               : "memory");
}
// bounded spin: a wrong byte count or descriptor must abort the launch, not hang the GPU.
// try_wait itself may suspend the thread for a system-dependent time, so the bound is wall
// time (4 s on %globaltimer, sampled every 256 failed polls), not a poll count.
__device__ __forceinline__ void mbar_wait_bounded(unsigned long long* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  unsigned long long t0 = 0;
  for (unsigned spin = 1;; ++spin) {
    unsigned ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) return;
    if ((spin & 255u) == 0u) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            unsigned long long* bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(d),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(b)
      : "memory");
}

template <int PC>
__global__ void __launch_bounds__(32 * (TILE_Y + 1), SPL_RING_MINB)
k_phaseA_tma(const __grid_constant__ CUtensorMap tm_zr, const __grid_constant__ CUtensorMap tm_sold,
             const __grid_constant__ CUtensorMap tm_snew, const __grid_constant__ CUtensorMap tm_codes,
             Geo g, const uint8_t* __restrict__ codes, const float* __restrict__ zr,
             const float* __restrict__ sold, float* __restrict__ snew, float* __restrict__ q,
             Scal* __restrict__ sc, double* __restrict__ partials, int par, int zchunk,
             int zb0, int zstr, unsigned nbt, Fused fz) {
  // z-chunk of this CTA = zb0 + blockIdx.z * zstr; nbt = CTAs of the whole phase (a slab runs
  // its interior chunks while the halo planes travel and the two boundary chunks afterwards:
  // two launches that share the partials array and the ticket)
  extern __shared__ __align__(1024) unsigned char smem_tma[];
  TmaSmem& sm_ = *reinterpret_cast<TmaSmem*>(smem_tma);
  constexpr int NTHR = 32 * (TILE_Y + 1);
  const int tid = threadIdx.x + threadIdx.y * 32;
  fill_inv_lut<float>(sm_.lut);
  if (tid == 0) {
    // z-slabs over peer memory: rho, max|r| of the other ranks
    if (!sc->done) fused_sync_B(fz, sc, par ^ 1, (blockIdx.x | blockIdx.y | blockIdx.z) == 0);
    phaseA_scalars(sc, par, &sm_.stop, &sm_.beta);
    mbar_init(&sm_.bar, NTHR);
#pragma unroll
    for (int s = 0; s < RING_D; ++s) mbar_init(&sm_.full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();
  if (sm_.stop) return;

  const float beta = (float)sm_.beta;
  const float* lut = sm_.lut;
  const int lane = threadIdx.x, ty = threadIdx.y;
  const int i0 = (blockIdx.x * 32 + lane) * 4;
  const int j0 = blockIdx.y * TILE_Y;
  // fz.on: one launch for the whole slab, with the two boundary chunks (the only ones that
  // read a halo plane) dispatched last -- long after the neighbours' planes have arrived
  const int bz = (fz.halo && !(fz.mode & 4)) ? (int)((blockIdx.z + 1u) % gridDim.z)
                                             : zb0 + (int)blockIdx.z * zstr;
  const int kb = bz * zchunk;
  const int ke = min(kb + zchunk, g.nz);
  const int nz = g.nz;
  const long long row = g.nx, plane = g.plane;
  if (fz.halo) {
    if (tid == 0) {     // plain polls: the planes are read from L2 (TMA, ld.cg) after the stamp
      const FusedBox* mine = fz.box[sc->rank];
      if (kb == 0 && sc->rank > 0) spin_local(&mine->hseq[0], fz.stamp);
      if (ke == nz && sc->rank < sc->world - 1) spin_local(&mine->hseq[1], fz.stamp);
    }
    __syncthreads();
    asm volatile("fence.proxy.async;\n" ::: "memory");   // the planes were written by generic stores
  }

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
  auto next_tile = [](int t) { return (t + 1 == 3) ? 0 : t + 1; };
  auto slot_of = [&](int p) { return (p - kb) % RING_D; };
  auto wait_full = [&](int p) {
    mbar_wait_bounded(&sm_.full[slot_of(p)], (unsigned)((p - kb) / RING_D) & 1u);
  };
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);

  double acc = 0.0;
  if (ty < TILE_Y) {
    // ===================== row warps ======================================================
    const int j = j0 + ty;
    const bool active = (i0 < g.nx) && (j < g.ny);
    const long long idx0 = ((long long)kb * g.ny + j) * row + i0;
    // raw operands of the thread's 4 cells on plane p (zero outside the box)
    auto combine_own = [&](int p, unsigned& code_out) {
      const TmaStage& st = sm_.st[slot_of(p)];
      const float4 a = *reinterpret_cast<const float4*>(&st.a[ty + 1][4 + 4 * lane]);
      code_out = 0u;
      if (p < 0 || p >= nz) return a;              // remote plane: s' itself
      code_out = *reinterpret_cast<const unsigned*>(&st.code[ty + 1][16 + 4 * lane]);
      return dir4(a, *reinterpret_cast<const float4*>(&st.old[ty + 1][4 + 4 * lane]), code_out);
    };
    float4 smv = z4, scur = z4, sp = z4;
    if (active) {                                     // plane kb-1: registers
      const long long idm = idx0 - plane;
      if (kb - 1 < 0) {
        smv = __ldcg(reinterpret_cast<const float4*>(snew + idm));
      } else {
        smv = dir4(*reinterpret_cast<const float4*>(zr + idm),
                   *reinterpret_cast<const float4*>(sold + idm),
                   *reinterpret_cast<const unsigned*>(codes + idm));
      }
    }
    wait_full(kb);
    unsigned codec = 0;
    int tcur = 0;
    scur = combine_own(kb, codec);
    sm_.tile[0][ty + 1][lane] = scur;
    mbar_arrive(&sm_.bar);                            // phase 0: tile(kb)
    long long idx = idx0;
    for (int c = kb; c < ke; ++c, idx += plane) {
      wait_full(c + 1);                               // raw tile of plane c+1 has landed
      mbar_wait(&sm_.bar, (unsigned)(c - kb) & 1u);   // tile(c), edge(c) complete
      const int tnext = next_tile(tcur);
      unsigned coden = 0;
      sp = combine_own(c + 1, coden);
      if (c + 1 < ke) sm_.tile[tnext][ty + 1][lane] = sp;
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
  } else {
    // ===================== producer / halo warp =============================================
    const int x0f = blockIdx.x * 128 - 4, x0c = blockIdx.x * 128 - 16, y0 = j0 - 1;
    auto issue = [&](int p) {                           // lane 0 only
      TmaStage& st = sm_.st[slot_of(p)];
      unsigned long long* fb = &sm_.full[slot_of(p)];
      const bool remote = (p < 0) || (p >= nz);
      mbar_expect_tx(fb, remote ? TMA_BYTES_F : 2 * TMA_BYTES_F + TMA_BYTES_C);
      if (remote) {
        tma_load_3d(&st.a[0][0], &tm_snew, x0f, y0, p + g.hz, fb);
      } else {
        tma_load_3d(&st.a[0][0], &tm_zr, x0f, y0, p + g.hz, fb);
        tma_load_3d(&st.old[0][0], &tm_sold, x0f, y0, p + g.hz, fb);
        tma_load_3d(&st.code[0][0], &tm_codes, x0c, y0, p + g.hz, fb);
      }
    };
    const int er = lane >> 1, eside = lane & 1;         // edge item of lanes 0..15
    auto publish = [&](int p, int t) {                  // halo rows + x-edge cells of a centre plane
      const TmaStage& st = sm_.st[slot_of(p)];
      sm_.tile[t][0][lane] =
          dir4(*reinterpret_cast<const float4*>(&st.a[0][4 + 4 * lane]),
               *reinterpret_cast<const float4*>(&st.old[0][4 + 4 * lane]),
               *reinterpret_cast<const unsigned*>(&st.code[0][16 + 4 * lane]));
      sm_.tile[t][TILE_Y + 1][lane] =
          dir4(*reinterpret_cast<const float4*>(&st.a[TILE_Y + 1][4 + 4 * lane]),
               *reinterpret_cast<const float4*>(&st.old[TILE_Y + 1][4 + 4 * lane]),
               *reinterpret_cast<const unsigned*>(&st.code[TILE_Y + 1][16 + 4 * lane]));
      if (lane < 2 * TILE_Y) {
        const int xf = eside ? 4 + 128 : 3, xc = eside ? 16 + 128 : 15;
        sm_.edge[t][lane] = dirf(st.a[er + 1][xf], st.old[er + 1][xf], st.code[er + 1][xc]);
      }
    };
    if (lane == 0)
      for (int p = kb; p < kb + RING_D - 1; ++p)
        if (p <= ke) issue(p);
    wait_full(kb);
    publish(kb, 0);
    mbar_arrive(&sm_.bar);                              // phase 0
    int tcur = 0;
    for (int c = kb; c < ke; ++c) {
      // the stage of plane c-1 was last read while tile(c-1) was formed; the hand-over phase
      // this warp waited for in the previous iteration covers that for every thread
      if (lane == 0 && c + RING_D - 1 <= ke) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        issue(c + RING_D - 1);
      }
      wait_full(c + 1);
      mbar_wait(&sm_.bar, (unsigned)(c - kb) & 1u);
      const int tnext = next_tile(tcur);
      if (c + 1 < ke) publish(c + 1, tnext);
      mbar_arrive(&sm_.bar);
      tcur = tnext;
    }
  }

  const double bs = block_sum(acc, sm_.red);
  const unsigned nb = nbt;
  if (tid == 0) partials[blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * (unsigned)bz)] = bs;
  if (last_block(&sc->ticketA, nb)) {
    double t = 0.0;
    if (tid < 32 * TILE_Y)
      for (unsigned b = tid; b < nb; b += 32 * TILE_Y) t += __ldcg(partials + b);
    t = block_sum(t, sm_.red);
    if (tid == 0) {
      sc->gA[sc->rank] = t;
      fused_push_A(fz, sc, t);
    }
  }
}

}  // namespace spl
