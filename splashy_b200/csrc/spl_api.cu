// spl_api.cu -- C ABI of libsplashy_cuda.so (include/splashy_cuda.h)
//
// Host-side runtime of the advect -> project path: owns the device arrays that
// replace the reference's global Dictionary<Coord,Cell> (Grid.fs:12), sequences
// the kernels of spl_advect.cuh / spl_stencil.cuh / spl_pcg.cuh on one CUDA
// stream, and -- when the box is one z-slab of a larger grid -- exchanges halo
// planes and the CG scalars with the neighbouring ranks through NCCL (loaded
// with dlopen so a single-GPU host needs no NCCL at all).
//
// There is no CPU path in this library.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#include <type_traits>

#include "../../include/splashy_cuda.h"
#include "spl_advect.cuh"
#include "spl_advect_tma.cuh"
#include "spl_common.cuh"
#include "spl_mg.cuh"
#include "spl_pcg.cuh"
#include "spl_pcg_ws.cuh"
#include "spl_pcg_tma.cuh"
#include "spl_mg_tma.cuh"
#include "spl_mg_march.cuh"
#include "spl_markers.cuh"
#include "spl_stencil.cuh"

using namespace spl;

// ---------------------------------------------------------------------------
// NCCL through dlopen
// ---------------------------------------------------------------------------
namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t,
                       cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t,
                       cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

bool load_nccl(std::string& err) {
  if (g_nccl.lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* n : names) {
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) {
    err = std::string("cannot load libnccl: ") + dlerror();
    return false;
  }
#define SPL_SYM(field, name)                                        \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(lib, name)); \
  if (!g_nccl.field) { err = std::string("libnccl lacks ") + name; return false; }
  SPL_SYM(GetUniqueId, "ncclGetUniqueId")
  SPL_SYM(CommInitRank, "ncclCommInitRank")
  SPL_SYM(CommDestroy, "ncclCommDestroy")
  SPL_SYM(AllGather, "ncclAllGather")
  SPL_SYM(AllReduce, "ncclAllReduce")
  SPL_SYM(Send, "ncclSend")
  SPL_SYM(Recv, "ncclRecv")
  SPL_SYM(GroupStart, "ncclGroupStart")
  SPL_SYM(GroupEnd, "ncclGroupEnd")
  SPL_SYM(GetErrorString, "ncclGetErrorString")
#undef SPL_SYM
  g_nccl.lib = lib;
  return true;
}

thread_local std::string g_create_error;

}  // namespace

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
// ---- peer-to-peer exchange of the per-rank scalars (world > 1) ------------------------------
// Every rank owns a PeerBox that the other ranks of the node map through CUDA IPC.  A gather
// is one tiny kernel: thread w stores this rank's value(s) and then the sequence number of the
// gather into rank w's box (release: __threadfence_system between the two), then waits for
// rank w's stamp in its own box and copies the value into the local slot array.  A rank
// writes before it waits, so no rank depends on a kernel that has not been launched yet by its
// own stream.  Four value slots indexed by the sequence number: a rank is at most one gather
// ahead of the slowest one.  The box is never reset by the host (a reset could race with a
// faster peer's store).
struct PeerBox {
  double val[4][MAXW][2];
  unsigned long long seq[4][MAXW];
};
struct PeerShared {        // one IPC-shared allocation per rank
  PeerBox gather;
  FusedBox fused;
};
struct PeerPtrs {
  PeerBox* p[MAXW];
};

__global__ void k_peer_gather(PeerPtrs peers, int rank, int world, unsigned long long seq,
                              double* __restrict__ slots, int per_rank) {
  const int w = threadIdx.x;
  if (w >= world) return;
  const int b = (int)(seq & 3ull);
  const double v0 = slots[rank * per_rank];
  const double v1 = per_rank > 1 ? slots[rank * per_rank + 1] : 0.0;
  if (w != rank) {
    PeerBox* dst = peers.p[w];
    volatile double* dv = &dst->val[b][rank][0];
    dv[0] = v0;
    dv[1] = v1;
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(&dst->seq[b][rank]) = seq;
    PeerBox* mine = peers.p[rank];
    volatile unsigned long long* f = &mine->seq[b][w];
    unsigned long long t0 = 0;
    for (unsigned spin = 1; *f != seq; ++spin) {
      if ((spin & 1023u) == 0u) {        // a missing rank must abort the launch, not hang the GPU
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 8000000000ull) __trap();
      }
    }
    __threadfence_system();
    volatile double* sv = &mine->val[b][w][0];
    slots[w * per_rank] = sv[0];
    if (per_rank > 1) slots[w * per_rank + 1] = sv[1];
  }
}

// ---- halo planes over peer memory: the mailbox exchange (world > 1) ---------------------------
// Every rank owns one IPC-shared mailbox: a header of sequence words and, per side (the
// neighbour below / above) two data slots.  An exchange of boundary planes is ONE launch:
//   sender CTAs    store this rank's bottom planes into the lower neighbour's mailbox and its
//                  top planes into the upper neighbour's (NVLink stores), one system-scope fence
//                  per CTA, and the last of them releases the exchange's sequence number into the
//                  neighbours' headers;
//   receiver CTAs  wait (bounded) for the neighbours' sequence number in the OWN header -- local
//                  memory -- and copy the slots into the halo planes.
// Senders never wait, and both kinds are resident together (<= 2 x 48 CTAs), so no rank depends
// on a launch its own stream has not issued.  Every rank issues the same exchanges in the same
// order on its compute stream; a neighbour can be at most one exchange ahead (it needs this
// rank's planes for the next one), hence two slots indexed by the parity of the sequence number.
// Works for any field of any multigrid level and any element size: the mailbox is the only
// mapped buffer.  An alternative to ncclSend / ncclRecv in exchange_halo / exchange_level, opt-in
// with SPL_MAILBOX=1 (see map_mailboxes: measured, not faster than NCCL on NVLink 5).
struct alignas(256) MailHdr {
  unsigned long long seq[2][2];      // [side][slot]: the exchange whose planes are complete
};
struct MailArgs {
  const char* src[2];                // my planes for the neighbour below / above
  char* dst[2];                      // its mailbox slot (mapped), nullptr = no neighbour
  unsigned long long* dst_seq[2];
  const char* slot[2];               // my mailbox slot filled by the neighbour below / above
  const unsigned long long* my_seq[2];
  char* halo[2];                     // my halo planes below / above
  size_t bytes;
  unsigned long long seq;
  int unit;                          // bytes per access: 16, 8, 4 or 1
};

template <typename V>
__device__ __forceinline__ void mail_copy(char* dst, const char* src, size_t bytes, int part,
                                          int parts, bool cg) {
  const size_t n = bytes / sizeof(V);
  V* d = reinterpret_cast<V*>(dst);
  const V* s = reinterpret_cast<const V*>(src);
  for (size_t i = (size_t)part * blockDim.x + threadIdx.x; i < n; i += (size_t)parts * blockDim.x)
    d[i] = cg ? __ldcg(s + i) : s[i];
}
__device__ __forceinline__ void mail_copy_any(char* dst, const char* src, size_t bytes, int unit,
                                              int part, int parts, bool cg) {
  if (unit == 16) mail_copy<uint4>(dst, src, bytes, part, parts, cg);
  else if (unit == 8) mail_copy<unsigned long long>(dst, src, bytes, part, parts, cg);
  else if (unit == 4) mail_copy<unsigned>(dst, src, bytes, part, parts, cg);
  else mail_copy<unsigned char>(dst, src, bytes, part, parts, cg);
}

__global__ void __launch_bounds__(256)
k_mail_exchange(MailArgs a, unsigned* ticket, int nsend) {
  if ((int)blockIdx.x < nsend) {
    for (int side = 0; side < 2; ++side)
      if (a.dst[side]) mail_copy_any(a.dst[side], a.src[side], a.bytes, a.unit, blockIdx.x, nsend, false);
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      const unsigned t = atomicAdd(ticket, 1u);
      if (t == (unsigned)nsend - 1u) {
        *ticket = 0u;
        __threadfence_system();
        for (int side = 0; side < 2; ++side)
          if (a.dst[side]) st_release_sys(a.dst_seq[side], a.seq);
      }
    }
  } else {
    const int part = (int)blockIdx.x - nsend, parts = (int)gridDim.x - nsend;
    if (threadIdx.x == 0)
      for (int side = 0; side < 2; ++side)
        if (a.my_seq[side]) spin_until(a.my_seq[side], a.seq);
    __syncthreads();
    for (int side = 0; side < 2; ++side)
      if (a.my_seq[side]) mail_copy_any(a.halo[side], a.slot[side], a.bytes, a.unit, part, parts, true);
  }
}

struct MgLevel {
  Geo g{};
  uint8_t *media = nullptr, *codes = nullptr;   // plane-0 pointers
  char *u = nullptr, *b = nullptr, *tmp = nullptr;
  double scale = 1.0;                           // operator = scale * integer Laplacian
  void* allocs[5]{};
};

struct spl_ctx {
  spl_desc d{};
  Geo g{};
  size_t esz = 4;              // bytes per real
  int vec = 1;                 // x cells per thread in the vector kernels
  int sms = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[6]{};
  cudaEvent_t tev[2]{};        // spl_timer_*
  cudaEvent_t* pev = nullptr;  // spl_profile_pcg: 4 events per iteration
  cudaEvent_t xev[128]{};      // spl_profile_pcg: around the x-flush launches
  int profile_flushes = 0;
  int profile_iters = 0;       // > 0 while spl_profile_pcg runs
  // fields, allocation start (plane -hz); *_0 = plane 0
  void* alloc[64]{};
  char *u = nullptr, *v = nullptr, *w = nullptr;     // current velocity (plane 0)
  char *u2 = nullptr, *v2 = nullptr, *w2 = nullptr;  // advection target
  char *su = nullptr, *sv = nullptr, *sw = nullptr;  // snapshot
  char *p = nullptr, *r = nullptr, *q = nullptr;
  char* sbuf[MAX_SBUF]{};      // ring of search-direction buffers (plane 0 pointers)
  int nsbuf = 16;              // m: x is updated every m iterations (SPL_NSBUF)
  char *z = nullptr, *einv = nullptr;
  uint8_t *media = nullptr, *codes = nullptr, *layer = nullptr;
  Scal* sc = nullptr;          // device
  Scal* hsc = nullptr;         // pinned mirror
  double* partials = nullptr;  // 2 * MAX_PARTIALS
  int* flag = nullptr;         // device int (quirk "too far")
  int* hflag = nullptr;        // pinned
  int* hdone = nullptr;        // pinned [2]: lagged convergence poll
  double* hagree = nullptr;    // pinned [MAXW + 1]: agree_max
  cudaEvent_t dev[2]{};
  // device aliases of the mapped pinned scalars below (to_host / from_host)
  Scal* hsc_d = nullptr;
  int *hflag_d = nullptr, *hdone_d = nullptr;
  unsigned long long *hesc = nullptr, *hesc_d = nullptr;   // trace escapes of the last advection
  ncclComm_t comm = nullptr;
  // z-slabs: the direction halo travels on a second (high-priority) stream while the interior
  // chunks of phase A run (SPL_OVERLAP=0: everything on one stream); the per-rank CG scalars
  // are exchanged by peer-to-peer stores into IPC-mapped boxes (SPL_P2P=0: ncclAllGather)
  cudaStream_t cstream = nullptr;
  // spl_upload_async / spl_download_async: one copy stream per direction (PCIe is full duplex)
  cudaStream_t h2d = nullptr, d2h = nullptr;
  cudaEvent_t ev_up = nullptr, ev_down = nullptr, ev_done = nullptr;
  bool upload_pending = false;   // copies are queued on h2d; the label-derived data is not built yet
  bool download_queued = false;  // ev_down marks the end of the last queued download
  cudaEvent_t ev_beta = nullptr, ev_halo = nullptr;
  bool overlap = true;
  bool split = false;          // SPL_SPLIT=1: interior / boundary launches of phase A around NCCL send / recv
  bool p2p = false;
  PeerShared* pbox = nullptr;
  PeerPtrs peers{};
  FusedBox* fboxes[MAXW]{};    // every rank's FusedBox (own + IPC-mapped)
  void* peer_maps[MAXW]{};     // IPC mappings to close
  char* nb_sbuf[2][MAX_SBUF]{};   // the neighbours' (below / above) direction buffers, allocation starts
  void* nb_maps[2][MAX_SBUF]{};
  // mailbox exchange of halo planes (k_mail_exchange)
  char* mbox = nullptr;           // own mailbox: MailHdr + 2 sides x 2 slots
  size_t mbox_slot = 0;           // bytes per slot
  char* nb_mbox[2]{};             // the neighbours' mailboxes (below / above), IPC-mapped
  unsigned long long xseq = 0;    // exchanges issued so far
  unsigned* xticket = nullptr;
  bool mbox_ok = false;
  bool fused_ok = false;       // the PCG iteration runs over peer memory (SPL_FUSED=0: separate gathers)
  unsigned long long pseq = 0;
  unsigned long long iter_seq = 0;
  unsigned* hticket = nullptr;
  std::vector<int> slab_geo;   // world > 1: {nz, z0} of every rank
  bool uploaded = false, have_rhs = false, have_snapshot = false;
  int fixed_iters = 0;
  // iterative refinement: after a converged solve, restart PCG from the true fp64 residual
  // (k_true_residual) keeping x; default 1 round for fp32 + SPL_PC_MG (SPL_REFINE, spl_set_refinement)
  int refine_rounds = 0;
  bool pcg_keep_x = false;
  double pcg_tol_first = 0.0;  // > 0: stopping tolerance of the pass before a refinement round
  double pcg_bmax_keep = 0.0;
  int iters_base = 0;
  double refine_first_tol = 1e-4;   // SPL_REFINE_FIRST_TOL
  // marker stage
  long long nmark = 0;
  double* mloc = nullptr;      // nmark x 3 locations (metres)
  long long *mold = nullptr, *mnew = nullptr;   // cells before / after the last move
  uint8_t* mark = nullptr;     // per cell: holds a marker / existed before relabelling (plane 0)
  void* mark_base = nullptr;   // its allocation (z-halo planes included)
  MarkerStats* mstats = nullptr;   // device
  WorldBox world{};            // World.contains as box indices
  std::vector<MgLevel> mg;     // SPL_PC_MG hierarchy; level 0 aliases codes / z / r / q
  int mg_nu = 2;               // pre- and post-smoothing sweeps (SPL_MG_NU)
  int mg_coarse = 30;          // sweeps on the coarsest level, even (SPL_MG_COARSE)
  std::vector<double> mg_omega;  // per-sweep Jacobi weights: Chebyshev on [1/3, 2] (SPL_MG_OMEGA=w: constant)
  bool mg_sumform = false;     // SPL_MG_SUMFORM=1: N s - sum Laplacian in the marching kernels
  // world > 1: the coarsest level is gathered and solved as ONE global grid on every rank
  bool mg_gbottom = false;
  Geo mg_gg{};                 // global geometry of the coarsest level
  uint8_t* mg_gcodes = nullptr;
  char *mg_gb = nullptr, *mg_gu = nullptr;
  int mg_gsweeps = 100;        // Jacobi sweeps of the global coarsest solve (SPL_MG_GSWEEPS)
  bool mg_no_fuse = false;     // SPL_MG_NO_FUSE=1: k_phaseB stays a separate launch
  bool mg_no_bottom = false;   // SPL_MG_NO_BOTTOM=1: per-level launches down to the coarsest level
  size_t mg_bottom_smem = 0;
  bool mg_no_march = false;    // SPL_MG_NO_MARCH=1: one-thread-per-cell kernels on every level
  int mg_attr_set = 0;
  int zchunk_override = 0;
  int ring_attr_set = 0;
  bool no_2d_view = false;     // SPL_NO_2D_VIEW=1: 2-D boxes use the 3-D tiling
  bool no_ring = false;        // SPL_NO_RING=1: force the generic phase-A kernel
  // TMA phase A (default when nx % 16 = 0; SPL_TMA=0 disables): tensor maps per field, keyed by
  // the field's plane-0 pointer
  bool use_tma = true;
  std::vector<std::pair<const void*, CUtensorMap>> tmaps;
  std::vector<int> tmap_kind;
  int tma_l2 = 3;              // CUtensorMapL2promotion of the tensor maps (SPL_TMA_L2: 0 none, 1 64B, 2 128B, 3 256B: 0.3821 vs 0.3848 ms)
  bool mg_no_tma = false;      // SPL_MG_TMA=0: multigrid passes with per-thread cp.async (k_mg_march)
  int mgt_attr_set = 0;
  bool no_ws = false;          // SPL_NO_WS=1: ring kernel without the halo warp (k_phaseA_ring)
  bool mg_no_fuse_slabs = false;   // SPL_MG_NO_FUSE_SLABS=1: k_phaseB instead of the fused update on slabs
  bool no_quad = false;        // SPL_NO_QUAD=1: one-cell-per-thread div / grad / check kernels
  bool no_rb_march = false;    // SPL_NO_RB_MARCH=1: k_rb_black4 / k_rb_red4 instead of k_rb_march
  bool no_gc_fuse = false;     // SPL_NO_GC_FUSE=1: k_grad_sub4 + k_check4 instead of k_grad_check4
  // cell-centred scalar fields advected with the velocity (spl_set_scalars); scal2 = target
  std::vector<char*> scal;
  char* scal2 = nullptr;
  // tile-staged advection (spl_advect_tma.cuh; default for fp32, nx % 4 = 0; SPL_ADV_TMA=0: the
  // L1-gather kernel k_advect)
  bool adv_tma = true;
  bool adv_fuse = true;        // SPL_ADV_FUSE=0: the first scalar gets its own pass
  int adv_zchunk = 0;          // SPL_ADV_ZCHUNK: planes per CTA (0 = automatic)
  int adv_attr_set = 0;
  int64_t launches = 0;
  double rhs_dt = 0.0;
  std::string err;
};

extern "C" {
static int finish_upload(spl_ctx* c);   // completes a pending spl_upload_async
}

namespace {

int fail(spl_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf; else g_create_error = buf;
  return code;
}

#define CU(call)                                                               \
  do {                                                                         \
    cudaError_t e_ = (call);                                                   \
    if (e_ != cudaSuccess)                                                     \
      return fail(c, SPL_E_CUDA, "CUDA error %s at %s:%d (%s)",                \
                  cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
  } while (0)

#define NC(call)                                                               \
  do {                                                                         \
    ncclResult_t r_ = (call);                                                  \
    if (r_ != ncclSuccess)                                                     \
      return fail(c, SPL_E_CUDA, "NCCL error %d at %s:%d (%s)", (int)r_,       \
                  __FILE__, __LINE__, g_nccl.GetErrorString(r_));              \
  } while (0)

#define LAUNCH(c, kernel, grid, block, ...)                                    \
  do {                                                                         \
    kernel<<<grid, block, 0, (c)->stream>>>(__VA_ARGS__);                      \
    (c)->launches++;                                                           \
  } while (0)

inline int grid_for(const spl_ctx* c, long long n, int block = 256) {
  long long need = (n + block - 1) / block;
  long long cap = (long long)c->sms * 8;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

size_t field_bytes(const spl_ctx* c, size_t es) {
  return (size_t)c->g.plane * (size_t)(c->g.nz + 2 * c->g.hz) * es;
}

int alloc_field(spl_ctx* c, int slot, size_t es, char** plane0, int fill) {
  void* ptr = nullptr;
  const size_t bytes = field_bytes(c, es);
  CU(cudaMalloc(&ptr, bytes));
  CU(cudaMemsetAsync(ptr, fill, bytes, c->stream));
  c->alloc[slot] = ptr;
  *plane0 = static_cast<char*>(ptr) + (size_t)c->g.hz * c->g.plane * es;
  return SPL_OK;
}

ncclDataType_t nccl_real(const spl_ctx* c) {
  return c->d.dtype == SPL_F64 ? ncclDouble : ncclFloat;
}

// boundary planes through the mailboxes: `lo_src` / `hi_src` = this rank's bottom / top planes,
// `lo_halo` / `hi_halo` = where the neighbours' planes belong, `bytes` each
static int gcd_align(size_t x) { return (x % 16 == 0) ? 16 : (x % 8 == 0) ? 8 : (x % 4 == 0) ? 4 : 1; }
int mail_exchange(spl_ctx* c, const char* lo_src, const char* hi_src, char* lo_halo, char* hi_halo,
                  size_t bytes) {
  const int rank = c->d.rank, world = c->d.world;
  const unsigned long long seq = ++c->xseq;
  const int sl = (int)(seq & 1ull);
  MailArgs a{};
  a.bytes = bytes;
  a.seq = seq;
  auto slot_of = [&](char* box, int side) {
    return box + sizeof(MailHdr) + ((size_t)side * 2 + sl) * c->mbox_slot;
  };
  auto seq_of = [&](char* box, int side) { return &reinterpret_cast<MailHdr*>(box)->seq[side][sl]; };
  int unit = gcd_align(bytes);
  auto fold = [&](const void* p) {
    const int al = gcd_align((size_t)reinterpret_cast<uintptr_t>(p));
    if (al < unit) unit = al;
  };
  if (rank > 0) {                       // I am the lower neighbour's "above" side (1)
    a.src[0] = lo_src;
    a.dst[0] = slot_of(c->nb_mbox[0], 1);
    a.dst_seq[0] = seq_of(c->nb_mbox[0], 1);
    a.slot[0] = slot_of(c->mbox, 0);
    a.my_seq[0] = seq_of(c->mbox, 0);
    a.halo[0] = lo_halo;
    fold(lo_src); fold(lo_halo);
  }
  if (rank < world - 1) {
    a.src[1] = hi_src;
    a.dst[1] = slot_of(c->nb_mbox[1], 0);
    a.dst_seq[1] = seq_of(c->nb_mbox[1], 0);
    a.slot[1] = slot_of(c->mbox, 1);
    a.my_seq[1] = seq_of(c->mbox, 1);
    a.halo[1] = hi_halo;
    fold(hi_src); fold(hi_halo);
  }
  a.unit = unit;
  // one sender + one receiver CTA per 64 KiB, at most 48 + 48 (all resident at once)
  int n = (int)((bytes + 65535) / 65536);
  if (n < 1) n = 1;
  if (n > 48) n = 48;
  k_mail_exchange<<<2 * n, 256, 0, c->stream>>>(a, c->xticket, n);
  c->launches++;
  CU(cudaGetLastError());
  return SPL_OK;
}

// exchange `width` z-planes of a field with both slab neighbours
int exchange_halo(spl_ctx* c, char* plane0, size_t es, ncclDataType_t ty, int width,
                  cudaStream_t st = nullptr) {
  if (c->d.world <= 1) return SPL_OK;
  if (!st) st = c->stream;
  const Geo& g = c->g;
  if (width > g.hz) width = g.hz;
  if (width > g.nz) width = g.nz;
  const size_t cnt = (size_t)g.plane * width;
  const size_t pb = (size_t)g.plane * es;
  if (c->mbox_ok && st == c->stream && cnt * es <= c->mbox_slot)
    return mail_exchange(c, plane0, plane0 + (size_t)(g.nz - width) * pb,
                         plane0 - (size_t)width * pb, plane0 + (size_t)g.nz * pb, cnt * es);
  NC(g_nccl.GroupStart());
  if (c->d.rank > 0) {
    NC(g_nccl.Send(plane0, cnt, ty, c->d.rank - 1, c->comm, st));
    NC(g_nccl.Recv(plane0 - (size_t)width * pb, cnt, ty, c->d.rank - 1, c->comm, st));
  }
  if (c->d.rank < c->d.world - 1) {
    NC(g_nccl.Send(plane0 + (size_t)(g.nz - width) * pb, cnt, ty, c->d.rank + 1, c->comm,
                   st));
    NC(g_nccl.Recv(plane0 + (size_t)g.nz * pb, cnt, ty, c->d.rank + 1, c->comm, st));
  }
  NC(g_nccl.GroupEnd());
  return SPL_OK;
}

// one z-plane of a field of any multigrid level (plane0 = its plane 0, halo >= 1)
int exchange_level(spl_ctx* c, const Geo& g, void* plane0_, size_t es, ncclDataType_t ty) {
  if (c->d.world <= 1) return SPL_OK;
  char* plane0 = static_cast<char*>(plane0_);
  const size_t cnt = (size_t)g.plane;
  const size_t pb = cnt * es;
  if (c->mbox_ok && pb <= c->mbox_slot)
    return mail_exchange(c, plane0, plane0 + (size_t)(g.nz - 1) * pb, plane0 - pb,
                         plane0 + (size_t)g.nz * pb, pb);
  NC(g_nccl.GroupStart());
  if (c->d.rank > 0) {
    NC(g_nccl.Send(plane0, cnt, ty, c->d.rank - 1, c->comm, c->stream));
    NC(g_nccl.Recv(plane0 - pb, cnt, ty, c->d.rank - 1, c->comm, c->stream));
  }
  if (c->d.rank < c->d.world - 1) {
    NC(g_nccl.Send(plane0 + (size_t)(g.nz - 1) * pb, cnt, ty, c->d.rank + 1, c->comm, c->stream));
    NC(g_nccl.Recv(plane0 + (size_t)g.nz * pb, cnt, ty, c->d.rank + 1, c->comm, c->stream));
  }
  NC(g_nccl.GroupEnd());
  return SPL_OK;
}

int exchange_velocity(spl_ctx* c, int width) {
  int rc;
  if ((rc = exchange_halo(c, c->u, c->esz, nccl_real(c), width))) return rc;
  if ((rc = exchange_halo(c, c->v, c->esz, nccl_real(c), width))) return rc;
  return exchange_halo(c, c->w, c->esz, nccl_real(c), width);
}

// complete the per-rank slots of the PCG scalars on every rank
int gather_slots(spl_ctx* c, double* slots, int per_rank) {
  if (c->d.world <= 1) return SPL_OK;
  if (c->p2p) {
    k_peer_gather<<<1, 32, 0, c->stream>>>(c->peers, c->d.rank, c->d.world, ++c->pseq, slots,
                                           per_rank);
    c->launches++;
    return SPL_OK;
  }
  NC(g_nccl.AllGather(slots + (size_t)c->d.rank * per_rank, slots, per_rank, ncclDouble,
                      c->comm, c->stream));
  return SPL_OK;
}

// world > 1: the maximum over all ranks of a per-rank value (one collective; every rank must
// call it).  Used to agree on an error before any rank returns early: a rank that left while
// the others went on into the next halo exchange would leave them waiting for ever.
int agree_max(spl_ctx* c, double v, double* out) {
  *out = v;
  if (c->d.world <= 1) return SPL_OK;
  c->hagree[MAXW] = v;
  CU(cudaMemcpyAsync(&c->sc->gD[c->d.rank], &c->hagree[MAXW], sizeof(double),
                     cudaMemcpyHostToDevice, c->stream));
  int rc = gather_slots(c, c->sc->gD, 1);
  if (rc) return rc;
  CU(cudaMemcpyAsync(c->hagree, c->sc->gD, sizeof(double) * c->d.world, cudaMemcpyDeviceToHost,
                     c->stream));
  CU(cudaStreamSynchronize(c->stream));
  for (int w = 0; w < c->d.world; ++w) *out = fmax(*out, c->hagree[w]);
  return SPL_OK;
}

// Scalars between host and device travel as a one-block kernel through MAPPED pinned memory,
// not as cudaMemcpyAsync: a copy of a few bytes queues on the copy engine behind whatever bulk
// transfer another stream has in flight (spl_upload_async / spl_download_async of a second box:
// 512 MiB pieces, ~10 ms each) and the compute stream would wait for it -- measured: 27 ms per
// step in bench.py's two-box end-to-end leg before this change.
__global__ void k_copy_words(unsigned* __restrict__ dst, const unsigned* __restrict__ src, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
}
// bytes % 4 == 0; `host_dev` = the device alias of a mapped pinned buffer
static int to_host(spl_ctx* c, void* host_dev, const void* dev, size_t bytes) {
  k_copy_words<<<1, 256, 0, c->stream>>>(static_cast<unsigned*>(host_dev),
                                        static_cast<const unsigned*>(dev), (int)(bytes / 4));
  CU(cudaGetLastError());
  return SPL_OK;
}
static int from_host(spl_ctx* c, void* dev, const void* host_dev, size_t bytes) {
  k_copy_words<<<1, 256, 0, c->stream>>>(static_cast<unsigned*>(dev),
                                        static_cast<const unsigned*>(host_dev), (int)(bytes / 4));
  CU(cudaGetLastError());
  return SPL_OK;
}

int pull_scal(spl_ctx* c) {
  int rc = to_host(c, c->hsc_d, c->sc, sizeof(Scal));
  if (rc) return rc;
  CU(cudaStreamSynchronize(c->stream));
  return SPL_OK;
}

__global__ void k_set_bmax(Scal* sc) { sc->bmax = rmax_of(sc, 0); }
__global__ void k_force_bmax(Scal* sc, double v) { sc->bmax = v; }

template <typename A, typename B>
__global__ void k_convert(long long n, const A* __restrict__ in, B* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = (B)in[i];
}

__global__ void k_finalize(Scal* sc, int par_last) {
  if (sc->done) return;
  const double rmax = rmax_of(sc, par_last);
  sc->rmax = rmax;
  sc->iters = sc->count;
  const bool bad = !(rmax == rmax);
  sc->breakdown = bad ? 1 : 0;
  sc->converged = (!bad && rmax <= sc->tol * sc->bmax) ? 1 : 0;
}

// planes each phase-A CTA marches through.  Long chunks amortise the pipeline
// prologue (2-3 extra planes of loads per chunk), short ones give more CTAs and a
// smaller tail; ~2048 CTAs (4-5 waves at 3 CTAs/SM) measured best at 512^3
// (z-chunk sweep in profiles/README.md).
int phaseA_zchunk(const spl_ctx* c, int tiles_xy) {
  if (c->zchunk_override > 0) return c->zchunk_override;
  const int nz = c->g.nz;
  const int target_blocks = 2048;
  int gz = (target_blocks + tiles_xy - 1) / tiles_xy;
  if (gz < 1) gz = 1;
  int zc = (nz + gz - 1) / gz;
  if (zc < 8) zc = 8;
  if (zc > nz) zc = nz;
  while (zc < nz && (long long)tiles_xy * ((nz + zc - 1) / zc) > MAX_PARTIALS) zc *= 2;
  return zc;
}

// ---- TMA tensor maps (k_phaseA_tma) -------------------------------------------------------
// A field is described as a rank-3 tensor (x fastest) over its whole allocation, halo planes
// included (z coordinate = plane + hz); the box is the raw tile of one CTA and one plane.
// cuTensorMapEncodeTiled is fetched from the driver at run time (no link against libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// fp32 raw tile, u8 code tile, coarse-correction tile, advection raw tiles (8 / 16 row warps)
enum { TM_F = 0, TM_C = 1, TM_E = 2, TM_A8 = 3, TM_A16 = 4 };

const CUtensorMap* tensor_map_for(spl_ctx* c, const Geo& g, const void* plane0, int kind) {
  for (auto& kv : c->tmaps)
    if (kv.first == plane0 && c->tmap_kind[&kv - &c->tmaps[0]] == kind) return &kv.second;
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) !=
            cudaSuccess || !fn)
      return nullptr;
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  if (c->tmaps.size() >= c->tmaps.capacity()) return nullptr;   // pointers handed out must stay valid
  const bool is_codes = kind == TM_C;
  const size_t es = is_codes ? 1 : 4;
  char* base = static_cast<char*>(const_cast<void*>(plane0)) - (size_t)g.hz * g.plane * es;
  const cuuint64_t dims[3] = {(cuuint64_t)g.nx, (cuuint64_t)g.ny, (cuuint64_t)(g.nz + 2 * g.hz)};
  const cuuint64_t strides[2] = {(cuuint64_t)g.nx * es, (cuuint64_t)g.plane * es};
  cuuint32_t box[3] = {(cuuint32_t)TMA_BX, (cuuint32_t)TMA_BY, 1u};
  if (kind == TM_C) box[0] = TMA_CX;
  if (kind == TM_E) { box[0] = MGT_EX; box[1] = MGT_EY; }
  if (kind == TM_A8) { box[0] = ADVT_ROW; box[1] = AdvT<8>::ROWS; }
  if (kind == TM_A16) { box[0] = ADVT_ROW; box[1] = AdvT<16>::ROWS; }
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  CUtensorMap m;
  const CUresult r = encode(&m, is_codes ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                            3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)c->tma_l2,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return nullptr;
  c->tmaps.emplace_back(plane0, m);
  c->tmap_kind.push_back(kind);
  return &c->tmaps.back().second;
}

// ---- typed implementation -------------------------------------------------------
template <typename T>
struct Impl {
  static T* P(char* p) { return reinterpret_cast<T*>(p); }

  // tile-staged advection (spl_advect_tma.cuh): one launch of k_advect_tma<NW, VEL, NSC>
  // followed by the worklist pass.  s_in / s_out: the scalar of an NSC = 1 pass.
  template <int NW, int VEL, int NSC>
  static bool advect_tma_pass(spl_ctx* c, float cc, const char* s_in, char* s_out) {
    if constexpr (std::is_same<T, float>::value) {
      const Geo& g = c->g;
      const int kind = NW == 8 ? TM_A8 : TM_A16;
      const CUtensorMap* mu = tensor_map_for(c, g, c->u, kind);
      const CUtensorMap* mv = tensor_map_for(c, g, c->v, kind);
      const CUtensorMap* mw = tensor_map_for(c, g, c->w, kind);
      const CUtensorMap* ms = NSC ? tensor_map_for(c, g, s_in, kind) : mu;
      if (!mu || !mv || !mw || !ms) return false;
      using Smem = AdvTSmem<NW, 3 + NSC>;
      const int bit = 1 << ((NW == 16 ? 4 : 0) + VEL * 2 + NSC);
      if (!(c->adv_attr_set & bit)) {
        if (cudaFuncSetAttribute(k_advect_tma<NW, VEL, NSC>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(Smem)) != cudaSuccess)
          return false;
        c->adv_attr_set |= bit;
      }
      const int tx = (g.nx + 31) / 32, ty = (g.ny + 2 * NW - 1) / (2 * NW);
      // planes per CTA: 64 (6 extra planes are staged per chunk; 2.40 ms at 512^3 against 2.45 /
      // 2.42 for ~100 / 128), shorter while the grid has fewer than two waves of CTAs
      int zc = c->adv_zchunk;
      if (zc <= 0) {
        const long long slots = (long long)c->sms * ((NW <= 8 && NSC == 0) ? 2 : 1);
        zc = 64;
        while (zc > 16 && (long long)tx * ty * ((g.nz + zc - 1) / zc) < 2 * slots) zc /= 2;
      }
      if (zc > g.nz) zc = g.nz;
      int gz = (g.nz + zc - 1) / zc;
      while (gz > 65535) { zc *= 2; gz = (g.nz + zc - 1) / zc; }
      int* worklist = reinterpret_cast<int*>(c->q);
      cudaMemsetAsync(&c->sc->wl_count, 0, sizeof(unsigned), c->stream);
      k_advect_tma<NW, VEL, NSC><<<dim3(tx, ty, gz), dim3(32, NW + 1), sizeof(Smem), c->stream>>>(
          *mu, *mv, *mw, *ms, g, c->media, P(c->u2), P(c->v2), P(c->w2), P(s_out), cc, worklist,
          c->sc, zc);
      c->launches++;
      LAUNCH(c, k_advect_rest<T>, grid_for(c, g.ncell) < 592 ? grid_for(c, g.ncell) : 592, 256, g,
             P(c->u), P(c->v), P(c->w), P(c->u2), P(c->v2), P(c->w2),
             NSC ? reinterpret_cast<const T*>(s_in) : nullptr, P(s_out), VEL, cc, worklist, c->sc);
      return true;
    }
    return false;
  }

  static int advect(spl_ctx* c, double dt) {
    const Geo& g = c->g;
    int rc = exchange_velocity(c, g.hz);
    if (rc) return rc;
    for (char* s : c->scal)
      if ((rc = exchange_halo(c, s, c->esz, nccl_real(c), g.hz))) return rc;
    const int grid = grid_for(c, g.ncell);
    const T cc = T(-(dt / c->d.h));
    size_t first_scalar = 0;
    if (c->d.quirks) {
      QuirkCfg q{c->d.quirks, c->d.origin[0], c->d.origin[1], c->d.origin[2] + g.z0, c->d.h};
      CU(cudaMemsetAsync(c->flag, 0, sizeof(int), c->stream));
      LAUNCH(c, k_advect_quirk<T>, grid, 256, g, q, c->media, P(c->u), P(c->v), P(c->w),
             P(c->u2), P(c->v2), P(c->w2), dt, c->sc, c->flag);
    } else {
      bool done = false;
      if constexpr (std::is_same<T, float>::value) {
        if (c->adv_tma && c->use_tma && g.nx % 4 == 0) {
          if (!c->scal.empty() && c->adv_fuse) {      // velocity + the first scalar in one pass
            done = advect_tma_pass<16, 1, 1>(c, cc, c->scal[0], c->scal2);
            if (done) { std::swap(c->scal[0], c->scal2); first_scalar = 1; }
          }
          if (!done) done = advect_tma_pass<8, 1, 0>(c, cc, nullptr, nullptr);
          if (done)
            for (size_t i = first_scalar; i < c->scal.size(); ++i) {
              if (!advect_tma_pass<16, 0, 1>(c, cc, c->scal[i], c->scal2)) break;
              std::swap(c->scal[i], c->scal2);
              first_scalar = i + 1;
            }
        }
      }
      if (!done) {
        const dim3 agrid((g.nx + 31) / 32, (g.ny + 7) / 8, g.nz < 65535 ? g.nz : 65535);
        // cells the hot kernel cannot prove safe go to a worklist kept in the PCG
        // scratch field q (>= 4 B per cell, free during advection)
        int* worklist = reinterpret_cast<int*>(c->q);
        CU(cudaMemsetAsync(&c->sc->wl_count, 0, sizeof(unsigned), c->stream));
        LAUNCH(c, k_advect<T>, agrid, dim3(32, 8), g, c->media, P(c->u), P(c->v), P(c->w),
               P(c->u2), P(c->v2), P(c->w2), cc, worklist, c->sc);
        LAUNCH(c, k_advect_rest<T>, grid, 256, g, P(c->u), P(c->v), P(c->w), P(c->u2),
               P(c->v2), P(c->w2), (const T*)nullptr, (T*)nullptr, 1, cc, worklist, c->sc);
      }
    }
    // scalars not covered above: generic bounds-tested pass, traced through the OLD velocity
    for (size_t i = first_scalar; i < c->scal.size(); ++i) {
      LAUNCH(c, k_advect_scalar<T>, grid, 256, g, c->media, P(c->u), P(c->v), P(c->w),
             P(c->scal[i]), P(c->scal2), cc, c->sc);
      std::swap(c->scal[i], c->scal2);
    }
    CU(cudaGetLastError());
    std::swap(c->u, c->u2);
    std::swap(c->v, c->v2);
    std::swap(c->w, c->w2);
    return SPL_OK;
  }

  static int forces(spl_ctx* c, double dt) {
    const Geo& g = c->g;
    LAUNCH(c, k_forces<T>, grid_for(c, g.ncell), 256, g, c->media, P(c->u), P(c->v), P(c->w),
           T(c->d.gravity[0] * dt), T(c->d.gravity[1] * dt), T(c->d.gravity[2] * dt));
    CU(cudaGetLastError());
    return SPL_OK;
  }

  static int viscosity(spl_ctx* c, double dt) {
    const Geo& g = c->g;
    int rc = exchange_velocity(c, 1);
    if (rc) return rc;
    LAUNCH(c, k_viscosity<T>, grid_for(c, g.ncell), 256, g, c->media, P(c->u), P(c->v), P(c->w),
           P(c->u2), P(c->v2), P(c->w2), T(c->d.viscosity * dt));
    CU(cudaGetLastError());
    std::swap(c->u, c->u2);
    std::swap(c->v, c->v2);
    std::swap(c->w, c->w2);
    return SPL_OK;
  }

  static int cleanup(spl_ctx* c) {
    const Geo& g = c->g;
    const int grid = grid_for(c, g.ncell);
    int rc;
    LAUNCH(c, k_layer_init, grid, 256, g, c->media, c->layer);
    for (int step = 1; step <= 2; ++step) {   // Build.fs:34 max_distance = 2
      if ((rc = exchange_halo(c, reinterpret_cast<char*>(c->layer), 1, ncclUint8, 1))) return rc;
      LAUNCH(c, k_layer_step, grid, 256, g, c->media, c->layer, step);
    }
    if ((rc = exchange_halo(c, reinterpret_cast<char*>(c->layer), 1, ncclUint8, 1))) return rc;
    if ((rc = exchange_velocity(c, 1))) return rc;
    LAUNCH(c, k_cleanup<T>, grid, 256, g, c->media, c->layer, P(c->u), P(c->v), P(c->w),
           P(c->u2), P(c->v2), P(c->w2));
    CU(cudaGetLastError());
    std::swap(c->u, c->u2);
    std::swap(c->v, c->v2);
    std::swap(c->w, c->w2);
    return SPL_OK;
  }

  static int rhs(spl_ctx* c, T scale, char* out) {
    const Geo& g = c->g;
    int rc = exchange_velocity(c, 1);
    if (rc) return rc;
    if constexpr (std::is_same<T, float>::value) {
      if (g.nx % 4 == 0 && !c->no_quad) {
        LAUNCH(c, k_div_rhs4, grid_for(c, g.ncell / 4), 256, g, c->codes, P(c->u), P(c->v),
               P(c->w), P(out), scale);
        CU(cudaGetLastError());
        return SPL_OK;
      }
    }
    LAUNCH(c, k_div_rhs<T>, grid_for(c, g.ncell), 256, g, c->codes, P(c->u), P(c->v),
           P(c->w), P(out), scale);
    CU(cudaGetLastError());
    return SPL_OK;
  }

  // fp32 / VEC = 4 uses the cp.async ring kernel; every other
  // combination (fp64, ragged nx) the register-pipelined generic kernel.
  // zb0 / zstr / nbt: z-chunks zb0 + blockIdx.z * zstr of a phase of nbt CTAs in total (the
  // TMA kernel only; every other variant runs the whole grid in one launch)
  static bool phaseA_splits(const spl_ctx* c) {
    return std::is_same<T, float>::value && c->use_tma && !c->no_ring && !c->no_ws &&
           c->g.nx % 16 == 0 && c->vec == 4;
  }
  template <int VEC, int PC>
  static void launch_phaseA(spl_ctx* c, dim3 gridA, dim3 blockA, const T* zr, const T* sold,
                            T* snew, int par, int zc, int zb0 = 0, int zstr = 1,
                            unsigned nbt = 0, const Fused& fz = Fused{}) {
    const Geo& g = c->g;
    if (nbt == 0) nbt = gridA.x * gridA.y * gridA.z;
    if constexpr (std::is_same<T, float>::value && VEC == 4) {
      if (c->use_tma && !c->no_ring && !c->no_ws && g.nx % 16 == 0) {   // TMA producer warp
        const CUtensorMap* mz = tensor_map_for(c, g, zr, TM_F);
        const CUtensorMap* mo = tensor_map_for(c, g, sold, TM_F);
        const CUtensorMap* mn = tensor_map_for(c, g, snew, TM_F);
        const CUtensorMap* mc = tensor_map_for(c, g, c->codes, TM_C);
        if (mz && mo && mn && mc) {
          if (!(c->ring_attr_set & (256 << PC))) {
            cudaFuncSetAttribute(k_phaseA_tma<PC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(TmaSmem));
            c->ring_attr_set |= (256 << PC);
          }
          k_phaseA_tma<PC><<<gridA, dim3(32, TILE_Y + 1), sizeof(TmaSmem), c->stream>>>(
              *mz, *mo, *mn, *mc, g, c->codes, zr, sold, snew, P(c->q), c->sc, c->partials, par, zc,
              zb0, zstr, nbt, fz);
          c->launches++;
          return;
        }
      }
      if (!c->no_ring && !c->no_ws) {   // warp-specialised variant: a ninth warp owns the CTA edges
        if (!(c->ring_attr_set & (16 << PC))) {
          cudaFuncSetAttribute(k_phaseA_ws<PC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)sizeof(RingSmemWS));
          c->ring_attr_set |= (16 << PC);
        }
        k_phaseA_ws<PC><<<gridA, dim3(32, TILE_Y + 1), sizeof(RingSmemWS), c->stream>>>(
            g, c->codes, zr, sold, snew, P(c->q), c->sc, c->partials, par, zc);
        c->launches++;
        return;
      }
      if (!c->no_ring) {
        if (!(c->ring_attr_set & (1 << PC))) {   // per context => per device
          cudaFuncSetAttribute(k_phaseA_ring<PC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)sizeof(RingSmem));
          c->ring_attr_set |= (1 << PC);
        }
        k_phaseA_ring<PC><<<gridA, blockA, sizeof(RingSmem), c->stream>>>(
            g, c->codes, zr, sold, snew, P(c->q), c->sc, c->partials, par, zc);
        c->launches++;
        return;
      }
    }
    LAUNCH(c, (k_phaseA<T, VEC, PC, TILE_Y>), gridA, blockA, g, c->codes, zr, sold, snew,
           P(c->q), c->sc, c->partials, par, zc);
  }

  // 2-D box (nz = 1): the same kernel on the view nx x 1 x ny, so it marches along y
  template <int VEC, int PC>
  static void launch_phaseA_2d(spl_ctx* c, const T* zr, const T* sold, T* snew, int par) {
    Geo v = c->g;
    v.ny = 1;
    v.nz = c->g.ny;
    v.nzg = c->g.ny;
    v.plane = c->g.nx;
    const int tx = (v.nx + 32 * VEC - 1) / (32 * VEC);
    int gz = (2048 + tx - 1) / tx;
    int zc = (v.nz + gz - 1) / gz;
    if (zc < 16) zc = 16;
    if (zc > v.nz) zc = v.nz;
    while (zc < v.nz && (long long)tx * ((v.nz + zc - 1) / zc) > MAX_PARTIALS) zc *= 2;
    const dim3 grid(tx, 1, (v.nz + zc - 1) / zc);
    LAUNCH(c, (k_phaseA<T, VEC, PC, 1>), grid, dim3(32, 1), v, c->codes, zr, sold, snew,
           P(c->q), c->sc, c->partials, par, zc);
  }

  // ---- multigrid V-cycle z = M^-1 r, rho = r.z into slot `par` ------------------------
  // A level runs the fused marching kernels (spl_mg_march.cuh) when it qualifies,
  // otherwise the one-thread-per-cell kernels of spl_mg.cuh (same arithmetic).
  static bool mg_fast(const spl_ctx* c, const MgLevel& l) {
    return std::is_same<T, float>::value && !c->mg_no_march && l.g.nx % 4 == 0 &&
           l.g.nx >= 64 && l.g.ny >= TILE_Y && l.g.nz >= 8;
  }
  static int mg_zchunk(const MgLevel& l) {
    const int tiles = ((l.g.nx + 127) / 128) * ((l.g.ny + TILE_Y - 1) / TILE_Y);
    const int gz = (2048 + tiles - 1) / tiles;
    int zc = (l.g.nz + gz - 1) / gz;
    if (zc < 8) zc = 8;
    zc += zc & 1;                      // even: z pairs of the restriction stay inside a CTA
    return zc;
  }
  template <int MODE, int OUT, int DOT, int RUPD = 0>
  static void mg_march(spl_ctx* c, const MgLevel& l, const char* a, char* out,
                       const MgLevel* coarse, const char* ec, int par, double omega,
                       double omega_v = 0.0) {
    if constexpr (std::is_same<T, float>::value) {
      MgMarchArgs p{};
      p.g = l.g;
      p.codes = l.codes;
      p.a = reinterpret_cast<const float*>(a);
      p.b = reinterpret_cast<const float*>(l.b);
      p.out = reinterpret_cast<float*>(out);
      p.inv_scale = (float)(1.0 / l.scale);
      p.scale = (float)l.scale;
      p.omega = (float)omega;
      p.omega_v = (float)omega_v;
      p.sumform = c->mg_sumform ? 1 : 0;
      if (coarse) {
        p.ec = reinterpret_cast<const float*>(ec);
        p.codes_c = coarse->codes;
        p.bc = reinterpret_cast<float*>(coarse->b);
        p.cnx = coarse->g.nx;
        p.cny = coarse->g.ny;
        p.cplane = coarse->g.plane;
      }
      p.sc = c->sc;
      p.partials = c->partials;
      p.par = par;
      p.zchunk = mg_zchunk(l);
      const dim3 grid((l.g.nx + 127) / 128, (l.g.ny + TILE_Y - 1) / TILE_Y,
                      (l.g.nz + p.zchunk - 1) / p.zchunk);
      if (RUPD) {   // level 0: r -= alpha q folded into the pass (replaces k_phaseB)
        p.q = reinterpret_cast<const float*>(c->q);
        p.r_out = reinterpret_cast<float*>(c->einv);   // not in place: neighbours still read r
        p.nsbuf = c->nsbuf;
      }
      if (c->use_tma && !c->mg_no_tma && l.g.nx % 16 == 0 &&
          (MODE != MG_PROLONG || (coarse->g.nx % 4 == 0 && coarse->g.hz == 1))) {
        // TMA producer warp (spl_mg_tma.cuh): one tensor map per field of the level
        const CUtensorMap* ma = tensor_map_for(c, l.g, a, TM_F);
        const CUtensorMap* mb = tensor_map_for(c, l.g, RUPD ? c->q : l.b, TM_F);
        const CUtensorMap* mc = tensor_map_for(c, l.g, l.codes, TM_C);
        const CUtensorMap* me = (MODE == MG_PROLONG) ? tensor_map_for(c, coarse->g, ec, TM_E) : ma;
        if (ma && mb && mc && me) {
          MgTmaMaps tm;
          tm.a = *ma; tm.b = *mb; tm.code = *mc; tm.ec = *me;
          using SmemT = MgtSmem<MODE, OUT, RUPD>;
          const int tbit = RUPD ? (1 << 30) : (1 << (MODE * 4 + OUT * 2 + DOT));
          if (!(c->mgt_attr_set & tbit)) {
            cudaFuncSetAttribute(k_mg_tma<MODE, OUT, DOT, RUPD>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemT));
            c->mgt_attr_set |= tbit;
          }
          k_mg_tma<MODE, OUT, DOT, RUPD><<<grid, dim3(32, TILE_Y + 1), sizeof(SmemT), c->stream>>>(tm, p);
          c->launches++;
          return;
        }
      }
      using Smem = MgSmem<MODE, OUT, RUPD>;
      const int bit = RUPD ? (1 << 30) : (1 << (MODE * 4 + OUT * 2 + DOT));
      if (!(c->mg_attr_set & bit)) {
        cudaFuncSetAttribute(k_mg_march<MODE, OUT, DOT, RUPD>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
        c->mg_attr_set |= bit;
      }
      k_mg_march<MODE, OUT, DOT, RUPD><<<grid, dim3(32, TILE_Y), sizeof(Smem), c->stream>>>(p);
      c->launches++;
    }
  }

  // fuse_r: level 0's first pass also performs the PCG update r -= alpha q (mg_fuse_ok)
  static int vcycle(spl_ctx* c, int par, bool fuse_r = false) {
    const int L = (int)c->mg.size();
    const int nu = c->mg_nu;
    const std::vector<double>& om = c->mg_omega;   // per-sweep weights (size nu)
    std::vector<char*> cur(L, nullptr), nxt(L, nullptr);
    auto other = [&](int l, char* b) { return b == c->mg[l].u ? c->mg[l].tmp : c->mg[l].u; };
    // z-slabs: the neighbouring slab's boundary plane of every field a later pass reads
    // through the z stencil travels into the halo right after the pass that wrote it
    int xrc = SPL_OK;
    auto xch = [&](int l, char* buf) {
      if (c->d.world > 1 && xrc == SPL_OK)
        xrc = exchange_level(c, c->mg[l].g, buf, sizeof(T), nccl_real(c));
    };
    auto advance_buf = [&](int l) {
      cur[l] = nxt[l];
      nxt[l] = other(l, cur[l]);
      xch(l, cur[l]);
    };
    auto simple_smooth = [&](int l, int first, double omega) {
      MgLevel& v = c->mg[l];
      LAUNCH(c, k_mg_smooth<T>, grid_for(c, v.g.ncell), 256, v.g, v.codes, P(cur[l]), P(v.b),
             P(nxt[l]), T(omega), T(1.0 / v.scale), first, c->sc);
      advance_buf(l);
    };
    bool dot_done = false;
    // `n` sweeps from a zero initial guess, sweep s with weight om[s % nu];
    // `last` = this ends with the final launch of level 0
    auto smooth_from_zero = [&](int l, int n, bool last) {
      MgLevel& v = c->mg[l];
      if (mg_fast(c, v) && n >= 2) {
        const bool d0 = last && n == 2;
        if (d0) mg_march<MG_FIRST2, MG_SMOOTH, 1>(c, v, v.b, nxt[l], nullptr, nullptr, par, om[1 % nu], om[0]);
        else if (l == 0 && fuse_r) {
          // r' = r - alpha q goes to the spare field (SPL_PC_MG does not use einv); the two
          // buffers then trade places, so every later pass and the next iteration see r'
          mg_march<MG_FIRST2, MG_SMOOTH, 0, 1>(c, v, v.b, nxt[l], nullptr, nullptr, par, om[1 % nu], om[0]);
          std::swap(c->r, c->einv);
          v.b = c->r;
        }
        else mg_march<MG_FIRST2, MG_SMOOTH, 0>(c, v, v.b, nxt[l], nullptr, nullptr, par, om[1 % nu], om[0]);
        advance_buf(l);
        for (int s = 2; s < n; ++s) {
          const bool d = last && s == n - 1;
          if (d) mg_march<MG_PLAIN, MG_SMOOTH, 1>(c, v, cur[l], nxt[l], nullptr, nullptr, par, om[s % nu]);
          else mg_march<MG_PLAIN, MG_SMOOTH, 0>(c, v, cur[l], nxt[l], nullptr, nullptr, par, om[s % nu]);
          advance_buf(l);
        }
        if (last) dot_done = true;
      } else {
        cur[l] = other(l, nxt[l]);       // not read by the first sweep
        for (int s = 0; s < n; ++s) simple_smooth(l, s == 0, om[s % nu]);
      }
    };
    // first output buffer of every level, chosen so that the result of the
    // post-smoothing lands in u (= z on level 0)
    for (int l = 0; l < L; ++l) {
      MgLevel& v = c->mg[l];
      int k;
      if (l == L - 1) k = (L == 1) ? 2 * nu : c->mg_coarse;
      else k = 2 * nu;
      if (mg_fast(c, v) && ((l == L - 1) ? k >= 2 : nu >= 2)) k -= 1;   // FIRST2 = 2 sweeps
      nxt[l] = (k & 1) ? v.u : v.tmp;
    }
    xch(0, c->mg[0].b);   // r (with fuse_r: the OLD r, combined with q on the fly)
    if (fuse_r && c->d.world > 1 && xrc == SPL_OK)   // ... and q's boundary planes next to it
      xrc = exchange_level(c, c->mg[0].g, c->q, sizeof(T), nccl_real(c));
    // levels [lb, L) fit one CTA's shared memory: a single launch runs their sub-cycle
    int lb = L;
    if (L >= 2 && !c->mg_no_bottom)
      for (int l = 0; l < L; ++l)
        if (c->mg[l].g.ncell <= MG_BOTTOM_CELLS && L - l <= MG_BOTTOM_LEVELS) { lb = l; break; }
    if (c->mg_gbottom) lb = L;               // z-slabs: global coarsest solve instead
    const int ldown = lb < L ? lb : L - 1;   // levels [0, ldown) smooth + restrict
    for (int l = 0; l < ldown; ++l) {
      MgLevel& f = c->mg[l];
      MgLevel& k = c->mg[l + 1];
      smooth_from_zero(l, nu, false);
      if (mg_fast(c, f)) {
        mg_march<MG_PLAIN, MG_RESTRICT, 0>(c, f, cur[l], nullptr, &k, nullptr, par, 0.0);
      } else {
        LAUNCH(c, k_mg_restrict<T>, grid_for(c, k.g.ncell), 256, f.g, f.codes, P(cur[l]),
               P(f.b), T(f.scale), k.g, k.codes, P(k.b), c->sc);
      }
      xch(l + 1, k.b);
    }
    if (c->mg_gbottom) {
      // every rank gathers the coarsest right-hand side (z-slabs concatenate into the
      // global array), solves the global coarsest grid in one CTA, and takes its own
      // planes plus the two halo planes out of the global solution
      MgLevel& k = c->mg[L - 1];
      const Geo& gg = c->mg_gg;
      if (xrc == SPL_OK) {
        ncclResult_t r_ = g_nccl.AllGather(k.b, c->mg_gb, (size_t)k.g.ncell, nccl_real(c),
                                           c->comm, c->stream);
        if (r_ != ncclSuccess) xrc = fail(c, SPL_E_CUDA, "NCCL error %d in the multigrid gather", (int)r_);
      }
      MgBottomArgs<T> a{};
      a.nlev = 1;
      a.nu = nu;
      a.coarse_sweeps = c->mg_gsweeps;
      a.g[0] = gg;
      a.codes[0] = c->mg_gcodes;
      a.scale[0] = T(k.scale);
      for (int i = 0; i < nu; ++i) a.omega[i] = T(om[i]);
      a.b0 = P(c->mg_gb);
      a.u0 = P(c->mg_gu);
      a.sc = c->sc;
      const size_t smem = mg_bottom_smem<T>(a.g, 1);
      if (smem > c->mg_bottom_smem) {
        cudaFuncSetAttribute(k_mg_bottom<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem);
        c->mg_bottom_smem = smem;
      }
      k_mg_bottom<T><<<1, 1024, smem, c->stream>>>(a);
      c->launches++;
      const int zlo = k.g.z0 > 0 ? k.g.z0 - 1 : 0;
      const int zhi = (k.g.z0 + k.g.nz < gg.nz) ? k.g.z0 + k.g.nz + 1 : gg.nz;
      const size_t pb = (size_t)gg.plane * sizeof(T);
      if (cudaMemcpyAsync(k.u + ((long long)zlo - k.g.z0) * (long long)pb,
                          c->mg_gu + (size_t)zlo * pb, (size_t)(zhi - zlo) * pb,
                          cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess && xrc == SPL_OK)
        xrc = fail(c, SPL_E_CUDA, "copy of the global coarsest solution failed");
      cur[L - 1] = k.u;
    } else if (lb < L) {
      MgBottomArgs<T> a{};
      a.nlev = L - lb;
      a.nu = nu;
      a.coarse_sweeps = c->mg_coarse;
      for (int l = lb; l < L; ++l) {
        a.g[l - lb] = c->mg[l].g;
        a.codes[l - lb] = c->mg[l].codes;
        a.scale[l - lb] = T(c->mg[l].scale);
      }
      for (int k = 0; k < nu; ++k) a.omega[k] = T(om[k]);
      a.b0 = P(c->mg[lb].b);
      a.u0 = P(c->mg[lb].u);
      a.sc = c->sc;
      const size_t smem = mg_bottom_smem<T>(a.g, a.nlev);
      if (smem > c->mg_bottom_smem) {
        cudaFuncSetAttribute(k_mg_bottom<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem);
        c->mg_bottom_smem = smem;
      }
      k_mg_bottom<T><<<1, 1024, smem, c->stream>>>(a);
      c->launches++;
      cur[lb] = c->mg[lb].u;
      xch(lb, cur[lb]);
    } else {
      smooth_from_zero(L - 1, (L == 1) ? 2 * nu : c->mg_coarse, L == 1);
    }
    for (int l = ldown - 1; l >= 0; --l) {      // post-smoothing: the weights in reverse order
      MgLevel& f = c->mg[l];
      MgLevel& k = c->mg[l + 1];
      if (mg_fast(c, f)) {
        for (int s = 0; s < nu; ++s) {
          const bool d = (l == 0) && s == nu - 1;
          const double w = om[nu - 1 - s];
          if (s == 0) {
            if (d) mg_march<MG_PROLONG, MG_SMOOTH, 1>(c, f, cur[l], nxt[l], &k, cur[l + 1], par, w);
            else mg_march<MG_PROLONG, MG_SMOOTH, 0>(c, f, cur[l], nxt[l], &k, cur[l + 1], par, w);
          } else {
            if (d) mg_march<MG_PLAIN, MG_SMOOTH, 1>(c, f, cur[l], nxt[l], nullptr, nullptr, par, w);
            else mg_march<MG_PLAIN, MG_SMOOTH, 0>(c, f, cur[l], nxt[l], nullptr, nullptr, par, w);
          }
          advance_buf(l);
        }
        if (l == 0) dot_done = true;
      } else {
        LAUNCH(c, k_mg_prolong<T>, grid_for(c, f.g.ncell), 256, f.g, f.codes, P(cur[l]), k.g,
               P(cur[l + 1]), c->sc);
        xch(l, cur[l]);
        for (int s = 0; s < nu; ++s) simple_smooth(l, 0, om[nu - 1 - s]);
      }
    }
    if (!dot_done)
      LAUNCH(c, k_dot_rz<T>, grid_for(c, c->g.ncell), 256, c->g, P(c->r), P(c->z), c->sc,
             c->partials, par);
    return xrc;
  }

  // the residual update can ride on the V-cycle's first pass when that pass is the fused
  // marching kernel on level 0 (fp32, at least two levels and two sweeps).  On z-slabs the pass
  // forms r' = r - alpha q for the neighbours' boundary planes too, from the halos of r and q:
  // one more plane exchange (q) instead of a whole k_phaseB pass (SPL_MG_NO_FUSE_SLABS=1: off).
  static bool mg_fuse_ok(const spl_ctx* c) {
    return c->d.precond == SPL_PC_MG && std::is_same<T, float>::value && !c->mg_no_fuse &&
           (c->d.world == 1 || !c->mg_no_fuse_slabs) && c->mg.size() >= 2 && c->mg_nu >= 2 &&
           mg_fast(c, c->mg[0]) &&
           c->profile_iters == 0;
  }
  static int precondition(spl_ctx* c, int par, bool fuse_r = false) {
    if (c->d.precond == SPL_PC_MG) return vcycle(c, par, fuse_r);
    const Geo& g = c->g;
    if constexpr (std::is_same<T, float>::value) {
      if (g.nx % 4 == 0 && !c->no_quad && !c->no_rb_march) {   // z-marching tiles, each array once
        const int tx = (g.nx + 127) / 128, ty = (g.ny + RBM_TY - 1) / RBM_TY;
        const int zc = phaseA_zchunk(c, tx * ty);
        const dim3 grid(tx, ty, (g.nz + zc - 1) / zc), block(32, RBM_TY);
        k_rb_march<0><<<grid, block, 0, c->stream>>>(g, c->codes, P(c->r), P(c->einv), P(c->z),
                                                    c->sc, c->partials, par, zc);
        k_rb_march<1><<<grid, block, 0, c->stream>>>(g, c->codes, P(c->r), P(c->einv), P(c->z),
                                                    c->sc, c->partials, par, zc);
        c->launches += 2;
        return SPL_OK;
      }
      if (g.nx % 4 == 0 && !c->no_quad) {     // quad kernels: same z, 30 B / cell
        const int gq = grid_for(c, g.ncell / 4);
        LAUNCH(c, k_rb_black4, gq, 256, g, c->codes, P(c->r), P(c->einv), P(c->z), c->sc);
        LAUNCH(c, k_rb_red4, gq, 256, g, c->codes, P(c->r), P(c->einv), P(c->z), c->sc,
               c->partials, par);
        return SPL_OK;
      }
    }
    const int grid = grid_for(c, g.ncell);
    LAUNCH(c, (k_rb_sweep<T, 1>), grid, 256, g, c->codes, P(c->r), P(c->einv), P(c->z), c->sc,
           c->partials, par);
    LAUNCH(c, (k_rb_sweep<T, 0>), grid, 256, g, c->codes, P(c->r), P(c->einv), P(c->z), c->sc,
           c->partials, par);
    return SPL_OK;
  }

  static int build_rbic0(spl_ctx* c) {
    const Geo& g = c->g;
    const int grid = grid_for(c, g.ncell);
    LAUNCH(c, k_rb_contrib<T>, grid, 256, g, c->codes, c->media, P(c->z), c->d.tau);
    LAUNCH(c, k_rb_einv<T>, grid, 256, g, c->codes, P(c->z), P(c->einv), 0.25);
    CU(cudaGetLastError());
    return SPL_OK;
  }

  template <int VEC, int PC>
  static int pcg(spl_ctx* c) {
    const Geo& g = c->g;
    const size_t fb = field_bytes(c, sizeof(T));
    // x = 0, s_0 = 0 (with halos), q = 0.  Only direction buffer 0 is read before
    // it is written (as s_0 with beta = 0); the halo planes of every buffer are
    // zero from allocation and only ever receive a neighbour slab's planes.
    const int m = c->nsbuf;
    if (!c->pcg_keep_x)
      CU(cudaMemsetAsync(c->alloc[6], 0, field_bytes(c, 8), c->stream));   // p (fp64)
    CU(cudaMemsetAsync(c->sbuf[0] - (size_t)g.hz * g.plane * sizeof(T), 0, fb, c->stream));
    CU(cudaMemsetAsync(c->alloc[10], 0, fb, c->stream));  // q
    SBufs bufs;
    for (int i = 0; i < MAX_SBUF; ++i) bufs.p[i] = c->sbuf[i < m ? i : 0];
    // scalars
    Scal* h = c->hsc;
    const unsigned long long keep_nf = 0;
    memset(h, 0, sizeof(Scal));
    h->world = c->d.world;
    h->rank = c->d.rank;
    h->tol = c->fixed_iters > 0 ? 0.0 : (c->pcg_tol_first > 0.0 ? c->pcg_tol_first : c->d.tol);
    h->nonfinite = keep_nf;
    for (int w = 0; w < c->d.world; ++w)
      h->gBM[1][w][0] = (w == 0) ? std::numeric_limits<double>::infinity() : 0.0;
    static_assert(sizeof(Scal) % 4 == 0, "Scal is copied in 32-bit words");
    if (int rc_ = from_host(c, c->sc, c->hsc_d, sizeof(Scal))) return rc_;
    CU(cudaStreamSynchronize(c->stream));   // h is reused below

    const int gridB = grid_for(c, g.ncell / VEC);
    const int tx = (g.nx + 32 * VEC - 1) / (32 * VEC);
    const int ty = (g.ny + TILE_Y - 1) / TILE_Y;
    int zc = phaseA_zchunk(c, tx * ty);
    // z-slabs over peer memory: at least four z-chunks, so that the two boundary chunks (which
    // wait for the neighbours' planes) are the second half of the launch at most
    if (c->d.world > 1 && c->fused_ok && !pc_explicit_z(PC) && phaseA_splits(c) &&
        c->zchunk_override <= 0 && (g.nz + zc - 1) / zc < 4) {
      zc = g.nz / 4;
      if (zc < 8) zc = 8;
    }
    const dim3 gridA(tx, ty, (g.nz + zc - 1) / zc);
    const dim3 blockA(32, TILE_Y);
    const T* zr = pc_explicit_z(PC) ? P(c->z) : P(c->r);

    // rho_0 = r.z, max|b|  (phase B with alpha = 0)
    LAUNCH(c, (k_phaseB<T, VEC, PC>), gridB, 256, g, c->codes, P(c->r), P(c->q), P(c->z),
           c->sc, c->partials, 0, 1, m, Fused{});
    int rc = SPL_OK;
    if (pc_explicit_z(PC) && (rc = precondition(c, 0))) return rc;
    if ((rc = gather_slots(c, &c->sc->gBM[0][0][0], 2))) return rc;
    LAUNCH(c, k_set_bmax, 1, 1, c->sc);
    if (c->pcg_keep_x) LAUNCH(c, k_force_bmax, 1, 1, c->sc, c->pcg_bmax_keep);   // refinement pass
    CU(cudaGetLastError());

    const int limit = c->fixed_iters > 0 ? c->fixed_iters : c->d.max_iters;
    int it = 0;
    int nflush = 0;
    int batch = 0;
    bool done = false;
    const int poll = (PC == PC_MG) ? 1 : 16;   // an MG iteration is ~70 launches
    Fused fz{};
    for (int w = 0; w < c->d.world && w < MAXW; ++w) fz.box[w] = c->fboxes[w];
    if (const char* e = getenv("SPL_FUSED_MODE")) fz.mode = atoi(e);
    const bool send_on_main = (fz.mode & 2) != 0;
    bool fused_prev = false;       // the previous iteration's phase B pushed its slots to the peers
    unsigned long long last_stamp = 0;
    while (it < limit && !done) {
      const int stop = (it + poll < limit) ? it + poll : limit;
      for (; it < stop;) {
        ++it;
        const int par = it & 1;
        char* sold = c->sbuf[(it - 1) % m];   // s_{it-1}
        char* snew = c->sbuf[it % m];         // s_it
        const bool two_d = g.nz == 1 && g.nzg == 1 && g.ny > 1 && !c->no_2d_view;
        const bool prof = c->profile_iters > 0 && it <= c->profile_iters;
        // z-slabs, three ways to get the boundary planes of s' to the neighbours:
        //  fused  the iteration runs over peer memory: k_send_boundary stores the planes into the
        //         neighbours' halos on the side stream, phase A is ONE launch whose boundary
        //         chunks come last and wait for the neighbour's stamp, the CG scalars travel
        //         as peer stores from the last CTA of each phase -- no NCCL call, no extra launch
        //  split  NCCL send / recv on the side stream while the interior chunks of phase A run
        //         (SPL_SPLIT=1; loses to the plain order at 512^3: the boundary launch is a
        //         1.15-wave tail)
        //  plain  boundary planes, send / recv, phase A, all on one stream
        const bool tma_ok = c->d.world > 1 && c->cstream && !two_d && phaseA_splits(c);
        const bool fused = tma_ok && c->fused_ok && !pc_explicit_z(PC) && gridA.z >= 3;
        const bool split = tma_ok && !fused && c->split && gridA.z >= 3;
        const bool fhalo = fused && !(fz.mode & 16);   // 16: scalars over peer memory, halo over NCCL
        fz.on = fused ? 1 : 0;
        fz.halo = fhalo ? 1 : 0;
        if (fused) {
          fz.stamp = ++c->iter_seq;
          fz.wait_b = fused_prev ? fz.stamp - 1 : 0;
          fused_prev = true;
          last_stamp = fz.stamp;
        }
        if (fhalo) {
          if constexpr (std::is_same<T, float>::value && !pc_explicit_z(PC)) {
            if (!send_on_main) {
              CU(cudaEventRecord(c->ev_beta, c->stream));        // r of iteration it-1 is final
              CU(cudaStreamWaitEvent(c->cstream, c->ev_beta, 0));
            }
            const int bi = it % m;
            const size_t pf = (size_t)g.plane * sizeof(float);
            float* lo = c->d.rank > 0 ? reinterpret_cast<float*>(
                c->nb_sbuf[0][bi] + (size_t)(g.hz + c->slab_geo[2 * (c->d.rank - 1)]) * pf) : nullptr;
            float* hi = c->d.rank < c->d.world - 1 ? reinterpret_cast<float*>(
                c->nb_sbuf[1][bi] + (size_t)(g.hz - 1) * pf) : nullptr;
            const int sgrid = grid_for(c, 2 * g.plane / 4) < c->sms ? grid_for(c, 2 * g.plane / 4) : c->sms;
            k_send_boundary<PC><<<sgrid, 256, 0,
                                  send_on_main ? c->stream : c->cstream>>>(
                g, c->codes, zr, P(sold), lo, hi, c->sc, par, fz, c->hticket);
            c->launches++;
            if (!send_on_main) CU(cudaEventRecord(c->ev_halo, c->cstream));   // it has read r: phase B may run
          }
        } else if (c->d.world > 1) {
          cudaStream_t hs = split ? c->cstream : c->stream;
          if (split) {
            CU(cudaEventRecord(c->ev_beta, c->stream));          // beta of this iteration is complete
            CU(cudaStreamWaitEvent(c->cstream, c->ev_beta, 0));
          }
          k_dir_boundary<T, PC><<<grid_for(c, 2 * g.plane), 256, 0, hs>>>(
              g, c->codes, zr, P(sold), P(snew), c->sc, par);
          c->launches++;
          if ((rc = exchange_halo(c, snew, sizeof(T), nccl_real(c), 1, hs))) return rc;
          if (split) CU(cudaEventRecord(c->ev_halo, c->cstream));
        }
        if (prof) CU(cudaEventRecord(c->pev[3 * (it - 1)], c->stream));
        if (two_d) {
          launch_phaseA_2d<VEC, PC>(c, zr, P(sold), P(snew), par);
        } else if (split) {
          const unsigned nbt = gridA.x * gridA.y * gridA.z;
          launch_phaseA<VEC, PC>(c, dim3(gridA.x, gridA.y, gridA.z - 2), blockA, zr, P(sold),
                                 P(snew), par, zc, 1, 1, nbt);
          CU(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
          launch_phaseA<VEC, PC>(c, dim3(gridA.x, gridA.y, 2), blockA, zr, P(sold), P(snew), par,
                                 zc, 0, (int)gridA.z - 1, nbt);
        } else {
          launch_phaseA<VEC, PC>(c, gridA, blockA, zr, P(sold), P(snew), par, zc, 0, 1, 0, fz);
          if (fhalo && !send_on_main) CU(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
        }
        if (prof) CU(cudaEventRecord(c->pev[3 * (it - 1) + 1], c->stream));
        if (!fz.on && (rc = gather_slots(c, c->sc->gA, 1))) return rc;
        if (prof) CU(cudaEventRecord(c->pev[3 * (it - 1) + 2], c->stream));
        const bool fuse_r = (PC == PC_MG) && mg_fuse_ok(c);
        if (!fuse_r)
          LAUNCH(c, (k_phaseB<T, VEC, PC>), gridB, 256, g, c->codes, P(c->r), P(c->q), P(c->z),
                 c->sc, c->partials, par, 0, m, fz);
        if (prof) CU(cudaEventRecord(c->pev[3 * c->profile_iters + (it - 1)], c->stream));
        if (pc_explicit_z(PC) && (rc = precondition(c, par, fuse_r))) return rc;   // z = M^-1 r, rho = r.z
        if (!fz.on && (rc = gather_slots(c, &c->sc->gBM[par][0][0], 2))) return rc;
        if (it % m == 0) {   // every buffer holds an unflushed direction: x += sum alpha s
          const bool pf = c->profile_iters > 0 && nflush < 64;
          if (pf) CU(cudaEventRecord(c->xev[2 * nflush], c->stream));
          LAUNCH(c, (k_xflush<T, VEC>), gridB, 256, g, reinterpret_cast<double*>(c->p), bufs,
                 c->sc, m);
          if (pf) CU(cudaEventRecord(c->xev[2 * nflush + 1], c->stream));
          ++nflush;
        }
      }
      CU(cudaGetLastError());
      if (c->fixed_iters > 0) continue;
      // lagged poll: the flag of batch b is read while batch b+1 is already queued, so
      // the GPU never waits for the host (a converged solve only runs no-op launches)
      const int slot = batch & 1;
      if (int rc_ = to_host(c, &c->hdone_d[slot], &c->sc->done, sizeof(int))) return rc_;
      CU(cudaEventRecord(c->dev[slot], c->stream));
      if (batch >= 1) {
        CU(cudaEventSynchronize(c->dev[slot ^ 1]));
        done = c->hdone[slot ^ 1] != 0;
      }
      ++batch;
    }
    if (fused_prev) {   // slots of the last phase B (k_finalize reads them)
      fz.on = 1;
      fz.wait_b = last_stamp;
      LAUNCH(c, k_fused_wait_B, 1, 32, c->sc, it & 1, fz);
    }
    // terms of the last (K mod m) iterations are still pending
    c->profile_flushes = nflush < 64 ? nflush : 64;
    LAUNCH(c, (k_xflush<T, VEC>), gridB, 256, g, reinterpret_cast<double*>(c->p), bufs, c->sc,
           m);
    LAUNCH(c, k_finalize, 1, 1, c->sc, it & 1);
    CU(cudaGetLastError());
    return SPL_OK;
  }

  static int pcg_dispatch(spl_ctx* c) {
    const int pc = c->d.precond;
    if (c->vec == 4) {
      if (pc == SPL_PC_NONE) return pcg<4, PC_NONE>(c);
      if (pc == SPL_PC_JACOBI) return pcg<4, PC_JACOBI>(c);
      if (pc == SPL_PC_RBIC0) return pcg<4, PC_RBIC0>(c);
      if (pc == SPL_PC_MG) return pcg<4, PC_MG>(c);
    } else {
      if (pc == SPL_PC_NONE) return pcg<1, PC_NONE>(c);
      if (pc == SPL_PC_JACOBI) return pcg<1, PC_JACOBI>(c);
      if (pc == SPL_PC_RBIC0) return pcg<1, PC_RBIC0>(c);
      if (pc == SPL_PC_MG) return pcg<1, PC_MG>(c);
    }
    return fail(c, SPL_E_BADARG, "preconditioner %d is not available", pc);
  }

  static int grad(spl_ctx* c, double dt) {
    const Geo& g = c->g;
    int rc = exchange_halo(c, c->p, 8, ncclDouble, 1);
    if (rc) return rc;
    const double kappa = dt / (c->d.h * c->d.rho);
    if constexpr (std::is_same<T, float>::value) {
      if (g.nx % 4 == 0 && !c->no_quad) {
        LAUNCH(c, k_grad_sub4, grid_for(c, g.ncell / 4), 256, g, c->media,
               reinterpret_cast<const double*>(c->p), P(c->u), P(c->v), P(c->w), kappa);
        CU(cudaGetLastError());
        return SPL_OK;
      }
    }
    LAUNCH(c, k_grad_sub<T>, grid_for(c, g.ncell), 256, g, c->media,
           reinterpret_cast<const double*>(c->p), P(c->u),
           P(c->v), P(c->w), kappa);
    CU(cudaGetLastError());
    return SPL_OK;
  }

  // Pressure.apply + the checks as one marching pass (k_grad_check4): fp32 quads only
  static bool grad_check_ok(const spl_ctx* c) {
    return std::is_same<T, float>::value && c->g.nx % 4 == 0 && !c->no_quad && !c->no_gc_fuse;
  }
  static int grad_check(spl_ctx* c, double dt) {
    if constexpr (std::is_same<T, float>::value) {
      const Geo& g = c->g;
      // halos: p of the neighbouring slabs' boundary planes, and the velocity below the slab
      // (the incoming z face of plane 0 is recomputed from the OLD w of plane -1)
      int rc = exchange_halo(c, c->p, 8, ncclDouble, 1);
      if (rc) return rc;
      if ((rc = exchange_velocity(c, 1))) return rc;
      const double kappa = dt / (c->d.h * c->d.rho);
      const int tx = (g.nx + 127) / 128, ty = (g.ny + GC_TY - 1) / GC_TY;
      const int zc = phaseA_zchunk(c, tx * ty);
      const dim3 grid(tx, ty, (g.nz + zc - 1) / zc);
      k_grad_check4<<<grid, dim3(32, GC_TY), 0, c->stream>>>(
          g, c->media, c->codes, reinterpret_cast<const double*>(c->p), P(c->u), P(c->v), P(c->w),
          P(c->u2), P(c->v2), P(c->w2), kappa, c->sc, c->partials, zc);
      c->launches++;
      CU(cudaGetLastError());
      std::swap(c->u, c->u2);
      std::swap(c->v, c->v2);
      std::swap(c->w, c->w2);
      return gather_slots(c, c->sc->gD, 1);
    }
    return SPL_E_BADARG;
  }

  static int check(spl_ctx* c) {
    const Geo& g = c->g;
    int rc = exchange_velocity(c, 1);
    if (rc) return rc;
    bool quad = false;
    if constexpr (std::is_same<T, float>::value) quad = g.nx % 4 == 0 && !c->no_quad;
    if constexpr (std::is_same<T, float>::value) {
      if (quad)
        LAUNCH(c, k_check4, grid_for(c, g.ncell / 4), 256, g, c->codes, P(c->u), P(c->v),
               P(c->w), reinterpret_cast<const double*>(c->p), c->sc, c->partials);
    }
    if (!quad)
      LAUNCH(c, k_check<T>, grid_for(c, g.ncell), 256, g, c->codes, P(c->u), P(c->v), P(c->w),
             reinterpret_cast<const double*>(c->p), c->sc, c->partials);
    CU(cudaGetLastError());
    return gather_slots(c, c->sc->gD, 1);
  }
};

// entry points that need the uploaded state: completes a pending spl_upload_async first
#define NEED_UPLOAD(c, who)                                                        \
  do {                                                                             \
    if ((c)->upload_pending) {                                                     \
      const int rc_ = finish_upload(c);                                            \
      if (rc_) return rc_;                                                         \
    }                                                                              \
    if (!(c)->uploaded) return fail(c, SPL_E_STATE, who " before spl_upload");     \
    /* a queued spl_download_async still reads the fields: later work waits for it */ \
    if ((c)->download_queued) cudaStreamWaitEvent((c)->stream, (c)->ev_down, 0);   \
  } while (0)

#define DISPATCH(c, expr_f32, expr_f64) ((c)->d.dtype == SPL_F64 ? (expr_f64) : (expr_f32))

int do_rhs(spl_ctx* c, double dt) {
  const double scale = (c->d.h * c->d.rho) / dt;   // Pressure.fs:64
  int rc = DISPATCH(c, Impl<float>::rhs(c, (float)scale, c->r),
                    Impl<double>::rhs(c, scale, c->r));
  if (rc) return rc;
  c->have_rhs = true;
  c->rhs_dt = dt;
  return SPL_OK;
}

float ev_ms(spl_ctx* c, int a, int b) {
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]);
  return ms;
}

// fill info from the device scalars after a solve / check
int read_back(spl_ctx* c, spl_info* info, bool solve, bool chk) {
  int rc = pull_scal(c);
  if (rc) return rc;
  const Scal* h = c->hsc;
  if (!info) return SPL_OK;
  if (solve) {
    info->iters = h->iters + c->iters_base;
    info->converged = h->converged;
    info->r_inf = h->rmax;
    info->b_inf = h->bmax;
  }
  if (chk) {
    double m = 0.0;
    for (int w = 0; w < c->d.world; ++w) m = fmax(m, h->gD[w]);
    info->max_abs_div = m;
    info->nonfinite = (int64_t)h->nonfinite;
  }
  info->trace_escapes = (int64_t)h->escapes;
  info->kernel_launches = c->launches;
  return SPL_OK;
}

}  // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

int spl_abi_version(void) { return SPL_ABI_VERSION; }

int spl_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

void spl_desc_default(spl_desc* d) {
  if (!d) return;
  memset(d, 0, sizeof *d);
  d->nx = d->ny = d->nz = 1;
  d->h = 10.0;                 // Constants.fs:11
  d->rho = 999.97;             // Constants.fs:15
  d->dtype = SPL_F32;
  d->precond = SPL_PC_JACOBI;
  d->tol = 1e-6;
  d->max_iters = 20000;
  d->tau = 0.0;
  d->quirks = 0;
  d->device = 0;
  d->rank = 0;
  d->world = 1;
  d->nz_global = 0;
  d->z0 = 0;
  d->nccl_id = nullptr;
  d->gravity[0] = 0.0;         // Constants.fs:26
  d->gravity[1] = -9.81;
  d->gravity[2] = 0.0;
  d->viscosity = 0.000001;     // Constants.fs:13
}

const char* spl_last_error(const spl_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

int spl_nccl_unique_id(void* out128) {
  spl_ctx* c = nullptr;
  std::string err;
  if (!out128) return fail(c, SPL_E_BADARG, "out128 is NULL");
  if (!load_nccl(err)) return fail(c, SPL_E_CUDA, "%s", err.c_str());
  ncclUniqueId id;
  NC(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(out128, &id, 128);
  return SPL_OK;
}

void spl_destroy(spl_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->d.device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->cstream) cudaStreamSynchronize(c->cstream);
  for (int w = 0; w < MAXW; ++w)
    if (c->peer_maps[w]) cudaIpcCloseMemHandle(c->peer_maps[w]);
  for (int side = 0; side < 2; ++side)
    for (int i = 0; i < MAX_SBUF; ++i)
      if (c->nb_maps[side][i]) cudaIpcCloseMemHandle(c->nb_maps[side][i]);
  for (int side = 0; side < 2; ++side)
    if (c->nb_mbox[side]) cudaIpcCloseMemHandle(c->nb_mbox[side]);
  if (c->mbox) cudaFree(c->mbox);
  if (c->xticket) cudaFree(c->xticket);
  if (c->pbox) cudaFree(c->pbox);
  if (c->hticket) cudaFree(c->hticket);
  if (c->h2d) { cudaStreamSynchronize(c->h2d); cudaStreamDestroy(c->h2d); }
  if (c->d2h) { cudaStreamSynchronize(c->d2h); cudaStreamDestroy(c->d2h); }
  for (cudaEvent_t e : {c->ev_up, c->ev_down, c->ev_done})
    if (e) cudaEventDestroy(e);
  if (c->cstream) cudaStreamDestroy(c->cstream);
  if (c->ev_beta) cudaEventDestroy(c->ev_beta);
  if (c->ev_halo) cudaEventDestroy(c->ev_halo);
  if (c->comm) g_nccl.CommDestroy(c->comm);
  for (void* p : c->alloc)
    if (p) cudaFree(p);
  for (MgLevel& l : c->mg)
    for (void* p : l.allocs)
      if (p) cudaFree(p);
  if (c->mg_gcodes) cudaFree(c->mg_gcodes);
  if (c->mg_gb) cudaFree(c->mg_gb);
  if (c->mg_gu) cudaFree(c->mg_gu);
  if (c->mloc) cudaFree(c->mloc);
  if (c->mold) cudaFree(c->mold);
  if (c->mnew) cudaFree(c->mnew);
  if (c->mark_base) cudaFree(c->mark_base);
  if (c->mstats) cudaFree(c->mstats);
  if (c->sc) cudaFree(c->sc);
  if (c->partials) cudaFree(c->partials);
  if (c->flag) cudaFree(c->flag);
  if (c->hsc) cudaFreeHost(c->hsc);
  if (c->hflag) cudaFreeHost(c->hflag);
  if (c->hdone) cudaFreeHost(c->hdone);
  if (c->hesc) cudaFreeHost(c->hesc);
  if (c->hagree) cudaFreeHost(c->hagree);
  for (cudaEvent_t e : c->dev)
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : c->ev)
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : c->tev)
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : c->xev)
    if (e) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

// z-slabs over peer memory: the two neighbours' direction buffers are mapped through CUDA IPC
// so that k_send_boundary can store this rank's boundary planes straight into their halos.
// Every rank exchanges the handles of its nsbuf buffers; the path is used only if every rank
// mapped what it needs (otherwise: ncclSend / ncclRecv as before).
static int map_neighbour_sbufs(spl_ctx* c) {
  const spl_desc& d = c->d;
  const int m = c->nsbuf;
  const char* fe = getenv("SPL_FUSED");
  bool ok = c->p2p && !(fe && atoi(fe) == 0) && d.dtype == SPL_F32;
  std::vector<unsigned char> hh((size_t)64 * m * d.world, 0);
  unsigned char* dh = nullptr;
  CU(cudaMalloc(&dh, hh.size()));
  for (int i = 0; i < m && ok; ++i) {
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, c->alloc[32 + i]) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
    memcpy(&hh[((size_t)d.rank * m + i) * 64], &h, 64);
  }
  CU(cudaMemcpyAsync(dh + (size_t)d.rank * m * 64, &hh[(size_t)d.rank * m * 64], (size_t)64 * m,
                     cudaMemcpyHostToDevice, c->stream));
  NC(g_nccl.AllGather(dh + (size_t)d.rank * m * 64, dh, (size_t)64 * m, ncclUint8, c->comm,
                      c->stream));
  CU(cudaMemcpyAsync(hh.data(), dh, hh.size(), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  for (int side = 0; side < 2 && ok; ++side) {
    const int nb = side == 0 ? d.rank - 1 : d.rank + 1;
    if (nb < 0 || nb >= d.world) continue;
    for (int i = 0; i < m; ++i) {
      cudaIpcMemHandle_t h;
      memcpy(&h, &hh[((size_t)nb * m + i) * 64], 64);
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = false;
        break;
      }
      c->nb_maps[side][i] = ptr;
      c->nb_sbuf[side][i] = static_cast<char*>(ptr);
    }
  }
  unsigned char mine = ok, *dm = dh;     // agree: all ranks or none
  CU(cudaMemcpyAsync(dm + d.rank, &mine, 1, cudaMemcpyHostToDevice, c->stream));
  NC(g_nccl.AllGather(dm + d.rank, dm, 1, ncclUint8, c->comm, c->stream));
  CU(cudaMemcpyAsync(hh.data(), dm, d.world, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  for (int w = 0; w < d.world; ++w) ok = ok && hh[w];
  c->fused_ok = ok;
  CU(cudaFree(dh));
  return SPL_OK;
}

// the mailboxes of k_mail_exchange: allocate, exchange the IPC handles through NCCL, map the two
// neighbours'; used only if every rank mapped what it needs (SPL_MAILBOX=0: ncclSend / ncclRecv)
static int map_mailboxes(spl_ctx* c) {
  const spl_desc& d = c->d;
  // Opt-in (SPL_MAILBOX=1): measured on 2 B200s with the multigrid step at 512^3 per GPU
  // (28 plane exchanges per V-cycle) it is no faster than NCCL's grouped send / recv -- 42.3 ms
  // against 41.5 ms per 14 iterations, profiles/README.md round 2 -- so NCCL stays the default.
  const char* me = getenv("SPL_MAILBOX");
  bool ok = c->p2p && me && atoi(me) != 0;
  // a slot holds the widest exchange: hz planes of a velocity component or one fp64 plane
  size_t slot = (size_t)c->g.plane * (size_t)std::max<size_t>((size_t)c->g.hz * c->esz, 8);
  slot = (slot + 255) / 256 * 256;
  c->mbox_slot = slot;
  const size_t total = sizeof(MailHdr) + 4 * slot;
  if (ok && cudaMalloc(&c->mbox, total) != cudaSuccess) { cudaGetLastError(); c->mbox = nullptr; ok = false; }
  if (ok) CU(cudaMemsetAsync(c->mbox, 0, sizeof(MailHdr), c->stream));
  CU(cudaMalloc(&c->xticket, sizeof(unsigned)));
  CU(cudaMemsetAsync(c->xticket, 0, sizeof(unsigned), c->stream));
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof mine);
  if (ok && cudaIpcGetMemHandle(&mine, c->mbox) != cudaSuccess) { cudaGetLastError(); ok = false; }
  std::vector<unsigned char> hh((size_t)65 * d.world, 0);
  unsigned char* dh = nullptr;
  CU(cudaMalloc(&dh, hh.size()));
  memcpy(&hh[(size_t)65 * d.rank], &mine, 64);
  hh[(size_t)65 * d.rank + 64] = ok;
  CU(cudaMemcpyAsync(dh + (size_t)65 * d.rank, &hh[(size_t)65 * d.rank], 65, cudaMemcpyHostToDevice,
                     c->stream));
  NC(g_nccl.AllGather(dh + (size_t)65 * d.rank, dh, 65, ncclUint8, c->comm, c->stream));
  CU(cudaMemcpyAsync(hh.data(), dh, hh.size(), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  for (int w = 0; w < d.world; ++w) ok = ok && hh[(size_t)65 * w + 64];
  for (int side = 0; side < 2 && ok; ++side) {
    const int nb = side == 0 ? d.rank - 1 : d.rank + 1;
    if (nb < 0 || nb >= d.world) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, &hh[(size_t)65 * nb], 64);
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      ok = false;
      break;
    }
    c->nb_mbox[side] = static_cast<char*>(ptr);
  }
  unsigned char mapped = ok;           // agree: all ranks or none
  CU(cudaMemcpyAsync(dh + d.rank, &mapped, 1, cudaMemcpyHostToDevice, c->stream));
  NC(g_nccl.AllGather(dh + d.rank, dh, 1, ncclUint8, c->comm, c->stream));
  CU(cudaMemcpyAsync(hh.data(), dh, d.world, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  for (int w = 0; w < d.world; ++w) ok = ok && hh[w];
  c->mbox_ok = ok;
  CU(cudaFree(dh));
  return SPL_OK;
}

static int create_impl(spl_ctx* c, const spl_desc* desc) {
  c->d = *desc;
  spl_desc& d = c->d;
  if (d.nx < 1 || d.ny < 1 || d.nz < 1)
    return fail(c, SPL_E_BADARG, "box %d x %d x %d is empty", d.nx, d.ny, d.nz);
  if ((long long)d.nx * d.ny * ((long long)d.nz + 8) >= 2147483647LL)
    return fail(c, SPL_E_BADARG, "box %d x %d x %d exceeds 2^31 cells per context: "
                "cut it into z-slabs", d.nx, d.ny, d.nz);
  if (d.dtype != SPL_F32 && d.dtype != SPL_F64)
    return fail(c, SPL_E_BADARG, "dtype %d is neither SPL_F32 nor SPL_F64", d.dtype);
  if (!(d.h > 0.0) || !(d.rho > 0.0))
    return fail(c, SPL_E_BADARG, "h and rho must be positive");
  if ((long long)((d.nx + 127) / 128) * ((d.ny + TILE_Y - 1) / TILE_Y) > MAX_PARTIALS)
    return fail(c, SPL_E_BADARG, "box %d x %d: more than %d x-y tiles per plane (cut it along y)",
                d.nx, d.ny, MAX_PARTIALS);
  if (d.world < 1) d.world = 1;
  if (d.world > MAXW) return fail(c, SPL_E_BADARG, "world %d exceeds %d", d.world, MAXW);
  if (d.rank < 0 || d.rank >= d.world)
    return fail(c, SPL_E_BADARG, "rank %d outside world %d", d.rank, d.world);
  if (d.nz_global <= 0) d.nz_global = d.nz;
  if (d.world == 1) { d.z0 = 0; d.nz_global = d.nz; }
  if (d.z0 < 0 || d.z0 + d.nz > d.nz_global)
    return fail(c, SPL_E_BADARG, "slab [%d,%d) outside nz_global %d", d.z0, d.z0 + d.nz,
                d.nz_global);
  if (d.max_iters < 0) d.max_iters = 0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(c, SPL_E_CUDA, "no CUDA device: libsplashy_cuda has no CPU fallback");
  if (d.device < 0 || d.device >= ndev)
    return fail(c, SPL_E_BADARG, "device %d of %d", d.device, ndev);
  CU(cudaSetDevice(d.device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, d.device));
  c->sms = prop.multiProcessorCount;
  c->esz = d.dtype == SPL_F64 ? 8 : 4;
  c->vec = (d.nx % 4 == 0) ? 4 : 1;
  Geo& g = c->g;
  g.nx = d.nx; g.ny = d.ny; g.nz = d.nz;
  g.hz = d.world > 1 ? 4 : 1;
  g.z0 = d.z0; g.nzg = d.nz_global;
  g.plane = (long long)d.nx * d.ny;
  g.ncell = g.plane * d.nz;
  if (const char* e = getenv("SPL_ZCHUNK")) c->zchunk_override = atoi(e);
  c->refine_rounds = (d.dtype == SPL_F32 && d.precond == SPL_PC_MG) ? 1 : 0;
  if (const char* e = getenv("SPL_REFINE")) c->refine_rounds = atoi(e) > 0 ? atoi(e) : 0;
  if (const char* e = getenv("SPL_REFINE_FIRST_TOL")) c->refine_first_tol = atof(e);
  if (const char* e = getenv("SPL_NO_RING")) c->no_ring = atoi(e) != 0;
  if (const char* e = getenv("SPL_NO_WS")) c->no_ws = atoi(e) != 0;
  if (const char* e = getenv("SPL_TMA")) c->use_tma = atoi(e) != 0;
  if (const char* e = getenv("SPL_MG_TMA")) c->mg_no_tma = atoi(e) == 0;
  if (const char* e = getenv("SPL_TMA_L2")) c->tma_l2 = atoi(e) & 3;
  c->tmaps.reserve(256);
  if (const char* e = getenv("SPL_NO_QUAD")) c->no_quad = atoi(e) != 0;
  if (const char* e = getenv("SPL_NO_GC_FUSE")) c->no_gc_fuse = atoi(e) != 0;
  if (const char* e = getenv("SPL_NO_RB_MARCH")) c->no_rb_march = atoi(e) != 0;
  if (const char* e = getenv("SPL_NO_2D_VIEW")) c->no_2d_view = atoi(e) != 0;
  if (const char* e = getenv("SPL_ADV_TMA")) c->adv_tma = atoi(e) != 0;
  if (const char* e = getenv("SPL_ADV_FUSE")) c->adv_fuse = atoi(e) != 0;
  if (const char* e = getenv("SPL_ADV_ZCHUNK")) c->adv_zchunk = atoi(e);

  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (auto& e : c->ev) CU(cudaEventCreate(&e));
  for (auto& e : c->tev) CU(cudaEventCreate(&e));
  if (d.world > 1) {
    std::string err;
    if (!d.nccl_id) return fail(c, SPL_E_BADARG, "world > 1 needs desc.nccl_id");
    if (!load_nccl(err)) return fail(c, SPL_E_CUDA, "%s", err.c_str());
    ncclUniqueId id;
    memcpy(&id, d.nccl_id, 128);
    NC(g_nccl.CommInitRank(&c->comm, d.world, id, d.rank));
    if (const char* e = getenv("SPL_OVERLAP")) c->overlap = atoi(e) != 0;
    if (const char* e = getenv("SPL_SPLIT")) c->split = atoi(e) != 0;
    int plo = 0, phi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&plo, &phi));
    CU(cudaStreamCreateWithPriority(&c->cstream, cudaStreamNonBlocking, phi));
    CU(cudaEventCreateWithFlags(&c->ev_beta, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming));
    // peer boxes for the CG scalars: allocate, exchange the IPC handles through NCCL, map the
    // peers' boxes; used only if every rank could map every box
    const char* pe = getenv("SPL_P2P");
    const bool want = !(pe && atoi(pe) == 0);
    CU(cudaMalloc(&c->pbox, sizeof(PeerShared)));
    CU(cudaMemsetAsync(c->pbox, 0, sizeof(PeerShared), c->stream));
    CU(cudaMalloc(&c->hticket, sizeof(unsigned)));
    CU(cudaMemsetAsync(c->hticket, 0, sizeof(unsigned), c->stream));
    cudaIpcMemHandle_t mine;
    unsigned char ok = want && cudaIpcGetMemHandle(&mine, c->pbox) == cudaSuccess;
    if (!ok) { cudaGetLastError(); memset(&mine, 0, sizeof mine); }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    unsigned char* dh = nullptr;
    CU(cudaMalloc(&dh, (size_t)65 * d.world));
    std::vector<unsigned char> hh((size_t)65 * d.world, 0);
    memcpy(&hh[(size_t)65 * d.rank], &mine, 64);
    hh[(size_t)65 * d.rank + 64] = ok;
    CU(cudaMemcpyAsync(dh + (size_t)65 * d.rank, &hh[(size_t)65 * d.rank], 65,
                       cudaMemcpyHostToDevice, c->stream));
    NC(g_nccl.AllGather(dh + (size_t)65 * d.rank, dh, 65, ncclUint8, c->comm, c->stream));
    CU(cudaMemcpyAsync(hh.data(), dh, hh.size(), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    bool all = true;
    for (int w = 0; w < d.world; ++w) all = all && hh[(size_t)65 * w + 64];
    c->peers.p[d.rank] = &c->pbox->gather;
    c->fboxes[d.rank] = &c->pbox->fused;
    for (int w = 0; w < d.world && all; ++w) {
      if (w == d.rank) continue;
      cudaIpcMemHandle_t h;
      memcpy(&h, &hh[(size_t)65 * w], 64);
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        all = false;
        break;
      }
      c->peer_maps[w] = ptr;
      c->peers.p[w] = &static_cast<PeerShared*>(ptr)->gather;
      c->fboxes[w] = &static_cast<PeerShared*>(ptr)->fused;
    }
    // second round: a rank that failed to map switches everybody back to ncclAllGather
    unsigned char mapped = all;
    CU(cudaMemcpyAsync(dh + d.rank, &mapped, 1, cudaMemcpyHostToDevice, c->stream));
    NC(g_nccl.AllGather(dh + d.rank, dh, 1, ncclUint8, c->comm, c->stream));
    CU(cudaMemcpyAsync(hh.data(), dh, d.world, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int w = 0; w < d.world; ++w) all = all && hh[w];
    c->p2p = all;
    CU(cudaFree(dh));
    // slab geometry of every rank (the multigrid hierarchy must coarsen in lock step)
    int geo[2] = {d.nz, d.z0};
    int* dg = nullptr;
    CU(cudaMalloc(&dg, sizeof(int) * 2 * d.world));
    CU(cudaMemcpyAsync(dg + 2 * d.rank, geo, sizeof geo, cudaMemcpyHostToDevice, c->stream));
    NC(g_nccl.AllGather(dg + 2 * d.rank, dg, 2, ncclInt32, c->comm, c->stream));
    c->slab_geo.resize(2 * d.world);
    CU(cudaMemcpyAsync(c->slab_geo.data(), dg, sizeof(int) * 2 * d.world, cudaMemcpyDeviceToHost,
                       c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaFree(dg));
    int zsum = 0;
    for (int w = 0; w < d.world; ++w) {
      if (c->slab_geo[2 * w + 1] != zsum)
        return fail(c, SPL_E_BADARG, "slab of rank %d starts at plane %d, expected %d", w,
                    c->slab_geo[2 * w + 1], zsum);
      zsum += c->slab_geo[2 * w];
      // the advection halo (hz = 4 planes) comes from the direct neighbour only
      if (c->slab_geo[2 * w] < 4)
        return fail(c, SPL_E_BADARG, "slab of rank %d has %d planes: z-slabs need at least 4", w,
                    c->slab_geo[2 * w]);
    }
    if (zsum != d.nz_global)
      return fail(c, SPL_E_BADARG, "slabs cover %d planes, nz_global is %d", zsum, d.nz_global);
  }
  int rc;
  char* unused8 = nullptr;
  char* unused9 = nullptr;
  char** fields[] = {&c->u, &c->v, &c->w, &c->u2, &c->v2, &c->w2, &c->p, &c->r,
                     &unused8, &unused9, &c->q};
  for (int i = 0; i < 11; ++i) {   // p (slot 6) is always fp64
    if (i == 8 || i == 9) continue;
    if ((rc = alloc_field(c, i, i == 6 ? 8 : c->esz, fields[i], 0))) return rc;
  }
  if (const char* e = getenv("SPL_NSBUF")) c->nsbuf = atoi(e);
  if (c->nsbuf < 2) c->nsbuf = 2;
  if (c->nsbuf > MAX_SBUF) c->nsbuf = MAX_SBUF;
  for (int i = 0; i < c->nsbuf; ++i)
    if ((rc = alloc_field(c, 32 + i, c->esz, &c->sbuf[i], 0))) return rc;
  if (d.world > 1 && (rc = map_neighbour_sbufs(c))) return rc;
  if (d.world > 1 && (rc = map_mailboxes(c))) return rc;
  if (d.precond == SPL_PC_RBIC0 || d.precond == SPL_PC_MG) {
    if ((rc = alloc_field(c, 11, c->esz, &c->z, 0))) return rc;
    if ((rc = alloc_field(c, 12, c->esz, &c->einv, 0))) return rc;
  }
  char* tmp;
  if ((rc = alloc_field(c, 13, 1, &tmp, 0xFF))) return rc;   // SPL_ABSENT everywhere
  c->media = reinterpret_cast<uint8_t*>(tmp);
  if ((rc = alloc_field(c, 14, 1, &tmp, 0))) return rc;
  c->codes = reinterpret_cast<uint8_t*>(tmp);
  if ((rc = alloc_field(c, 24, 1, &tmp, 0xFF))) return rc;   // BFS layers (spl_cleanup)
  c->layer = reinterpret_cast<uint8_t*>(tmp);
  if (d.precond == SPL_PC_MG) {
    if (const char* e = getenv("SPL_MG_NU")) c->mg_nu = atoi(e) > 0 ? atoi(e) : 2;
    if (const char* e = getenv("SPL_MG_NO_MARCH")) c->mg_no_march = atoi(e) != 0;
    if (const char* e = getenv("SPL_MG_SUMFORM")) c->mg_sumform = atoi(e) != 0;
    if (const char* e = getenv("SPL_MG_NO_BOTTOM")) c->mg_no_bottom = atoi(e) != 0;
    if (const char* e = getenv("SPL_MG_NO_FUSE")) c->mg_no_fuse = atoi(e) != 0;
    if (const char* e = getenv("SPL_MG_NO_FUSE_SLABS")) c->mg_no_fuse_slabs = atoi(e) != 0;
    if (c->mg_nu > MG_MAX_NU) c->mg_nu = MG_MAX_NU;
    // Chebyshev weights for the eigenvalues of D^-1 A in [1/3, 2] (the high-frequency
    // band of the 7-point operator under 2x coarsening), ascending; the post-smoother
    // applies them in reverse, so the cycle stays symmetric.  oracle: dense_ref.mg_weights
    c->mg_omega.assign(c->mg_nu, 2.0 / 3.0);
    const char* ow = getenv("SPL_MG_OMEGA");
    for (int k = 0; k < c->mg_nu; ++k) {
      if (ow && atof(ow) > 0.0) c->mg_omega[k] = atof(ow);
      else c->mg_omega[k] = 1.0 / (7.0 / 6.0 + 5.0 / 6.0 * cos(M_PI * (2 * k + 1) / (2.0 * c->mg_nu)));
    }
    if (const char* e = getenv("SPL_MG_COARSE")) c->mg_coarse = atoi(e) > 1 ? atoi(e) & ~1 : 30;
    // level 0 aliases the fine arrays; coarser levels until the smallest extent <= 8
    MgLevel l0;
    l0.g = c->g;
    l0.media = c->media; l0.codes = c->codes;
    l0.u = c->z; l0.b = c->r; l0.tmp = c->q;
    c->mg.push_back(l0);
    while ((int)c->mg.size() < MG_MAX_LEVELS) {
      const Geo& f = c->mg.back().g;
      int mind = 1 << 30;
      if (f.nx > 1 && f.nx < mind) mind = f.nx;
      if (f.ny > 1 && f.ny < mind) mind = f.ny;
      if (f.nz > 1 && f.nz < mind) mind = f.nz;
      // z-slabs coarsen in lock step: every rank must hold whole coarse planes, and the
      // decision must be the same on every rank (each level adds collectives to the cycle),
      // so it is taken from the geometry of ALL slabs at this level
      if (d.world > 1) {
        const int lev = (int)c->mg.size() - 1;
        bool odd = (f.nzg & 1) != 0;
        for (int w = 0; w < d.world; ++w) {
          const int wnz = c->slab_geo[2 * w] >> lev, wz0 = c->slab_geo[2 * w + 1] >> lev;
          if ((wnz & 1) || (wz0 & 1)) odd = true;
          if (wnz > 1 && wnz < mind) mind = wnz;
        }
        if (odd) break;
      }
      if (mind <= 8 || mind == (1 << 30)) break;
      MgLevel l;
      l.g.nx = (f.nx + 1) / 2; l.g.ny = (f.ny + 1) / 2; l.g.nz = (f.nz + 1) / 2;
      l.g.hz = 1;
      l.g.z0 = (d.world > 1) ? f.z0 / 2 : 0;
      l.g.nzg = (d.world > 1) ? f.nzg / 2 : l.g.nz;
      l.g.plane = (long long)l.g.nx * l.g.ny;
      l.g.ncell = l.g.plane * l.g.nz;
      l.scale = c->mg.back().scale * 0.25;
      const size_t cells = (size_t)l.g.plane * (l.g.nz + 2);
      const size_t sizes[5] = {cells, cells, cells * c->esz, cells * c->esz, cells * c->esz};
      char* base[5];
      for (int a = 0; a < 5; ++a) {
        CU(cudaMalloc(&l.allocs[a], sizes[a]));
        CU(cudaMemsetAsync(l.allocs[a], a == 0 ? 0xFF : 0, sizes[a], c->stream));
        base[a] = static_cast<char*>(l.allocs[a]) + (size_t)l.g.plane * (a < 2 ? 1 : c->esz);
      }
      l.media = reinterpret_cast<uint8_t*>(base[0]);
      l.codes = reinterpret_cast<uint8_t*>(base[1]);
      l.u = base[2]; l.b = base[3]; l.tmp = base[4];
      c->mg.push_back(l);
    }
    // z-slabs: gather the coarsest level into one global grid when the slabs are equal
    // and the global grid fits one CTA's shared memory (otherwise the bottom of the cycle
    // stays local to each slab: a weaker, still symmetric, preconditioner)
    if (d.world > 1 && c->mg.size() >= 2 && !getenv("SPL_MG_LOCAL_BOTTOM")) {
      const Geo& k = c->mg.back().g;
      Geo gg = k;
      gg.nz = k.nzg; gg.z0 = 0; gg.ncell = gg.plane * gg.nz;
      const bool equal = (long long)k.nz * d.world == k.nzg && k.z0 == d.rank * k.nz;
      const size_t smem = (size_t)gg.ncell * (3 * c->esz + 1) + 64;
      if (equal && smem <= 216 * 1024) {
        c->mg_gbottom = true;
        c->mg_gg = gg;
        CU(cudaMalloc(&c->mg_gcodes, (size_t)gg.ncell));
        CU(cudaMalloc(&c->mg_gb, (size_t)gg.ncell * c->esz));
        CU(cudaMalloc(&c->mg_gu, (size_t)gg.ncell * c->esz));
        if (const char* e = getenv("SPL_MG_GSWEEPS")) c->mg_gsweeps = atoi(e) > 1 ? atoi(e) : 100;
      }
    }
  }
  CU(cudaMalloc(&c->sc, sizeof(Scal)));
  CU(cudaMemsetAsync(c->sc, 0, sizeof(Scal), c->stream));
  CU(cudaMalloc(&c->partials, sizeof(double) * 2 * MAX_PARTIALS));
  CU(cudaMalloc(&c->flag, sizeof(int)));
  CU(cudaHostAlloc(&c->hsc, sizeof(Scal), cudaHostAllocMapped));
  CU(cudaHostAlloc(&c->hflag, sizeof(int), cudaHostAllocMapped));
  CU(cudaHostAlloc(&c->hdone, 2 * sizeof(int), cudaHostAllocMapped));
  CU(cudaHostAlloc(&c->hesc, sizeof(unsigned long long), cudaHostAllocMapped));
  CU(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->hsc_d), c->hsc, 0));
  CU(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->hflag_d), c->hflag, 0));
  CU(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->hdone_d), c->hdone, 0));
  CU(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->hesc_d), c->hesc, 0));
  *c->hesc = 0;
  *c->hflag = 0;
  CU(cudaHostAlloc(&c->hagree, (MAXW + 1) * sizeof(double), cudaHostAllocDefault));
  for (auto& e : c->dev) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));

  c->d.nccl_id = nullptr;   // caller-owned bytes are not kept
  for (int a = 0; a < 3; ++a) c->world.lo[a] = 0;
  c->world.hi[0] = c->g.nx - 1; c->world.hi[1] = c->g.ny - 1; c->world.hi[2] = c->g.nzg - 1;
  CU(cudaStreamSynchronize(c->stream));
  return SPL_OK;
}

int spl_create(spl_ctx** out, const spl_desc* desc) {
  spl_ctx* none = nullptr;
  if (!out || !desc) return fail(none, SPL_E_BADARG, "spl_create: NULL argument");
  *out = nullptr;
  spl_ctx* c = new spl_ctx();
  const int rc = create_impl(c, desc);
  if (rc != SPL_OK) {
    g_create_error = c->err;
    spl_destroy(c);
    return rc;
  }
  *out = c;
  return SPL_OK;
}

// everything derived from the labels: stencil codes, preconditioner data
static int rebuild_operators(spl_ctx* c) {
  const Geo& g = c->g;
  int rc;
  LAUNCH(c, k_codes, grid_for(c, g.ncell), 256, g, c->media, c->codes);
  CU(cudaGetLastError());
  if (c->d.precond == SPL_PC_RBIC0) {
    rc = DISPATCH(c, Impl<float>::build_rbic0(c), Impl<double>::build_rbic0(c));
    if (rc) return rc;
  }
  // SPL_PC_MG: coarse labels and stencil codes.  With z-slabs the smoothers read the
  // labels / codes of the neighbouring slab's boundary plane through the halo.
  if (!c->mg.empty() && (rc = exchange_level(c, g, c->codes, 1, ncclUint8))) return rc;
  for (size_t l = 1; l < c->mg.size(); ++l) {
    MgLevel& f = c->mg[l - 1];
    MgLevel& k = c->mg[l];
    LAUNCH(c, k_mg_coarsen, grid_for(c, k.g.ncell), 256, f.g, f.media, k.g, k.media);
    if ((rc = exchange_level(c, k.g, k.media, 1, ncclUint8))) return rc;
    LAUNCH(c, k_codes, grid_for(c, k.g.ncell), 256, k.g, k.media, k.codes);
    CU(cudaGetLastError());
    if ((rc = exchange_level(c, k.g, k.codes, 1, ncclUint8))) return rc;
  }
  if (c->mg_gbottom) {
    MgLevel& k = c->mg.back();
    NC(g_nccl.AllGather(k.codes, c->mg_gcodes, (size_t)k.g.ncell, ncclUint8, c->comm, c->stream));
  }
  return SPL_OK;
}

// host -> device copies of one upload on stream `st`
static int enqueue_upload(spl_ctx* c, cudaStream_t st, const uint8_t* media, const void* u,
                          const void* v, const void* w, const void* p) {
  const Geo& g = c->g;
  const size_t nb = (size_t)g.ncell * c->esz;
  CU(cudaMemcpyAsync(c->media, media, (size_t)g.ncell, cudaMemcpyHostToDevice, st));
  const void* src[3] = {u, v, w};
  char* dst[3] = {c->u, c->v, c->w};
  for (int i = 0; i < 3; ++i) {
    if (src[i]) CU(cudaMemcpyAsync(dst[i], src[i], nb, cudaMemcpyHostToDevice, st));
    else CU(cudaMemsetAsync(dst[i], 0, nb, st));
  }
  // pressure is kept in fp64 on the device
  if (!p) {
    CU(cudaMemsetAsync(c->p, 0, (size_t)g.ncell * 8, st));
  } else if (c->d.dtype == SPL_F64) {
    CU(cudaMemcpyAsync(c->p, p, (size_t)g.ncell * 8, cudaMemcpyHostToDevice, st));
  } else {
    CU(cudaMemcpyAsync(c->q, p, nb, cudaMemcpyHostToDevice, st));
    k_convert<float, double><<<grid_for(c, g.ncell), 256, 0, st>>>(
        g.ncell, reinterpret_cast<const float*>(c->q), reinterpret_cast<double*>(c->p));
    c->launches++;
  }
  return SPL_OK;
}

// what follows the copies, on the compute stream: label halos and everything derived from
// the labels (collectives when world > 1, hence always in program order on c->stream)
static int finish_upload(spl_ctx* c) {
  CU(cudaSetDevice(c->d.device));
  if (c->upload_pending) {
    CU(cudaStreamWaitEvent(c->stream, c->ev_up, 0));
    c->upload_pending = false;
  }
  const Geo& g = c->g;
  int rc = exchange_halo(c, reinterpret_cast<char*>(c->media), 1, ncclUint8, g.hz);
  if (rc) return rc;
  if ((rc = rebuild_operators(c))) return rc;
  CU(cudaStreamSynchronize(c->stream));
  c->uploaded = true;
  c->have_rhs = false;
  return SPL_OK;
}

int spl_upload(spl_ctx* c, const uint8_t* media, const void* u, const void* v,
               const void* w, const void* p) {
  if (!c) return SPL_E_BADARG;
  if (!media) return fail(c, SPL_E_BADARG, "spl_upload: media is required");
  CU(cudaSetDevice(c->d.device));
  if (c->h2d) CU(cudaStreamSynchronize(c->h2d));      // an earlier asynchronous upload / download
  if (c->d2h) CU(cudaStreamSynchronize(c->d2h));
  c->upload_pending = false;
  int rc = enqueue_upload(c, c->stream, media, u, v, w, p);
  if (rc) return rc;
  return finish_upload(c);
}

static int async_streams(spl_ctx* c) {
  if (c->h2d) return SPL_OK;
  CU(cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&c->ev_up, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_down, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
  return SPL_OK;
}

// spl_upload without waiting: the copies are queued on the context's host-to-device stream
// and the call returns; the first entry point that needs the state (spl_step, ...) waits for
// them on the device and builds the label-derived data.  The host buffers (pinned, or the
// call degenerates to a synchronous copy) must stay untouched until then.  Copies of one
// context overlap the kernels of ANOTHER context: two boxes in flight per GPU hide the PCIe
// time of upload -> step -> download behind the step of the other box.
int spl_upload_async(spl_ctx* c, const uint8_t* media, const void* u, const void* v,
                     const void* w, const void* p) {
  if (!c) return SPL_E_BADARG;
  if (!media) return fail(c, SPL_E_BADARG, "spl_upload_async: media is required");
  CU(cudaSetDevice(c->d.device));
  int rc = async_streams(c);
  if (rc) return rc;
  // the fields may still be read by kernels of this context or by a queued download
  CU(cudaEventRecord(c->ev_done, c->stream));
  CU(cudaStreamWaitEvent(c->h2d, c->ev_done, 0));
  if (c->download_queued) CU(cudaStreamWaitEvent(c->h2d, c->ev_down, 0));
  if ((rc = enqueue_upload(c, c->h2d, media, u, v, w, p))) return rc;
  CU(cudaEventRecord(c->ev_up, c->h2d));
  c->upload_pending = true;
  c->uploaded = false;
  return SPL_OK;
}

static int enqueue_download(spl_ctx* c, cudaStream_t st, void* u, void* v, void* w, void* p) {
  const size_t nb = (size_t)c->g.ncell * c->esz;
  void* dst[3] = {u, v, w};
  char* src[3] = {c->u, c->v, c->w};
  for (int i = 0; i < 3; ++i)
    if (dst[i]) CU(cudaMemcpyAsync(dst[i], src[i], nb, cudaMemcpyDeviceToHost, st));
  if (p) {
    if (c->d.dtype == SPL_F64) {
      CU(cudaMemcpyAsync(p, c->p, nb, cudaMemcpyDeviceToHost, st));
    } else {   // fp64 accumulator -> fp32 at the boundary
      k_convert<double, float><<<grid_for(c, c->g.ncell), 256, 0, st>>>(
          c->g.ncell, reinterpret_cast<const double*>(c->p), reinterpret_cast<float*>(c->q));
      c->launches++;
      CU(cudaMemcpyAsync(p, c->q, nb, cudaMemcpyDeviceToHost, st));
    }
  }
  return SPL_OK;
}

int spl_download(spl_ctx* c, void* u, void* v, void* w, void* p, void* div) {
  if (!c) return SPL_E_BADARG;
  NEED_UPLOAD(c, "spl_download");
  CU(cudaSetDevice(c->d.device));
  if (c->d2h) CU(cudaStreamSynchronize(c->d2h));      // c->q may hold a queued download's pressure
  int rc = enqueue_download(c, c->stream, u, v, w, p);
  if (rc) return rc;
  if (div) {
    const size_t nb = (size_t)c->g.ncell * c->esz;
    rc = DISPATCH(c, Impl<float>::rhs(c, 1.0f, c->q), Impl<double>::rhs(c, 1.0, c->q));
    if (rc) return rc;
    CU(cudaMemcpyAsync(div, c->q, nb, cudaMemcpyDeviceToHost, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  return SPL_OK;
}

// spl_download (without div) queued on the device-to-host stream; spl_sync waits for it.
int spl_download_async(spl_ctx* c, void* u, void* v, void* w, void* p) {
  if (!c) return SPL_E_BADARG;
  NEED_UPLOAD(c, "spl_download_async");
  CU(cudaSetDevice(c->d.device));
  int rc = async_streams(c);
  if (rc) return rc;
  CU(cudaEventRecord(c->ev_done, c->stream));
  CU(cudaStreamWaitEvent(c->d2h, c->ev_done, 0));
  if ((rc = enqueue_download(c, c->d2h, u, v, w, p))) return rc;
  CU(cudaEventRecord(c->ev_down, c->d2h));
  c->download_queued = true;
  return SPL_OK;
}

// waits for every queued copy of this context (not for a pending upload's derived data:
// that is built by the next entry point that needs it)
int spl_sync(spl_ctx* c) {
  if (!c) return SPL_E_BADARG;
  CU(cudaSetDevice(c->d.device));
  if (c->h2d) CU(cudaStreamSynchronize(c->h2d));
  if (c->d2h) CU(cudaStreamSynchronize(c->d2h));
  c->download_queued = false;
  return SPL_OK;
}

int spl_advect(spl_ctx* c, double dt, spl_info* info) {
  if (!c) return SPL_E_BADARG;
  NEED_UPLOAD(c, "spl_advect");
  CU(cudaSetDevice(c->d.device));
  c->launches = 0;
  CU(cudaMemsetAsync(&c->sc->escapes, 0, sizeof(unsigned long long), c->stream));
  CU(cudaEventRecord(c->ev[0], c->stream));
  int rc = DISPATCH(c, Impl<float>::advect(c, dt), Impl<double>::advect(c, dt));
  if (rc) return rc;
  CU(cudaEventRecord(c->ev[1], c->stream));
  if ((rc = to_host(c, c->hflag_d, c->flag, sizeof(int)))) return rc;
  if ((rc = pull_scal(c))) return rc;
  c->have_rhs = false;
  if (info) {
    memset(info, 0, sizeof *info);
    info->ms_advect = info->ms_total = ev_ms(c, 0, 1);
    info->trace_escapes = (int64_t)c->hsc->escapes;
    info->kernel_launches = c->launches;
  }
  const bool bad = (c->d.quirks && *c->hflag) || c->hsc->escapes;
  double any = 0.0;
  if ((rc = agree_max(c, bad ? 1.0 : 0.0, &any))) return rc;
  if (any > 0.0)   // Convection.fs:65; every rank of a slab decomposition reports it
    return fail(c, SPL_E_TRACE, bad ? "Backwards particle trace went too far."
                                    : "Backwards particle trace went too far. (on another rank)");
  return SPL_OK;
}

int spl_set_scalars(spl_ctx* c, int n, const void* const* fields) {
  if (!c || n < 0 || n > SPL_MAX_SCALARS || (n > 0 && !fields)) return SPL_E_BADARG;
  NEED_UPLOAD(c, "spl_set_scalars");
  CU(cudaSetDevice(c->d.device));
  int rc;
  while ((int)c->scal.size() < n) {
    char* p0 = nullptr;
    if ((rc = alloc_field(c, 48 + (int)c->scal.size(), c->esz, &p0, 0))) return rc;
    c->scal.push_back(p0);
  }
  if (n > 0 && !c->scal2 && (rc = alloc_field(c, 48 + SPL_MAX_SCALARS, c->esz, &c->scal2, 0)))
    return rc;
  c->scal.resize(n);        // dropped fields keep their allocation until spl_destroy
  const size_t nb = (size_t)c->g.ncell * c->esz;
  for (int i = 0; i < n; ++i) {
    if (fields[i]) CU(cudaMemcpyAsync(c->scal[i], fields[i], nb, cudaMemcpyHostToDevice, c->stream));
    else CU(cudaMemsetAsync(c->scal[i], 0, nb, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  return SPL_OK;
}

int spl_get_scalars(spl_ctx* c, int n, void* const* fields) {
  if (!c || n < 0 || (n > 0 && !fields)) return SPL_E_BADARG;
  if (n > (int)c->scal.size())
    return fail(c, SPL_E_BADARG, "spl_get_scalars: %d requested, %d set", n, (int)c->scal.size());
  CU(cudaSetDevice(c->d.device));
  const size_t nb = (size_t)c->g.ncell * c->esz;
  for (int i = 0; i < n; ++i)
    if (fields[i]) CU(cudaMemcpyAsync(fields[i], c->scal[i], nb, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SPL_OK;
}

int spl_add_forces(spl_ctx* c, double dt) {
  if (!c) return SPL_E_BADARG;
  NEED_UPLOAD(c, "spl_add_forces");
  CU(cudaSetDevice(c->d.device));
  int rc = DISPATCH(c, Impl<float>::forces(c, dt), Impl<double>::forces(c, dt));
  if (rc) return rc;
  CU(cudaStreamSynchronize(c->stream));
  c->have_rhs = false;
  return SPL_OK;
}

int spl_add_viscosity(spl_ctx* c, double dt) {
  if (!c) return SPL_E_BADARG;
  NEED_UPLOAD(c, "spl_add_viscosity");
  CU(cudaSetDevice(c->d.device));
  int rc = DISPATCH(c, Impl<float>::viscosity(c, dt), Impl<double>::viscosity(c, dt));
  if (rc) return rc;
  CU(cudaStreamSynchronize(c->stream));
  c->have_rhs = false;
  return SPL_OK;
}

int spl_cleanup(spl_ctx* c) {
  if (!c) return SPL_E_BADARG;
  NEED_UPLOAD(c, "spl_cleanup");
  CU(cudaSetDevice(c->d.device));
  int rc = DISPATCH(c, Impl<float>::cleanup(c), Impl<double>::cleanup(c));
  if (rc) return rc;
  CU(cudaStreamSynchronize(c->stream));
  c->have_rhs = false;
  return SPL_OK;
}

int spl_divergence(spl_ctx* c, double dt, spl_info* info) {
  if (!c) return SPL_E_BADARG;
  NEED_UPLOAD(c, "spl_divergence");
  if (!(dt > 0.0)) return fail(c, SPL_E_BADARG, "dt must be positive");
  CU(cudaSetDevice(c->d.device));
  c->launches = 0;
  CU(cudaEventRecord(c->ev[0], c->stream));
  int rc = do_rhs(c, dt);
  if (rc) return rc;
  CU(cudaEventRecord(c->ev[1], c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (info) {
    memset(info, 0, sizeof *info);
    info->ms_rhs = info->ms_total = ev_ms(c, 0, 1);
    info->kernel_launches = c->launches;
  }
  return SPL_OK;
}

static int project_impl(spl_ctx* c, double dt, spl_info* info, bool with_advect) {
  if (!(dt > 0.0)) return fail(c, SPL_E_BADARG, "dt must be positive");
  CU(cudaSetDevice(c->d.device));
  c->launches = 0;
  int rc;
  CU(cudaMemsetAsync(&c->sc->escapes, 0, sizeof(unsigned long long), c->stream));
  CU(cudaEventRecord(c->ev[0], c->stream));
  if (with_advect) {
    if ((rc = DISPATCH(c, Impl<float>::advect(c, dt), Impl<double>::advect(c, dt)))) return rc;
    // the escape counter and the quirk flag leave the device now (the solve resets the scalar
    // block) but are only looked at after the step: no host round trip between the stages
    if ((rc = to_host(c, c->hesc_d, &c->sc->escapes, sizeof(unsigned long long)))) return rc;
    if ((rc = to_host(c, c->hflag_d, c->flag, sizeof(int)))) return rc;
  }
  CU(cudaEventRecord(c->ev[1], c->stream));
  if ((rc = do_rhs(c, dt))) return rc;
  CU(cudaEventRecord(c->ev[2], c->stream));
  c->iters_base = 0;
  // with a refinement round to follow, the first pass only needs to reach the level where
  // the fp32 recursive residual stops tracking the true one (~1e-4 max|b| on large grids);
  // the restart from the fp64 residual does the rest
  const bool refine = c->refine_rounds > 0 && c->fixed_iters == 0;
  c->pcg_tol_first = (refine && c->d.tol < c->refine_first_tol) ? c->refine_first_tol : 0.0;
  rc = DISPATCH(c, Impl<float>::pcg_dispatch(c), Impl<double>::pcg_dispatch(c));
  c->pcg_tol_first = 0.0;
  if (rc) return rc;
  for (int round = 0; round < c->refine_rounds && c->fixed_iters == 0; ++round) {
    // iterative refinement: r := b - A x in fp64, PCG restarted on it with x kept
    if ((rc = pull_scal(c))) return rc;
    if (!c->hsc->converged || c->hsc->breakdown || !(c->hsc->bmax > 0.0)) break;
    const double bmax = c->hsc->bmax;
    const int it0 = c->hsc->iters;
    if ((rc = do_rhs(c, dt))) return rc;
    if ((rc = exchange_halo(c, c->p, 8, ncclDouble, 1))) return rc;
    if (c->d.dtype == SPL_F64)
      LAUNCH(c, k_true_residual<double>, grid_for(c, c->g.ncell), 256, c->g, c->codes,
             reinterpret_cast<const double*>(c->p), reinterpret_cast<double*>(c->r));
    else if (c->g.nx % 4 == 0 && !c->no_quad)
      LAUNCH(c, k_true_residual4, grid_for(c, c->g.ncell / 4), 256, c->g, c->codes,
             reinterpret_cast<const double*>(c->p), reinterpret_cast<float*>(c->r));
    else
      LAUNCH(c, k_true_residual<float>, grid_for(c, c->g.ncell), 256, c->g, c->codes,
             reinterpret_cast<const double*>(c->p), reinterpret_cast<float*>(c->r));
    CU(cudaGetLastError());
    c->pcg_keep_x = true;
    c->pcg_bmax_keep = bmax;
    rc = DISPATCH(c, Impl<float>::pcg_dispatch(c), Impl<double>::pcg_dispatch(c));
    c->pcg_keep_x = false;
    if (rc) return rc;
    c->iters_base += it0;
  }
  CU(cudaEventRecord(c->ev[3], c->stream));
  if (c->d.dtype == SPL_F32 && Impl<float>::grad_check_ok(c)) {
    if ((rc = Impl<float>::grad_check(c, dt))) return rc;       // one pass, 34 B / cell
    CU(cudaEventRecord(c->ev[4], c->stream));
  } else {
    if ((rc = DISPATCH(c, Impl<float>::grad(c, dt), Impl<double>::grad(c, dt)))) return rc;
    CU(cudaEventRecord(c->ev[4], c->stream));
    if ((rc = DISPATCH(c, Impl<float>::check(c), Impl<double>::check(c)))) return rc;
  }
  CU(cudaEventRecord(c->ev[5], c->stream));
  spl_info local;
  memset(&local, 0, sizeof local);
  if ((rc = read_back(c, &local, true, true))) return rc;      // synchronises the stream
  const unsigned long long escapes = with_advect ? *c->hesc : 0ull;
  const int too_far = (with_advect && c->d.quirks) ? *c->hflag : 0;
  local.trace_escapes = (int64_t)escapes;
  local.ms_advect = ev_ms(c, 0, 1);
  local.ms_rhs = ev_ms(c, 1, 2);
  local.ms_solve = ev_ms(c, 2, 3);
  local.ms_grad = ev_ms(c, 3, 4);
  local.ms_total = ev_ms(c, 0, 5);
  if (info) *info = local;
  c->have_rhs = false;
  if (escapes || too_far)
    return fail(c, SPL_E_TRACE, "Backwards particle trace went too far.");
  if (c->hsc->breakdown)
    return fail(c, SPL_E_NONFINITE, "Invalid pressure.");                 // Pressure.fs:138
  if (!local.converged && c->fixed_iters == 0)
    return fail(c, SPL_E_NOT_CONVERGED,
                "PCG stopped after %d iterations with max|r| = %g (max|b| = %g)", local.iters,
                local.r_inf, local.b_inf);
  return SPL_OK;
}

int spl_project(spl_ctx* c, double dt, spl_info* info) {
  if (!c) return SPL_E_BADARG;
  NEED_UPLOAD(c, "spl_project");
  return project_impl(c, dt, info, false);
}

int spl_step(spl_ctx* c, double dt, spl_info* info) {
  if (!c) return SPL_E_BADARG;
  NEED_UPLOAD(c, "spl_step");
  return project_impl(c, dt, info, true);
}

// ---- marker stage ---------------------------------------------------------
static bool integral_h(const spl_ctx* c) {
  return c->d.h >= 1.0 && c->d.h == std::floor(c->d.h) && c->d.h < 2147483647.0;
}

static int markers_need(spl_ctx* c, const char* who) {
  if (c->upload_pending) { const int rc_ = finish_upload(c); if (rc_) return rc_; }
  if (!c->uploaded) return fail(c, SPL_E_STATE, "%s before spl_upload", who);
  if (!integral_h(c))   // Coord.nearest works on int<m> (Coord.fs:49-57, Constants.fs:11 h = 10)
    return fail(c, SPL_E_BADARG, "%s: the marker stage needs an integral cell size h >= 1", who);
  if (c->nmark <= 0) return fail(c, SPL_E_STATE, "%s before spl_set_markers", who);
  if (c->d.world > 1 && (c->d.quirks & SPL_Q_MOVE_CELLS))   // a record would have to change rank
    return fail(c, SPL_E_BADARG, "%s: SPL_Q_MOVE_CELLS needs world = 1", who);
  return SPL_OK;
}

// the counters of the marker kernels, summed over the ranks: every rank takes the same
// decision (an error that one rank alone acted on would leave the others in a collective)
static int read_marker_stats(spl_ctx* c, MarkerStats* out) {
  static_assert(sizeof(MarkerStats) == 3 * sizeof(unsigned long long), "three counters");
  if (c->d.world > 1)
    NC(g_nccl.AllReduce(c->mstats, c->mstats, 3, ncclUint64, ncclSum, c->comm, c->stream));
  CU(cudaMemcpyAsync(out, c->mstats, sizeof(MarkerStats), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SPL_OK;
}

int spl_set_world(spl_ctx* c, const int lo[3], const int hi[3]) {
  if (!c || !lo || !hi) return SPL_E_BADARG;
  for (int a = 0; a < 3; ++a) {
    c->world.lo[a] = lo[a];
    c->world.hi[a] = hi[a];
  }
  return SPL_OK;
}

int spl_set_markers(spl_ctx* c, int64_t n, const double* xyz) {
  if (!c || n < 0 || (n > 0 && !xyz)) return SPL_E_BADARG;
  if (!integral_h(c))
    return fail(c, SPL_E_BADARG, "spl_set_markers: the marker stage needs an integral cell size h >= 1");
  CU(cudaSetDevice(c->d.device));
  if (c->mloc) { cudaFree(c->mloc); cudaFree(c->mold); cudaFree(c->mnew); }
  c->mloc = nullptr; c->mold = c->mnew = nullptr;
  c->nmark = n;
  if (!c->mstats) CU(cudaMalloc(&c->mstats, sizeof(MarkerStats)));
  if (!c->mark) {   // allocated like a field: the relabelling reads the neighbour slab's plane
    CU(cudaMalloc(&c->mark_base, field_bytes(c, 1)));
    CU(cudaMemsetAsync(c->mark_base, 0, field_bytes(c, 1), c->stream));
    c->mark = static_cast<uint8_t*>(c->mark_base) + (size_t)c->g.hz * c->g.plane;
  }
  if (n == 0) return SPL_OK;
  CU(cudaMalloc(&c->mloc, (size_t)n * 3 * sizeof(double)));
  CU(cudaMalloc(&c->mold, (size_t)n * sizeof(long long)));
  CU(cudaMalloc(&c->mnew, (size_t)n * sizeof(long long)));
  CU(cudaMemcpyAsync(c->mloc, xyz, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice,
                     c->stream));
  CU(cudaMemsetAsync(c->mstats, 0, sizeof(MarkerStats), c->stream));
  LAUNCH(c, k_marker_cells, grid_for(c, n), 256, c->g, c->mloc, (long long)n, (int)c->d.h,
         c->d.origin[0], c->d.origin[1], c->d.origin[2], c->mnew, c->mstats, 1);
  CU(cudaMemcpyAsync(c->mold, c->mnew, (size_t)n * sizeof(long long), cudaMemcpyDeviceToDevice,
                     c->stream));
  MarkerStats st;   // every rank holds every marker and counts the same ones: no reduction
  CU(cudaMemcpyAsync(&st, c->mstats, sizeof(MarkerStats), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (st.outside)
    return fail(c, SPL_E_MARKER, "%llu marker(s) lie outside the dense box", st.outside);
  return SPL_OK;
}

int spl_get_markers(spl_ctx* c, double* xyz) {
  if (!c || !xyz) return SPL_E_BADARG;
  if (c->nmark <= 0) return fail(c, SPL_E_STATE, "spl_get_markers before spl_set_markers");
  CU(cudaSetDevice(c->d.device));
  CU(cudaMemcpyAsync(xyz, c->mloc, (size_t)c->nmark * 3 * sizeof(double),
                     cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SPL_OK;
}

int spl_move_markers(spl_ctx* c, double dt, int64_t* moved) {
  if (!c) return SPL_E_BADARG;
  int rc = markers_need(c, "spl_move_markers");
  if (rc) return rc;
  CU(cudaSetDevice(c->d.device));
  const long long n = c->nmark;
  const int sum_faces = (c->d.quirks & SPL_Q_MARKER_SUM) ? 1 : 0;
  // a marker in plane 0 of a slab reads the -z face of the neighbour's top plane
  if ((rc = exchange_velocity(c, 1))) return rc;
  CU(cudaMemsetAsync(c->mstats, 0, sizeof(MarkerStats), c->stream));
  if (c->d.dtype == SPL_F64) {
    LAUNCH(c, k_move_markers<double>, grid_for(c, n), 256, c->g, c->media,
           Impl<double>::P(c->u), Impl<double>::P(c->v), Impl<double>::P(c->w), c->mloc, n, dt,
           (int)c->d.h, c->d.origin[0], c->d.origin[1], c->d.origin[2], sum_faces, c->mold,
           c->mnew, c->mstats, c->d.rank, c->d.world);
  } else {
    LAUNCH(c, k_move_markers<float>, grid_for(c, n), 256, c->g, c->media, Impl<float>::P(c->u),
           Impl<float>::P(c->v), Impl<float>::P(c->w), c->mloc, n, dt, (int)c->d.h,
           c->d.origin[0], c->d.origin[1], c->d.origin[2], sum_faces, c->mold, c->mnew,
           c->mstats, c->d.rank, c->d.world);
  }
  CU(cudaGetLastError());
  if (c->d.world > 1) {
    // each location was written by exactly one rank (zeros elsewhere): the sum is exact
    NC(g_nccl.AllReduce(c->mloc, c->mloc, (size_t)n * 3, ncclDouble, ncclSum, c->comm, c->stream));
    LAUNCH(c, k_marker_cells, grid_for(c, n), 256, c->g, c->mloc, n, (int)c->d.h, c->d.origin[0],
           c->d.origin[1], c->d.origin[2], c->mnew, c->mstats, 0);
    CU(cudaGetLastError());
  }
  MarkerStats st;
  if ((rc = read_marker_stats(c, &st))) return rc;
  if (moved) *moved = (int64_t)st.moved;
  if (st.outside)   // Grid.raw_get on a coordinate the box cannot hold
    return fail(c, SPL_E_MARKER, "%llu marker(s) left the dense box", st.outside);
  if ((c->d.quirks & SPL_Q_MOVE_CELLS) && st.moved) {
    double* p = reinterpret_cast<double*>(c->p);
    if (c->d.dtype == SPL_F64)
      LAUNCH(c, k_move_records<double>, grid_for(c, n), 256, c->g, c->media,
             Impl<double>::P(c->u), Impl<double>::P(c->v), Impl<double>::P(c->w), p, c->mold,
             c->mnew, n);
    else
      LAUNCH(c, k_move_records<float>, grid_for(c, n), 256, c->g, c->media, Impl<float>::P(c->u),
             Impl<float>::P(c->v), Impl<float>::P(c->w), p, c->mold, c->mnew, n);
    CU(cudaGetLastError());
    c->have_rhs = false;
  }
  return SPL_OK;
}

int spl_relabel(spl_ctx* c, int sanity) {
  if (!c) return SPL_E_BADARG;
  int rc = markers_need(c, "spl_relabel");
  if (rc) return rc;
  CU(cudaSetDevice(c->d.device));
  const Geo& g = c->g;
  const int grid = grid_for(c, g.ncell);
  const int lazy = (c->d.quirks & SPL_Q_LAZY_BUFFER) ? 1 : 0;
  CU(cudaMemsetAsync(c->mark, 0, (size_t)g.ncell, c->stream));
  CU(cudaMemsetAsync(c->mstats, 0, sizeof(MarkerStats), c->stream));
  LAUNCH(c, k_mark_cells, grid_for(c, c->nmark), 256, c->mnew, c->nmark, c->mark);
  LAUNCH(c, k_relabel_init, grid, 256, g, c->world, c->media, c->layer, c->mark, c->mstats);
  for (int step = 1; step <= 2; ++step) {   // Build.max_distance = max 2 (ceil time_step_constant)
    // slabs: the sources of a ring may sit in the neighbour's boundary plane
    if ((rc = exchange_halo(c, reinterpret_cast<char*>(c->layer), 1, ncclUint8, 1))) return rc;
    if ((rc = exchange_halo(c, reinterpret_cast<char*>(c->media), 1, ncclUint8, 1))) return rc;
    if (lazy && (rc = exchange_halo(c, reinterpret_cast<char*>(c->mark), 1, ncclUint8, 1))) return rc;
    LAUNCH(c, k_relabel_step, grid, 256, g, c->world, c->media, c->layer, c->mark, step, lazy);
  }
  double* p = reinterpret_cast<double*>(c->p);
  if (c->d.dtype == SPL_F64)
    LAUNCH(c, k_relabel_finish<double>, grid, 256, g, c->media, c->layer, c->mark,
           Impl<double>::P(c->u), Impl<double>::P(c->v), Impl<double>::P(c->w), p);
  else
    LAUNCH(c, k_relabel_finish<float>, grid, 256, g, c->media, c->layer, c->mark,
           Impl<float>::P(c->u), Impl<float>::P(c->v), Impl<float>::P(c->w), p);
  CU(cudaGetLastError());
  if ((rc = exchange_halo(c, reinterpret_cast<char*>(c->media), 1, ncclUint8, g.hz))) return rc;
  if ((rc = rebuild_operators(c))) return rc;
  c->have_rhs = false;
  MarkerStats st;
  if (sanity) {
    unsigned long long* missing = &c->mstats->moved;   // reuse: zero since the memset above
    LAUNCH(c, k_check_surroundings, grid_for(c, c->nmark), 256, g, c->media, c->mnew, c->nmark,
           missing);
  }
  if ((rc = read_marker_stats(c, &st))) return rc;
  if (sanity && st.escaped)      // Build.fs:162-166
    return fail(c, SPL_E_MARKER, "Error: Fluid marker went outside bounds (%llu marker cells).",
                st.escaped);
  if (sanity && st.moved)        // Build.fs:168-174
    return fail(c, SPL_E_MARKER, "Fluid marker does not have a neighbor. (%llu markers)",
                st.moved);
  return SPL_OK;
}

int spl_download_media(spl_ctx* c, uint8_t* media) {
  if (!c || !media) return SPL_E_BADARG;
  NEED_UPLOAD(c, "spl_download_media");
  CU(cudaSetDevice(c->d.device));
  CU(cudaMemcpyAsync(media, c->media, (size_t)c->g.ncell, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SPL_OK;
}

int spl_advance(spl_ctx* c, double dt, int sanity, spl_info* info) {
  if (!c) return SPL_E_BADARG;
  NEED_UPLOAD(c, "spl_advance");
  if (!(dt > 0.0)) return fail(c, SPL_E_BADARG, "dt must be positive");
  CU(cudaSetDevice(c->d.device));
  int rc = SPL_OK;
  if (c->nmark > 0) {   // Simulator.fs:60-79: move markers, rebuild the cell labels
    if ((rc = spl_move_markers(c, dt, nullptr))) return rc;
    if ((rc = spl_relabel(c, sanity))) return rc;
  }
  // Simulator.fs:82-86: convection, forces, viscosity (advect's error checks need the
  // escape counter, so it goes through the public entry point)
  spl_info adv;
  rc = spl_advect(c, dt, &adv);
  if (rc) return rc;
  if ((rc = DISPATCH(c, Impl<float>::forces(c, dt), Impl<double>::forces(c, dt)))) return rc;
  if ((rc = DISPATCH(c, Impl<float>::viscosity(c, dt), Impl<double>::viscosity(c, dt)))) return rc;
  // Simulator.fs:88-89 (+ the validators of :91-97 on the device)
  spl_info local;
  memset(&local, 0, sizeof local);
  rc = project_impl(c, dt, &local, false);
  local.ms_advect = adv.ms_advect;
  local.ms_total += adv.ms_advect;
  local.trace_escapes = adv.trace_escapes;
  local.kernel_launches += adv.kernel_launches + 2;
  if (info) *info = local;
  if (rc) return rc;
  if (sanity) {
    double nf = 0.0;   // per-rank count: agree before anybody leaves (cleanup has collectives)
    if ((rc = agree_max(c, local.nonfinite ? 1.0 : 0.0, &nf))) return rc;
    if (nf > 0.0) return fail(c, SPL_E_NONFINITE, "Invalid pressure.");
    if (local.max_abs_div > 0.0001)   // Pressure.fs:148
      return fail(c, SPL_E_DIVERGENCE, "Velocity field is not correct: divergence %g.",
                  local.max_abs_div);
  }
  // Simulator.fs:100-110
  if ((rc = DISPATCH(c, Impl<float>::cleanup(c), Impl<double>::cleanup(c)))) return rc;
  CU(cudaStreamSynchronize(c->stream));
  return SPL_OK;
}

int spl_check(spl_ctx* c, double div_tol, spl_info* info) {
  if (!c) return SPL_E_BADARG;
  NEED_UPLOAD(c, "spl_check");
  CU(cudaSetDevice(c->d.device));
  c->launches = 0;
  CU(cudaMemsetAsync(&c->sc->nonfinite, 0, sizeof(unsigned long long), c->stream));
  int rc = DISPATCH(c, Impl<float>::check(c), Impl<double>::check(c));
  if (rc) return rc;
  spl_info local;
  memset(&local, 0, sizeof local);
  if ((rc = read_back(c, &local, false, true))) return rc;
  if (info) *info = local;
  if (local.nonfinite)   // Vector.fs:17, Pressure.fs:138-142
    return fail(c, SPL_E_NONFINITE,
                "Vector3d has an invalid quantity: either NaN or ±Infinity. (%lld entries)",
                (long long)local.nonfinite);
  if (local.max_abs_div > div_tol)   // Pressure.fs:149
    return fail(c, SPL_E_DIVERGENCE, "Velocity field is not correct: divergence %g.",
                local.max_abs_div);
  return SPL_OK;
}

void* spl_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}

void spl_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int spl_snapshot(spl_ctx* c) {
  if (!c) return SPL_E_BADARG;
  NEED_UPLOAD(c, "spl_snapshot");
  CU(cudaSetDevice(c->d.device));
  if (!c->have_snapshot) {
    int rc;
    if ((rc = alloc_field(c, 15, c->esz * 3, &c->su, 0))) return rc;
    c->have_snapshot = true;
  }
  const size_t nb = (size_t)c->g.ncell * c->esz;
  const size_t stride = field_bytes(c, c->esz);
  char* base = static_cast<char*>(c->alloc[15]);
  CU(cudaMemcpyAsync(base, c->u, nb, cudaMemcpyDeviceToDevice, c->stream));
  CU(cudaMemcpyAsync(base + stride, c->v, nb, cudaMemcpyDeviceToDevice, c->stream));
  CU(cudaMemcpyAsync(base + 2 * stride, c->w, nb, cudaMemcpyDeviceToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SPL_OK;
}

int spl_restore(spl_ctx* c) {
  if (!c) return SPL_E_BADARG;
  if (!c->have_snapshot) return fail(c, SPL_E_STATE, "spl_restore before spl_snapshot");
  CU(cudaSetDevice(c->d.device));
  const size_t nb = (size_t)c->g.ncell * c->esz;
  const size_t stride = field_bytes(c, c->esz);
  char* base = static_cast<char*>(c->alloc[15]);
  CU(cudaMemcpyAsync(c->u, base, nb, cudaMemcpyDeviceToDevice, c->stream));
  CU(cudaMemcpyAsync(c->v, base + stride, nb, cudaMemcpyDeviceToDevice, c->stream));
  CU(cudaMemcpyAsync(c->w, base + 2 * stride, nb, cudaMemcpyDeviceToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->have_rhs = false;
  return SPL_OK;
}

int spl_timer_start(spl_ctx* c) {
  if (!c) return SPL_E_BADARG;
  CU(cudaSetDevice(c->d.device));
  CU(cudaEventRecord(c->tev[0], c->stream));
  return SPL_OK;
}

int spl_timer_stop(spl_ctx* c, float* ms) {
  if (!c || !ms) return SPL_E_BADARG;
  CU(cudaSetDevice(c->d.device));
  CU(cudaEventRecord(c->tev[1], c->stream));
  CU(cudaEventSynchronize(c->tev[1]));
  CU(cudaEventElapsedTime(ms, c->tev[0], c->tev[1]));
  return SPL_OK;
}

int spl_apply_preconditioner(spl_ctx* c, const void* r, void* z, float* ms) {
  if (!c || !r || !z) return SPL_E_BADARG;
  NEED_UPLOAD(c, "spl_apply_preconditioner");
  if (c->d.precond != SPL_PC_RBIC0 && c->d.precond != SPL_PC_MG)
    return fail(c, SPL_E_BADARG, "preconditioner %d has no stored z", c->d.precond);
  if (c->d.world > 1) return fail(c, SPL_E_BADARG, "spl_apply_preconditioner needs world = 1");
  CU(cudaSetDevice(c->d.device));
  const size_t bytes = (size_t)c->g.ncell * c->esz;
  CU(cudaMemcpyAsync(c->r, r, bytes, cudaMemcpyHostToDevice, c->stream));
  Scal* h = c->hsc;
  memset(h, 0, sizeof(Scal));
  h->world = 1;
  CU(cudaMemcpyAsync(c->sc, h, sizeof(Scal), cudaMemcpyHostToDevice, c->stream));
  CU(cudaEventRecord(c->tev[0], c->stream));
  int prc = (c->d.dtype == SPL_F64) ? Impl<double>::precondition(c, 0)
                                    : Impl<float>::precondition(c, 0);
  if (prc) return prc;
  CU(cudaEventRecord(c->tev[1], c->stream));
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(z, c->z, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (ms) CU(cudaEventElapsedTime(ms, c->tev[0], c->tev[1]));
  c->have_rhs = false;      // r was overwritten
  return SPL_OK;
}

int spl_profile_pcg(spl_ctx* c, double dt, int iters, float* ms_a, float* ms_b, float* ms_x) {
  if (!c || !ms_a || !ms_b || !ms_x) return SPL_E_BADARG;
  NEED_UPLOAD(c, "spl_profile_pcg");
  if (iters < 1 || iters > 256) return fail(c, SPL_E_BADARG, "iters must be in [1, 256]");
  if (!(dt > 0.0)) return fail(c, SPL_E_BADARG, "dt must be positive");
  CU(cudaSetDevice(c->d.device));
  int rc = do_rhs(c, dt);
  if (rc) return rc;
  const int n_ev = 4 * iters;
  cudaEvent_t* evs = new cudaEvent_t[n_ev];
  for (int i = 0; i < n_ev; ++i) CU(cudaEventCreate(&evs[i]));
  for (auto& e : c->xev)
    if (!e) CU(cudaEventCreate(&e));
  c->profile_flushes = 0;
  c->pev = evs;
  c->profile_iters = iters;
  const int keep_fixed = c->fixed_iters;
  c->fixed_iters = iters;
  rc = DISPATCH(c, Impl<float>::pcg_dispatch(c), Impl<double>::pcg_dispatch(c));
  c->fixed_iters = keep_fixed;
  c->profile_iters = 0;
  c->pev = nullptr;
  if (rc == SPL_OK) {
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) rc = SPL_E_CUDA;
  }
  double a = 0.0, b = 0.0;
  if (rc == SPL_OK) {
    for (int i = 0; i < iters; ++i) {
      float t = 0.f;
      cudaEventElapsedTime(&t, evs[3 * i], evs[3 * i + 1]);
      a += t;
      cudaEventElapsedTime(&t, evs[3 * i + 2], evs[3 * iters + i]);
      b += t;
    }
    *ms_a = (float)(a / iters);
    *ms_b = (float)(b / iters);
    if (getenv("SPL_PROFILE_VERBOSE") && iters > 1) {   // whole iteration period, start to start
      float t = 0.f;
      cudaEventElapsedTime(&t, evs[0], evs[3 * (iters - 1)]);
      fprintf(stderr, "[spl_profile_pcg] rank %d: iteration period %.1f us (phase A %.1f, B %.1f)\n",
              c->d.rank, 1e3 * t / (iters - 1), 1e3 * a / iters, 1e3 * b / iters);
    }
    double xs = 0.0;
    for (int i = 0; i < c->profile_flushes; ++i) {
      float t = 0.f;
      cudaEventElapsedTime(&t, c->xev[2 * i], c->xev[2 * i + 1]);
      xs += t;
    }
    *ms_x = c->profile_flushes ? (float)(xs / c->profile_flushes) : 0.f;
  }
  for (int i = 0; i < n_ev; ++i) cudaEventDestroy(evs[i]);
  delete[] evs;
  c->have_rhs = false;
  return rc;
}

int spl_set_refinement(spl_ctx* c, int rounds) {
  if (!c || rounds < 0 || rounds > 8) return SPL_E_BADARG;
  c->refine_rounds = rounds;
  return SPL_OK;
}

int spl_set_fixed_iters(spl_ctx* c, int iters) {
  if (!c) return SPL_E_BADARG;
  c->fixed_iters = iters > 0 ? iters : 0;
  return SPL_OK;
}

void* spl_device_ptr(spl_ctx* c, int which, int* halo_planes) {
  if (!c) return nullptr;
  if (halo_planes) *halo_planes = c->g.hz;
  switch (which) {
    case 0: return c->u - (size_t)c->g.hz * c->g.plane * c->esz;
    case 1: return c->v - (size_t)c->g.hz * c->g.plane * c->esz;
    case 2: return c->w - (size_t)c->g.hz * c->g.plane * c->esz;
    case 3: return c->p - (size_t)c->g.hz * c->g.plane * 8;   // always fp64
    case 4: return c->r - (size_t)c->g.hz * c->g.plane * c->esz;
    case 5: return c->media - (size_t)c->g.hz * c->g.plane;
    case 6: return c->codes - (size_t)c->g.hz * c->g.plane;
    default: return nullptr;
  }
}

}  // extern "C"
