// comm.cu -- communicator of the column-sharded large-n mode: one process per GPU, NCCL over
// NVLink/NVSwitch.  Exchanged, and nothing else: the m x m Gram partials (once per outer iteration), the m-vector
// t = sum_g J_g v_g (once per projection / pcg matvec), row sums of c(x), and the packed CG / line-search scalars.
// J and every n-vector never leave their GPU.  NCCL is bound at run time (dlopen) so that the process uses the same
// libnccl.so.2 torch.distributed already loaded; without a communicator every hook is a no-op.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "ctx.h"
#include "large_state.h"
#include "large_device.cuh"

using namespace lfpsqp;

namespace {
typedef int ncclResult_t;
typedef void *ncclComm_t;
struct ncclUniqueId { char internal[128]; };
enum { ncclSum = 0, ncclMax = 2, ncclFloat64 = 8 };
struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

bool load_nccl(const char *path, std::string &err) {
  if (g_nccl.h) return true;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // already loaded by torch.distributed?
  if (!h && path && path[0]) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { err = std::string("cannot load NCCL: ") + dlerror(); return false; }
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) { err = "NCCL symbols missing"; return false; }
  g_nccl.h = h;
  return true;
}

// out[0..cnt) = for each masked slot k (ascending): sum / max of part[k][0..np)   (single CTA)
__global__ void __launch_bounds__(256) pack_slots_kernel(const double *part, int np, unsigned mask, int domax, double *out) {
  __shared__ double sh[33];
  int o = 0;
  for (int k = 0; k < NSLOT; k++) {
    if (!(mask & (1u << k))) continue;
    double s = 0.0;
    const double *p = part + (size_t)k * MAXP;
    if (domax) { for (int i = threadIdx.x; i < np; i += 256) { double w = p[i]; s = (w > s || isnan(w)) ? w : s; } }
    else for (int i = threadIdx.x; i < np; i += 256) s += p[i];
    // block reduce
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int d = 16; d > 0; d >>= 1) { double t = __shfl_xor_sync(0xffffffffu, s, d); s = domax ? ((t > s || isnan(t)) ? t : s) : s + t; }
    __syncthreads();
    if (lane == 0) sh[w] = s;
    __syncthreads();
    double r = (lane < 8) ? sh[lane] : 0.0;
    for (int d = 16; d > 0; d >>= 1) { double t = __shfl_xor_sync(0xffffffffu, r, d); r = domax ? ((t > r || isnan(t)) ? t : r) : r + t; }
    if (threadIdx.x == 0) out[o] = r;
    o++;
    __syncthreads();
  }
}
// scatter the reduced values back to part[k][0] (consumers then read np = 1 partial)
__global__ void unpack_slots_kernel(double *part, unsigned mask, const double *in) {
  int o = 0;
  for (int k = 0; k < NSLOT; k++) if (mask & (1u << k)) { part[(size_t)k * MAXP] = in[o]; o++; }
}
__global__ void pack_scalars_kernel(const LargeCtrl *ctrl, unsigned mask, double *out) {
  int o = 0;
  for (int k = 0; k < 16; k++) if (mask & (1u << k)) { out[o] = ctrl->s[k]; o++; }
}
__global__ void unpack_scalars_kernel(LargeCtrl *ctrl, unsigned mask, const double *in) {
  int o = 0;
  for (int k = 0; k < 16; k++) if (mask & (1u << k)) { ctrl->s[k] = in[o]; o++; }
}

void ar(LargeState &S, double *buf, size_t count, int op) {
  if (S.world <= 1 || !S.comm || !S.comm->nccl) return;
  ncclResult_t r = g_nccl.AllReduce(buf, buf, count, ncclFloat64, op, (ncclComm_t)S.comm->nccl, S.stream);
  if (r != 0) fprintf(stderr, "lfpsqp: ncclAllReduce failed: %s\n", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  S.collectives++;
}

// ------------------------------------------------------------------ peer-memory all-reduce (one kernel, no NCCL)
// Region layout per rank (doubles): data[2][PC_MAX + PC_SCAL] then flags[PC_RANKS][PC_COLS] (unsigned long long).
// Protocol per call (epoch e, parity e&1): every CTA (1) writes its slice of the local contribution into its OWN
// region, (2) __threadfence_system, (3) stores e into flags[my_rank][cta] of EVERY peer's region (remote stores over
// NVLink), (4) spins until flags[r][cta] >= e for all r in its own region, (5) sums the slices of all ranks in rank
// order through the mapped peer pointers (remote loads) -- the same order everywhere, so every rank gets the
// bitwise-identical result.  A region is reused two calls later; a peer's flag for call e+1 implies that its reads of
// call e are complete (stream order), so the parity double buffer is race-free.
// (layout constants PC_* / FZ_* live in large_ctrl.h: the fused projcg kernel shares the exported region)

struct PeerPtrs { double *r[PC_RANKS]; };

__device__ __forceinline__ unsigned long long *pc_flags(double *region) { return reinterpret_cast<unsigned long long *>(region + PC_DATA); }
__device__ __forceinline__ double ld_sys(const double *p) { double v; asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }

__device__ __forceinline__ bool pc_exchange(const PeerPtrs &P, int rank, int world, unsigned long long epoch, int col, LargeCtrl *ctrl) {
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world) {
    unsigned long long *f = pc_flags(P.r[threadIdx.x]) + (size_t)rank * PC_COLS + col;   // my flag slot in peer threadIdx.x
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
  }
  bool ok = true;
  if ((int)threadIdx.x < world) {
    const unsigned long long *f = pc_flags(P.r[rank]) + (size_t)threadIdx.x * PC_COLS + col;  // peer's flag in MY region
    long long t0 = clock64();
    unsigned long long v;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v >= epoch) break;
      if (clock64() - t0 > 8000000000LL) { ok = false; ctrl->commfail = 1; break; }   // ~4 s: a peer died; do not hang the GPU (the host reports LFPSQP_ERR_COMM)
    } while (true);
  }
  __syncthreads();
  return ok;
}

// vector flavour: dst[i] = sum_r src_r[i], i < count <= PC_MAX ; CTA c handles [c*1024, c*1024+1024)
__global__ void __launch_bounds__(256) peer_allreduce_vec_kernel(PeerPtrs P, int rank, int world, unsigned long long epoch,
                                                                 const double *src, double *dst, int count, LargeCtrl *ctrl) {
  const int par = (int)(epoch & 1), base = blockIdx.x * 1024;
  double *mine = P.r[rank] + (size_t)par * (PC_MAX + PC_SCAL);
  for (int i = base + threadIdx.x; i < min(count, base + 1024); i += 256) mine[i] = src[i];
  pc_exchange(P, rank, world, epoch, blockIdx.x, ctrl);
  for (int i = base + threadIdx.x; i < min(count, base + 1024); i += 256) {
    double s = 0.0;
    for (int r = 0; r < world; r++) s += ld_sys(P.r[r] + (size_t)par * (PC_MAX + PC_SCAL) + i);
    dst[i] = s;
  }
}

// scalar flavour (one CTA): for every masked slot k: v_k = reduce(part[k][0..np)) locally (sum or max), all-reduced,
// and written back to part[k][0] (consumers then read np = 1) and, if ctrl_out, to ctrl->s[k].
// from_ctrl: take the local value from ctrl->s[k] instead of reducing partials.
__global__ void __launch_bounds__(256) peer_allreduce_slots_kernel(PeerPtrs P, int rank, int world, unsigned long long epoch,
                                                                   double *part, int np, unsigned mask, int domax, int from_ctrl,
                                                                   int ctrl_out, LargeCtrl *ctrl) {
  __shared__ double sh[33];
  const int par = (int)(epoch & 1);
  double *mine = P.r[rank] + (size_t)par * (PC_MAX + PC_SCAL) + PC_MAX;
  for (int k = 0; k < 16; k++) {
    if (!(mask & (1u << k))) continue;
    double r;
    if (from_ctrl) r = ctrl->s[k];
    else r = domax ? reduce_partials_max(part + (size_t)k * MAXP, np, sh) : reduce_partials(part + (size_t)k * MAXP, np, sh);
    if (threadIdx.x == 0) mine[k] = r;
    __syncthreads();
  }
  pc_exchange(P, rank, world, epoch, PC_COLS - 1, ctrl);
  if (threadIdx.x < 16 && (mask & (1u << threadIdx.x))) {
    const int k = threadIdx.x;
    double s = domax ? 0.0 : 0.0;
    for (int r = 0; r < world; r++) {
      double v = ld_sys(P.r[r] + (size_t)par * (PC_MAX + PC_SCAL) + PC_MAX + k);
      s = domax ? ((v > s || isnan(v)) ? v : s) : s + v;
    }
    if (!from_ctrl) part[(size_t)k * MAXP] = s;
    if (ctrl_out || from_ctrl) ctrl->s[k] = s;
  }
}

PeerPtrs peer_ptrs(LargeState &S) { PeerPtrs P; for (int r = 0; r < PC_RANKS; r++) P.r[r] = S.comm->peer_map[r]; return P; }
bool use_peer(LargeState &S, size_t count) { return S.comm && S.comm->peer_ready && count <= (size_t)PC_MAX; }
}  // namespace

void comm_allreduce(LargeState &S, double *buf, size_t count) {
  if (S.world <= 1) return;
  if (use_peer(S, count)) {
    unsigned long long e = ++S.comm->epoch;
    peer_allreduce_vec_kernel<<<(unsigned)((count + 1023) / 1024), 256, 0, S.stream>>>(peer_ptrs(S), S.rank, S.world, e, buf, buf, (int)count, S.ctrl);
    S.launches++; S.collectives++;
    return;
  }
  ar(S, buf, count, ncclSum);
}

void comm_allreduce_scalars(LargeState &S, unsigned summask, unsigned maxmask) {
  if (S.world <= 1) return;
  if (use_peer(S, 1)) {
    if (summask) { unsigned long long e = ++S.comm->epoch; peer_allreduce_slots_kernel<<<1, 256, 0, S.stream>>>(peer_ptrs(S), S.rank, S.world, e, nullptr, 0, summask, 0, 1, 1, S.ctrl); }
    if (maxmask) { unsigned long long e = ++S.comm->epoch; peer_allreduce_slots_kernel<<<1, 256, 0, S.stream>>>(peer_ptrs(S), S.rank, S.world, e, nullptr, 0, maxmask, 1, 1, 1, S.ctrl); }
    S.launches += 2; S.collectives += 2;
    return;
  }
  if (summask) {
    pack_scalars_kernel<<<1, 1, 0, S.stream>>>(S.ctrl, summask, S.commbuf);
    ar(S, S.commbuf, __builtin_popcount(summask), ncclSum);
    unpack_scalars_kernel<<<1, 1, 0, S.stream>>>(S.ctrl, summask, S.commbuf);
  }
  if (maxmask) {
    pack_scalars_kernel<<<1, 1, 0, S.stream>>>(S.ctrl, maxmask, S.commbuf + 16);
    ar(S, S.commbuf + 16, __builtin_popcount(maxmask), ncclMax);
    unpack_scalars_kernel<<<1, 1, 0, S.stream>>>(S.ctrl, maxmask, S.commbuf + 16);
  }
  S.launches += 4;
}

static void slots_allreduce(LargeState &S, unsigned mask, int np) {
  if (use_peer(S, 1)) {
    unsigned long long e = ++S.comm->epoch;
    peer_allreduce_slots_kernel<<<1, 256, 0, S.stream>>>(peer_ptrs(S), S.rank, S.world, e, S.lp, np, mask, 0, 0, 0, S.ctrl);
    S.launches++; S.collectives++;
    return;
  }
  pack_slots_kernel<<<1, 256, 0, S.stream>>>(S.lp, np, mask, 0, S.commbuf + 32);
  ar(S, S.commbuf + 32, __builtin_popcount(mask), ncclSum);
  unpack_slots_kernel<<<1, 1, 0, S.stream>>>(S.lp, mask, S.commbuf + 32);
  S.launches += 2;
}
void comm_allreduce_loop_slot(LargeState &S, int slot, int np) {
  if (S.world <= 1) return;
  slots_allreduce(S, 1u << slot, np);
}
void comm_allreduce_loop_slots_cg(LargeState &S, int par) {
  if (S.world <= 1) return;
  slots_allreduce(S, (1u << 1) | (1u << (2 + par)), S.np_loop_raw);
}

void comm_release(LargeState &) {}

// ---- CUDA IPC plumbing of the peer-memory all-reduce: every rank exports one region, maps everybody else's
extern "C" int lfpsqp_comm_ipc_export(lfpsqp_ctx *c, void *handle64_out) {
  if (!c || !handle64_out) return LFPSQP_ERR_ARG;
  cudaSetDevice(c->device);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!c->comm.peer_local) {
    if (cudaMalloc((void **)&c->comm.peer_local, PC_REGION_BYTES) != cudaSuccess) { cudaGetLastError(); return c->fail(LFPSQP_ERR_NOMEM, "peer region allocation failed"); }
    cudaMemset(c->comm.peer_local, 0, PC_REGION_BYTES);
    cudaDeviceSynchronize();
  }
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, c->comm.peer_local);
  if (e != cudaSuccess) return c->cuda_fail(e, "cudaIpcGetMemHandle");
  memcpy(handle64_out, &h, 64);
  return LFPSQP_OK;
}
// handles: world x 64 bytes, in rank order.  Call after every rank exported and after lfpsqp_comm_init.
extern "C" int lfpsqp_comm_ipc_import(lfpsqp_ctx *c, const void *handles) {
  if (!c || !handles) return LFPSQP_ERR_ARG;
  if (c->comm.world < 2 || c->comm.world > PC_RANKS || !c->comm.peer_local) return c->fail(LFPSQP_ERR_ARG, "peer all-reduce needs 2..8 ranks and an exported region");
  cudaSetDevice(c->device);
  for (int r = 0; r < c->comm.world; r++) {
    if (r == c->comm.rank) { c->comm.peer_map[r] = c->comm.peer_local; continue; }
    cudaIpcMemHandle_t h; memcpy(&h, (const char *)handles + 64 * r, 64);
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { c->comm.peer_ready = false; return c->cuda_fail(e, "cudaIpcOpenMemHandle (peer-memory all-reduce unavailable; NCCL is used instead)"); }
    c->comm.peer_map[r] = (double *)p;
  }
  // the epoch stays monotone over re-imports: the exported region keeps the flags of earlier epochs
  c->comm.peer_ready = true;
  return LFPSQP_OK;
}
// 0 = single GPU, 1 = NCCL only, 2 = peer-memory kernels for small messages + NCCL for the Gram
extern "C" int lfpsqp_comm_mode(lfpsqp_ctx *c) {
  if (!c || c->comm.world <= 1) return 0;
  return c->comm.peer_ready ? 2 : 1;
}

extern "C" int lfpsqp_comm_unique_id(void *out128, const char *nccl_lib_path) {
  std::string err;
  if (!out128 || !load_nccl(nccl_lib_path, err)) { fprintf(stderr, "lfpsqp_comm_unique_id: %s\n", err.c_str()); return LFPSQP_ERR_COMM; }
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != 0) return LFPSQP_ERR_COMM;
  memcpy(out128, &id, 128);
  return LFPSQP_OK;
}

extern "C" int lfpsqp_comm_init(lfpsqp_ctx *c, int rank, int world, const void *unique_id128, const char *nccl_lib_path) {
  if (!c || !unique_id128 || world < 1 || rank < 0 || rank >= world) return LFPSQP_ERR_ARG;
  std::string err;
  if (!load_nccl(nccl_lib_path, err)) return c->fail(LFPSQP_ERR_COMM, "%s", err.c_str());
  cudaSetDevice(c->device);
  if (c->comm.nccl) { g_nccl.CommDestroy((ncclComm_t)c->comm.nccl); c->comm.nccl = nullptr; }
  ncclUniqueId id; memcpy(&id, unique_id128, 128);
  ncclComm_t comm = nullptr;
  ncclResult_t r = g_nccl.CommInitRank(&comm, world, id, rank);
  if (r != 0) return c->fail(LFPSQP_ERR_COMM, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  c->comm.nccl = comm; c->comm.rank = rank; c->comm.world = world;
  return LFPSQP_OK;
}

extern "C" int lfpsqp_comm_destroy(lfpsqp_ctx *c) {
  if (!c) return LFPSQP_ERR_ARG;
  cudaSetDevice(c->device);
  for (int r = 0; r < 8; r++) if (c->comm.peer_map[r] && c->comm.peer_map[r] != c->comm.peer_local) cudaIpcCloseMemHandle(c->comm.peer_map[r]);
  if (c->comm.peer_local) cudaFree(c->comm.peer_local);
  if (c->comm.nccl && g_nccl.CommDestroy) { cudaSetDevice(c->device); g_nccl.CommDestroy((ncclComm_t)c->comm.nccl); }
  c->comm = CommState();
  return LFPSQP_OK;
}
