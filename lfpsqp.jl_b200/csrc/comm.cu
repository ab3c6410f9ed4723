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
}  // namespace

void comm_allreduce(LargeState &S, double *buf, size_t count) { ar(S, buf, count, ncclSum); }

void comm_allreduce_scalars(LargeState &S, unsigned summask, unsigned maxmask) {
  if (S.world <= 1) return;
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

void comm_allreduce_loop_slot(LargeState &S, int slot, int np) {
  if (S.world <= 1) return;
  pack_slots_kernel<<<1, 256, 0, S.stream>>>(S.lp, np, 1u << slot, 0, S.commbuf + 32);
  ar(S, S.commbuf + 32, 1, ncclSum);
  unpack_slots_kernel<<<1, 1, 0, S.stream>>>(S.lp, 1u << slot, S.commbuf + 32);
  S.launches += 2;
}

void comm_allreduce_loop_slots_cg(LargeState &S, int par) {
  if (S.world <= 1) return;
  unsigned mask = (1u << 1) | (1u << (2 + par));
  pack_slots_kernel<<<1, 256, 0, S.stream>>>(S.lp, S.np_loop_raw, mask, 0, S.commbuf + 40);
  ar(S, S.commbuf + 40, 2, ncclSum);
  unpack_slots_kernel<<<1, 1, 0, S.stream>>>(S.lp, mask, S.commbuf + 40);
  S.launches += 2;
}

void comm_release(LargeState &) {}

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
  if (c->comm.nccl && g_nccl.CommDestroy) { cudaSetDevice(c->device); g_nccl.CommDestroy((ncclComm_t)c->comm.nccl); }
  c->comm = CommState();
  return LFPSQP_OK;
}
