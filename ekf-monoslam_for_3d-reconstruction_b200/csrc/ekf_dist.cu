// csrc/ekf_dist.cu — NCCL plumbing of the row-block partitioned large-map update (BASELINE config 4,
// SURVEY.md §8(e)).  One process per GPU; every rank owns an ekf_handle holding a replica of the filter
// and attaches it to a communicator created here from an id that the host distributes (the Python host
// broadcasts it with torch.distributed).  libnccl is opened at run time (the copy torch already
// loaded), so libekf_b200.so has no link-time NCCL dependency and single-GPU users never touch it.
// The data path: ekf_api.cu::stacked_update updates only the rank's rows of Sigma and calls
// ekf_dist_allgather_rows for the W_b / V_b panels (n x 128 fp64 per 64 features), delta, and once per
// stacked update the row blocks of Sigma.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/ekf_b200.h"
#include "ekf_handle.h"

// the handful of NCCL declarations used (nccl.h, ABI-stable since NCCL 2.0)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;     // ncclSuccess = 0
enum { kNcclDouble = 8 };     // ncclFloat64
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;   // op 0 = ncclSum
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_fail(ekf_handle* h, ncclResult_t r, const char* what) {
  char buf[256];
  snprintf(buf, sizeof buf, "NCCL error %d (%s) in %s", r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", what);
  if (h) h->err = buf;
  return EKF_ERR_CUDA;
}

extern "C" {

static int p2p_setup(ekf_handle* h);

int ekf_dist_load_nccl(const char* path) {
  if (g_nccl.lib) return EKF_OK;
  void* lib = dlopen(path && path[0] ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { fprintf(stderr, "ekf_dist_load_nccl: %s\n", dlerror()); return EKF_ERR_UNSUPPORTED; }
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(lib, "ncclCommInitRank");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(lib, "ncclCommDestroy");
  g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(lib, "ncclAllGather");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(lib, "ncclAllReduce");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather || !g_nccl.AllReduce) return EKF_ERR_UNSUPPORTED;
  g_nccl.lib = lib;
  return EKF_OK;
}

int ekf_dist_unique_id(char out[128]) {
  if (!g_nccl.lib || !out) return EKF_ERR_STATE;
  ncclUniqueId id;
  const ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != 0) return nccl_fail(nullptr, r, "ncclGetUniqueId");
  memcpy(out, id.internal, 128);
  return EKF_OK;
}

int ekf_dist_attach(ekf_handle* h, const char id_bytes[128], int rank, int world) {
  if (!h || !id_bytes || world < 1 || rank < 0 || rank >= world) return EKF_ERR_ARG;
  if (!g_nccl.lib) return EKF_ERR_STATE;
  if (h->nccl_comm) return EKF_ERR_STATE;
  if (cudaSetDevice(h->device) != cudaSuccess) return EKF_ERR_CUDA;
  ncclUniqueId id;
  memcpy(id.internal, id_bytes, 128);
  ncclComm_t comm = nullptr;
  const ncclResult_t r = g_nccl.CommInitRank(&comm, world, id, rank);
  if (r != 0) return nccl_fail(h, r, "ncclCommInitRank");
  h->nccl_comm = comm; h->rank = rank; h->world = world;
  return p2p_setup(h);
}

// ---- peer-memory exchange ------------------------------------------------------------------------
// Every rank exports its three W / V panel buffers and one small allocation (partial-S slots + flag words) with CUDA IPC;
// the handles travel once over the fresh communicator; afterwards the producing kernels store straight into the peers
// (ekf_update.cu: k_blk_S / k_blk_V with a P2PView) and no NCCL call remains inside a block of the partitioned
// look-ahead update.  EKF_DIST_P2P=0 keeps the NCCL collectives.
typedef int (*cuMemGetAddressRange_t)(unsigned long long*, size_t*, unsigned long long);
struct P2PExport { cudaIpcMemHandle_t hnd; unsigned long long off; unsigned long long valid; };   // 80 bytes

static void p2p_teardown(ekf_handle* h);

// Every rank ALWAYS takes part in the two collectives below (the handle all-gather and the agreement all-reduce), whatever
// failed locally: a rank that cannot export sends records marked invalid, a rank that cannot map a peer votes 0.  Peer
// memory is switched on only when every rank voted 1; otherwise every rank closes what it opened and the update keeps the
// NCCL exchange on ALL ranks (mixed paths would hang: one side spinning on flag words, the other inside ncclAllReduce).
static int p2p_setup(ekf_handle* h) {
  const char* e = getenv("EKF_DIST_P2P");
  if ((e && atoi(e) == 0) || h->world > 8 || h->world < 2) return EKF_OK;   // same environment / world on every rank
  ncclComm_t comm = (ncclComm_t)h->nccl_comm;
  int ok = 1;
  void* cu = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  cuMemGetAddressRange_t range = cu ? (cuMemGetAddressRange_t)dlsym(cu, "cuMemGetAddressRange_v2") : nullptr;
  if (!range) ok = 0;
  const size_t spart_doubles = (size_t)h->world * EKF_UB * EKF_UB;
  const size_t xs_bytes = spart_doubles * sizeof(double) + 16 * sizeof(unsigned long long) + 64;
  if (ok && cudaMalloc((void**)&h->p2p.xs, xs_bytes) != cudaSuccess) { cudaGetLastError(); h->p2p.xs = nullptr; ok = 0; }
  if (ok) { cudaMemset(h->p2p.xs, 0, xs_bytes); cudaDeviceSynchronize(); }
  void* local[4] = {h->Wbuf[0], h->Wbuf[1], h->Wbuf[2], h->p2p.xs};
  P2PExport mine[4];
  memset(mine, 0, sizeof mine);
  for (int k = 0; k < 4 && ok; ++k) {
    unsigned long long base = 0; size_t sz = 0;
    if (range(&base, &sz, (unsigned long long)local[k]) != 0) { ok = 0; break; }
    if (cudaIpcGetMemHandle(&mine[k].hnd, (void*)base) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
    mine[k].off = (unsigned long long)local[k] - base;
  }
  for (int k = 0; k < 4; ++k) mine[k].valid = ok ? 1ull : 0ull;
  const size_t rec = sizeof(mine);
  char *dsend = nullptr, *drecv = nullptr;
  int* dvote = nullptr;
  int rc = EKF_OK;
  std::vector<P2PExport> all((size_t)4 * h->world);
  // the staging buffers are a few hundred bytes: if even they cannot be allocated the communicator is unusable anyway
  if (cudaMalloc((void**)&dsend, rec) != cudaSuccess || cudaMalloc((void**)&drecv, rec * h->world) != cudaSuccess ||
      cudaMalloc((void**)&dvote, sizeof(int)) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(dsend); cudaFree(drecv); cudaFree(dvote);
    p2p_teardown(h);
    h->err = "ekf_dist: no device memory for the peer-mapping handshake";
    return EKF_ERR_CUDA;
  }
  cudaMemcpy(dsend, mine, rec, cudaMemcpyHostToDevice);
  ncclResult_t r = g_nccl.AllGather(dsend, drecv, rec, /*ncclInt8*/ 0, comm, h->stream);
  if (r != 0) { rc = nccl_fail(h, r, "ncclAllGather (IPC handles)"); ok = 0; }
  cudaStreamSynchronize(h->stream);
  if (rc == EKF_OK) cudaMemcpy(all.data(), drecv, rec * h->world, cudaMemcpyDeviceToHost);
  for (int q = 0; q < h->world && ok; ++q)
    for (int k = 0; k < 4; ++k)
      if (!all[(size_t)q * 4 + k].valid) ok = 0;   // a peer could not export: nobody maps anything
  for (int q = 0; q < h->world && ok; ++q) {
    void* ptr[4];
    for (int k = 0; k < 4 && ok; ++k) {
      if (q == h->rank) { ptr[k] = local[k]; continue; }
      void* base = nullptr;
      if (cudaIpcOpenMemHandle(&base, all[(size_t)q * 4 + k].hnd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        fprintf(stderr, "ekf_dist: cudaIpcOpenMemHandle failed (rank %d <- %d)\n", h->rank, q);
        ok = 0;
        break;
      }
      h->p2p.mapped[q][k] = base;
      ptr[k] = (char*)base + all[(size_t)q * 4 + k].off;
    }
    if (!ok) break;
    for (int k = 0; k < 3; ++k) h->p2p.peerW[k][q] = (double*)ptr[k];
    h->p2p.peerSpart[q] = (double*)ptr[3];
    h->p2p.peerFlags[q] = (unsigned long long*)((double*)ptr[3] + spart_doubles);
  }
  // agreement: MIN over the ranks' votes (skipped only when the communicator itself already failed above)
  int agreed = 0;
  if (rc == EKF_OK) {
    cudaMemcpy(dvote, &ok, sizeof(int), cudaMemcpyHostToDevice);
    r = g_nccl.AllReduce(dvote, dvote, 1, /*ncclInt32*/ 2, /*ncclMin*/ 3, comm, h->stream);
    if (r != 0) rc = nccl_fail(h, r, "ncclAllReduce (peer-mapping agreement)");
    cudaStreamSynchronize(h->stream);
    if (rc == EKF_OK) cudaMemcpy(&agreed, dvote, sizeof(int), cudaMemcpyDeviceToHost);
  }
  cudaFree(dsend); cudaFree(drecv); cudaFree(dvote);
  if (agreed == 1) {
    h->p2p.on = true;
  } else {
    if (rc == EKF_OK && h->rank == 0) fprintf(stderr, "ekf_dist: peer mapping not available on every rank: staying on NCCL\n");
    p2p_teardown(h);   // closes every mapping opened so far, frees the slots, p2p.on = false
  }
  return rc;
}

static void p2p_teardown(ekf_handle* h) {
  for (int q = 0; q < 8; ++q)
    for (int k = 0; k < 4; ++k)
      if (h->p2p.mapped[q][k]) { cudaIpcCloseMemHandle(h->p2p.mapped[q][k]); h->p2p.mapped[q][k] = nullptr; }
  if (h->p2p.xs) { cudaFree(h->p2p.xs); h->p2p.xs = nullptr; }
  for (int q = 0; q < 8; ++q) {
    for (int k = 0; k < 3; ++k) h->p2p.peerW[k][q] = nullptr;
    h->p2p.peerSpart[q] = nullptr; h->p2p.peerFlags[q] = nullptr;
  }
  h->p2p.on = false;
}

int ekf_dist_detach(ekf_handle* h) {
  if (!h) return EKF_ERR_ARG;
  if (h->nccl_comm) {
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->p2p.xs) {   // no peer may still be storing into this rank when its buffers go away
      g_nccl.AllReduce(h->p2p.xs, h->p2p.xs, 1, kNcclDouble, 0, (ncclComm_t)h->nccl_comm, h->stream);
      cudaStreamSynchronize(h->stream);
      p2p_teardown(h);
    }
    g_nccl.CommDestroy((ncclComm_t)h->nccl_comm);
    h->nccl_comm = nullptr; h->rank = 0; h->world = 1;
  }
  return EKF_OK;
}

int ekf_dist_peer_memory(const ekf_handle* h) { return h && h->p2p.on ? 1 : 0; }

int ekf_dist_info(const ekf_handle* h, int* rank, int* world, int64_t* allgather_bytes) {
  if (!h) return EKF_ERR_ARG;
  if (rank) *rank = h->rank;
  if (world) *world = h->world;
  if (allgather_bytes) *allgather_bytes = h->dist_bytes;
  return EKF_OK;
}

}  // extern "C"

int ekf_dist_allgather_rows(ekf_handle* h, double* buf, int rows_per_rank, size_t row_elems) {
  const size_t count = (size_t)rows_per_rank * row_elems;
  const ncclResult_t r = g_nccl.AllGather(buf + (size_t)h->rank * count, buf, count, kNcclDouble, (ncclComm_t)h->nccl_comm, h->stream);
  if (r != 0) return nccl_fail(h, r, "ncclAllGather");
  h->dist_bytes += (long long)(count * sizeof(double));
  return 0;
}

// In-place sum over the ranks (the 128 x 128 partial innovation blocks of the row-block partition).
int ekf_dist_allreduce_sum(ekf_handle* h, double* buf, size_t count) {
  const ncclResult_t r = g_nccl.AllReduce(buf, buf, count, kNcclDouble, /*ncclSum*/ 0, (ncclComm_t)h->nccl_comm, h->stream);
  if (r != 0) return nccl_fail(h, r, "ncclAllReduce");
  h->dist_bytes += (long long)(count * sizeof(double));
  return 0;
}
