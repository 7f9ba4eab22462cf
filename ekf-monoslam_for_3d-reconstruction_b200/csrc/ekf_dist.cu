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
#include <cstring>

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
  return EKF_OK;
}

int ekf_dist_detach(ekf_handle* h) {
  if (!h) return EKF_ERR_ARG;
  if (h->nccl_comm) {
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    g_nccl.CommDestroy((ncclComm_t)h->nccl_comm);
    h->nccl_comm = nullptr; h->rank = 0; h->world = 1;
  }
  return EKF_OK;
}

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
