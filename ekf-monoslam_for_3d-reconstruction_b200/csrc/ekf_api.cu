// csrc/ekf_api.cu — the C ABI of include/ekf_b200.h: handle, host-side sequencing of the kernels,
// feature-table management (add / remove), accessors.  No CPU fallback: every entry point needs a
// CUDA device and fails with EKF_ERR_CUDA otherwise.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ekf_b200.h"
#include "ekf_handle.h"
#include "ekf_kernels.h"

static int ekf_fail_cuda(ekf_handle* h, cudaError_t e, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  if (h) h->err = buf;
  return EKF_ERR_CUDA;
}
static int ekf_fail(ekf_handle* h, int code, const char* msg) {
  if (h) h->err = msg;
  return code;
}

// ---- profiling helpers: one event pair per kernel class instance -----------------------------
static cudaEvent_t prof_event(ekf_handle* h) {
  if (!h->prof_pool.empty()) { cudaEvent_t e = h->prof_pool.back(); h->prof_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct ProfScope {
  ekf_handle* h; int cls; long long l0; cudaEvent_t a; cudaStream_t s;
  ProfScope(ekf_handle* hh, int c, cudaStream_t st = nullptr) : h(hh), cls(c), l0(hh->launches), a(nullptr), s(st ? st : hh->stream) {
    if (h->prof_on) { a = prof_event(h); cudaEventRecord(a, s); }
  }
  ~ProfScope() {
    if (h->prof_on) {
      cudaEvent_t b = prof_event(h);
      cudaEventRecord(b, s);
      h->prof_pending.push_back({cls, (int)(h->launches - l0), a, b});
    }
  }
};
// call after a stream synchronize
static void prof_flush(ekf_handle* h) {
  for (auto& r : h->prof_pending) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { h->prof_ms[r.cls] += ms; h->prof_launches[r.cls] += r.nl; }
    h->prof_pool.push_back(r.a); h->prof_pool.push_back(r.b);
  }
  h->prof_pending.clear();
}

// ---- EKF_TRACE=1: per-launch timeline of the pipelined stacked update (start / end of every kernel on its stream, relative to the
// first one), printed to stderr after the step's final synchronize.  Diagnostic only: the event records perturb the schedule a little.
struct TraceRec { const char* name; int b; cudaEvent_t a, e; };
static std::vector<TraceRec> g_trace;
static const bool g_trace_on = getenv("EKF_TRACE") != nullptr;
struct TraceScope {
  const char* name; int b; cudaStream_t s; cudaEvent_t a = nullptr;
  TraceScope(const char* nm, int blk, cudaStream_t st) : name(nm), b(blk), s(st) {
    if (g_trace_on) { cudaEventCreate(&a); cudaEventRecord(a, s); }
  }
  ~TraceScope() {
    if (g_trace_on) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); g_trace.push_back({name, b, a, e}); }
  }
};
static void trace_flush() {
  if (!g_trace_on || g_trace.empty()) return;
  for (auto& r : g_trace) {
    float t0 = 0, t1 = 0;
    cudaEventElapsedTime(&t0, g_trace[0].a, r.a); cudaEventElapsedTime(&t1, g_trace[0].a, r.e);
    fprintf(stderr, "trace %-8s b=%d  %8.1f -> %8.1f us  (%.1f)\n", r.name, r.b, t0 * 1e3, t1 * 1e3, (t1 - t0) * 1e3);
  }
  fprintf(stderr, "trace end\n");
  for (auto& r : g_trace) { cudaEventDestroy(r.a); cudaEventDestroy(r.e); }
  g_trace.clear();
}

template <class T>
static cudaError_t dalloc(T** p, size_t count) { return cudaMalloc((void**)p, sizeof(T) * std::max<size_t>(count, 1)); }

// the filter's stream waits for a host frame still being uploaded on the copy stream (no-op otherwise)
static void frame_ready(ekf_handle* h) {
  if (h->frame_pending) { cudaStreamWaitEvent(h->stream, h->ev_frame, 0); h->frame_pending = false; }
}

static cudaError_t alloc_feattab(FeatTab& t, int cap, int w2) {
  cudaError_t e;
#define A(field, cnt) if ((e = dalloc(&t.field, (size_t)(cnt))) != cudaSuccess) return e;
  A(pos, cap) A(coding, cap) A(innov, cap) A(li, cap) A(hi, cap) A(removef, cap) A(n_tot, cap) A(n_find, cap)
  A(real_index, cap) A(pos_in_z, cap) A(sel, cap) A(center, 2 * cap) A(quality, cap) A(last_ncc, cap)
  A(z, 2 * cap) A(h, 2 * cap) A(Hc, 26 * cap) A(S2, 4 * cap) A(patch, (size_t)cap * w2) A(mpatch, (size_t)cap * w2)
#undef A
  return cudaSuccess;
}
static void free_feattab(FeatTab& t) {
  cudaFree(t.pos); cudaFree(t.coding); cudaFree(t.innov); cudaFree(t.li); cudaFree(t.hi); cudaFree(t.removef);
  cudaFree(t.n_tot); cudaFree(t.n_find); cudaFree(t.real_index); cudaFree(t.pos_in_z); cudaFree(t.sel);
  cudaFree(t.center); cudaFree(t.quality); cudaFree(t.last_ncc); cudaFree(t.z); cudaFree(t.h); cudaFree(t.Hc);
  cudaFree(t.S2); cudaFree(t.patch); cudaFree(t.mpatch);
  t = FeatTab{};
}

extern "C" {

const char* ekf_build_info(void) { return "libekf_b200: sm_100a, fp64, update block " "128" " rows, built " __DATE__; }

void ekf_config_default(ekf_config* c) {
  if (!c) return;
  c->sigma_vx = c->sigma_vy = c->sigma_vz = 0.01;
  c->sigma_wx = c->sigma_wy = c->sigma_wz = 0.01;
  c->rho_0 = 0.1; c->sigma_rho_0 = 0.25; c->T_camera = 0.5;
  c->fx = 592.2860; c->fy = 584.9968; c->u0 = 362.1059; c->v0 = 275.9642;
  c->k1 = -0.3954; c->k2 = 0.5521; c->k3 = 0; c->p1 = -0.0075; c->p2 = 0.0140;
  c->ncc_threshold = 0.8; c->search_clamp = 20; c->ransac_p = 0.99; c->li_threshold_factor = 2;
  c->hi_chi2_threshold = 1; c->quality_ratio = 0.2; c->linearity_threshold = 0.01;
  c->window_size = 21; c->sigma_pixel = 2; c->kernel_size = 1000000000; c->sigma_size = 2; c->scale = 1;
  c->nInitFeatures = 5; c->min_features = 30; c->max_features = 100; c->forsePlane = 0;
  c->ransac_nhyp0 = 10000; c->xyz_conversion = 1; c->abs_int_quirk = 0;
}

const char* ekf_last_error(const ekf_handle* h) { return h ? h->err.c_str() : "null handle"; }

int ekf_destroy(ekf_handle* h) {
  if (!h) return EKF_OK;
  cudaSetDevice(h->device);
  if (h->own_stream) cudaStreamSynchronize(h->own_stream);
  if (h->nccl_comm) ekf_dist_detach(h);
  for (auto& r : h->prof_pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : h->prof_pool) cudaEventDestroy(e);
  cudaFree(h->mu); cudaFree(h->muB); cudaFree(h->Sigma); cudaFree(h->SigmaB); cudaFree(h->W); cudaFree(h->nu);
  cudaFree(h->Lb); cudaFree(h->Dinv); cudaFree(h->Dblk); cudaFree(h->yb); cudaFree(h->delta); cudaFree(h->mu_i); cudaFree(h->cand);
  cudaFree(h->map_dev); cudaFree(h->keep_dev); cudaFree(h->newpos_dev); cudaFree(h->ctl); cudaFree(h->frame); cudaFree(h->raw);
  cudaFree(h->picks_dev); cudaFree(h->out_dev); cudaFree(h->gemm_counters);
  if (h->gemm_stream) { cudaStreamSynchronize(h->gemm_stream); cudaStreamDestroy(h->gemm_stream); }
  cudaFree(h->Wbuf[1]); cudaFree(h->Wbuf[2]); cudaFree(h->Wbuf[3]); cudaFree(h->Wbuf[4]); cudaFree(h->Wbuf[5]); cudaFree(h->Wbuf[6]); cudaFree(h->Wbuf[7]); cudaFree(h->Gbuf);
  cudaFree(h->G2buf); cudaFree(h->Sg2buf[0]); cudaFree(h->Sg2buf[1]);
  for (int i = 0; i < 2; ++i) { if (h->ev_mini[i]) cudaEventDestroy(h->ev_mini[i]); if (h->ev_Sg2[i]) cudaEventDestroy(h->ev_Sg2[i]); }
  cudaFree(h->Dinv2); cudaFree(h->Dblk2); cudaFree(h->yb2); cudaFree(h->delta1); cudaFree(h->delta2); cudaFree(h->gy); cudaFree(h->bt_H); cudaFree(h->bt_zmh); cudaFree(h->Sgbuf); cudaFree(h->bt_pos); cudaFree(h->bt_nd); cudaFree(h->chain_flags); cudaFree(h->tile_order); cudaFree(h->tile_nhot); cudaFree(h->tile_counters);
  if (h->v_stream) { cudaStreamSynchronize(h->v_stream); cudaStreamDestroy(h->v_stream); }
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
  if (h->chain_stream) { cudaStreamSynchronize(h->chain_stream); cudaStreamDestroy(h->chain_stream); }
  if (h->ev_chain) cudaEventDestroy(h->ev_chain);
  if (h->ev_frame) cudaEventDestroy(h->ev_frame);
  if (h->ev_prev) cudaEventDestroy(h->ev_prev);
  if (h->gather_stream) { cudaStreamSynchronize(h->gather_stream); cudaStreamDestroy(h->gather_stream); }
  for (int i = 0; i < 2; ++i) { if (h->ev_F[i]) cudaEventDestroy(h->ev_F[i]); if (h->ev_corr2[i]) cudaEventDestroy(h->ev_corr2[i]); if (h->ev_dd[i]) cudaEventDestroy(h->ev_dd[i]); }
  if (h->ev_A) cudaEventDestroy(h->ev_A);
  if (h->corr_stream) { cudaStreamSynchronize(h->corr_stream); cudaStreamDestroy(h->corr_stream); }
  if (h->ev_S) cudaEventDestroy(h->ev_S);
  if (h->ev_G) cudaEventDestroy(h->ev_G);
  if (h->ev_corr) cudaEventDestroy(h->ev_corr);
  for (int i = 0; i < 3; ++i) { if (h->ev_gather[i]) cudaEventDestroy(h->ev_gather[i]); if (h->ev_V[i]) cudaEventDestroy(h->ev_V[i]); }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  cudaFree(h->det_mask); cudaFree(h->det_eig); cudaFree(h->det_keys); cudaFree(h->det_counters); cudaFree(h->det_xy);
  cudaFree(h->xyz_flag); cudaFree(h->xyz_rmap); cudaFree(h->xyz_pos); cudaFree(h->xyz_coding); cudaFree(h->xyz_y); cudaFree(h->xyz_J);
  if (h->out_host) cudaFreeHost(h->out_host);
  if (h->ctl_host) cudaFreeHost(h->ctl_host);
  free_feattab(h->ft); free_feattab(h->ftB);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return EKF_OK;
}

int ekf_create(const ekf_config* cfg, int feature_capacity, int device, ekf_handle** out) {
  if (!cfg || !out || feature_capacity < 1 || feature_capacity > 8192) return EKF_ERR_ARG;
  *out = nullptr;
  // parts of the reference outside this path are rejected, never emulated on the CPU
  if (cfg->scale < 1 || cfg->scale > 64) return EKF_ERR_ARG;
  if (cfg->window_size < 3 || cfg->window_size > 31) return EKF_ERR_UNSUPPORTED;
  if (cfg->search_clamp > 20 || cfg->search_clamp < 0) return EKF_ERR_UNSUPPORTED;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || device < 0 || device >= ndev) return EKF_ERR_CUDA;
  ekf_handle* h = new ekf_handle;
  h->cfg = *cfg;
  h->device = device;
  auto bail = [&](cudaError_t ee, const char* what) {
    ekf_fail_cuda(h, ee, what, __FILE__, __LINE__);
    fprintf(stderr, "ekf_create: %s\n", h->err.c_str());
    ekf_destroy(h);
    return EKF_ERR_CUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e, "cudaSetDevice");
  {
    // the filter's own stream carries the serial gain chain: highest priority, so that its CTAs are placed before
    // those of the downdate running on the second (lowest-priority) stream
    int plo = 0, phi = 0;
    cudaDeviceGetStreamPriorityRange(&plo, &phi);
    if ((e = cudaStreamCreateWithPriority(&h->own_stream, cudaStreamNonBlocking, phi)) != cudaSuccess) return bail(e, "stream");
  }
  h->stream = h->own_stream;
  if (update_kernels_init() != 0) return bail(cudaGetLastError(), "kernel attributes");
  h->Ncap = feature_capacity;
  h->ncap = EKF_CAM + 6 * feature_capacity;
  h->ld = (h->ncap + 7) & ~7;
  const int w2 = (cfg->window_size * cfg->window_size + 15) & ~15;  // template record stride
  const size_t ssz = (size_t)(h->ncap + EKF_DIST_PAD_ROWS) * h->ld;
#define TRY(x) if ((e = (x)) != cudaSuccess) return bail(e, #x);
  TRY(dalloc(&h->mu, h->ld)) TRY(dalloc(&h->muB, h->ld)) TRY(dalloc(&h->Sigma, ssz)) TRY(dalloc(&h->SigmaB, ssz))
  TRY(dalloc(&h->W, (size_t)(h->ncap + 1 + EKF_DIST_PAD_ROWS) * EKF_UB)) TRY(dalloc(&h->nu, EKF_UB)) TRY(dalloc(&h->Lb, EKF_UB * EKF_UB))
  TRY(dalloc(&h->Dinv, EKF_UB * EKF_UB)) TRY(dalloc(&h->Dblk, EKF_UB * 32)) TRY(dalloc(&h->yb, EKF_UB)) TRY(dalloc(&h->delta, h->ld + EKF_DIST_PAD_ROWS)) TRY(dalloc(&h->mu_i, h->ld))
  TRY(dalloc(&h->cand, h->Ncap)) TRY(dalloc(&h->map_dev, h->ncap)) TRY(dalloc(&h->keep_dev, h->Ncap))
  TRY(dalloc(&h->newpos_dev, h->Ncap)) TRY(dalloc(&h->ctl, 1)) TRY(dalloc(&h->gemm_counters, 2))
  h->Wbuf[0] = h->W;
  TRY(dalloc(&h->Wbuf[1], (size_t)(h->ncap + 1 + EKF_DIST_PAD_ROWS) * EKF_UB)) TRY(dalloc(&h->Wbuf[2], (size_t)(h->ncap + 1 + EKF_DIST_PAD_ROWS) * EKF_UB)) TRY(dalloc(&h->Wbuf[3], (size_t)(h->ncap + 1) * EKF_UB))
  TRY(dalloc(&h->Gbuf, EKF_UB * EKF_UB))
  TRY(dalloc(&h->Wbuf[4], (size_t)(h->ncap + 1) * EKF_UB)) TRY(dalloc(&h->Wbuf[5], (size_t)(h->ncap + 1) * EKF_UB))
  TRY(dalloc(&h->Wbuf[6], (size_t)(h->ncap + 1) * EKF_UB)) TRY(dalloc(&h->Wbuf[7], (size_t)(h->ncap + 1) * EKF_UB))
  TRY(dalloc(&h->G2buf, EKF_UB * EKF_UB)) TRY(dalloc(&h->Sg2buf[0], EKF_UB * EKF_UB)) TRY(dalloc(&h->Sg2buf[1], EKF_UB * EKF_UB))
  TRY(dalloc(&h->Dinv2, EKF_UB * EKF_UB)) TRY(dalloc(&h->Dblk2, EKF_UB * 32)) TRY(dalloc(&h->yb2, EKF_UB)) TRY(dalloc(&h->gy, EKF_UB))
  TRY(dalloc(&h->bt_H, 26 * (size_t)feature_capacity)) TRY(dalloc(&h->bt_zmh, 2 * (size_t)feature_capacity)) TRY(dalloc(&h->bt_pos, feature_capacity))
  TRY(dalloc(&h->bt_nd, feature_capacity)) TRY(dalloc(&h->Sgbuf, EKF_UB * EKF_UB))
  h->tile_blk_cap = (feature_capacity + EKF_UB / 2 - 1) / (EKF_UB / 2) + 1;
  h->tile_T_cap = std::min(128, (h->ncap + 63) / 64);   // the hot-first tile lists exist for n <= 8192 (the schedule using them stops at n = 6000)
  TRY(dalloc(&h->tile_order, (size_t)h->tile_blk_cap * (h->tile_T_cap * (h->tile_T_cap + 1) / 2))) TRY(dalloc(&h->tile_nhot, h->tile_blk_cap))
  TRY(dalloc(&h->tile_counters, h->tile_blk_cap))
  TRY(dalloc(&h->chain_flags, 3 * (size_t)h->tile_blk_cap + 8)) TRY(cudaMemset(h->chain_flags, 0, sizeof(unsigned int) * (3 * (size_t)h->tile_blk_cap + 8)))
  TRY(dalloc(&h->delta1, h->ld + EKF_DIST_PAD_ROWS)) TRY(dalloc(&h->delta2, h->ld + EKF_DIST_PAD_ROWS))
  {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);   // lo = least priority
    TRY(cudaStreamCreateWithPriority(&h->gemm_stream, cudaStreamNonBlocking, lo))
    TRY(cudaStreamCreateWithPriority(&h->corr_stream, cudaStreamNonBlocking, hi))
    TRY(cudaStreamCreateWithPriority(&h->copy_stream, cudaStreamNonBlocking, hi))
    TRY(cudaStreamCreateWithPriority(&h->chain_stream, cudaStreamNonBlocking, hi))
    TRY(cudaEventCreateWithFlags(&h->ev_chain, cudaEventDisableTiming))
    TRY(cudaEventCreateWithFlags(&h->ev_frame, cudaEventDisableTiming)) TRY(cudaEventCreateWithFlags(&h->ev_prev, cudaEventDisableTiming))
    TRY(cudaStreamCreateWithPriority(&h->v_stream, cudaStreamNonBlocking, hi))
    TRY(cudaStreamCreateWithPriority(&h->gather_stream, cudaStreamNonBlocking, hi))
    for (int i = 0; i < 2; ++i) {
      TRY(cudaEventCreateWithFlags(&h->ev_F[i], cudaEventDisableTiming)) TRY(cudaEventCreateWithFlags(&h->ev_corr2[i], cudaEventDisableTiming))
      TRY(cudaEventCreateWithFlags(&h->ev_dd[i], cudaEventDisableTiming))
    }
    TRY(cudaEventCreateWithFlags(&h->ev_A, cudaEventDisableTiming))
    for (int i = 0; i < 2; ++i) {
      TRY(cudaEventCreateWithFlags(&h->ev_mini[i], cudaEventDisableTiming)) TRY(cudaEventCreateWithFlags(&h->ev_Sg2[i], cudaEventDisableTiming))
    }
    TRY(cudaEventCreateWithFlags(&h->ev_G, cudaEventDisableTiming)) TRY(cudaEventCreateWithFlags(&h->ev_corr, cudaEventDisableTiming))
    for (int i = 0; i < 3; ++i) {
      TRY(cudaEventCreateWithFlags(&h->ev_gather[i], cudaEventDisableTiming)) TRY(cudaEventCreateWithFlags(&h->ev_V[i], cudaEventDisableTiming))
    }
    TRY(cudaEventCreateWithFlags(&h->ev_S, cudaEventDisableTiming)) TRY(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming)) TRY(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming))
    // look-ahead pays off once the downdate of a block (n x n x 128) is long against the block's gain chain:
    // measured +27 % at n = 11972; at n = 3014 only +6 % (the chain kernels' large shared-memory CTAs queue behind the
    // short, resident downdate) against +33 % for the schedule below; EKF_LOOKAHEAD_MIN_N overrides the threshold
    const char* e = getenv("EKF_LOOKAHEAD_MIN_N");
    h->lookahead = e ? atoi(e) : 6000;
    // below that threshold: the schedule that starts the downdate of a block beside the NEXT block's Cholesky
    // (stacked_update_factor_beside_downdate); EKF_PIPE_MIN_N overrides (0 = never)
    e = getenv("EKF_PIPE_MIN_N");
    h->pipe_small = e ? atoi(e) : 1000;
    // which of the two schedules for that range: 1 = chain-short (default: V off the critical chain; same-box A/B at n = 3014:
    // 0.764 against 0.786 ms per cfg2 step), 0 = factor-beside-downdate — see DESIGN.md section 4
    e = getenv("EKF_SCHED");
    h->sched = e ? atoi(e) : 1;
    e = getenv("EKF_V_AFTER_DD");
    h->v_after_dd = e ? atoi(e) : 0;
    e = getenv("EKF_GATHER_HP");
    h->gather_hp = e ? atoi(e) : 0;
    e = getenv("EKF_DD_RELEASE");
    h->dd_release = e ? atoi(e) : 2;
    e = getenv("EKF_S_LOOKAHEAD");
    h->s_lookahead = e ? atoi(e) : 0;   // measured slower (0.772 against 0.726 ms per cfg2 step): see DESIGN.md section 4
    e = getenv("EKF_PRELAUNCH");
    h->prelaunch_on = e ? atoi(e) : 1;
    e = getenv("EKF_SPLIT_DD");
    h->split_dd = e ? atoi(e) : 0;   // 0: 2-D grid; 1: tile list, the next gather gated on its hot tiles; 2: tile list only
  }
  TRY(dalloc(&h->xyz_flag, h->Ncap)) TRY(dalloc(&h->xyz_rmap, 2 * (size_t)h->ncap)) TRY(dalloc(&h->xyz_pos, h->Ncap))
  TRY(dalloc(&h->xyz_coding, h->Ncap)) TRY(dalloc(&h->xyz_y, 3 * (size_t)h->Ncap)) TRY(dalloc(&h->xyz_J, 18 * (size_t)h->Ncap))
  TRY(alloc_feattab(h->ft, h->Ncap, w2)) TRY(alloc_feattab(h->ftB, h->Ncap, w2))
  h->out_bytes = sizeof(double) * 210 + sizeof(int) * (16 + 3 * (size_t)h->Ncap);
  h->picks_cap = EKF_PICKS_CAP;
  TRY(dalloc(&h->picks_dev, (size_t)h->picks_cap))
  TRY(cudaMalloc((void**)&h->out_dev, h->out_bytes)) TRY(cudaMallocHost((void**)&h->out_host, h->out_bytes)) TRY(cudaMallocHost((void**)&h->ctl_host, sizeof(DevCtl)))
  TRY(cudaMemsetAsync(h->ctl, 0, sizeof(DevCtl), h->stream))
  TRY(cudaMemsetAsync(h->gemm_counters, 0, 2 * sizeof(int), h->stream))
  TRY(cudaMemsetAsync(h->Sigma, 0, sizeof(double) * ssz, h->stream))
  TRY(cudaMemsetAsync(h->SigmaB, 0, sizeof(double) * ssz, h->stream))
  TRY(cudaMemsetAsync(h->mu, 0, sizeof(double) * h->ld, h->stream))
  // device copy of the configuration scalars
  DevCfg& d = h->dcfg;
  d.cam = CamParams{cfg->fx, cfg->fy, cfg->u0, cfg->v0, cfg->k1, cfg->k2, cfg->k3, cfg->p1, cfg->p2};
  d.Vmax[0] = cfg->sigma_vx * cfg->sigma_vx; d.Vmax[1] = cfg->sigma_vy * cfg->sigma_vy; d.Vmax[2] = cfg->sigma_vz * cfg->sigma_vz;
  d.Vmax[3] = cfg->sigma_wx * cfg->sigma_wx; d.Vmax[4] = cfg->sigma_wy * cfg->sigma_wy; d.Vmax[5] = cfg->sigma_wz * cfg->sigma_wz;
  d.sigma_pixel_2 = (double)(cfg->sigma_pixel * cfg->sigma_pixel);
  d.th_low = cfg->li_threshold_factor * cfg->sigma_pixel;
  d.th_hi = cfg->hi_chi2_threshold;
  d.ransac_p = cfg->ransac_p;
  d.linearity_threshold = cfg->linearity_threshold;
  d.rho_0 = cfg->rho_0; d.sigma_rho_0 = cfg->sigma_rho_0;
  d.T_camera = cfg->T_camera; d.kernel_min_size = cfg->kernel_size;
  d.ncc_threshold = (float)cfg->ncc_threshold; d.search_clamp = (float)cfg->search_clamp;
  d.sigma_size_f = (float)cfg->sigma_size; d.quality_ratio = (float)cfg->quality_ratio;
  d.window = cfg->window_size; d.sigma_pixel = cfg->sigma_pixel; d.nhyp0 = cfg->ransac_nhyp0;
  d.forsePlane = cfg->forsePlane; d.abs_int_quirk = cfg->abs_int_quirk;
  d.tstride = w2;
  // initial state (vslamRansac.cpp:163-216)
  double mu0[EKF_CAM] = {0};
  mu0[3] = 0.0; mu0[4] = 0.0; mu0[5] = -0.707106781; mu0[6] = 0.707106781; mu0[13] = 1;
  std::vector<double> S0((size_t)EKF_CAM * EKF_CAM, 0.0);
  for (int i = 0; i < EKF_CAM; ++i) S0[i * EKF_CAM + i] = 0.0000000004 * 1.0;
  S0[13 * EKF_CAM + 13] = 0.09;
  const double svv = 0.0004, sww = 0.0004;
  for (int i = 0; i < 3; ++i) { S0[(7 + i) * EKF_CAM + 7 + i] = svv * svv * 1.0; S0[(10 + i) * EKF_CAM + 10 + i] = sww * sww * 1.0; }
  TRY(cudaMemcpyAsync(h->mu, mu0, sizeof mu0, cudaMemcpyHostToDevice, h->stream))
  TRY(cudaMemcpy2DAsync(h->Sigma, sizeof(double) * h->ld, S0.data(), sizeof(double) * EKF_CAM, sizeof(double) * EKF_CAM, EKF_CAM,
                        cudaMemcpyHostToDevice, h->stream))
  TRY(cudaStreamSynchronize(h->stream))
#undef TRY
  h->n = EKF_CAM; h->N = 0;
  *out = h;
  return EKF_OK;
}

int ekf_set_stream(ekf_handle* h, void* s) {
  if (!h) return EKF_ERR_ARG;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  h->stream = s ? (cudaStream_t)s : h->own_stream;
  return EKF_OK;
}
int ekf_sync(ekf_handle* h) {
  if (!h) return EKF_ERR_ARG;
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  frame_ready(h);
  EKF_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return EKF_OK;
}

// captureNewFrame(Mat[, double]) (vslamRansac.cpp:226-245): time stamp, resize by 1 / scale, BGR -> gray
static int capture_common(ekf_handle* h, const uint8_t* img, int width, int height, int stride, int channels, double stamp,
                          bool device_src) {
  if (!h || !img || width < 8 || height < 8 || (channels != 1 && channels != 3) || stride < width * channels) return EKF_ERR_ARG;
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  const int scale = h->cfg.scale;
  const int dw = width / scale, dh = height / scale;   // cv::Size(width / scale, height / scale), integer division
  if (dw < 8 || dh < 8) return EKF_ERR_ARG;
  if (stamp >= 0) {  // captureNewFrame(Mat, double), vslamRansac.cpp:226-233
    if (h->old_ts > 0) h->dT = (stamp - h->old_ts);
    h->old_ts = stamp;
  }
  const int dstride = (dw + 15) & ~15;
  const size_t need = (size_t)dstride * dh;
  // the frame goes into the filter's buffer on the copy stream, ordered after everything already enqueued on the filter's stream
  // (earlier readers of the frame buffer)
  cudaStream_t cs = h->copy_stream;   // device frames too: the copy is ordered after the filter's stream and runs beside predict
  EKF_CUDA_CHECK(cudaEventRecord(h->ev_prev, h->stream));
  EKF_CUDA_CHECK(cudaStreamWaitEvent(cs, h->ev_prev, 0));
  if (need > h->frame_cap) {
    EKF_CUDA_CHECK(cudaStreamSynchronize(cs));
    EKF_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    cudaFree(h->frame);
    h->frame = nullptr;
    EKF_CUDA_CHECK(cudaMalloc((void**)&h->frame, need));
    h->frame_cap = need;
  }
  if (scale == 1 && channels == 1) {
    EKF_CUDA_CHECK(cudaMemcpy2DAsync(h->frame, dstride, img, stride, width, height,
                                     device_src ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, cs));
  } else {
    const uint8_t* src = img;
    int sstride = stride;
    if (!device_src) {
      const size_t rawneed = (size_t)width * channels * height;
      if (rawneed > h->raw_cap) {
        EKF_CUDA_CHECK(cudaStreamSynchronize(cs));
        EKF_CUDA_CHECK(cudaStreamSynchronize(h->stream));
        cudaFree(h->raw);
        h->raw = nullptr;
        EKF_CUDA_CHECK(cudaMalloc((void**)&h->raw, rawneed));
        h->raw_cap = rawneed;
      }
      EKF_CUDA_CHECK(cudaMemcpy2DAsync(h->raw, (size_t)width * channels, img, stride, (size_t)width * channels, height,
                                       cudaMemcpyHostToDevice, cs));
      src = h->raw; sstride = width * channels;
    }
    launch_capture_resize_gray(cs, src, width, height, sstride, channels, h->frame, dw, dh, dstride, &h->launches);
    EKF_CUDA_CHECK(cudaGetLastError());
  }
  if (h->fv.px != h->frame || h->fv.w != dw || h->fv.h != dh || h->fv.stride != dstride)
    match_make_tensor_map(&h->frame_map, h->frame, dw, dh, dstride, 1, h->cfg.window_size, (int)h->cfg.search_clamp);
  h->fv = FrameView{h->frame, dw, dh, dstride};
  h->have_frame = true;
  EKF_CUDA_CHECK(cudaEventRecord(h->ev_frame, cs));
  h->frame_pending = true;
  return EKF_OK;
}
int ekf_capture_frame(ekf_handle* h, const uint8_t* gray, int width, int height, int stride, double stamp) {
  return capture_common(h, gray, width, height, stride, 1, stamp, false);
}
int ekf_capture_frame_device(ekf_handle* h, const uint8_t* gray, int width, int height, int stride, double stamp) {
  return capture_common(h, gray, width, height, stride, 1, stamp, true);
}
int ekf_capture_frame_bgr(ekf_handle* h, const uint8_t* bgr, int width, int height, int stride, double stamp) {
  return capture_common(h, bgr, width, height, stride, 3, stamp, false);
}
int ekf_get_frame(ekf_handle* h, uint8_t* out, int* width, int* height) {   // returnGrayImg (vslamRansac.cpp:1364)
  if (!h || !width || !height) return EKF_ERR_ARG;
  if (!h->have_frame) return ekf_fail(h, EKF_ERR_STATE, "no frame captured");
  *width = h->fv.w; *height = h->fv.h;
  if (out) {
    EKF_CUDA_CHECK(cudaSetDevice(h->device));
    frame_ready(h);
    EKF_CUDA_CHECK(cudaMemcpy2DAsync(out, h->fv.w, h->frame, h->fv.stride, h->fv.w, h->fv.h, cudaMemcpyDeviceToHost, h->stream));
    EKF_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  }
  return EKF_OK;
}

int ekf_predict(ekf_handle* h, const double dv[3], const double dw[3], int vcontrol) {
  if (!h) return EKF_ERR_ARG;
  if (!h->have_frame) return ekf_fail(h, EKF_ERR_STATE, "predict before captureNewFrame");
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  const double z3[3] = {0, 0, 0};
  h->cam_cache_ok = false;
  if (vcontrol) h->noise_cov_factor = 0; else h->noise_cov_factor++;
  if (h->cfg.kernel_size < 100000) frame_ready(h);   // motion-blur template prediction on: keep the whole step behind the upload
  {
    ProfScope ps(h, 0);
    launch_predict(h->stream, h->Sigma, h->ld, h->n, h->mu, h->ft, h->N, h->fv, h->ctl, h->dcfg, h->dT, dv ? dv : z3, dw ? dw : z3,
                   vcontrol, &h->launches);
  }
  EKF_CUDA_CHECK(cudaGetLastError());
  h->predicted = true;
  h->cache_ok = false;
  return EKF_OK;
}

static int refresh_cache(ekf_handle* h) {
  if (h->cache_ok) return EKF_OK;
  const int N = h->N;
  auto get = [&](auto& vec, const auto* src, size_t cnt) -> cudaError_t {
    vec.resize(cnt);
    if (cnt == 0) return cudaSuccess;
    return cudaMemcpyAsync(vec.data(), src, cnt * sizeof(vec[0]), cudaMemcpyDeviceToHost, h->stream);
  };
  EKF_CUDA_CHECK(get(h->c_pos, h->ft.pos, N)); EKF_CUDA_CHECK(get(h->c_coding, h->ft.coding, N));
  EKF_CUDA_CHECK(get(h->c_innov, h->ft.innov, N)); EKF_CUDA_CHECK(get(h->c_li, h->ft.li, N));
  EKF_CUDA_CHECK(get(h->c_hi, h->ft.hi, N)); EKF_CUDA_CHECK(get(h->c_removef, h->ft.removef, N));
  EKF_CUDA_CHECK(get(h->c_ntot, h->ft.n_tot, N)); EKF_CUDA_CHECK(get(h->c_nfind, h->ft.n_find, N));
  EKF_CUDA_CHECK(get(h->c_real, h->ft.real_index, N)); EKF_CUDA_CHECK(get(h->c_posz, h->ft.pos_in_z, N));
  EKF_CUDA_CHECK(get(h->c_center, h->ft.center, 2 * (size_t)N)); EKF_CUDA_CHECK(get(h->c_quality, h->ft.quality, N));
  EKF_CUDA_CHECK(get(h->c_ncc, h->ft.last_ncc, N)); EKF_CUDA_CHECK(get(h->c_z, h->ft.z, 2 * (size_t)N));
  EKF_CUDA_CHECK(get(h->c_h, h->ft.h, 2 * (size_t)N)); EKF_CUDA_CHECK(get(h->c_Hc, h->ft.Hc, 26 * (size_t)N));
  EKF_CUDA_CHECK(get(h->c_S2, h->ft.S2, 4 * (size_t)N));
  EKF_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  h->cache_ok = true;
  return EKF_OK;
}

int ekf_match(ekf_handle* h, int* n_matched) {
  if (!h) return EKF_ERR_ARG;
  if (!h->predicted) return ekf_fail(h, EKF_ERR_STATE, "match before predict");
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  frame_ready(h);
  {
    ProfScope ps(h, 1);
    launch_match_filter(h->stream, h->ft, h->N, h->fv, h->dcfg, &h->frame_map, &h->launches);
  }
  EKF_CUDA_CHECK(cudaGetLastError());
  h->cache_ok = false;
  if (n_matched) {
    int rc = refresh_cache(h);
    if (rc) return rc;
    int c = 0;
    for (int i = 0; i < h->N; ++i) c += h->c_innov[i] ? 1 : 0;
    *n_matched = c;
  }
  return EKF_OK;
}

// removeFeature for a set of feature indices at once (vslamRansac.cpp:373-421; order-independent).
static int remove_features(ekf_handle* h, const std::vector<int>& victims) {
  if (victims.empty()) return EKF_OK;
  int rc = refresh_cache(h);
  if (rc) return rc;
  std::vector<char> dead(h->N, 0);
  for (int v : victims) {
    if (v < 0 || v >= h->N) return EKF_ERR_ARG;
    dead[v] = 1;
  }
  // archive good XYZ features (V:394-404)
  for (int v = 0; v < h->N; ++v) {
    if (!dead[v] || !(h->c_nfind[v] > 5 && h->m_coding[v] == 1)) continue;
    DeletedPatch dp;
    dp.real_index = h->c_real[v];
    const int pos = h->m_pos[v];
    EKF_CUDA_CHECK(cudaMemcpyAsync(dp.XYZ_pos, h->mu + pos, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    double blk[9];
    EKF_CUDA_CHECK(cudaMemcpy2DAsync(blk, 3 * sizeof(double), h->Sigma + (size_t)pos * h->ld + pos, h->ld * sizeof(double),
                                     3 * sizeof(double), 3, cudaMemcpyDeviceToHost, h->stream));
    EKF_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) dp.cov_4_delete[c * 3 + r] = blk[c * 3 + r];
    h->deleted.push_back(dp);
  }
  std::vector<int> keep, newpos, map;
  for (int i = 0; i < EKF_CAM; ++i) map.push_back(i);
  std::vector<int> npos2, ncod2;
  for (int f = 0; f < h->N; ++f) {
    if (dead[f]) continue;
    const int fs = h->m_coding[f] ? 3 : 6;
    keep.push_back(f);
    newpos.push_back((int)map.size());
    npos2.push_back((int)map.size());
    ncod2.push_back(h->m_coding[f]);
    for (int c = 0; c < fs; ++c) map.push_back(h->m_pos[f] + c);
  }
  const int n2 = (int)map.size(), N2 = (int)keep.size();
  EKF_CUDA_CHECK(cudaMemcpyAsync(h->map_dev, map.data(), sizeof(int) * n2, cudaMemcpyHostToDevice, h->stream));
  if (N2 > 0) {
    EKF_CUDA_CHECK(cudaMemcpyAsync(h->keep_dev, keep.data(), sizeof(int) * N2, cudaMemcpyHostToDevice, h->stream));
    EKF_CUDA_CHECK(cudaMemcpyAsync(h->newpos_dev, newpos.data(), sizeof(int) * N2, cudaMemcpyHostToDevice, h->stream));
  }
  launch_gather_state(h->stream, h->Sigma, h->SigmaB, h->ld, h->mu, h->muB, n2, h->map_dev, &h->launches);
  launch_gather_features(h->stream, h->ft, h->ftB, N2, h->keep_dev, h->newpos_dev, h->dcfg.tstride, &h->launches);
  EKF_CUDA_CHECK(cudaGetLastError());
  EKF_CUDA_CHECK(cudaStreamSynchronize(h->stream));  // host vectors above go out of scope
  std::swap(h->Sigma, h->SigmaB);
  std::swap(h->mu, h->muB);
  std::swap(h->ft, h->ftB);
  h->n = n2; h->N = N2;
  h->m_pos = npos2; h->m_coding = ncod2;
  h->cache_ok = false;
  return EKF_OK;
}

// convert2XYZ_ifLinear (only >= 0) / convert2XYZ_ifLinearAll (only < 0), V:741-780
static int convert_xyz(ekf_handle* h, int only) {
  if (h->N == 0) return EKF_OK;
  cudaStream_t st = h->stream;
  launch_xyz_decide(st, h->Sigma, h->ld, h->mu, h->ft, h->N, h->dcfg, only, h->xyz_flag, h->xyz_y, h->xyz_J, &h->launches);
  EKF_CUDA_CHECK(cudaGetLastError());
  std::vector<int> flag(h->N);
  EKF_CUDA_CHECK(cudaMemcpyAsync(flag.data(), h->xyz_flag, sizeof(int) * h->N, cudaMemcpyDeviceToHost, st));
  EKF_CUDA_CHECK(cudaStreamSynchronize(st));
  int nconv = 0;
  for (int f = 0; f < h->N; ++f) nconv += flag[f] ? 1 : 0;
  if (nconv == 0) return EKF_OK;
  std::vector<int> rmap, npos(h->N), ncod(h->N);
  for (int i = 0; i < EKF_CAM; ++i) { rmap.push_back(i); rmap.push_back(-1); }
  for (int f = 0; f < h->N; ++f) {
    npos[f] = (int)rmap.size() / 2;
    if (flag[f]) {
      for (int x = 0; x < 3; ++x) { rmap.push_back(h->m_pos[f]); rmap.push_back(4 * f + x); }
      ncod[f] = 1;
    } else {
      const int fs = h->m_coding[f] ? 3 : 6;
      for (int c = 0; c < fs; ++c) { rmap.push_back(h->m_pos[f] + c); rmap.push_back(-1); }
      ncod[f] = h->m_coding[f];
    }
  }
  const int n2 = (int)rmap.size() / 2;
  EKF_CUDA_CHECK(cudaMemcpyAsync(h->xyz_rmap, rmap.data(), sizeof(int) * rmap.size(), cudaMemcpyHostToDevice, st));
  EKF_CUDA_CHECK(cudaMemcpyAsync(h->xyz_pos, npos.data(), sizeof(int) * h->N, cudaMemcpyHostToDevice, st));
  EKF_CUDA_CHECK(cudaMemcpyAsync(h->xyz_coding, ncod.data(), sizeof(int) * h->N, cudaMemcpyHostToDevice, st));
  launch_xyz_apply(st, h->Sigma, h->SigmaB, h->ld, h->mu, h->muB, n2, h->xyz_rmap, h->xyz_J, h->xyz_y, h->ft, h->N, h->xyz_pos,
                   h->xyz_coding, &h->launches);
  EKF_CUDA_CHECK(cudaGetLastError());
  EKF_CUDA_CHECK(cudaStreamSynchronize(st));  // host vectors above go out of scope
  std::swap(h->Sigma, h->SigmaB);
  std::swap(h->mu, h->muB);
  h->n = n2;
  h->m_pos = npos; h->m_coding = ncod;
  h->cache_ok = false;
  return EKF_OK;
}

// One stacked update over `cnt` selected features (ft.sel), block by block (see ekf_update.cu).
// Stacked update with one block of look-ahead.  The chain of block b+1 (gain: S, Cholesky, V) does not
// wait for the full covariance downdate of block b: W_{b+1} = Sigma_b H^T is formed as
//     W'_{b+1} - V_b (H_{b+1} V_b)^T,   W'_{b+1} gathered from Sigma_{b-1}
// (one n x 128 x 128 correction GEMM instead of the n x n x 128 downdate on the critical path), while the
// downdate Sigma -= V_b V_b^T and the next gather run on a second, low-priority stream beside the
// single-CTA Cholesky.  Same arithmetic per element up to the order of the two subtractions.
static int stacked_update_lookahead(ekf_handle* h, int cnt) {
  cudaStream_t sm = h->stream, sg = h->gemm_stream;
  const int nblk = (cnt + EKF_UB / 2 - 1) / (EKF_UB / 2);
  // Row-block partition (ekf_dist): this rank gathers, corrects, solves and downdates rows [r0, r1) only; the partial
  // S blocks are all-reduced and V_b is all-gathered on the main stream, beside the downdate of block b-1 on the second.
  const bool dist = h->nccl_comm != nullptr && h->world > 1;
  int rpr = h->n, r0 = 0, r1 = h->n;
  if (dist) {
    rpr = (((h->n + h->world - 1) / h->world) + 31) & ~31;
    if ((long long)rpr * h->world > (long long)h->n + EKF_DIST_PAD_ROWS) return (int)cudaErrorInvalidValue;
    r0 = std::min(h->n, h->rank * rpr);
    r1 = std::min(h->n, r0 + rpr);
  }
  cudaMemsetAsync(h->delta, 0, sizeof(double) * (size_t)(h->n + (dist ? EKF_DIST_PAD_ROWS : 0)), sm);
  cudaEventRecord(h->ev_fork, sm);
  cudaStreamWaitEvent(sg, h->ev_fork, 0);
  for (int b = 0; b < nblk && b < 2; ++b) {
    ProfScope ps(h, 3, sg);
    launch_blk_gather(sg, h->Sigma, h->ld, r0, r1, h->ft, b * (EKF_UB / 2), cnt, nullptr, h->Wbuf[b], nullptr, &h->launches);
    cudaEventRecord(h->ev_gather[b], sg);
  }
  for (int b = 0; b < nblk; ++b) {
    const int f0 = b * (EKF_UB / 2);
    double* Wb = h->Wbuf[b % 3];
    cudaStreamWaitEvent(sm, h->ev_gather[b % 3], 0);
    if (b > 0 && r1 > r0) {
      ProfScope ps(h, 3);
      double* Vp = h->Wbuf[(b - 1) % 3];
      launch_blk_G(sm, Vp, h->ft, f0, cnt, h->Gbuf, &h->launches);
      const int rc = launch_gemm_nt_sub(sm, Wb + (size_t)r0 * EKF_UB, EKF_UB, Vp + (size_t)r0 * EKF_UB, EKF_UB, h->Gbuf, EKF_UB, r1 - r0, EKF_UB,
                                        EKF_UB, nullptr, 0, h->gemm_counters, &h->launches);
      if (rc) return rc;
    }
    if (dist && h->p2p.on) {
      // peer-memory exchange: the S kernel stores its partial block into every rank's slot, the factor kernel waits for
      // the peers' epochs and sums the slots, the V kernel stores its finished rows into every rank's panel — no NCCL call
      P2PView pv;
      for (int q = 0; q < 8; ++q) { pv.w[q] = h->p2p.peerW[b % 3][q]; pv.spart[q] = h->p2p.peerSpart[q]; pv.flags[q] = h->p2p.peerFlags[q]; }
      pv.rank = h->rank; pv.world = h->world; pv.epoch = ++h->p2p.epoch;
      const unsigned long long* lflags = h->p2p.peerFlags[h->rank];
      unsigned int* tickets = (unsigned int*)(h->p2p.peerFlags[h->rank] + 16);
      { ProfScope ps(h, 4); launch_blk_S_part_p2p(sm, Wb, h->ft, f0, cnt, h->dcfg, r0, r1, h->rank == 0 ? 1 : 0, h->delta, h->nu, h->Lb, pv, tickets, &h->launches); }
      { ProfScope ps(h, 4); launch_blk_factor_p2p(sm, h->p2p.peerSpart[h->rank], lflags, h->world, pv.epoch, h->Lb, h->nu, h->Dinv, h->Dblk, h->yb, h->ctl, &h->launches); }
      { ProfScope ps(h, 5); launch_blk_V_p2p(sm, Wb, r0, r1, h->Dinv, h->Dblk, h->yb, pv, tickets + 1, lflags, h->world, pv.epoch, h->ctl, &h->launches); }
      h->dist_bytes += (long long)sizeof(double) * ((size_t)EKF_UB * EKF_UB + (size_t)(r1 - r0) * EKF_UB) * (h->world - 1);
      { ProfScope ps(h, 5); launch_delta_rows(sm, Wb, h->yb, h->delta, h->n, &h->launches); }
    } else if (dist) {
      { ProfScope ps(h, 4); launch_blk_S_part(sm, Wb, h->ft, f0, cnt, h->dcfg, r0, r1, h->rank == 0 ? 1 : 0, h->delta, h->nu, h->Lb, &h->launches); }
      { ProfScope ps(h, 11); if (ekf_dist_allreduce_sum(h, h->Lb, (size_t)EKF_UB * EKF_UB)) return (int)cudaErrorUnknown; }
      { ProfScope ps(h, 4); launch_blk_factor_only(sm, h->Lb, h->nu, h->Dinv, h->Dblk, h->yb, h->ctl, &h->launches); }
      { ProfScope ps(h, 5); launch_blk_V(sm, Wb, r0, r1, h->Dinv, h->Dblk, h->yb, nullptr, &h->launches); }
      { ProfScope ps(h, 11); if (ekf_dist_allgather_rows(h, Wb, rpr, EKF_UB)) return (int)cudaErrorUnknown; }
      { ProfScope ps(h, 5); launch_delta_rows(sm, Wb, h->yb, h->delta, h->n, &h->launches); }
    } else {
      { ProfScope ps(h, 4);
        launch_blk_S_nu(sm, Wb, h->ft, f0, cnt, h->dcfg, h->delta, h->Lb, h->nu, &h->launches);
        launch_blk_factor_only(sm, h->Lb, h->nu, h->Dinv, h->Dblk, h->yb, h->ctl, &h->launches); }
      { ProfScope ps(h, 5); launch_blk_V(sm, Wb, r0, r1, h->Dinv, h->Dblk, h->yb, h->delta, &h->launches); }
    }
    cudaEventRecord(h->ev_V[b % 3], sm);
    cudaStreamWaitEvent(sg, h->ev_V[b % 3], 0);
    if (r1 > r0) {
      ProfScope ps(h, 6, sg);
      const int rc = launch_gemm_nt_sub(sg, h->Sigma + (size_t)r0 * h->ld, h->ld, Wb + (size_t)r0 * EKF_UB, EKF_UB, Wb, EKF_UB, r1 - r0, h->n, EKF_UB,
                                        nullptr, dist ? 0 : h->lower_only, h->gemm_counters, &h->launches);
      if (rc) return rc;
    }
    if (b + 2 < nblk) {
      ProfScope ps(h, 3, sg);
      launch_blk_gather(sg, h->Sigma, h->ld, r0, r1, h->ft, (b + 2) * (EKF_UB / 2), cnt, nullptr, h->Wbuf[(b + 2) % 3], nullptr, &h->launches);
      cudaEventRecord(h->ev_gather[(b + 2) % 3], sg);
    }
  }
  cudaEventRecord(h->ev_join, sg);
  cudaStreamWaitEvent(sm, h->ev_join, 0);
  if (dist) { ProfScope ps(h, 11); if (ekf_dist_allgather_rows(h, h->Sigma, rpr, (size_t)h->ld)) return (int)cudaErrorUnknown; }
  {
    ProfScope ps(h, 7);
    launch_finish_update(sm, h->Sigma, h->ld, h->n, h->mu, h->delta, h->ctl, &h->launches);
  }
  return 0;
}

// Second schedule of the same look-ahead algebra, for maps whose downdate is SHORT (one or two waves of tiles):
// there the low-priority downdate, once resident, keeps every SM busy for its whole duration and the single-CTA
// Cholesky of the next block (140 KB of shared memory: needs a drained SM) queues behind it.  Here the downdate
// of block b-1 is held back until S_b is ready and is launched right AFTER the Cholesky of block b, so the
// Cholesky takes its SM first and the downdate fills the other 147: the longest serial kernel of the chain is
// hidden behind the widest one.  The gather of W'_{b+1} follows the downdate on the second stream and runs beside V_b.
//   main:  [G_b, corr_b, S_b] -> factor_b -> V_b            second:  downdate_{b-1} -> gather_{b+1}
static int stacked_update_factor_beside_downdate(ekf_handle* h, int cnt) {
  cudaStream_t sm = h->stream, sg = h->gemm_stream, sc = h->corr_stream;
  const int nblk = (cnt + EKF_UB / 2 - 1) / (EKF_UB / 2);
  // raw[b & 1]: W'_b as gathered (read by S_b); cor[b & 1]: second copy, corrected in place to W_b, then V_b
  double* raw[2] = {h->Wbuf[1], h->Wbuf[2]};
  double* cor[2] = {h->Wbuf[0], h->Wbuf[3]};
  cudaMemsetAsync(h->delta, 0, sizeof(double) * (size_t)h->n, sm);
  cudaEventRecord(h->ev_fork, sm);
  cudaStreamWaitEvent(sg, h->ev_fork, 0);
  for (int b = 0; b < nblk && b < 2; ++b) {   // W'_0 = W_0 and W'_1, both from the prior covariance
    ProfScope ps(h, 3, sg);
    launch_blk_gather2(sg, h->Sigma, h->ld, h->n, h->ft, b * (EKF_UB / 2), cnt, raw[b], cor[b], &h->launches);
    cudaEventRecord(h->ev_gather[b], sg);
  }
  for (int b = 0; b < nblk; ++b) {
    const int f0 = b * (EKF_UB / 2);
    double* Vp = cor[(b - 1) & 1];
    cudaStreamWaitEvent(sm, h->ev_gather[b & 1], 0);
    if (b > 0) {
      { ProfScope ps(h, 3); launch_blk_G(sm, Vp, h->ft, f0, cnt, h->Gbuf, &h->launches); }
      cudaEventRecord(h->ev_G, sm);
      cudaStreamWaitEvent(sc, h->ev_G, 0);
      // W_b = W'_b - V_{b-1} G^T on the third stream, beside S_b and the Cholesky: only V_b needs it
      const int rc = launch_gemm_nt_sub(sc, cor[b & 1], EKF_UB, Vp, EKF_UB, h->Gbuf, EKF_UB, h->n, EKF_UB, EKF_UB, nullptr, 0, h->gemm_counters, &h->launches);
      if (rc) return rc;
      cudaEventRecord(h->ev_corr, sc);
    }
    { ProfScope ps(h, 4); launch_blk_S_nu_G(sm, raw[b & 1], h->ft, f0, cnt, h->dcfg, h->delta, b > 0 ? h->Gbuf : nullptr, h->Lb, h->nu, &h->launches); }
    if (b > 0) {   // release the downdate of block b-1 (it already waits for V_{b-1})
      cudaEventRecord(h->ev_S, sm);
      cudaStreamWaitEvent(sg, h->ev_S, 0);
    }
    { ProfScope ps(h, 4); launch_blk_factor_only(sm, h->Lb, h->nu, h->Dinv, h->Dblk, h->yb, h->ctl, &h->launches); }
    if (b > 0) {
      {
        ProfScope ps(h, 6, sg);
        const int rc = launch_gemm_nt_sub(sg, h->Sigma, h->ld, Vp, EKF_UB, Vp, EKF_UB, h->n, h->n, EKF_UB, nullptr, h->lower_only, h->gemm_counters, &h->launches);
        if (rc) return rc;
      }
      if (b + 1 < nblk) {
        cudaStreamWaitEvent(sg, h->ev_corr, 0);   // the next gather overwrites V_{b-1}, which the correction of W_b reads
        ProfScope ps(h, 3, sg);
        launch_blk_gather2(sg, h->Sigma, h->ld, h->n, h->ft, (b + 1) * (EKF_UB / 2), cnt, raw[(b + 1) & 1], cor[(b + 1) & 1], &h->launches);
        cudaEventRecord(h->ev_gather[(b + 1) & 1], sg);
      }
      cudaStreamWaitEvent(sm, h->ev_corr, 0);
    }
    { ProfScope ps(h, 5); launch_blk_V(sm, cor[b & 1], 0, h->n, h->Dinv, h->Dblk, h->yb, h->delta, &h->launches); }
    cudaEventRecord(h->ev_V[b % 3], sm);
    cudaStreamWaitEvent(sg, h->ev_V[b % 3], 0);
  }
  {
    double* Vl = cor[(nblk - 1) & 1];
    ProfScope ps(h, 6, sg);
    const int rc = launch_gemm_nt_sub(sg, h->Sigma, h->ld, Vl, EKF_UB, Vl, EKF_UB, h->n, h->n, EKF_UB, nullptr, h->lower_only, h->gemm_counters, &h->launches);
    if (rc) return rc;
  }
  cudaEventRecord(h->ev_join, sg);
  cudaStreamWaitEvent(sm, h->ev_join, 0);
  {
    ProfScope ps(h, 7);
    launch_finish_update(sm, h->Sigma, h->ld, h->n, h->mu, h->delta, h->ctl, &h->launches);
  }
  return 0;
}

// Third schedule, same look-ahead algebra, for the same range of n: the n-row solve V_b LEAVES the critical chain.  The next
// block needs V_b only through G_{b+1} = H_{b+1} V_b = (H_{b+1} W_b) L_b^-T, a 128-row solve on the rows of W_b that block b+1
// touches (k_blk_Gx), and through nu_{b+1}, where the missing term of delta is G_{b+1} y_b.  Chain per block:
//   main:  Gx_b -> S_b -> factor_b                (8 + 4 + 42 us instead of V 21 + G 8 + S 4 + factor 42)
//   v:     V_b (beside factor_{b+1}; out of place: Gx_{b+1} reads W_b meanwhile), delta_b = delta_{b-1} + V_b y_b (ping-pong)
//   corr:  W_b = W'_b - V_{b-1} G_b^T             (beside factor_b; needs all of V_{b-1})
//   second: downdate_{b-1} (released by S_b, as before) -> gather of W'_{b+1}
// Two sets of factor outputs (L, D, y) alternate: V_{b-1} reads one while factor_b writes the other.
// Whether the low-innovation update of this step will take the chain-short schedule, as far as the host knows BEFORE it has read
// n_li back: if so, ekf_update_after_match starts the block tables and the first two gathers (which read only the prior covariance and
// the inlier list the RANSAC kernel left on the device) while that read-back is in flight.
static ChainFlags chain_flags_view(ekf_handle* h);
static bool chain_short_likely(const ekf_handle* h) {
  const bool partitioned = h->nccl_comm && h->world > 1;
  if (partitioned || (h->sched != 1 && h->sched != 3) || h->pipe_small <= 0 || h->n < h->pipe_small) return false;
  if (h->lookahead > 0 && h->n >= h->lookahead) return false;
  return h->N > EKF_UB / 2;
}
static void chain_short_prelaunch(ekf_handle* h) {
  cudaStream_t sm = h->stream, sg = h->gemm_stream;
  const BlkTab bt{h->bt_H, h->bt_zmh, h->bt_pos, h->bt_nd};
  cudaEventRecord(h->ev_fork, sm);
  cudaStreamWaitEvent(sg, h->ev_fork, 0);
  // everything on the second stream: the host's wait on the main stream must not include it
  launch_blk_prep(sg, h->ft, h->N, h->bt_H, h->bt_zmh, h->bt_pos, h->bt_nd, &h->launches, &h->ctl->n_li);
  double* raw[2] = {h->Wbuf[1], h->Wbuf[2]};
  double* cor[2] = {h->Wbuf[0], h->Wbuf[3]};
  for (int b = 0; b < 2; ++b) {
    ProfScope ps(h, 3, sg); TraceScope ts("gather", b, sg);
    launch_blk_gather2(sg, h->Sigma, h->ld, h->n, h->ft, b * (EKF_UB / 2), 0, raw[b], cor[b], &h->launches, bt, &h->ctl->n_li);
    cudaEventRecord(h->ev_gather[b], sg);
  }
  if (h->s_lookahead && h->N > EKF_UB) {   // hot rows of W''_2 from the prior covariance (S look-ahead)
    ProfScope ps(h, 3, sg); TraceScope ts("mini", 2, sg);
    launch_blk_gather_hot(sg, h->Sigma, h->ld, h->n, h->ft, 2 * (EKF_UB / 2), 0, h->Wbuf[6], &h->launches, bt, &h->ctl->n_li);
    cudaEventRecord(h->ev_mini[0], sg);
  }
  h->prelaunched = true;
}

static int stacked_update_chain_short(ekf_handle* h, int cnt) {
  cudaStream_t sm = h->stream, sg = h->gemm_stream, sc = h->corr_stream, sv = h->v_stream;
  const int nblk = (cnt + EKF_UB / 2 - 1) / (EKF_UB / 2);
  double* raw[2] = {h->Wbuf[1], h->Wbuf[2]};   // W'_b as gathered (read by S_b)
  double* cor[2] = {h->Wbuf[0], h->Wbuf[3]};   // second copy, corrected in place to W_b (read by Gx_{b+1} and V_b)
  double* Vb[2] = {h->Wbuf[4], h->Wbuf[5]};    // V_b (read by corr_{b+1} and downdate_b)
  double* Ls[2] = {h->Dinv, h->Dinv2};
  double* Ds[2] = {h->Dblk, h->Dblk2};
  double* ys[2] = {h->yb, h->yb2};
  double* dl[3] = {h->delta, h->delta1, h->delta2};   // delta_b lives in dl[b % 3]; delta_{-1} = delta_{-2} = 0
  // S look-ahead: from block 2 on, S_b is formed from the hot rows of W''_b gathered one block EARLIER (from Sigma_{b-3}, next to the
  // full gather of block b-1) with two correction terms, -G2 G2^T (G2 = H_b V_{b-2}, formed off the chain as soon as V_{b-2}
  // exists) and -G G^T (k_blk_Gx), so that S_b does not wait for the downdate of block b-2 and the gather behind it
  double* rawH[2] = {h->Wbuf[6], h->Wbuf[7]};
  const bool sla = h->s_lookahead && nblk > 2;
  // EKF_SCHED=3: pre-positioned factor.  factor_b is launched on its own stream as soon as S_b has been enqueued, takes over the SM
  // factor_{b-1} just gave back and waits there for the flag the last CTA of the S_b kernel raises: no launch gap between S_b and
  // the factor, and the downdate of block b-1 need not be held back until S_b is done (EKF_DD_RELEASE: 3 after S_b as in the plain
  // chain-short schedule, 2 after -G G^T, 1 after Gx, 0 not at all).
  const bool prepos = h->sched == 3 && !sla && nblk <= std::min(h->tile_blk_cap, 1000);
  cudaStream_t sf = h->chain_stream;
  if (prepos) h->chain_seq++;
  const ChainFlags fl = chain_flags_view(h);
  const int rel = prepos ? h->dd_release : 3;
  for (int i = 0; i < 3; ++i) cudaMemsetAsync(dl[i], 0, sizeof(double) * (size_t)h->n, sm);
  // gather beside the downdate (hot tiles first): lower-triangle mode with square tiles only
  const int T = (h->n + 63) / 64, Ltiles = T * (T + 1) / 2;
  const bool gate = h->split_dd == 1;
  const bool split = h->split_dd && nblk > 2 && h->lower_only && T <= h->tile_T_cap && nblk <= h->tile_blk_cap && gemm_uses_square_tiles();
  // The gather of W'_{b+1} sits on the cycle S_b -> downdate_{b-1} -> gather -> S_{b+1}; on the downdate's own low-priority stream its CTAs
  // queue behind those of the correction GEMM and of V (21 us instead of 11).  EKF_GATHER_HP=1 moves it to a high-priority stream behind
  // an event of the downdate.
  const bool ghp = h->gather_hp != 0;
  cudaStream_t sgat = ((split && gate) || ghp) ? h->gather_stream : sg;
  const BlkTab bt{h->bt_H, h->bt_zmh, h->bt_pos, h->bt_nd};
  const bool pre = h->prelaunched;   // tables and the first two gathers are already under way
  h->prelaunched = false;
  if (!pre) launch_blk_prep(sm, h->ft, cnt, h->bt_H, h->bt_zmh, h->bt_pos, h->bt_nd, &h->launches);
  if (split) launch_blk_tile_order(sm, h->ft, cnt, T, h->tile_order, h->tile_nhot, h->tile_counters, &h->launches);
  cudaEventRecord(h->ev_fork, sm);
  cudaStreamWaitEvent(sg, h->ev_fork, 0);
  cudaStreamWaitEvent(sv, h->ev_fork, 0);
  if (prepos) cudaStreamWaitEvent(sf, h->ev_fork, 0);
  for (int b = 0; b < nblk && b < 2 && !pre; ++b) {   // W'_0 = W_0 and W'_1, both from the prior covariance
    ProfScope ps(h, 3, sg); TraceScope ts("gather", b, sg);
    launch_blk_gather2(sg, h->Sigma, h->ld, h->n, h->ft, b * (EKF_UB / 2), cnt, raw[b], cor[b], &h->launches, bt);
    cudaEventRecord(h->ev_gather[b], sg);
  }
  if (sla && !(pre && h->N > EKF_UB)) {
    ProfScope ps(h, 3, sg); TraceScope ts("mini", 2, sg);
    launch_blk_gather_hot(sg, h->Sigma, h->ld, h->n, h->ft, 2 * (EKF_UB / 2), cnt, rawH[0], &h->launches, bt);
    cudaEventRecord(h->ev_mini[0], sg);
  }
  for (int b = 0; b < nblk; ++b) {
    const int f0 = b * (EKF_UB / 2), p = b & 1, q = p ^ 1;
    if (b > 0) {
      if (prepos) cudaStreamWaitEvent(sm, h->ev_F[q], 0);   // factor_{b-1} ran on its own stream
      if (b > 1) {
        cudaStreamWaitEvent(sm, h->ev_corr2[q], 0);   // W_{b-1} is corrected (b - 1 = 0 needs no correction)
        cudaStreamWaitEvent(sm, h->ev_V[(b - 2) % 3], 0);   // delta_{b-2} is complete; V_{b-2} no longer reads the factor set p
      }
      { ProfScope ps(h, 4); TraceScope ts("Gx", b, sm); launch_blk_Gx(sm, cor[q], h->ft, f0, cnt, Ls[q], Ds[q], ys[q], h->Gbuf, h->gy, &h->launches, bt); }
      cudaEventRecord(h->ev_G, sm);
      if (rel == 1) { cudaEventRecord(h->ev_S, sm); cudaStreamWaitEvent(sg, h->ev_S, 0); }
      // W_b = W'_b - V_{b-1} G_b^T beside S_b and the Cholesky
      cudaStreamWaitEvent(sc, h->ev_G, 0);
      cudaStreamWaitEvent(sc, h->ev_V[(b - 1) % 3], 0);
      cudaStreamWaitEvent(sc, h->ev_gather[p], 0);   // the second copy of W'_b
      TraceScope ts("corr", b, sc);
      const int rc = launch_gemm_nt_sub(sc, cor[p], EKF_UB, Vb[q], EKF_UB, h->Gbuf, EKF_UB, h->n, EKF_UB, EKF_UB, nullptr, 0, h->gemm_counters, &h->launches);
      if (rc) return rc;
      cudaEventRecord(h->ev_corr2[p], sc);
      if (sla && b + 1 < nblk) {   // G2_{b+1} = H_{b+1} V_{b-1} and -G2 G2^T, a block ahead of their use
        launch_blk_G(sc, Vb[q], h->ft, (b + 1) * (EKF_UB / 2), cnt, h->G2buf, &h->launches);
        launch_blk_Sg(sc, h->G2buf, h->Sg2buf[q], &h->launches);
        cudaEventRecord(h->ev_Sg2[q], sc);
      }
      // -G_b G_b^T now, in the shadow of the gather this block still waits for
      { ProfScope ps(h, 4); TraceScope ts("Sg", b, sm); launch_blk_Sg(sm, h->Gbuf, h->Sgbuf, &h->launches); }
      if (rel == 2) { cudaEventRecord(h->ev_S, sm); cudaStreamWaitEvent(sg, h->ev_S, 0); }
    }
    if (sla && b >= 2) {   // hot rows of W''_b and -G2 G2^T were formed a block ago
      cudaStreamWaitEvent(sm, h->ev_mini[p], 0);
      cudaStreamWaitEvent(sm, h->ev_Sg2[p], 0);
      ProfScope ps(h, 4); TraceScope ts("S", b, sm);
      launch_blk_S_nu_G(sm, rawH[p], h->ft, f0, cnt, h->dcfg, dl[(b + 1) % 3] /* delta_{b-2} */, h->Gbuf, h->Lb, h->nu, &h->launches, h->gy, bt, h->Sgbuf,
                        nullptr, nullptr, 0, h->Sg2buf[p]);
    } else {
      cudaStreamWaitEvent(sm, h->ev_gather[p], 0);   // only S_b reads W'_b: Gx_b ran in the shadow of the gather
      ProfScope ps(h, 4); TraceScope ts("S", b, sm);
      launch_blk_S_nu_G(sm, raw[p], h->ft, f0, cnt, h->dcfg, dl[(b + 1) % 3] /* delta_{b-2} */, b > 0 ? h->Gbuf : nullptr, h->Lb, h->nu, &h->launches,
                        b > 0 ? h->gy : nullptr, bt, b > 0 ? h->Sgbuf : nullptr, prepos ? fl.tickets + 1 : nullptr, prepos ? fl.sg + b : nullptr,
                        fl.token0 + b);
    }
    if (b > 0 && rel == 3) {   // release the downdate of block b-1 (it also waits for V_{b-1})
      cudaEventRecord(h->ev_S, sm);
      cudaStreamWaitEvent(sg, h->ev_S, 0);
    }
    if (prepos) {   // every producer the kernel waits for is already enqueued (S_b above): the wait cannot hold back what it needs
      ProfScope ps(h, 4, sf); TraceScope ts("factor", b, sf);
      launch_blk_factor_wait(sf, h->Lb, h->nu, Ls[p], Ds[p], ys[p], h->ctl, fl.sg + b, fl.token0 + b, &h->launches);
    } else {
      ProfScope ps(h, 4); TraceScope ts("factor", b, sm); launch_blk_factor_only(sm, h->Lb, h->nu, Ls[p], Ds[p], ys[p], h->ctl, &h->launches);
    }
    cudaEventRecord(h->ev_F[p], prepos ? sf : sm);
    if (b > 0) {
      cudaStreamWaitEvent(sg, h->ev_V[(b - 1) % 3], 0);
      const bool two = split && b + 1 < nblk;   // the gather of W'_{b+1} starts beside this downdate, after its hot tiles
      if (two && gate) {
        cudaEventRecord(h->ev_A, sg);            // everything the downdate waits for
        cudaStreamWaitEvent(sgat, h->ev_A, 0);
      }
      {
        ProfScope ps(h, 6, sg); TraceScope ts("dd", b - 1, sg);
        const int rc = launch_gemm_nt_sub(sg, h->Sigma, h->ld, Vb[q], EKF_UB, Vb[q], EKF_UB, h->n, h->n, EKF_UB, nullptr, h->lower_only, h->gemm_counters, &h->launches,
                                          two ? h->tile_order + (size_t)(b + 1) * Ltiles : nullptr, Ltiles, h->tile_nhot + (b + 1), h->tile_counters + (b + 1));
        if (rc) return rc;
      }
      if (ghp && !(two && gate) && b + 1 < nblk) {
        cudaEventRecord(h->ev_A, sg);            // the downdate of block b-1 is complete
        cudaStreamWaitEvent(sgat, h->ev_A, 0);
      }
      if (b + 1 < nblk) {
        // cor[q] / raw[q] are free: Gx_b (before S_b, which released the downdate) and V_{b-1} (waited for above) have read them
        ProfScope ps(h, 3, sgat); TraceScope ts("gather", b + 1, sgat);
        if (two && gate) launch_blk_gather2_after_tiles(sgat, h->Sigma, h->ld, h->n, h->ft, (b + 1) * (EKF_UB / 2), cnt, raw[q], cor[q],
                                                h->tile_counters + (b + 1), h->tile_nhot + (b + 1), h->ctl, &h->launches, bt);
        else launch_blk_gather2(sgat, h->Sigma, h->ld, h->n, h->ft, (b + 1) * (EKF_UB / 2), cnt, raw[q], cor[q], &h->launches, bt);
        cudaEventRecord(h->ev_gather[q], sgat);
      }
      if (sla && b + 2 < nblk) {   // hot rows of W''_{b+2} from the same covariance (Sigma_{b-1}); rawH[p] was last read by S_b
        ProfScope ps(h, 3, sgat); TraceScope ts("mini", b + 2, sgat);
        launch_blk_gather_hot(sgat, h->Sigma, h->ld, h->n, h->ft, (b + 2) * (EKF_UB / 2), cnt, rawH[p], &h->launches, bt);
        cudaEventRecord(h->ev_mini[p], sgat);
      }
      cudaEventRecord(h->ev_dd[q], sg);
    }
    // V_b on its own stream: after factor_b, the correction of W_b and the downdate that still reads Vb[p] (block b-2)
    cudaStreamWaitEvent(sv, h->ev_F[p], 0);
    if (b > 0) cudaStreamWaitEvent(sv, h->ev_corr2[p], 0);
    if (b > 1) cudaStreamWaitEvent(sv, h->ev_dd[p], 0);
    if (sla && b > 1) cudaStreamWaitEvent(sv, h->ev_Sg2[p], 0);   // G2_b has read V_{b-2} out of the panel V_b is about to overwrite
    // EKF_V_AFTER_DD=1: V_b is not needed before the downdate of block b (released by S_{b+1}); holding it back until the downdate of
    // block b-1 is through keeps its 95 CTAs out of that downdate's last wave
    if (h->v_after_dd && b > 0) cudaStreamWaitEvent(sv, h->ev_dd[q], 0);
    { ProfScope ps(h, 5, sv); TraceScope ts("V", b, sv);
      launch_blk_V(sv, cor[p], 0, h->n, Ls[p], Ds[p], ys[p], dl[b % 3], &h->launches, Vb[p], dl[(b + 2) % 3] /* delta_{b-1} */); }
    cudaEventRecord(h->ev_V[b % 3], sv);
  }
  {
    const int p = (nblk - 1) & 1;
    cudaStreamWaitEvent(sg, h->ev_V[(nblk - 1) % 3], 0);
    ProfScope ps(h, 6, sg); TraceScope ts("dd", nblk - 1, sg);
    const int rc = launch_gemm_nt_sub(sg, h->Sigma, h->ld, Vb[p], EKF_UB, Vb[p], EKF_UB, h->n, h->n, EKF_UB, nullptr, h->lower_only, h->gemm_counters, &h->launches);
    if (rc) return rc;
  }
  cudaEventRecord(h->ev_join, sg);
  cudaStreamWaitEvent(sm, h->ev_join, 0);
  {
    ProfScope ps(h, 7); TraceScope ts("finish", nblk, sm);
    launch_finish_update(sm, h->Sigma, h->ld, h->n, h->mu, dl[(nblk - 1) % 3], h->ctl, &h->launches);
  }
  return 0;
}

// Fourth schedule (EKF_SCHED=2), the chain-short algebra with a RESIDENT factor CTA: k_chain_factor is launched once per stacked
// update, keeps its SM, assembles S_b itself and factors every block, talking to the other kernels through flag words
// (gather_b / Sg_b -> factor_b -> Gx_{b+1} and, through a one-warp gate, V_b).  What that buys:
//  * the downdate of block b-1 starts as soon as V_{b-1} exists — the other schedules hold it back until S_b is formed so that the
//    next factor KERNEL finds a free SM first;
//  * the chain per block loses a kernel and two launch gaps: Gx_b -> Sg_b -> [flag] -> assemble + factor -> [flag] -> Gx_{b+1}.
// Every wait is bounded and every producer is launched before the kernel that waits for it, so the waits resolve in launch order.
static ChainFlags chain_flags_view(ekf_handle* h) {
  ChainFlags f;
  f.gather = h->chain_flags; f.sg = h->chain_flags + h->tile_blk_cap; f.fact = h->chain_flags + 2 * (size_t)h->tile_blk_cap;
  f.tickets = h->chain_flags + 3 * (size_t)h->tile_blk_cap;
  f.token0 = h->chain_seq * 1024u + 1u;
  return f;
}
static bool resident_chain_likely(const ekf_handle* h) {
  const bool partitioned = h->nccl_comm && h->world > 1;
  if (partitioned || h->sched != 2 || h->pipe_small <= 0 || h->n < h->pipe_small) return false;
  if (h->lookahead > 0 && h->n >= h->lookahead) return false;
  return h->N > EKF_UB / 2 && (h->N + EKF_UB / 2 - 1) / (EKF_UB / 2) <= std::min(h->tile_blk_cap, 1000);
}
static void resident_chain_prelaunch(ekf_handle* h) {
  cudaStream_t sm = h->stream, sg = h->gemm_stream;
  const BlkTab bt{h->bt_H, h->bt_zmh, h->bt_pos, h->bt_nd};
  h->chain_seq++;
  cudaEventRecord(h->ev_fork, sm);
  cudaStreamWaitEvent(sg, h->ev_fork, 0);
  launch_blk_prep(sg, h->ft, h->N, h->bt_H, h->bt_zmh, h->bt_pos, h->bt_nd, &h->launches, &h->ctl->n_li);
  double* raw[2] = {h->Wbuf[1], h->Wbuf[2]};
  double* cor[2] = {h->Wbuf[0], h->Wbuf[3]};
  for (int b = 0; b < 2; ++b) {
    ProfScope ps(h, 3, sg); TraceScope ts("gather", b, sg);
    launch_blk_gather2(sg, h->Sigma, h->ld, h->n, h->ft, b * (EKF_UB / 2), 0, raw[b], cor[b], &h->launches, bt, &h->ctl->n_li);
    cudaEventRecord(h->ev_gather[b], sg);
  }
  h->prelaunched = true;
}
static int stacked_update_resident_chain(ekf_handle* h, int cnt) {
  cudaStream_t sm = h->stream, sg = h->gemm_stream, sc = h->corr_stream, sv = h->v_stream, sf = h->chain_stream;
  const int nblk = (cnt + EKF_UB / 2 - 1) / (EKF_UB / 2);
  double* raw[2] = {h->Wbuf[1], h->Wbuf[2]};
  double* cor[2] = {h->Wbuf[0], h->Wbuf[3]};
  double* Vb[2] = {h->Wbuf[4], h->Wbuf[5]};
  double* Ls[2] = {h->Dinv, h->Dinv2};
  double* Ds[2] = {h->Dblk, h->Dblk2};
  double* ys[2] = {h->yb, h->yb2};
  double* dl[3] = {h->delta, h->delta1, h->delta2};
  for (int i = 0; i < 3; ++i) cudaMemsetAsync(dl[i], 0, sizeof(double) * (size_t)h->n, sm);
  const BlkTab bt{h->bt_H, h->bt_zmh, h->bt_pos, h->bt_nd};
  const bool pre = h->prelaunched;
  h->prelaunched = false;
  if (!pre) {
    h->chain_seq++;
    launch_blk_prep(sm, h->ft, cnt, h->bt_H, h->bt_zmh, h->bt_pos, h->bt_nd, &h->launches);
  }
  const ChainFlags fl = chain_flags_view(h);
  cudaEventRecord(h->ev_fork, sm);
  cudaStreamWaitEvent(sg, h->ev_fork, 0);
  cudaStreamWaitEvent(sv, h->ev_fork, 0);
  cudaStreamWaitEvent(sf, h->ev_fork, 0);
  {   // the resident factor CTA: first on the device, before any downdate exists
    ChainFactorArgs a;
    a.S = h->Lb; a.nu = h->nu;
    for (int i = 0; i < 2; ++i) { a.L[i] = Ls[i]; a.D[i] = Ds[i]; a.y[i] = ys[i]; }
    a.fl = fl; a.nblk = nblk;
    // NOTHING may follow this launch in its stream until every producer it waits for has been enqueued: an event record behind it
    // completes only with the kernel, and when the stream shares a hardware queue with another one (more streams than
    // connections) it would hold back that stream's later launches — the gathers this kernel waits for (measured: every wait
    // ran into its time-out).  No profiling / trace events here for the same reason; ev_chain is recorded at the end.
    launch_chain_factor(sf, a, h->ctl, &h->launches);
  }
  for (int b = 0; b < nblk && b < 2 && !pre; ++b) {
    ProfScope ps(h, 3, sg); TraceScope ts("gather", b, sg);
    launch_blk_gather2(sg, h->Sigma, h->ld, h->n, h->ft, b * (EKF_UB / 2), cnt, raw[b], cor[b], &h->launches, bt);
    cudaEventRecord(h->ev_gather[b], sg);
  }
  for (int b = 0; b < nblk; ++b) {
    const int f0 = b * (EKF_UB / 2), p = b & 1, q = p ^ 1;
    if (b > 0) {
      if (b > 1) {
        cudaStreamWaitEvent(sm, h->ev_corr2[q], 0);         // W_{b-1} is corrected
        cudaStreamWaitEvent(sm, h->ev_V[(b - 2) % 3], 0);   // delta_{b-2} is complete (the resident CTA reads it once Sg_b is published)
      }
      { ProfScope ps(h, 4); TraceScope ts("Gx", b, sm);
        launch_blk_Gx(sm, cor[q], h->ft, f0, cnt, Ls[q], Ds[q], ys[q], h->Gbuf, h->gy, &h->launches, bt, fl.fact + (b - 1), fl.token0 + (b - 1), h->ctl); }
      cudaEventRecord(h->ev_G, sm);
      cudaStreamWaitEvent(sc, h->ev_G, 0);
      cudaStreamWaitEvent(sc, h->ev_V[(b - 1) % 3], 0);
      cudaStreamWaitEvent(sc, h->ev_gather[p], 0);
      {
        TraceScope ts("corr", b, sc);
        const int rc = launch_gemm_nt_sub(sc, cor[p], EKF_UB, Vb[q], EKF_UB, h->Gbuf, EKF_UB, h->n, EKF_UB, EKF_UB, nullptr, 0, h->gemm_counters, &h->launches);
        if (rc) return rc;
      }
      cudaEventRecord(h->ev_corr2[p], sc);
    }
    // S_b and nu_b by the ten-CTA kernel (G G^T on DMMA + the 13-row gather), whose last CTA raises the flag the resident CTA waits for
    cudaStreamWaitEvent(sm, h->ev_gather[p], 0);
    { ProfScope ps(h, 4); TraceScope ts("S", b, sm);
      launch_blk_S_nu_G(sm, raw[p], h->ft, f0, cnt, h->dcfg, dl[(b + 1) % 3], b > 0 ? h->Gbuf : nullptr, h->Lb, h->nu, &h->launches,
                        b > 0 ? h->gy : nullptr, bt, nullptr, fl.tickets + 1, fl.sg + b, fl.token0 + b); }
    if (b > 0) {
      // downdate of block b-1 as soon as V_{b-1} exists, then the gather of W'_{b+1}
      cudaStreamWaitEvent(sg, h->ev_V[(b - 1) % 3], 0);
      {
        ProfScope ps(h, 6, sg); TraceScope ts("dd", b - 1, sg);
        const int rc = launch_gemm_nt_sub(sg, h->Sigma, h->ld, Vb[q], EKF_UB, Vb[q], EKF_UB, h->n, h->n, EKF_UB, nullptr, h->lower_only, h->gemm_counters, &h->launches);
        if (rc) return rc;
      }
      cudaEventRecord(h->ev_dd[q], sg);
      if (b + 1 < nblk) {
        // raw[q] / cor[q] are free: the resident CTA is past S_{b-1} (V_{b-1}, waited for above, ran behind factor_{b-1}), Gx_b has
        // read cor[q] (ev_G) and so has V_{b-1}
        cudaStreamWaitEvent(sg, h->ev_G, 0);
        ProfScope ps(h, 3, sg); TraceScope ts("gather", b + 1, sg);
        launch_blk_gather2(sg, h->Sigma, h->ld, h->n, h->ft, (b + 1) * (EKF_UB / 2), cnt, raw[q], cor[q], &h->launches, bt);
        cudaEventRecord(h->ev_gather[q], sg);
      }
    }
    // V_b behind a one-warp gate on factor_b
    if (b > 0) cudaStreamWaitEvent(sv, h->ev_corr2[p], 0);
    else cudaStreamWaitEvent(sv, h->ev_gather[0], 0);
    if (b > 1) cudaStreamWaitEvent(sv, h->ev_dd[p], 0);
    { ProfScope ps(h, 5, sv); TraceScope ts("V", b, sv);
      launch_wait_flag(sv, fl.fact + b, fl.token0 + b, h->ctl, &h->launches);
      launch_blk_V(sv, cor[p], 0, h->n, Ls[p], Ds[p], ys[p], dl[b % 3], &h->launches, Vb[p], dl[(b + 2) % 3]); }
    cudaEventRecord(h->ev_V[b % 3], sv);
  }
  {
    const int p = (nblk - 1) & 1;
    cudaStreamWaitEvent(sg, h->ev_V[(nblk - 1) % 3], 0);
    ProfScope ps(h, 6, sg); TraceScope ts("dd", nblk - 1, sg);
    const int rc = launch_gemm_nt_sub(sg, h->Sigma, h->ld, Vb[p], EKF_UB, Vb[p], EKF_UB, h->n, h->n, EKF_UB, nullptr, h->lower_only, h->gemm_counters, &h->launches);
    if (rc) return rc;
  }
  cudaEventRecord(h->ev_join, sg);
  cudaEventRecord(h->ev_chain, sf);
  cudaStreamWaitEvent(sm, h->ev_join, 0);
  cudaStreamWaitEvent(sm, h->ev_chain, 0);
  {
    ProfScope ps(h, 7); TraceScope ts("finish", nblk, sm);
    launch_finish_update(sm, h->Sigma, h->ld, h->n, h->mu, dl[(nblk - 1) % 3], h->ctl, &h->launches);
  }
  return 0;
}

static int stacked_update(ekf_handle* h, int cnt, bool plane = false) {
  if (cnt <= 0 && !plane) return 0;
  if (cnt < 0) cnt = 0;
  if (!plane && cnt > EKF_UB / 2) {
    const bool partitioned = h->nccl_comm && h->world > 1;
    if (h->lookahead > 0 && h->n >= h->lookahead) return stacked_update_lookahead(h, cnt);   // also row-block partitioned
    if (!partitioned && h->pipe_small > 0 && h->n >= h->pipe_small)
      return (h->sched == 2 && (cnt + EKF_UB / 2 - 1) / (EKF_UB / 2) <= std::min(h->tile_blk_cap, 1000)) ? stacked_update_resident_chain(h, cnt)
           : h->sched >= 1 ? stacked_update_chain_short(h, cnt) : stacked_update_factor_beside_downdate(h, cnt);
  }
  if (h->prelaunched) {   // the speculative gathers of the chain-short schedule wrote the W panels on the second stream: order after them
    cudaStreamWaitEvent(h->stream, h->ev_gather[0], 0);
    cudaStreamWaitEvent(h->stream, h->ev_gather[1], 0);
    h->prelaunched = false;
  }
  cudaStream_t st = h->stream;
  // Row-block partition (BASELINE config 4): every rank holds a replica of Sigma, updates only its
  // rows [r0, r1) and exchanges the small panels: W_b rows (S_b needs the camera / feature rows of W_b),
  // V_b rows (the downdate of a row block needs all of V_b as its right factor) and delta; the updated
  // row blocks of Sigma are all-gathered once at the end.  world == 1: r0 = 0, r1 = n, no exchange.
  const bool dist = h->nccl_comm != nullptr && h->world > 1;
  int rpr = h->n, r0 = 0, r1 = h->n;
  if (dist) {
    rpr = (((h->n + h->world - 1) / h->world) + 31) & ~31;
    if ((long long)rpr * h->world > (long long)h->n + EKF_DIST_PAD_ROWS) return (int)cudaErrorInvalidValue;
    r0 = std::min(h->n, h->rank * rpr);
    r1 = std::min(h->n, r0 + rpr);
  }
  cudaMemsetAsync(h->delta, 0, sizeof(double) * (size_t)(h->n + (dist ? EKF_DIST_PAD_ROWS : 0)), st);
  for (int f0 = 0; f0 < cnt; f0 += EKF_UB / 2) {
    { ProfScope ps(h, 3); launch_blk_gather(st, h->Sigma, h->ld, r0, r1, h->ft, f0, cnt, h->delta, h->W, h->nu, &h->launches); }
    if (dist) {
      // S_b = H_b W_b + R needs the camera / feature rows of W_b, which live on several ranks: every rank sums the terms of
      // the rows it owns and the 128 x 128 partial blocks are all-reduced (128 KB instead of all-gathering the 12 MB panel)
      { ProfScope ps(h, 4); launch_blk_S_part(st, h->W, h->ft, f0, cnt, h->dcfg, r0, r1, h->rank == 0 ? 1 : 0, nullptr, nullptr, h->Lb, &h->launches); }
      { ProfScope ps(h, 11); if (ekf_dist_allreduce_sum(h, h->Lb, (size_t)EKF_UB * EKF_UB)) return (int)cudaErrorUnknown; }
      { ProfScope ps(h, 4); launch_blk_factor_only(st, h->Lb, h->nu, h->Dinv, h->Dblk, h->yb, h->ctl, &h->launches); }
    } else {
      ProfScope ps(h, 4);
      launch_blk_factor(st, h->W, h->ft, f0, cnt, h->nu, h->dcfg, h->Lb, h->Dinv, h->Dblk, h->yb, h->ctl, &h->launches);  /* Lb = S_b scratch, Dinv = L, Dblk = diagonal-block inverses */
    }
    { ProfScope ps(h, 5); launch_blk_V(st, h->W, r0, r1, h->Dinv, h->Dblk, h->yb, dist ? nullptr : h->delta, &h->launches); }
    if (dist) {
      { ProfScope ps(h, 11); if (ekf_dist_allgather_rows(h, h->W, rpr, EKF_UB)) return (int)cudaErrorUnknown; }
      // delta += V_b y_b for all rows from the gathered panel, the same kernel on every rank: replicas stay bit-identical
      { ProfScope ps(h, 5); launch_delta_rows(st, h->W, h->yb, h->delta, h->n, &h->launches); }
    }
    {
      ProfScope ps(h, 6);
      const int rc = launch_gemm_nt_sub(st, h->Sigma + (size_t)r0 * h->ld, h->ld, h->W + (size_t)r0 * EKF_UB, EKF_UB, h->W, EKF_UB,
                                        r1 - r0, h->n, EKF_UB, nullptr, dist ? 0 : h->lower_only, h->gemm_counters, &h->launches);
      if (rc) return rc;
    }
  }
  if (plane) {   // forsePlane rows as one more block (V:1250-1263)
    { ProfScope ps(h, 3); launch_plane_gather(st, h->Sigma, h->ld, r0, r1, h->mu, h->delta, h->W, h->nu, h->ctl, &h->launches); }
    if (dist) { ProfScope ps(h, 11); if (ekf_dist_allgather_rows(h, h->W, rpr, EKF_UB)) return (int)cudaErrorUnknown; }
    {
      ProfScope ps(h, 4);
      launch_plane_S(st, h->W, h->Lb, &h->launches);
      launch_blk_factor_only(st, h->Lb, h->nu, h->Dinv, h->Dblk, h->yb, h->ctl, &h->launches);
    }
    { ProfScope ps(h, 5); launch_blk_V(st, h->W, r0, r1, h->Dinv, h->Dblk, h->yb, h->delta, &h->launches); }
    if (dist) {
      ProfScope ps(h, 11);
      if (ekf_dist_allgather_rows(h, h->W, rpr, EKF_UB)) return (int)cudaErrorUnknown;
      if (ekf_dist_allgather_rows(h, h->delta, rpr, 1)) return (int)cudaErrorUnknown;
    }
    {
      ProfScope ps(h, 6);
      const int rc = launch_gemm_nt_sub(st, h->Sigma + (size_t)r0 * h->ld, h->ld, h->W + (size_t)r0 * EKF_UB, EKF_UB, h->W, EKF_UB,
                                        r1 - r0, h->n, EKF_UB, nullptr, dist ? 0 : h->lower_only, h->gemm_counters, &h->launches);
      if (rc) return rc;
    }
  }
  if (dist) { ProfScope ps(h, 11); if (ekf_dist_allgather_rows(h, h->Sigma, rpr, (size_t)h->ld)) return (int)cudaErrorUnknown; }
  {
    ProfScope ps(h, 7);
    launch_finish_update(st, h->Sigma, h->ld, h->n, h->mu, h->delta, h->ctl, &h->launches);
  }
  return 0;
}

int ekf_update_after_match(ekf_handle* h, const uint32_t* picks, int n_picks) {
  if (!h || n_picks < 0 || (n_picks > 0 && !picks)) return EKF_ERR_ARG;
  if (!h->predicted) return ekf_fail(h, EKF_ERR_STATE, "update before predict");
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  h->cam_cache_ok = false;
  cudaStream_t st = h->stream;
  // picks_dev is sized once by ekf_create (EKF_PICKS_CAP draws): no allocation inside the step
  if (n_picks > h->picks_cap) return ekf_fail(h, EKF_ERR_CAPACITY, "more RANSAC picks than EKF_PICKS_CAP: pass at most that many draws per update");
  if (n_picks > 0) EKF_CUDA_CHECK(cudaMemcpyAsync(h->picks_dev, picks, sizeof(uint32_t) * n_picks, cudaMemcpyHostToDevice, st));
  DevCtl& hc = *h->ctl_host;   // pinned: the read-backs below are true asynchronous copies followed by one stream wait
  // 1-point RANSAC, then the low-innovation update
  {
    ProfScope ps(h, 2);
    launch_ransac(st, h->Sigma, h->ld, h->n, h->mu, h->ft, h->N, h->ctl, h->dcfg, h->picks_dev, n_picks, h->mu_i, h->cand, &h->launches);
  }
  if (h->prelaunch_on && chain_short_likely(h)) chain_short_prelaunch(h);   // GPU work for the duration of the read-back
  else if (h->prelaunch_on && resident_chain_likely(h)) resident_chain_prelaunch(h);
  EKF_CUDA_CHECK(cudaMemcpyAsync(h->ctl_host, h->ctl, sizeof(DevCtl), cudaMemcpyDeviceToHost, st));
  EKF_CUDA_CHECK(cudaStreamSynchronize(st));
  const int n_li = hc.n_li;
  if (n_li <= 0 && h->prelaunched) {   // no update after all: keep later work on the main stream behind the speculative gathers
    cudaStreamWaitEvent(st, h->ev_gather[0], 0);
    cudaStreamWaitEvent(st, h->ev_gather[1], 0);
    h->prelaunched = false;
  }
  if (n_li > 0) {
    static const bool host_timing = getenv("EKF_HOST_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    int rc = stacked_update(h, n_li);
    if (host_timing) {
      const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
      fprintf(stderr, "stacked_update(li, %d rows): host enqueue %.1f us, %lld launches so far\n", n_li, us, h->launches);
    }
    if (rc) return ekf_fail_cuda(h, (cudaError_t)rc, "stacked update (li)", __FILE__, __LINE__);
  }
  // High-innovation rescue.  The common case is n_hi == 0 (nothing rescued) without forsePlane: there is no second update then, and
  // the last CTA of the rescue kernel itself does the step's book-keeping and writes the packed result record, which is copied
  // back BEFORE the host learns n_hi — the step ends with one synchronize and one launch less.  Otherwise the host runs the second
  // update and k_bookkeeping below.
  const bool plane = h->cfg.forsePlane != 0;
  int* outi_dev = reinterpret_cast<int*>(h->out_dev + 210);
  const size_t bytes = sizeof(double) * 210 + sizeof(int) * (16 + 3 * (size_t)h->N);
  {
    ProfScope ps(h, 8);
    launch_hi_rescue(st, h->Sigma, h->ld, h->mu, h->ft, h->N, h->ctl, h->dcfg, &h->launches, plane ? nullptr : h->out_dev, plane ? nullptr : outi_dev);
  }
  if (!plane) EKF_CUDA_CHECK(cudaMemcpyAsync(h->out_host, h->out_dev, bytes, cudaMemcpyDeviceToHost, st));
  EKF_CUDA_CHECK(cudaMemcpyAsync(h->ctl_host, h->ctl, sizeof(DevCtl), cudaMemcpyDeviceToHost, st));
  EKF_CUDA_CHECK(cudaStreamSynchronize(st));
  const int n_hi = hc.n_hi;
  if (n_hi > 0 || plane) {
    int rc = stacked_update(h, n_hi, plane);
    if (rc) return ekf_fail_cuda(h, (cudaError_t)rc, "stacked update (hi)", __FILE__, __LINE__);
    // book-keeping + packed result record
    {
      ProfScope ps(h, 9);
      launch_bookkeeping(st, h->Sigma, h->ld, h->mu, h->ft, h->N, h->ctl, h->dcfg, h->out_dev, outi_dev, &h->launches);
    }
    EKF_CUDA_CHECK(cudaGetLastError());
    EKF_CUDA_CHECK(cudaMemcpyAsync(h->out_host, h->out_dev, bytes, cudaMemcpyDeviceToHost, st));
    EKF_CUDA_CHECK(cudaStreamSynchronize(st));
  }
  EKF_CUDA_CHECK(cudaGetLastError());
  prof_flush(h);
  trace_flush();
  const int* outi = reinterpret_cast<const int*>(h->out_host + 210);
  h->stats.n_in_innovation_predict = outi[0];
  h->stats.n_matched = outi[1];
  h->stats.n_li = n_li;
  h->stats.n_hi = n_hi;
  h->stats.ransac_hypotheses = outi[4];
  h->stats.blur_requests = outi[6];
  if (outi[5] & 64) {
    char msg[160];
    snprintf(msg, sizeof msg, "pipelined update: a flag a kernel waits for did not arrive within the wait limit (code 0x%x)", outi[5]);
    return ekf_fail(h, EKF_ERR_CUDA, msg);
  }
  if (outi[5] & 16) return ekf_fail(h, EKF_ERR_CUDA, "row-block partition: a peer's panel did not arrive within the wait limit (peer-memory exchange)");
  if (outi[5]) return ekf_fail(h, EKF_ERR_STATE, "innovation covariance not positive definite");
  if (outi[7]) return ekf_fail(h, EKF_ERR_UNSUPPORTED, "a motion-blur kernel exceeded 256 x 256 pixels");
  h->cache_ok = false;
  // delete flagged features (V:1296-1299), then the visibility top-up hook (V:1301-1315)
  std::vector<int> victims;
  const int N = h->N;
  for (int i = 0; i < N; ++i)
    if (outi[16 + i] & 8) victims.push_back(i);
  int nvis = 0;
  for (int i = 0; i < N; ++i)
    if ((outi[16 + i] & 1) && !(outi[16 + i] & 8)) nvis++;
  h->stats.n_removed = (int)victims.size();
  int rc = remove_features(h, victims);
  if (rc) return rc;
  h->stats.topup_request = 0;
  if (nvis < h->cfg.min_features) {
    if (h->N > h->cfg.max_features) {
      rc = remove_features(h, std::vector<int>{0});
      if (rc) return rc;
      h->stats.n_removed += 1;
    }
    h->stats.topup_request = h->cfg.min_features - nvis;
    rc = ekf_find_new_features(h, h->stats.topup_request);   // vslamRansac.cpp:1314
    if (rc < 0) return rc;
  }
  if (h->cfg.xyz_conversion) {  // V:1317
    rc = convert_xyz(h, -1);
    if (rc) return rc;
  }
  h->stats.kernel_launches = h->launches;
  h->predicted = false;
  h->cam_cache_ok = true;
  return EKF_OK;
}

int ekf_update(ekf_handle* h, const uint32_t* picks, int n_picks) {
  int rc = ekf_match(h, nullptr);
  if (rc) return rc;
  return ekf_update_after_match(h, picks, n_picks);
}

int ekf_inject_match(ekf_handle* h, int idx, double zu, double zv, int accepted) {
  if (!h || idx < 0 || idx >= h->N) return EKF_ERR_ARG;
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  const int one = accepted ? 1 : 0, zero = 0;
  const double z[2] = {zu, zv};
  const float c[2] = {accepted ? (float)zu : -1.0f, accepted ? (float)zv : -1.0f};
  cudaStream_t st = h->stream;
  EKF_CUDA_CHECK(cudaMemcpyAsync(h->ft.innov + idx, &one, sizeof(int), cudaMemcpyHostToDevice, st));
  EKF_CUDA_CHECK(cudaMemcpyAsync(h->ft.li + idx, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
  EKF_CUDA_CHECK(cudaMemcpyAsync(h->ft.hi + idx, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
  if (accepted) EKF_CUDA_CHECK(cudaMemcpyAsync(h->ft.z + 2 * idx, z, sizeof z, cudaMemcpyHostToDevice, st));
  EKF_CUDA_CHECK(cudaMemcpyAsync(h->ft.center + 2 * idx, c, sizeof c, cudaMemcpyHostToDevice, st));
  EKF_CUDA_CHECK(cudaStreamSynchronize(st));
  h->cache_ok = false;
  return EKF_OK;
}

int ekf_add_feature(ekf_handle* h, float u, float v) {
  if (!h) return EKF_ERR_ARG;
  if (!h->have_frame) return ekf_fail(h, EKF_ERR_STATE, "addFeature before captureNewFrame");
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  const int half = h->cfg.window_size / 2;
  // isInsideImage (vslamRansac.cpp:314,1644-1652); in the fp64 build hd is (double)pf
  const double x = (double)u, y = (double)v;
  if (!(x > half && y > half && x < h->fv.w - half && y < h->fv.h - half)) return 0;
  if (h->N >= h->Ncap || h->n + 6 > h->ncap) return ekf_fail(h, EKF_ERR_CAPACITY, "feature capacity exceeded");
  frame_ready(h);
  launch_add_feature(h->stream, h->Sigma, h->ld, h->n, h->mu, h->ft, h->N, h->fv, h->dcfg, u, v, h->patchnumbre, &h->launches);
  EKF_CUDA_CHECK(cudaGetLastError());
  h->m_pos.push_back(h->n);
  h->m_coding.push_back(0);
  h->patchnumbre += 1;
  h->N += 1;
  h->n += 6;
  h->cache_ok = false;
  return 1;
}

int ekf_remove_feature(ekf_handle* h, int index) {
  if (!h || index < 0 || index >= h->N) return EKF_ERR_ARG;
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  return remove_features(h, std::vector<int>{index});
}

#define EKF_DET_KEYS 65536
#define EKF_DET_MAX 1024
// goodFeaturesToTrack(frame, corners, num, 0.01, 12, mask) with the reference's mask (vslamRansac.cpp:788-831)
static int detect_corners(ekf_handle* h, int num, std::vector<float>& xy) {
  xy.clear();
  if (!h->have_frame) return ekf_fail(h, EKF_ERR_STATE, "findNewFeatures before captureNewFrame");
  if (num <= 0) num = h->cfg.nInitFeatures;   // vslamRansac.cpp:786
  if (num > EKF_DET_MAX) num = EKF_DET_MAX;
  const size_t px = (size_t)h->fv.w * h->fv.h;
  cudaStream_t st = h->stream;
  frame_ready(h);
  if (px > h->det_cap) {
    EKF_CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(h->det_mask); cudaFree(h->det_eig);
    h->det_mask = nullptr; h->det_eig = nullptr;
    EKF_CUDA_CHECK(cudaMalloc((void**)&h->det_mask, px));
    EKF_CUDA_CHECK(cudaMalloc((void**)&h->det_eig, px * sizeof(float)));
    h->det_cap = px;
  }
  if (!h->det_keys) {
    EKF_CUDA_CHECK(cudaMalloc((void**)&h->det_keys, sizeof(unsigned long long) * EKF_DET_KEYS));
    EKF_CUDA_CHECK(cudaMalloc((void**)&h->det_counters, 4 * sizeof(int)));
    EKF_CUDA_CHECK(cudaMalloc((void**)&h->det_xy, 2 * EKF_DET_MAX * sizeof(float)));
  }
  launch_detect_corners(st, h->fv, h->ft, h->N, h->cfg.window_size, h->det_mask, h->det_eig, h->det_keys, EKF_DET_KEYS, h->det_counters,
                        num, h->det_xy, &h->launches);
  EKF_CUDA_CHECK(cudaGetLastError());
  int counters[4];
  EKF_CUDA_CHECK(cudaMemcpyAsync(counters, h->det_counters, sizeof counters, cudaMemcpyDeviceToHost, st));
  EKF_CUDA_CHECK(cudaStreamSynchronize(st));
  const int n = counters[2];
  xy.resize(2 * (size_t)n);
  if (n > 0) {
    EKF_CUDA_CHECK(cudaMemcpyAsync(xy.data(), h->det_xy, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, st));
    EKF_CUDA_CHECK(cudaStreamSynchronize(st));
  }
  return EKF_OK;
}

int ekf_detect_corners(ekf_handle* h, int num, float* out_xy, int* out_n) {
  if (!h || !out_xy || !out_n) return EKF_ERR_ARG;
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  std::vector<float> xy;
  const int rc = detect_corners(h, num, xy);
  if (rc) return rc;
  *out_n = (int)(xy.size() / 2);
  for (size_t i = 0; i < xy.size(); ++i) out_xy[i] = xy[i];
  return EKF_OK;
}

int ekf_find_new_features(ekf_handle* h, int num) {
  if (!h) return EKF_ERR_ARG;
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  std::vector<float> xy;
  int rc = detect_corners(h, num, xy);
  if (rc) return rc;
  int added = 0;
  for (size_t i = 0; i + 1 < xy.size(); i += 2) {   // vslamRansac.cpp:832-835
    if (h->N >= h->Ncap) break;
    rc = ekf_add_feature(h, xy[i], xy[i + 1]);
    if (rc < 0) return rc;
    added += rc;
  }
  return added;
}

int ekf_convert2xyz_if_linear(ekf_handle* h, int index) {
  if (!h || index < 0 || index >= h->N) return EKF_ERR_ARG;
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  if (h->m_coding[index]) return EKF_OK;
  return convert_xyz(h, index);
}
int ekf_convert2xyz_if_linear_all(ekf_handle* h) {
  if (!h) return EKF_ERR_ARG;
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  return convert_xyz(h, -1);
}

int ekf_num_features(const ekf_handle* h) { return h ? h->N : EKF_ERR_ARG; }
int ekf_state_dim(const ekf_handle* h) { return h ? h->n : EKF_ERR_ARG; }
double ekf_get_dt(const ekf_handle* h) { return h ? h->dT : 0.0; }

int ekf_get_state(ekf_handle* h, double out[EKF_STATE_DIM]) {
  if (!h || !out) return EKF_ERR_ARG;
  // The packed result record of update() already carries the camera state and its covariance block (k_bookkeeping), and nothing
  // between two predicts changes them (removeFeature / addFeature / XYZ conversion touch feature rows only): answer from the
  // host copy instead of another device round trip.
  if (h->cam_cache_ok) { memcpy(out, h->out_host, sizeof(double) * EKF_CAM); return EKF_OK; }
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  EKF_CUDA_CHECK(cudaMemcpyAsync(out, h->mu, sizeof(double) * EKF_CAM, cudaMemcpyDeviceToHost, h->stream));
  EKF_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return EKF_OK;
}
int ekf_get_sigma(ekf_handle* h, double out[EKF_STATE_DIM * EKF_STATE_DIM]) {
  if (!h || !out) return EKF_ERR_ARG;
  if (h->cam_cache_ok) { memcpy(out, h->out_host + EKF_CAM, sizeof(double) * EKF_CAM * EKF_CAM); return EKF_OK; }
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  EKF_CUDA_CHECK(cudaMemcpy2DAsync(out, sizeof(double) * EKF_CAM, h->Sigma, sizeof(double) * h->ld, sizeof(double) * EKF_CAM, EKF_CAM,
                                   cudaMemcpyDeviceToHost, h->stream));
  EKF_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return EKF_OK;
}
int ekf_covariance_parameter(ekf_handle* h, double* out) {
  if (!h || !out) return EKF_ERR_ARG;
  double S[EKF_CAM * EKF_CAM];
  int rc = ekf_get_sigma(h, S);
  if (rc) return rc;
  double P = 0;  // vslamRansac.cpp:854-855
  P += S[0 * 14 + 0] + S[1 * 14 + 1] + S[2 * 14 + 2];
  P += S[4 * 14 + 4] + S[5 * 14 + 5] + S[6 * 14 + 6] + S[3 * 14 + 3];
  *out = P;
  return EKF_OK;
}
int ekf_num_deleted(const ekf_handle* h) { return h ? (int)h->deleted.size() : EKF_ERR_ARG; }
int ekf_get_deleted(ekf_handle* h, int i, ekf_deleted_info* out) {
  if (!h || !out || i < 0 || i >= (int)h->deleted.size()) return EKF_ERR_ARG;
  const DeletedPatch& dp = h->deleted[i];
  out->real_index = dp.real_index; out->_pad = 0;
  for (int c = 0; c < 3; ++c) out->xyz_pos[c] = dp.XYZ_pos[c];
  for (int c = 0; c < 9; ++c) out->cov_4_delete[c] = dp.cov_4_delete[c];
  return EKF_OK;
}
int ekf_get_points_features(ekf_handle* h, double* out, int rows_cap, int* rows) {
  if (!h || !rows) return EKF_ERR_ARG;
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  int rc = refresh_cache(h);
  if (rc) return rc;
  const int psize = h->N > 0 ? h->c_real[h->N - 1] : 0;   // R:349 (the reference reads patches[size-1] unguarded)
  *rows = psize + 1;
  if (!out) return EKF_OK;
  if (rows_cap < psize + 1) return EKF_ERR_ARG;
  const size_t bytes = sizeof(double) * 12 * (size_t)(psize + 1);
  double* dev = nullptr;
  EKF_CUDA_CHECK(cudaMallocAsync((void**)&dev, bytes, h->stream));
  EKF_CUDA_CHECK(cudaMemsetAsync(dev, 0, bytes, h->stream));
  launch_points_features(h->stream, h->Sigma, h->ld, h->mu, h->ft, h->N, dev, psize + 1, &h->launches);
  EKF_CUDA_CHECK(cudaGetLastError());
  EKF_CUDA_CHECK(cudaMemcpyAsync(out, dev, bytes, cudaMemcpyDeviceToHost, h->stream));
  double map_scale = 1.0;
  EKF_CUDA_CHECK(cudaMemcpyAsync(&map_scale, h->mu + 13, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  EKF_CUDA_CHECK(cudaFreeAsync(dev, h->stream));
  EKF_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  for (const DeletedPatch& dp : h->deleted) {             // R:397-416, host-resident archive
    if (dp.real_index < 0 || dp.real_index > psize) continue;
    double* o = out + (size_t)dp.real_index * 12;
    for (int c = 0; c < 9; ++c) o[3 + c] = dp.cov_4_delete[c];
    for (int c = 0; c < 3; ++c) o[c] = dp.XYZ_pos[c] * map_scale;
  }
  if (h->deleted.size() > 7000) h->deleted.clear();
  return EKF_OK;
}
int ekf_rts_epoch(ekf_handle* h, double mu[13], double sigma[169], const double mu_s[13], const double sigma_s[169],
                  const double dTspeed[3], const double dRspeed[3], double deltaT) {
  if (!h || !mu || !sigma || !mu_s || !sigma_s || !dTspeed || !dRspeed || !(deltaT > 0)) return EKF_ERR_ARG;
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  double io[370];
  memcpy(io, mu, 13 * sizeof(double)); memcpy(io + 13, sigma, 169 * sizeof(double));
  memcpy(io + 182, mu_s, 13 * sizeof(double)); memcpy(io + 195, sigma_s, 169 * sizeof(double));
  memcpy(io + 364, dTspeed, 3 * sizeof(double)); memcpy(io + 367, dRspeed, 3 * sizeof(double));
  double* dev = nullptr;
  EKF_CUDA_CHECK(cudaMallocAsync((void**)&dev, sizeof(io) + 16, h->stream));
  int* flag = reinterpret_cast<int*>(dev + 370);
  EKF_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), h->stream));
  EKF_CUDA_CHECK(cudaMemcpyAsync(dev, io, sizeof(io), cudaMemcpyHostToDevice, h->stream));
  launch_rts_epoch(h->stream, dev, h->dcfg, deltaT, flag, &h->launches);
  EKF_CUDA_CHECK(cudaGetLastError());
  int singular = 0;
  EKF_CUDA_CHECK(cudaMemcpyAsync(io, dev, 182 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  EKF_CUDA_CHECK(cudaMemcpyAsync(&singular, flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  EKF_CUDA_CHECK(cudaFreeAsync(dev, h->stream));
  EKF_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  if (singular) return ekf_fail(h, EKF_ERR_STATE, "rts_epoch: predicted covariance is singular");
  memcpy(mu, io, 13 * sizeof(double)); memcpy(sigma, io + 13, 169 * sizeof(double));
  return EKF_OK;
}
int ekf_get_center(ekf_handle* h, int idx, float out[2]) {
  if (!h || !out || idx < 0 || idx >= h->N) return EKF_ERR_ARG;
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  int rc = refresh_cache(h);
  if (rc) return rc;
  out[0] = h->c_center[2 * idx]; out[1] = h->c_center[2 * idx + 1];
  return EKF_OK;
}
int ekf_get_feature(ekf_handle* h, int idx, ekf_feature_info* o) {
  if (!h || !o || idx < 0 || idx >= h->N) return EKF_ERR_ARG;
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  int rc = refresh_cache(h);
  if (rc) return rc;
  memset(o, 0, sizeof *o);
  const int pos = h->c_pos[idx], fs = h->c_coding[idx] ? 3 : 6;
  o->position_in_state = pos; o->position_in_z = h->c_posz[idx]; o->coding = h->c_coding[idx];
  o->n_tot = h->c_ntot[idx]; o->n_find = h->c_nfind[idx]; o->real_index = h->c_real[idx];
  o->is_in_innovation = h->c_innov[idx]; o->is_in_li = h->c_li[idx]; o->is_in_hi = h->c_hi[idx]; o->remove_flag = h->c_removef[idx];
  o->center[0] = h->c_center[2 * idx]; o->center[1] = h->c_center[2 * idx + 1];
  o->quality_index = h->c_quality[idx]; o->last_ncc = h->c_ncc[idx];
  for (int a = 0; a < 2; ++a) { o->z[a] = h->c_z[2 * idx + a]; o->h[a] = h->c_h[2 * idx + a]; }
  for (int a = 0; a < 26; ++a) o->H[a] = h->c_Hc[26 * idx + a];
  EKF_CUDA_CHECK(cudaMemcpyAsync(o->state, h->mu + pos, sizeof(double) * fs, cudaMemcpyDeviceToHost, h->stream));
  double blk[36];
  EKF_CUDA_CHECK(cudaMemcpy2DAsync(blk, sizeof(double) * fs, h->Sigma + (size_t)pos * h->ld + pos, sizeof(double) * h->ld,
                                   sizeof(double) * fs, fs, cudaMemcpyDeviceToHost, h->stream));
  EKF_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  for (int a = 0; a < fs; ++a)
    for (int b = 0; b < fs; ++b) o->cov[a * 6 + b] = blk[a * fs + b];
  return EKF_OK;
}
int ekf_get_template(ekf_handle* h, int idx, int which, uint8_t* out) {
  if (!h || !out || idx < 0 || idx >= h->N) return EKF_ERR_ARG;
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  const int w2 = h->cfg.window_size * h->cfg.window_size;
  const uint8_t* src = (which ? h->ft.mpatch : h->ft.patch) + (size_t)idx * h->dcfg.tstride;
  EKF_CUDA_CHECK(cudaMemcpyAsync(out, src, w2, cudaMemcpyDeviceToHost, h->stream));
  EKF_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return EKF_OK;
}
int ekf_get_step_stats(ekf_handle* h, ekf_step_stats* out) {
  if (!h || !out) return EKF_ERR_ARG;
  h->stats.kernel_launches = h->launches;
  *out = h->stats;
  return EKF_OK;
}
int ekf_set_profiling(ekf_handle* h, int on) {
  if (!h) return EKF_ERR_ARG;
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  EKF_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  prof_flush(h);
  h->prof_on = on != 0;
  return EKF_OK;
}
int ekf_get_profile(ekf_handle* h, ekf_profile* out, int reset) {
  if (!h || !out) return EKF_ERR_ARG;
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  EKF_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  prof_flush(h);
  for (int c = 0; c < EKF_PROF_CLASSES; ++c) { out->ms[c] = h->prof_ms[c]; out->launches[c] = h->prof_launches[c]; }
  if (reset) for (int c = 0; c < EKF_PROF_CLASSES; ++c) { h->prof_ms[c] = 0; h->prof_launches[c] = 0; }
  return EKF_OK;
}
int ekf_set_symmetric_downdate(ekf_handle* h, int on) {
  if (!h) return EKF_ERR_ARG;
  h->lower_only = on ? 1 : 0;
  return EKF_OK;
}
int ekf_debug_time_downdate(ekf_handle* h, int reps, float* ms_per_launch) {
  if (!h || !ms_per_launch || reps < 1) return EKF_ERR_ARG;
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  EKF_CUDA_CHECK(cudaStreamSynchronize(st));
  // a small panel (entries 1e-6 x a row / column pattern) so that the covariance stays positive definite whatever it holds
  std::vector<double> v((size_t)h->n * EKF_UB);
  for (int i = 0; i < h->n; ++i)
    for (int k = 0; k < EKF_UB; ++k) v[(size_t)i * EKF_UB + k] = 1e-9 * (double)(((i * 131 + k * 17) % 97) - 48);
  EKF_CUDA_CHECK(cudaMemcpyAsync(h->Wbuf[1], v.data(), sizeof(double) * v.size(), cudaMemcpyHostToDevice, st));
  cudaEvent_t a, b;
  EKF_CUDA_CHECK(cudaEventCreate(&a)); EKF_CUDA_CHECK(cudaEventCreate(&b));
  for (int i = 0; i < 2; ++i) {
    const int rc = launch_gemm_nt_sub(st, h->Sigma, h->ld, h->Wbuf[1], EKF_UB, h->Wbuf[1], EKF_UB, h->n, h->n, EKF_UB, nullptr, h->lower_only, h->gemm_counters, &h->launches);
    if (rc) return ekf_fail_cuda(h, (cudaError_t)rc, "downdate (timing)", __FILE__, __LINE__);
  }
  EKF_CUDA_CHECK(cudaEventRecord(a, st));
  for (int i = 0; i < reps; ++i) {
    const int rc = launch_gemm_nt_sub(st, h->Sigma, h->ld, h->Wbuf[1], EKF_UB, h->Wbuf[1], EKF_UB, h->n, h->n, EKF_UB, nullptr, h->lower_only, h->gemm_counters, &h->launches);
    if (rc) return ekf_fail_cuda(h, (cudaError_t)rc, "downdate (timing)", __FILE__, __LINE__);
  }
  EKF_CUDA_CHECK(cudaEventRecord(b, st));
  EKF_CUDA_CHECK(cudaStreamSynchronize(st));
  float ms = 0;
  EKF_CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
  cudaEventDestroy(a); cudaEventDestroy(b);
  *ms_per_launch = ms / (float)reps;
  h->cam_cache_ok = false;
  return EKF_OK;
}
int ekf_get_full(ekf_handle* h, double* mu, double* sigma, int ld) {
  if (!h || !mu || (sigma && ld < h->n)) return EKF_ERR_ARG;
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  EKF_CUDA_CHECK(cudaMemcpyAsync(mu, h->mu, sizeof(double) * h->n, cudaMemcpyDeviceToHost, h->stream));
  if (sigma)
    EKF_CUDA_CHECK(cudaMemcpy2DAsync(sigma, sizeof(double) * ld, h->Sigma, sizeof(double) * h->ld, sizeof(double) * h->n, h->n,
                                     cudaMemcpyDeviceToHost, h->stream));
  EKF_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return EKF_OK;
}
int ekf_set_full(ekf_handle* h, const double* mu, const double* sigma, int ld) {
  if (!h || !mu || !sigma || ld < h->n) return EKF_ERR_ARG;
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  h->cam_cache_ok = false;
  EKF_CUDA_CHECK(cudaMemcpyAsync(h->mu, mu, sizeof(double) * h->n, cudaMemcpyHostToDevice, h->stream));
  EKF_CUDA_CHECK(cudaMemcpy2DAsync(h->Sigma, sizeof(double) * h->ld, sigma, sizeof(double) * ld, sizeof(double) * h->n, h->n,
                                   cudaMemcpyHostToDevice, h->stream));
  EKF_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  return EKF_OK;
}
int ekf_get_S_blocks(ekf_handle* h, double* out) {
  if (!h || !out) return EKF_ERR_ARG;
  EKF_CUDA_CHECK(cudaSetDevice(h->device));
  int rc = refresh_cache(h);
  if (rc) return rc;
  for (int i = 0; i < h->N; ++i)
    for (int c = 0; c < 4; ++c) out[4 * i + c] = h->c_innov[i] ? h->c_S2[4 * i + c] : 0.0;
  return EKF_OK;
}

int ekf_match_batch(const uint8_t* frames, int n_frames, int width, int height, int stride, const uint8_t* templates,
                    int features_per_frame, int window_size, const double* hh, const double* S, float sigma_size,
                    float ncc_threshold, float search_clamp, int32_t* out_uv, float* out_score, void* stream) {
  if (!frames || !templates || !hh || !S || !out_uv || !out_score || n_frames < 0 || features_per_frame < 0) return EKF_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return EKF_ERR_CUDA;
  const int rc = launch_match_batch((cudaStream_t)stream, frames, n_frames, width, height, stride, templates, features_per_frame,
                                    window_size, hh, S, sigma_size, ncc_threshold, search_clamp, out_uv, out_score);
  if (rc < 0) return EKF_ERR_UNSUPPORTED;
  return cudaGetLastError() == cudaSuccess ? EKF_OK : EKF_ERR_CUDA;
}

}  // extern "C"
