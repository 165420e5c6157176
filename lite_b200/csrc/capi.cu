// capi.cu — implementation of include/lite_b200.h (host side: context, upload,
// kernel launches, NCCL allreduce).  No CPU fallback: every compute entry point
// needs a CUDA device.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <emmintrin.h>
#include <nvtx3/nvToolsExt.h>

#include "kernels.cuh"
#include "gen_launch.cuh"

// NVTX range over one C-ABI call (header-only NVTX 3: a no-op unless a profiler is attached)
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

// ---- minimal NCCL surface, resolved at run time (libnccl.so.2) --------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclFloat64 = 8 };
enum { ncclSum = 0 };
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static const bool g_trace = getenv("LITE_B200_TRACE") != nullptr;
static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CU(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess)                                                            \
      return fail(LITE_ECUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

static int nccl_load() {
  if (g_nccl.lib) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *lib = nullptr;
  for (const char *n : names) {
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) return fail(LITE_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(lib, "ncclCommInitRank");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(lib, "ncclCommDestroy");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(lib, "ncclAllReduce");
  g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(lib, "ncclAllGather");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce)
    return fail(LITE_ENCCL, "libnccl.so.2 lacks the expected symbols");
  g_nccl.lib = lib;
  return 0;
}

// layout of the result block (doubles): two 128-word sums, 8 status words, 2 exchange words, the validation word
#define LITE_RES_STATUS 256
#define LITE_RES_PEER 264
#define LITE_RES_ERR 266
#define LITE_RES_WORDS 268

template <typename T>
struct DBuf {
  T *p = nullptr;
  size_t n = 0;
  bool view = false;  // points into the upload arena, not owned
  int alloc(size_t count) {
    if (count <= n && p && !view) return 0;
    if (p && !view) cudaFree(p);
    view = false;
    p = nullptr; n = 0;
    if (count == 0) return 0;
    cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
    if (e != cudaSuccess) return fail(LITE_ECUDA, "cudaMalloc(%zu B): %s", count * sizeof(T), cudaGetErrorString(e));
    n = count;
    return 0;
  }
  void release() { if (p && !view) cudaFree(p); p = nullptr; n = 0; view = false; }
};

// One device block + one pinned host block for everything lite_tess_upload copies:
// the host arrays are packed into the pinned block and cross PCIe in a single copy.
struct Arena {
  char *d = nullptr, *h = nullptr;
  size_t cap = 0, used = 0;
};
static inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

// A few persistent host threads per context: an upload's packing and export assertions run
// on them (spawning threads per call cost about as much as the work itself).
static double now_ms();
struct HostPool {
  std::vector<std::thread> th;
  std::mutex m;
  std::condition_variable cv_go, cv_done;
  std::function<void(int)> job;
  std::atomic<unsigned long> gen{0};
  std::atomic<int> pending{0};
  bool stop = false;
  // A worker that has just finished a job polls for the next one for this long before it goes
  // to sleep on the condition variable: in a loop of uploads (one every few hundred
  // microseconds) the workers then start within a microsecond instead of after a futex wake-up
  // each (tens of microseconds, one after the other).  LITE_B200_SPIN_US overrides it, 0 = never poll.
  double spin_ms = 1.0;
  void start(int n) {
    if (const char *e = getenv("LITE_B200_SPIN_US")) spin_ms = atof(e) * 1e-3;
    for (int i = 0; i < n; i++)
      th.emplace_back([this, i] {
        unsigned long seen = 0;
        bool hot = false;
        for (;;) {
          if (hot && spin_ms > 0) {
            const double t_end = now_ms() + spin_ms;
            for (int k = 0; gen.load(std::memory_order_acquire) == seen; k++) {
              _mm_pause();
              if ((k & 63) == 63 && now_ms() > t_end) break;
            }
          }
          std::function<void(int)> f;
          {
            std::unique_lock<std::mutex> lk(m);
            cv_go.wait(lk, [&] { return stop || gen.load(std::memory_order_relaxed) != seen; });
            if (stop) return;
            seen = gen.load(std::memory_order_relaxed);
            f = job;
          }
          f(i);
          hot = true;
          if (pending.fetch_sub(1, std::memory_order_acq_rel) == 1) {
            std::lock_guard<std::mutex> lk(m);
            cv_done.notify_all();
          }
        }
      });
  }
  // run f(0..n-1) on the workers; returns at once, wait() joins
  void run(const std::function<void(int)> &f) {
    {
      std::lock_guard<std::mutex> lk(m);
      job = f; pending.store((int)th.size(), std::memory_order_relaxed);
      gen.fetch_add(1, std::memory_order_release);
    }
    cv_go.notify_all();
  }
  void wait() {
    for (int k = 0; k < 20000 && pending.load(std::memory_order_acquire) != 0; k++) _mm_pause();
    std::unique_lock<std::mutex> lk(m);
    cv_done.wait(lk, [&] { return pending.load(std::memory_order_acquire) == 0; });
  }
  void shutdown() {
    { std::lock_guard<std::mutex> lk(m); stop = true; }
    cv_go.notify_all();
    for (auto &t : th) if (t.joinable()) t.join();
    th.clear();
  }
};

struct SmfState;  // ensemble of SMF chains, smf_ensemble.cuh

struct lite_ctx {
  int device = 0, rank = 0, world = 1;
  SmfState *smf = nullptr;
  HostPool *pool = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;   // merges+flips run here, beside the split kernel
  cudaEvent_t ev_fork = nullptr, ev_mf = nullptr;
  cudaEvent_t ev_up0 = nullptr, ev_guide = nullptr;  // upload: the sampler's tables go over the second stream
  bool mf_inflight = false;         // stream must wait for ev_mf before using the flip statistics
  bool fork_pending = false;        // an upload since the merge/flip stream last waited for ev_fork
  ncclComm_t comm = nullptr;
  // cross-GPU sum over peer memory (kernels.cuh: PeerArgs); off -> ncclAllReduce
  bool peer_on = false;
  double *peer_local = nullptr;                 // this rank's mailbox
  double *peer_box[LITE_MAX_WORLD] = {};        // every rank's mailbox as mapped here
  bool peer_ipc[LITE_MAX_WORLD] = {};           // mapped with cudaIpcOpenMemHandle (to be closed)
  unsigned long long peer_seq = 0;
  DBuf<unsigned long long> peer_info;           // [0] timeout marker, [1] wait ns
  double peer_wait_us = 0;
  float ms_exchange = 0;                        // NCCL path: device time of the allreduce
  bool have_tess = false, have_prog = false, have_mf = false;
  bool mf_deferred = false;  // lite_candidates_merges_flips was called, its kernel is not enqueued yet
  bool record = false;
  bool validate_host = false;  // LITE_OPT_VALIDATE_HOST: export assertions on the host, errors returned by lite_tess_upload itself
  bool iid = false;         // LITE_OPT_SAMPLER: i.i.d. positions + sort instead of the stratified sampler
  DBuf<double> sel_a, sel_b;   // positions of the current batch shard, unsorted / sorted
  DBuf<long long> id_a, id_b;  // their batch indices
  DBuf<char> sort_tmp;
  bool mf_pending = false;  // sums of the last merges_flips still in flight
  bool rev = false;  // direction of the next theta-reduce pass over the split statistics
  int64_t launches = 0;
  // raw SoA
  DBuf<double> vx, vy, he_length;
  DBuf<int> he_src, he_twin, he_next, he_face, he_seg, he_next_hf, he_prev_hf;
  DBuf<uint8_t> he_dir, face_inside, seg_bnd;
  DBuf<int> f_off, f_cnt, f_list, seg_he_start, seg_he_end, seg_n_edges, nb_list, b_list;
  // derived
  DBuf<double2> pt, udir;
  DBuf<double> ang, len, ca, el, f_area, f_perim, f_sumang, f_minang, cum, pre, plen, f_tot2, stot2;
  DBuf<int2> sfo;
  DBuf<uint32_t> flg;
  DBuf<int> slot_he, slot_seg, slot_face, he_slot, guide, err;
  // faces that are not simple polygons (windows with holes): gen_face.h / gen_launch.cuh
  int n_general = 0;                 // how many the uploaded tessellation has
  gen::Cycles cyc;                   // their boundary cycles (host)
  std::vector<int> x_off, x_cnt, x_list;   // per-face halfedge lists extended by the inner boundaries
  std::vector<uint8_t> x_gen;
  DBuf<int> g_outer, g_hole_off, g_hole_cnt, g_hole_he;
  DBuf<uint8_t> g_fgen;
  gen::Tess GT{};
  gen::Pt *gen_scratch = nullptr;    // LITE_GEN_PTS points per thread of the general kernels
  int gen_blocks = 0;
  Arena arena;
  struct PackJob { char *dst; const void *src; size_t bytes; };
  std::vector<PackJob> pack_jobs;  // host->pinned copies of one upload, run on a few threads
  TessDev T{};
  int n_nb = 0, n_b = 0;
  double int_length = 0;
  Prog G{};
  // splits
  DBuf<double> stat;
  size_t cap = 0;
  int64_t n_local = 0, n_global = 0;
  DBuf<int> r_e1, r_e2, in_he;
  DBuf<double2> r_p1, r_p2;
  DBuf<double> r_x, r_angle, r_u, in_x, in_angle;
  size_t rcap = 0;
  // merges / flips
  DBuf<double> mstat, fstat, sums;  // sums: per-block column sums of mstat then fstat, [blocks][LITE_MAX_D]
  double *h_part = nullptr; size_t h_part_cap = 0;  // pinned copy of them
  int mf_gm = 0, mf_blocks = 0;
  DBuf<int> m_he, fl_e1, fl_e2, fl_right;
  DBuf<double2> fl_p2;
  std::vector<double> h_sum_m, h_sum_f;
  // reduce
  DBuf<double> partials, red_out;
  DBuf<double> res;  // one block behind red_out | status | peer_info | err: a single copy brings an evaluation back
  DBuf<unsigned int> counter;
  DBuf<unsigned long long> status;  // [0] split no-exit, [1] flip no-exit, [2..3] checksum
  double *h_pinned = nullptr;       // pinned staging for results
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t tev[LITE_TIMER_SLOTS] = {};  // user timer marks (lite_timer_mark)
  float ms_split = 0, ms_eval = 0, ms_mf = 0, ms_reduce = 0;
  double acc_ms[5] = {0, 0, 0, 0, 0};  // sums over evaluations: split kernel, reduce kernel, merges+flips, exchange, count
  int64_t fail_splits = 0, fail_flips = 0;  // no-exit counts behind the last evaluation
  bool ev01 = false;  // ev[0]/ev[1] have been recorded
  std::vector<uint8_t> h_he_inside;
  std::vector<double> h_cum;        // cumulative slot lengths of the last upload (host copy)
  cudaEvent_t ev_upload = nullptr;  // the H2D copy of the last upload has left the pinned block
  cudaEvent_t ev_tr[3] = {nullptr, nullptr, nullptr};  // LITE_B200_TRACE: start of the copy, end of the copy, end of the preparation
  bool upload_inflight = false;
  bool early_unfenced = false;  // an upload that failed after its bulk copy had been started: no ev_upload covers that copy
  int sm_count = 148;
};

extern "C" const char *lite_last_error(void) { return g_err.c_str(); }
// shared with host_capi.cpp
extern "C" int lite_set_error_(int code, const char *msg) { return fail(code, "%s", msg); }

extern "C" int lite_nccl_unique_id(void *out128) {
  if (!out128) return fail(LITE_EINVAL, "null output");
  int r = nccl_load();
  if (r) return r;
  ncclUniqueId id;
  ncclResult_t e = g_nccl.GetUniqueId(&id);
  if (e != 0) return fail(LITE_ENCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "?");
  memcpy(out128, &id, 128);
  return 0;
}

// Map every rank's mailbox into this process.  Collective over the context's NCCL
// communicator (used here for the plumbing only: one all-gather of the IPC handles, one
// all-reduce so that all ranks agree on the outcome).  Any failure leaves peer_on false and
// lite_pl_eval uses ncclAllReduce.
struct PeerCard {
  cudaIpcMemHandle_t handle;
  long long pid;
  unsigned long long ptr;
  int device, ok;
  char host[64];
};
static int peer_setup(lite_ctx *c) {
  const char *force = getenv("LITE_B200_ALLREDUCE");
  const bool want = c->world <= LITE_MAX_WORLD && g_nccl.AllGather && !(force && !strcmp(force, "nccl"));
  PeerCard mine{};
  mine.pid = (long long)getpid(); mine.device = c->device; mine.ok = 0;
  gethostname(mine.host, sizeof mine.host - 1);
  if (want && cudaMalloc((void **)&c->peer_local, LITE_BOX_DOUBLES * sizeof(double)) == cudaSuccess) {
    cudaMemsetAsync(c->peer_local, 0, LITE_BOX_DOUBLES * sizeof(double), c->stream);
    if (cudaIpcGetMemHandle(&mine.handle, c->peer_local) == cudaSuccess) mine.ok = 1;
    mine.ptr = (unsigned long long)c->peer_local;
  }
  (void)cudaGetLastError();
  DBuf<char> dsend, drecv;
  if (dsend.alloc(sizeof(PeerCard)) || drecv.alloc(sizeof(PeerCard) * c->world)) return LITE_ECUDA;
  std::vector<PeerCard> cards(c->world);
  CU(cudaMemcpyAsync(dsend.p, &mine, sizeof mine, cudaMemcpyHostToDevice, c->stream));
  if (!g_nccl.AllGather) { dsend.release(); drecv.release(); return 0; }
  ncclResult_t ne = g_nccl.AllGather(dsend.p, drecv.p, sizeof(PeerCard), ncclInt8, c->comm, c->stream);
  if (ne != 0) return fail(LITE_ENCCL, "ncclAllGather: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(ne) : "?");
  CU(cudaMemcpyAsync(cards.data(), drecv.p, sizeof(PeerCard) * c->world, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  int ok = mine.ok;
  for (int r = 0; r < c->world && ok; r++) {
    const PeerCard &k = cards[r];
    if (!k.ok || strncmp(k.host, mine.host, sizeof mine.host)) { ok = 0; break; }
    if (r == c->rank) { c->peer_box[r] = c->peer_local; continue; }
    if (k.pid == mine.pid) {  // another context of this process: address it directly
      int can = 0;
      if (k.device != c->device) {
        if (cudaDeviceCanAccessPeer(&can, c->device, k.device) != cudaSuccess || !can) { ok = 0; break; }
        cudaError_t e = cudaDeviceEnablePeerAccess(k.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { ok = 0; break; }
        (void)cudaGetLastError();
      }
      c->peer_box[r] = (double *)k.ptr;
    } else {
      void *q = nullptr;
      if (cudaIpcOpenMemHandle(&q, k.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
      c->peer_box[r] = (double *)q; c->peer_ipc[r] = true;
    }
  }
  (void)cudaGetLastError();
  // all or nothing: the ranks must take the same path
  double votes = ok ? 1.0 : 0.0, *dv = (double *)dsend.p;
  CU(cudaMemcpyAsync(dv, &votes, sizeof votes, cudaMemcpyHostToDevice, c->stream));
  ne = g_nccl.AllReduce(dv, dv, 1, ncclFloat64, ncclSum, c->comm, c->stream);
  if (ne != 0) return fail(LITE_ENCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(ne) : "?");
  CU(cudaMemcpyAsync(&votes, dv, sizeof votes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->peer_on = (votes == (double)c->world);
  dsend.release(); drecv.release();
  if (g_trace) fprintf(stderr, "[lite_b200] rank %d/%d: cross-GPU sum over %s\n", c->rank, c->world,
                       c->peer_on ? "peer memory (NVLink)" : "ncclAllReduce");
  return 0;
}
static void peer_release(lite_ctx *c) {
  for (int r = 0; r < LITE_MAX_WORLD; r++) {
    if (c->peer_ipc[r] && c->peer_box[r]) cudaIpcCloseMemHandle(c->peer_box[r]);
    c->peer_box[r] = nullptr; c->peer_ipc[r] = false;
  }
  if (c->peer_local) cudaFree(c->peer_local);
  c->peer_local = nullptr; c->peer_on = false;
}

static int ctx_init(lite_ctx *c, const void *nccl_unique_id) {
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_up0, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_guide, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_upload, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_mf, cudaEventDisableTiming));
  for (int i = 0; i < 8; i++) CU(cudaEventCreate(&c->ev[i]));
  for (int i = 0; i < LITE_TIMER_SLOTS; i++) CU(cudaEventCreate(&c->tev[i]));
  CU(cudaMallocHost((void **)&c->h_pinned, 4096 * sizeof(double)));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, c->device));
  c->sm_count = prop.multiProcessorCount;
  if (c->res.alloc(LITE_RES_WORDS) || c->counter.alloc(4)) return LITE_ECUDA;
  c->red_out.p = c->res.p; c->red_out.n = 256; c->red_out.view = true;
  c->status.p = (unsigned long long *)(c->res.p + LITE_RES_STATUS); c->status.n = 8; c->status.view = true;
  c->peer_info.p = (unsigned long long *)(c->res.p + LITE_RES_PEER); c->peer_info.n = 2; c->peer_info.view = true;
  c->err.p = (int *)(c->res.p + LITE_RES_ERR); c->err.n = 2; c->err.view = true;
  CU(cudaMemsetAsync(c->res.p, 0, LITE_RES_WORDS * sizeof(double), c->stream));
  CU(cudaMemsetAsync(c->status.p, 0, 8 * sizeof(unsigned long long), c->stream));
  CU(cudaMemsetAsync(c->counter.p, 0, 4 * sizeof(unsigned int), c->stream));
  CU(cudaMemsetAsync(c->peer_info.p, 0, 2 * sizeof(unsigned long long), c->stream));
  if (c->world > 1) {
    if (!nccl_unique_id) return fail(LITE_EINVAL, "world_size > 1 needs a NCCL unique id");
    int r = nccl_load();
    if (r) return r;
    ncclUniqueId id;
    memcpy(&id, nccl_unique_id, 128);
    ncclResult_t ne = g_nccl.CommInitRank(&c->comm, c->world, id, c->rank);
    if (ne != 0) return fail(LITE_ENCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(ne) : "?");
    r = peer_setup(c);
    if (r) return r;
  }
  c->pool = new HostPool();
  {
    // packing threads: four, fewer when the ranks of one box have to share its cores (eight ranks
    // with five busy threads each on 16-32 cores stretched the end-to-end step by a quarter)
    int hc = (int)std::thread::hardware_concurrency();
    int nt = hc > 0 ? hc / c->world - 1 : 4;
    if (nt > 4) nt = 4;
    if (const char *e = getenv("LITE_B200_PACK_THREADS")) nt = atoi(e);
    if (nt < 1) nt = 1;
    if (nt > 16) nt = 16;
    c->pool->start(nt);
  }
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int lite_ctx_destroy(lite_ctx *c);

extern "C" int lite_ctx_create(lite_ctx **out, int device, int rank, int world_size,
                               const void *nccl_unique_id) {
  if (!out) return fail(LITE_EINVAL, "null output");
  *out = nullptr;
  if (world_size < 1 || rank < 0 || rank >= world_size) return fail(LITE_EINVAL, "bad rank/world_size");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(LITE_ECUDA, "no CUDA device available (%s): this library has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(LITE_EINVAL, "device %d out of range (%d devices)", device, ndev);
  CU(cudaSetDevice(device));
  lite_ctx *c = new lite_ctx();
  c->device = device; c->rank = rank; c->world = world_size;
  const int r = ctx_init(c, nccl_unique_id);
  if (r) {  // release whatever was created; the message of the failure survives
    const std::string msg = g_err;
    lite_ctx_destroy(c);
    g_err = msg;
    return r;
  }
  *out = c;
  return 0;
}

static void smf_release(lite_ctx *c);

extern "C" int lite_ctx_destroy(lite_ctx *c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  smf_release(c);
  peer_release(c);
  c->peer_info.release();
  if (c->gen_scratch) cudaFree(c->gen_scratch);
  c->g_outer.release(); c->g_hole_off.release(); c->g_hole_cnt.release(); c->g_hole_he.release(); c->g_fgen.release();
  if (c->pool) { c->pool->shutdown(); delete c->pool; c->pool = nullptr; }
  if (c->stream2) cudaStreamSynchronize(c->stream2);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  for (int i = 0; i < 8; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  for (int i = 0; i < LITE_TIMER_SLOTS; i++) if (c->tev[i]) cudaEventDestroy(c->tev[i]);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  if (c->h_part) cudaFreeHost(c->h_part);
  // DBuf members are released explicitly (no destructors: keep the struct POD-like)
  c->vx.release(); c->vy.release(); c->he_length.release();
  c->he_src.release(); c->he_twin.release(); c->he_next.release(); c->he_face.release();
  c->he_seg.release(); c->he_next_hf.release(); c->he_prev_hf.release();
  c->he_dir.release(); c->face_inside.release(); c->seg_bnd.release();
  c->f_off.release(); c->f_cnt.release(); c->f_list.release(); c->seg_he_start.release();
  c->seg_he_end.release(); c->seg_n_edges.release(); c->nb_list.release(); c->b_list.release();
  c->sel_a.release(); c->sel_b.release(); c->id_a.release(); c->id_b.release(); c->sort_tmp.release();
  c->pt.release(); c->udir.release(); c->ang.release(); c->len.release(); c->ca.release(); c->el.release(); c->f_area.release();
  c->f_perim.release(); c->f_sumang.release(); c->f_minang.release(); c->cum.release(); c->pre.release(); c->plen.release(); c->f_tot2.release(); c->stot2.release(); c->sfo.release();
  c->flg.release(); c->slot_he.release(); c->slot_seg.release(); c->slot_face.release();
  c->he_slot.release(); c->guide.release(); c->err.release();
  c->stat.release(); c->r_e1.release(); c->r_e2.release(); c->in_he.release();
  c->r_p1.release(); c->r_p2.release(); c->r_x.release(); c->r_angle.release(); c->r_u.release();
  c->in_x.release(); c->in_angle.release();
  c->mstat.release(); c->fstat.release(); c->sums.release(); c->m_he.release();
  c->fl_e1.release(); c->fl_e2.release(); c->fl_right.release(); c->fl_p2.release();
  c->partials.release(); c->red_out.release(); c->counter.release(); c->status.release(); c->res.release();
  if (c->arena.d) cudaFree(c->arena.d);
  if (c->arena.h) cudaFreeHost(c->arena.h);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_up0) cudaEventDestroy(c->ev_up0);
  if (c->ev_guide) cudaEventDestroy(c->ev_guide);
  if (c->ev_upload) cudaEventDestroy(c->ev_upload);
  if (c->ev_mf) cudaEventDestroy(c->ev_mf);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  if (c->stream) cudaStreamDestroy(c->stream);
  (void)cudaGetLastError();
  delete c;
  return 0;
}

static int tess_err_sync(lite_ctx *c);
static int mf_flush(lite_ctx *c);

extern "C" int lite_ctx_sync(lite_ctx *c) {
  if (!c) return fail(LITE_EINVAL, "null context");
  CU(cudaSetDevice(c->device));
  { const int fr = mf_flush(c); if (fr) return fr; }
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaStreamSynchronize(c->stream2));
  return tess_err_sync(c);
}

extern "C" int lite_ctx_set_option(lite_ctx *c, int option, int64_t value) {
  if (!c) return fail(LITE_EINVAL, "null context");
  if (option == LITE_OPT_RECORD_SPLITS) {
    if (c->n_local != 0) return fail(LITE_ESTATE, "clear the splits before changing the record option");
    c->record = value != 0;
    return 0;
  }
  if (option == LITE_OPT_VALIDATE_HOST) { c->validate_host = value != 0; return 0; }
  if (option == LITE_OPT_SAMPLER) {
    if (value != LITE_SAMPLER_STRATIFIED && value != LITE_SAMPLER_IID) return fail(LITE_EINVAL, "unknown sampler %lld", (long long)value);
    c->iid = value == LITE_SAMPLER_IID;
    return 0;
  }
  return fail(LITE_EINVAL, "unknown option %d", option);
}

// stage one host array into the arena; the single H2D copy happens afterwards
template <typename T>
static int up(lite_ctx *c, DBuf<T> &b, const T *src, size_t n) {
  Arena &A = c->arena;
  const size_t bytes = (n ? n : 1) * sizeof(T);
  if (A.used + align256(bytes) > A.cap) return fail(LITE_ESTATE, "upload arena overflow");
  if (!b.view) b.release();
  if (n) c->pack_jobs.push_back({A.h + A.used, (const void *)src, n * sizeof(T)});
  b.p = (T *)(A.d + A.used); b.n = n ? n : 1; b.view = true;
  A.used += align256(bytes);
  return 0;
}
static int arena_reserve(lite_ctx *c, size_t need) {
  Arena &A = c->arena;
  A.used = 0;
  if (need <= A.cap) return 0;
  CU(cudaStreamSynchronize(c->stream));
  if (A.d) cudaFree(A.d);
  if (A.h) cudaFreeHost(A.h);
  A.d = A.h = nullptr; A.cap = 0;
  c->have_tess = false;  // the previous tessellation lived in the freed block
  const size_t cap = need + need / 4;
  CU(cudaMalloc((void **)&A.d, cap));
  CU(cudaMallocHost((void **)&A.h, cap));
  A.cap = cap;
  return 0;
}

// copy into the pinned upload block with non-temporal stores (dst 16-byte aligned): the lines do
// not stay dirty in the writing core's cache, where the copy engine would have to snoop them out
static void memcpy_stream(char *dst, const char *src, size_t n) {
  size_t i = 0;
  for (; i + 16 <= n; i += 16) _mm_stream_si128((__m128i *)(dst + i), _mm_loadu_si128((const __m128i *)(src + i)));
  for (; i < n; i++) dst[i] = src[i];
}

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// The export assertions on the host (faces, lists, links), over `nt` threads of the pool: the
// path of tessellations with holes (their boundary cycles are derived on the host, which must
// not walk unchecked links) and of LITE_OPT_VALIDATE_HOST.  Every other upload is checked on
// the device (k_validate).  Returns 0 or fails with the first inconsistency.
static int host_validate(lite_ctx *c, const lite_tess_soa *s) {
  const int nH = s->n_halfedges, nF = s->n_faces, nS = s->n_segments, nV = s->n_vertices;
  const int nt = (int)c->pool->th.size();
  std::vector<int> bad_face(nt, -1), bad_he(nt, -1), bad_list_f(nt, -1), bad_list_k(nt, -1);
  c->pool->run([=, &bad_face, &bad_he, &bad_list_f, &bad_list_k](int t) {
    const int f0 = (int)((int64_t)nF * t / nt), f1 = (int)((int64_t)nF * (t + 1) / nt);
    for (int f = f0; f < f1 && bad_face[t] < 0 && bad_list_f[t] < 0; f++) {
      const int off = s->face_he_offset[f], cnt = s->face_he_count[f];
      if (cnt < 0 || off < 0 || off + cnt > s->n_face_he || (s->face_inside[f] && cnt < 3)) { bad_face[t] = f; break; }
      for (int k = 0; k < cnt; k++) {
        const int h = s->face_he_list[off + k];
        if (h < 0 || h >= nH ||
            (s->face_inside[f] && (s->he_face[h] != f || s->he_next[h] != s->face_he_list[off + (k + 1 == cnt ? 0 : k + 1)]))) {
          bad_list_f[t] = f; bad_list_k[t] = k; break;
        }
      }
    }
    const int h0 = (int)((int64_t)nH * t / nt), h1 = (int)((int64_t)nH * (t + 1) / nt);
    for (int h = h0; h < h1; h++) {
      const int f = s->he_face[h], tw = s->he_twin[h];
      if (f < 0 || f >= nF || tw < 0 || tw >= nH || s->he_twin[tw] != h || s->he_src[h] < 0 ||
          s->he_src[h] >= nV || s->he_seg[h] < 0 || s->he_seg[h] >= nS || s->he_next[h] < 0 || s->he_next[h] >= nH) { bad_he[t] = h; break; }
    }
  });
  c->pool->wait();
  for (int t = 0; t < nt; t++) {
    if (bad_face[t] >= 0) return fail(LITE_EINVAL, "face %d: halfedge list out of range or fewer than 3 halfedges", bad_face[t]);
    if (bad_list_f[t] >= 0)
      return fail(LITE_EINVAL, "face %d: halfedge list disagrees with he_next/he_face at position %d", bad_list_f[t], bad_list_k[t]);
    if (bad_he[t] >= 0) return fail(LITE_EINVAL, "halfedge %d: inconsistent links", bad_he[t]);
  }
  for (int i = 0; i < nS; i++)
    if (s->seg_he_start[i] < 0 || s->seg_he_start[i] >= nH || s->seg_he_end[i] < 0 || s->seg_he_end[i] >= nH)
      return fail(LITE_EINVAL, "segment %d: halfedge out of range", i);
  for (int i = 0; i < s->n_non_blocking; i++) {
    int sg = s->non_blocking[i];
    if (sg < 0 || sg >= nS || s->seg_n_edges[sg] != 1 || s->seg_on_boundary[sg])
      return fail(LITE_EINVAL, "non-blocking segment %d must be internal with exactly one edge", sg);
  }
  for (int i = 0; i < s->n_blocking; i++) {
    int sg = s->blocking[i];
    if (sg < 0 || sg >= nS || s->seg_n_edges[sg] < 2 || s->seg_on_boundary[sg])
      return fail(LITE_EINVAL, "blocking segment %d must be internal with at least two edges", sg);
  }
  return 0;
}

extern "C" int lite_tess_upload(lite_ctx *c, const lite_tess_soa *s) {
  NvtxRange nvtx_("lite_tess_upload");
  if (!c || !s) return fail(LITE_EINVAL, "null argument");
  if (c->mf_inflight) { cudaStreamSynchronize(c->stream2); c->mf_inflight = false; }
  // From here on the context holds no tessellation until this upload has succeeded: the
  // views into the arena are re-pointed as the arrays are staged, so after a failure the
  // previous tessellation is gone and every call that needs one returns LITE_ESTATE.
  c->have_tess = false; c->have_mf = false; c->mf_pending = false; c->mf_deferred = false;
  c->n_local = 0; c->n_global = 0;
  const double t0 = now_ms();
  double t1 = t0, t2 = t0, t3 = t0, t4 = t0;
  CU(cudaSetDevice(c->device));
  const int nH = s->n_halfedges, nF = s->n_faces, nS = s->n_segments, nV = s->n_vertices;
  if (nH <= 0 || nF <= 0 || nS <= 0 || nV <= 0 || s->n_face_he <= 0 || s->n_non_blocking < 0 || s->n_blocking < 0)
    return fail(LITE_EINVAL, "empty tessellation or negative counts");
  // ---- copies: everything is packed into one pinned block and sent in one transfer
  int n_guide = 1;  // sized once the number of slots is known
  {
    size_t need = 0;
    need += 2 * align256((size_t)nV * 8);
    need += 8 * align256((size_t)nH * 4) + align256((size_t)nH) + align256((size_t)nH * 8);
    need += align256((size_t)nF) + 2 * align256((size_t)nF * 4) + align256((size_t)s->n_face_he * 4);
    need += 3 * align256((size_t)nS * 4) + align256((size_t)nS);
    need += align256((size_t)(s->n_non_blocking + 1) * 4) + align256((size_t)(s->n_blocking + 1) * 4);
    need += align256((size_t)s->n_face_he * 8) + 64 * 256;
    // room for the extended lists of a tessellation with holes (staged only if it has any)
    need += 2 * align256((size_t)nH * 4) + align256((size_t)nH * 8) + 5 * align256((size_t)nF * 4) + align256((size_t)nF);
    int ra = arena_reserve(c, need);
    if (ra) return ra;
  }
  if (c->upload_inflight) { CU(cudaEventSynchronize(c->ev_upload)); c->upload_inflight = false; }
  if (c->early_unfenced) { CU(cudaStreamSynchronize(c->stream)); CU(cudaStreamSynchronize(c->stream2)); c->early_unfenced = false; }
  c->pack_jobs.clear();
  int r = 0;
  r |= up(c, c->vx, s->vx, nV); r |= up(c, c->vy, s->vy, nV);
  r |= up(c, c->he_src, s->he_src, nH); r |= up(c, c->he_twin, s->he_twin, nH);
  r |= up(c, c->he_next, s->he_next, nH); r |= up(c, c->he_face, s->he_face, nH);
  r |= up(c, c->he_seg, s->he_seg, nH); r |= up(c, c->he_next_hf, s->he_next_hf, nH);
  r |= up(c, c->he_prev_hf, s->he_prev_hf, nH); r |= up(c, c->he_dir, s->he_dir, nH);
  r |= up(c, c->he_length, s->he_length, nH);
  r |= up(c, c->face_inside, s->face_inside, nF); r |= up(c, c->f_off, s->face_he_offset, nF);
  r |= up(c, c->f_cnt, s->face_he_count, nF); r |= up(c, c->f_list, s->face_he_list, s->n_face_he);
  r |= up(c, c->seg_he_start, s->seg_he_start, nS); r |= up(c, c->seg_he_end, s->seg_he_end, nS);
  r |= up(c, c->seg_n_edges, s->seg_n_edges, nS); r |= up(c, c->seg_bnd, s->seg_on_boundary, nS);
  r |= up(c, c->nb_list, s->non_blocking, s->n_non_blocking);
  r |= up(c, c->b_list, s->blocking, s->n_blocking);
  if (r) return r;
  // Host side of an upload, on a few threads: packing into the pinned block, ONE light pass
  // that counts the halfedges bounding inside faces (does the tessellation have faces with
  // holes?) and the gather of the halfedge lengths in slot order.  The export assertions proper
  // run on the device behind the copy (k_validate); only a tessellation with holes, whose
  // boundary cycles have to be walked here, is checked on the host first.
  //
  // Packing uses non-temporal stores.  With ordinary stores the lines written by several
  // threads stay dirty in their cores' private caches and the copy engine has to snoop them out:
  // the 3.8 MB copy of a 10^4-cell tessellation then takes 0.5 ms instead of the link's 0.08 ms
  // (scripts/h2d_probe4.py; the whole end-to-end step went from 1.43 to 0.80 ms with this alone).
  const int nt = (int)c->pool->th.size();
  std::vector<int> bad_he(nt, -1);
  std::vector<int64_t> part_inside(nt, 0), part_twice(nt, 0);
  c->h_he_inside.resize(nH);  // every entry is written by the count below
  std::vector<double> &cum = c->h_cum;  // only a tessellation with holes uses it (extended slots)
  // The worker that finishes last starts the copy of everything packed so far (all but the
  // cumulative lengths, which this thread is still summing): the 0.08 ms on the link overlap the
  // rest of the host work.  Not for a tessellation with holes: its extended lists are built below.
  const size_t bulk_bytes = c->arena.used;
  // the cumulative lengths are written straight into the pinned block, behind the bulk
  const size_t cum_off = c->arena.used;
  if (cum_off + align256((size_t)s->n_face_he * sizeof(double)) > c->arena.cap) return fail(LITE_ESTATE, "upload arena overflow");
  c->arena.used += align256((size_t)s->n_face_he * sizeof(double));
  double *const cum_pin = (double *)(c->arena.h + cum_off);
  bool early_copy = false;
  if (g_trace) {
    if (!c->ev_tr[0]) for (int k = 0; k < 3; k++) cudaEventCreate(&c->ev_tr[k]);
    else {
      float h2d = 0, prep = 0;
      if (cudaEventSynchronize(c->ev_tr[2]) == cudaSuccess && cudaEventElapsedTime(&h2d, c->ev_tr[0], c->ev_tr[1]) == cudaSuccess &&
          cudaEventElapsedTime(&prep, c->ev_tr[1], c->ev_tr[2]) == cudaSuccess)
        fprintf(stderr, "[lite_b200] previous upload on the device: copy %.3f ms, rest of the copy + checks + preparation %.3f ms\n", h2d, prep);
    }
  }
  // Everything the device does with an upload except the sampler's tables needs only the bulk:
  // the memsets, the checks and the slot / face preparation.  They are enqueued by whoever
  // enqueues the bulk copy -- the last packing worker, while this thread is still summing the
  // cumulative lengths, or this thread further down.
  auto alloc_slots = [c, nH, nF](size_t nSl) -> int {
    return c->pt.alloc(nSl) || c->udir.alloc(nSl) || c->ang.alloc(nSl) || c->len.alloc(nSl) || c->ca.alloc(nSl) || c->el.alloc(nSl) ||
           c->flg.alloc(nSl) || c->slot_he.alloc(nSl) || c->slot_seg.alloc(nSl) ||
           c->slot_face.alloc(nSl) || c->he_slot.alloc(nH) || c->f_area.alloc(nF) ||
           c->f_perim.alloc(nF) || c->f_sumang.alloc(nF) || c->f_minang.alloc(nF) || c->pre.alloc(nSl) ||
           c->plen.alloc(nSl) || c->f_tot2.alloc(nF) || c->sfo.alloc(nSl) || c->stot2.alloc(nSl);
  };
  auto enqueue_prep = [c, s, nH, nF, nS, nV](size_t nSl, bool general) -> cudaError_t {
    // one memset: the no-exit counters, the general-face overflow count, the exchange's report and
    // the validation word are the tail of the result block (the other status words are zeroed
    // where they are used); he_slot is initialised by k_validate
    cudaError_t e = cudaMemsetAsync(c->status.p, 0, (LITE_RES_WORDS - LITE_RES_STATUS) * sizeof(double), c->stream);
    if (e != cudaSuccess) return e;
    PrepArgs A{};
    A.vx = c->vx.p; A.vy = c->vy.p; A.he_src = c->he_src.p; A.he_twin = c->he_twin.p;
    A.he_next = c->he_next.p; A.he_face = c->he_face.p; A.he_seg = c->he_seg.p;
    A.he_next_hf = c->he_next_hf.p; A.he_dir = c->he_dir.p; A.face_inside = c->face_inside.p;
    A.f_gen = general ? c->g_fgen.p : nullptr;
    A.he_length = c->he_length.p; A.face_he_offset = c->f_off.p; A.face_he_count = c->f_cnt.p;
    A.face_he_list = c->f_list.p; A.n_faces = nF;
    A.pt = c->pt.p; A.udir = c->udir.p; A.ang = c->ang.p; A.len = c->len.p; A.ca = c->ca.p; A.el = c->el.p; A.flg = c->flg.p;
    A.slot_he = c->slot_he.p; A.slot_seg = c->slot_seg.p; A.slot_face = c->slot_face.p;
    A.he_slot = c->he_slot.p; A.err = c->err.p;
    TessDev &T = c->T;
    T.n_slots = (int)nSl; T.n_faces = nF; T.n_he = nH; T.n_seg = nS; T.err = c->err.p;
    T.pt = c->pt.p; T.udir = c->udir.p; T.ang = c->ang.p; T.len = c->len.p; T.ca = c->ca.p; T.flg = c->flg.p;
    T.slot_he = c->slot_he.p; T.slot_seg = c->slot_seg.p; T.slot_face = c->slot_face.p;
    T.he_slot = c->he_slot.p; T.he_twin = c->he_twin.p; T.he_src = c->he_src.p;
    T.he_next_hf = c->he_next_hf.p; T.he_prev_hf = c->he_prev_hf.p;
    T.f_off = c->f_off.p; T.f_cnt = c->f_cnt.p; T.f_gen = general ? c->g_fgen.p : nullptr;
    T.f_area = c->f_area.p; T.f_perim = c->f_perim.p; T.f_sumang = c->f_sumang.p; T.f_minang = c->f_minang.p;
    T.seg_n_edges = c->seg_n_edges.p; T.seg_he_start = c->seg_he_start.p; T.seg_he_end = c->seg_he_end.p;
    T.seg_bnd = c->seg_bnd.p;
    T.sfo = c->sfo.p; T.pre = c->pre.p; T.plen = c->plen.p; T.f_tot2 = c->f_tot2.p; T.stot2 = c->stot2.p;
    if (general) {
      gen::Tess &GT = c->GT;
      GT.n_he = nH; GT.n_faces = nF; GT.vx = c->vx.p; GT.vy = c->vy.p;
      GT.he_src = c->he_src.p; GT.he_twin = c->he_twin.p; GT.he_next = c->he_next.p; GT.he_face = c->he_face.p;
      GT.he_seg = c->he_seg.p; GT.he_next_hf = c->he_next_hf.p; GT.he_prev_hf = c->he_prev_hf.p;
      GT.he_dir = c->he_dir.p; GT.face_inside = c->face_inside.p; GT.seg_bnd = c->seg_bnd.p;
      GT.seg_n_edges = c->seg_n_edges.p; GT.f_outer = c->g_outer.p; GT.f_hole_off = c->g_hole_off.p;
      GT.f_hole_cnt = c->g_hole_cnt.p; GT.hole_he = c->g_hole_he.p;
      if (!c->gen_scratch) {
        c->gen_blocks = c->sm_count;
        // two blocks of scratch: the merge/flip kernel runs on the second stream, beside the split kernel
        e = cudaMalloc((void **)&c->gen_scratch, 2 * (size_t)c->gen_blocks * LITE_GEN_TB * LITE_GEN_PTS * sizeof(gen::Pt));
        if (e != cudaSuccess) { c->gen_scratch = nullptr; return e; }
      }
    }
    ValidateArgs V{};
    V.nV = nV; V.nH = nH; V.nF = nF; V.nS = nS; V.nFH = (int)nSl; V.nNB = s->n_non_blocking; V.nB = s->n_blocking;
    V.he_src = c->he_src.p; V.he_twin = c->he_twin.p; V.he_next = c->he_next.p; V.he_face = c->he_face.p;
    V.he_seg = c->he_seg.p; V.he_next_hf = c->he_next_hf.p; V.he_prev_hf = c->he_prev_hf.p;
    V.face_inside = c->face_inside.p; V.seg_bnd = c->seg_bnd.p; V.f_gen = general ? c->g_fgen.p : nullptr;
    V.f_off = c->f_off.p; V.f_cnt = c->f_cnt.p; V.f_list = c->f_list.p; V.seg_he_start = c->seg_he_start.p;
    V.seg_he_end = c->seg_he_end.p; V.seg_n_edges = c->seg_n_edges.p; V.nb_list = c->nb_list.p; V.b_list = c->b_list.p;
    V.err = c->err.p; V.he_slot = c->he_slot.p;
    int most = nH;
    if (nF > most) most = nF;
    if (nS > most) most = nS;
    if ((int)nSl > most) most = (int)nSl;
    k_validate<<<(most + 255) / 256, 256, 0, c->stream>>>(V);
    k_prep_slots<<<(unsigned)((nSl + 127) / 128), 128, 0, c->stream>>>(A, (int)nSl);
    // outside faces: count 0 in the device view
    const int tb = 64, gb = (nF + tb - 1) / tb;   // one thread per face, latency bound: small blocks reach every SM at 10^4 faces
    k_prep_faces<<<gb, tb, 0, c->stream>>>(T, c->ca.p, c->el.p, c->f_area.p, c->f_perim.p, c->f_sumang.p, c->f_minang.p,
                                           c->pre.p, c->plen.p, c->f_tot2.p, c->sfo.p, c->stot2.p);
    c->launches += 3;
    return cudaGetLastError();
  };
  bool early_prep = false;
  // The bulk leaves in a few groups of 64 KB pieces: a group is sent as soon as its last piece is
  // packed, so the link works while the rest is still being packed.
  std::vector<lite_ctx::PackJob> pieces;
  for (const auto &j : c->pack_jobs)
    for (size_t o = 0; o < j.bytes; o += (64u << 10))
      pieces.push_back({j.dst + o, (const char *)j.src + o, j.bytes - o < (64u << 10) ? j.bytes - o : (size_t)(64u << 10)});
  enum { MAX_GROUPS = 8 };
  int n_groups = 4;
  if (const char *e = getenv("LITE_B200_UPLOAD_GROUPS")) n_groups = atoi(e);
  if (n_groups > MAX_GROUPS) n_groups = MAX_GROUPS;
  if (n_groups > (int)pieces.size()) n_groups = (int)pieces.size();
  if (n_groups < 1) n_groups = 1;
  std::vector<uint8_t> piece_group(pieces.size(), 0);
  size_t g_b0[MAX_GROUPS + 1];
  std::atomic<int> g_left[MAX_GROUPS];
  {
    size_t total = 0, acc = 0;
    for (const auto &q : pieces) total += q.bytes;
    int g = 0, cnt = 0;
    g_b0[0] = 0;
    for (size_t k = 0; k < pieces.size(); k++) {
      if (g + 1 < n_groups && cnt > 0 && acc >= total * (size_t)(g + 1) / (size_t)n_groups) {
        g_left[g].store(cnt, std::memory_order_relaxed);
        g++; cnt = 0;
        g_b0[g] = (size_t)(pieces[k].dst - c->arena.h);
      }
      piece_group[k] = (uint8_t)g;
      cnt++; acc += pieces[k].bytes;
    }
    g_left[g].store(cnt, std::memory_order_relaxed);
    n_groups = g + 1;
    g_b0[n_groups] = bulk_bytes;
  }
  std::atomic<size_t> next_piece{0};
  std::atomic<int> next_count{0};
  const int n_count_chunks = 4 * nt;
  std::atomic<int> arrived_pack{0};
  std::atomic<bool> copy_failed{false}, copies_started{false};
  double wk_start[16] = {0};
  double wk_t[3] = {0, 0, 0};  // trace: the workers' clock when the first group left, when all had, after the preparation
  if (!c->validate_host && alloc_slots((size_t)s->n_face_he)) return LITE_ECUDA;
  if (g_trace) cudaEventRecord(c->ev_tr[0], c->stream);
  CU(cudaEventRecord(c->ev_up0, c->stream));  // whatever still reads the previous tessellation is in front of this
  c->pool->run([=, &bad_he, &part_inside, &part_twice, &wk_t, &wk_start, &pieces,
                &next_piece, &next_count, &arrived_pack, &piece_group, &g_b0, &g_left, &copy_failed, &copies_started](int t) {
    if (g_trace) wk_start[t & 15] = now_ms();
    // pieces handed out one at a time: the arrays differ in size by a factor of 100
    for (;;) {
      const size_t k = next_piece.fetch_add(1, std::memory_order_relaxed);
      if (k >= pieces.size()) break;
      memcpy_stream(pieces[k].dst, (const char *)pieces[k].src, pieces[k].bytes);
      _mm_sfence();
      const int g = piece_group[k];
      if (g_left[g].fetch_sub(1, std::memory_order_acq_rel) == 1 && !c->validate_host) {
        copies_started.store(true, std::memory_order_relaxed);
        if (cudaSetDevice(c->device) != cudaSuccess ||
            cudaMemcpyAsync(c->arena.d + g_b0[g], c->arena.h + g_b0[g], g_b0[g + 1] - g_b0[g], cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
          copy_failed.store(true, std::memory_order_relaxed);
        if (g_trace && wk_t[0] == 0) wk_t[0] = now_ms();
      }
    }
    arrived_pack.fetch_add(1, std::memory_order_acq_rel);   // every group this worker completed has been sent
    if (g_trace) wk_t[1] = now_ms();
    // Does a halfedge bound an inside face, and does a face lie on both sides of an edge (with
    // the listed halfedges, this tells a tessellation with holes)?  In chunks, so that a worker
    // that was busy sending a group takes fewer.
    int64_t inside = 0, twice = 0;
    for (;;) {
      const int q = next_count.fetch_add(1, std::memory_order_relaxed);
      if (q >= n_count_chunks || bad_he[t] >= 0) break;
      const int h0 = (int)((int64_t)nH * q / n_count_chunks), h1 = (int)((int64_t)nH * (q + 1) / n_count_chunks);
      for (int h = h0; h < h1; h++) {
        const int f = s->he_face[h], tw = s->he_twin[h];
        if ((unsigned)f >= (unsigned)nF || (unsigned)tw >= (unsigned)nH) { bad_he[t] = h; break; }
        const bool in = s->face_inside[f] != 0;
        c->h_he_inside[h] = in;
        if (!in) continue;
        inside++;
        if (s->he_face[tw] == f) twice++;   // the face runs along both sides of this edge
      }
    }
    part_inside[t] = inside;
    part_twice[t] = twice;
  });
  // the job reads this frame; a copy that has started must be fenced before the block is reused
  struct Joiner {
    lite_ctx *c; std::atomic<bool> *started;
    ~Joiner() { c->pool->wait(); if (started->load()) c->early_unfenced = true; }
  } joiner{c, &copies_started};
  // the face lists must tile [0, n_face_he) in face order: slot order == face order
  int64_t listed = 0;
  // maximal runs of slots whose faces are all inside (or all outside): the cumulative lengths
  // below are summed run by run, in one flat loop each
  struct SlotRun { int begin, end; bool inside; };
  std::vector<SlotRun> runs;
  {
    int64_t run = 0;
    for (int f = 0; f < nF; f++) {
      const int off = s->face_he_offset[f], cnt = s->face_he_count[f];
      if (off != run || cnt < 0 || cnt > s->n_face_he - off || (s->face_inside[f] && cnt < 3))
        return fail(LITE_EINVAL, "face %d: its halfedge list must follow the previous face's (offset %d, expected %lld) and hold "
                    "at least 3 halfedges when the face is inside", f, off, (long long)run);
      run += cnt;
      const bool in = s->face_inside[f] != 0;
      if (in) listed += cnt;
      if (cnt > 0) {
        if (!runs.empty() && runs.back().inside == in) runs.back().end = off + cnt;
        else runs.push_back({off, off + cnt, in});
      }
    }
    if (run != s->n_face_he) return fail(LITE_EINVAL, "the face lists hold %lld halfedges, n_face_he is %d", (long long)run, s->n_face_he);
  }
  const double t_stage = now_ms();
  // cumulative halfedge lengths of the sampler (cumlen += e->get_length(), lib/ttessel.cpp:1980),
  // accumulated sequentially in SLOT order (faces in index order, each in ccb order from
  // outer_ccb()): the summation order is part of the sampler's definition.  Entry m is slot m.
  // One flat loop per run: the chain of additions is the only dependency left (the loop over
  // faces mispredicted its exit once per face and took twice as long).  Non-temporal stores, as
  // for the rest of the block.
  double cl = 0.0;
  int bad_list = -1;
  // Between blocks of the sum this thread looks whether every group of the bulk has been sent
  // and, once, queues the checks and the preparation behind it -- for a tessellation without
  // holes, which the count in progress has yet to confirm; if it finds holes, everything is
  // prepared again below with the extended lists.  (On a 10^5-cell tessellation the sum takes
  // longer than the copy.)
  bool prep_tried = false;
  auto try_prep = [&]() {
    if (prep_tried || c->validate_host || arrived_pack.load(std::memory_order_acquire) < nt) return;
    prep_tried = true;
    if (copy_failed.load(std::memory_order_relaxed)) return;
    early_copy = true;
    if (g_trace) cudaEventRecord(c->ev_tr[1], c->stream);
    early_prep = enqueue_prep((size_t)s->n_face_he, false) == cudaSuccess;
    if (g_trace) wk_t[2] = now_ms();
  };
  for (const SlotRun &R : runs) {
    if (!R.inside) {
      for (int k = R.begin; k < R.end; k++) _mm_stream_si64((long long *)(cum_pin + k), (long long)__builtin_bit_cast(int64_t, cl));
      continue;
    }
    const int32_t *lst = s->face_he_list;
    const double *hl = s->he_length;
    for (int k0 = R.begin; k0 < R.end && bad_list < 0; k0 += 8192) {
      const int k1 = (R.end - k0 > 8192) ? k0 + 8192 : R.end;
      for (int k = k0; k < k1; k++) {
        const int h = lst[k];
        if ((unsigned)h >= (unsigned)nH) { bad_list = k; break; }
        cl += hl[h];
        _mm_stream_si64((long long *)(cum_pin + k), (long long)__builtin_bit_cast(int64_t, cl));
      }
      try_prep();
    }
    if (bad_list >= 0) break;
  }
  _mm_sfence();
  const double t_cum = now_ms();
  // The cumulative lengths and the guide table go over the second stream, beside the
  // preparation (which does not read them).
  bool side_done = false;
  int n_guide_simple = 1;
  while (n_guide_simple < 2 * s->n_face_he) n_guide_simple <<= 1;
  if (!c->validate_host && bad_list < 0 && cl > 0.0) {
    while (arrived_pack.load(std::memory_order_acquire) < nt) _mm_pause();
    try_prep();
    if (early_prep && c->guide.alloc((size_t)n_guide_simple) == 0) {
      const size_t cum_bytes = align256((size_t)s->n_face_he * sizeof(double));
      bool ok = cudaStreamWaitEvent(c->stream2, c->ev_up0, 0) == cudaSuccess &&
                cudaMemcpyAsync(c->arena.d + cum_off, c->arena.h + cum_off, cum_bytes, cudaMemcpyHostToDevice, c->stream2) == cudaSuccess;
      if (ok) {
        k_build_guide<<<(n_guide_simple + 255) / 256, 256, 0, c->stream2>>>((const double *)(c->arena.d + cum_off), s->n_face_he,
                                                                           cl / (double)n_guide_simple, n_guide_simple, c->guide.p);
        c->launches += 1;
        ok = cudaEventRecord(c->ev_guide, c->stream2) == cudaSuccess;
      }
      side_done = ok;
      if (g_trace) wk_t[2] = now_ms();
    }
  }
  c->pool->wait();
  c->early_unfenced = copies_started.load();
  if (g_trace) fprintf(stderr, "[lite_b200] tess_upload host: staging %.3f ms, cumulative lengths %.3f ms, wait for the workers (pack, count) %.3f ms; "
                       "workers: first group sent at %.3f ms, last one packed at %.3f; preparation and tables queued at %.3f\n",
                       t_stage - t0, t_cum - t_stage, now_ms() - t_cum, wk_t[0] - t0, wk_t[1] - t0, wk_t[2] - t0);
  if (g_trace) {
    fprintf(stderr, "[lite_b200] workers started at");
    for (int t = 0; t < nt && t < 16; t++) fprintf(stderr, " %.3f", wk_start[t] - t0);
    fprintf(stderr, " ms\n");
  }
  if (bad_list >= 0) return fail(LITE_EINVAL, "face list entry %d: not a halfedge", bad_list);
  for (int t = 0; t < nt; t++)
    if (bad_he[t] >= 0) return fail(LITE_EINVAL, "halfedge %d: inconsistent links", bad_he[t]);
  int64_t inside_he = 0, twice = 0;
  for (int t = 0; t < nt; t++) { inside_he += part_inside[t]; twice += part_twice[t]; }
  if (inside_he < listed)
    return fail(LITE_EINVAL, "%lld halfedges bound inside faces but %lld are listed on outer boundaries",
                (long long)inside_he, (long long)listed);
  // Faces that are not simple polygons: more halfedges bound inside faces than the outer
  // boundaries list (inner boundaries: holes), or a face lies on both sides of an edge.  Their
  // slots are the outer boundary followed by the inner ones, their candidates go through
  // gen_face.h.  A rectangular window never has any and skips all of this.
  const bool general = inside_he != listed || twice != 0;
  if (side_done) CU(cudaStreamWaitEvent(c->stream, c->ev_guide, 0));
  if (general) { early_prep = false; side_done = false; }  // what was queued assumed simple faces: prepared again below
  c->n_general = 0;
  if (general || c->validate_host) {
    const int hv = host_validate(c, s);
    if (hv) return hv;
  }
  const int32_t *l_off = s->face_he_offset, *l_cnt = s->face_he_count, *l_list = s->face_he_list;
  size_t n_slots = (size_t)s->n_face_he;
  if (general) {
    gen::derive_cycles(nH, nF, s->he_next, s->he_face, s->he_twin, s->face_inside, s->face_he_offset,
                       s->face_he_count, s->face_he_list, c->cyc);
    c->n_general = c->cyc.n_general;
    c->x_off.assign(nF, 0); c->x_cnt.assign(nF, 0); c->x_gen.assign(nF, 0);
    c->x_list.clear();
    c->x_list.reserve((size_t)inside_he);
    for (int f = 0; f < nF; f++) {
      c->x_off[f] = (int)c->x_list.size();
      if (!s->face_inside[f]) continue;
      bool tw = false;
      for (int k = 0; k < s->face_he_count[f]; k++) {
        const int h = s->face_he_list[s->face_he_offset[f] + k];
        c->x_list.push_back(h);
        tw |= s->he_face[s->he_twin[h]] == f;
      }
      for (int q = 0; q < c->cyc.f_hole_cnt[f]; q++) {
        const int h0 = c->cyc.hole_he[c->cyc.f_hole_off[f] + q];
        int h = h0, guard = 0;
        do { c->x_list.push_back(h); h = s->he_next[h]; } while (h != h0 && ++guard < nH);
      }
      c->x_cnt[f] = (int)c->x_list.size() - c->x_off[f];
      c->x_gen[f] = (c->cyc.f_hole_cnt[f] > 0 || tw) ? 1 : 0;
    }
    if ((int64_t)c->x_list.size() != inside_he)
      return fail(LITE_EINVAL, "boundary cycles of the inside faces hold %zu halfedges, %lld bound inside faces",
                  c->x_list.size(), (long long)inside_he);
    l_off = c->x_off.data(); l_cnt = c->x_cnt.data(); l_list = c->x_list.data();
    n_slots = c->x_list.size();
    // the sampler's cumulative lengths over the extended slots
    cum.assign(n_slots, 0.0);
    cl = 0.0;
    for (int f = 0; f < nF; f++)
      for (int k = 0; k < l_cnt[f]; k++) { cl += s->he_length[l_list[l_off[f] + k]]; cum[(size_t)l_off[f] + k] = cl; }
  }
  while (n_guide < 2 * (int)n_slots) n_guide <<= 1;
  if (!(fabs(cl - s->int_length) <= 1e-9 * fmax(1.0, fabs(cl))))
    return fail(LITE_EINVAL, "int_length %.17g differs from the sum of inside halfedge lengths %.17g",
                s->int_length, cl);
  t1 = now_ms();
  c->pack_jobs.clear();
  size_t n_cum = (size_t)s->n_face_he;
  if (general) {
    r |= up(c, c->cum, cum.data(), cum.size());
    n_cum = cum.size();
  } else {
    if (!c->cum.view) c->cum.release();
    c->cum.p = (double *)(c->arena.d + cum_off); c->cum.n = n_cum; c->cum.view = true;
  }
  if (general) {  // the extended lists replace the caller's, the cycles go along
    r |= up(c, c->f_off, l_off, (size_t)nF); r |= up(c, c->f_cnt, l_cnt, (size_t)nF);
    r |= up(c, c->f_list, l_list, n_slots);
    r |= up(c, c->g_outer, c->cyc.f_outer.data(), (size_t)nF);
    r |= up(c, c->g_hole_off, c->cyc.f_hole_off.data(), (size_t)nF);
    r |= up(c, c->g_hole_cnt, c->cyc.f_hole_cnt.data(), (size_t)nF);
    r |= up(c, c->g_hole_he, c->cyc.hole_he.data(), c->cyc.hole_he.size());
    r |= up(c, c->g_fgen, c->x_gen.data(), (size_t)nF);
  }
  if (c->guide.alloc((size_t)n_guide)) return LITE_ECUDA;
  if (r) return r;
  for (size_t k = 0; k < c->pack_jobs.size(); k++) memcpy_stream(c->pack_jobs[k].dst, (const char *)c->pack_jobs[k].src, c->pack_jobs[k].bytes);
  _mm_sfence();
  c->pack_jobs.clear();
  t2 = now_ms();
  // what was staged since the bulk left (the cumulative lengths, unless they already went over
  // the second stream; the extended lists of a tessellation with holes) -- or everything
  if (side_done) {
    // nothing left: the bulk went in groups, the cumulative lengths over the second stream
  } else if (early_copy) {
    CU(cudaMemcpyAsync(c->arena.d + bulk_bytes, c->arena.h + bulk_bytes, c->arena.used - bulk_bytes, cudaMemcpyHostToDevice, c->stream));
  } else {
    if (g_trace) cudaEventRecord(c->ev_tr[0], c->stream);
    CU(cudaMemcpyAsync(c->arena.d, c->arena.h, c->arena.used, cudaMemcpyHostToDevice, c->stream));
    if (g_trace) cudaEventRecord(c->ev_tr[1], c->stream);
  }
  const size_t nSl = n_slots;
  if (!early_prep) {
    if (alloc_slots(nSl)) return LITE_ECUDA;
    if (enqueue_prep(nSl, general) != cudaSuccess) return fail(LITE_ECUDA, "preparation kernels: %s", cudaGetErrorString(cudaGetLastError()));
  }
  TessDev &T = c->T;
  T.cum = c->cum.p; T.guide = c->guide.p; T.n_cum = (int)n_cum; T.n_guide = n_guide;
  T.total_len = cl; T.guide_scale = (double)n_guide / cl;
  if (!side_done) {
    k_build_guide<<<(n_guide + 255) / 256, 256, 0, c->stream>>>(c->cum.p, (int)n_cum, cl / (double)n_guide, n_guide, c->guide.p);
    c->launches += 1;
  }
  CU(cudaEventRecord(c->ev_upload, c->stream));                     // both copies have left the pinned block
  c->upload_inflight = true; c->early_unfenced = false;
  // the merge/flip stream only depends on the tessellation being in place: it forks here,
  // not from whatever the main stream is doing when SetEnergy is called (the split kernel,
  // if AddSplits came first: the two then run side by side)
  CU(cudaEventRecord(c->ev_fork, c->stream));
  c->fork_pending = true;
  if (g_trace) cudaEventRecord(c->ev_tr[2], c->stream);
  t3 = now_ms();
  // no synchronisation: the copy, the checks and the preparation kernels run behind the
  // caller's next calls on the same stream
  CU(cudaGetLastError());
  t4 = now_ms();
  if (g_trace)
    fprintf(stderr, "[lite_b200] tess_upload: checks %.3f ms, pack %.3f ms, enqueue %.3f ms, device wait %.3f ms (%zu bytes)\n",
            t1 - t0, t2 - t1, t3 - t2, t4 - t3, c->arena.used);
  c->n_nb = s->n_non_blocking; c->n_b = s->n_blocking; c->int_length = s->int_length;
  c->have_tess = true; c->have_mf = false; c->mf_pending = false;
  c->n_local = 0; c->n_global = 0;
  return 0;
}

// after a synchronisation of c->stream: did the uploaded tessellation pass k_validate?
static int tess_err_report(lite_ctx *c, int code) {
  if (!code) return 0;
  static const char *what[] = {"?", "halfedge with links out of range or twin(twin) != itself", "face with a halfedge list out of range",
                               "face list entry that is not a halfedge of the face followed by its next", "segment with halfedges out of range",
                               "non-blocking segment that is not internal with exactly one edge",
                               "blocking segment that is not internal with at least two edges"};
  const int k = (code >> 24) & 0xff;
  c->have_tess = false; c->have_mf = false;
  return fail(LITE_EINVAL, "the uploaded tessellation is inconsistent: %s (element %d); upload a valid one",
              what[(k >= 1 && k <= 6) ? k : 0], code & 0xffffff);
}
static int tess_err_sync(lite_ctx *c) {
  if (!c->have_tess) return 0;
  int code = 0;
  CU(cudaMemcpyAsync(&code, c->err.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return tess_err_report(c, code);
}

static int category_of(int fid) {
  switch (fid) {
    case LITE_FEAT_IS_POINT_INSIDE_WINDOW: return 0;
    case LITE_FEAT_EDGE_LENGTH: case LITE_FEAT_EDGE_LENGTH_INTERNAL: return 1;
    case LITE_FEAT_FACE_NUMBER: case LITE_FEAT_FACE_AREA_2: case LITE_FEAT_FACE_PERIMETER:
    case LITE_FEAT_FACE_SHAPE: case LITE_FEAT_FACE_SUM_OF_ANGLES: case LITE_FEAT_MIN_ANGLE: return 2;
    case LITE_FEAT_SEG_NUMBER: case LITE_FEAT_IS_SEGMENT_INTERNAL:
    case LITE_FEAT_MINUS_IS_SEGMENT_INTERNAL: case LITE_FEAT_SEGMENT_SIZE_2: return 3;
    default: return -1;
  }
}

extern "C" int lite_energy_set(lite_ctx *c, int d, const int32_t *fid) {
  if (!c || !fid) return fail(LITE_EINVAL, "null argument");
  if (d < 1 || d > LITE_MAX_D) return fail(LITE_EINVAL, "d=%d outside [1,%d]", d, LITE_MAX_D);
  int lastcat = 0;
  Prog G{};
  G.d = d;
  for (int i = 0; i < d; i++) {
    int cat = category_of(fid[i]);
    if (cat < 0) return fail(LITE_EINVAL, "unknown feature id %d: only the built-in feature functions run on the device", fid[i]);
    if (cat < lastcat) return fail(LITE_EINVAL, "features must come in CatVector order vertices, edges, faces, segs");
    lastcat = cat;
    G.fid[i] = fid[i];
    G.mask |= 1u << fid[i];
  }
  if (c->n_local != 0) return fail(LITE_ESTATE, "clear the splits before changing the energy");
  if (c->mf_inflight) { cudaStreamSynchronize(c->stream2); c->mf_inflight = false; }
  if (c->G.d != G.d) {  // the statistic block is d columns of `cap` rows: re-dimension it
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    c->stat.release();
    c->cap = 0;
  }
  c->G = G;
  c->have_prog = true; c->have_mf = false; c->mf_pending = false; c->mf_deferred = false;
  return 0;
}

// after a stream synchronisation: pick up the merge/flip sums copied by merges_flips
static void finalize_mf(lite_ctx *c) {
  if (!c->mf_pending) return;
  const int d = c->G.d;
  // blocks [0, gm) hold merges, the others flips; summed in block order (deterministic)
  c->h_sum_m.assign(d, 0.0); c->h_sum_f.assign(d, 0.0);
  for (int b = 0; b < c->mf_blocks; b++) {
    double *dst = (b < c->mf_gm) ? c->h_sum_m.data() : c->h_sum_f.data();
    for (int k = 0; k < d; k++) dst[k] += c->h_part[(size_t)b * LITE_MAX_D + k];
  }
  cudaEventElapsedTime(&c->ms_mf, c->ev[4], c->ev[5]);
  if (g_trace && c->ev01) {
    float a = 0, b = 0;
    if (cudaEventElapsedTime(&a, c->ev[0], c->ev[4]) == cudaSuccess && cudaEventElapsedTime(&b, c->ev[1], c->ev[5]) == cudaSuccess)
      fprintf(stderr, "[lite_b200] merge/flip kernels start %.3f ms after the split kernel starts, end %.3f ms after it ends\n", a, b);
  }
  c->mf_pending = false;
}

// The merge/flip kernel is enqueued at once, on the second stream, BEFORE the split kernel of the
// same step: k_split fills every SM with blocks that stay until the batch is done (6 x 128 threads
// at 80 registers take the whole register file), so anything launched behind it only starts when
// it drains.  Deferring this launch until after the split kernel (to get k_split onto the GPU ~15 us
// earlier) was measured: the merge/flip kernel then ends 0.27 ms late and the step loses 20 us.
static int mf_flush(lite_ctx *c);

extern "C" int lite_candidates_merges_flips(lite_ctx *c, int32_t *n_merges, int32_t *n_flips) {
  NvtxRange nvtx_("lite_candidates_merges_flips");
  if (!c) return fail(LITE_EINVAL, "null context");
  if (!c->have_tess || !c->have_prog) return fail(LITE_ESTATE, "upload a tessellation and set the energy first");
  c->mf_deferred = true;
  c->have_mf = true;
  if (n_merges) *n_merges = c->n_nb;
  if (n_flips) *n_flips = 2 * c->n_b;
  return mf_flush(c);
}

static int mf_flush(lite_ctx *c) {
  if (!c->mf_deferred) return 0;
  c->mf_deferred = false;
  int32_t *n_merges = nullptr, *n_flips = nullptr;
  CU(cudaSetDevice(c->device));
  const int d = c->G.d, nm = c->n_nb, nf = 2 * c->n_b;
  if (c->mstat.alloc((size_t)d * (nm ? nm : 1)) || c->fstat.alloc((size_t)d * (nf ? nf : 1)) ||
      c->m_he.alloc(nm ? nm : 1) || c->fl_e1.alloc(nf ? nf : 1) ||
      c->fl_e2.alloc(nf ? nf : 1) || c->fl_right.alloc(nf ? nf : 1) || c->fl_p2.alloc(nf ? nf : 1))
    return LITE_ECUDA;
  const int blocks = (nm + 127) / 128 + (nf + 127) / 128;
  c->mf_gm = (nm + 127) / 128; c->mf_blocks = blocks;
  if (c->sums.alloc((size_t)(blocks ? blocks : 1) * LITE_MAX_D)) return LITE_ECUDA;
  if ((size_t)blocks * LITE_MAX_D > c->h_part_cap) {
    if (c->mf_inflight) CU(cudaStreamSynchronize(c->stream2));
    if (c->h_part) cudaFreeHost(c->h_part);
    c->h_part = nullptr; c->h_part_cap = 0;
    CU(cudaMallocHost((void **)&c->h_part, (size_t)blocks * LITE_MAX_D * sizeof(double)));
    c->h_part_cap = (size_t)blocks * LITE_MAX_D;
  }
  // fork: the merge/flip kernels are independent of the split kernel and run beside it
  if (c->mf_inflight) CU(cudaStreamSynchronize(c->stream2));
  if (c->fork_pending) { CU(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0)); c->fork_pending = false; }  // recorded by lite_tess_upload
  CU(cudaEventRecord(c->ev[4], c->stream2));
  CU(cudaMemsetAsync(c->status.p + 1, 0, sizeof(unsigned long long), c->stream2));
  if (nm + nf > 0) {
    MFArgs A{};
    A.nb_seg = c->nb_list.p; A.nm = nm; A.mstat = c->mstat.p; A.m_he = c->m_he.p; A.gm = (nm + 127) / 128;
    A.b_seg = c->b_list.p; A.nb = c->n_b; A.fstat = c->fstat.p; A.f_e1 = c->fl_e1.p; A.f_e2 = c->fl_e2.p;
    A.f_right = c->fl_right.p; A.f_p2 = c->fl_p2.p; A.status = c->status.p + 1; A.part = c->sums.p;
    if (c->n_general > 0) {  // rows of the merges / flips next to a face with holes, first
      k_gen_mf<<<c->gen_blocks, LITE_GEN_TB, 0, c->stream2>>>(c->T, c->GT, c->G, A,
                                                               c->gen_scratch + (size_t)c->gen_blocks * LITE_GEN_TB * LITE_GEN_PTS, c->status.p + 6);
      c->launches += 1;
    }
    k_merges_flips<<<blocks, 128, 0, c->stream2>>>(c->T, c->G, A);
    c->launches += 1;
  }
  CU(cudaEventRecord(c->ev[5], c->stream2));
  // the partial sums come back with the next synchronisation (lite_pl_eval / lite_pl_sums):
  // SetEnergy itself does not stall the stream
  if (blocks > 0)
    CU(cudaMemcpyAsync(c->h_part, c->sums.p, (size_t)blocks * LITE_MAX_D * sizeof(double), cudaMemcpyDeviceToHost, c->stream2));
  CU(cudaEventRecord(c->ev_mf, c->stream2));
  c->mf_pending = true; c->mf_inflight = true;
  c->have_mf = true;
  if (n_merges) *n_merges = nm;
  if (n_flips) *n_flips = nf;
  return 0;
}

static int grow_splits(lite_ctx *c, int64_t add) {
  const int d = c->G.d;
  size_t need = (size_t)(c->n_local + add);
  if (need > c->cap) {
    size_t ncap = c->cap ? c->cap * 2 : 0;
    if (ncap < need) ncap = need;
    ncap = (ncap + 255) & ~(size_t)255;
    DBuf<double> nb;
    if (nb.alloc(ncap * d)) return LITE_ECUDA;
    if (c->n_local)
      CU(cudaMemcpy2DAsync(nb.p, ncap * sizeof(double), c->stat.p, c->cap * sizeof(double),
                           (size_t)c->n_local * sizeof(double), d, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->stat.release();
    c->stat = nb;
    c->cap = ncap;
  }
  if (c->record && need > c->rcap) {
    size_t ncap = c->rcap ? c->rcap * 2 : 0;
    if (ncap < need) ncap = need;
    DBuf<int> e1, e2; DBuf<double2> p1, p2; DBuf<double> x, a, u;
    if (e1.alloc(ncap) || e2.alloc(ncap) || p1.alloc(ncap) || p2.alloc(ncap) || x.alloc(ncap) ||
        a.alloc(ncap) || u.alloc(ncap))
      return LITE_ECUDA;
    size_t n = (size_t)c->n_local;
    if (n) {
      CU(cudaMemcpyAsync(e1.p, c->r_e1.p, n * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
      CU(cudaMemcpyAsync(e2.p, c->r_e2.p, n * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
      CU(cudaMemcpyAsync(p1.p, c->r_p1.p, n * sizeof(double2), cudaMemcpyDeviceToDevice, c->stream));
      CU(cudaMemcpyAsync(p2.p, c->r_p2.p, n * sizeof(double2), cudaMemcpyDeviceToDevice, c->stream));
      CU(cudaMemcpyAsync(x.p, c->r_x.p, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
      CU(cudaMemcpyAsync(a.p, c->r_angle.p, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
      CU(cudaMemcpyAsync(u.p, c->r_u.p, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    c->r_e1.release(); c->r_e2.release(); c->r_p1.release(); c->r_p2.release();
    c->r_x.release(); c->r_angle.release(); c->r_u.release();
    c->r_e1 = e1; c->r_e2 = e2; c->r_p1 = p1; c->r_p2 = p2; c->r_x = x; c->r_angle = a; c->r_u = u;
    c->rcap = ncap;
  }
  return 0;
}

// Room for n more dummy splits on this rank without reallocating (PLInferenceNOIS grows the
// set by n_nb per iteration, lib/ttessel.cpp:3500: a caller that knows how many iterations it will
// run reserves once instead of paying a doubling of a multi-hundred-MB block inside the loop).
extern "C" int lite_candidates_reserve(lite_ctx *c, int64_t n_more_local) {
  if (!c) return fail(LITE_EINVAL, "null context");
  if (!c->have_prog) return fail(LITE_ESTATE, "set the energy first");
  if (n_more_local < 0) return fail(LITE_EINVAL, "negative count");
  CU(cudaSetDevice(c->device));
  const size_t need = (size_t)(c->n_local + n_more_local);
  if (need <= c->cap) return 0;
  const int d = c->G.d;
  const size_t ncap = (need + 255) & ~(size_t)255;
  DBuf<double> nb;
  if (nb.alloc(ncap * d)) return LITE_ECUDA;
  if (c->n_local)
    CU(cudaMemcpy2DAsync(nb.p, ncap * sizeof(double), c->stat.p, c->cap * sizeof(double),
                         (size_t)c->n_local * sizeof(double), d, cudaMemcpyDeviceToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->stat.release();
  c->stat = nb;
  c->cap = ncap;
  return 0;
}

static int split_grid(lite_ctx *c, int64_t n) {
  // one resident wave: every warp then owns one contiguous run of candidates
  int64_t blocks = (n + LITE_SPLIT_TB - 1) / LITE_SPLIT_TB;
  int64_t maxb = (int64_t)c->sm_count * LITE_SPLIT_MINB;
  if (blocks > maxb) blocks = maxb;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

static bool prog_is_acs(const Prog &G) {
  return G.d == 4 && G.fid[0] == LITE_FEAT_EDGE_LENGTH && G.fid[1] == LITE_FEAT_FACE_AREA_2 &&
         G.fid[2] == LITE_FEAT_FACE_SUM_OF_ANGLES && G.fid[3] == LITE_FEAT_IS_SEGMENT_INTERNAL;
}
// the sim_acs energy gets the compile-time program, everything else the run-time one
template <int MODE>
static void launch_split(lite_ctx *c, const Prog &G, const SplitArgs &A, int grid) {
  if ((MODE & 4) && prog_is_acs(G)) k_split<MODE, ProgACS><<<grid, LITE_SPLIT_TB, 0, c->stream>>>(c->T, G, A);
  else k_split<MODE, ProgDyn><<<grid, LITE_SPLIT_TB, 0, c->stream>>>(c->T, G, A);
  c->launches += 1;
  if (c->n_general > 0) {  // the candidates on faces with holes, which k_split left out
    GenSplitArgs X{};
    X.A = A; X.philox = MODE & 1; X.record = (MODE & 2) ? 1 : 0; X.stat = (MODE & 4) ? 1 : 0; X.check = (MODE & 8) ? 1 : 0;
    X.scratch = c->gen_scratch; X.overflow = c->status.p + 6;
    k_gen_splits<<<c->gen_blocks, LITE_GEN_TB, 0, c->stream>>>(c->T, c->GT, G, X);
    c->launches += 1;
  }
}

static void fill_split_args(lite_ctx *c, SplitArgs &A) {
  A.stat = c->stat.p; A.cap = c->cap; A.base = (long)c->n_local;
  A.r_e1 = c->r_e1.p; A.r_e2 = c->r_e2.p; A.r_p1 = c->r_p1.p; A.r_p2 = c->r_p2.p;
  A.r_x = c->r_x.p; A.r_angle = c->r_angle.p; A.r_u = c->r_u.p;
  A.status = c->status.p; A.checksum = c->status.p + 2; A.ticket = c->status.p + 4;
}

extern "C" int lite_candidates_add_splits_given(lite_ctx *c, int64_t n, const int32_t *he,
                                                const double *x, const double *angle,
                                                int64_t n_global_total) {
  NvtxRange nvtx_("lite_candidates_add_splits_given");
  if (!c) return fail(LITE_EINVAL, "null context");
  if (!c->have_tess || !c->have_prog) return fail(LITE_ESTATE, "upload a tessellation and set the energy first");
  if (n < 0 || n_global_total < n) return fail(LITE_EINVAL, "bad candidate counts");
  CU(cudaSetDevice(c->device));
  if (n == 0) { c->n_global += n_global_total; return 0; }
  if (!he || !x || !angle) return fail(LITE_EINVAL, "null candidate arrays");
  // validate the halfedge ids (they index device memory): a split starts on a
  // halfedge bounding an inside face (lib/ttessel.cpp:1978)
  for (int64_t i = 0; i < n; i++)
    if (he[i] < 0 || he[i] >= c->T.n_he || !c->h_he_inside[he[i]])
      return fail(LITE_EINVAL, "candidate %lld: halfedge id %d does not bound an inside face", (long long)i, he[i]);
  int r = grow_splits(c, n);
  if (r) return r;
  if (c->in_he.alloc(n) || c->in_x.alloc(n) || c->in_angle.alloc(n)) return LITE_ECUDA;
  CU(cudaEventRecord(c->ev[0], c->stream));
  CU(cudaMemcpyAsync(c->in_he.p, he, n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->in_x.p, x, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->in_angle.p, angle, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  SplitArgs A{};
  fill_split_args(c, A);
  A.n_local = n; A.j0 = 0; A.sel_step = 0;
  A.in_he = c->in_he.p; A.in_x = c->in_x.p; A.in_angle = c->in_angle.p;
  int g = split_grid(c, n);
  if (c->record) launch_split<4 | 2>(c, c->G, A, g);
  else launch_split<4>(c, c->G, A, g);
  CU(cudaEventRecord(c->ev[1], c->stream));
  c->ev01 = true;
  { const int fr = mf_flush(c); if (fr) return fr; }
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  cudaEventElapsedTime(&c->ms_split, c->ev[0], c->ev[1]);
  c->n_local += n; c->n_global += n_global_total;
  return 0;
}

static void shard_range(int64_t n, int rank, int world, int64_t &j0, int64_t &cnt) {
  // contiguous shards [r*n/G, (r+1)*n/G)
  j0 = (int64_t)(((__int128)n * rank) / world);
  int64_t j1 = (int64_t)(((__int128)n * (rank + 1)) / world);
  cnt = j1 - j0;
}

// LITE_SAMPLER_IID: draw this shard's positions independently, sort them with their batch
// indices (split_sample sorts its draws so that the candidates come out in halfedge order,
// lib/ttessel.cpp:1974; here the order also keeps the lanes of a warp on one face)
static int iid_prepare(lite_ctx *c, SplitArgs &A, int64_t n, int64_t j0, uint64_t seed, uint64_t offset) {
  if (c->sel_a.alloc((size_t)n) || c->sel_b.alloc((size_t)n) || c->id_a.alloc((size_t)n) || c->id_b.alloc((size_t)n))
    return LITE_ECUDA;
  k_iid_draw<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((long)n, (long)j0, seed, offset, c->T.total_len,
                                                                 c->sel_a.p, c->id_a.p);
  size_t tmp = 0;
  CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp, c->sel_a.p, c->sel_b.p, c->id_a.p, c->id_b.p, (int)n, 0, 64, c->stream));
  if (c->sort_tmp.alloc(tmp + 256)) return LITE_ECUDA;
  CU(cub::DeviceRadixSort::SortPairs(c->sort_tmp.p, tmp, c->sel_a.p, c->sel_b.p, c->id_a.p, c->id_b.p, (int)n, 0, 64, c->stream));
  c->launches += 2;
  A.sel_sorted = c->sel_b.p; A.sel_id = c->id_b.p;
  return 0;
}

extern "C" int lite_candidates_add_splits_philox(lite_ctx *c, int64_t n_global, uint64_t seed,
                                                 uint64_t offset) {
  NvtxRange nvtx_("lite_candidates_add_splits_philox");
  if (!c) return fail(LITE_EINVAL, "null context");
  if (!c->have_tess || !c->have_prog) return fail(LITE_ESTATE, "upload a tessellation and set the energy first");
  if (n_global < 0) return fail(LITE_EINVAL, "negative count");
  CU(cudaSetDevice(c->device));
  int64_t j0, n;
  shard_range(n_global, c->rank, c->world, j0, n);
  if (n > 0) {
    int r = grow_splits(c, n);
    if (r) return r;
    SplitArgs A{};
    fill_split_args(c, A);
    A.n_local = n; A.j0 = j0; A.sel_step = c->T.total_len / (double)n_global; A.seed = seed; A.ctr0 = offset;
    int g = split_grid(c, n);
    CU(cudaEventRecord(c->ev[0], c->stream));
    if (c->iid) {
      if (n > 0x7fffffff) return fail(LITE_EINVAL, "the i.i.d. sampler sorts at most 2^31-1 positions per rank and batch");
      r = iid_prepare(c, A, n, j0, seed, offset);
      if (r) return r;
    }
    if (c->record) launch_split<4 | 2 | 1>(c, c->G, A, g);
    else launch_split<4 | 1>(c, c->G, A, g);
    CU(cudaEventRecord(c->ev[1], c->stream));
  c->ev01 = true;
  }
  c->n_local += n; c->n_global += n_global;
  c->rev = false;  // k_split wrote the tail last: the next reduce starts from it
  return mf_flush(c);  // the merge/flip kernel, if SetEnergy asked for it, goes out behind the split kernel
}

extern "C" int lite_lines_sample_clip(lite_ctx *c, int64_t n_global, uint64_t seed, uint64_t offset,
                                      int mode, uint64_t *checksum2) {
  NvtxRange nvtx_("lite_lines_sample_clip");
  if (!c) return fail(LITE_EINVAL, "null context");
  if (!c->have_tess) return fail(LITE_ESTATE, "upload a tessellation first");
  if (n_global < 0 || (mode != 0 && mode != 1)) return fail(LITE_EINVAL, "bad arguments");
  CU(cudaSetDevice(c->device));
  int64_t j0, n;
  shard_range(n_global, c->rank, c->world, j0, n);
  CU(cudaMemsetAsync(c->status.p + 2, 0, 2 * sizeof(unsigned long long), c->stream));
  if (n > 0) {
    Prog G0{};  // no features
    bool rec_save = c->record;
    if (mode == 1) {
      if (c->n_local != 0) return fail(LITE_ESTATE, "store mode needs an empty split list");
      c->record = true;
      int r = grow_splits(c, n);
      c->record = rec_save;
      if (r) return r;
    }
    SplitArgs A{};
    fill_split_args(c, A);
    A.base = 0;
    A.n_local = n; A.j0 = j0; A.sel_step = c->T.total_len / (double)n_global; A.seed = seed; A.ctr0 = offset;
    int g = split_grid(c, n);
    CU(cudaEventRecord(c->ev[0], c->stream));
    if (c->iid) {
      if (n > 0x7fffffff) return fail(LITE_EINVAL, "the i.i.d. sampler sorts at most 2^31-1 positions per rank and batch");
      int r = iid_prepare(c, A, n, j0, seed, offset);
      if (r) return r;
    }
    if (mode == 1) launch_split<8 | 2 | 1>(c, G0, A, g);
    else launch_split<8 | 1>(c, G0, A, g);
    CU(cudaEventRecord(c->ev[1], c->stream));
  c->ev01 = true;
  }
  CU(cudaMemcpyAsync(c->h_pinned, c->status.p + 2, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  { const int tr = tess_err_sync(c); if (tr) return tr; }
  if (n > 0) cudaEventElapsedTime(&c->ms_split, c->ev[0], c->ev[1]);
  if (checksum2) memcpy(checksum2, c->h_pinned, 2 * sizeof(uint64_t));
  return 0;
}

extern "C" int lite_candidates_clear_splits(lite_ctx *c) {
  if (!c) return fail(LITE_EINVAL, "null context");
  c->n_local = 0; c->n_global = 0;
  CU(cudaSetDevice(c->device));
  CU(cudaMemsetAsync(c->status.p, 0, sizeof(unsigned long long), c->stream));
  return 0;
}

extern "C" int lite_candidates_count(lite_ctx *c, int64_t *nl, int64_t *ng) {
  if (!c) return fail(LITE_EINVAL, "null context");
  if (nl) *nl = c->n_local;
  if (ng) *ng = c->n_global;
  return 0;
}

template <int D>
static void launch_reduce(lite_ctx *c, const RedSeg &A, const RedSeg &B, const Theta &th, const PeerArgs &P) {
  k_reduce<D><<<A.grid + B.grid, 256, 0, c->stream>>>(A, B, th, P);
}
static void reduce_dispatch(lite_ctx *c, int d, const RedSeg &A, const RedSeg &B, const Theta &th, const PeerArgs &P) {
  switch (d) {
    case 1: launch_reduce<1>(c, A, B, th, P); break;
    case 2: launch_reduce<2>(c, A, B, th, P); break;
    case 3: launch_reduce<3>(c, A, B, th, P); break;
    case 4: launch_reduce<4>(c, A, B, th, P); break;
    case 5: launch_reduce<5>(c, A, B, th, P); break;
    case 6: launch_reduce<6>(c, A, B, th, P); break;
    case 7: launch_reduce<7>(c, A, B, th, P); break;
    case 8: launch_reduce<8>(c, A, B, th, P); break;
    case 9: launch_reduce<9>(c, A, B, th, P); break;
    case 10: launch_reduce<10>(c, A, B, th, P); break;
    case 11: launch_reduce<11>(c, A, B, th, P); break;
    case 12: launch_reduce<12>(c, A, B, th, P); break;
    default: break;
  }
  c->launches += 1;
}

// GetValue / GetGradient / GetHessian from the reduced sums (lib/ttessel.cpp:3387-3443).
// S: split sums over ALL ranks [S0, S1_k, S2_kl (k<=l)], F: the same over the flips,
// M / Fl: sum_stat_merges / sum_stat_flips.  The normaliser uses the GLOBAL number of
// dummy splits (:3396).  Host arithmetic only: exported so that the CPU tests can drive it.
extern "C" int lite_pl_combine(int d, int64_t n_splits_global, double int_length, const double *theta,
                               const double *S, const double *F, const double *M, const double *Fl,
                               double *lpl, double *grad, double *hess) {
  if (d < 1 || d > LITE_MAX_D || !theta || !S || !F || !M || !Fl) return fail(LITE_EINVAL, "bad argument");
  if (n_splits_global <= 0) return fail(LITE_ENOSPLIT, "no dummy splits: the reference divides by stat_splits.size()");
  // lib/ttessel.cpp:3396: int_length/(pi*|S|)
  const double cn = int_length / (LITE_PI * (double)n_splits_global);
  if (lpl) {
    double tm = 0, tf = 0;
    for (int i = 0; i < d; i++) { tm += theta[i] * M[i]; tf += theta[i] * Fl[i]; }
    double v = -tm;
    v -= cn * S[0];
    v -= tf;
    v -= F[0];
    *lpl = v;
  }
  if (grad)
    for (int i = 0; i < d; i++) grad[i] = -M[i] - cn * S[1 + i] - Fl[i] - F[1 + i];
  if (hess) {
    int t = 1 + d;
    for (int k = 0; k < d; k++)
      for (int l = k; l < d; l++, t++) {
        double v = -cn * S[t] - F[t];
        hess[k * d + l] = v;
        hess[l * d + k] = v;
      }
  }
  return 0;
}

// the contiguous shard of rank r; exported for the CPU tests
extern "C" int lite_shard_range(int64_t n, int rank, int world, int64_t *first, int64_t *count) {
  if (n < 0 || world < 1 || rank < 0 || rank >= world || !first || !count) return fail(LITE_EINVAL, "bad argument");
  shard_range(n, rank, world, *first, *count);
  return 0;
}

extern "C" int lite_pl_eval(lite_ctx *c, const double *theta, double *lpl, double *grad, double *hess) {
  NvtxRange nvtx_("lite_pl_eval");
  if (!c || !theta) return fail(LITE_EINVAL, "null argument");
  if (!c->have_tess || !c->have_prog) return fail(LITE_ESTATE, "upload a tessellation and set the energy first");
  if (!c->have_mf) return fail(LITE_ESTATE, "call lite_candidates_merges_flips (SetEnergy) before evaluating");
  if (c->n_global == 0) return fail(LITE_ENOSPLIT, "no dummy splits: the reference divides by stat_splits.size()");
  CU(cudaSetDevice(c->device));
  { const int fr = mf_flush(c); if (fr) return fr; }
  const int d = c->G.d;
  const int NT = 1 + d + d * (d + 1) / 2;
  Theta th{};
  for (int i = 0; i < d; i++) th.v[i] = theta[i];
  const int maxgrid = c->sm_count * 2;
  if (c->partials.alloc((size_t)2 * maxgrid * NT)) return LITE_ECUDA;
  if (c->mf_inflight) CU(cudaStreamWaitEvent(c->stream, c->ev_mf, 0));  // join the merge/flip stream
  CU(cudaEventRecord(c->ev[2], c->stream));
  // one launch: this rank's shard of the splits, then the flips (replicated on every rank)
  const long n = (long)c->n_local, nf = 2L * c->n_b;
  int grid = (int)((n + 2047) / 2048), gf = (int)((nf + 1023) / 1024);
  if (grid > maxgrid) grid = maxgrid;
  if (gf > maxgrid) gf = maxgrid;
  c->rev = !c->rev;  // boustrophedon: the end the previous pass (or k_split) touched last is still in L2
  // with peers the split segment always runs (at least one block): its last block does the exchange
  const bool exch = c->world > 1 && c->peer_on;
  RedSeg A{c->stat.p, c->cap, n, c->rev ? 1 : 0, c->partials.p, c->counter.p, c->red_out.p,
           (n > 0 || exch) ? (grid < 1 ? 1 : grid) : 0};
  RedSeg B{c->fstat.p, (size_t)nf, nf, 0, c->partials.p + (size_t)maxgrid * NT, c->counter.p + 1, c->red_out.p + 128,
           nf > 0 ? (gf < 1 ? 1 : gf) : 0};
  PeerArgs P{};
  P.world = exch ? c->world : 1; P.rank = c->rank; P.info = c->peer_info.p;
  if (exch) {
    P.seq = ++c->peer_seq;
    for (int r = 0; r < c->world; r++) P.box[r] = c->peer_box[r];
  }
  // timing events only where something separates them: ev[2] also opens the reduce unless a
  // memset precedes it, ev[7] also closes the evaluation unless NCCL follows
  cudaEvent_t red0 = c->ev[2], eval1 = c->ev[7];
  if (A.grid == 0 || nf == 0) {
    if (A.grid == 0) CU(cudaMemsetAsync(c->red_out.p, 0, 128 * sizeof(double), c->stream));
    if (nf == 0) CU(cudaMemsetAsync(c->red_out.p + 128, 0, 128 * sizeof(double), c->stream));
    CU(cudaEventRecord(c->ev[6], c->stream));
    red0 = c->ev[6];
  }
  if (A.grid + B.grid > 0) reduce_dispatch(c, d, A, B, th, P);
  CU(cudaEventRecord(c->ev[7], c->stream));
  if (c->world > 1 && !exch) {
    ncclResult_t ne = g_nccl.AllReduce(c->red_out.p, c->red_out.p, (size_t)NT, ncclFloat64, ncclSum, c->comm, c->stream);
    if (ne != 0) return fail(LITE_ENCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(ne) : "?");
    CU(cudaEventRecord(c->ev[3], c->stream));
    eval1 = c->ev[3];
  }
  // the sums, the geometric failures of the candidates behind them (no exit edge found, see
  // lite_pl_failures), the exchange's report and the validation word come back in one copy
  CU(cudaMemcpyAsync(c->h_pinned, c->res.p, LITE_RES_WORDS * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  {
    const int tr = tess_err_report(c, *(const int *)(c->h_pinned + LITE_RES_ERR));
    if (tr) return tr;
  }
  c->mf_inflight = false;
  finalize_mf(c);
  cudaEventElapsedTime(&c->ms_eval, c->ev[2], eval1);
  cudaEventElapsedTime(&c->ms_reduce, red0, c->ev[7]);
  c->ms_exchange = 0; c->peer_wait_us = 0;
  if (c->world > 1 && !exch) cudaEventElapsedTime(&c->ms_exchange, c->ev[7], c->ev[3]);
  if (c->ev01) { float ms = 0; if (cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]) == cudaSuccess) c->ms_split = ms; }
  {
    const unsigned long long *st = (const unsigned long long *)(c->h_pinned + LITE_RES_STATUS);
    c->fail_splits = (int64_t)st[0]; c->fail_flips = (int64_t)st[1];
    if (st[6] != 0)
      return fail(LITE_ECAPACITY, "%llu candidates on faces with holes needed polygons of more than %d vertices",
                  st[6], LITE_GEN_PTS);
  }
  if (exch) {
    const unsigned long long *pi = (const unsigned long long *)(c->h_pinned + LITE_RES_PEER);
    if (pi[0] == P.seq)
      return fail(LITE_ENCCL, "rank %d: the peers' sums of evaluation %llu did not arrive within %.1f s "
                  "(did every rank call lite_pl_eval?)", c->rank, P.seq, LITE_PEER_TIMEOUT_NS * 1e-9);
    c->peer_wait_us = (double)pi[1] * 1e-3;
    c->ms_exchange = (float)(c->peer_wait_us * 1e-3);
  }
  c->acc_ms[0] += c->ms_split; c->acc_ms[1] += c->ms_reduce; c->acc_ms[2] += c->ms_mf; c->acc_ms[3] += c->ms_exchange; c->acc_ms[4] += 1;
  return lite_pl_combine(d, c->n_global, c->int_length, theta, c->h_pinned, c->h_pinned + 128, c->h_sum_m.data(),
                         c->h_sum_f.data(), lpl, grad, hess);
}

// No-exit candidates among those the last lite_pl_eval summed: a split or flip whose ray found
// no edge to land on contributes an all-zero statistic row (e^0 = 1 to S0), a bias the caller
// should know about.  Counted since the last lite_candidates_clear_splits / SetEnergy.
extern "C" int lite_pl_failures(lite_ctx *c, int64_t *split_no_exit, int64_t *flip_no_exit) {
  if (!c) return fail(LITE_EINVAL, "null context");
  if (split_no_exit) *split_no_exit = c->fail_splits;
  if (flip_no_exit) *flip_no_exit = c->fail_flips;
  return 0;
}

// How lite_pl_eval combines the ranks: 0 single rank, 1 ncclAllReduce, 2 peer memory inside
// k_reduce; exchange_ms = device time the last evaluation spent in that exchange (NCCL: the
// allreduce; peer memory: push + wait for the slowest rank + sum, measured by the kernel).
extern "C" int lite_comm_info(lite_ctx *c, int32_t *mode, float *exchange_ms) {
  if (!c) return fail(LITE_EINVAL, "null context");
  if (mode) *mode = c->world <= 1 ? 0 : (c->peer_on ? 2 : 1);
  if (exchange_ms) *exchange_ms = c->ms_exchange;
  return 0;
}

extern "C" int lite_pl_sums(lite_ctx *c, double *sm, double *sf) {
  if (!c) return fail(LITE_EINVAL, "null context");
  if (!c->have_mf) return fail(LITE_ESTATE, "call lite_candidates_merges_flips first");
  CU(cudaSetDevice(c->device));
  { const int fr = mf_flush(c); if (fr) return fr; }
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaStreamSynchronize(c->stream2));
  c->mf_inflight = false;
  finalize_mf(c);
  { const int tr = tess_err_sync(c); if (tr) return tr; }
  for (int i = 0; i < c->G.d; i++) {
    if (sm) sm[i] = c->h_sum_m[i];
    if (sf) sf[i] = c->h_sum_f[i];
  }
  return 0;
}

template <typename T>
static int fetch_plain(lite_ctx *c, const T *src, size_t have, int64_t first, int64_t count, void *out) {
  if (first < 0 || count < 0 || (size_t)(first + count) > have) return fail(LITE_EINVAL, "fetch range out of bounds");
  if (count == 0) return 0;
  CU(cudaMemcpyAsync(out, src + first, count * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
static int fetch_stat(lite_ctx *c, const double *stat, size_t cap, size_t have, int d, int64_t first,
                      int64_t count, double *out) {
  if (first < 0 || count < 0 || (size_t)(first + count) > have) return fail(LITE_EINVAL, "fetch range out of bounds");
  if (count == 0) return 0;
  std::vector<double> tmp((size_t)count * d);
  CU(cudaMemcpy2DAsync(tmp.data(), count * sizeof(double), stat + first, cap * sizeof(double),
                       count * sizeof(double), d, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  for (int64_t i = 0; i < count; i++)
    for (int k = 0; k < d; k++) out[i * d + k] = tmp[(size_t)k * count + i];
  return 0;
}

extern "C" int lite_debug_fetch(lite_ctx *c, int which, int64_t first, int64_t count, void *out) {
  if (!c || !out) return fail(LITE_EINVAL, "null argument");
  CU(cudaSetDevice(c->device));
  { const int fr = mf_flush(c); if (fr) return fr; }
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaStreamSynchronize(c->stream2));
  c->mf_inflight = false;
  finalize_mf(c);
  { const int tr = tess_err_sync(c); if (tr) return tr; }
  const size_t ns = (size_t)c->n_local, nm = (size_t)c->n_nb, nf = (size_t)2 * c->n_b;
  const bool rec = c->rcap > 0;
  switch (which) {
    case LITE_FETCH_SPLIT_STAT: return fetch_stat(c, c->stat.p, c->cap, ns, c->G.d, first, count, (double *)out);
    case LITE_FETCH_MERGE_STAT:
      if (!c->have_mf) break;
      return fetch_stat(c, c->mstat.p, nm, nm, c->G.d, first, count, (double *)out);
    case LITE_FETCH_FLIP_STAT:
      if (!c->have_mf) break;
      return fetch_stat(c, c->fstat.p, nf, nf, c->G.d, first, count, (double *)out);
    case LITE_FETCH_MERGE_HE: if (!c->have_mf) break; return fetch_plain(c, c->m_he.p, nm, first, count, out);
    case LITE_FETCH_FLIP_E1: if (!c->have_mf) break; return fetch_plain(c, c->fl_e1.p, nf, first, count, out);
    case LITE_FETCH_FLIP_E2: if (!c->have_mf) break; return fetch_plain(c, c->fl_e2.p, nf, first, count, out);
    case LITE_FETCH_FLIP_P2: if (!c->have_mf) break; return fetch_plain(c, c->fl_p2.p, nf, first, count, out);
    case LITE_FETCH_FLIP_RIGHT: if (!c->have_mf) break; return fetch_plain(c, c->fl_right.p, nf, first, count, out);
    case LITE_FETCH_STATUS: {
      CU(cudaMemcpyAsync(out, c->status.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
      CU(cudaStreamSynchronize(c->stream));
      return 0;
    }
    default: break;
  }
  if (which == LITE_FETCH_SPLIT_E1 || which == LITE_FETCH_SPLIT_E2 || which == LITE_FETCH_SPLIT_P1 ||
      which == LITE_FETCH_SPLIT_P2 || which == LITE_FETCH_SPLIT_X || which == LITE_FETCH_SPLIT_ANGLE ||
      which == LITE_FETCH_SPLIT_U) {
    if (!rec) return fail(LITE_ESTATE, "per-split records need LITE_OPT_RECORD_SPLITS before adding splits");
    size_t have = c->rcap;
    switch (which) {
      case LITE_FETCH_SPLIT_E1: return fetch_plain(c, c->r_e1.p, have, first, count, out);
      case LITE_FETCH_SPLIT_E2: return fetch_plain(c, c->r_e2.p, have, first, count, out);
      case LITE_FETCH_SPLIT_P1: return fetch_plain(c, c->r_p1.p, have, first, count, out);
      case LITE_FETCH_SPLIT_P2: return fetch_plain(c, c->r_p2.p, have, first, count, out);
      case LITE_FETCH_SPLIT_X: return fetch_plain(c, c->r_x.p, have, first, count, out);
      case LITE_FETCH_SPLIT_ANGLE: return fetch_plain(c, c->r_angle.p, have, first, count, out);
      default: return fetch_plain(c, c->r_u.p, have, first, count, out);
    }
  }
  return fail(LITE_ESTATE, "nothing to fetch for selector %d (call order?)", which);
}

extern "C" int lite_last_kernel_ms(lite_ctx *c, float *a, float *b, float *m) {
  if (!c) return fail(LITE_EINVAL, "null context");
  CU(cudaSetDevice(c->device));
  { const int fr = mf_flush(c); if (fr) return fr; }
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaStreamSynchronize(c->stream2));
  c->mf_inflight = false;
  finalize_mf(c);
  float ms = 0;
  if (c->ev01 && cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]) == cudaSuccess) c->ms_split = ms;
  if (a) *a = c->ms_split;
  if (b) *b = c->ms_eval;
  if (m) *m = c->ms_mf;
  return 0;
}

// Device times summed over the lite_pl_eval calls since the last reset (so that a timed loop
// need not query after every step): [0] split kernel, [1] theta-reduce kernel, [2] merges+flips
// kernel, [3] cross-GPU exchange, [4] number of evaluations.
extern "C" int lite_timers_accumulated(lite_ctx *c, double out5[5], int reset) {
  if (!c || !out5) return fail(LITE_EINVAL, "null argument");
  for (int k = 0; k < 5; k++) out5[k] = c->acc_ms[k];
  if (reset) for (int k = 0; k < 5; k++) c->acc_ms[k] = 0;
  return 0;
}

extern "C" int lite_last_reduce_ms(lite_ctx *c, float *ms) {
  if (!c || !ms) return fail(LITE_EINVAL, "null argument");
  *ms = c->ms_reduce;
  return 0;
}

extern "C" int lite_timer_mark(lite_ctx *c, int slot) {
  if (!c || slot < 0 || slot >= LITE_TIMER_SLOTS) return fail(LITE_EINVAL, "bad timer slot");
  CU(cudaSetDevice(c->device));
  CU(cudaEventRecord(c->tev[slot], c->stream));
  return 0;
}

extern "C" int lite_timer_elapsed(lite_ctx *c, int slot_a, int slot_b, float *ms) {
  if (!c || !ms || slot_a < 0 || slot_a >= LITE_TIMER_SLOTS || slot_b < 0 || slot_b >= LITE_TIMER_SLOTS)
    return fail(LITE_EINVAL, "bad timer slot");
  CU(cudaSetDevice(c->device));
  CU(cudaEventSynchronize(c->tev[slot_b]));
  CU(cudaEventElapsedTime(ms, c->tev[slot_a], c->tev[slot_b]));
  return 0;
}

extern "C" int64_t lite_kernel_launch_count(lite_ctx *c) { return c ? c->launches : 0; }

extern "C" int lite_measure_fp64_peak(lite_ctx *c, double *tflops) {
  if (!c || !tflops) return fail(LITE_EINVAL, "null argument");
  CU(cudaSetDevice(c->device));
  const int iters = 1 << 14, tb = 256, gb = c->sm_count * 8;
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    CU(cudaEventRecord(c->ev[4], c->stream));
    k_dfma_peak<<<gb, tb, 0, c->stream>>>(c->red_out.p ? c->red_out.p : (double *)c->status.p, iters, 1.0000001, 1e-9);
    CU(cudaEventRecord(c->ev[5], c->stream));
    CU(cudaStreamSynchronize(c->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]);
    double fl = 2.0 * 8.0 * (double)iters * tb * gb;
    double tf = fl / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  c->launches += 5;
  *tflops = best;
  return 0;
}

#include "smf_ensemble.cuh"
