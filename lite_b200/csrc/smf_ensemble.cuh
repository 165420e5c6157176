// smf_ensemble.cuh — an ensemble of independent SMF chains on the GPU (included at
// the end of capi.cu).  Config "ensemble of 65536 SMF chains" of BASELINE.json: the
// reference runs applications/checkACS.cpp once per seed (one chain per process);
// here chain i is one thread, its tessellation one block of HBM
// (smf_chain.h), its random numbers the Philox stream (seed, global chain id,
// proposal index), so a chain's trajectory does not depend on how the ensemble is
// sharded over GPUs.  No collective: chains never exchange anything.
#pragma once
#include "smf_chain.h"

#ifndef LITE_SMF_TB
#define LITE_SMF_TB 64
#endif

struct SmfState {
  char *blob = nullptr;
  size_t stride = 0;
  int cap = 0;
  int64_t n_local = 0, n_global = 0, first = 0;  // this rank holds chains [first, first + n_local)
  uint64_t seed = 0;
  double rect[4] = {0, 0, 1, 1};
  smf::Energy G{};
  smf::ChainParams cp{0.33, 0.33};
  bool have_model = false;
  double *rows = nullptr; size_t rows_cap = 0;         // device [n_samples][12][n_local]
  smf::TraceRec *trace = nullptr; size_t trace_cap = 0;
  int *bad = nullptr;  // [0] chains stopped by capacity, [1] broken
  float ms_last = 0;
};

__global__ void __launch_bounds__(LITE_SMF_TB) k_smf_init(char *blob, size_t stride, int cap, int n, double x0,
                                                          double y0, double x1, double y1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  smf::Header h_loc;
  smf::Chain c;
  c.H = &h_loc;
  smf::bind(c, blob + (size_t)i * stride, cap);
  smf::init_window(c, cap, x0, y0, x1, y1);
  smf::store(c);
}

// n_samples batches of `steps` proposals per chain; after each batch the chain
// writes the checkACS columns (applications/checkACS.cpp:136-148) of its state,
// with the proposal / acceptance counts of that batch (SMFChain::step returns the
// counts of one call, lib/ttessel.cpp:3257-3311).
//
// The kernel is bound by the latency of dependent loads (a DCEL is pointer chasing),
// so what matters is how many chains are in flight.  MINB blocks of 64 threads per SM
// fix the register budget (4 -> 255 registers, 8 -> 128, 16 -> 64); lpc_shift lets only
// every (1 << lpc_shift)-th thread take a chain (an experiment knob: spreading the chains
// over more warps measured slower than packing them, see lite_smf_run).
template <int MINB>
__global__ void __launch_bounds__(LITE_SMF_TB, MINB) k_smf_run(char *blob, size_t stride, int cap, int n, uint64_t first,
                                                               smf::Energy G_arg, smf::ChainParams cp, uint64_t seed,
                                                               int n_samples, long steps, double *rows, int *bad,
                                                               int lpc_shift) {
  // the energy is read through a reference by the out-of-line steps: one copy per block
  // in shared memory instead of one per thread in local memory
  __shared__ smf::Energy sh_G;
  if (threadIdx.x == 0) sh_G = G_arg;
  __syncthreads();
  const smf::Energy &G = sh_G;
  const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gt & ((1L << lpc_shift) - 1)) return;
  const long i = gt >> lpc_shift;
  if (i >= n) return;
  smf::Chain c;
  smf::Header h_loc;  // working copy of the header (shared memory measured no faster)
  c.H = &h_loc;
  smf::bind(c, blob + (size_t)i * stride, cap);
  smf::load(c);
  smf::StepRng R;
  R.seed = seed; R.chain = first + (uint64_t)i;
  for (int s = 0; s < n_samples; s++) {
    for (int k = 0; k < 6; k++) c.H->counts[k] = 0;
    for (long t = 0; t < steps; t++)
      if (!smf::smf_step(c, G, cp, R, nullptr)) break;
    if (rows) smf::stat_row(c, rows + (size_t)s * smf::N_STAT_COLS * n + i, (size_t)n);
  }
  if (c.H->status == smf::ST_CAPACITY) atomicAdd(bad, 1);
  else if (c.H->status != smf::ST_OK) atomicAdd(bad + 1, 1);
  smf::store(c);
}

// one chain, every proposal recorded (parity tests)
__global__ void k_smf_trace(char *blob, size_t stride, int cap, int i, uint64_t first, smf::Energy G,
                            smf::ChainParams cp, uint64_t seed, long steps, smf::TraceRec *out, int *done) {
  smf::Header h_loc;
  smf::Chain c;
  c.H = &h_loc;
  smf::bind(c, blob + (size_t)i * stride, cap);
  smf::load(c);
  smf::StepRng R;
  R.seed = seed; R.chain = first + (uint64_t)i;
  long t = 0;
  for (; t < steps; t++)
    if (!smf::smf_step(c, G, cp, R, out + t)) break;
  *done = (int)t;
  smf::store(c);
}

__global__ void __launch_bounds__(LITE_SMF_TB) k_smf_stats(char *blob, size_t stride, int cap, int n, double *rows,
                                                           int *check_fail) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  smf::Header h_loc;
  smf::Chain c;
  c.H = &h_loc;
  smf::bind(c, blob + (size_t)i * stride, cap);
  smf::load(c);
  smf::stat_row(c, rows + i, (size_t)n);
  if (check_fail && smf::check(c) != 0) atomicAdd(check_fail, 1);
}

static void smf_release(lite_ctx *c) {
  SmfState *S = c->smf;
  if (!S) return;
  if (S->blob) cudaFree(S->blob);
  if (S->rows) cudaFree(S->rows);
  if (S->trace) cudaFree(S->trace);
  if (S->bad) cudaFree(S->bad);
  delete S;
  c->smf = nullptr;
}

extern "C" int lite_smf_create(lite_ctx *c, int64_t n_chains_global, int32_t max_segments, double x0, double y0,
                               double x1, double y1, uint64_t seed) {
  if (!c) return fail(LITE_EINVAL, "null context");
  if (n_chains_global < 1 || !(x1 > x0) || !(y1 > y0)) return fail(LITE_EINVAL, "bad ensemble size or window");
  // 16-bit halfedge ids: 6 * cap < 65535
  if (max_segments < 1 || max_segments + 4 > 10900)
    return fail(LITE_EINVAL, "max_segments must be in [1, 10896] (16-bit element ids of the chain layout)");
  CU(cudaSetDevice(c->device));
  smf_release(c);
  SmfState *S = new SmfState();
  c->smf = S;
  S->cap = max_segments + 4;
  S->stride = smf::chain_bytes(S->cap);
  S->n_global = n_chains_global;
  shard_range(n_chains_global, c->rank, c->world, S->first, S->n_local);
  S->seed = seed;
  S->rect[0] = x0; S->rect[1] = y0; S->rect[2] = x1; S->rect[3] = y1;
  if (S->n_local > 0x7fffffff) return fail(LITE_EINVAL, "too many chains for one GPU");
  const size_t bytes = S->stride * (size_t)(S->n_local > 0 ? S->n_local : 1);
  cudaError_t e = cudaMalloc((void **)&S->blob, bytes);
  if (e != cudaSuccess) {
    smf_release(c);
    return fail(LITE_ECUDA, "cudaMalloc(%zu B) for %lld chains of capacity %d segments: %s", bytes,
                (long long)S->n_local, max_segments, cudaGetErrorString(e));
  }
  CU(cudaMalloc((void **)&S->bad, 4 * sizeof(int)));
  CU(cudaMemsetAsync(S->bad, 0, 4 * sizeof(int), c->stream));
  if (S->n_local > 0) {
    const int n = (int)S->n_local;
    k_smf_init<<<(n + LITE_SMF_TB - 1) / LITE_SMF_TB, LITE_SMF_TB, 0, c->stream>>>(S->blob, S->stride, S->cap, n, x0, y0, x1, y1);
    c->launches++;
    CU(cudaGetLastError());
  }
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int lite_smf_destroy(lite_ctx *c) {
  if (!c) return fail(LITE_EINVAL, "null context");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  smf_release(c);
  return 0;
}

extern "C" int lite_smf_set_model(lite_ctx *c, int d, const int32_t *fid, const double *theta, double p_split,
                                  double p_merge) {
  if (!c || !c->smf) return fail(LITE_ESTATE, "lite_smf_create first");
  if (d < 1 || d > LITE_MAX_D || !fid || !theta) return fail(LITE_EINVAL, "bad energy (1 <= d <= %d)", LITE_MAX_D);
  if (!(p_split > 0) || !(p_merge > 0) || p_split + p_merge > 1)
    return fail(LITE_EINVAL, "SMFChain: need p_split > 0, p_merge > 0, p_split + p_merge <= 1 (lib/ttessel.cpp:3122-3127)");
  SmfState *S = c->smf;
  S->G.d = d; S->G.mask = 0;
  for (int i = 0; i < d; i++) {
    if (fid[i] < 0 || fid[i] >= LITE_FEAT_COUNT) return fail(LITE_EINVAL, "unknown feature id %d: no device twin, no CPU fallback", fid[i]);
    S->G.fid[i] = fid[i]; S->G.theta[i] = theta[i]; S->G.mask |= 1u << fid[i];
  }
  S->cp.p_split = p_split; S->cp.p_merge = p_merge;
  S->have_model = true;
  return 0;
}

static int smf_report_bad(lite_ctx *c) {
  SmfState *S = c->smf;
  int bad[2] = {0, 0};
  CU(cudaMemcpyAsync(bad, S->bad, sizeof bad, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (bad[0] || bad[1]) {
    CU(cudaMemsetAsync(S->bad, 0, 4 * sizeof(int), c->stream));
    if (bad[1]) return fail(LITE_EINVAL, "%d chain(s) reached an inconsistent state", bad[1]);
    return fail(LITE_ECAPACITY, "%d chain(s) stopped: more than %d internal segments; recreate the ensemble with a larger max_segments",
                bad[0], S->cap - 4);
  }
  return 0;
}

extern "C" int lite_smf_run(lite_ctx *c, int32_t n_samples, int64_t steps_per_sample, double *rows_out) {
  NvtxRange nvtx_("lite_smf_run");
  if (!c || !c->smf) return fail(LITE_ESTATE, "lite_smf_create first");
  SmfState *S = c->smf;
  if (!S->have_model) return fail(LITE_ESTATE, "lite_smf_set_model first");
  if (n_samples < 1 || steps_per_sample < 0) return fail(LITE_EINVAL, "bad sample count / batch length");
  CU(cudaSetDevice(c->device));
  const int n = (int)S->n_local;
  const size_t need = (size_t)n_samples * smf::N_STAT_COLS * (size_t)(n > 0 ? n : 1);
  if (rows_out && need > S->rows_cap) {
    if (S->rows) cudaFree(S->rows);
    S->rows = nullptr; S->rows_cap = 0;
    CU(cudaMalloc((void **)&S->rows, need * sizeof(double)));
    S->rows_cap = need;
  }
  CU(cudaEventRecord(c->ev[6], c->stream));
  if (n > 0) {
    // Launch shape, from sweeps on B200 (gpurun_out/smf_sweep1.log, smf_ab2.log, smf_ab3.log;
    // LITE_SMF_MINB / LITE_SMF_LPC override it for such sweeps).  The kernel is bound by
    // the latency of dependent loads, so what counts is how the chains share warps and SMs:
    //  * the diverged lanes of a warp overlap their misses (32 chains per warp cost 1.8x one
    //    chain per warp), so chains are packed, 16 per warp while the ensemble still fits
    //    the GPU that way and 32 otherwise; thinner warps lose again (the lanes of a warp
    //    share local-memory lines, sparse warps waste L1);
    //  * an ensemble that fits at 255 registers per thread runs the unconstrained build in
    //    one-warp blocks spread over all SMs; a larger one runs at 128 registers, 512 chains
    //    resident per SM (64 registers measured slower).
    static const char *env_minb = getenv("LITE_SMF_MINB"), *env_lpc = getenv("LITE_SMF_LPC");

    int shift = ((long)n * 2 <= (long)c->sm_count * 4 * LITE_SMF_TB) ? 1 : 0;
    if (env_lpc) { shift = 0; while ((1 << (shift + 1)) <= atoi(env_lpc) && shift < 5) shift++; }
    const long threads = (long)n << shift;
    int minb = (threads <= (long)c->sm_count * 4 * LITE_SMF_TB) ? 4 : 8;
    if (env_minb) minb = atoi(env_minb);
    const int tb = (minb <= 4) ? 32 : LITE_SMF_TB;
    // requests for the lines a move is about to touch (smf::prefetch_around): +2.4 % for an
    // ensemble that leaves most of the GPU's load slots idle (8192 chains: 19.47 -> 19.02 ms per
    // 250 proposals), -4 % for one that fills it (65 536 chains: the extra requests compete with
    // the loads), scripts/ab_smf_prefetch.sh
    S->cp.prefetch = (minb <= 4) ? 7 : 0;
    if (const char *e = getenv("LITE_SMF_PREFETCH")) S->cp.prefetch = atoi(e);
    const unsigned grid = (unsigned)((threads + tb - 1) / tb);
    double *rows = rows_out ? S->rows : nullptr;
#define SMF_LAUNCH(M)                                                                                             \
  k_smf_run<M><<<grid, tb, 0, c->stream>>>(S->blob, S->stride, S->cap, n, (uint64_t)S->first, S->G, S->cp, S->seed, \
                                           n_samples, (long)steps_per_sample, rows, S->bad, shift)
    if (minb <= 4) SMF_LAUNCH(4);
    else if (minb <= 8) SMF_LAUNCH(8);
    else SMF_LAUNCH(16);
#undef SMF_LAUNCH
    c->launches++;
    CU(cudaGetLastError());
  }
  CU(cudaEventRecord(c->ev[7], c->stream));
  if (rows_out && n > 0)
    CU(cudaMemcpyAsync(rows_out, S->rows, need * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  int r = smf_report_bad(c);
  cudaEventElapsedTime(&S->ms_last, c->ev[6], c->ev[7]);
  return r;
}

extern "C" int lite_smf_last_ms(lite_ctx *c, float *ms) {
  if (!c || !c->smf || !ms) return fail(LITE_ESTATE, "lite_smf_create first");
  *ms = c->smf->ms_last;
  return 0;
}

extern "C" int lite_smf_stats(lite_ctx *c, double *rows_out, int32_t validate) {
  if (!c || !c->smf || !rows_out) return fail(LITE_ESTATE, "lite_smf_create first");
  SmfState *S = c->smf;
  CU(cudaSetDevice(c->device));
  const int n = (int)S->n_local;
  if (n == 0) return 0;
  const size_t need = (size_t)smf::N_STAT_COLS * n;
  if (need > S->rows_cap) {
    if (S->rows) cudaFree(S->rows);
    S->rows = nullptr; S->rows_cap = 0;
    CU(cudaMalloc((void **)&S->rows, need * sizeof(double)));
    S->rows_cap = need;
  }
  CU(cudaMemsetAsync(S->bad + 2, 0, sizeof(int), c->stream));
  k_smf_stats<<<(n + LITE_SMF_TB - 1) / LITE_SMF_TB, LITE_SMF_TB, 0, c->stream>>>(S->blob, S->stride, S->cap, n, S->rows,
                                                                             validate ? S->bad + 2 : nullptr);
  c->launches++;
  CU(cudaGetLastError());
  int nfail = 0;
  CU(cudaMemcpyAsync(rows_out, S->rows, need * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(&nfail, S->bad + 2, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (nfail) return fail(LITE_EINVAL, "%d chain(s) fail the consistency check", nfail);
  return 0;
}

extern "C" int lite_smf_trace(lite_ctx *c, int64_t chain_local, int64_t n_steps, void *trace_out, int64_t *n_done) {
  if (!c || !c->smf || !trace_out) return fail(LITE_ESTATE, "lite_smf_create first");
  SmfState *S = c->smf;
  if (!S->have_model) return fail(LITE_ESTATE, "lite_smf_set_model first");
  if (chain_local < 0 || chain_local >= S->n_local || n_steps < 1) return fail(LITE_EINVAL, "bad chain index or step count");
  CU(cudaSetDevice(c->device));
  if ((size_t)n_steps > S->trace_cap) {
    if (S->trace) cudaFree(S->trace);
    S->trace = nullptr; S->trace_cap = 0;
    CU(cudaMalloc((void **)&S->trace, (size_t)n_steps * sizeof(smf::TraceRec)));
    S->trace_cap = (size_t)n_steps;
  }
  k_smf_trace<<<1, 1, 0, c->stream>>>(S->blob, S->stride, S->cap, (int)chain_local, (uint64_t)S->first, S->G, S->cp,
                                      S->seed, (long)n_steps, S->trace, S->bad + 3);
  c->launches++;
  CU(cudaGetLastError());
  int done = 0;
  CU(cudaMemcpyAsync(&done, S->bad + 3, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaMemcpy(trace_out, S->trace, (size_t)done * sizeof(smf::TraceRec), cudaMemcpyDeviceToHost));
  if (n_done) *n_done = done;
  return 0;
}

// Intersections of the internal edges of every chain with a few probe segments: the
// `inter_*.txt` output of applications/checkACS.cpp:151-172 (there the probes are the four window
// edges and the two mid-lines, and a row is written when CGAL::intersection(probe, edge) is a single
// point: crossings and touching end points, not collinear overlaps).
__global__ void __launch_bounds__(LITE_SMF_TB) k_smf_intersect(char *blob, size_t stride, int cap, int n, int n_probes,
                                                               const double *probes, long long out_cap, int *o_chain,
                                                               int *o_probe, double2 *o_xy, unsigned long long *count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  smf::Header h_loc;
  smf::Chain c;
  c.H = &h_loc;
  smf::bind(c, blob + (size_t)i * stride, cap);
  smf::load(c);
  for (int p = 0; p < c.H->n_p; p++) {
    if (!smf::pair_alive(c, p) || smf::h_bnd(c, 2 * p)) continue;   // both faces bounded: an internal edge
    const smf::Pt a = smf::P_src(c, 2 * p), b = smf::P_src(c, 2 * p + 1);
    for (int k = 0; k < n_probes; k++) {
      const double cx = probes[4 * k], cy = probes[4 * k + 1], dx = probes[4 * k + 2], dy = probes[4 * k + 3];
      const double o1 = (b.x - a.x) * (cy - a.y) - (b.y - a.y) * (cx - a.x);   // probe ends against the edge's line
      const double o2 = (b.x - a.x) * (dy - a.y) - (b.y - a.y) * (dx - a.x);
      const double o3 = (dx - cx) * (a.y - cy) - (dy - cy) * (a.x - cx);       // edge ends against the probe's line
      const double o4 = (dx - cx) * (b.y - cy) - (dy - cy) * (b.x - cx);
      if (o1 == 0.0 && o2 == 0.0) continue;                                     // collinear: not a point
      const bool s12 = (o1 <= 0.0 && o2 >= 0.0) || (o1 >= 0.0 && o2 <= 0.0);
      const bool s34 = (o3 <= 0.0 && o4 >= 0.0) || (o3 >= 0.0 && o4 <= 0.0);
      if (!s12 || !s34) continue;
      const double t = o3 / (o3 - o4);                                          // along the edge a -> b
      const unsigned long long at = atomicAdd(count, 1ull);
      if ((long long)at < out_cap) {
        o_chain[at] = i; o_probe[at] = k;
        o_xy[at] = make_double2(fma(t, b.x - a.x, a.x), fma(t, b.y - a.y, a.y));
      }
    }
  }
}

extern "C" int lite_smf_intersections(lite_ctx *c, int32_t n_probes, const double *probes, int64_t cap,
                                      int32_t *chain_out, int32_t *probe_out, double *xy_out, int64_t *n_found) {
  if (!c || !c->smf || !probes || !n_found) return fail(LITE_ESTATE, "lite_smf_create first");
  if (n_probes < 1 || n_probes > 64 || cap < 0) return fail(LITE_EINVAL, "bad probe count or capacity");
  SmfState *S = c->smf;
  CU(cudaSetDevice(c->device));
  const int n = (int)S->n_local;
  *n_found = 0;
  if (n == 0) return 0;
  double *d_probes = nullptr; int *d_chain = nullptr, *d_probe = nullptr; double2 *d_xy = nullptr;
  unsigned long long *d_count = nullptr, found = 0;
  const size_t m = cap > 0 ? (size_t)cap : 1;
  int rc = 0;
  if (cudaMalloc((void **)&d_probes, 4 * n_probes * sizeof(double)) != cudaSuccess || cudaMalloc((void **)&d_chain, m * sizeof(int)) != cudaSuccess ||
      cudaMalloc((void **)&d_probe, m * sizeof(int)) != cudaSuccess || cudaMalloc((void **)&d_xy, m * sizeof(double2)) != cudaSuccess ||
      cudaMalloc((void **)&d_count, sizeof(unsigned long long)) != cudaSuccess)
    rc = fail(LITE_ECUDA, "cudaMalloc for %lld intersection records failed", (long long)cap);
  if (!rc) {
    cudaMemcpyAsync(d_probes, probes, 4 * n_probes * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), c->stream);
    k_smf_intersect<<<(n + LITE_SMF_TB - 1) / LITE_SMF_TB, LITE_SMF_TB, 0, c->stream>>>(S->blob, S->stride, S->cap, n, n_probes, d_probes,
                                                                                  (long long)cap, d_chain, d_probe, d_xy, d_count);
    c->launches++;
    cudaMemcpyAsync(&found, d_count, sizeof found, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(LITE_ECUDA, "k_smf_intersect: %s", cudaGetErrorString(cudaGetLastError()));
  }
  if (!rc) {
    *n_found = (int64_t)found;
    const size_t k = (size_t)((int64_t)found < cap ? (int64_t)found : cap);
    if (k && chain_out) cudaMemcpy(chain_out, d_chain, k * sizeof(int), cudaMemcpyDeviceToHost);
    if (k && probe_out) cudaMemcpy(probe_out, d_probe, k * sizeof(int), cudaMemcpyDeviceToHost);
    if (k && xy_out) cudaMemcpy(xy_out, d_xy, k * sizeof(double2), cudaMemcpyDeviceToHost);
    if ((int64_t)found > cap && (chain_out || probe_out || xy_out))
      rc = fail(LITE_ECAPACITY, "%lld intersections found, room for %lld: call again with that capacity", (long long)found, (long long)cap);
  }
  cudaFree(d_probes); cudaFree(d_chain); cudaFree(d_probe); cudaFree(d_xy); cudaFree(d_count);
  return rc;
}

struct lite_host_tess;
extern "C" int lite_smf_blob_to_host_(char *blob, int cap, lite_host_tess **out);  // host_capi.cpp

extern "C" int lite_smf_export(lite_ctx *c, int64_t chain_local, lite_host_tess **out) {
  if (!c || !c->smf || !out) return fail(LITE_ESTATE, "lite_smf_create first");
  SmfState *S = c->smf;
  if (chain_local < 0 || chain_local >= S->n_local) return fail(LITE_EINVAL, "bad chain index");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  std::vector<char> raw(S->stride + 256);
  char *base = (char *)(((uintptr_t)raw.data() + 255) & ~(uintptr_t)255);
  CU(cudaMemcpy(base, S->blob + (size_t)chain_local * S->stride, S->stride, cudaMemcpyDeviceToHost));
  return lite_smf_blob_to_host_(base, S->cap, out);
}

extern "C" int lite_smf_count(lite_ctx *c, int64_t *n_local, int64_t *n_global, int64_t *first_global,
                              int64_t *bytes_per_chain) {
  if (!c || !c->smf) return fail(LITE_ESTATE, "lite_smf_create first");
  if (n_local) *n_local = c->smf->n_local;
  if (n_global) *n_global = c->smf->n_global;
  if (first_global) *first_global = c->smf->first;
  if (bytes_per_chain) *bytes_per_chain = (int64_t)c->smf->stride;
  return 0;
}
