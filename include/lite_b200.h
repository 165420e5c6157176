/* lite_b200.h — C ABI of the B200-native pseudo-likelihood hot path of LiTe.
 *
 * The reference (kien-kieu/lite) has no FFI boundary: its "operator API" for
 * this path is the C++ class API PseudoLikDiscrete / Energy / TTessel
 * (reference include/ttessel.h:966-1020, :1201-1240, :747-753) with host
 * function pointers as features (:932-937).  This header is the thin C layer a
 * binding to that API would call; include/ttessel.h (this repo) is the
 * source-compatible C++ facade built on it.  Each entry point cites the
 * reference member it replaces.
 *
 * Conventions: every call returns 0 on success, a negative LITE_E* code
 * otherwise; lite_last_error() gives the message (thread local).  A context
 * owns one CUDA device, one stream and (world_size > 1) one NCCL communicator;
 * it is not thread safe, distinct contexts are independent.  All host arrays
 * are caller owned and copied during the call.  Indices are int32, -1 = null.
 * There is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef LITE_B200_H
#define LITE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LITE_OK 0
#define LITE_EINVAL (-1)   /* bad argument / inconsistent tessellation      */
#define LITE_ECUDA (-2)    /* CUDA runtime error or no device               */
#define LITE_ENCCL (-3)    /* NCCL unavailable or failed                    */
#define LITE_ESTATE (-4)   /* call sequence error (e.g. eval before upload) */
#define LITE_ENOSPLIT (-5) /* pl_eval with zero dummy splits: the reference
                              divides by stat_splits.size() (lib/ttessel.cpp:3396) */

#define LITE_ECAPACITY (-6) /* an SMF chain outgrew the capacity the ensemble was created with */

#define LITE_MAX_D 12

/* Device feature ids.  The facade maps the addresses of the reference's
 * built-in feature functions (include/ttessel.h:1384-1405) to these. */
enum lite_feature {
  LITE_FEAT_IS_POINT_INSIDE_WINDOW = 0,    /* vertices  lib/ttessel.cpp:4131 */
  LITE_FEAT_EDGE_LENGTH = 1,               /* edges     :4141 (as written: boundary edges only) */
  LITE_FEAT_SEG_NUMBER = 2,                /* segs      :4157 */
  LITE_FEAT_IS_SEGMENT_INTERNAL = 3,       /* segs      :4172 */
  LITE_FEAT_MINUS_IS_SEGMENT_INTERNAL = 4, /* segs      include/ttessel.h:1396 */
  LITE_FEAT_SEGMENT_SIZE_2 = 5,            /* segs      :4424 (as written: boundary segments only) */
  LITE_FEAT_FACE_NUMBER = 6,               /* faces     :4206 */
  LITE_FEAT_FACE_AREA_2 = 7,               /* faces     :4215 */
  LITE_FEAT_FACE_PERIMETER = 8,            /* faces     :4229 */
  LITE_FEAT_FACE_SHAPE = 9,                /* faces     :4247 */
  LITE_FEAT_FACE_SUM_OF_ANGLES = 10,       /* faces     :4285 */
  LITE_FEAT_MIN_ANGLE = 11,                /* faces     :4348 */
  LITE_FEAT_EDGE_LENGTH_INTERNAL = 12,     /* edges     EXTENSION, not in the reference: length of
                                              internal edges (the evident intent of edge_length) */
  LITE_FEAT_COUNT = 13
};

/* Flattened tessellation: the payload of LineTes_halfedge
 * (include/ttessel.h:445-450) plus the DCEL links, in the reference's halfedge
 * container order (TTessel::alt_length_weighted_random_halfedge depends on it,
 * lib/ttessel.cpp:1977-1989).  Per-face halfedge lists start at outer_ccb(). */
typedef struct lite_tess_soa {
  int32_t n_vertices;
  const double *vx, *vy;
  int32_t n_halfedges;
  const int32_t *he_src, *he_twin, *he_next, *he_face, *he_seg;
  const int32_t *he_next_hf, *he_prev_hf; /* along-segment links, -1 = null */
  const uint8_t *he_dir;                  /* LineTes_halfedge::get_dir()    */
  const double *he_length;                /* LineTes_halfedge::get_length() */
  int32_t n_faces;
  const uint8_t *face_inside;             /* Face::data()                   */
  const int32_t *face_he_offset, *face_he_count; /* into face_he_list; count 0 for outside faces */
  int32_t n_face_he;
  const int32_t *face_he_list;
  int32_t n_segments;
  const int32_t *seg_he_start, *seg_he_end; /* Seg::halfedges_start()/end() */
  const int32_t *seg_n_edges;
  const uint8_t *seg_on_boundary;
  int32_t n_non_blocking, n_blocking;       /* TTessel private lists, container order */
  const int32_t *non_blocking, *blocking;   /* segment ids */
  double int_length;                        /* TTessel::get_total_internal_length() */
} lite_tess_soa;

typedef struct lite_ctx lite_ctx;

const char *lite_last_error(void);

/* 128-byte NCCL unique id for rank 0 to broadcast (out-of-band, e.g. through
 * torch.distributed) before lite_ctx_create on every rank. */
int lite_nccl_unique_id(void *out128);

/* One context per process/GPU.  nccl_unique_id may be NULL when world_size==1. */
int lite_ctx_create(lite_ctx **out, int device, int rank, int world_size,
                    const void *nccl_unique_id);
int lite_ctx_destroy(lite_ctx *ctx);
int lite_ctx_sync(lite_ctx *ctx);
/* option ids */
#define LITE_OPT_RECORD_SPLITS 1 /* keep per-split e1,e2,p1,p2,x,angle for lite_debug_fetch */
/* Which joint law lite_candidates_add_splits_philox / lite_lines_sample_clip give the positions
 * of a batch along the cumulative halfedge length (the marginal law of every candidate is the
 * reference's in both: halfedge with probability proportional to its length, then x, angle):
 *   LITE_SAMPLER_STRATIFIED (default): one position per stratum of width L/n, jittered inside it.
 *     Unbiased for the pseudo-likelihood sums, lower variance, no sort; NOT the joint law of the
 *     reference (its n positions are independent).
 *   LITE_SAMPLER_IID: n independent uniform positions, sorted, exactly as TTessel::split_sample /
 *     alt_length_weighted_random_halfedge(n) draw and sort them (lib/ttessel.cpp:1968-1991). */
/* lite_tess_upload checks the consistency of what it is given (index ranges, twin links, face
 * lists against he_next / he_face, segment lists: the export assertions, in the spirit of
 * TTessel::is_valid).  By default the checks run on the DEVICE behind the copy: the upload returns
 * at once, nothing dereferences an unchecked index, every kernel does nothing on a tessellation
 * that failed, and the failure is reported (LITE_EINVAL) by the next call that synchronises
 * (lite_pl_eval, lite_pl_sums, lite_debug_fetch, lite_ctx_sync, ...).  With this option set they run
 * on the host inside lite_tess_upload, which then reports them itself (tessellations with holes
 * always take this route). */
#define LITE_OPT_VALIDATE_HOST 3
#define LITE_OPT_SAMPLER 2
#define LITE_SAMPLER_STRATIFIED 0
#define LITE_SAMPLER_IID 1
int lite_ctx_set_option(lite_ctx *ctx, int option, int64_t value);

/* Energy::set_ttessel (lib/ttessel.cpp:2862): replicate the tessellation on the GPU. */
int lite_tess_upload(lite_ctx *ctx, const lite_tess_soa *tess);

/* Energy::add_features_* / add_theta_* (lib/ttessel.cpp:3016-3088): the d
 * features in flattened CatVector order vertices, edges, faces, segs
 * (lib/ttessel.cpp:2750-2761). */
int lite_energy_set(lite_ctx *ctx, int d, const int32_t *feature_id);

/* PseudoLikDiscrete::SetEnergy (lib/ttessel.cpp:3327-3348): all_merges +
 * all_flips, their statistic variations and the two sums. */
int lite_candidates_merges_flips(lite_ctx *ctx, int32_t *n_merges, int32_t *n_flips);

/* PseudoLikDiscrete::AddGivenSplit for a list (lib/ttessel.cpp:3371-3374);
 * split i = Split(halfedge he[i], x[i], angle[i]) (lib/ttessel.cpp:2090).
 * The list is this rank's shard; n_global_total is the size of the whole
 * batch over all ranks (== n when world_size == 1). */
int lite_candidates_add_splits_given(lite_ctx *ctx, int64_t n, const int32_t *he,
                                     const double *x, const double *angle,
                                     int64_t n_global_total);

/* PseudoLikDiscrete::AddSplits(n) (lib/ttessel.cpp:3356-3365) with the
 * reference's global RNG replaced by a counter-based Philox4x32-10 stream:
 * split j of the batch uses counter offset+j, whatever the number of ranks. */
int lite_candidates_add_splits_philox(lite_ctx *ctx, int64_t n_global, uint64_t seed,
                                      uint64_t offset);

/* Room for n_more_local further dummy splits on this rank, so that the AddSplits(n_nb) of every
 * PLInferenceNOIS iteration (lib/ttessel.cpp:3500) appends without reallocating. */
int lite_candidates_reserve(lite_ctx *ctx, int64_t n_more_local);
/* PseudoLikDiscrete::ClearSplits (lib/ttessel.cpp:3378-3380). */
int lite_candidates_clear_splits(lite_ctx *ctx);
int lite_candidates_count(lite_ctx *ctx, int64_t *n_splits_local, int64_t *n_splits_global);

/* PseudoLikDiscrete::GetValue / GetGradient / GetHessian in one pass
 * (lib/ttessel.cpp:3387-3443).  grad: d doubles, hess: d*d row-major; any of
 * the three outputs may be NULL.  With world_size > 1 the shards' 1+d+d(d+1)/2
 * sums (S0, S1, upper triangle of S2) are added across the GPUs inside the reduce
 * kernel, over NVLink peer memory (lite_comm_info; ncclAllReduce when the peers
 * cannot be mapped or LITE_B200_ALLREDUCE=nccl); every rank gets the bitwise
 * identical result.  Collective: every rank must make the call. */
int lite_pl_eval(lite_ctx *ctx, const double *theta, double *lpl, double *grad, double *hess);
/* The host arithmetic at the end of lite_pl_eval (lib/ttessel.cpp:3387-3443) on sums
 * supplied by the caller: S, F = [S0, S1_k, S2_kl (k<=l)] over the dummy splits of ALL
 * ranks / over the flips, M, Fl = sum_stat_merges, sum_stat_flips.  Needs no device. */
int lite_pl_combine(int d, int64_t n_splits_global, double int_length, const double *theta,
                    const double *S, const double *F, const double *M, const double *Fl,
                    double *lpl, double *grad, double *hess);
/* Shard [first, first+count) of a batch of n held by `rank` of `world`. */
int lite_shard_range(int64_t n, int rank, int world, int64_t *first, int64_t *count);
/* Candidates behind the last lite_pl_eval whose ray found no exit edge (their statistic
 * row is all zero, which biases S0 by 1 each): counts since the last ClearSplits /
 * SetEnergy.  0 on every tessellation the tests know; see DESIGN.md. */
int lite_pl_failures(lite_ctx *ctx, int64_t *split_no_exit, int64_t *flip_no_exit);
/* mode: 0 single rank, 1 ncclAllReduce, 2 peer memory (NVLink) inside the reduce kernel;
 * exchange_ms: device time the last lite_pl_eval spent combining the ranks. */
int lite_comm_info(lite_ctx *ctx, int32_t *mode, float *exchange_ms);

/* Energy::statistic_variation sums kept by SetEnergy: sum_stat_merges,
 * sum_stat_flips (d doubles each, already negated as in the reference). */
int lite_pl_sums(lite_ctx *ctx, double *sum_stat_merges, double *sum_stat_flips);

/* Config "random_lines sampling and face clipping": sample n_global lines
 * (Philox, as add_splits_philox) and clip them against their faces without
 * computing statistics.  mode 0: checksum only, mode 1: store e1,e2,p1,p2
 * (readable with lite_debug_fetch).  checksum[0] = xor of (e1<<32|e2) hashes,
 * checksum[1] = sum of e2 ids. */
int lite_lines_sample_clip(lite_ctx *ctx, int64_t n_global, uint64_t seed, uint64_t offset,
                           int mode, uint64_t *checksum2);

/* Per-candidate outputs for parity tests.  `first`/`count` index this rank's
 * local candidates. */
enum lite_fetch {
  LITE_FETCH_SPLIT_E1 = 0,    /* int32  */
  LITE_FETCH_SPLIT_E2 = 1,    /* int32  (-1: no exit found) */
  LITE_FETCH_SPLIT_P1 = 2,    /* double[2] */
  LITE_FETCH_SPLIT_P2 = 3,    /* double[2] */
  LITE_FETCH_SPLIT_X = 4,     /* double */
  LITE_FETCH_SPLIT_ANGLE = 5, /* double */
  LITE_FETCH_SPLIT_STAT = 6,  /* double[d]  (-statistic_variation, as stored by the reference) */
  LITE_FETCH_MERGE_HE = 7,    /* int32 */
  LITE_FETCH_MERGE_STAT = 8,  /* double[d] */
  LITE_FETCH_FLIP_E1 = 9,     /* int32 */
  LITE_FETCH_FLIP_E2 = 10,    /* int32 */
  LITE_FETCH_FLIP_P2 = 11,    /* double[2] */
  LITE_FETCH_FLIP_RIGHT = 12, /* int32 */
  LITE_FETCH_FLIP_STAT = 13,  /* double[d] */
  LITE_FETCH_SPLIT_U = 14,    /* double: the uniform behind angle = acos(1-2U) (philox path) */
  LITE_FETCH_STATUS = 15      /* int64[4]: no-exit count, oversize-face count, ... */
};
int lite_debug_fetch(lite_ctx *ctx, int which, int64_t first, int64_t count, void *out);

/* Timing helpers for bench.py: device time (ms, CUDA events on the context's
 * stream) of the last add_splits / pl_eval / merges_flips kernels, and the
 * number of kernel launches issued by the context so far. */
int lite_last_kernel_ms(lite_ctx *ctx, float *add_splits_ms, float *pl_eval_ms,
                        float *merges_flips_ms);
int64_t lite_kernel_launch_count(lite_ctx *ctx);
/* The same times summed over the lite_pl_eval calls since the last reset: out5 = split kernel,
 * theta-reduce kernel, merges+flips kernel, cross-GPU exchange (ms), number of evaluations. */
int lite_timers_accumulated(lite_ctx *ctx, double out5[5], int reset);
/* Device time (ms) of the split-statistics theta-reduce kernel of the last lite_pl_eval. */
int lite_last_reduce_ms(lite_ctx *ctx, float *ms);
/* CUDA-event marks on the context's stream (the stream every kernel of the
 * context is launched on): mark slot a, do work, mark slot b, read b - a. */
#define LITE_TIMER_SLOTS 8
int lite_timer_mark(lite_ctx *ctx, int slot);
int lite_timer_elapsed(lite_ctx *ctx, int slot_a, int slot_b, float *ms);
/* Measured DFMA throughput (TFLOP/s) of this GPU: the FP64 roofline denominator. */
int lite_measure_fp64_peak(lite_ctx *ctx, double *tflops);

/* ------------------------------------------------------------------------
 * Ensemble of independent SMF chains (BASELINE.json config 5).  The reference
 * simulates ONE chain per process: SMFChain(&energy, p_split, p_merge).step(n)
 * (lib/ttessel.cpp:3121-3311) on a TTessel that starts as the empty window, and
 * applications/checkACS.cpp runs it once per seed, writing a row of statistics
 * every `nipl` proposals.  Here chain i of the ensemble is one GPU thread with
 * its own tessellation in device memory; its random numbers are the Philox4x32-10
 * stream (seed, global chain id, proposal index), so its trajectory does not
 * depend on the number of ranks.  Chains [r*n/G, (r+1)*n/G) live on rank r; no
 * collective is involved.
 * ------------------------------------------------------------------------ */
struct lite_host_tess; /* include/lite_b200_host.h */

/* TTessel::insert_window(Rectangle) (lib/ttessel.cpp:1544-1550) for every chain.
 * max_segments: capacity in internal segments per chain (<= 10896); a chain that
 * would exceed it stops and lite_smf_run returns LITE_ECAPACITY. */
int lite_smf_create(lite_ctx *ctx, int64_t n_chains_global, int32_t max_segments, double x0,
                    double y0, double x1, double y1, uint64_t seed);
int lite_smf_destroy(lite_ctx *ctx);
/* Energy::add_features_* / add_theta_* + SMFChain::SMFChain(e, p_split, p_merge)
 * (lib/ttessel.cpp:3016-3088, :3121-3127): d device feature ids with their theta. */
int lite_smf_set_model(lite_ctx *ctx, int d, const int32_t *feature_id, const double *theta,
                       double p_split, double p_merge);
#define LITE_SMF_COLS 12
/* n_samples x SMFChain::step(steps_per_sample) on every local chain.  rows_out (may
 * be NULL): n_samples x LITE_SMF_COLS x n_local doubles, the columns of
 * applications/checkACS.cpp:136-148 after each batch: l(T), n(T), n_bl, n_nbl,
 * m(T), proposed/accepted S, M, F of that batch, then the energy accumulated
 * since creation (Energy::add_value, lib/ttessel.cpp:3284). */
int lite_smf_run(lite_ctx *ctx, int32_t n_samples, int64_t steps_per_sample, double *rows_out);
/* The same columns for the current state (counts: since the last batch started);
 * validate != 0 also runs the consistency check of TTessel::is_valid
 * (lib/ttessel.cpp:1721-1753) on every chain. */
int lite_smf_stats(lite_ctx *ctx, double *rows_out, int32_t validate);
/* Parity hook: advance ONE local chain by n_steps proposals, recording each one
 * (layout: smf::TraceRec, 23 doubles; tests/smf_replay.py). */
int lite_smf_trace(lite_ctx *ctx, int64_t chain_local, int64_t n_steps, void *trace_out,
                   int64_t *n_done);
/* Copy one chain's tessellation into a host model (lite_host_export gives the
 * lite_tess_soa to upload for a pseudo-likelihood fit, lite_host_write_text the
 * reference's text format). */
int lite_smf_export(lite_ctx *ctx, int64_t chain_local, struct lite_host_tess **out);
int lite_smf_count(lite_ctx *ctx, int64_t *n_local, int64_t *n_global, int64_t *first_global,
                   int64_t *bytes_per_chain);
/* The intersection output of applications/checkACS.cpp:151-172 for the current state of every
 * local chain: one record (local chain index, probe index, x, y) for each internal edge (both
 * faces bounded) that meets probe segment k = (probes[4k..4k+1]) -> (probes[4k+2..4k+3]) in a single
 * point (crossing or touching; collinear overlaps are not points).  checkACS probes with the four
 * window edges and the two mid-lines.  Records come in no particular order.  n_found receives the
 * number of intersections; if it exceeds cap the first cap are returned with LITE_ECAPACITY
 * (call with cap = 0 and NULL outputs to count). */
int lite_smf_intersections(lite_ctx *ctx, int32_t n_probes, const double *probes, int64_t cap,
                           int32_t *chain_out, int32_t *probe_out, double *xy_out, int64_t *n_found);
/* Device time (ms) of the last lite_smf_run kernel. */
int lite_smf_last_ms(lite_ctx *ctx, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* LITE_B200_H */
