/* lite_b200_host.h — C entry points of the CGAL-free host model (the state
 * behind the TTessel facade of include/ttessel.h) for callers that cannot use
 * C++: build a T-tessellation in a polygonal window, run the CRTT chain that
 * synthesises benchmark inputs, apply the three moves, flatten it for
 * lite_tess_upload().  All of this is host-side, sequential bookkeeping (the
 * reference's LineTes/TTessel layer, lib/ttessel.cpp:32-251, :606-780,
 * :1506-1710); the pseudo-likelihood path itself only runs on the GPU. */
#ifndef LITE_B200_HOST_H
#define LITE_B200_HOST_H
#include "lite_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct lite_host_tess lite_host_tess;

/* TTessel::insert_window(Rectangle), lib/ttessel.cpp:1544-1550 */
int lite_host_create_rect(lite_host_tess **out, double x0, double y0, double x1, double y1);
/* LineTes::read, lib/ttessel.cpp:1377-1441: the window (one or several polygons, outer boundaries
 * counter-clockwise, holes clockwise) then one maximal segment per line.  A window with holes or
 * several polygons can be checked, exported (lite_host_export -> lite_tess_upload) and written;
 * the moves and the chain below need a single hole-free polygon. */
int lite_host_create_from_text(lite_host_tess **out, const char *text);
int lite_host_destroy(lite_host_tess *t);
/* SMFChain(&mod,p_split,p_merge).step(n) under the CRTT energy
 * log(tau)*minus_is_segment_internal (demos/pseudo_lik_inference.cpp:56-69).
 * counts: proposed/accepted S, M, F (ModifCounts, include/ttessel.h:1027-1034). */
int lite_host_crtt_run(lite_host_tess *t, double tau, uint64_t n_steps, uint64_t seed,
                       int64_t counts6[6]);
/* TTessel::update(Split(e,x,angle)) / update(Merge(e)) / update(Flip(e)) on
 * halfedge ids of the last export (lib/ttessel.cpp:1580-1701). */
int lite_host_update_split(lite_host_tess *t, int32_t he, double x, double angle);
int lite_host_update_merge(lite_host_tess *t, int32_t nb_index);
int lite_host_update_flip(lite_host_tess *t, int32_t b_index, int side);
/* number of cells, non-blocking and blocking internal segments */
int lite_host_counts(lite_host_tess *t, int32_t *n_cells, int32_t *n_nb, int32_t *n_b);
/* consistency check in the spirit of TTessel::is_valid (lib/ttessel.cpp:1721-1753);
 * returns 0 when valid, LITE_EINVAL with a message otherwise */
int lite_host_check(lite_host_tess *t);
/* Flatten; the pointers stay valid until the next call on t. */
int lite_host_export(lite_host_tess *t, lite_tess_soa *out);
/* LineTes::write (lib/ttessel.cpp:1348-1370), decimal doubles as n/1; returns the
 * needed size; writes at most cap bytes */
int64_t lite_host_write_text(lite_host_tess *t, char *buf, int64_t cap);

/* clip_segment_by_polygon(Segment, HPolygons), lib/ttessel.cpp:3762-3786 (and :3696-3760 for one
 * polygon with holes): the longest piece of seg = (x0, y0, x1, y1) inside the open window of t,
 * written to out in the direction of seg.  LITE_EINVAL when the segment does not hit the window
 * (the reference throws std::domain_error). */
int lite_host_clip_segment(lite_host_tess *t, const double seg[4], double out[4]);
/* LineTes::insert_segment, lib/ttessel.cpp:904-980: the segment is clipped by the window and
 * added to the maximal segments of t.  While the segments do not form a T-tessellation (free
 * ends, crossings, corners: what remove_[ilx]vertices would clean, not built) the arrangement
 * can be inspected (lite_host_edges), written and inserted into; checks, moves and the export
 * answer LITE_EINVAL / LITE_ESTATE until they do.  Ids of an earlier export are invalidated.
 * seg_id (may be null) receives the id of the new segment. */
int lite_host_insert_segment(lite_host_tess *t, const double seg[4], int32_t *seg_id);
/* LineTes::is_a_T_tessellation, lib/ttessel.cpp:1307-1329: 1 or 0 */
int lite_host_is_t_tessellation(lite_host_tess *t);
/* Every halfedge of the arrangement, T-tessellation or not: end points (source x, y, target x,
 * y), segment id, next / previous halfedge along its segment (-1 at the segment's ends:
 * LineTes_halfedge::get_next_hf / get_prev_hf).  Returns the number of halfedges; writes only if
 * cap holds them all; any array may be null. */
int64_t lite_host_edges(lite_host_tess *t, int64_t cap, double *xy, int32_t *seg, int32_t *next_hf,
                        int32_t *prev_hf);

/* LineTes::remove_ivertices / remove_lvertices / remove_xvertices, lib/ttessel.cpp:982-1296: the
 * passes that turn what a series of lite_host_insert_segment calls left into a T-tessellation.
 * Free ends (degree 1) are cut back to the previous vertex or prolonged to the next edge,
 * whichever changes the smaller length, smallest change first; corners (degree 2, inside the
 * window) are prolonged likewise; crossings inside the window with no free end next to them are
 * broken into two T-vertices at most `step` apart.  imax: 0 = all (otherwise imax + 1 removals,
 * as the reference's loop counts).  The lengths (may be null) are what the verbose mode of the
 * reference prints.  Ids of an earlier export are invalidated. */
int lite_host_remove_ivertices(lite_host_tess *t, int64_t imax, double *removed_length, double *added_length);
int lite_host_remove_lvertices(lite_host_tess *t, int64_t imax, double *added_length);
int lite_host_remove_xvertices(lite_host_tess *t, double step);
/* free ends, corners inside the window, crossings (LineTes::number_of_xvertices(false, false),
 * lib/ttessel.cpp:272-278) and the crossings remove_xvertices would break
 * (number_of_xvertices(true, true)); any pointer may be null */
int lite_host_vertex_census(lite_host_tess *t, int32_t *n_free_ends, int32_t *n_corners, int32_t *n_crossings,
                            int32_t *n_crossings_to_remove);

/* The vertices that keep t from being a T-tessellation (is_a_T_vertex per vertex,
 * lib/ttessel.cpp:1307-1329, :4002-4046): the census above plus what no pass removes (three ends
 * on one point; two ends on the same point of a third segment).  first_xy (may be null): where
 * the first one is. */
int lite_host_non_t_vertices(lite_host_tess *t, int32_t *n, double first_xy[2]);

#ifdef __cplusplus
}
#endif
#endif
