/*
 * pixie_host.h — C ABI of the CPU-side producers that sit ABOVE the raster hot path.
 *
 * In the reference these stay in Nim (treeform/pixie src/pixie/paths.nim: parsePath :119-262,
 * path builders :346-652, commandsToShapes :654-1057, strokeShapes :1922-2082,
 * shapesToSegments :1059-1090; src/pixie/internal.nim gaussianKernel :17-34).  There is no Nim
 * toolchain in this image, so the host side is mirrored in C++ (libpixie_host.so) to be able
 * to feed the C-ABI device library (pixie_cuda.h) with exactly what the Nim callers would
 * pass: y-quantised, oriented segments with winding, and the Gaussian LUT.
 *
 * Nothing here touches a GPU.  All functions return 0 on success, non-zero on error
 * (message via pixie_host_last_error(), the analogue of raising PixieError, common.nim:4).
 */
#ifndef PIXIE_HOST_H
#define PIXIE_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pixie_path pixie_path;         /* paths.nim:24-27 Path */
typedef struct pixie_segments pixie_segments; /* seq[(Segment, int16)] paths.nim:1059-1062 */

enum { PIXIE_BUTT_CAP = 0, PIXIE_ROUND_CAP = 1, PIXIE_SQUARE_CAP = 2 };   /* paths.nim:10-12 */
enum { PIXIE_MITER_JOIN = 0, PIXIE_ROUND_JOIN = 1, PIXIE_BEVEL_JOIN = 2 }; /* paths.nim:14-16 */

const char* pixie_host_last_error(void);

/* Path construction (paths.nim:51, :119, :339-652). */
pixie_path* pixie_host_path_new(void);
void pixie_host_path_free(pixie_path* p);
int pixie_host_path_parse(const char* svg_path, pixie_path** out);
int pixie_host_path_num_commands(const pixie_path* p);               /* length of the float command stream */
int pixie_host_path_commands(const pixie_path* p, float* out, int cap);
void pixie_host_path_move_to(pixie_path* p, float x, float y);
void pixie_host_path_line_to(pixie_path* p, float x, float y);
void pixie_host_path_bezier_curve_to(pixie_path* p, float x1, float y1, float x2, float y2, float x3, float y3);
void pixie_host_path_quadratic_curve_to(pixie_path* p, float x1, float y1, float x2, float y2);
void pixie_host_path_elliptical_arc_to(pixie_path* p, float rx, float ry, float rot, int large, int sweep, float x, float y);
int pixie_host_path_arc(pixie_path* p, float x, float y, float r, float a0, float a1, int ccw);
int pixie_host_path_arc_to(pixie_path* p, float x1, float y1, float x2, float y2, float r);
void pixie_host_path_rect(pixie_path* p, float x, float y, float w, float h, int clockwise);
void pixie_host_path_rounded_rect(pixie_path* p, float x, float y, float w, float h,
                                  float nw, float ne, float se, float sw, int clockwise);
void pixie_host_path_ellipse(pixie_path* p, float cx, float cy, float rx, float ry);
int pixie_host_path_polygon(pixie_path* p, float x, float y, float size, int sides);
void pixie_host_path_close(pixie_path* p);

/*
 * fillPath's producer chain for a solid paint (paths.nim:2108-2109 + :1604):
 * commandsToShapes(closeSubpaths=true, pixelScale(mat)) -> transform(mat) -> shapesToSegments.
 * mat is vmath Mat3 storage order: mat[0..2] = column 0 (m00,m01,m02), mat[3..5] = column 1,
 * mat[6..8] = column 2 (translation in mat[6], mat[7]).  NULL = identity.
 */
int pixie_host_fill_segments(const pixie_path* p, const float* mat, pixie_segments** out);

/* strokePath's producer chain (paths.nim:2163-2172 + :1604). */
int pixie_host_stroke_segments(const pixie_path* p, const float* mat, float stroke_width,
                               int line_cap, int line_join, float miter_limit,
                               const float* dashes, int num_dashes, pixie_segments** out);

int pixie_host_segments_count(const pixie_segments* s);
const float* pixie_host_segments_xyxy(const pixie_segments* s);     /* n x 4: at.x, at.y, to.x, to.y */
const int16_t* pixie_host_segments_winding(const pixie_segments* s); /* n */
void pixie_host_segments_free(pixie_segments* s);

/* internal.nim:17-34 — out must hold 2*radius+1 entries. */
int pixie_host_gaussian_kernel(int radius, uint16_t* out);

#ifdef __cplusplus
}
#endif
#endif
