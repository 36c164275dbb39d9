/* Wrapper translation unit: pulls the UNMODIFIED reference encoder/slicetype.c
 * in by #include so that its file-static lookahead driver can be called from the
 * test harness.  It is linked INSTEAD of the plain slicetype.o.
 * Test infrastructure only.
 */
#include "encoder/slicetype.c"

/* public doorway to the static x264_slicetype_frame_cost (encoder/slicetype.c:223) */
int xref_slicetype_frame_cost( x264_t *h, x264_frame_t **frames, int p0, int p1, int b )
{
    return x264_slicetype_frame_cost( h, frames, p0, p1, b );
}
