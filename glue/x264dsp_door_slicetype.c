/* Wrapper translation unit: pulls the UNMODIFIED reference encoder/slicetype.c
 * in by #include so that its file-static lookahead driver can be called from the
 * test harness.  It is linked INSTEAD of the plain slicetype.o.
 * Test infrastructure only.
 */
#define x264_slicetype_decide xref_orig_slicetype_decide
#include "encoder/slicetype.c"
#undef x264_slicetype_decide

/* public doorway to the static x264_slicetype_frame_cost (encoder/slicetype.c:223) */
int xref_slicetype_frame_cost( x264_t *h, x264_frame_t **frames, int p0, int p1, int b )
{
    return x264_slicetype_frame_cost( h, frames, p0, p1, b );
}

/* Driver-level door (see x264dsp_doors.c): x264_slicetype_frame_cost caches its result in the frame
 * (slicetype.c:238), so a batched implementation only has to fill that cache before the reference
 * asks.  x264_slicetype_analyse is about to ask for next.list[0] against last_nonb (slicetype.c:408-429),
 * and x264_rc_analyse_slice asks for the same pair later (slicetype.c:605-642). */
typedef void (*xref_cost_cb)( void *h, void *p0, void *b, int want_intra, int16_t *mvs, int *costs, int *sums );
extern xref_cost_cb xref_hook_cost;
extern int xref_hook_calls[3];

void x264_slicetype_decide( x264_t *h )
{
    if( xref_hook_cost && h->lookahead->last_nonb && h->lookahead->next.i_size > 0 )
    {
        x264_frame_t *p0 = h->lookahead->last_nonb, *b = h->lookahead->next.list[0];
        if( b->i_type == X264_TYPE_AUTO && b->i_cost_est[1][0] < 0 )
        {
            const int n = h->mb.i_mb_count;
            int sums[8] = { 0 };
            const int want_intra = !b->b_intra_calculated;
            /* results land directly in the arrays the reference keeps them in */
            xref_hook_cost( h, p0, b, want_intra, &b->lowres_mvs[0][0][0][0], b->lowres_mv_costs[0][0], sums );
            xref_hook_calls[2]++;
            (void)n;
            if( want_intra )
            {
                b->i_cost_est[0][0] = sums[1];
                b->i_cost_est_aq[0][0] = sums[1];
            }
            b->i_cost_est[1][0] = sums[0];
            b->i_cost_est_aq[1][0] = sums[0];
            b->i_intra_mbs[1] = sums[2];
            b->b_intra_calculated = 1;
        }
    }
    xref_orig_slicetype_decide( h );
}
