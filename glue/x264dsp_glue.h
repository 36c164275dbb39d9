/* x264dsp_glue.h -- the reference-side glue of libx264dsp_b200.so (see x264dsp_glue.c). */
#ifndef X264DSP_GLUE_H
#define X264DSP_GLUE_H
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* puts the device behind every door of x264dsp_doors.c / x264dsp_door_slicetype.c; call once before x264_encoder_open
 * (the context is created on CUDA device 0 at the first door call, when the picture size is known) */
void x264dsp_glue_install( void );
/* on top of x264dsp_glue_install: the macroblock loop of every P slice (x264_macroblock_analyse + x264_macroblock_encode,
 * one reference frame, analyse.inter == 0) runs as ONE x264dsp_p_frames_dev call per frame; the host keeps the entropy
 * coder.  Other slices and settings fall back to the per-macroblock doors. */
void x264dsp_glue_install_pframe( void );
/* likewise the macroblock loop of every I slice (x264_mb_analyse_intra + the intra branches of x264_macroblock_encode) as ONE
 * x264dsp_i_frames_dev call per frame */
void x264dsp_glue_install_iframe( void );
/* every door forwards to the reference's own code again */
void x264dsp_glue_uninstall( void );
/* one JSON line: calls served per door, the doors' own {entered, eligible, served} counters, kernel launches */
int x264dsp_glue_report( FILE *out );

#ifdef __cplusplus
}
#endif
#endif
