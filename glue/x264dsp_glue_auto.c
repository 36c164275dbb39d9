/* Linked into x264ref_gpu only: installs the glue before main() runs, so that the reference's CLI (x264.c, input.c,
 * output.c -- compiled unmodified) needs no change at all.  X264DSP_GLUE_PFRAME=0 keeps the P-slice macroblock loop on the
 * host (per-macroblock doors only); X264DSP_GLUE=0 leaves the doors closed (the binary then
 * behaves exactly like the reference CLI); X264DSP_GLUE_STATS=<file> writes x264dsp_glue_report() there at exit. */
#include <execinfo.h>
#include <signal.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "x264dsp_glue.h"

static void glue_atexit( void )
{
    const char *path = getenv( "X264DSP_GLUE_STATS" );
    if( path )
    {
        FILE *f = fopen( path, "w" );
        if( f )
        {
            x264dsp_glue_report( f );
            fclose( f );
        }
    }
}

/* X264DSP_GLUE_DEBUG=1: a fault prints its call chain before the process dies */
static void glue_fault( int sig )
{
    void *bt[48];
    const int n = backtrace( bt, 48 );
    backtrace_symbols_fd( bt, n, 2 );
    _exit( 128 + sig );
}

__attribute__((constructor)) static void glue_auto( void )
{
    if( getenv( "X264DSP_GLUE_DEBUG" ) )
    {
        signal( SIGSEGV, glue_fault );
        signal( SIGBUS, glue_fault );
    }
    const char *e = getenv( "X264DSP_GLUE" );
    if( e && !strcmp( e, "0" ) )
        return;
    x264dsp_glue_install();
    e = getenv( "X264DSP_GLUE_PFRAME" );           /* 0: per-macroblock doors only */
    if( !e || strcmp( e, "0" ) )
    {
        x264dsp_glue_install_pframe();
        x264dsp_glue_install_iframe();
    }
    atexit( glue_atexit );
}
