/* Force-included (-include) when compiling the UNMODIFIED reference sources
 * under /root/reference for x86.  Three TI C6000 intrinsics leak into the
 * portable branch of encoder/analyse.c (x264_memset_uint16, analyse.c:217-223);
 * these are plain-C stand-ins with the documented TI semantics.  Test
 * infrastructure only -- nothing here ships in the product library.
 */
#ifndef X264DSP_TI_SHIM_H
#define X264DSP_TI_SHIM_H
#include <stdint.h>

/* _pack2(a,b): low halfword of a -> bits 31..16, low halfword of b -> bits 15..0 */
#define _pack2( a, b ) ( (((uint32_t)(a) & 0xffffu) << 16) | ((uint32_t)(b) & 0xffffu) )
/* _itoll(hi,lo): build a 64-bit value from two 32-bit halves */
#define _itoll( hi, lo ) ( (int64_t)( ((uint64_t)(uint32_t)(hi) << 32) | (uint32_t)(lo) ) )
/* _mem8(p): unaligned 8-byte lvalue */
typedef int64_t __attribute__((may_alias, aligned(1))) x264dsp_mem8_t;
#define _mem8( p ) ( *(x264dsp_mem8_t *)(p) )

#endif
