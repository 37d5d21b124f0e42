/*
 * clownresampler.h -- drop-in declaration header for the B200 (sm_100a) implementation of
 * clownresampler's Lanczos FIR path.
 *
 * This header declares exactly the API surface of the reference's single header
 * (/root/reference/clownresampler.h, cited below as H:line) so that code written against the
 * reference -- including the reference's own tests/test-low-level.c and tests/test-high-level.c,
 * unmodified -- compiles against it and links with libclownresampler_b200.so.  There is no
 * implementation in this file: every function lives in the shared library, whose hot loop
 * (H:986-1035 inside H:1058-1092) runs as hand-written CUDA kernels.  There is no CPU fallback:
 * if no CUDA device is usable the resample calls report an error through
 * ClownResamplerB200_GetLastError() (clownresampler_b200.h) and emit no frames.
 *
 * ABI: the library is built for the reference's DEFAULT integer mode (H:546-560), i.e. the
 * C89 types below, on LP64: sizeof(ClownResampler_Precomputed) = 49152,
 * ..._LowestLevel_Configuration = 32, ..._LowLevel_State = 64, ..._HighLevel_State = 8296
 * (SURVEY.md section 8).  CC_USE_C99_INTEGERS (H:483) changes those layouts and is rejected.
 *
 * The reference's usage macros are accepted and ignored: CLOWNRESAMPLER_IMPLEMENTATION
 * (H:859) has nothing to instantiate, and CLOWNRESAMPLER_STATIC (H:436) cannot make functions
 * of a shared library static, so both modes resolve to the same external symbols.
 */
#ifndef CLOWNRESAMPLER_B200_DROPIN_H
#define CLOWNRESAMPLER_B200_DROPIN_H

#include <stddef.h>

#ifdef CC_USE_C99_INTEGERS
#error "libclownresampler_b200 is built for the reference's default (C89) integer mode; do not define CC_USE_C99_INTEGERS"
#endif

/* Compile-time configuration, fixed at the reference's defaults (H:445-460). */
#ifndef CLOWNRESAMPLER_KERNEL_RADIUS
#define CLOWNRESAMPLER_KERNEL_RADIUS 3
#endif
#ifndef CLOWNRESAMPLER_KERNEL_RESOLUTION
#define CLOWNRESAMPLER_KERNEL_RESOLUTION 0x400
#endif
#ifndef CLOWNRESAMPLER_MAXIMUM_CHANNELS
#define CLOWNRESAMPLER_MAXIMUM_CHANNELS 16
#endif
#if CLOWNRESAMPLER_KERNEL_RADIUS != 3 || CLOWNRESAMPLER_KERNEL_RESOLUTION != 0x400 || CLOWNRESAMPLER_MAXIMUM_CHANNELS != 16
#error "libclownresampler_b200 is built for KERNEL_RADIUS=3, KERNEL_RESOLUTION=0x400, MAXIMUM_CHANNELS=16"
#endif

/* Integer types of the default mode (H:546-560) and the boolean (H:606-611). */
#ifndef CC_INTEGERS_DEFINED
#define CC_INTEGERS_DEFINED
typedef signed char    cc_s8l;
typedef signed short   cc_s16l;
typedef signed long    cc_s32l;
typedef unsigned char  cc_u8l;
typedef unsigned short cc_u16l;
typedef unsigned long  cc_u32l;
typedef signed int     cc_s8f;
typedef signed int     cc_s16f;
typedef signed long    cc_s32f;
typedef unsigned int   cc_u8f;
typedef unsigned int   cc_u16f;
typedef unsigned long  cc_u32f;
typedef cc_u8l cc_bool;
enum { cc_false = 0, cc_true = 1 };
#endif

/* The precomputed Lanczos table (H:627-630): 6144 entries of 65536 * L(x), x in [-3, 3). */
typedef struct ClownResampler_Precomputed
{
	cc_s32l lanczos_kernel_table[CLOWNRESAMPLER_KERNEL_RADIUS * 2 * CLOWNRESAMPLER_KERNEL_RESOLUTION];
} ClownResampler_Precomputed;

/* Tap geometry derived from the three rates (H:632-638).  Callers read
   integer_stretched_kernel_radius to size their padding (H:728-729). */
typedef struct ClownResampler_LowestLevel_Configuration
{
	size_t stretched_kernel_radius;         /* 16.16 */
	size_t integer_stretched_kernel_radius;
	size_t stretched_kernel_radius_delta;   /* 16.16 */
	size_t kernel_step_size;
} ClownResampler_LowestLevel_Configuration;

/* Caller-owned, plain-data stream state (H:640-648). */
typedef struct ClownResampler_LowLevel_State
{
	ClownResampler_LowestLevel_Configuration lowest_level;
	cc_u8f channels;
	size_t position_integer;
	cc_u32f position_fractional;            /* 16.16 */
	cc_u32f increment;                      /* 16.16 */
} ClownResampler_LowLevel_State;

/* Streaming wrapper state with its 4096-sample window buffer (H:650-659). */
typedef struct ClownResampler_HighLevel_State
{
	ClownResampler_LowLevel_State low_level;
	cc_s16l input_buffer[0x1000];
	cc_s16l *input_buffer_start;
	cc_s16l *input_buffer_end;
	size_t maximum_integer_stretched_kernel_radius;
	size_t leading_padding_frames_needed, trailing_padding_frames_remaining;
} ClownResampler_HighLevel_State;

/* H:661-662 */
typedef size_t (*ClownResampler_InputCallback)(void *user_data, cc_s16l *buffer, size_t total_frames);
typedef cc_bool (*ClownResampler_OutputCallback)(void *user_data, const cc_s32f *frame, cc_u8f total_samples);

#ifdef __cplusplus
extern "C" {
#endif

/* Fills the table on the host in double precision with libm, bit-identical to H:955-961.
   The device copy is made lazily, keyed by table contents, at the first resample call. */
void ClownResampler_Precompute(ClownResampler_Precomputed *precomputed);                                   /* replaces H:682 */

cc_bool ClownResampler_LowestLevel_Configure(ClownResampler_LowestLevel_Configuration *configuration,
	cc_u32f input_sample_rate, cc_u32f output_sample_rate, cc_u32f low_pass_filter_sample_rate);       /* replaces H:687 */

/* One output frame (H:688).  Computed on the GPU like everything else; meant for spot checks,
   it costs a full host<->device round trip per call. */
void ClownResampler_LowestLevel_Resample(const ClownResampler_LowestLevel_Configuration *configuration,
	const ClownResampler_Precomputed *precomputed, cc_s32f *output_frame, cc_u8f channels,
	const cc_s16l *input_buffer, size_t position_integer, cc_u32f position_fractional);               /* replaces H:688 */

cc_bool ClownResampler_LowLevel_Init(ClownResampler_LowLevel_State *resampler, cc_u8f channels,
	cc_u32f input_sample_rate, cc_u32f output_sample_rate, cc_u32f low_pass_filter_sample_rate);       /* replaces H:711 */
cc_bool ClownResampler_LowLevel_Adjust(ClownResampler_LowLevel_State *resampler,
	cc_u32f input_sample_rate, cc_u32f output_sample_rate, cc_u32f low_pass_filter_sample_rate);       /* replaces H:719 */

/* Same contract as H:721-749: `input_buffer` points at the leading padding, R =
   lowest_level.integer_stretched_kernel_radius padding frames each side, not counted in
   *total_input_frames; one callback per output frame, in order, stop when it returns 0;
   *total_input_frames and the state are updated exactly as the reference's loop would.
   Frames are computed on the GPU in growing speculative chunks and delivered from pinned memory. */
cc_bool ClownResampler_LowLevel_Resample(ClownResampler_LowLevel_State *resampler,
	const ClownResampler_Precomputed *precomputed, const cc_s16l *input_buffer, size_t *total_input_frames,
	ClownResampler_OutputCallback output_callback, const void *user_data);                             /* replaces H:749 */

cc_bool ClownResampler_HighLevel_Init(ClownResampler_HighLevel_State *resampler, cc_u8f channels,
	cc_u32f input_sample_rate, cc_u32f output_sample_rate, cc_u32f low_pass_filter_sample_rate);       /* replaces H:770 */
cc_bool ClownResampler_HighLevel_Resample(ClownResampler_HighLevel_State *resampler,
	const ClownResampler_Precomputed *precomputed, ClownResampler_InputCallback input_callback,
	ClownResampler_OutputCallback output_callback, const void *user_data);                             /* replaces H:825 */
cc_bool ClownResampler_HighLevel_Adjust(ClownResampler_HighLevel_State *resampler,
	cc_u32f input_sample_rate, cc_u32f output_sample_rate, cc_u32f low_pass_filter_sample_rate);       /* replaces H:839 */
cc_bool ClownResampler_HighLevel_ResampleEnd(ClownResampler_HighLevel_State *resampler,
	const ClownResampler_Precomputed *precomputed, ClownResampler_OutputCallback output_callback,
	const void *user_data);                                                                            /* replaces H:847 */

#ifdef __cplusplus
}
#endif

#endif /* CLOWNRESAMPLER_B200_DROPIN_H */
