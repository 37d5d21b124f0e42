/*
 * cr_oracle.h -- CPU restatement of clownresampler's Lanczos FIR hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity checker for the CUDA path in
 * clownresampler_b200/.  Nothing shipped may include, link or call it: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs do.  It is a from-scratch restatement (own structure, fixed-width types,
 * bulk output instead of per-frame callbacks) of the algorithm in the reference
 * header, cited below as H = /root/reference/clownresampler.h.
 *
 * Parity is PINNED: tests/test_oracle.py checks this file against
 *   (1) the unmodified reference compiled into oracle/_ref/libclownref.so
 *       (bit-exact on the reference's own test workload and on random cases),
 *   (2) the reference's golden file tests/test3 through the legacy normaliser,
 *   (3) committed vectors generated from the reference (tests/golden/).
 */
#ifndef CR_ORACLE_H
#define CR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRO_KERNEL_RADIUS      3      /* H:445-447 */
#define CRO_KERNEL_RESOLUTION  1024   /* H:452-454 */
#define CRO_MAX_CHANNELS       16     /* H:458-460 */
#define CRO_TABLE_SIZE         (CRO_KERNEL_RADIUS * 2 * CRO_KERNEL_RESOLUTION) /* H:629 */
#define CRO_RATIO_INVALID      0xFFFFFFFFu /* H:920, H:940 */

/* H:632-638.  Same four quantities, fixed-width. */
typedef struct cro_config {
    uint64_t radius_fx;      /* stretched_kernel_radius, 16.16 */
    uint64_t radius_int;     /* integer_stretched_kernel_radius */
    uint64_t radius_delta;   /* stretched_kernel_radius_delta, 16.16 */
    uint64_t step;           /* kernel_step_size */
} cro_config;

/* H:640-648. */
typedef struct cro_state {
    cro_config cfg;
    uint32_t channels;
    uint64_t pos_int;
    uint64_t pos_frac;   /* 16.16 fractional part, < 65536 between calls */
    uint64_t increment;  /* 16.16 */
} cro_state;

enum { CRO_NORM_CURRENT = 0, CRO_NORM_LEGACY = 1, CRO_NORM_NONE = 2 };

void     cro_precompute(int32_t table[CRO_TABLE_SIZE]);                       /* H:892-908, H:955-961 */
uint64_t cro_ratio(uint64_t a, uint64_t b);                                   /* H:913-953 */
int      cro_configure(cro_config *cfg, uint64_t in_rate, uint64_t out_rate, uint64_t lpf_rate); /* H:963-984 */
int      cro_init(cro_state *st, uint32_t channels, uint64_t in_rate, uint64_t out_rate, uint64_t lpf_rate); /* H:1044-1056 */
int      cro_adjust(cro_state *st, uint64_t in_rate, uint64_t out_rate, uint64_t lpf_rate);                  /* H:1052-1056 */

/* One output frame, H:986-1035.  `norm_mode` selects the current per-frame tap-sum
 * normaliser (H:1025,1033), the legacy fixed normaliser the shipped goldens were made
 * with (`legacy_scale` = 16.16 inverse kernel scale), or none. */
void cro_frame(const cro_config *cfg, const int32_t *table, int64_t *frame, uint32_t channels,
               const int16_t *padded_input, uint64_t pos_int, uint64_t pos_frac,
               int norm_mode, uint64_t legacy_scale);

/* The frame loop, H:1058-1092, with the output callback replaced by a bulk s32 buffer
 * plus a frame limit: the "callback" accepts frame k and returns 0 exactly when
 * k == max_frames (1-based), like examples/low-level.c:84.  max_frames == 0 means
 * "never stop".  Returns 1 when the input ran out (H:1067), 0 when the limit stopped it
 * (H:1088).  *frames_written receives the number of frames stored.  `out` may be NULL
 * (count / state only). */
int cro_lowlevel_resample(cro_state *st, const int32_t *table, const int16_t *padded_input,
                          uint64_t *total_input_frames, int32_t *out, uint64_t max_frames,
                          uint64_t *frames_written, int norm_mode, uint64_t legacy_scale);

/* Streaming wrapper, H:1101-1176 and H:1216-1250, restated over a whole in-memory
 * stream: feeds `input` (n_input_frames, interleaved, unpadded) through a 4096-sample
 * window buffer in at most `max_chunk_frames` frames per refill (0 = as many as fit),
 * then flushes radius_int zero frames.  Writes every output frame to `out`.
 * Returns the number of output frames (or (uint64_t)-1 on configuration failure). */
uint64_t cro_highlevel_stream(uint32_t channels, uint64_t in_rate, uint64_t out_rate, uint64_t lpf_rate,
                              const int32_t *table, const int16_t *input, uint64_t n_input_frames,
                              uint64_t max_chunk_frames, int32_t *out, uint64_t out_capacity_frames);

/* Closed forms used by the CUDA path; the tests check them against the loop. */
uint64_t cro_count_output_frames(uint64_t pos_int, uint64_t pos_frac, uint64_t increment, uint64_t total_input_frames);

/* Deterministic synthetic s16 generator shared with the device-side generator
 * (counter-based hash so any window can be regenerated): sample(stream, frame, channel, seed). */
int16_t  cro_noise_sample(uint32_t seed, uint32_t stream, uint64_t frame, uint32_t channel);
void     cro_fill_noise(int16_t *dst, uint32_t seed, uint32_t stream, uint64_t first_frame, uint64_t n_frames, uint32_t channels);

/* Clamp convention of the reference's playback callbacks (examples/low-level.c:74-77). */
static inline int16_t cro_clamp_s16(int64_t v) { return (int16_t)(v < -0x7FFF ? -0x7FFF : (v > 0x7FFF ? 0x7FFF : v)); }

#ifdef __cplusplus
}
#endif
#endif
