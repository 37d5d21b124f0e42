/*
 * ref_shim.c -- thin bulk-buffer wrappers around the UNMODIFIED reference, compiled from
 * where it lies (-I$(REFERENCE_DIR)) into oracle/_ref/libclownref.so by oracle/Makefile.
 * TEST INFRASTRUCTURE ONLY: used to pin oracle/cr_oracle.c, to generate tests/golden/, and
 * as the "reference" CPU baseline of bench.py.  No reference source is copied: the header
 * is included by path and this file only adds callbacks that store frames in arrays.
 */
#define CLOWNRESAMPLER_IMPLEMENTATION
#include <clownresampler.h>

#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <time.h>

/* ---- ABI facts of the reference build (default C89 integer mode) ---- */
void ref_abi(uint64_t out[16])
{
    out[0] = sizeof(cc_s16l);  out[1] = sizeof(cc_s32l); out[2] = sizeof(cc_s32f); out[3] = sizeof(cc_u32f);
    out[4] = sizeof(cc_u8f);   out[5] = sizeof(cc_bool); out[6] = sizeof(size_t);
    out[7] = sizeof(ClownResampler_Precomputed);
    out[8] = sizeof(ClownResampler_LowestLevel_Configuration);
    out[9] = sizeof(ClownResampler_LowLevel_State);
    out[10] = sizeof(ClownResampler_HighLevel_State);
    out[11] = offsetof(ClownResampler_LowLevel_State, channels);
    out[12] = offsetof(ClownResampler_LowLevel_State, position_integer);
    out[13] = offsetof(ClownResampler_LowLevel_State, position_fractional);
    out[14] = offsetof(ClownResampler_LowLevel_State, increment);
    out[15] = offsetof(ClownResampler_HighLevel_State, input_buffer);
}

static ClownResampler_Precomputed g_pre;
static int g_pre_ready;
static const ClownResampler_Precomputed *pre(void)
{
    if (!g_pre_ready) { ClownResampler_Precompute(&g_pre); g_pre_ready = 1; }
    return &g_pre;
}

void ref_table(int32_t out[CLOWNRESAMPLER_KERNEL_RADIUS * 2 * CLOWNRESAMPLER_KERNEL_RESOLUTION])
{
    size_t i;
    const ClownResampler_Precomputed *p = pre();
    for (i = 0; i < CLOWNRESAMPLER_COUNT_OF(p->lanczos_kernel_table); ++i)
        out[i] = (int32_t)p->lanczos_kernel_table[i];
}

uint64_t ref_ratio(uint64_t a, uint64_t b) { return ClownResampler_CalculateRatio(a, b); }

int ref_configure(uint64_t out[4], uint64_t in_rate, uint64_t out_rate, uint64_t lpf)
{
    ClownResampler_LowestLevel_Configuration c;
    memset(&c, 0, sizeof c);
    if (!ClownResampler_LowestLevel_Configure(&c, in_rate, out_rate, lpf))
        return 0;
    out[0] = c.stretched_kernel_radius; out[1] = c.integer_stretched_kernel_radius;
    out[2] = c.stretched_kernel_radius_delta; out[3] = c.kernel_step_size;
    return 1;
}

/* ---- low-level bulk ---- */
typedef struct sink {
    int32_t *out;
    uint64_t written, limit;
} sink;

static cc_bool store_frame(void *user, const cc_s32f *frame, cc_u8f n)
{
    sink *s = (sink *)user;
    cc_u8f i;
    if (s->out)
        for (i = 0; i < n; ++i)
            s->out[s->written * n + i] = (int32_t)frame[i];
    ++s->written;
    return (cc_bool)(s->limit == 0 || s->written != s->limit);
}

/* state_io = {position_integer, position_fractional}; returns the reference's return value. */
int ref_lowlevel_bulk(uint32_t channels, uint64_t in_rate, uint64_t out_rate, uint64_t lpf,
                      const int16_t *padded_input, uint64_t *total_input_frames, uint64_t state_io[2],
                      int32_t *out, uint64_t max_frames, uint64_t *frames_written)
{
    ClownResampler_LowLevel_State st;
    sink s;
    size_t frames = (size_t)*total_input_frames;
    cc_bool r;
    if (!ClownResampler_LowLevel_Init(&st, channels, in_rate, out_rate, lpf))
        return -1;
    st.position_integer = (size_t)state_io[0];
    st.position_fractional = (cc_u32f)state_io[1];
    s.out = out; s.written = 0; s.limit = max_frames;
    r = ClownResampler_LowLevel_Resample(&st, pre(), padded_input, &frames, store_frame, &s);
    *total_input_frames = frames;
    state_io[0] = st.position_integer;
    state_io[1] = st.position_fractional;
    *frames_written = s.written;
    return r;
}

/* A pitch-bend style sequence on one buffer: segment k calls ClownResampler_LowLevel_Adjust(rates[3k..3k+2]) and
   then resamples until `limits[k]` frames were emitted (0 = until the input runs out), carrying the state and
   the remaining input over, exactly as a caller of H:719 + H:749 would.  Returns frames written in total. */
uint64_t ref_lowlevel_adjust_sequence(uint32_t channels, const uint64_t *rates, const uint64_t *limits, uint32_t segments,
                                      const int16_t *padded_input, uint64_t total_input_frames, int32_t *out, uint64_t state_out[3])
{
    ClownResampler_LowLevel_State st;
    sink s;
    size_t frames = (size_t)total_input_frames;
    const int16_t *in = padded_input;
    uint32_t k;
    if (!ClownResampler_LowLevel_Init(&st, channels, rates[0], rates[1], rates[2]))
        return (uint64_t)-1;
    s.out = out; s.written = 0;
    for (k = 0; k < segments && frames != 0; ++k) {
        const size_t before = frames;
        const uint64_t start = s.written;
        if (!ClownResampler_LowLevel_Adjust(&st, rates[3 * k], rates[3 * k + 1], rates[3 * k + 2]))
            return (uint64_t)-1;
        s.limit = limits[k] ? start + limits[k] : 0;
        ClownResampler_LowLevel_Resample(&st, pre(), in, &frames, store_frame, &s);
        in += (before - frames) * channels;     /* the unconsumed frames are the next call's input */
    }
    state_out[0] = st.position_integer; state_out[1] = st.position_fractional; state_out[2] = frames;
    return s.written;
}

/* ---- high-level streaming ---- */
typedef struct stream_ctx {
    sink s;
    const int16_t *data;
    uint64_t frames_left, chunk_limit;
    uint32_t channels;
} stream_ctx;

static size_t feed(void *user, cc_s16l *buffer, size_t total_frames)
{
    stream_ctx *c = (stream_ctx *)user;
    size_t n = total_frames;
    if (c->chunk_limit != 0 && n > c->chunk_limit) n = (size_t)c->chunk_limit;
    if (n > c->frames_left) n = (size_t)c->frames_left;
    memcpy(buffer, c->data, n * c->channels * sizeof(int16_t));
    c->data += n * c->channels;
    c->frames_left -= n;
    return n;
}

static cc_bool store_frame_hl(void *user, const cc_s32f *frame, cc_u8f n)
{
    return store_frame(&((stream_ctx *)user)->s, frame, n);
}

uint64_t ref_highlevel_stream(uint32_t channels, uint64_t in_rate, uint64_t out_rate, uint64_t lpf,
                              const int16_t *input, uint64_t n_input_frames, uint64_t max_chunk_frames,
                              int32_t *out, uint64_t out_capacity_frames)
{
    static ClownResampler_HighLevel_State st;
    stream_ctx c;
    if (!ClownResampler_HighLevel_Init(&st, channels, in_rate, out_rate, lpf))
        return (uint64_t)-1;
    c.s.out = out; c.s.written = 0; c.s.limit = out_capacity_frames;
    c.data = input; c.frames_left = n_input_frames; c.chunk_limit = max_chunk_frames; c.channels = channels;
    if (ClownResampler_HighLevel_Resample(&st, pre(), feed, store_frame_hl, &c))
        ClownResampler_HighLevel_ResampleEnd(&st, pre(), store_frame_hl, &c);
    return c.s.written;
}

/* Streaming with mid-stream ClownResampler_HighLevel_Adjust (pitch bend): segment k's rates apply from output frame
   switch_at[k-1] on (the output callback returns 0 there, the caller adjusts and calls Resample again, as a game mixer
   would between ticks).  switch_at has segments-1 strictly increasing entries.  Returns frames written, -1 on a refused
   Init, -2 on a refused Adjust. */
uint64_t ref_highlevel_adjust_stream(uint32_t channels, const uint64_t *rates, const uint64_t *switch_at, uint32_t segments,
                                     const int16_t *input, uint64_t n_input_frames, int32_t *out, uint64_t out_capacity_frames)
{
    static ClownResampler_HighLevel_State st;
    stream_ctx c;
    uint32_t k = 0;
    int ending = 0;
    if (!ClownResampler_HighLevel_Init(&st, channels, rates[0], rates[1], rates[2]))
        return (uint64_t)-1;
    c.s.out = out; c.s.written = 0;
    c.data = input; c.frames_left = n_input_frames; c.chunk_limit = 0; c.channels = channels;
    for (;;) {
        cc_bool ran_out;
        c.s.limit = k + 1 < segments ? switch_at[k] : out_capacity_frames;
        ran_out = ending ? ClownResampler_HighLevel_ResampleEnd(&st, pre(), store_frame_hl, &c)
                         : ClownResampler_HighLevel_Resample(&st, pre(), feed, store_frame_hl, &c);
        if (ran_out) {
            if (ending) break;
            ending = 1;
            continue;
        }
        if (c.s.written >= out_capacity_frames || k + 1 >= segments) break;
        ++k;
        if (!ClownResampler_HighLevel_Adjust(&st, rates[3 * k], rates[3 * k + 1], rates[3 * k + 2]))
            return (uint64_t)-2;
    }
    return c.s.written;
}

/* ---- single-thread timing of the reference's own loop ---- */
typedef struct clamp_sink { int16_t *out; uint64_t written; } clamp_sink;

static cc_bool clamp_store(void *user, const cc_s32f *frame, cc_u8f n)
{
    /* the clamp convention of examples/low-level.c:74-77, storing s16 */
    clamp_sink *s = (clamp_sink *)user;
    cc_u8f i;
    for (i = 0; i < n; ++i) {
        const cc_s32f v = frame[i];
        s->out[s->written * n + i] = (int16_t)(v < -0x7FFF ? -0x7FFF : (v > 0x7FFF ? 0x7FFF : v));
    }
    ++s->written;
    return cc_true;
}

/* Resamples one padded buffer with the LowLevel API, returns seconds spent inside
 * ClownResampler_LowLevel_Resample only; *frames_out = output frames produced. */
double ref_time_lowlevel(uint32_t channels, uint64_t in_rate, uint64_t out_rate, uint64_t lpf,
                         const int16_t *padded_input, uint64_t total_input_frames, int16_t *out_s16, uint64_t *frames_out)
{
    ClownResampler_LowLevel_State st;
    clamp_sink s;
    size_t frames = (size_t)total_input_frames;
    struct timespec t0, t1;
    if (!ClownResampler_LowLevel_Init(&st, channels, in_rate, out_rate, lpf))
        return -1.0;
    s.out = out_s16; s.written = 0;
    (void)pre();
    clock_gettime(CLOCK_MONOTONIC, &t0);
    ClownResampler_LowLevel_Resample(&st, pre(), padded_input, &frames, clamp_store, &s);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    *frames_out = s.written;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
