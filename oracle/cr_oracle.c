/*
 * cr_oracle.c -- CPU restatement of clownresampler's Lanczos FIR hot path.
 * TEST INFRASTRUCTURE ONLY (see cr_oracle.h).  Parity pinned by tests/test_oracle.py.
 *
 * H = /root/reference/clownresampler.h.  All hot-path arithmetic is integer; the
 * reference runs it in `long` (64-bit on LP64), this file in int64_t/uint64_t, so
 * there is no overflow difference.  C division truncates toward zero, which is
 * exactly what H:625 relies on, so `/` is used the same way here.
 */
#include "cr_oracle.h"

#include <math.h>
#include <string.h>

#define FX_ONE 65536 /* 16.16, H:620 */

/* ------------------------------------------------------------------ table */

/* H:892-908: Lanczos-3 window, L(0) = 1. */
static double lanczos3(double x)
{
    const double pi = 3.1415926535897932384626433832795028841971693993751058209749445923078164062862089986280348253421170679;
    const double a = x * pi;
    const double b = a / (double)CRO_KERNEL_RADIUS;
    if (x == 0.0)
        return 1.0;
    return (sin(a) * sin(b)) / (a * b);
}

/* H:955-961: entry i samples x = (i/N*2 - 1) * radius; the double is scaled by 65536
 * and converted to integer with C's truncating cast. */
void cro_precompute(int32_t table[CRO_TABLE_SIZE])
{
    size_t i;
    for (i = 0; i < CRO_TABLE_SIZE; ++i) {
        const double x = ((double)i / (double)CRO_TABLE_SIZE * 2.0 - 1.0) * (double)CRO_KERNEL_RADIUS;
        table[i] = (int32_t)(lanczos3(x) * FX_ONE);
    }
}

/* ------------------------------------------------------------------ ratio */

/* H:913-953: 16.16 quotient floor(a * 65536 / b) by base-65536 long division on the
 * digits (a >> 16, a & 0xFFFF, 0).  Sentinel when a or b is 0 or the quotient needs
 * more than 32 bits; a zero quotient is bumped to 1. */
uint64_t cro_ratio(uint64_t a, uint64_t b)
{
    uint64_t d2, d1, d0, q2, q1, q0, rem, q;
    if (a == 0 || b == 0)
        return CRO_RATIO_INVALID;
    d2 = a / FX_ONE;
    d1 = a % FX_ONE;
    d0 = 0;
    q2 = d2 / b;               rem = d2 % b;
    d1 |= rem * FX_ONE;        /* H:929 uses |=, identical to + because d1 < 65536 */
    q1 = d1 / b;               rem = d1 % b;
    d0 |= rem * FX_ONE;
    q0 = d0 / b;
    if (q2 != 0 || q1 >= FX_ONE)
        return CRO_RATIO_INVALID;
    q = q1 * FX_ONE + q0;
    return q == 0 ? 1 : q;
}

/* ------------------------------------------------------------------ config */

static uint64_t min_u64(uint64_t a, uint64_t b) { return a < b ? a : b; }

/* H:963-984. */
int cro_configure(cro_config *cfg, uint64_t in_rate, uint64_t out_rate, uint64_t lpf_rate)
{
    const uint64_t low_pass = min_u64(in_rate, min_u64(out_rate, lpf_rate));   /* H:968 */
    const uint64_t scale = cro_ratio(in_rate, low_pass);                       /* H:969 */
    const uint64_t inv_scale = cro_ratio(low_pass, in_rate);                   /* H:970 */
    if (scale >= (uint64_t)0x1000 * FX_ONE)                                    /* H:974 */
        return 0;
    cfg->radius_fx = CRO_KERNEL_RADIUS * scale;                                /* H:977 */
    cfg->radius_int = (cfg->radius_fx + (FX_ONE - 1)) / FX_ONE;                /* H:978 */
    cfg->radius_delta = cfg->radius_int * FX_ONE - cfg->radius_fx;             /* H:979 */
    cfg->step = (uint64_t)((int64_t)CRO_KERNEL_RESOLUTION * (int64_t)inv_scale / FX_ONE); /* H:981 */
    return 1;
}

int cro_adjust(cro_state *st, uint64_t in_rate, uint64_t out_rate, uint64_t lpf_rate)
{
    st->increment = cro_ratio(in_rate, out_rate);                              /* H:1054 */
    return cro_configure(&st->cfg, in_rate, out_rate, lpf_rate);               /* H:1055 */
}

int cro_init(cro_state *st, uint32_t channels, uint64_t in_rate, uint64_t out_rate, uint64_t lpf_rate)
{
    st->channels = channels;                                                   /* H:1046-1048 */
    st->pos_int = 0;
    st->pos_frac = 0;
    return cro_adjust(st, in_rate, out_rate, lpf_rate);
}

/* ------------------------------------------------------------------ one frame */

/* H:986-1035.  `frame` holds `channels` accumulators that the caller zeroed (H:1071). */
void cro_frame(const cro_config *cfg, const int32_t *table, int64_t *frame, uint32_t channels,
               const int16_t *in, uint64_t pos_int, uint64_t pos_frac, int norm_mode, uint64_t legacy_scale)
{
    /* window of padded-buffer frames [first, last) -- H:993-996 */
    const uint64_t lo_rel = (pos_frac + cfg->radius_delta + (FX_ONE - 1)) / FX_ONE;
    const uint64_t hi_rel = (pos_frac + cfg->radius_fx) / FX_ONE;
    const uint64_t first = pos_int + lo_rel;
    const uint64_t last = pos_int + cfg->radius_int + hi_rel;
    /* table index of the first tap -- H:1001 (unsigned arithmetic, operands non-negative) */
    uint64_t kidx = cfg->step * (lo_rel * FX_ONE - pos_frac) / FX_ONE;
    int64_t tap_sum = 0;
    uint64_t f;
    uint32_t c;

    for (f = first; f < last; ++f, kidx += cfg->step) {                        /* H:1008 */
        const int64_t k = table[kidx];                                         /* H:1015 */
        const int16_t *s = in + f * channels;
        tap_sum += k;                                                          /* H:1016 */
        for (c = 0; c < channels; ++c)
            frame[c] += (int64_t)s[c] * k / FX_ONE;                            /* H:1020: per-tap truncation */
    }

    if (norm_mode == CRO_NORM_CURRENT) {
        const int64_t recip = (int64_t)0x80000000 / tap_sum;                   /* H:1025: 17.15 reciprocal */
        for (c = 0; c < channels; ++c)
            frame[c] = frame[c] * recip / (1 << 15);                           /* H:1033 */
    } else if (norm_mode == CRO_NORM_LEGACY) {
        /* Pre-normaliser revision that produced tests/test3. */
        for (c = 0; c < channels; ++c)
            frame[c] = frame[c] * (int64_t)legacy_scale / FX_ONE;
    }
}

/* ------------------------------------------------------------------ frame loop */

/* H:1058-1092. */
int cro_lowlevel_resample(cro_state *st, const int32_t *table, const int16_t *in,
                          uint64_t *total_input_frames, int32_t *out, uint64_t max_frames,
                          uint64_t *frames_written, int norm_mode, uint64_t legacy_scale)
{
    uint64_t n = 0;
    for (;;) {
        int64_t acc[CRO_MAX_CHANNELS];
        uint32_t c;

        if (st->pos_int >= *total_input_frames) {                              /* H:1063-1067 */
            st->pos_int -= *total_input_frames;
            *total_input_frames = 0;
            if (frames_written) *frames_written = n;
            return 1;
        }
        memset(acc, 0, sizeof acc);                                            /* H:1071 */
        cro_frame(&st->cfg, table, acc, st->channels, in, st->pos_int, st->pos_frac, norm_mode, legacy_scale);

        st->pos_frac += st->increment;                                         /* H:1076-1078 */
        st->pos_int += st->pos_frac / FX_ONE;
        st->pos_frac %= FX_ONE;

        if (out)
            for (c = 0; c < st->channels; ++c)
                out[n * st->channels + c] = (int32_t)acc[c];
        ++n;

        if (max_frames != 0 && n == max_frames) {                              /* callback returned 0: H:1081-1088 */
            const uint64_t consumed = min_u64(st->pos_int, *total_input_frames);
            *total_input_frames -= consumed;
            st->pos_int -= consumed;
            if (frames_written) *frames_written = n;
            return 0;
        }
    }
}

/* number of frames the loop above emits when never stopped. */
uint64_t cro_count_output_frames(uint64_t pos_int, uint64_t pos_frac, uint64_t increment, uint64_t total_input_frames)
{
    const unsigned __int128 start = ((unsigned __int128)pos_int << 16) + pos_frac;
    const unsigned __int128 end = (unsigned __int128)total_input_frames << 16;
    if (start >= end)
        return 0;
    return (uint64_t)((end - start + increment - 1) / increment);
}

/* ------------------------------------------------------------------ streaming wrapper */

/* H:1101-1176 + H:1216-1250 restated for an in-memory stream.  The reference keeps a
 * 4096-sample buffer laid out as [R carried frames][R look-ahead frames][new frames...],
 * where R is the integer kernel radius at init time: the stream is delayed by R frames,
 * each refill first moves the last 2R frames to the front (H:1150), and the low-level
 * loop is run on the frames between the two R-frame dead zones (H:1165-1171).  The end
 * flush appends R zero frames (H:1223-1233). */
#define CRO_HL_BUFFER_SAMPLES 0x1000 /* H:654 */

typedef struct hl_source {
    const int16_t *data;
    uint64_t frames_left;
    uint64_t zero_frames_left;
    uint64_t chunk_limit;
    uint32_t channels;
} hl_source;

static uint64_t hl_pull(hl_source *src, int16_t *dst, uint64_t want)
{
    uint64_t n;
    if (src->chunk_limit != 0 && want > src->chunk_limit)
        want = src->chunk_limit;
    if (src->frames_left != 0) {
        n = min_u64(want, src->frames_left);
        memcpy(dst, src->data, n * src->channels * sizeof(int16_t));
        src->data += n * src->channels;
        src->frames_left -= n;
        return n;
    }
    n = min_u64(want, src->zero_frames_left);
    memset(dst, 0, n * src->channels * sizeof(int16_t));
    src->zero_frames_left -= n;
    return n;
}

uint64_t cro_highlevel_stream(uint32_t channels, uint64_t in_rate, uint64_t out_rate, uint64_t lpf_rate,
                              const int32_t *table, const int16_t *input, uint64_t n_input_frames,
                              uint64_t max_chunk_frames, int32_t *out, uint64_t out_capacity_frames)
{
    cro_state st;
    int16_t buf[CRO_HL_BUFFER_SAMPLES];
    hl_source src;
    uint64_t R, lead_needed, produced = 0;
    uint64_t start, end; /* sample offsets into buf, H:655-656 */

    if (channels == 0 || channels > CRO_MAX_CHANNELS)                          /* H:1103 */
        return (uint64_t)-1;
    if (!cro_init(&st, channels, in_rate, out_rate, lpf_rate))                 /* H:1106 */
        return (uint64_t)-1;
    R = st.cfg.radius_int;                                                     /* H:1109 */
    if (2 * R * channels >= CRO_HL_BUFFER_SAMPLES)
        return (uint64_t)-1;                                                   /* the reference would overrun its buffer here */
    memset(buf, 0, R * channels * sizeof(int16_t));                            /* H:1112 */
    start = end = R * channels;                                                /* H:1115 */
    lead_needed = R;

    src.data = input;
    src.frames_left = n_input_frames;
    src.zero_frames_left = R;                                                  /* trailing flush, H:1226 */
    src.chunk_limit = max_chunk_frames;
    src.channels = channels;

    /* The reference is driven as Resample(real input) then ResampleEnd(zero input); since the
     * state carries over unchanged between the two calls, one loop over a source that turns
     * to zeros when the real input is exhausted is equivalent.  The only observable
     * difference is that a 0-frame read ends a call (H:1132, H:1157), after which the caller
     * re-enters with the padding source; hl_pull never returns 0 until both are exhausted. */
    for (;;) {
        while (lead_needed != 0) {                                             /* H:1127-1136 */
            const uint64_t got = hl_pull(&src, buf + (2 * R - lead_needed) * channels, lead_needed);
            if (got == 0)
                return produced;
            lead_needed -= got;
        }
        if (start == end) {                                                    /* H:1141-1158 */
            uint64_t got;
            memmove(buf, buf + end - R * channels, 2 * R * channels * sizeof(int16_t));
            start = R * channels;
            got = hl_pull(&src, buf + 2 * R * channels, (CRO_HL_BUFFER_SAMPLES - 2 * R * channels) / channels);
            end = start + got * channels;
            if (got == 0)
                return produced;
        }
        {
            uint64_t frames = (end - start) / channels;                        /* H:1167 */
            uint64_t wrote = 0;
            const uint64_t room = out_capacity_frames - produced;
            if (room == 0)
                return produced;
            cro_lowlevel_resample(&st, table, buf + start - st.cfg.radius_int * channels, &frames,
                                  out ? out + produced * channels : NULL, room, &wrote, CRO_NORM_CURRENT, 0);
            produced += wrote;
            start = end - frames * channels;                                   /* H:1171 */
        }
    }
}

/* ------------------------------------------------------------------ synthetic input */

/* Counter-based generator (lowbias32-style integer hash of (seed, stream, channel, frame)),
 * top 16 bits as the sample.  The device-side generator in clownresampler_b200/csrc uses the
 * same constants; tests compare the two bit-for-bit. */
static uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7FEB352Du;
    x ^= x >> 15; x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}

int16_t cro_noise_sample(uint32_t seed, uint32_t stream, uint64_t frame, uint32_t channel)
{
    uint32_t h = seed ^ 0x9E3779B9u;
    h = mix32(h + stream * 0x85EBCA6Bu);
    h = mix32(h + channel * 0xC2B2AE35u);
    h = mix32(h + (uint32_t)frame);
    h = mix32(h + (uint32_t)(frame >> 32) * 0x27D4EB2Fu);
    return (int16_t)(h >> 16);
}

void cro_fill_noise(int16_t *dst, uint32_t seed, uint32_t stream, uint64_t first_frame, uint64_t n_frames, uint32_t channels)
{
    uint64_t f;
    uint32_t c;
    for (f = 0; f < n_frames; ++f)
        for (c = 0; c < channels; ++c)
            dst[f * channels + c] = cro_noise_sample(seed, stream, first_frame + f, c);
}
